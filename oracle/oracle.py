"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE, not product code).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; nothing under
``uclchem_b200/`` does.  It loads ``oracle/liboracle.so`` (built by
``oracle/Makefile`` from the plain-C restatement of the reference algorithm),
hands it a network as arrays, and exposes the same operations the reference
exposes through ``uclchemwrap``: one model, a grid of models, ``get_rates``
and ``get_odes``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from uclchem_b200.network import TYPE_NAMES, Network
from uclchem_b200.params import NPARAM

_HERE = Path(__file__).resolve().parent

NAMED = [
    "nh", "nh2", "nc", "ncx", "no", "nn", "nmg", "nmgx", "np", "nf", "nna", "nli", "npah", "nsx",
    "nsix", "nclx", "nd", "nhe", "n18o", "n15n", "n13c", "nelec", "nco", "nbulk", "nsurface", "ngn",
    "ngo", "ngoh", "nsi",
    "nR_H2Form_CT", "nR_H2Form_ER", "nR_H2Form_ERDes", "nR_HFreeze", "nR_EFreeze", "nR_H2Freeze",
    "nR_H2_hv", "nR_CO_hv", "nR_C_hv", "nR_H2_crp",
]

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)


class _OrcNetwork(C.Structure):
    _fields_ = [
        ("nspec", C.c_int32), ("nreac", C.c_int32), ("nice", C.c_int32), ("nsurf", C.c_int32),
        ("ngar", C.c_int32), ("n_loss", C.c_int32), ("n_gain", C.c_int32), ("n_refractory", C.c_int32),
        ("mass", _pd), ("atom_counts", _pi),
        ("surface_list", _pi), ("bulk_list", _pi), ("ice_list", _pi), ("gas_ice_list", _pi),
        ("binding_energy", _pd), ("formation_enthalpy", _pd),
        ("re", _pi), ("pr", _pi),
        ("alpha", _pd), ("beta", _pd), ("gama", _pd), ("min_temps", _pd), ("max_temps", _pd),
        ("reduced_masses", _pd),
        ("extrapolate", _pi), ("rtype", _pi), ("flux_factors", _pi),
        ("loss_species", _pi), ("loss_reaction", _pi), ("gain_species", _pi), ("gain_reaction", _pi),
        ("freeze_partners", _pi), ("gar_params", _pd),
        ("type_lo", _pi), ("type_hi", _pi), ("named", _pi), ("refractory_list", _pi), ("is_ion", _pi),
    ]


class OrcStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("nst", "nfe", "nje", "nlu", "nni", "ncfn", "netf", "nintervals")]


_SRCS = ("orc_vode.c", "orc_chem.c", "orc_model.c", "orc_cshock.c", "orc_collapse.c", "orc_vode.h", "orc_internal.h", "uclchem_oracle.h")


def build(force: bool = False) -> Path:
    so = _HERE / "liboracle.so"
    srcs = [_HERE / f for f in _SRCS]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(_HERE), "-B", "liboracle.so"], check=True, capture_output=True)
    return so


def build_native() -> tuple[Path, str]:
    """A copy compiled for the host it runs on (`-O3 -march=native`), for the CPU timing legs of bench.py: the
    portable library is built for x86-64-v3 so that it travels.  Falls back to the portable one if the
    compiler is missing.  Returns (path, compile flags)."""
    out = _HERE / "_native"
    so = out / "liboracle.so"
    flags = "-O3 -march=native -fPIC -std=c11 -ffp-contract=off"
    try:
        out.mkdir(exist_ok=True)
        srcs = [_HERE / f for f in _SRCS]
        if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
            subprocess.run(["gcc", *flags.split(), "-shared", "-o", str(so), *[str(_HERE / f) for f in _SRCS if f.endswith(".c")],
                            "-lm", "-lpthread"], check=True, capture_output=True)
        return so, flags
    except Exception:
        return build(), "-O3 -march=x86-64-v3 -fPIC -std=c11 -ffp-contract=off (portable build: native compile failed)"


class Oracle:
    """The reference algorithm on the CPU, for one network."""

    def __init__(self, net: Network, native: bool = False):
        self.net = net
        if native:
            so, self.cflags = build_native()
        else:
            so, self.cflags = build(), "-O3 -march=x86-64-v3 -fPIC -std=c11 -ffp-contract=off"
        self.lib = C.CDLL(str(so))
        self._keep = []
        self._c = self._pack(net)
        L = self.lib
        L.orc_run_model.restype = C.c_int
        L.orc_run_grid.restype = C.c_int
        L.orc_run_grid_timed.restype = C.c_int
        L.orc_get_rates.restype = C.c_int
        L.orc_get_odes.restype = C.c_int
        L.orc_h2_photo_diss_rate.restype = C.c_double
        L.orc_h2_photo_diss_rate.argtypes = [C.c_double] * 4
        L.orc_co_photo_diss_rate.restype = C.c_double
        L.orc_co_photo_diss_rate.argtypes = [C.c_double] * 4
        L.orc_set_deadline.restype = None
        L.orc_set_deadline.argtypes = [C.c_double]

    FLAG_DEADLINE = -98

    def set_deadline(self, seconds: float) -> None:
        """Benchmark guard (not reference behaviour): models still running `seconds` from now stop with
        FLAG_DEADLINE; 0 switches the guard off."""
        self.lib.orc_set_deadline(float(seconds))

    # ------------------------------------------------------------------
    def _arr(self, a, dtype):
        a = np.ascontiguousarray(a, dtype=dtype)
        self._keep.append(a)
        return a.ctypes.data_as(_pd if dtype == np.float64 else _pi)

    def _pack(self, net: Network) -> _OrcNetwork:
        ls, lr, gs, gr = net.stoichiometry()
        lo = [-1 if net.type_ranges[t] is None else net.type_ranges[t][0] for t in TYPE_NAMES]
        hi = [-1 if net.type_ranges[t] is None else net.type_ranges[t][1] for t in TYPE_NAMES]
        named = []
        for n in NAMED:
            if n.startswith("nR_"):
                named.append(net.reaction_idx[n])
            else:
                named.append(net.species_idx[n])
        refr = net.refractory_list if len(net.refractory_list) else np.zeros(1, np.int32)
        is_ion = np.array([1 if "+" in n else 0 for n in net.names], np.int32)
        s = _OrcNetwork()
        s.nspec, s.nreac, s.nice, s.nsurf = net.nspec, net.nreac, len(net.ice_list), len(net.surface_list)
        s.ngar = net.gar_params.shape[0]
        s.n_loss, s.n_gain, s.n_refractory = len(ls), len(gs), len(net.refractory_list)
        f64, i32 = np.float64, np.int32
        s.mass = self._arr(net.mass, f64)
        s.atom_counts = self._arr(net.atom_counts, i32)
        s.surface_list = self._arr(net.surface_list, i32)
        s.bulk_list = self._arr(net.bulk_list, i32)
        s.ice_list = self._arr(net.ice_list, i32)
        s.gas_ice_list = self._arr(net.gas_ice_list, i32)
        s.binding_energy = self._arr(net.binding_energy, f64)
        s.formation_enthalpy = self._arr(net.formation_enthalpy, f64)
        s.re = self._arr(net.re, i32)
        s.pr = self._arr(net.pr, i32)
        for k in ("alpha", "beta", "gama", "min_temps", "max_temps", "reduced_masses"):
            setattr(s, k, self._arr(getattr(net, k), f64))
        s.extrapolate = self._arr(net.extrapolate.astype(np.int32), i32)
        s.rtype = self._arr(net.rtype, i32)
        s.flux_factors = self._arr(net.flux_factors(5), i32)
        s.loss_species, s.loss_reaction = self._arr(ls, i32), self._arr(lr, i32)
        s.gain_species, s.gain_reaction = self._arr(gs, i32), self._arr(gr, i32)
        s.freeze_partners = self._arr(net.freeze_partners, i32)
        s.gar_params = self._arr(net.gar_params, f64)
        s.type_lo, s.type_hi = self._arr(lo, i32), self._arr(hi, i32)
        s.named = self._arr(named, i32)
        s.refractory_list = self._arr(refr, i32)
        s.is_ion = self._arr(is_ion, i32)
        return s

    # ------------------------------------------------------------------
    def getydot(self, rate, y, blr, cov, safe_mantle, safe_bulk, dens):
        ydot = np.zeros(self.net.neq)
        sg = C.c_double(0.0)
        rate = np.ascontiguousarray(rate, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        self.lib.orc_getydot(C.byref(self._c), rate.ctypes.data_as(_pd), y.ctypes.data_as(_pd),
                             C.c_double(blr), C.c_double(cov), C.c_double(safe_mantle),
                             C.c_double(safe_bulk), C.c_double(dens), ydot.ctypes.data_as(_pd), C.byref(sg))
        return ydot, sg.value

    def get_rates(self, params: np.ndarray, y: np.ndarray) -> np.ndarray:
        params = np.ascontiguousarray(params, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        out = np.zeros(self.net.nreac)
        self.lib.orc_get_rates(C.byref(self._c), params.ctypes.data_as(_pd), y.ctypes.data_as(_pd),
                               out.ctypes.data_as(_pd))
        return out

    def get_odes(self, params: np.ndarray, y: np.ndarray) -> np.ndarray:
        params = np.ascontiguousarray(params, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        out = np.zeros(self.net.neq)
        self.lib.orc_get_odes(C.byref(self._c), params.ctypes.data_as(_pd), y.ctypes.data_as(_pd),
                              out.ctypes.data_as(_pd))
        return out

    def probe_rhs(self, params: np.ndarray, y: np.ndarray) -> np.ndarray:
        """F at exactly the given state (no pre-integration)."""
        params = np.ascontiguousarray(params, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        out = np.zeros(self.net.neq)
        self.lib.orc_probe_rhs(C.byref(self._c), params.ctypes.data_as(_pd), y.ctypes.data_as(_pd),
                               out.ctypes.data_as(_pd))
        return out

    def run_model(self, kind: int, params: np.ndarray, y0=None, timepoints: int = 500, rates: bool = False,
                  pp_grid=None, pp_coldens: bool = False):
        """Returns dict(flag, y_final, phys_final, physics[nrows,8], abund[nrows,nspec], rates, stats, t_diss).
        kind 5 (postprocess) takes the tracer history pp_grid [10, ntime] (time in s, density, gas T, dust T,
        radfield, zeta, N_H, N_H2, N_CO, N_C; the columns are only read with pp_coldens)."""
        net = self.net
        params = np.ascontiguousarray(params, np.float64)
        assert params.shape == (NPARAM,)
        yfin = np.zeros(net.neq)
        pfin = np.zeros(8)
        phys = np.zeros((timepoints + 1, 8))
        chem = np.zeros((timepoints + 1, net.nspec))
        rts = np.zeros((timepoints + 1, net.nreac)) if rates else None
        nrows = C.c_int(0)
        tdiss = C.c_double(0.0)
        st = OrcStats()
        y0p = None
        if y0 is not None:
            y0 = np.ascontiguousarray(y0, np.float64)
            y0p = y0.ctypes.data_as(_pd)
        ntime, gp = 0, None
        if pp_grid is not None:
            pp_grid = np.ascontiguousarray(pp_grid, np.float64)
            assert pp_grid.ndim == 2 and pp_grid.shape[0] == 10
            ntime, gp = pp_grid.shape[1], pp_grid.ctypes.data_as(_pd)
        self.lib.orc_run_model_pp.restype = C.c_int
        flag = self.lib.orc_run_model_pp(
            C.byref(self._c), C.c_int(kind), params.ctypes.data_as(_pd), y0p, yfin.ctypes.data_as(_pd),
            pfin.ctypes.data_as(_pd), C.c_int(timepoints), phys.ctypes.data_as(_pd), chem.ctypes.data_as(_pd),
            rts.ctypes.data_as(_pd) if rates else None, C.byref(nrows), C.byref(tdiss), C.byref(st),
            C.c_int(ntime), gp, C.c_int(1 if pp_coldens else 0))
        n = nrows.value
        return dict(flag=flag, y_final=yfin, phys_final=pfin, physics=phys[:n], abund=chem[:n],
                    rates=rts[:n] if rates else None,
                    stats={k: getattr(st, k) for k, _ in OrcStats._fields_}, dissipation_time=tdiss.value)

    def run_grid(self, kind: int, params: np.ndarray, y0=None, nthreads: int | None = None, timed: bool = False):
        """params [NPARAM, ncell]; returns (y_final[ncell,neq], phys[ncell,8], flag[ncell], stats) and, with
        `timed`, the wall seconds of every model (-1 for cells the deadline guard never let start)."""
        net = self.net
        params = np.ascontiguousarray(params, np.float64)
        ncell = params.shape[1]
        yfin = np.zeros((ncell, net.neq))
        pfin = np.zeros((ncell, 8))
        flag = np.zeros(ncell, np.int32)
        stats = (OrcStats * ncell)()
        y0p = None
        if y0 is not None:
            y0 = np.ascontiguousarray(y0, np.float64)
            assert y0.shape == (ncell, net.neq)
            y0p = y0.ctypes.data_as(_pd)
        if nthreads is None:
            nthreads = os.cpu_count() or 1
        secs = np.full(ncell, -1.0)
        self.lib.orc_run_grid_timed(C.byref(self._c), C.c_int(kind), C.c_int64(ncell), params.ctypes.data_as(_pd),
                                    y0p, yfin.ctypes.data_as(_pd), pfin.ctypes.data_as(_pd),
                                    flag.ctypes.data_as(_pi), stats, C.c_int(nthreads), secs.ctypes.data_as(_pd))
        st = np.array([[getattr(s, k) for k, _ in OrcStats._fields_] for s in stats], np.int64)
        if timed:
            return yfin, pfin, flag, st, secs
        return yfin, pfin, flag, st
