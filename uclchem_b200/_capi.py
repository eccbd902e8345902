"""ctypes binding of the C ABI declared in ``include/uclgpu.h``.

This is the Python stand-in for the Fortran ``ISO_C_BINDING`` interface block a
UCLCHEM maintainer adds to ``wrap.f90``: the same entry
points, the same plain pointers.  There is no fallback of any kind: if the CUDA
library has not been built, or no sm_100 device is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from .params import NPARAM

_LIBDIR = Path(__file__).resolve().parent / "lib"
_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)

N_PHYS = 8
STAT_FIELDS = ("nst", "nfe", "nje", "nlu", "nni", "ncfn", "netf", "nintervals", "nsing", "nmaxcor", "ndiverge",
               "nfailcall", "cyc_rates", "cyc_rhs", "cyc_jac", "cyc_factor", "cyc_dense", "cyc_solve", "cyc_total",
               "reserved")

# every symbol include/uclgpu.h declares (checked by the CPU test-suite)
EXPORTED = [
    "uclgpu_init", "uclgpu_shutdown", "uclgpu_strerror", "uclgpu_nspec", "uclgpu_nreac",
    "uclgpu_species_name", "uclgpu_network_tag", "uclgpu_default_params", "uclgpu_run_grid",
    "uclgpu_get_rates", "uclgpu_get_odes", "uclgpu_run_grid_device", "uclgpu_last_kernel_ms",
    "uclgpu_naug", "uclgpu_probe_rhs", "uclgpu_probe_newton", "uclgpu_work_model", "uclgpu_fp64_peak",
]


class UclgpuStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in STAT_FIELDS]


class UclgpuOpts(C.Structure):
    _fields_ = [("timepoints", C.c_int32), ("physics_traj", _pd), ("chem_traj", _pd), ("rates_traj", _pd),
                ("dissipation_time", _pd), ("reserved0", C.c_int32), ("step_budget", C.c_int32),
                ("cost_hint", _pd), ("y0_index", _pi), ("ny0", C.c_int64),
                ("n_coeff", C.c_int64), ("coeff_which", _pi), ("coeff_index", _pi), ("coeff_value", _pd),
                ("pp_grid", _pd), ("pp_ntime", C.c_int32), ("pp_coldens", C.c_int32),
                ("chunk_bytes", C.c_int64)]


class UclgpuError(RuntimeError):
    pass


def library_path(tag: str = "default") -> Path:
    return _LIBDIR / f"libuclgpu_{tag}.so"


class Library:
    """One compiled network (= one shared library, like one uclchemwrap build)."""

    def __init__(self, tag: str = "default"):
        path = library_path(tag)
        if not path.exists():
            raise UclgpuError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(uclchem_b200 has no CPU fallback)")
        self.path = path
        L = self.lib = C.CDLL(str(path))
        L.uclgpu_strerror.restype = C.c_char_p
        L.uclgpu_species_name.restype = C.c_char_p
        L.uclgpu_network_tag.restype = C.c_char_p
        for f in ("uclgpu_init", "uclgpu_nspec", "uclgpu_nreac", "uclgpu_default_params", "uclgpu_run_grid",
                  "uclgpu_get_rates", "uclgpu_get_odes", "uclgpu_run_grid_device", "uclgpu_last_kernel_ms",
                  "uclgpu_probe_rhs", "uclgpu_probe_newton", "uclgpu_naug"):
            getattr(L, f).restype = C.c_int
        L.uclgpu_default_params.argtypes = [C.c_int64, _pd]
        L.uclgpu_run_grid.argtypes = [C.c_int, C.c_int64, _pd, _pd, _pd, _pd, _pi, C.POINTER(UclgpuStats),
                                      C.POINTER(UclgpuOpts)]
        L.uclgpu_get_rates.argtypes = [C.c_int64, _pd, _pd, _pd]
        L.uclgpu_get_odes.argtypes = [C.c_int64, _pd, _pd, _pd]
        L.uclgpu_probe_rhs.argtypes = [C.c_int64, _pd, _pd, _pd]
        L.uclgpu_probe_newton.argtypes = [C.c_int64, _pd, _pd, C.c_double, _pd, _pd]
        L.uclgpu_run_grid_device.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.uclgpu_last_kernel_ms.argtypes = [C.c_int, _pd, C.POINTER(C.c_int64)]
        L.uclgpu_init.argtypes = [C.c_int, _pi]
        L.uclgpu_work_model.restype = C.c_int
        L.uclgpu_work_model.argtypes = [_pd]
        L.uclgpu_fp64_peak.restype = C.c_int
        L.uclgpu_fp64_peak.argtypes = [C.c_int, _pd]
        self.nspec = L.uclgpu_nspec()
        self.nreac = L.uclgpu_nreac()
        self.neq = self.nspec + 1
        self.naug = L.uclgpu_naug()
        self.tag = L.uclgpu_network_tag().decode()
        self.species = [L.uclgpu_species_name(i).decode() for i in range(self.nspec)]

    # ------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise UclgpuError(f"uclgpu error {rc}: {self.lib.uclgpu_strerror(rc).decode()}")

    def init(self, devices=None):
        if devices is None:
            self._check(self.lib.uclgpu_init(0, None))
        else:
            arr = np.ascontiguousarray(devices, np.int32)
            self._check(self.lib.uclgpu_init(len(arr), arr.ctypes.data_as(_pi)))

    def shutdown(self):
        self.lib.uclgpu_shutdown()

    def default_params(self, ncell: int) -> np.ndarray:
        out = np.zeros((NPARAM, ncell))
        self._check(self.lib.uclgpu_default_params(ncell, out.ctypes.data_as(_pd)))
        return out

    def run_grid(self, kind: int, params: np.ndarray, y0=None, timepoints: int = 0, want_physics=False,
                 want_chem=False, want_rates=False, step_budget: int = 0, cost_hint=None,
                 chunk_bytes: int = 0, y0_index=None, coefficients=None, pp_grid=None, pp_coldens=False):
        params = np.ascontiguousarray(params, np.float64)
        assert params.ndim == 2 and params.shape[0] == NPARAM
        ncell = params.shape[1]
        y_final = np.zeros((ncell, self.neq))
        phys = np.zeros((ncell, N_PHYS))
        flag = np.zeros(ncell, np.int32)
        stats = (UclgpuStats * max(1, ncell))()
        y0p = None
        opts = UclgpuOpts()
        if y0 is not None:
            y0 = np.ascontiguousarray(y0, np.float64)
            if y0_index is not None:      # y0 is a table of starting states shared between cells
                y0_index = np.ascontiguousarray(y0_index, np.int32)
                assert y0_index.shape == (ncell,) and y0.ndim == 2 and y0.shape[1] == self.neq
                assert ncell == 0 or (0 <= y0_index.min() and y0_index.max() < y0.shape[0])
                opts.y0_index = y0_index.ctypes.data_as(_pi)
                opts.ny0 = y0.shape[0]
            else:
                assert y0.shape == (ncell, self.neq)
            y0p = y0.ctypes.data_as(_pd)
        opts.timepoints = timepoints
        opts.step_budget = step_budget
        opts.chunk_bytes = chunk_bytes
        if pp_grid is not None:   # postprocess: [ncell, 10, ntime] tracer histories
            pp_grid = np.ascontiguousarray(pp_grid, np.float64)
            assert pp_grid.ndim == 3 and pp_grid.shape[:2] == (ncell, 10)
            opts.pp_grid, opts.pp_ntime, opts.pp_coldens = pp_grid.ctypes.data_as(_pd), pp_grid.shape[2], int(bool(pp_coldens))
        if coefficients:   # [(which, 0-based reaction, value), ...] with which in 0 alpha / 1 beta / 2 gamma
            cw = np.ascontiguousarray([c[0] for c in coefficients], np.int32)
            ci = np.ascontiguousarray([c[1] for c in coefficients], np.int32)
            cv = np.ascontiguousarray([c[2] for c in coefficients], np.float64)
            opts.n_coeff = len(cw)
            opts.coeff_which, opts.coeff_index, opts.coeff_value = cw.ctypes.data_as(_pi), ci.ctypes.data_as(_pi), cv.ctypes.data_as(_pd)
        if cost_hint is not None:
            cost_hint = np.ascontiguousarray(cost_hint, np.float64)
            assert cost_hint.shape == (ncell,)
            opts.cost_hint = cost_hint.ctypes.data_as(_pd)
        out = {}
        tdiss = np.zeros(ncell)
        opts.dissipation_time = tdiss.ctypes.data_as(_pd)
        if want_physics:
            out["physics"] = np.zeros((ncell, timepoints + 1, N_PHYS))
            opts.physics_traj = out["physics"].ctypes.data_as(_pd)
        if want_chem:
            out["abund"] = np.zeros((ncell, timepoints + 1, self.nspec))
            opts.chem_traj = out["abund"].ctypes.data_as(_pd)
        if want_rates:
            out["rates"] = np.zeros((ncell, timepoints + 1, self.nreac))
            opts.rates_traj = out["rates"].ctypes.data_as(_pd)
        self._check(self.lib.uclgpu_run_grid(kind, ncell, params.ctypes.data_as(_pd), y0p,
                                             y_final.ctypes.data_as(_pd), phys.ctypes.data_as(_pd),
                                             flag.ctypes.data_as(_pi), stats, C.byref(opts)))
        st = np.array([[getattr(s, k) for k in STAT_FIELDS] for s in stats[:ncell]], np.int64).reshape(ncell, len(STAT_FIELDS))
        out.update(y_final=y_final, phys_final=phys, flag=flag, stats=st, dissipation_time=tdiss)
        return out

    def _probe(self, fn, params, y, width, *extra):
        params = np.ascontiguousarray(params, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        ncell = params.shape[1]
        assert y.shape == (ncell, self.neq)
        out = np.zeros((ncell, width))
        self._check(fn(ncell, params.ctypes.data_as(_pd), y.ctypes.data_as(_pd), *extra, out.ctypes.data_as(_pd)))
        return out

    def get_rates(self, params, y):
        return self._probe(self.lib.uclgpu_get_rates, params, y, self.nreac)

    def get_odes(self, params, y):
        return self._probe(self.lib.uclgpu_get_odes, params, y, self.neq)

    def probe_rhs(self, params, y):
        return self._probe(self.lib.uclgpu_probe_rhs, params, y, self.neq)

    def probe_newton(self, params, y, gamma, b):
        b = np.ascontiguousarray(b, np.float64)
        return self._probe(self.lib.uclgpu_probe_newton, params, y, self.naug, C.c_double(gamma),
                           b.ctypes.data_as(_pd))

    def last_kernel_ms(self, dev: int = 0):
        ms = C.c_double(0.0)
        n = C.c_int64(0)
        self._check(self.lib.uclgpu_last_kernel_ms(dev, C.byref(ms), C.byref(n)))
        return ms.value, n.value


_LIBS: dict = {}


def get_library(tag: str = "default") -> Library:
    if tag not in _LIBS:
        _LIBS[tag] = Library(tag)
    return _LIBS[tag]
