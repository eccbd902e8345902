"""Network contract reader: turns MakeRates' ``network.f90`` into plain arrays.

MakeRates (reference ``src/uclchem/makerates/io_functions.py:801-951``,
``write_network_file``) publishes a chemical network as Fortran array
constructors in ``network.f90``.  That file is the *authoritative* statement of
every constant the hot path reads: the float literals in it are default-real
(single precision) literals written with ``%.4e``
(``io_functions.py:1021,1047``), so the value the reference Fortran actually
computes with is ``double(float32(round_5sig(x)))`` -- not what
``reactions.csv`` holds.  We therefore parse the file
itself and reproduce that rounding.

The result is a :class:`Network` of numpy arrays (0-based indices, ``-1`` for
"no species") plus the derived structure the B200 back-end needs: per-reaction
type ids, flux factor lists (``reaction.py:779-819``), and loss/gain
stoichiometry (``io_functions.py:562-581``).  Nothing here is imported by the
reference; it is the front half of our MakeRates CUDA back-end
(:mod:`uclchem_b200.makerates_cuda`).
"""
from __future__ import annotations

import json
import re
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

# Reaction type ids, ordered as the ``<type>Reacs`` ranges appear in
# network.f90 (io_functions.py:939-949 iterates reaction_types + TWOBODY).
TYPE_NAMES = [
    "PHOTON", "CRP", "CRPHOT", "FREEZE", "DESORB", "THERM", "DESOH2", "DESCR",
    "DEUVCR", "H2FORM", "ER", "ERDES", "LH", "LHDES", "BULKSWAP", "SURFSWAP",
    "IONOPOL1", "IONOPOL2", "CRS", "EXSOLID", "EXRELAX", "GAR", "TWOBODY",
]
TYPE_ID = {n: i for i, n in enumerate(TYPE_NAMES)}

# Pseudo-factor slots appended after the NEQ real unknowns in the "extended
# state" vector used by flux tables (see Network.flux_factors):
#   y_ext[NEQ+0] = 1.0
#   y_ext[NEQ+1] = bulkLayersReciprocal
#   y_ext[NEQ+2] = 1/safeMantle
#   y_ext[NEQ+3] = totalSwap/safeMantle
EXT_ONE, EXT_BLR, EXT_INV_SM, EXT_SWAP_SM = 0, 1, 2, 3
N_EXT = 4


def _join_continuations(text: str) -> str:
    """Fortran free-form continuation: trailing '&' + leading '&' splice exactly."""
    out = []
    cur = ""
    for raw in text.splitlines():
        line = raw.rstrip("\n")
        if cur:
            s = line.lstrip()
            if s.startswith("&"):
                s = s[1:]
            line = s
        if line.rstrip().endswith("&"):
            cur += line.rstrip()[:-1]
            continue
        out.append(cur + line)
        cur = ""
    if cur:
        out.append(cur)
    return "\n".join(out)


def _f32(x: float) -> float:
    """Value of a default-real Fortran literal promoted to double."""
    return float(np.float32(x))


_ARRAY_RE = re.compile(
    r"::\s*(\w+)\s*\(([\d,\s]+)\)\s*=\s*(?:RESHAPE\(\s*)?\(/(.*?)/\)", re.S
)


def parse_network_f90(path: str | Path) -> dict:
    """Return {name: python list / dict} for every array and named index."""
    text = _join_continuations(Path(path).read_text())
    out: dict = {}
    for line in text.splitlines():
        m = _ARRAY_RE.search(line)
        if m:
            name, shape, body = m.group(1), m.group(2), m.group(3)
            decl = line.split("::")[0].upper()
            if "CHARACTER" in decl:
                vals = re.findall(r'"([^"]*)"', body)
                out[name] = [v.strip() for v in vals]
            elif "LOGICAL" in decl:
                out[name] = [v.strip().upper() == ".TRUE." for v in body.split(",")]
            elif "INTEGER" in decl:
                out[name] = [int(v) for v in body.split(",")]
            else:
                out[name] = [_f32(float(v)) for v in body.split(",")]
            continue
        if "PARAMETER" in line.upper() and "::" in line and "=" in line and "(/" not in line:
            rhs = line.split("::", 1)[1]
            pairs = re.findall(r"(\w+)\s*=\s*(-?\d+)\b", rhs)
            if pairs and all(p[0][0].lower() == "n" for p in pairs):
                out.setdefault("_named", {}).update({k: int(v) for k, v in pairs})
        m2 = re.search(r"THREE_PHASE\s*=\s*\.(\w+)\.", line)
        if m2:
            out["THREE_PHASE"] = m2.group(1).upper() == "TRUE"
    return out


@dataclass
class Network:
    """Plain-array view of one MakeRates network (all indices 0-based)."""

    names: list
    mass: np.ndarray            # [nspec]
    atom_counts: np.ndarray     # [nspec]
    surface_list: np.ndarray    # [nsurf] species idx of '#'
    bulk_list: np.ndarray       # [nbulk] species idx of '@'
    ice_list: np.ndarray        # [nice]
    gas_ice_list: np.ndarray    # [nice]
    binding_energy: np.ndarray  # [nice]
    formation_enthalpy: np.ndarray  # [nice]
    solid_fractions: np.ndarray
    mono_fractions: np.ndarray
    volcanic_fractions: np.ndarray
    refractory_list: np.ndarray  # empty if none
    re: np.ndarray              # [nreac,3]  (-1 = none)
    pr: np.ndarray              # [nreac,4]
    alpha: np.ndarray
    beta: np.ndarray
    gama: np.ndarray
    min_temps: np.ndarray
    max_temps: np.ndarray
    reduced_masses: np.ndarray
    extrapolate: np.ndarray     # bool
    freeze_partners: np.ndarray  # [nsurf] reaction idx
    gar_params: np.ndarray      # [ngar,7]
    type_ranges: dict           # type name -> (first, last) 0-based inclusive, or None
    species_idx: dict           # 'nh2' -> 0-based idx (nspec == "absent", the density slot)
    reaction_idx: dict          # 'nR_H2_hv' -> 0-based
    three_phase: bool = True
    # derived
    rtype: np.ndarray = field(default=None)      # [nreac] TYPE_ID
    body_count: np.ndarray = field(default=None)  # [nreac] number of *D factors

    @property
    def nspec(self) -> int:
        return len(self.names)

    @property
    def neq(self) -> int:
        return self.nspec + 1

    @property
    def nreac(self) -> int:
        return len(self.alpha)

    # ------------------------------------------------------------------
    @classmethod
    def from_network_f90(cls, path: str | Path) -> "Network":
        d = parse_network_f90(path)
        names = d["specname"]
        nspec = len(names)

        def idx(a):  # 1-based Fortran index (9999 = none) -> 0-based (-1)
            a = np.asarray(a, dtype=np.int64)
            return np.where(a >= 9999, -1, a - 1).astype(np.int32)

        ranges = {}
        for t in TYPE_NAMES:
            key = t.lower() + "Reacs"
            lo, hi = d[key]
            ranges[t] = None if lo >= 99999 else (lo - 1, hi - 1)
        named = d["_named"]
        sp_idx = {k: v - 1 for k, v in named.items() if not k.startswith("nR_")}
        re_idx = {k: v - 1 for k, v in named.items() if k.startswith("nR_")}
        refr = np.asarray(d["refractoryList"], dtype=np.int64)
        refr = (refr[refr > 0] - 1).astype(np.int32)
        net = cls(
            names=names,
            mass=np.asarray(d["mass"], dtype=np.float64),
            atom_counts=np.asarray(d["atomCounts"], dtype=np.int32),
            surface_list=idx(d["surfaceList"]),
            bulk_list=idx(d["bulkList"]),
            ice_list=idx(d["iceList"]),
            gas_ice_list=idx(d["gasIceList"]),
            binding_energy=np.asarray(d["bindingEnergy"], dtype=np.float64),
            formation_enthalpy=np.asarray(d["formationEnthalpy"], dtype=np.float64),
            solid_fractions=np.asarray(d["solidFractions"], dtype=np.float64),
            mono_fractions=np.asarray(d["monoFractions"], dtype=np.float64),
            volcanic_fractions=np.asarray(d["volcanicFractions"], dtype=np.float64),
            refractory_list=refr,
            re=np.stack([idx(d["re1"]), idx(d["re2"]), idx(d["re3"])], axis=1),
            pr=np.stack([idx(d["p1"]), idx(d["p2"]), idx(d["p3"]), idx(d["p4"])], axis=1),
            alpha=np.asarray(d["alpha"], dtype=np.float64),
            beta=np.asarray(d["beta"], dtype=np.float64),
            gama=np.asarray(d["gama"], dtype=np.float64),
            min_temps=np.asarray(d["minTemps"], dtype=np.float64),
            max_temps=np.asarray(d["maxTemps"], dtype=np.float64),
            reduced_masses=np.asarray(d["reducedMasses"], dtype=np.float64),
            extrapolate=np.asarray(d["ExtrapolateRates"], dtype=bool),
            freeze_partners=idx(d["freezePartners"]),
            gar_params=np.asarray(d["garParams"], dtype=np.float64).reshape(7, -1).T.copy(),
            type_ranges=ranges,
            species_idx=sp_idx,
            reaction_idx=re_idx,
            three_phase=d.get("THREE_PHASE", True),
        )
        assert len(net.mass) == nspec
        net._derive()
        return net

    # ------------------------------------------------------------------
    def _derive(self) -> None:
        nreac = self.nreac
        rtype = np.full(nreac, -1, dtype=np.int32)
        for t, r in self.type_ranges.items():
            if r is not None:
                rtype[r[0]: r[1] + 1] = TYPE_ID[t]
        assert (rtype >= 0).all(), "reaction outside every type range"
        self.rtype = rtype
        # reaction.py:93-103 -- number of density factors
        nre = (self.re >= 0).sum(axis=1)
        bc = nre - 1
        for t in ("DESOH2", "FREEZE"):
            bc = bc + (rtype == TYPE_ID[t])
        for t in ("LH", "LHDES"):
            bc = bc - (rtype == TYPE_ID[t])
        # reaction.py:792-794 -- GAR carries one more factor of density
        bc = bc + (rtype == TYPE_ID["GAR"])
        self.body_count = bc.astype(np.int32)

    # ------------------------------------------------------------------
    def flux_factors(self, width: int = 5) -> np.ndarray:
        """[nreac,width] indices into the extended state giving
        flux_r = rate_r * prod_k y_ext[f[r,k]]   (reaction.py:779-819).

        Extended state = y[0:neq] (density is y[neq-1]) followed by the N_EXT
        pseudo factors documented at the top of this module.  Unused slots
        point at the constant-one entry.
        """
        neq = self.neq
        one = neq + EXT_ONE
        dens = neq - 1
        F = np.full((self.nreac, width), one, dtype=np.int32)
        is_bulk = np.zeros(self.nspec + 1, dtype=bool)
        is_bulk[self.bulk_list] = True
        nh = self.species_idx["nh"]
        for r in range(self.nreac):
            t = TYPE_NAMES[self.rtype[r]]
            fac = [dens] * int(self.body_count[r])
            reacts = [int(s) for s in self.re[r] if s >= 0]
            if t == "H2FORM":
                reacts = reacts[:1]  # only one factor of H (reaction.py:812-814)
            fac += reacts
            if t == "BULKSWAP":
                fac.append(neq + EXT_BLR)
            elif t == "SURFSWAP":
                fac.append(neq + EXT_SWAP_SM)
            elif t in ("DEUVCR", "DESCR", "DESOH2", "ER", "ERDES"):
                fac.append(neq + EXT_INV_SM)
                if t == "DESOH2":
                    fac.append(nh)
            if t in ("LH", "LHDES") and is_bulk[self.re[r, 0]]:
                fac.append(neq + EXT_BLR)
            assert len(fac) <= width, (r, t, fac)
            F[r, : len(fac)] = fac
        return F

    def stoichiometry(self):
        """(loss_species, loss_reaction, gain_species, gain_reaction) term lists
        following io_functions.py:562-581 (one entry per occurrence; ER gas
        reactant losses are charged to the '#' partner)."""
        ls, lr, gs, gr = [], [], [], []
        name_to_idx = {n: i for i, n in enumerate(self.names)}
        surf = set(int(s) for s in self.surface_list) | set(int(s) for s in self.bulk_list)
        for r in range(self.nreac):
            t = TYPE_NAMES[self.rtype[r]]
            for s in self.re[r]:
                if s < 0:
                    continue
                s = int(s)
                if t == "ER" and s not in surf:
                    s = name_to_idx["#" + self.names[s]]
                ls.append(s)
                lr.append(r)
            for s in self.pr[r]:
                if s >= 0:
                    gs.append(int(s))
                    gr.append(r)
        return (np.asarray(ls, np.int32), np.asarray(lr, np.int32),
                np.asarray(gs, np.int32), np.asarray(gr, np.int32))

    # ------------------------------------------------------------------
    def element_matrix(self):
        """(elements, counts[nelem,nspec], charge[nspec]) parsed from species
        names the way species.py:263-355 / analysis.py:629-666 do.  BULK and
        SURFACE get zero rows."""
        elements = ["H", "HE", "C", "N", "O", "S", "SI", "MG", "CL", "P", "F", "D", "NA", "LI", "FE"]
        two = sorted([e for e in elements if len(e) == 2], key=len, reverse=True)
        counts = np.zeros((len(elements), self.nspec), dtype=np.float64)
        charge = np.zeros(self.nspec, dtype=np.float64)
        for i, raw in enumerate(self.names):
            if raw in ("BULK", "SURFACE"):
                continue
            n = raw.lstrip("#@")
            if n == "E-":
                charge[i] = -1
                continue
            if n.endswith("+"):
                charge[i] = 1
                n = n[:-1]
            elif n.endswith("-"):
                charge[i] = -1
                n = n[:-1]
            k = 0
            while k < len(n):
                sym = None
                for e in two:
                    if n.startswith(e, k):
                        sym = e
                        break
                if sym is None:
                    sym = n[k]
                k += len(sym)
                m = re.match(r"\d+", n[k:])
                mult = 1
                if m:
                    mult = int(m.group(0))
                    k += len(m.group(0))
                if sym in elements:
                    counts[elements.index(sym), i] += mult
                else:
                    raise ValueError(f"unknown element {sym!r} in {raw!r}")
        keep = counts.sum(axis=1) > 0
        return [e for e, k_ in zip(elements, keep) if k_], counts[keep], charge

    # ------------------------------------------------------------------
    def to_json(self, path: str | Path) -> None:
        d = {}
        for k, v in self.__dict__.items():
            if isinstance(v, np.ndarray):
                d[k] = {"dtype": str(v.dtype), "shape": list(v.shape), "data": v.ravel().tolist()}
            else:
                d[k] = v
        Path(path).write_text(json.dumps(d))

    @classmethod
    def from_json(cls, path: str | Path) -> "Network":
        d = json.loads(Path(path).read_text())
        kw = {}
        for k, v in d.items():
            if isinstance(v, dict) and "dtype" in v:
                kw[k] = np.asarray(v["data"], dtype=v["dtype"]).reshape(v["shape"])
            else:
                kw[k] = v
        kw["type_ranges"] = {k: (tuple(v) if v is not None else None) for k, v in kw["type_ranges"].items()}
        net = cls(**kw)
        return net


_DEFAULT_JSON = Path(__file__).parent / "networks" / "default.json"


def load_default() -> Network:
    """The reference's shipped default network (335 species / 3203 reactions)."""
    return Network.from_json(_DEFAULT_JSON)
