"""Build tests/golden/ from the reference checkout (authoring container only).

Inputs (read-only, /root/reference):
  examples/example-output/{static,phase1,phase2}-full.dat   golden trajectories G1-G3
  examples/example-output/{startcollapse,startstatic,shockstart}.dat   abundance vectors G2/G4
  src/fortran_src/odes.f90   evaluated (not copied) by tools/ref_odes_eval.py -> RHS known answers
Outputs: small .npz tables; see tests/golden/README.md.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
from uclchem_b200.network import Network  # noqa: E402
from ref_odes_eval import compile_getydot  # noqa: E402

REF = Path("/root/reference")
OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)

net = Network.from_network_f90(REF / "src/fortran_src/network.f90")


def read_full(path):
    lines = path.read_text().splitlines()
    header = [h.strip() for h in lines[0].split(",")]
    assert header[8:] == net.names, "golden header != network species order"
    data = np.array([[float(v) for v in ln.split(",")] for ln in lines[1:]])
    return data[:, :8], data[:, 8:]


for name in ("static", "phase1", "phase2"):
    phys, ab = read_full(REF / f"examples/example-output/{name}-full.dat")
    np.savez_compressed(OUT / f"{name}_full.npz", physics=phys, abund=ab)
    print(name, phys.shape, ab.shape)

for name in ("startcollapse", "startstatic", "shockstart"):
    v = np.array([float(x) for x in (REF / f"examples/example-output/{name}.dat").read_text().split(",")])
    assert v.shape == (net.nspec,)
    np.save(OUT / f"{name}.npy", v)

# RHS known answers from the reference's own generated odes.f90
f = compile_getydot()
rng = np.random.default_rng(20261017)
cases = []
for k in range(6):
    y = 10 ** rng.uniform(-14, -4, net.neq)
    y[net.species_idx["nh2"]] = 0.4
    y[net.species_idx["nh"]] = 10 ** rng.uniform(-5, -1)
    y[net.species_idx["nbulk"]] = y[net.bulk_list].sum()
    y[net.species_idx["nsurface"]] = y[net.surface_list].sum()
    dens = 10 ** rng.uniform(2, 7)
    y[net.nspec] = dens
    rate = 10 ** rng.uniform(-14, -9, net.nreac)
    if k % 2 == 1:  # force the mantle-loss branch (YDOT(SURFACE) < 0)
        lo, hi = net.type_ranges["FREEZE"]
        rate[lo:hi + 1] = 0.0
        lo, hi = net.type_ranges["THERM"]
        rate[lo:hi + 1] *= 1e8
    safe_mantle = max(1e-30, y[net.species_idx["nsurface"]])
    safe_bulk = max(1e-30, y[net.species_idx["nbulk"]])
    blr = min(1.0, 10 ** rng.uniform(-2, 0.5))
    cov = 10 ** rng.uniform(-3, 0)
    ydot = f(rate, y, blr, cov, safe_mantle, safe_bulk, dens)
    cases.append(dict(y=y, rate=rate, blr=blr, cov=cov, safe_mantle=safe_mantle, safe_bulk=safe_bulk,
                      dens=dens, ydot=ydot))
np.savez_compressed(OUT / "getydot_cases.npz", **{f"{k}_{i}": np.asarray(c[k]) for i, c in enumerate(cases) for k in c})
print("rhs cases", len(cases), [float(c["ydot"][net.species_idx["nsurface"]]) for c in cases])
