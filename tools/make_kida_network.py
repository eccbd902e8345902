"""Fourth network: the default species and grain network on the KIDA gas-phase database (kida.uva.2024) instead of
UMIST22 -- the database that carries the IONOPOL1 / IONOPOL2 reaction types (io_functions.py:184-189, rates.f90:301-314).
Authoring-container only; same recipe as tools/make_gar_network.py.

OUTCOME (this container): NOT BUILDABLE.  The reference's KIDA reader fails on both KIDA files it ships
(`gas_reactions_kida.uva.2024.in` and `legacy/kida.uva.2014.dat`): `IndexError: list index out of range` in
`io_functions.check_reaction` (row[12] of a 12-entry KIDA row), so the reference's MakeRates produces no network
with IONOPOL1 / IONOPOL2 reactions to pin on.  Kept as the evidence for DESIGN.md section 9 item 3.
"""
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
REF = Path("/root/reference")
OUT = Path("/tmp/mk_kida"); OUT.mkdir(parents=True, exist_ok=True)

settings = OUT / "settings.yaml"
settings.write_text(f"""species_file: {REF}/Makerates/data/default/default_species.csv
database_reaction_file: {REF}/Makerates/data/databases/gas_reactions_kida.uva.2024.in
database_reaction_type: KIDA
custom_reaction_file: {REF}/Makerates/data/default/default_grain_network.csv
custom_reaction_type: UCL
output_directory: {OUT}
add_crp_photo_to_grain: False
enable_rates_to_disk: False
""")
pkg = types.ModuleType("uclchem"); pkg.__path__ = [str(REF / "src/uclchem")]; sys.modules["uclchem"] = pkg
from uclchem.makerates import run_makerates  # noqa: E402
try:
    run_makerates(str(settings))
except FileNotFoundError as e:   # the last step edits src/uclchem/constants.py relative to the cwd: not needed
    print("ignored:", e)

from uclchem_b200.network import Network  # noqa: E402
from ref_odes_eval import compile_getydot  # noqa: E402
net = Network.from_network_f90(OUT / "network.f90")
net.to_json(ROOT / "uclchem_b200/networks/kida.json")
print("network", net.nspec, net.nreac)

f = compile_getydot(OUT / "odes.f90")
rng = np.random.default_rng(20261018)
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
        lo, hi = net.type_ranges["FREEZE"]; rate[lo:hi + 1] = 0.0
        lo, hi = net.type_ranges["THERM"]; rate[lo:hi + 1] *= 1e8
    safe_mantle = max(1e-30, y[net.species_idx["nsurface"]])
    safe_bulk = max(1e-30, y[net.species_idx["nbulk"]])
    blr = min(1.0, 10 ** rng.uniform(-2, 0.5))
    cov = 10 ** rng.uniform(-3, 0)
    ydot = f(rate, y, blr, cov, safe_mantle, safe_bulk, dens)
    cases.append(dict(y=y, rate=rate, blr=blr, cov=cov, safe_mantle=safe_mantle, safe_bulk=safe_bulk, dens=dens, ydot=ydot))
np.savez_compressed(ROOT / "tests/golden/getydot_cases_kida.npz",
                    **{f"{k}_{i}": np.asarray(c[k]) for i, c in enumerate(cases) for k in c})
print("rhs cases", [float(c["ydot"][net.species_idx["nsurface"]]) for c in cases])
