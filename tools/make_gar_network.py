"""Third network: the default network plus grain-assisted recombination (GAR, Weingartner & Draine 2001;
`Makerates/data/grain_assisted_recombination/gar_settings.yaml` of the reference): 335 species / 3209 reactions.

Authoring-container only.  Runs the reference's own MakeRates (Python, imported from /root/reference/src
without its compiled wrapper) into a scratch directory, then
  * parses the produced network.f90 into uclchem_b200/networks/gar.json (what our generator reads),
  * evaluates the produced odes.f90 (tools/ref_odes_eval.py: read and interpreted, never copied) on six
    random states -> tests/golden/getydot_cases_gar.npz, the RHS known answers for this network.
"""
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
REF = Path("/root/reference")
OUT = Path("/tmp/mk_gar"); OUT.mkdir(parents=True, exist_ok=True)

settings = OUT / "settings.yaml"
settings.write_text(f"""species_file: {REF}/Makerates/data/default/default_species.csv
database_reaction_file: {REF}/Makerates/data/databases/umist22.csv
database_reaction_type: UMIST
custom_reaction_file: {REF}/Makerates/data/default/default_grain_network.csv
custom_reaction_type: UCL
output_directory: {OUT}
add_crp_photo_to_grain: False
gar_reaction_file: {REF}/Makerates/data/grain_assisted_recombination/grain_assisted_recombination.csv
gar_reaction_type: UCL
grain_assisted_recombination_file: {REF}/Makerates/data/databases/weingartner01_grain_assisted_recombination.yaml
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
net.to_json(ROOT / "uclchem_b200/networks/gar.json")
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
np.savez_compressed(ROOT / "tests/golden/getydot_cases_gar.npz",
                    **{f"{k}_{i}": np.asarray(c[k]) for i, c in enumerate(cases) for k in c})
print("rhs cases", [float(c["ydot"][net.species_idx["nsurface"]]) for c in cases])
