import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
from uclchem_b200.network import load_default
from uclchem_b200.params import params_from_dict
from uclchem_b200._capi import get_library
from oracle.oracle import Oracle
net = load_default(); orc = Oracle(net); lib = get_library(); lib.init()
p1 = params_from_dict({"initialDens": 1e4, "initialTemp": 10.0, "finalTime": 1e-7})
r = orc.run_model(0, p1[:, 0])
y0 = r['y_final'].copy(); y0[-1] = 1e4
iB, iS = net.species_idx['nbulk'], net.species_idx['nsurface']
y0[iB] = y0[net.bulk_list].sum(); y0[iS] = y0[net.surface_list].sum()
ewt = 1 / (1e-8 * np.abs(y0) + np.maximum(1e-14 * y0, 1e-25))
f = lambda y: lib.probe_rhs(p1, y[None, :])[0]
f0 = f(y0)
for h in (0.0469, 0.5, 2.0):
    y = y0 + h * f0
    out = []
    for it in range(4):
        G = y - y0 - h * f(y)
        d = lib.probe_newton(p1, y0[None, :], h, (-G)[None, :])[0][:336]
        y = y + d
        out.append(np.sqrt(np.mean((d * ewt) ** 2)))
    print(h, ' '.join(f"{v:.3e}" for v in out))
