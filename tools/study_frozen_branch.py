"""CPU study (no GPU): 'one transfer branch per corrector pass'.  The stall cells of DESIGN.md section 6 sit in a
quasi-steady state of the mantle in which the uncorrected surface growth S is a difference of 1e-13 terms at the
1e-21 level -- below the integrator's own tolerance -- so the branch `S < 0` of the three-phase transfer
(odes.f90:4815-5153) flips from one Newton iterate to the next.  Here the branch is decided ONCE per corrector pass,
by the sign of S at the predicted state, and kept for the Jacobian and all Newton iterates of that pass (oracle
experiment switches g_orc_force_branch / orc_set_pass_hook; the reference's rule is untouched in every test).

    python tools/study_frozen_branch.py dens temp zeta [final_time] [freeze=1|0]
"""
import ctypes as C, os, sys, functools, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
print = functools.partial(print, flush=True)
from oracle.oracle import Oracle
from uclchem_b200.network import load_default
from uclchem_b200.params import params_from_dict
dens, temp, zeta = (float(a) for a in sys.argv[1:4])
tfin = float(sys.argv[4]) if len(sys.argv) > 4 else 1e6
freeze = int(sys.argv[5]) if len(sys.argv) > 5 else 1
net = load_default(); orc = Oracle(net)
L = orc.lib
L.orc_ctx_surfgrowth.restype = C.c_double; L.orc_ctx_surfgrowth.argtypes = [C.c_void_p]
force = C.c_int.in_dll(L, "g_orc_force_branch")
HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_int)
count = {"pos": 0, "neg": 0}

def hook(ctx, phase):
    if phase == 0:
        force.value = 0
    else:
        sgn = -1 if L.orc_ctx_surfgrowth(ctx) < 0 else 1
        force.value = sgn
        count["neg" if sgn < 0 else "pos"] += 1

cb = HOOK(hook)
if freeze:
    L.orc_set_pass_hook(cb)
p = params_from_dict(dict({"initialDens": dens, "initialTemp": temp, "zeta": zeta, "finalTime": tfin},
                          **json.loads(os.environ.get("STUDY_EXTRA", '{"radfield": 1.0, "baseAv": 2.0, "rout": 0.05}'))))
orc.set_deadline(float(os.environ.get("STUDY_SECONDS", "900")))
t0 = time.time()
r = orc.run_model(0, p[:, 0])
force.value = 0
print(f"freeze {freeze}: flag {r['flag']} stats {r['stats']} passes {count}  {time.time() - t0:.0f} s")
np.save(f"/tmp/frozen_{freeze}_{dens:g}_{temp:g}_{zeta:g}.npy", r["y_final"])
