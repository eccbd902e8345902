"""Where and why does the engine stall?  (run on the GPU box)
1. all 10^4 config-2 cells under a step budget -> per-cell counters (gpurun_out/stall_grid.npz);
2. for the first few cells that exhausted the budget: a single-cell run with a full Newton-iteration trace
   (gpurun_out/stall_trace_<cell>.npz), then a second, identical run that dumps the solver state (y, Nordsieck
   columns, saved Jacobian, Newton matrix, right-hand side) at an iteration deep inside the stalled DVODE call
   (gpurun_out/stall_dump_<cell>.bin) for offline comparison with the oracle's F and its finite-difference Jacobian.
usage: gpu_stall_probe.py [budget] [ncells_to_trace] [tag]"""
import os, sys, time, functools
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
print = functools.partial(print, flush=True)
from bench import config2_params
from uclchem_b200._capi import Library, STAT_FIELDS
budget = int(float(sys.argv[1])) if len(sys.argv) > 1 else 30000
ntrace = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tag = sys.argv[3] if len(sys.argv) > 3 else "default"
out = ROOT / "gpurun_out"; out.mkdir(exist_ok=True)
lib = Library(tag); lib.init([0])
P = config2_params()
lib.run_grid(0, P[:, ::68][:, :148], step_budget=2000)
t = time.time(); o = lib.run_grid(0, P, step_budget=budget); dt = time.time() - t
ms, _ = lib.last_kernel_ms(0)
st = o["stats"]; S = {k: st[:, i] for i, k in enumerate(STAT_FIELDS)}
sec = S["cyc_total"] / 1.965e9
print(f"[{tag}] budget {budget}: wall {dt:.1f} s kernel {ms/1e3:.1f} s flags {dict(zip(*np.unique(o['flag'], return_counts=True)))}")
print("   per-cell s: sum/148 %.1f max %.1f; nst pct 1/50/99/max" % (sec.sum() / 148, sec.max()),
      np.percentile(S["nst"], [1, 50, 99, 100]).astype(int), " cyc/step %.0f" % (S["cyc_total"].sum() / S["nst"].sum()))
np.savez_compressed(out / "stall_grid.npz", stats=st, flag=o["flag"], y_final=o["y_final"], kernel_ms=ms)
heavy = np.where(o["flag"] == -5)[0]
# spread the traced cells over the stall regions (temperature index = (cell // 20) % 20)
pick = []
for c in heavy:
    if all(abs(((c // 20) % 20) - ((q // 20) % 20)) >= 2 for q in pick):
        pick.append(int(c))
    if len(pick) == ntrace:
        break
print("heavy cells", len(heavy), "tracing", pick)
cap = 4 * budget
for c in pick:
    p = np.ascontiguousarray(P[:, c:c + 1])
    os.environ["UCLGPU_TRACE"] = str(cap); os.environ["UCLGPU_TRACE_FILE"] = str(out / "trace_tmp.bin")
    os.environ.pop("UCLGPU_DUMP_AT", None)
    r = lib.run_grid(0, p, step_budget=budget)
    tr = np.fromfile(out / "trace_tmp.bin").reshape(-1, 12)
    tr = tr[tr[:, 1] != 0]
    np.savez_compressed(out / f"stall_trace_{c}.npz", trace=tr, stats=r["stats"], flag=r["flag"], params=p)
    nstc = np.floor(tr[:, 7])
    deep = np.where(nstc >= 3000)[0]
    print(f"cell {c}: flag {r['flag'][0]} records {len(tr)} nst {r['stats'][0][0]}; first record deep in a stalled call: {deep[0] if len(deep) else None}")
    if len(deep):
        k = int(deep[0])
        print("   trace there: tn %.6e h %.3e nq %d m %d del %.3e dcon %.3e rc %.4f S %.3e" % tuple(tr[k][[0, 1, 2, 3, 4, 5, 6, 10]]))
        os.environ["UCLGPU_DUMP_AT"] = str(k); os.environ["UCLGPU_DUMP_FILE"] = str(out / f"stall_dump_{c}.bin")
        lib.run_grid(0, p, step_budget=budget)
        os.environ.pop("UCLGPU_DUMP_AT", None)
    (out / "trace_tmp.bin").unlink(missing_ok=True)
os.environ.pop("UCLGPU_TRACE", None)
