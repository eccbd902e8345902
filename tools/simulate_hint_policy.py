"""Offline check of the bench's cost-hint policy (no GPU): list scheduling of the MEASURED per-cell costs of the whole
config-2 grid (profiles/r02_grid_per_cell_costs.npz: step attempts and SM cycles of all 10^4 cells, one B200 run of
tools/gpu_grid_full.py with the 1e5-step budget) on 148 queues, longest-expected-first by the key a step of the bench
would pass, using bench.py's own slicing and hint functions.  Prints the makespan of every step of two passes next to
what a perfect hint and a perfectly balanced step would give."""
import argparse, heapq, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
sys.argv = sys.argv[:1]
import numpy as np
import bench
from uclchem_b200.params import PARAM_INDEX

d = np.load(ROOT / "profiles" / "r02_grid_per_cell_costs.npz")
att, cost = d["attempts"].astype(float), d["sm_cycles"] / 1.965e9          # seconds on one SM
w = bench.Config2(0, 1, argparse.Namespace(cells=0))
P = w.params
generic = np.log10(P[PARAM_INDEX["initialdens"]]) + 0.2 * np.log10(P[PARAM_INDEX["finaltime"]])   # uclgpu.cu cost_order


def makespan(cells, key):
    order = cells[np.argsort(-key, kind="stable")]
    h = [0.0] * 148
    heapq.heapify(h)
    for c in order:
        heapq.heappush(h, heapq.heappop(h) + cost[c])
    return max(h)


attempts = np.full(len(att), np.nan)
for k in range(2 * bench.NSLICE):
    if k % bench.NSLICE == 0:
        attempts[:] = np.nan
    idx = w.slices[k % bench.NSLICE]
    hint = w.cost_hint(k, attempts)
    got = makespan(idx, generic[idx] if hint is None else hint)
    print(f"step {k}: {'no hint' if hint is None else 'hinted '}  makespan {got:6.2f} s   perfect hint {makespan(idx, att[idx]):6.2f} s   "
          f"balanced {cost[idx].sum() / 148:6.2f} s")
    attempts[idx] = att[idx]
