"""bench.py host logic that does not need a GPU: workload definition, the bounded CPU work-queue sample (oracle
wall-clock guard, per-model timing) and the assembly of the JSON line from measured quantities."""
import argparse
import ctypes as C
import json

import numpy as np

import bench
from uclchem_b200._capi import STAT_FIELDS, get_library
from uclchem_b200.params import PARAM_INDEX


def test_workload_is_the_whole_config2_grid_in_four_interleaved_steps():
    p = bench.config2_params()
    assert p.shape[1] == 25 * 20 * 20
    sl = bench.Config2(0, 1, argparse.Namespace(cells=0)).slices
    assert sorted(np.concatenate(sl).tolist()) == list(range(10000)) and all(len(s) == 2500 for s in sl)
    # every slice sees every density, temperature and zeta
    for s in sl:
        assert len(np.unique(p[PARAM_INDEX["initialdens"], s])) == 25 and len(np.unique(p[PARAM_INDEX["initialtemp"], s])) == 20
        assert len(np.unique(p[PARAM_INDEX["zeta"], s])) == 20
    # ... and slice q holds the cells with (density index + zeta index) mod 4 == q: density / zeta neighbours are in other slices
    D, T, Z = np.unravel_index(np.arange(10000), bench.GRID_SHAPE)
    for q, s in enumerate(sl):
        assert ((D[s] + Z[s]) % bench.NSLICE == q).all()
    # N ranks: the zeta axis is refined N-fold and dealt round-robin; rank grids are disjoint, same density/temperature axes
    z = [np.unique(bench.config2_params(rank=r, world=4)[PARAM_INDEX["zeta"]]) for r in range(4)]
    allz = np.sort(np.concatenate(z))
    assert len(allz) == 80 and len(np.unique(allz)) == 80 and np.isclose(allz[0], 1.0) and np.isclose(allz[-1], 1e3)
    assert all(bench.config2_params(rank=r, world=4).shape[1] == 10000 for r in range(4))
    order = bench.cpu_sample_order(10000)
    assert sorted(order.tolist()) == list(range(10000)) and np.array_equal(order, bench.cpu_sample_order(10000))


def test_oracle_work_queue_sample_is_bounded_and_timed(oracle):
    """All threads pull from one queue until the bound; models still running then stop with flag -98 and are left
    out; cells the queue never reached keep -98 and a negative time."""
    from uclchem_b200.params import params_from_dict
    p = params_from_dict({"initialDens": np.full(12, 1e4), "initialTemp": np.linspace(10, 30, 12), "finalTime": 1.0})
    r = bench.run_oracle_sample(p, 2, 30.0, cells=np.arange(12))
    assert r["n_finished"] == 12 and r["n_cut"] == 0 and r["rate"] > 0 and r["ms_per_bdf_step"] > 0
    assert abs(r["rate"] - 2 * 12 / r["core_seconds"]) < 1e-9
    p = params_from_dict({"initialDens": np.full(6, 1e5), "initialTemp": np.full(6, 20.0), "finalTime": 1.0e6})
    r = bench.run_oracle_sample(p, 2, 0.5, cells=np.arange(6))
    assert r["n_finished"] == 0 and r["n_cut"] == 2 and r["rate"] == 0.0
    d = bench.cpu_baseline_dict(r, 2, 0.5)
    assert d["kind"] == "port" and d["cores"] == 2 and "2 still running" in d["sample"] and "-march=native" in d["compile_flags"]


def test_json_line_assembly():
    lib = get_library("default")          # loads without a GPU; uclgpu_work_model is host-only
    flop = (C.c_double * 8)()
    assert lib.lib.uclgpu_work_model(flop) == 0
    m = 91
    assert flop[2] < flop[6] and abs((flop[6] - flop[2]) - (2 * m ** 3 - (2 * m ** 3) // 3)) < 1e5   # dense block counted as an LU
    assert flop[5] == 64 * 8 + 336 * 8 + 64 + 4 + 160
    ncell = 100
    stats = np.zeros((ncell, len(STAT_FIELDS)), np.int64)
    S = {k: i for i, k in enumerate(STAT_FIELDS)}
    stats[:, S["nst"]] = 8000; stats[:, S["nfe"]] = 12000; stats[:, S["nje"]] = 300; stats[:, S["nlu"]] = 2000
    stats[:, S["nni"]] = 12000; stats[:, S["nintervals"]] = 46
    a = argparse.Namespace(steps=2, warmup=3)
    line = bench.assemble_line(a=a, world=2, n_ok=190, n_cells=200, n_budget=8, workload={"workload": "x"}, wall_s=5.0,
                               kernel_ms=3900.0, launches=2, stats=stats, clocks={"sm_mhz": 1900.0},
                               work_model=list(flop), fp64_peak_tflops=30.0, h2d_bytes=51200, d2h_bytes=300000,
                               cpu={"value": 1.0, "unit": bench.UNIT, "cores": 4, "kind": "port", "sample": "s"},
                               parity={"gpu_flags_nonzero": 10}, traffic=123, stat_fields=STAT_FIELDS,
                               per_rank_kernel_ms=[1900.0, 1950.0], gather_bytes=7)
    line = json.loads(json.dumps(line))
    assert line["value"] == 190 / 3.9 and line["e2e"]["value"] == 190 / 5.0          # only integrated models count
    assert line["unit"] == "models/s" and line["n_gpus"] == 2 and line["scaling"] == "weak" and line["dtype"] == "f64"
    assert line["ms_per_step"] == 2500.0 and line["kernel_ms_per_step"] == 1950.0 and line["gpu_launches"] == 2
    assert line["models"] == {"cells_processed": 200, "integrated": 190, "abandoned_at_step_budget": 8, "other_failures": 2}
    r = line["roofline"]
    assert r["bound"] == "fp64" and r["unit"] == "TFLOP/s" and r["traffic"] == 123
    w_flop = ncell * (12000 * flop[0] + 300 * flop[1] + 2000 * flop[2] + 12000 * flop[3] + 46 * flop[4])
    assert abs(r["achieved"] - w_flop / 3.9 / 1e12) < 1e-9 and abs(r["frac"] - r["achieved"] / 30.0) < 1e-15   # per GPU
    assert r["executed_flop_per_model"] > r["algorithmic_flop_per_model"]
    h = r["hbm"]
    assert abs(h["achieved"] - flop[5] * 100 / 3.9 / 1e9) < 1e-12 and abs(h["frac"] - h["achieved"] / h["peak"]) < 1e-15
    assert line["solver"]["steps_per_model"] == 8000 and line["cpu_baseline"]["kind"] == "port"
    for key in ("metric", "steps", "warmup", "higher_is_better", "vs_baseline", "data", "config", "clocks", "parity"):
        assert key in line


def test_cost_hint_comes_from_neighbouring_cells_of_the_same_pass_only():
    """The hint the bench passes to uclgpu_opts.cost_hint for a cell is the largest step-attempt count among its 26
    neighbours in the (density, temperature, zeta) index box that were integrated EARLIER IN THE SAME PASS over the
    grid -- never the cell's own cost (the bench returns to the same cells every NSLICE steps; nobody integrates the
    same model twice), and main() forgets everything when a new pass begins."""
    w = bench.Config2(0, 1, argparse.Namespace(cells=0))
    flat = np.arange(10000).reshape(bench.GRID_SHAPE)
    att = np.full(10000, np.nan)
    assert w.cost_hint(0, att) is None                       # first step of a pass: nothing known
    att[w.slices[0]] = 100.0
    c = flat[10, 7, 3]                                       # (10 + 3) mod 4 = 1: a cell of slice 1
    assert c in w.slices[1]
    att[flat[11, 7, 3]] = 9.0e4                              # density neighbour, (11 + 3) mod 4 = 2 ... pretend it is known
    att[flat[10, 8, 4]] = 5.0e3
    h1 = w.cost_hint(1, att)
    assert h1[w.slices[1].tolist().index(c)] == 9.0e4       # the largest neighbour wins
    att[c] = 1.0e6                                           # the cell's OWN figure is never its hint
    assert w.cost_hint(1, att)[w.slices[1].tolist().index(c)] == 9.0e4
    # neighbours do not wrap around the grid edges: cell (0, 0, 0) must not see (24, 19, 19)
    att[:] = np.nan
    att[flat[24, 19, 19]] = 7.0e4
    att[flat[1, 1, 1]] = 3.0
    assert w.cost_hint(0, att)[w.slices[0].tolist().index(flat[0, 0, 0])] == 3.0
    # cells without a known neighbour get the median of the known hints
    h = w.cost_hint(2, att)
    assert np.isfinite(h).all()
    # neighbourhood_max itself: brute force on random data
    rng = np.random.default_rng(0)
    v = np.where(rng.random(10000) < 0.3, rng.random(10000), np.nan)
    got = bench.neighbourhood_max(v).reshape(bench.GRID_SHAPE)
    V = v.reshape(bench.GRID_SHAPE)
    for (i, j, k) in [(0, 0, 0), (24, 19, 19), (5, 6, 7), (12, 0, 19), (24, 10, 0)]:
        box = [V[a, b, c] for a in range(max(0, i - 1), min(25, i + 2)) for b in range(max(0, j - 1), min(20, j + 2))
               for c in range(max(0, k - 1), min(20, k + 2)) if (a, b, c) != (i, j, k) and not np.isnan(V[a, b, c])]
        assert (np.isnan(got[i, j, k]) and not box) or got[i, j, k] == max(box)
    # the debug subset (--cells) has no grid structure: no hint
    assert bench.Config2(0, 1, argparse.Namespace(cells=50)).cost_hint(1, np.ones(50)) is None


def test_within_pass_hint_schedules_like_a_perfect_hint_on_the_measured_costs():
    """List scheduling (148 queues, longest expected first) of the per-cell costs measured on the B200 for the whole
    config-2 grid (profiles/r02_grid_per_cell_costs.npz): with the bench's within-pass neighbourhood hint, steps 2-4
    of a pass finish within 3 % of what a perfect hint (the cells' own step counts) gives; the unhinted first step
    does not (that is the tail the hint removes)."""
    import heapq
    from conftest import ROOT
    d = np.load(ROOT / "profiles" / "r02_grid_per_cell_costs.npz")
    att, cost = d["attempts"].astype(float), d["sm_cycles"] / 1.965e9
    w = bench.Config2(0, 1, argparse.Namespace(cells=0))
    generic = np.log10(w.params[PARAM_INDEX["initialdens"]])

    def makespan(cells, key):
        h = [0.0] * 148
        heapq.heapify(h)
        for c in cells[np.argsort(-key, kind="stable")]:
            heapq.heappush(h, heapq.heappop(h) + cost[c])
        return max(h)

    attempts = np.full(10000, np.nan)
    for k in range(bench.NSLICE):
        idx = w.slices[k]
        hint = w.cost_hint(k, attempts)
        got, best = makespan(idx, generic[idx] if hint is None else hint), makespan(idx, att[idx])
        if k == 0:
            assert hint is None and got > 1.15 * best
        else:
            assert hint is not None and got < 1.03 * best, (k, got, best)
        attempts[idx] = att[idx]
