"""bench.py host logic that does not need a GPU: workload definition, the bounded CPU sample (oracle
wall-clock guard) and the assembly of the JSON line from measured quantities."""
import argparse
import ctypes as C
import json

import numpy as np

import bench
from uclchem_b200._capi import STAT_FIELDS, get_library


def test_workload_is_config2_minus_stall_cells():
    p = bench.config2_params()
    assert p.shape[1] == 25 * 20 * 20
    pool = bench.bounded_cells(p.shape[1])
    heavy = np.load(bench.ROOT / "tools" / "config2_heavy_cells.npy")
    assert len(pool) == 10000 - len(heavy) == 9839 and not set(pool) & set(heavy)
    idx = bench.sample_cells(len(pool), 16)
    assert idx[0] == 0 and idx[-1] == len(pool) - 1 and len(set(idx)) == 16
    assert np.array_equal(bench.sample_cells(len(pool), 16, offset=3), (idx + 3) % len(pool))


def test_oracle_deadline_guard(oracle):
    """A model still running at the bound stops with flag -98; without the bound it runs to the end."""
    from uclchem_b200.params import params_from_dict
    p = params_from_dict({"initialDens": [1e4, 1e5], "initialTemp": [10.0, 20.0], "finalTime": 1.0e6})
    oracle.set_deadline(0.3)
    _, _, flag, st = oracle.run_grid(0, p, nthreads=2)
    oracle.set_deadline(0.0)
    assert (flag == oracle.FLAG_DEADLINE).all() and (st[:, 0] > 0).all()
    p[:, :] = params_from_dict({"initialDens": [1e4, 1e5], "initialTemp": [10.0, 20.0], "finalTime": 1.0})
    _, _, flag, _ = oracle.run_grid(0, p, nthreads=2)
    assert (flag == 0).all()
    assert "2 of them were still running" in bench.sample_text(np.arange(3), np.array([0, -98, -98]), 5.0, 4.0)


def test_json_line_assembly():
    lib = get_library("default")          # loads without a GPU; uclgpu_work_model is host-only
    flop = (C.c_double * 8)()
    assert lib.lib.uclgpu_work_model(flop) == 0
    ncell = 100
    stats = np.zeros((ncell, len(STAT_FIELDS)), np.int64)
    S = {k: i for i, k in enumerate(STAT_FIELDS)}
    stats[:, S["nst"]] = 8000; stats[:, S["nfe"]] = 12000; stats[:, S["nje"]] = 300; stats[:, S["nlu"]] = 2000
    stats[:, S["nni"]] = 12000; stats[:, S["nintervals"]] = 46
    a = argparse.Namespace(steps=2, warmup=3)
    line = bench.assemble_line(a=a, world=2, ncell=ncell, workload={"workload": "x"}, dt=4.0, kernel_ms=3900.0,
                               dt_e2e=5.0, launches=2, launches_e2e=2, stats=stats, flags=np.zeros(ncell, np.int32),
                               clocks={"sm_mhz": 1900.0}, work_model=list(flop), fp64_peak_tflops=30.0,
                               h2d_bytes=51200, d2h_bytes=300000,
                               cpu={"value": 1.0, "unit": bench.UNIT, "cores": 4, "kind": "port", "sample": "s"},
                               parity={"flags_nonzero": 0}, traffic=123, stat_fields=STAT_FIELDS)
    line = json.loads(json.dumps(line))
    assert line["value"] == 2 * ncell * 2 / 4.0 and line["e2e"]["value"] == 2 * ncell * 2 / 5.0
    assert line["unit"] == "models/s" and line["n_gpus"] == 2 and line["scaling"] == "weak" and line["dtype"] == "f64"
    assert line["ms_per_step"] == 2000.0 and line["kernel_ms_per_step"] == 1950.0 and line["gpu_launches"] == 4
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["traffic"] == 123
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-15
    w_bytes = ncell * 46 * flop[5]
    assert abs(r["achieved"] - w_bytes / 1.95 / 1e9) < 1e-12
    f = line["fp64"]
    w_flop = ncell * (12000 * flop[0] + 300 * flop[1] + 2000 * flop[2] + 12000 * flop[3] + 46 * flop[4])
    assert abs(f["achieved_tflops"] - w_flop / 1.95 / 1e12) < 1e-9 and abs(f["frac"] - f["achieved_tflops"] / 30.0) < 1e-15
    assert line["solver"]["steps_per_model"] == 8000 and line["cpu_baseline"]["kind"] == "port"
    for key in ("metric", "steps", "warmup", "higher_is_better", "vs_baseline", "data", "config", "clocks", "parity"):
        assert key in line
