#!/usr/bin/env python
"""bench.py -- cloud models integrated per second on B200 (BASELINE.json metric).

Workload (N = 1): BASELINE config[1], the 10^4-point static-cloud grid (25 densities x 20 temperatures x 20
cosmic-ray rates, 1 Myr each, default network), ALL of its cells.  With N ranks the zeta axis is refined N-fold
(25 x 20 x 20N points, rank r owns every N-th zeta plane), so per-GPU work is fixed (weak scaling), the job
integrates N x 10^4 distinct models, and rank 0 gathers every model's final abundances and flag.

A *step* is one call of the hot path over one batch: an interleaved quarter of the rank's grid (cells with
(density index + zeta index) mod 4 == step mod 4: every density, temperature and zeta in every step), 2 500 models
per GPU; four consecutive steps are one pass over the grid.  (A whole pass is ~40 s, and the driver's `--steps 20
--warmup 5` has to finish in minutes; a quarter keeps ~17 cells per SM in the work queue.)  From the second step
of a pass on the call carries `uclgpu_opts.cost_hint`: per cell the largest step count measured on its neighbouring
grid cells EARLIER IN THE SAME PASS (what a user who integrates a grid in a few chunks knows); nothing is carried
from one pass to the next, and a cell's own earlier integration is never used.

The reference algorithm does not terminate in bounded work on ~1 % of these cells: the three-phase surface/bulk
transfer has a kink at zero net surface growth, DVODE hits MXSTEP in every retry there, and UCLCHEM's own stall
guard (chemistry.f90:224-237) is disabled by its always-on `usepostprocess`.  Both arms run
every cell; `--step-budget` (default 100 000 BDF steps = 10 x MXSTEP, the most one output interval may take in
the reference) bounds a single cell like the reference's guard was meant to, and cells that exhaust it are
returned with INT_TOO_MANY_FAILS_ERROR and are NOT counted as integrated models -- their time is.

  value  : models integrated / CUDA-event time of k_integrate in the same calls (inputs resident in HBM)
  e2e    : models integrated / wall time of the calls a user makes -- uclgpu_run_grid through the C ABI with
           pinned HOST buffers (H2D, kernel, D2H inside) plus, for N > 1, the gather to rank 0
  roofline : algorithmic fp64 work (solver counters x per-operation counts from the MakeRates CUDA back-end,
           LU of the dense block at 2/3 m^3) over kernel time, against a DFMA peak measured in this process;
           `hbm` gives the algorithmic bytes against the measured copy bandwidth for completeness
  cpu_baseline : the oracle (CPU restatement of the reference algorithm: dense finite-difference Jacobian DVODE,
           cold restart per interval), work queue over all host cores, bounded sample

`--impl reference` times that CPU restatement alone (the reference Fortran cannot be compiled in this image: no
Fortran compiler).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def _baseline_metric():
    """The metric string of BASELINE.json (the line must carry the reference's headline metric verbatim)."""
    try:
        return json.loads((ROOT / "BASELINE.json").read_text())["metric"]
    except Exception:
        return "cloud models integrated/sec (10^5-pt grid, 1 Myr) at 1-8 B200 vs CPU DVODE"


METRIC = _baseline_metric()
UNIT = "models/s"
NSLICE = 4            # a step is one interleaved quarter of the rank's grid
STEP_BUDGET = 100000  # BDF steps per cell: 10 x MXSTEP


def config2_params(ncell_side=(25, 20, 20), rank=0, world=1):
    """regular grid, no RNG.  With `world` ranks the zeta axis has 20*world points
    and rank r takes every world-th one (world = 1: the config-2 grid itself)."""
    from uclchem_b200.params import params_from_dict
    nd, nt, nz = ncell_side
    dens = 10 ** np.linspace(3, 7, nd)
    temp = np.linspace(10, 100, nt)
    zeta = (10 ** np.linspace(0, 3, nz * world))[rank::world]
    D, T, Z = np.meshgrid(dens, temp, zeta, indexing="ij")
    return params_from_dict({"initialDens": D.ravel(), "initialTemp": T.ravel(), "zeta": Z.ravel(), "radfield": 1.0,
                             "baseAv": 2.0, "rout": 0.05, "finalTime": 1.0e6, "freefall": False,
                             "endAtFinalDensity": False})


def slice_cells(ncell, q, nslice=NSLICE):
    """Cells of step slice q: flat index congruent to q (mod nslice)."""
    return np.arange(q % nslice, ncell, nslice)


GRID_SHAPE = (25, 20, 20)   # config 2: densities x temperatures x zetas of one rank's grid (flat index, zeta fastest)


def checkerboard_slices(shape=GRID_SHAPE, nslice=NSLICE):
    """Config-2 steps: slice q holds the cells with (density index + zeta index) mod nslice == q.  Every slice sees
    every density, temperature and zeta (2 500 cells each), and every cell of slices 1..3 has density / zeta
    neighbours in the slices integrated before it in the same pass over the grid."""
    nd, nt, nz = shape
    D, _, Z = np.meshgrid(np.arange(nd), np.arange(nt), np.arange(nz), indexing="ij")
    lab = ((D + Z) % nslice).ravel()
    return [np.where(lab == q)[0] for q in range(nslice)]


def neighbourhood_max(values, shape=GRID_SHAPE):
    """Per cell the largest finite entry of `values` (flat, NaN = unknown) among its 26 neighbours in the
    (density, temperature, zeta) index box, the cell itself excluded; NaN where no neighbour is known."""
    nd, nt, nz = shape
    pad = np.full((nd + 2, nt + 2, nz + 2), -np.inf)
    pad[1:-1, 1:-1, 1:-1] = np.where(np.isnan(values), -np.inf, values).reshape(shape)
    best = np.full(shape, -np.inf)
    for a in (0, 1, 2):
        for b in (0, 1, 2):
            for c in (0, 1, 2):
                if (a, b, c) != (1, 1, 1):
                    best = np.maximum(best, pad[a:a + nd, b:b + nt, c:c + nz])
    return np.where(np.isinf(best), np.nan, best).ravel()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "500"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text())["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def log(msg):
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_sample_order(ncell):
    """Deterministic shuffle of the workload's cells: the CPU work queue takes them in this order, so that cheap
    and expensive regions of the grid are mixed and a time-bounded prefix is a fair sample."""
    return np.random.default_rng(20261017).permutation(ncell)


_ORACLES = {}


def oracle_for(tag):
    """The CPU restatement for one network (native build), cached."""
    if tag not in _ORACLES:
        from oracle.oracle import Oracle
        from uclchem_b200.network import Network, load_default
        net = load_default() if tag == "default" else Network.from_json(ROOT / "uclchem_b200" / "networks" / f"{tag}.json")
        _ORACLES[tag] = Oracle(net, native=True)
    return _ORACLES[tag]


def run_oracle_sample(params, cores, seconds, cells=None, kind=0, y0=None, tag="default"):
    """Time the CPU restatement on the workload: one work queue over `cores` threads (every core stays busy until
    the bound), cells in `cpu_sample_order`, stopped `seconds` after the start.  A model still running at the
    bound stops (oracle guard, flag -98) and is left out of numerator and denominator -- which favours the CPU
    figure, because the cells that get cut are the slow ones.  Returns a dict."""
    from oracle.oracle import Oracle
    order = cpu_sample_order(params.shape[1]) if cells is None else np.asarray(cells)
    order = order[: max(cores, int(cores * seconds / 1.5))]   # more than the queue can finish (>= 1.5 s per model)
    orc = oracle_for(tag)
    orc.set_deadline(seconds)
    t0 = time.perf_counter()
    y, _, flag, st, secs = orc.run_grid(kind, np.ascontiguousarray(params[:, order]), y0=None if y0 is None else y0[order],
                                        nthreads=cores, timed=True)
    wall = time.perf_counter() - t0
    orc.set_deadline(0.0)
    fin = (flag != Oracle.FLAG_DEADLINE) & (secs >= 0)
    cut = (flag == Oracle.FLAG_DEADLINE) & (secs >= 0)
    core_s = float(secs[fin].sum())
    n_fin = int(fin.sum())
    rate = cores * n_fin / core_s if core_s > 0 else 0.0
    nst = st[:, 0].astype(np.float64)
    return {"rate": rate, "wall": wall, "cells": order, "y": y, "flag": flag, "finished": fin, "n_finished": n_fin, "stats": st,
            "n_cut": int(cut.sum()), "core_seconds": core_s,
            "ms_per_bdf_step": 1e3 * core_s / nst[fin].sum() if n_fin else None,
            "steps_per_model": float(nst[fin].mean()) if n_fin else None,
            "seconds_per_model": core_s / n_fin if n_fin else None, "cflags": orc.cflags}


def cpu_baseline_dict(r, cores, seconds):
    sample = (f"work queue over {cores} threads on a fixed shuffle of the workload's cells, stopped after {seconds:.0f} s: "
              f"{r['n_finished']} models finished in {r['core_seconds']:.0f} core-seconds; {r['n_cut']} still running at "
              "the bound are left out of numerator and denominator (they are the slow ones: this favours the CPU figure); "
              "value = cores x finished / core-seconds of the finished models")
    return {"value": r["rate"], "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            "what": "CPU restatement of the reference algorithm (oracle/: DVODE MF=22, dense finite-difference Jacobian, "
                    "LINPACK LU, cold restart per interval); the reference Fortran cannot be built in this image",
            "ms_per_bdf_step_per_core": r["ms_per_bdf_step"], "bdf_steps_per_model": r["steps_per_model"],
            "seconds_per_model_per_core": r["seconds_per_model"], "compile_flags": r["cflags"]}


# ------------------------------------------------------------------------------------------------ JSON line
def assemble_line(*, a, world, n_ok, n_cells, n_budget, workload, wall_s, kernel_ms, launches, stats, clocks,
                  work_model, fp64_peak_tflops, h2d_bytes, d2h_bytes, cpu, parity, traffic, stat_fields,
                  per_rank_kernel_ms=None, gather_bytes=0):
    """The bench JSON line from measured quantities (pure function: unit-tested on the CPU).
    n_ok: models integrated (flag 0) over all ranks and timed steps; n_cells: cells processed; wall_s / kernel_ms:
    totals over the timed steps (max over ranks); stats: rank 0's per-cell counters of the timed steps."""
    f_rhs, f_jac, f_lu, f_solve, f_rates, b_cell, f_lu_exec, f_solve_exec = list(work_model)[:8]
    S = {k: stats[:, i].astype(np.float64).sum() for i, k in enumerate(stat_fields)}
    n_stat = max(1, stats.shape[0])
    w_flop = (S["nfe"] * f_rhs + S["nje"] * f_jac + S["nlu"] * f_lu + S["nni"] * f_solve + S["nintervals"] * f_rates)
    w_exec = (S["nfe"] * f_rhs + S["nje"] * f_jac + S["nlu"] * f_lu_exec + S["nni"] * f_solve_exec + S["nintervals"] * f_rates)
    kern_s = kernel_ms / 1e3
    # `stats` holds every cell rank 0 launched in the timed steps (all stages); kernel time is the max over ranks
    tfl = w_flop / kern_s / 1e12
    hbm_peak, how = peak_hbm()
    gbs = b_cell * n_stat / kern_s / 1e9
    roof = {"bound": "fp64", "achieved": tfl, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
            "frac": tfl / fp64_peak_tflops if fp64_peak_tflops else None, "traffic": traffic,
            "peak_source": "DFMA micro-benchmark run by this process (uclgpu_fp64_peak); MEASURED_PEAKS.json has no fp64 figure",
            "algorithmic_flop_per_model": w_flop / n_stat, "executed_flop_per_model": w_exec / n_stat,
            "note": "per GPU. Algorithmic counts: F_rhs, F_jac, F_lu = sparse factor terms + 2/3 m^3 for the dense "
                    "block, F_solve = substitution, F_rates (uclgpu_work_model); the explicit dense inverse and the "
                    "product-form programs this build executes are NOT counted. The kernel is latency bound "
                    "(one sequential chain of ~9 k BDF steps per model, working set chip-resident)",
            "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "peak_source": how,
                    "algorithmic_bytes_per_model": b_cell,
                    "note": "state is chip-resident: algorithmic HBM traffic is the parameters in and the result row out"}}
    return {
        "metric": METRIC, "value": n_ok / kern_s, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * wall_s / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload, "clocks": clocks,
        "e2e": {"value": n_ok / wall_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": int(d2h_bytes), "gather_bytes_per_step": int(gather_bytes)},
        "gpu_launches": int(launches),
        "kernel_ms_per_step": kernel_ms / a.steps,   # CUDA events around k_integrate on its stream (max over ranks)
        "per_rank_kernel_ms_per_step": per_rank_kernel_ms,
        "models": {"cells_processed": int(n_cells), "integrated": int(n_ok), "abandoned_at_step_budget": int(n_budget),
                   "other_failures": int(n_cells - n_ok - n_budget)},
        "roofline": roof, "cpu_baseline": cpu, "parity": parity,
        "solver": {"steps_per_model": S["nst"] / n_stat, "lu_per_model": S["nlu"] / n_stat,
                   "jac_per_model": S["nje"] / n_stat, "newton_iters_per_model": S["nni"] / n_stat,
                   "failed_dvode_calls": S["nfailcall"],
                   "sm_cycles_per_bdf_step": S["cyc_total"] / max(1.0, S["nst"])},
    }



# ------------------------------------------------------------------------------------------------ workloads
class StepSpec:
    """Inputs of one step: `params` [NPARAM, n] of the models that are counted, optionally preceded by a first
    stage (`stage1` = (kind, params)) whose final abundances are the starting states: cell c starts from
    stage-1 row y0_index[c]."""

    def __init__(self, kind, params, stage1=None, y0_index=None):
        self.kind, self.params, self.stage1, self.y0_index = kind, np.ascontiguousarray(params), stage1, y0_index


class Workload:
    def cost_hint(self, k, attempts):
        return None

    tag = "default"
    nslice = NSLICE
    step_budget = STEP_BUDGET  # BDF steps per model before it is abandoned: ~10x a typical model of the workload
    disjoint_ranks = True     # ranks take different steps of one grid (False: every rank has its own grid)

    def step(self, k):
        raise NotImplementedError


class Config1(Workload):
    """BASELINE configs[0]: the reference's own single static cloud (n_H = 1e4, T = 10 K, 1 Myr); a step is one
    such model per SM (148 identical models), the line also quotes the time of ONE model."""
    nslice = 1

    def __init__(self, rank, world, a):
        from uclchem_b200.params import params_from_dict
        self.p = params_from_dict({"initialDens": np.full(148, 1e4), "initialTemp": 10.0, "finalTime": 1.0e6,
                                   "freefall": False, "endAtFinalDensity": False})
        self.desc = {"workload": "config[0]: single static cloud model, n_H = 1e4, T = 10 K, 1 Myr, default network",
                     "step": "148 identical models, one per SM (a model is one sequential chain: its wall time is the step time)",
                     "cells_per_gpu_per_step": 148}

    def step(self, k):
        return StepSpec(0, self.p)


class Config2(Workload):
    disjoint_ranks = False

    def __init__(self, rank, world, a):
        self.params = config2_params(rank=rank, world=world)
        if a.cells:
            self.params = np.ascontiguousarray(self.params[:, np.linspace(0, self.params.shape[1] - 1, a.cells).astype(int)])
        n = self.params.shape[1]
        self.grid = not a.cells     # the regular 25 x 20 x 20 grid (False: debug subset, flat slicing, no hint)
        self.slices = checkerboard_slices() if self.grid else [slice_cells(n, q) for q in range(NSLICE)]
        self.desc = {"workload": "config[1]: 10^4-point static cloud grid (25 n_H x 20 T x 20 zeta), 1 Myr, default network "
                                 "335 species / 3203 reactions, reltol 1e-8, ALL cells"
                                 + (f"; {world} ranks: zeta axis refined to {20 * world} points, one 10^4-cell grid per GPU" if world > 1 else ""),
                     "step": f"one interleaved quarter of the rank's grid ({len(self.slices[0])} cells per GPU: (density index + zeta index) mod {NSLICE} "
                             f"= step mod {NSLICE}, every density, temperature and zeta in every step); {NSLICE} consecutive steps are one pass over the grid",
                     "cells_per_gpu_per_step": int(len(self.slices[0])), "cells_total": int(world * n)}

    def step(self, k):
        return StepSpec(0, self.params[:, self.slices[k % NSLICE]])

    def cost_hint(self, k, attempts):
        """Expected cost of the cells of step k from what THIS pass over the grid has measured so far: `attempts`
        holds, per flat grid index, the BDF step attempts of the cells integrated since the pass began (NaN = not
        yet; the caller forgets everything when a new pass begins).  The hint of a cell is the largest figure among
        its 26 neighbours in the (density, temperature, zeta) index box -- the cells on which DVODE stalls form
        bands along the density axis, so a stalled neighbour is a good warning -- and the median where no
        neighbour is known.  The first step of every pass runs without a hint (generic longest-first order of the
        library).  This is what a user who integrates a large grid in a few chunks can do with
        `uclgpu_opts.cost_hint`; nothing is ever taken from an earlier integration of the same model or grid."""
        if not self.grid or np.isnan(attempts).all():
            return None
        h = neighbourhood_max(attempts)[self.slices[k % NSLICE]]
        if np.isnan(h).all():
            return None
        return np.where(np.isnan(h), np.nanmedian(h), h)


class Config3(Workload):
    """BASELINE configs[2]: stage 1 free-fall clouds 1e2 -> n_f for 40 n_f x 25 zeta (1 000 runs),
    stage 2 hot_core(temp_indx 1..5, max_temperature 100..400 K in 20 steps) from each: 100 000 models.  A step
    takes 8 of the 1 000 (n_f, zeta) pairs (a fixed shuffle) with all 100 hot cores of each: 8 stage-1 models,
    then 800 hot cores whose starting states are read from the stage-1 result table by index."""
    nslice = 125
    step_budget = 1000000     # a 1 Myr hot core is 282 output intervals, ~1.1e5 steps when nothing stalls

    def __init__(self, rank, world, a):
        nf, ze = np.meshgrid(10 ** np.linspace(4, 7, 40), 10 ** np.linspace(0, 2, 25), indexing="ij")
        self.pairs = np.stack([nf.ravel(), ze.ravel()], axis=1)[np.random.default_rng(3).permutation(1000)]
        ti, mt = np.meshgrid(np.arange(1, 6), np.linspace(100, 400, 20), indexing="ij")
        self.ti, self.mt = ti.ravel().astype(float), mt.ravel()
        self.desc = {"workload": "config[2]: two-stage grid, 1 000 free-fall clouds (40 n_f x 25 zeta, 1e2 -> n_f) each feeding 100 "
                                 "hot cores (5 temp_indx x 20 max_temperature, freezeFactor 0, 1 Myr): 100 000 models",
                     "step": "8 (n_f, zeta) pairs of a fixed shuffle: 8 stage-1 models, then their 800 hot cores starting from the "
                             "stage-1 result table (uclgpu_opts.y0_index); 125 steps cover the grid; counted models = hot cores",
                     "cells_per_gpu_per_step": 800, "cells_total": 100000}

    def step(self, k):
        from uclchem_b200.params import params_from_dict
        pr = self.pairs[(k % self.nslice) * 8:(k % self.nslice) * 8 + 8]
        s1 = params_from_dict({"freefall": True, "endAtFinalDensity": True, "initialDens": 1e2, "finalDens": pr[:, 0],
                               "zeta": pr[:, 1], "initialTemp": 10.0, "finalTime": 1.0e7})
        idx = np.repeat(np.arange(8), 100)
        s2 = params_from_dict({"initialDens": pr[idx, 0], "zeta": pr[idx, 1], "initialTemp": 10.0, "finalTime": 1.0e6,
                               "freezeFactor": 0.0, "freefall": False, "endAtFinalDensity": False,
                               "temp_indx": np.tile(self.ti, 8), "max_temperature": np.tile(self.mt, 8)})
        return StepSpec(1, s2, stage1=(0, s1), y0_index=idx.astype(np.int32))


class Config4(Workload):
    """BASELINE configs[3]: 10^4 C-shocks, 100 velocities (10..45 km/s) x 100 pre-shock densities (10^3.5..10^6),
    pre-shock abundances from a free-fall collapse to each density, tolerances of notebooks/3_running_a_grid.py."""
    nslice = 10
    step_budget = 400000      # ~240 output intervals, 3.7e4 steps for a typical shock

    def __init__(self, rank, world, a):
        self.vs = np.linspace(10, 45, 100)
        self.n0 = 10 ** np.linspace(3.5, 6, 100)
        self.desc = {"workload": "config[3]: C-shock grid, 100 shock velocities (10-45 km/s) x 100 pre-shock densities (10^3.5-10^6), "
                                 "timestep_factor 0.01, finalTime 1e5 yr, reltol 1e-6, abstol_min 1e-20, pre-shock abundances from a "
                                 "free-fall collapse 1e2 -> n0",
                     "step": "10 densities (every tenth) x 100 velocities: 10 stage-1 free-fall models, then 1 000 shocks; 10 steps cover the grid",
                     "cells_per_gpu_per_step": 1000, "cells_total": 10000}

    def step(self, k):
        from uclchem_b200.params import params_from_dict
        n0 = self.n0[(k % 10)::10]
        s1 = params_from_dict({"freefall": True, "endAtFinalDensity": True, "initialDens": 1e2, "finalDens": n0,
                               "initialTemp": 10.0, "finalTime": 1.0e7})
        idx = np.repeat(np.arange(10), 100)
        s2 = params_from_dict({"initialDens": n0[idx], "initialTemp": 10.0, "finalTime": 1.0e5, "shock_vel": np.tile(self.vs, 10),
                               "timestep_factor": 0.01, "minimum_temperature": 0.0, "reltol": 1e-6, "abstol_min": 1e-20})
        return StepSpec(2, s2, stage1=(0, s1), y0_index=idx.astype(np.int32))


class Config5(Workload):
    """BASELINE configs[4]: 10^6 static clouds (100^3 over the config-2 ranges) on the network MakeRates generates
    with add_crp_photo_to_grain (335 species / 3453 reactions).  Run at the DEFAULT tolerances (reltol 1e-8,
    abstol_min 1e-25), which is stricter than SURVEY.md 8(d) asks (the reference's own test of this network loosens
    them to 1e-5 / 1e-15): the engine's no-pivot factorisation is validated at the default tolerances only
    (DESIGN.md section 9)."""
    tag = "crp_photo"
    nslice = 400

    def __init__(self, rank, world, a):
        self.axes = (10 ** np.linspace(3, 7, 100), np.linspace(10, 100, 100), 10 ** np.linspace(0, 3, 100))
        self.desc = {"workload": "config[4]: 10^6-point static cloud grid (100 n_H x 100 T x 100 zeta over the config-2 ranges), 1 Myr, "
                                 "crp-photo network 335 species / 3453 reactions, default tolerances (reltol 1e-8, abstol_min 1e-25: "
                                 "stricter than the reference's test of this network, which uses 1e-5 / 1e-15)",
                     "step": "every 400th cell of the grid (2 500 cells per GPU); 400 steps cover the grid",
                     "cells_per_gpu_per_step": 2500, "cells_total": 1000000}

    def step(self, k):
        from uclchem_b200.params import params_from_dict
        i = np.arange(k % 400, 1000000, 400)
        d, t, z = np.unravel_index(i, (100, 100, 100))
        return StepSpec(0, params_from_dict({"initialDens": self.axes[0][d], "initialTemp": self.axes[1][t], "zeta": self.axes[2][z],
                                             "radfield": 1.0, "baseAv": 2.0, "rout": 0.05, "finalTime": 1.0e6, "freefall": False,
                                             "endAtFinalDensity": False}))


WORKLOADS = {1: Config1, 2: Config2, 3: Config3, 4: Config4, 5: Config5}


_JSON_FD = None


def quiet_stdout():
    """stdout carries exactly ONE JSON line (the driver parses it): everything else that libraries print there
    (NCCL's version banner, for one) goes to stderr from here on; emit_json writes to the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_json(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=NSLICE)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", type=int, default=2, choices=sorted(WORKLOADS),
                    help="BASELINE.json configs[workload-1]; 2 (the 10^4 static-cloud grid) is the headline and the default")
    ap.add_argument("--step-budget", type=int, default=-1,
                    help="BDF steps after which a cell is abandoned (flag -5, not counted); 0 = the reference's unbounded crawl; "
                         "default: the workload's own (about ten times what a typical model of the workload needs)")
    ap.add_argument("--cpu-seconds", type=float, default=60.0, help="bound of the cpu_baseline sample")
    ap.add_argument("--no-cost-hint", action="store_true",
                    help="do not pass uclgpu_opts.cost_hint (step counts measured on neighbouring grid cells earlier in the same pass, workload 2)")
    ap.add_argument("--cells", type=int, default=0, help="debug: override the grid size (not a valid bench line)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[a.workload](rank, world, a)
    if a.step_budget < 0:
        a.step_budget = wl.step_budget
    workload = dict(wl.desc)
    workload.update({"step_budget": a.step_budget,
                     "timing": "L2 flushed (256 MiB write) between timed steps; a step's inputs are read once per cell",
                     "warmup_step": "one pass over a 296-cell stride of the first step's cells (same kernel and launch shape), "
                                    "step budget 20 000"})

    def step_id(k):   # ranks share one grid: consecutive steps are dealt round-robin; own grid per rank: same step index
        return k * world + rank if wl.disjoint_ranks else k

    # ------------------------------------------------------------------ reference arm (CPU restatement)
    if a.impl == "reference":
        if rank != 0:
            return
        cores = cpu_cores()
        seconds = float(min(200.0, max(60.0, 30.0 * a.steps)))   # one continuous sample, reported per step
        spec = wl.step(0)
        y0 = None
        if spec.stage1 is not None:   # starting states of the sampled cells: the CPU runs their first stage itself
            log("reference arm: first stage on the CPU")
            y0 = oracle_for(wl.tag).run_grid(spec.stage1[0], spec.stage1[1], nthreads=cores)[0][spec.y0_index]
        if a.warmup:
            log(f"reference arm: warm-up sample ({min(5.0, a.warmup * 1.0):.0f} s) on {cores} cores")
            run_oracle_sample(spec.params, cores, min(5.0, a.warmup * 1.0), kind=spec.kind, y0=y0, tag=wl.tag)
        log(f"reference arm: {seconds:.0f} s sample on {cores} cores")
        r = run_oracle_sample(spec.params, cores, seconds, kind=spec.kind, y0=y0, tag=wl.tag)
        cpu = cpu_baseline_dict(r, cores, seconds)
        workload["step"] = (f"one continuous {seconds:.0f} s sample of the first step's cells (shuffled work queue, all cores busy), "
                            f"reported as {a.steps} equal steps")
        emit_json({
            "impl": "reference", "metric": METRIC, "value": r["rate"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * r["wall"] / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload, "cpu_baseline": cpu,
            "e2e": {"value": r["rate"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    from uclchem_b200._capi import STAT_FIELDS, UclgpuOpts, UclgpuStats, get_library
    from uclchem_b200.sharding import gather_rows, max_over_ranks

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = get_library(wl.tag)          # raises if the CUDA extension is missing: no fallback
    lib.init([local_rank])
    neq = lib.neq
    nstat = len(STAT_FIELDS)
    pd_, pi_ = C.POINTER(C.c_double), C.POINTER(C.c_int32)

    class Batch:
        """Pinned host buffers of one call of uclgpu_run_grid (what a caller of the C ABI owns)."""

        def __init__(self, kind, p, y0_index=None):
            self.kind, self.n = kind, p.shape[1]
            self.params = torch.from_numpy(np.ascontiguousarray(p)).pin_memory()
            self.y = torch.zeros((self.n, neq), dtype=torch.float64).pin_memory()
            self.phys = torch.zeros((self.n, 8), dtype=torch.float64).pin_memory()
            self.flag = torch.zeros(self.n, dtype=torch.int32).pin_memory()
            self.stats = torch.zeros((self.n, nstat), dtype=torch.int64).pin_memory()
            self.y0_index = None if y0_index is None else torch.from_numpy(np.ascontiguousarray(y0_index, np.int32)).pin_memory()

        def run(self, budget, y0_table=None, cost_hint=None):
            opts = UclgpuOpts()
            opts.step_budget = budget
            if cost_hint is not None:
                self._hint = np.ascontiguousarray(cost_hint, np.float64)
                opts.cost_hint = C.cast(self._hint.ctypes.data, pd_)
            y0p = None
            if y0_table is not None:
                opts.y0_index = C.cast(self.y0_index.data_ptr(), pi_)
                opts.ny0 = y0_table.shape[0]
                y0p = C.cast(y0_table.data_ptr(), pd_)
            rc = lib.lib.uclgpu_run_grid(self.kind, self.n, C.cast(self.params.data_ptr(), pd_), y0p,
                                         C.cast(self.y.data_ptr(), pd_), C.cast(self.phys.data_ptr(), pd_),
                                         C.cast(self.flag.data_ptr(), pi_),
                                         C.cast(self.stats.data_ptr(), C.POINTER(UclgpuStats)), C.byref(opts))
            lib._check(rc)
            return lib.last_kernel_ms(local_rank)

        def h2d_bytes(self):
            return self.params.numel() * 8 + (0 if self.y0_index is None else self.y0_index.numel() * 4)

        def d2h_bytes(self):
            return self.y.numel() * 8 + self.phys.numel() * 8 + self.flag.numel() * 4 + self.stats.numel() * 8

    class Step:
        """One step: an optional first stage, then the counted models (which start from the first stage's table)."""

        def __init__(self, spec):
            self.spec = spec
            self.s1 = Batch(spec.stage1[0], spec.stage1[1]) if spec.stage1 is not None else None
            self.main = Batch(spec.kind, spec.params, spec.y0_index)
            self.n = self.main.n

        def run(self, budget, cost_hint=None):
            ms = nl = 0.0
            if self.s1 is not None:
                m1, n1 = self.s1.run(budget)
                ms, nl = ms + m1, nl + n1
            m2, n2 = self.main.run(budget, None if self.s1 is None else self.s1.y, cost_hint)
            return ms + m2, int(nl + n2)

        def all_stats(self):
            parts = [self.main.stats.numpy().copy()] + ([self.s1.stats.numpy().copy()] if self.s1 is not None else [])
            return np.concatenate(parts)

        def h2d_bytes(self):
            return self.main.h2d_bytes() + (0 if self.s1 is None else self.s1.h2d_bytes() + self.s1.y.numel() * 8)

        def d2h_bytes(self):
            return self.main.d2h_bytes() + (0 if self.s1 is None else self.s1.d2h_bytes())

    n_distinct = min(a.steps, wl.nslice if not wl.disjoint_ranks else a.steps)
    steps = {}
    for k in range(a.steps):
        sid = step_id(k) % wl.nslice
        if sid not in steps:
            steps[sid] = Step(wl.step(sid))
    first = steps[step_id(0) % wl.nslice]
    wsel = np.linspace(0, first.n - 1, min(first.n, 296)).astype(int)
    wspec = first.spec
    warm = Step(StepSpec(wspec.kind, wspec.params[:, wsel], wspec.stage1, None if wspec.y0_index is None else wspec.y0_index[wsel]))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n_max = max(st_.n for st_ in steps.values())
    stage = torch.empty((n_max, neq + 1), dtype=torch.float64, device=dev)       # y_final + flag, for the gather
    gathered = [torch.empty_like(stage) for _ in range(world)] if (world > 1 and rank == 0) else None
    h_gathered = torch.empty((world, n_max, neq + 1), dtype=torch.float64).pin_memory() if (world > 1 and rank == 0) else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        warm.run(20000)
        if world > 1:   # the communicator is created lazily by the first collective: not inside the timed region
            gather_rows(stage, rank, world, out=gathered)
        flush.fill_(1)
    log(f"warm-up done ({a.warmup} x {warm.n} cells); timing {a.steps} steps of {first.n} cells")

    kernel_ms, launches = 0.0, 0
    n_ok = n_budget = n_cells = 0
    step_stats = []
    attempts = np.full(getattr(wl, "params", np.zeros((1, 0))).shape[1], np.nan)   # per grid cell, for the cost hint
    i_att = [STAT_FIELDS.index(f) for f in ("nst", "netf", "ncfn")]
    hinted_steps = 0
    with ClockSampler(local_rank) as clk:
        barrier()
        t0 = time.perf_counter()
        for k in range(a.steps):
            b = steps[step_id(k) % wl.nslice]
            if step_id(k) % wl.nslice == 0:
                attempts[:] = np.nan    # a new pass over the grid: nothing from the previous one is used
            hint = wl.cost_hint(step_id(k), attempts) if not a.no_cost_hint else None
            hinted_steps += hint is not None
            ms, nl = b.run(a.step_budget, hint)
            if os.environ.get("UCLCHEM_BENCH_VERBOSE"):
                log(f"rank {rank} step {k}: kernel {ms / 1e3:.2f} s, {'hinted' if hint is not None else 'no hint'}")
            if len(attempts):
                attempts[wl.slices[step_id(k) % NSLICE]] = b.main.stats.numpy()[:, i_att].sum(axis=1)
            kernel_ms += ms
            launches += nl
            if world > 1:
                # the only collective of the path: every model's final abundances and flag go to rank 0
                stage[: b.n, :neq].copy_(b.main.y, non_blocking=True)
                stage[: b.n, neq].copy_(b.main.flag.to(torch.float64), non_blocking=True)
                gather_rows(stage, rank, world, out=gathered)
                if rank == 0:
                    for r_ in range(world):
                        h_gathered[r_].copy_(gathered[r_], non_blocking=True)
                    torch.cuda.synchronize()
            fl = b.main.flag.numpy()
            n_ok += int((fl == 0).sum())
            n_budget += int((fl == -5).sum())
            n_cells += b.n
            step_stats.append(b.all_stats())
            flush.fill_(1)  # L2 flush between timed iterations
        barrier()
        wall = time.perf_counter() - t0
    clocks = clk.summary()
    cnt = torch.tensor([n_ok, n_budget, n_cells, launches], dtype=torch.float64, device=dev)
    per_rank = [kernel_ms / a.steps]
    if world > 1:
        mine = torch.tensor([kernel_ms / a.steps], dtype=torch.float64, device=dev)
        allk = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allk, mine)
        per_rank = [float(x.item()) for x in allk]
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    wall, kernel_ms = max_over_ranks([wall, kernel_ms], world, dev)
    n_ok, n_budget, n_cells, launches = (int(x) for x in cnt.tolist())
    log(f"rank {rank}: {a.steps} steps in {wall:.1f} s, kernel {kernel_ms / 1e3:.1f} s")

    if rank == 0:
        log(f"value {n_ok / (kernel_ms / 1e3):.2f} {UNIT}, e2e {n_ok / wall:.2f} {UNIT}; {n_budget} of {n_cells} cells hit the step budget")
        flop = (C.c_double * 8)()
        lib.lib.uclgpu_work_model(flop)
        pk = C.c_double(0.0)
        lib.lib.uclgpu_fp64_peak(local_rank, C.byref(pk))
        traffic = None   # DRAM bytes of one k_integrate launch of the default workload, from the committed ncu launch list
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists() and not a.cells and a.workload == 2:
            traffic = json.loads(tf.read_text()).get("k_integrate_dram_bytes_per_step_launch")
        cores = cpu_cores()
        cpu, parity = None, {"gpu_flags_nonzero": int(n_cells - n_ok)}
        if world == 1:
            # cpu_baseline on cells the timed steps processed (the first step's); the finished ones double as the
            # parity sample.  Staged workloads: the sampled cells start from the GPU's own first-stage results.
            b = first
            log(f"cpu_baseline: oracle work queue on {cores} cores, bounded at {a.cpu_seconds:.0f} s")
            try:
                y0 = None if b.s1 is None else b.s1.y.numpy()[b.spec.y0_index]
                r = run_oracle_sample(b.spec.params, cores, a.cpu_seconds, kind=b.spec.kind, y0=y0, tag=wl.tag)
                cpu = cpu_baseline_dict(r, cores, a.cpu_seconds)
                cells = r["cells"]
                f_gpu, y_gpu = b.main.flag.numpy(), b.main.y.numpy()
                ok = r["finished"] & (r["flag"] == 0) & (f_gpu[cells] == 0)
                yr, yg = r["y"][ok][:, :335], y_gpu[cells[ok]][:, :335]
                m = yr > 1e-15
                dex = np.abs(np.log10(np.where(m, yg / np.where(m, yr, 1.0), 1.0)))
                if ok.any():
                    # the worst cell with what each arm did on it: a cell on which either arm had failed DVODE calls
                    # re-evaluated its interval-frozen rates mid-interval, so the two arms are
                    # no longer integrating the same piecewise problem there
                    kk = np.where(ok)[0][int(dex.max(axis=1).argmax())]
                    gs = dict(zip(STAT_FIELDS, b.main.stats.numpy()[cells[kk]].tolist()))
                    pidx = {k_: i_ for i_, k_ in enumerate(("initialtemp", "initialdens"))}
                    parity["worst_cell"] = {"dex": float(dex.max()), "cell_in_step": int(cells[kk]),
                                            "initialTemp": float(b.spec.params[0, cells[kk]]), "initialDens": float(b.spec.params[1, cells[kk]]),
                                            "gpu_steps": gs["nst"], "gpu_failed_dvode_calls": gs["nfailcall"], "gpu_error_test_failures": gs["netf"],
                                            "oracle_steps": int(r["stats"][kk, 0]), "oracle_convergence_failures": int(r["stats"][kk, 5])}
                    nf = b.main.stats.numpy()[cells[ok]][:, STAT_FIELDS.index("nfailcall")]
                    clean = nf == 0
                    parity["max_dex_on_cells_without_failed_gpu_calls"] = float(dex[clean].max()) if clean.any() else None
                    parity["cells_with_failed_gpu_calls_in_sample"] = int((~clean).sum())
                parity.update({"max_dex_vs_oracle_on_sample": float(dex.max()) if ok.any() else None,
                               "cells_above_0.01_dex": int((dex.max(axis=1) > 0.01).sum()) if ok.any() else None,
                               "sample_cells_compared": int(ok.sum()),
                               "sample": "every cell of the cpu_baseline sample that finished in both arms with flag 0",
                               "oracle_flags_nonzero": int(((r["flag"] != 0) & r["finished"]).sum())})
                log(f"cpu_baseline: {r['wall']:.1f} s, {r['rate']:.3f} models/s on {cores} cores; parity on {int(ok.sum())} cells: "
                    f"{parity['max_dex_vs_oracle_on_sample']}")
            except Exception as e:   # the GPU numbers must not be lost to a CPU-side problem
                cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"failed: {e!r}"}
        line = assemble_line(a=a, world=world, n_ok=n_ok, n_cells=n_cells, n_budget=n_budget, workload=workload,
                             wall_s=wall, kernel_ms=kernel_ms, launches=launches, stats=np.concatenate(step_stats),
                             clocks=clocks, work_model=list(flop), fp64_peak_tflops=pk.value,
                             h2d_bytes=first.h2d_bytes(), d2h_bytes=first.d2h_bytes(), cpu=cpu, parity=parity,
                             traffic=traffic, stat_fields=STAT_FIELDS, per_rank_kernel_ms=per_rank,
                             gather_bytes=(world - 1) * n_max * (neq + 1) * 8 if world > 1 else 0)
        line["config"]["cost_hint"] = (f"{hinted_steps} of {a.steps} steps ran with uclgpu_opts.cost_hint = step attempts measured on the "
                                       "neighbouring grid cells (26-cell index box) earlier in the SAME pass over the grid; the first step of every pass has none, and nothing is carried from one pass to the next"
                                       if hinted_steps else "none (generic longest-first order)")
        if a.workload == 1:
            line["single_model_seconds"] = kernel_ms / 1e3 / a.steps
        emit_json(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
