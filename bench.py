#!/usr/bin/env python
"""bench.py -- cloud models integrated per second on B200 (BASELINE.json metric).

A *step* is one pass of the hot path over one batch: the BASELINE config-2 workload, the
10^4-point static-cloud grid (25 densities x 20 temperatures x 20 cosmic-ray rates, 1 Myr
each, default network).  161 of its cells (1.6 %) are cells on which the reference algorithm
itself does not terminate in bounded work: DVODE hits MXSTEP in every retry, and because
`usepostprocess` is always .true. UCLCHEM's own stall guard (chemistry.f90:224-237) never
fires, so the model crawls on for 1e5..1e7 steps (measured here: the slowest single cell needs
~8e6 BDF steps = 573 s on one SM, hours on a CPU core, while the other 9 839 cells together need
~40 s on 148 SMs).  A cell is one sequential chain of steps, so those few cells would set the
pass time of ANY implementation and make the default run take tens of minutes.  Both arms
therefore time the same bounded workload: the grid minus those 161 cells (indices committed in
tools/config2_heavy_cells.npy).  `--full-grid` times all 10^4 cells (DESIGN.md section 6 has
that measurement).  Per rank the work is fixed (weak scaling): with N ranks the job integrates
N x 9 839 independent models and rank 0 gathers the results.

  value  : device-resident leg (parameters already in HBM, results left in HBM)
  e2e    : the call a user makes -- uclgpu_run_grid through the C ABI with pinned HOST
           buffers, H2D and D2H inside the timed region
  roofline / fp64 : algorithmic work from the solver counters x the per-operation counts
           the MakeRates CUDA back-end emits, over the CUDA-event time of the kernel
  cpu_baseline : the oracle (CPU restatement of the reference algorithm: dense FD
           Jacobian DVODE, cold restart per interval) on the box's host cores

`--impl reference` times that CPU restatement alone (the reference Fortran cannot be
compiled in this image: no Fortran compiler).
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


def config2_params(ncell_side=(25, 20, 20)):
    """regular grid, no RNG."""
    from uclchem_b200.params import params_from_dict
    nd, nt, nz = ncell_side
    dens = 10 ** np.linspace(3, 7, nd)
    temp = np.linspace(10, 100, nt)
    zeta = 10 ** np.linspace(0, 3, nz)
    D, T, Z = np.meshgrid(dens, temp, zeta, indexing="ij")
    return params_from_dict({"initialDens": D.ravel(), "initialTemp": T.ravel(), "zeta": Z.ravel(), "radfield": 1.0,
                             "baseAv": 2.0, "rout": 0.05, "finalTime": 1.0e6, "freefall": False,
                             "endAtFinalDensity": False})


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


def bounded_cells(ncell):
    """Indices of the config-2 cells with bounded work (see the module docstring)."""
    ok = np.ones(ncell, bool)
    f = ROOT / "tools" / "config2_heavy_cells.npy"
    if ncell == 10000:
        ok[np.load(f)] = False
    return np.where(ok)[0]


def sample_cells(ncell, n, offset=0):
    """`n` evenly spaced cells of the workload (deterministic; `offset` shifts the comb)."""
    return (np.linspace(0, ncell - 1, n).astype(int) + offset) % ncell


def run_oracle_sample(params, cores, offset=0, deadline_s=90.0):
    """Time the CPU restatement on `cores` evenly spaced cells of the workload, one per core, with a
    wall-clock bound: a model still running after `deadline_s` is stopped (oracle guard, flag -98) and
    does not count.  Returns (models/s over the cells that finished, seconds, cell indices, y, flag)."""
    from oracle.oracle import Oracle
    from uclchem_b200.network import load_default
    idx = sample_cells(params.shape[1], cores, offset)
    orc = Oracle(load_default())
    orc.set_deadline(deadline_s)
    t0 = time.perf_counter()
    y, _, flag, _ = orc.run_grid(0, np.ascontiguousarray(params[:, idx]), nthreads=cores)
    dt = time.perf_counter() - t0
    orc.set_deadline(0.0)
    done = int((flag != Oracle.FLAG_DEADLINE).sum())
    return done / dt, dt, idx, y, flag


def sample_text(idx, flag, dt, deadline_s):
    cut = int((flag == -98).sum())
    txt = f"{len(idx)} evenly spaced cells of the workload, one per host core, {dt:.1f} s"
    if cut:
        txt += (f"; {cut} of them were still running at the {deadline_s:.0f} s bound and are not counted "
                "(the figure is then an upper bound of the CPU rate)")
    return txt


def assemble_line(*, a, world, ncell, workload, dt, kernel_ms, dt_e2e, launches, launches_e2e, stats, flags,
                  clocks, work_model, fp64_peak_tflops, h2d_bytes, d2h_bytes, cpu, parity, traffic, stat_fields):
    """The bench JSON line from measured quantities (pure function: unit-tested on the CPU)."""
    f_rhs, f_jac, f_lu, f_solve, f_rates, b_interval = list(work_model)[:6]
    S = {k: stats[:, i].astype(np.float64).sum() for i, k in enumerate(stat_fields)}
    w_flop = (S["nfe"] * f_rhs + S["nje"] * f_jac + S["nlu"] * f_lu + S["nni"] * f_solve + S["nintervals"] * f_rates)
    w_bytes = S["nintervals"] * b_interval
    kern_s = kernel_ms / 1e3 / a.steps          # one launch per step
    peak, how = peak_hbm()
    roof = {"bound": "hbm", "achieved": w_bytes / kern_s / 1e9, "peak": peak, "unit": "GB/s",
            "frac": w_bytes / kern_s / 1e9 / peak, "traffic": traffic, "peak_source": how,
            "note": "state is chip-resident per cell: HBM traffic is only cell load/store, the kernel is bound "
                    "by dependent-instruction latency (shared memory, barriers) and the fp64 pipe, see fp64"}
    fp64 = {"achieved_tflops": w_flop / kern_s / 1e12, "peak_tflops": fp64_peak_tflops,
            "frac": w_flop / kern_s / 1e12 / fp64_peak_tflops if fp64_peak_tflops else None,
            "peak_source": "DFMA micro-benchmark run by this process (uclgpu_fp64_peak)"}
    return {
        "metric": METRIC, "value": world * ncell * a.steps / dt, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload, "clocks": clocks,
        "e2e": {"value": world * ncell * a.steps / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": int(d2h_bytes)},
        "gpu_launches": int(launches + launches_e2e),
        "kernel_ms_per_step": kernel_ms / a.steps,   # CUDA events around k_integrate on its stream (max over ranks)
        "roofline": roof, "fp64": fp64, "cpu_baseline": cpu, "parity": parity,
        "solver": {"steps_per_model": S["nst"] / ncell, "lu_per_model": S["nlu"] / ncell,
                   "jac_per_model": S["nje"] / ncell, "newton_iters_per_model": S["nni"] / ncell,
                   "failed_dvode_calls": S["nfailcall"]},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=0, help="debug: override the grid size (not a valid bench line)")
    ap.add_argument("--full-grid", action="store_true",
                    help="time all 10^4 cells, including the 161 on which the reference algorithm stalls (~10 min per pass)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    params = config2_params()
    n_grid = params.shape[1]
    if not a.full_grid:
        params = np.ascontiguousarray(params[:, bounded_cells(n_grid)])
    if a.cells:
        params = np.ascontiguousarray(params[:, np.linspace(0, params.shape[1] - 1, a.cells).astype(int)])
    ncell = params.shape[1]
    # warm-up steps run the same kernel over a bounded stride of the same grid (clocks, instruction
    # cache, lazy module load); a full pass is tens of seconds and has no state a warm-up could prime
    n_warm = min(ncell, 592)
    warm_idx = np.linspace(0, ncell - 1, n_warm).astype(int)
    excluded = n_grid - params.shape[1] if not a.cells else None
    workload = {"workload": f"config[1]: {n_grid}-point static cloud grid (25 n_H x 20 T x 20 zeta), 1 Myr, default "
                            "network 335 species / 3203 reactions, reltol 1e-8"
                            + ("" if a.full_grid else f"; both arms time the {ncell} cells with bounded work: the "
                               f"{excluded} cells (1.6 %) on which the reference algorithm itself stalls (MXSTEP in "
                               "every DVODE retry, 1e5-1e7 steps, up to 573 s for ONE cell on one SM) are left out, "
                               "see bench.py docstring / DESIGN.md section 6"),
                "cells_per_gpu": ncell,
                "full_grid_measured": "all 10000 cells, one device-resident pass on one B200: 573.2 s = 17.4 models/s "
                                      "(round 1, profiles/r01b_bench_full_grid_stderr.log); `--full-grid` re-measures it",
                "timing": "L2 flushed (256 MiB write) between timed steps",
                "warmup_step": f"one pass over a {n_warm}-cell stride of the same grid (same kernel and launch shape)"}

    # ------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return
        cores = cpu_cores()
        deadline = max(30.0, min(120.0, 240.0 / max(1, a.steps)))   # the whole run stays within a few minutes
        for w in range(a.warmup):   # results discarded: a short bound is enough to page the library in
            log(f"reference arm: warm-up sample {w + 1}/{a.warmup} on {cores} cores")
            run_oracle_sample(params, cores, offset=1 + w, deadline_s=30.0)
        n_done, dt, cut = 0, 0.0, 0
        for k in range(a.steps):
            v_k, dt_k, idx, _, flag = run_oracle_sample(params, cores, offset=100 + k, deadline_s=deadline)
            n_done += int((flag != -98).sum())
            cut += int((flag == -98).sum())
            dt += dt_k
            log(f"reference arm: step {k + 1}/{a.steps}: {dt_k:.1f} s, {int((flag != -98).sum())} of {len(idx)} cells finished")
        v = n_done / dt
        sample = (f"{cores} evenly spaced cells of the workload per step, one per host core, bounded at {deadline:.0f} s "
                  f"per step ({cut} cells cut off and not counted)")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    from uclchem_b200._capi import STAT_FIELDS, UclgpuStats, get_library
    from uclchem_b200.sharding import gather_results

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = get_library("default")          # raises if the CUDA extension is missing: no fallback
    lib.init([local_rank])
    neq = lib.neq
    nstat = len(STAT_FIELDS)

    # device-resident buffers (value leg)
    d_params = torch.from_numpy(params).to(dev)
    d_y = torch.empty((ncell, neq), dtype=torch.float64, device=dev)
    d_phys = torch.empty((ncell, 8), dtype=torch.float64, device=dev)
    d_flag = torch.empty(ncell, dtype=torch.int32, device=dev)
    d_stats = torch.zeros((ncell, nstat), dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # pinned host buffers (e2e leg)
    h_params = torch.from_numpy(params).pin_memory()
    h_y = torch.empty((ncell, neq), dtype=torch.float64).pin_memory()
    h_phys = torch.empty((ncell, 8), dtype=torch.float64).pin_memory()
    h_flag = torch.empty(ncell, dtype=torch.int32).pin_memory()
    h_stats = torch.zeros((ncell, nstat), dtype=torch.int64).pin_memory()
    pd_, pi_ = C.POINTER(C.c_double), C.POINTER(C.c_int32)

    def step_device():
        rc = lib.lib.uclgpu_run_grid_device(local_rank, 0, ncell, d_params.data_ptr(), None, d_y.data_ptr(),
                                            d_phys.data_ptr(), d_flag.data_ptr(), d_stats.data_ptr(), None)
        lib._check(rc)
        return lib.last_kernel_ms(local_rank)

    def step_e2e():
        rc = lib.lib.uclgpu_run_grid(0, ncell, C.cast(h_params.data_ptr(), pd_), None, C.cast(h_y.data_ptr(), pd_),
                                     C.cast(h_phys.data_ptr(), pd_), C.cast(h_flag.data_ptr(), pi_),
                                     C.cast(h_stats.data_ptr(), C.POINTER(UclgpuStats)), None)
        lib._check(rc)
        return lib.last_kernel_ms(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        kernel_ms, launches = 0.0, 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            ms, nl = fn()
            kernel_ms += ms
            launches += nl
            flush.fill_(1)  # L2 flush between timed iterations
        if world > 1:   # the only collective of the path: final result gather on rank 0
            gather_results(d_flag.to(torch.float64).unsqueeze(1), world * ncell, rank, world)
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt, kernel_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item(), launches

    d_wparams = torch.from_numpy(np.ascontiguousarray(params[:, warm_idx])).to(dev)
    for _ in range(a.warmup):
        lib._check(lib.lib.uclgpu_run_grid_device(local_rank, 0, n_warm, d_wparams.data_ptr(), None, d_y.data_ptr(),
                                                  d_phys.data_ptr(), d_flag.data_ptr(), d_stats.data_ptr(), None))
        flush.fill_(1)
    log(f"warm-up done ({a.warmup} x {n_warm} cells); timing {a.steps} device-resident pass(es) over {ncell} cells")
    with ClockSampler(local_rank) as clk:
        dt, kernel_ms, launches = timed(step_device, a.steps)
    log(f"device-resident leg: {dt:.1f} s ({world * ncell * a.steps / dt:.1f} models/s); timing the e2e leg")
    clocks = clk.summary()
    value = world * ncell * a.steps / dt
    stats = d_stats.cpu().numpy()
    flags = d_flag.cpu().numpy()
    dt_e2e, _, launches_e2e = timed(step_e2e, a.steps)
    e2e_value = world * ncell * a.steps / dt_e2e
    log(f"e2e leg: {dt_e2e:.1f} s ({e2e_value:.1f} models/s)")
    assert np.array_equal(h_flag.numpy(), flags)

    if rank == 0:
        # ---- algorithmic work: solver counters x per-operation counts emitted by the generator
        flop = (C.c_double * 8)()
        lib.lib.uclgpu_work_model(flop)
        pk = C.c_double(0.0)
        lib.lib.uclgpu_fp64_peak(local_rank, C.byref(pk))
        traffic = None   # DRAM bytes of one k_integrate launch of this workload, from the committed ncu launch list
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists() and not a.cells and not a.full_grid:
            traffic = json.loads(tf.read_text()).get("k_integrate_dram_bytes_per_launch")
        cores = cpu_cores()
        deadline = 120.0
        log(f"cpu_baseline: oracle on {cores} cores, bounded at {deadline:.0f} s")
        try:
            cv, cdt, idx, yref, cflag = run_oracle_sample(params, cores, deadline_s=deadline)
            ok = cflag == 0
            y_gpu = h_y.numpy()[idx][:, :335]
            m = (yref[:, :335] > 1e-15) & ok[:, None]
            dex = float(np.abs(np.log10(y_gpu[m] / yref[:, :335][m])).max()) if m.any() else None
            cpu = {"value": cv, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": sample_text(idx, cflag, cdt, deadline)}
            parity = {"max_dex_vs_oracle_on_sample": dex, "sample_cells_compared": int(ok.sum()),
                      "flags_nonzero": int((flags != 0).sum()),
                      "oracle_flags_nonzero": int(((cflag != 0) & (cflag != -98)).sum())}
            log(f"cpu_baseline: {cdt:.1f} s, {cv:.3f} models/s")
        except Exception as e:   # the GPU numbers must not be lost to a CPU-side problem
            cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"failed: {e!r}"}
            parity = {"flags_nonzero": int((flags != 0).sum())}
        # the timed numbers first, on stderr, in case anything below goes wrong
        log(f"value {value:.2f} {UNIT}, e2e {e2e_value:.2f} {UNIT}, kernel {kernel_ms / a.steps:.1f} ms/step")
        line = assemble_line(a=a, world=world, ncell=ncell, workload=workload, dt=dt, kernel_ms=kernel_ms,
                             dt_e2e=dt_e2e, launches=launches, launches_e2e=launches_e2e, stats=stats, flags=flags,
                             clocks=clocks, work_model=list(flop), fp64_peak_tflops=pk.value,
                             h2d_bytes=params.nbytes,
                             d2h_bytes=h_y.numel() * 8 + h_phys.numel() * 8 + h_flag.numel() * 4 + h_stats.numel() * 8,
                             cpu=cpu, parity=parity, traffic=traffic, stat_fields=STAT_FIELDS)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
