"""Opt-in blended three-phase transfer (uclgpu_opts.transfer_band / UCLGPU_TRANSFER_BAND): effect on the stall
cells and on ordinary cells (run on the GPU box).
usage: gpu_band_probe.py <budget> <band> [<band> ...]"""
import os, sys, time, functools
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
print = functools.partial(print, flush=True)
from bench import config2_params
from uclchem_b200._capi import Library, STAT_FIELDS
budget = int(float(sys.argv[1])); bands = [float(x) for x in sys.argv[2:]]
lib = Library("default"); lib.init([0])
P = config2_params()
heavy = np.load(ROOT / "tools" / "config2_stall_cells_r02.npy")   # cells beyond 30 000 steps in the round-2 default build
easy = np.setdiff1d(np.linspace(0, 9999, 600).astype(int), heavy)
res = {}
for band in bands:
    os.environ["UCLGPU_TRANSFER_BAND"] = repr(band)
    for name, idx in (("heavy", heavy), ("easy", easy)):
        p = np.ascontiguousarray(P[:, idx])
        t = time.time(); o = lib.run_grid(0, p, step_budget=budget); dt = time.time() - t
        S = {k: o["stats"][:, i] for i, k in enumerate(STAT_FIELDS)}
        print(f"band {band:g} {name} ({len(idx)} cells): {dt:.1f} s flags {dict(zip(*np.unique(o['flag'], return_counts=True)))} "
              f"nst pct 50/90/max {np.percentile(S['nst'], [50, 90, 100]).astype(int)} netf/nst {S['netf'].sum() / S['nst'].sum():.3f} "
              f"ncfn {S['ncfn'].sum()} failcalls {S['nfailcall'].sum()}")
        res[(band, name)] = o
    np.savez_compressed(ROOT / "gpurun_out" / f"band_{band:g}.npz", heavy=heavy, easy=easy,
                        y_heavy=res[(band, "heavy")]["y_final"], flag_heavy=res[(band, "heavy")]["flag"], stats_heavy=res[(band, "heavy")]["stats"],
                        y_easy=res[(band, "easy")]["y_final"], flag_easy=res[(band, "easy")]["flag"], stats_easy=res[(band, "easy")]["stats"])
b0 = bands[0]
for band in bands[1:]:
    for name in ("heavy", "easy"):
        a, b = res[(b0, name)], res[(band, name)]
        ok = (a["flag"] == 0) & (b["flag"] == 0)
        ya, yb = a["y_final"][ok][:, :335], b["y_final"][ok][:, :335]
        m = ya > 1e-15
        dex = np.abs(np.log10(np.where(m, yb / ya, 1.0)))
        print(f"band {band:g} vs {b0:g} on {name}: {ok.sum()} cells finished in both, max dex {dex.max():.3e}, cells > 0.01 dex: {(dex.max(axis=1) > 0.01).sum()}, 99th pct of per-cell max {np.percentile(dex.max(axis=1), 99):.2e}")
