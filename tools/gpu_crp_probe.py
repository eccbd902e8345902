"""Sanity probe of a library on a few cells (run on the GPU box under an outer `timeout`).
usage: gpu_crp_probe.py <tag> <ncells> <finalTime>"""
import sys, time, functools
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
print = functools.partial(print, flush=True)
from uclchem_b200._capi import Library, STAT_FIELDS
from uclchem_b200.params import params_from_dict
tag, n, ft = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
L = Library(tag); L.init([0])
p = params_from_dict({"initialDens": 10 ** np.linspace(3, 7, n), "initialTemp": np.linspace(10, 100, n), "zeta": 10 ** np.linspace(0, 3, n),
                      "radfield": 1.0, "baseAv": 2.0, "rout": 0.05, "finalTime": ft, "reltol": 1e-5, "abstol_min": 1e-15})
print("launching", tag, n, ft)
t = time.time(); o = L.run_grid(0, p, step_budget=20000); dt = time.time() - t
print(f"[{tag}] {n} cells to {ft:g} yr: {dt:.2f} s flags {dict(zip(*np.unique(o['flag'], return_counts=True)))} nst max {o['stats'][:, 0].max()}")
