"""Golden fixture for the 1 Myr config-2 parity test: the CPU oracle (oracle/, the restatement of the
reference's DVODE path) run TO COMPLETION on a fixed list of cells of the 10^4-point static-cloud grid.

The list mixes cells evenly spaced over the grid with cells from the regions where the integration is
hard: 8679 and 7245 (the oracle needs 4-5x the engine's steps there) and cells of
tools/config2_heavy_cells.npy (the round-1 engine exceeded 30 000 steps on them).  Each cell is one
process (no deadline); results are written after every finished cell, so the script can be stopped and
restarted (finished cells are kept).

usage: python tools/make_config2_fixture.py [nproc]      -> tests/golden/config2_1myr_cells.npz
"""
import sys
import time
from multiprocessing import Pool
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
OUT = ROOT / "tests" / "golden" / "config2_1myr_cells.npz"

EASY = [137, 1422, 2707, 3992, 5277, 6562, 7847, 9132]
HARD = [8679, 7245, 8539, 6179, 4146, 5721, 9877, 8361]
CELLS = EASY + HARD


def run(cell):
    from bench import config2_params
    from oracle.oracle import Oracle
    from uclchem_b200.network import load_default
    P = config2_params()
    o = Oracle(load_default())
    t = time.time()
    r = o.run_model(0, P[:, cell].copy())
    return cell, r["flag"], r["y_final"], np.array(list(r["stats"].values()), np.int64), time.time() - t


def save(done):
    cells = sorted(done)
    np.savez_compressed(OUT, cells=np.array(cells), flag=np.array([done[c][0] for c in cells], np.int32),
                        y_final=np.array([done[c][1] for c in cells]), stats=np.array([done[c][2] for c in cells]),
                        seconds=np.array([done[c][3] for c in cells]),
                        stat_fields=np.array(["nst", "nfe", "nje", "nlu", "nni", "ncfn", "netf", "nintervals"]))


if __name__ == "__main__":
    nproc = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    done = {}
    if OUT.exists():
        d = np.load(OUT)
        done = {int(c): (int(f), y, s, float(t)) for c, f, y, s, t in zip(d["cells"], d["flag"], d["y_final"], d["stats"], d["seconds"])}
    todo = [c for c in CELLS if c not in done]
    print(f"{len(done)} cells already done, {len(todo)} to run on {nproc} processes", flush=True)
    with Pool(nproc) as pool:
        for cell, flag, y, st, dt in pool.imap_unordered(run, todo):
            done[cell] = (flag, y, st, dt)
            save(done)
            print(f"cell {cell}: flag {flag} nst {st[0]} netf {st[6]} ncfn {st[5]} {dt:.0f} s", flush=True)
