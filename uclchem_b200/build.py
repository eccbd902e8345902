"""Build the CUDA library for one MakeRates network (in-tree, sm_100a only).

``generate(tag)`` runs the MakeRates CUDA back-end on the network (from the
reference's ``network.f90`` when a path is given, else from the committed
``networks/<tag>.json``); ``compile(tag)`` runs nvcc.  The result is
``uclchem_b200/lib/libuclgpu_<tag>.so`` -- one shared library per network, the way
the reference compiles one ``uclchemwrap`` per network.
"""
from __future__ import annotations

import shutil
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
CSRC = _PKG / "csrc"
LIBDIR = _PKG / "lib"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


# per-network build options: dense-tail threshold of the symbolic LU and extra defines.  The crp-photo
# network (config 5) needs the compact shared-memory layout and a slightly smaller dense block to fit
# the 227 KB of an SM.
TAG_OPTIONS = {
    "crp_photo": {"dense_threshold": 0.95, "defines": ["-DUCLGPU_COMPACT_SMEM"]},
    # default network + grain-assisted recombination (tools/make_gar_network.py): default layout and options
    "gar": {},
}


def generate(tag: str = "default", network_f90: str | None = None) -> Path:
    from .makerates_cuda import Generated, emit
    from .network import Network

    js = _PKG / "networks" / f"{TAG_OPTIONS.get(tag, {}).get('network', tag)}.json"
    if network_f90 is not None:
        net = Network.from_network_f90(network_f90)
        js.parent.mkdir(parents=True, exist_ok=True)
        net.to_json(js)
    else:
        net = Network.from_json(js)
    opt = TAG_OPTIONS.get(tag, {})
    gen = Generated(net, dense_threshold=opt.get("dense_threshold", 0.9), factor_terms_per_lane=opt.get("factor_terms_per_lane", 8))
    return emit(gen, CSRC / "generated" / tag, tag)


# build variants: suffix of the library name -> extra nvcc defines.  The default build uses the product-form
# solves with the inverse program overlapped with the dense Gauss-Jordan inverse (engine_core.cuh); "sub" keeps
# the level-scheduled substitution for A/B runs (`python tools/gpu_ab.py default_sub default 592`).
VARIANTS = {
    "": [],
    "sub": ["-DUCLGPU_NO_PRODUCT_FORM"],
    # inverse program of the product form on the whole CTA after the dense inverse instead of next to it
    "nov": ["-DUCLGPU_NO_PF_OVERLAP"],
}


def compile(tag: str = "default", force: bool = False, verbose: bool = False, variant: str = "") -> Path:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    gen = CSRC / "generated" / tag / "net_tables.cuh"
    generator = [_PKG / f for f in ("makerates_cuda.py", "symbolic.py", "product_form.py", "network.py")]
    if not gen.exists() or any(g.stat().st_mtime > gen.stat().st_mtime for g in generator):
        generate(tag)   # tables are a pure function of the network and the generator
    LIBDIR.mkdir(exist_ok=True)
    out = LIBDIR / (f"libuclgpu_{tag}_{variant}.so" if variant else f"libuclgpu_{tag}.so")
    srcs = [CSRC / "uclgpu.cu", CSRC / "engine_core.cuh", CSRC / "engine_la.cuh", CSRC / "engine_gj.cuh", CSRC / "engine_bdf.cuh", CSRC / "engine_collapse.cuh",
            CSRC / "engine_model.cuh", gen, _PKG.parent / "include" / "uclgpu.h"]
    if not force and out.exists() and all(out.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return out
    cmd = [nvcc, *NVCC_FLAGS, *VARIANTS[variant], *TAG_OPTIONS.get(tag, {}).get("defines", []), f"-I{gen.parent}",
           "-o", str(out), str(CSRC / "uclgpu.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return out
