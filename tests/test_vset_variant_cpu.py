"""The DVSET-in-registers build variant (-DUCLGPU_VSET_REG, engine_bdf.cuh) against the shared-memory
version it replaces: both are plain scalar C++, so they are extracted from the header, compiled for
the host with g++ and compared bit for bit on random (order, step-history) inputs."""
import shutil
import subprocess

import pytest
from conftest import ROOT

HARNESS = r'''
#include <cstdio>
#include <cstring>
#include <cmath>
#include <cstdlib>
#define __device__
#define __noinline__
#define LMAXORD 6
#define V_CORTES 0.1
struct Scalars { double tau[14], el[14], tq[6]; double h; int nq, l, nqwait; };
%s
int main() {
    srand(12345);
    long bad = 0, n = 0;
    for (int trial = 0; trial < 100000; trial++) {
        Scalars a;
        memset(&a, 0, sizeof(a));
        a.nq = 1 + rand() %% 5; a.l = a.nq + 1; a.nqwait = rand() %% 4;
        a.h = pow(10.0, -3 + 12.0 * rand() / RAND_MAX);
        for (int i = 0; i < 14; i++) { a.tau[i] = a.h * pow(10.0, -1.0 + 2.0 * rand() / RAND_MAX); a.el[i] = 1e300 * (rand() %% 3); }
        for (int i = 0; i < 6; i++) a.tq[i] = -7.0;
        Scalars b = a;
        vset_ref(a); vset_reg(b);
        n++;
        if (memcmp(a.tq, b.tq, sizeof(a.tq)) != 0 || memcmp(a.el, b.el, sizeof(a.el)) != 0) bad++;
    }
    printf("%%ld %%ld\n", n, bad);
    return 0;
}
'''


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_register_dvset_is_bitwise_the_shared_memory_dvset(tmp_path):
    src = (ROOT / "uclchem_b200" / "csrc" / "engine_bdf.cuh").read_text()
    blk = src[src.index("#ifdef UCLGPU_VSET_REG"): src.index("#endif // UCLGPU_VSET_REG")]
    reg = blk[blk.index("__device__ __noinline__ void vset_dev"): blk.index("#else")].replace("vset_dev", "vset_reg")
    ref = blk[blk.index("#else"):]
    ref = ref[ref.index("__device__ __noinline__ void vset_dev"):].replace("vset_dev", "vset_ref")
    (tmp_path / "t.cpp").write_text(HARNESS % (ref + reg))
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-o", str(tmp_path / "t"), str(tmp_path / "t.cpp")], check=True)
    n, bad = map(int, subprocess.run([str(tmp_path / "t")], check=True, capture_output=True, text=True).stdout.split())
    assert n == 100000 and bad == 0
