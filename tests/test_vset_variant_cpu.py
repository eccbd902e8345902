"""The device DVSET (engine_bdf.cuh: recurrences unrolled with predicates, el[] / tau[] in registers)
against the oracle's restatement of dvode.f90 DVSET (orc_vode.c): both are plain scalar code, so the
device function is extracted from the header, compiled for the host with g++ next to the oracle's C
source, and the two are compared bit for bit on random (order, step-history) inputs."""
import shutil
import subprocess

import pytest
from conftest import ROOT

DEVICE_SIDE = r'''
#include <cmath>
#include <cstring>
#define __device__
#define __noinline__
#define LMAXORD 6
#define V_CORTES 0.1
struct Scalars { double tau[14], el[14], tq[6]; double h; int nq, l, nqwait; };
%s
extern "C" void dev_vset(double *tau, double *el, double *tq, double h, int nq, int l, int nqwait)
{
    Scalars s;
    memcpy(s.tau, tau, sizeof(s.tau)); memcpy(s.el, el, sizeof(s.el)); memcpy(s.tq, tq, sizeof(s.tq));
    s.h = h; s.nq = nq; s.l = l; s.nqwait = nqwait;
    vset_dev(s);
    memcpy(el, s.el, sizeof(s.el)); memcpy(tq, s.tq, sizeof(s.tq));
}
'''

ORACLE_SIDE = r'''
#include "%s"
#include <stdio.h>
void dev_vset(double *tau, double *el, double *tq, double h, int nq, int l, int nqwait);
int main(void)
{
    srand(12345);
    long bad = 0, n = 0;
    for (int trial = 0; trial < 100000; trial++) {
        vode_t a;
        memset(&a, 0, sizeof(a));
        a.nq = 1 + rand() %% 5; a.l = a.nq + 1; a.nqwait = rand() %% 4;
        a.h = pow(10.0, -3 + 12.0 * rand() / RAND_MAX);
        for (int i = 0; i < 14; i++) { a.tau[i] = a.h * pow(10.0, -1.0 + 2.0 * rand() / RAND_MAX); a.el[i] = 1e300 * (rand() %% 3); }
        for (int i = 0; i < 6; i++) a.tq[i] = -7.0;
        double el[14], tq[6];
        memcpy(el, a.el, sizeof(el)); memcpy(tq, a.tq, sizeof(tq));
        dev_vset(a.tau, el, tq, a.h, a.nq, a.l, a.nqwait);
        vset(&a);
        n++;
        /* entries of EL above L are scratch in both versions */
        if (memcmp(a.tq, tq, sizeof(tq)) != 0 || memcmp(a.el + 1, el + 1, sizeof(double) * a.l) != 0) bad++;
    }
    printf("%%ld %%ld\n", n, bad);
    return 0;
}
'''


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_device_dvset_is_bitwise_the_oracle_dvset(tmp_path):
    src = (ROOT / "uclchem_b200" / "csrc" / "engine_bdf.cuh").read_text()
    fn = src[src.index("__device__ __noinline__ void vset_dev"):]
    fn = fn[: fn.index("\n}\n") + 3]
    (tmp_path / "dev.cpp").write_text(DEVICE_SIDE % fn)
    (tmp_path / "orc.c").write_text(ORACLE_SIDE % (ROOT / "oracle" / "orc_vode.c"))
    flags = ["-O2", "-ffp-contract=off"]
    subprocess.run(["g++", *flags, "-c", "-o", str(tmp_path / "dev.o"), str(tmp_path / "dev.cpp")], check=True)
    subprocess.run(["gcc", *flags, "-std=gnu11", "-I", str(ROOT / "oracle"), "-c", "-o", str(tmp_path / "orc.o"),
                    str(tmp_path / "orc.c")], check=True)
    subprocess.run(["g++", "-o", str(tmp_path / "t"), str(tmp_path / "dev.o"), str(tmp_path / "orc.o"), "-lm", "-lpthread"],
                   check=True)
    n, bad = map(int, subprocess.run([str(tmp_path / "t")], check=True, capture_output=True, text=True).stdout.split())
    assert n == 100000 and bad == 0
