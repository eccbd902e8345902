"""The blocked Gauss-Jordan inverse of the dense trailing block (uclchem_b200/csrc/engine_gj.cuh) is written as
pure per-thread phases, so the header is compiled for the host here and the phases are run thread by thread in
the order the device barriers impose (publish | scale | update + publish), against numpy's inverse."""
import shutil
import subprocess

import numpy as np
import pytest
from conftest import ROOT

HARNESS = r'''
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
using std::isfinite;
#define GJ_HOST_TEST
#define MDENSE %(m)d
#define GJ_B 5
%(macros)s
#include "%(hdr)s"
int main()
{
    const int M = MDENSE, NTHR = (GJ_NT * GJ_NT + 31) & ~31;
    std::vector<double> T(M * M), pan(GJ_PANEL, 1e300);
    if (fread(T.data(), sizeof(double), M * M, stdin) != (size_t)(M * M)) return 2;
    std::vector<GjTile> t(NTHR);
    for (int tid = 0; tid < NTHR; tid++) { gj_load(t[tid], T.data(), tid); gj_init_panel(pan.data(), tid); }
    for (int tid = 0; tid < NTHR; tid++) gj_publish(t[tid], pan.data(), tid, 0);
    for (int kb = 0; kb < GJ_NT; kb++) {
        for (int tid = NTHR - 1; tid >= 0; tid--) gj_scale_row_panel(pan.data(), tid, NTHR, kb);
        // update and look-ahead publish are not separated by a barrier on the device: run them per thread, in
        // an order that exposes a missing double buffer (publishers of step kb+1 first)
        for (int pass = 0; pass < 2; pass++)
            for (int tid = 0; tid < NTHR; tid++) {
                const int tr = tid / GJ_NT, tc = tid %% GJ_NT;
                const bool pub = tr == kb + 1 || tc == kb + 1;
                if (pub != (pass == 0)) continue;
                gj_update(t[tid], pan.data(), tid, kb);
                gj_publish(t[tid], pan.data(), tid, kb + 1);
            }
    }
    for (int tid = 0; tid < NTHR; tid++) gj_store(t[tid], T.data(), tid);
    fwrite(T.data(), sizeof(double), M * M, stdout);
    fprintf(stderr, "%%d\n", pan[GJ_OK] != 0.0);
    return 0;
}
'''


def _macros():
    src = (ROOT / "uclchem_b200" / "csrc" / "engine_core.cuh").read_text()
    keep = [l for l in src.splitlines() if l.startswith("#define GJ_") and not l.startswith("#define GJ_B ")]
    assert any("GJ_PANEL" in l for l in keep)
    return "\n".join(keep)


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
@pytest.mark.parametrize("m", [91, 89, 5, 12])
def test_blocked_gauss_jordan_matches_dense_inverse(tmp_path, m):
    (tmp_path / "t.cpp").write_text(HARNESS % {"m": m, "macros": _macros(),
                                               "hdr": ROOT / "uclchem_b200" / "csrc" / "engine_gj.cuh"})
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-o", str(tmp_path / "t"), str(tmp_path / "t.cpp")], check=True)
    rng = np.random.default_rng(m)
    A = rng.standard_normal((m, m)) * 0.05 + np.eye(m) * (1.0 + rng.random(m))   # diagonally dominated, like I - gamma J
    r = subprocess.run([str(tmp_path / "t")], input=A.tobytes(), capture_output=True, check=True)
    inv = np.frombuffer(r.stdout, np.float64).reshape(m, m)
    assert r.stderr.decode().strip() == "1"
    assert np.abs(inv @ A - np.eye(m)).max() < 1e-13
    assert np.abs(inv - np.linalg.inv(A)).max() < 1e-12
    # a zero pivot is reported
    A[0, 0] = 0.0
    r = subprocess.run([str(tmp_path / "t")], input=A.tobytes(), capture_output=True, check=True)
    assert r.stderr.decode().strip() == "0"
