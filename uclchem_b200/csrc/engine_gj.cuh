// engine_gj.cuh -- blocked Gauss-Jordan inverse of the dense trailing block (DGEFA/DGESL replacement for
// the last NET_M unknowns, dvode.f90:11982,12091), the per-thread phases.
//
// The M x M block is cut into GJ_NT x GJ_NT tiles of GJ_B x GJ_B; thread tid = tr * GJ_NT + tc keeps tile
// (tr, tc) in registers for the whole inversion (rows / columns beyond M are padded with the identity).
// Block step kb eliminates the GJ_B pivots of diagonal tile kb at once:
//
//     D^-1          = inverse of the diagonal tile (one thread, in registers)
//     R_j           = D^-1 A[kb][j]                   row panel
//     A[i][j]      -= A[i][kb] R_j       (i, j != kb)
//     A[kb][j]      = R_j,   A[i][kb] = -A[i][kb] D^-1,   A[kb][kb] = D^-1
//
// which is the in-place Gauss-Jordan inverse without pivoting, GJ_B scalar steps at a time.  Two barriers per
// block step (2 x 19 instead of one per pivot, 91): the column panel A[:][kb], the unscaled row panel A[kb][:]
// and D^-1 of step kb are published through shared memory at the END of step kb-1 (look-ahead, right after the
// owning threads have updated their tiles), so step kb is
//
//     barrier | scale: R = D^-1 * row panel, spread over all threads | barrier | update + publish for kb+1
//
// The update is branch-free: tiles on the pivot tile row / column start from zero and read "minus identity"
// (row) or D^-1 (column) as their other factor, so every thread runs the same GJ_B^3 multiply-adds.
//
// Everything here is a pure function of (tid, registers, panel) so that tests/test_dense_inverse_cpu.py can
// compile this header for the host and run the phases thread by thread against a dense inverse.
#pragma once

#ifndef GJ_HOST_TEST
#include "engine_core.cuh"
#define GJ_FN __device__ __forceinline__
#else
#define GJ_FN static inline
#endif

struct GjTile {
    double a[GJ_B][GJ_B];
};

// tile of the block owned by thread tid, from the row-major M x M matrix T (identity padding)
GJ_FN void gj_load(GjTile &t, const double *T, int tid)
{
    const int tr = tid / GJ_NT, tc = tid - tr * GJ_NT;
    const bool active = tid < GJ_NT * GJ_NT;
#pragma unroll
    for (int r = 0; r < GJ_B; r++)
#pragma unroll
        for (int c = 0; c < GJ_B; c++) {
            const int i = tr * GJ_B + r, j = tc * GJ_B + c;
            t.a[r][c] = (active && i < MDENSE && j < MDENSE) ? T[i * MDENSE + j] : ((i == j) ? 1.0 : 0.0);
        }
}

GJ_FN void gj_store(const GjTile &t, double *T, int tid)
{
    const int tr = tid / GJ_NT, tc = tid - tr * GJ_NT;
    if (tid >= GJ_NT * GJ_NT) return;
#pragma unroll
    for (int r = 0; r < GJ_B; r++)
#pragma unroll
        for (int c = 0; c < GJ_B; c++) {
            const int i = tr * GJ_B + r, j = tc * GJ_B + c;
            if (i < MDENSE && j < MDENSE) T[i * MDENSE + j] = t.a[r][c];
        }
}

// in-place Gauss-Jordan inverse of one tile in registers; false if a pivot is zero / not finite
GJ_FN bool gj_invert_tile(GjTile &t)
{
    bool ok = true;
#pragma unroll
    for (int k = 0; k < GJ_B; k++) {
        const double p = 1.0 / t.a[k][k];
        if (!isfinite(p) || p == 0.0) ok = false;
#pragma unroll
        for (int c = 0; c < GJ_B; c++) t.a[k][c] = (c == k) ? p : t.a[k][c] * p;
#pragma unroll
        for (int r = 0; r < GJ_B; r++) {
            if (r == k) continue;
            const double f = t.a[r][k];
#pragma unroll
            for (int c = 0; c < GJ_B; c++) t.a[r][c] = (c == k) ? -f * p : t.a[r][c] - f * t.a[k][c];
        }
    }
    return ok;
}

// constants of the panel (once, before the first barrier)
GJ_FN void gj_init_panel(double *pan, int tid)
{
    if (tid < GJ_TS) pan[GJ_MI + tid] = (tid / GJ_B == tid % GJ_B) ? -1.0 : 0.0;
    if (tid == 0) pan[GJ_OK] = 1.0;
}

// Publish what block step kn needs from this thread's tile (called with kn = 0 before the loop and with
// kn = kb + 1 at the end of step kb): the diagonal tile is inverted in place and D^-1 goes to its buffer,
// tiles of tile column kn go to the column panel ([k][r]), tiles of tile row kn to the row panel ([k][c]).
GJ_FN void gj_publish(GjTile &t, double *pan, int tid, int kn)
{
    const int tr = tid / GJ_NT, tc = tid - tr * GJ_NT;
    if (tid >= GJ_NT * GJ_NT || kn >= GJ_NT) return;
    if (tr == kn && tc == kn) {
        if (!gj_invert_tile(t)) pan[GJ_OK] = 0.0;
        double *d = pan + GJ_DI + 26 * (kn & 1);
#pragma unroll
        for (int k = 0; k < GJ_B; k++)
#pragma unroll
            for (int c = 0; c < GJ_B; c++) d[k * GJ_B + c] = t.a[k][c];
    } else if (tc == kn) {
        double *d = pan + GJ_CP + (kn & 1) * (GJ_NT * GJ_TS) + tr * GJ_TS;
#pragma unroll
        for (int k = 0; k < GJ_B; k++)
#pragma unroll
            for (int r = 0; r < GJ_B; r++) d[k * GJ_B + r] = t.a[r][k];
    } else if (tr == kn) {
        double *d = pan + GJ_RO + tc * GJ_TS;
#pragma unroll
        for (int k = 0; k < GJ_B; k++)
#pragma unroll
            for (int c = 0; c < GJ_B; c++) d[k * GJ_B + c] = t.a[k][c];
    }
}

// R = D^-1 * (row panel), one output element per (thread, pass); nthr threads share the GJ_NT * GJ_TS elements
GJ_FN void gj_scale_row_panel(double *pan, int tid, int nthr, int kb)
{
    const double *di = pan + GJ_DI + 26 * (kb & 1);
    for (int o = tid; o < GJ_NT * GJ_TS; o += nthr) {
        const int t = o / GJ_TS, e = o - t * GJ_TS, r = e / GJ_B, c = e - r * GJ_B;
        const double *ro = pan + GJ_RO + t * GJ_TS + c;
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < GJ_B; k++) v += di[r * GJ_B + k] * ro[k * GJ_B];
        pan[GJ_RN + o] = v;
    }
}

// block step kb on this thread's tile (see the header comment)
GJ_FN void gj_update(GjTile &t, const double *pan, int tid, int kb)
{
    const int tr = tid / GJ_NT, tc = tid - tr * GJ_NT;
    if (tid >= GJ_NT * GJ_NT) return;
    const bool on_row = tr == kb, on_col = tc == kb;
    const double *cp = on_row ? pan + GJ_MI : pan + GJ_CP + (kb & 1) * (GJ_NT * GJ_TS) + tr * GJ_TS;
    const double *rp = on_col ? pan + GJ_DI + 26 * (kb & 1) : pan + GJ_RN + tc * GJ_TS;
    if (on_row || on_col) {
#pragma unroll
        for (int r = 0; r < GJ_B; r++)
#pragma unroll
            for (int c = 0; c < GJ_B; c++) t.a[r][c] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < GJ_B; k++) {
        double cc[GJ_B], rr[GJ_B];
#pragma unroll
        for (int r = 0; r < GJ_B; r++) cc[r] = cp[k * GJ_B + r];
#pragma unroll
        for (int c = 0; c < GJ_B; c++) rr[c] = rp[k * GJ_B + c];
#pragma unroll
        for (int r = 0; r < GJ_B; r++)
#pragma unroll
            for (int c = 0; c < GJ_B; c++) t.a[r][c] -= cc[r] * rr[c];
    }
}
