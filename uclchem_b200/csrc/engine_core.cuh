// engine_core.cuh -- block-per-cell building blocks of the B200 UCLCHEM engine.
//
// One CTA (NET_NTHREADS threads, one per SM) owns one grid cell at a time and keeps
// the cell's whole working set in shared memory: the Newton matrix / LU factors
// (val), the rate coefficients, the reaction fluxes, the Nordsieck history and
// the work vectors.  HBM is touched only when a cell is loaded or stored.
// All irregular work (RHS gather, Jacobian assembly, sparse LU, triangular
// solves) runs as generated "team programs" (see makerates_cuda.py).
//
// Reference semantics reproduced here (file:line of the Fortran reference):
//   calculateReactionRates  rates.f90:21-343     -> calc_rates()
//   photoreactions.f90:46-304                    -> photo_rates_warp()
//   F / GETYDOT  chemistry.f90:294-352, odes.f90 -> rhs_eval()
//   DVJAC+DGEFA (dvode.f90:8182,11982) replaced by analytic J + fixed-pattern LU -> jac_factor()
//   DVSOL/DGESL (dvode.f90:8698,12091)           -> lin_solve()
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "../../include/uclgpu.h"
#include "net_tables.cuh"

#define NT NET_NTHREADS
#define NWARPS (NT / 32)
#define NEQ NET_NEQ
#define NSPEC NET_NSPEC
#define NAUG NET_NAUG
#define NREAC NET_NREAC
#define NSURF NET_NSURF
#define NICE NET_NICE
#define MDENSE NET_M
#define LMAXORD 6
#define NEQP ((NEQ + 7) & ~7)
#define NAUGP ((NAUG + 7) & ~7)
#define GJ_B 5                              /* square register tile of the dense Gauss-Jordan inverse */
#define GJ_NT ((MDENSE + GJ_B - 1) / GJ_B)  /* tiles per side */
#define GJ_TS (GJ_B * GJ_B)                 /* doubles per tile; panels are arrays of tiles (odd pitch: conflict-free) */
// shared-memory panels of the blocked Gauss-Jordan inverse (dense_inverse, engine_la.cuh), offsets in doubles
#define GJ_CP 0                             /* column panel, 2 x GJ_NT tiles [k][r] (double buffered) */
#define GJ_RO (2 * GJ_NT * GJ_TS)           /* row panel before scaling, GJ_NT tiles [k][c] */
#define GJ_RN (3 * GJ_NT * GJ_TS)           /* row panel times the inverse of the diagonal tile */
#define GJ_DI (4 * GJ_NT * GJ_TS)           /* inverse of the diagonal tile, 2 x 26 (double buffered) */
#define GJ_MI (GJ_DI + 52)                  /* minus the identity tile */
#define GJ_OK (GJ_MI + 26)                  /* "all pivots finite" flag */
#define GJ_PANEL (GJ_OK + 2)

// ---- constants.f90:2-18, surfacereactions.f90:34-52 (single-precision literals kept)
#define C_KBOLTZ 1.38065040e-16
#define C_AMU 1.66053892e-24
#define C_PI ((double)3.141592654f)
#define C_PC 3.086e18
#define C_SPY 3.16e7
#define C_MH 1.67262164e-24
#define C_GRAIN_RADIUS 1.e-5
#define C_MAX_GRAIN_TEMP 150.0
#define C_MIN_SURFACE_ABUND 1.0e-20
#define C_GCS (0.5 * (7.908e-22 + 8.473e-22)) /* GRAIN_CROSSSECTION_PER_H */
#define C_GSA (4.0 * C_GCS)                   /* GRAIN_SURFACEAREA_PER_H */
#define C_COV0 (0.5 * NET_GDR / NET_NSITES)   /* bulkGainFromMantleBuildUp, surfacereactions.f90:120-123 */
#define C_MIN_ABUND 1.0e-30

struct Scalars {
    double p[UCLGPU_NPARAM];
    // physicscore module state (physics-core.f90:11-20)
    double gastemp, dusttemp, density, av, coldens, cloudsize, radfield, zeta, zetascale, h2crprate;
    double time_in_years, current_time, target_time, current_time_old;
    // chemistry state (chemistry.f90:28-44)
    double h2col, cocol, ccol, safe_mantle, safe_bulk, blr, last_temp, phi, abstol_factor;
    // per-interval rate prefactors
    double k_desoh2, k_descr, k_deuvcr, stick_h, stick_h2, h2form_dust, scat_h2_pre, thermal_vel;
    int lh_on, swap_off, mxstep, kind;
    // rate-coefficient tables in use: the compiled-in network.f90 arrays, or the caller's overridden copy
    // (wrap.f90:744-761 alpha / beta / gamma dictionaries, uclgpu_opts.coeff_*)
    const double *c_alpha, *c_beta, *c_gama;
    long long step_budget;
    // RHS ext quantities at the last evaluated state
    double e_sm, e_sb, e_blr, e_ism, e_tsw, e_dblr, e_dism, e_S;
    double dflux[2]; // fluxes of the deferred photo reactions (H2 + hv, CO + hv)
    // hotcore / cshock
    double max_temp, vs, timestep_factor, min_postshock_temp;
    double cs_dlength, cs_z1, cs_z2, cs_z3, cs_v0, cs_at, cs_vn0, cs_zn0, cs_dissipation_time, cs_max_temp,
        cs_drift_vel, cs_zn, cs_vn;
    int temp_indx;
    // postprocess.f90 module state: this cell's tracer history [10][pp_ntime] (time s, density, gas T, dust T,
    // radfield, zeta, N_H, N_H2, N_CO, N_C) in global memory, read by thread 0 at output times
    const double *pp_grid;
    int pp_ntime, pp_coldens, pp_tstep;
    // jshock.f90 module state
    double js_max_temp, js_vmin, js_tshock, js_tcool, js_max_dens, js_t_lambda, js_n_lambda, js_v0;
    // collapse.f90 module state
    int col_mode;
    double col_max_time, col_parcel_radius, col_mass_in_radius;
    // BDF state (names follow dvode.f90)
    double tau[14], el[14], tq[6];
    double h, hu, hscal, hnew, tn, rc, prl1, rl1, eta, etamax, crate, drc, acnrm, conp, told, rtol, dsm, del,
        delp;
    int nq, l, lmax, nqwait, newq, newh, jstart, kflag, jcur, icf, ipup, nslp, nslj, nqu, ncf, nflag, m_iter;
    int nst_call; // NST of the current DVODE call
    int kuth, force_j, hist_valid, use_tcrit; // ITASK=4 style stop at TOUT; warm restart bookkeeping
    long long nst, nfe, nje, nlu, nni, ncfn, netf, nintervals;
    long long nsing, nmaxcor, ndiverge, nfailcall;
    long long cyc_rates, cyc_rhs, cyc_jac, cyc_factor, cyc_dense, cyc_solve, cyc_total;
    int flag, flag2; // block-wide decisions published by thread 0
    int fz_level, fz_save, fz_head; // how far the fused scalar sections got (engine_bdf.cuh)
    double dbg[64];
};

// Product-form solves (uclchem_b200/product_form.py): factor_p also forms the explicit
// sparse inverses of the factors of the sparse pivots (in place + NET_NVAL_PF - NET_NVAL fill slots) on the
// warps that have no tile in the dense Gauss-Jordan inverse, and lin_solve is five wide levels; the result is
// left in s.tmpv.  Default since round 2 (B200: solve 27.5 k -> 12.0 k cycles, 134.6 k -> 114.9 k cycles per
// BDF step, <= 8e-4 dex against the substitution build on 584 config-2 cells).  Networks that need the compact
// shared-memory layout (their flux array is aliased, and it is the staging buffer of the inverse program)
// and builds with -DUCLGPU_NO_PRODUCT_FORM keep the level-scheduled substitution.
#if !defined(UCLGPU_COMPACT_SMEM) && !defined(UCLGPU_NO_PRODUCT_FORM)
#define UCLGPU_PRODUCT_FORM
#ifndef UCLGPU_NO_PF_OVERLAP
#define UCLGPU_PF_OVERLAP
#endif
#endif
#ifdef UCLGPU_PRODUCT_FORM
#define NET_NVAL_STORE NET_NVAL_PF
#define SOLVE_RESULT(s) ((s).tmpv)
#else
#define NET_NVAL_STORE NET_NVAL
#define SOLVE_RESULT(s) ((s).xs)
#endif

struct __align__(16) Smem {
    double val[(NET_NVAL_STORE + 7) & ~7];
    double rate[(NREAC + 7) & ~7];
#ifdef UCLGPU_COMPACT_SMEM
    // Larger networks (crp_photo: 136 KB Newton matrix) do not leave room for everything.  The flux
    // array is live only inside rhs_eval; the solve vectors and the Gauss-Jordan panels are live only
    // between two RHS evaluations (newton_rhs -> lin_solve -> read-out, and dense_inverse), so they
    // share its storage.  Not combinable with the product form (its staging buffer is flux).
    union {
        double flux[(NREAC + 7) & ~7];
        struct {
            double xs[NAUGP];
            double tmpv[NAUGP];
            double gj_pan[GJ_PANEL];
        };
    };
#else
    double flux[(NREAC + 7) & ~7];
#endif
    double y[NEQP + 8];        // iterate + ext slots y[NEQ+0..3] = {1, blr, 1/safeMantle, tau}
    double yh[LMAXORD][NEQP];  // Nordsieck array
    double ewt[NEQP], savf[NEQP], acor[NEQP], atol[NEQP], abund[NEQP];
#ifndef UCLGPU_COMPACT_SMEM
    // xs / tmpv: linear-solve vectors (elimination ordering); invd: reciprocal sparse pivots (copied out of
    // val at the end of factor_p).  None of the three is live while the dense block is inverted -- the
    // Newton right-hand side is formed after the factorisation -- so the Gauss-Jordan panels (15.8 KB)
    // alias them.
    union {
        struct {
            double xs[NAUGP];
            double tmpv[NAUGP];
            double invd[NAUGP];
        };
        double gj_pan[GJ_PANEL];
    };
#else
    double invd[NAUGP];        // reciprocal sparse pivots (copied out of val after each factorisation)
#endif
    double red[2][32];
    Scalars st;
};

// Block barrier.  __syncthreads() is the *aligned* barrier (barrier.sync.aligned): every
// thread of a warp must arrive together, and a warp that is still diverged when it arrives
// (thread 0's scalar sections, the single lane that evaluates the CO photo rate) is undefined
// behaviour -- compute-sanitizer synccheck/racecheck showed warps slipping one barrier ahead.
// The non-aligned form counts arrivals per thread and is safe for diverged warps.
#define BLOCK_SYNC() asm volatile("barrier.sync 0;" ::: "memory")

struct Blk {
    unsigned phase; // ping-pong index for the reduction scratch
    double *jsv;    // this CTA's saved-Jacobian scratch in global memory (stays L2 resident)
    double *trace;  // debug trace buffer (null unless UCLGPU_TRACE is set and this CTA owns cell 0)
    int trace_cap, trace_n, dump_at;
    double *dump;
};

// phase timers: thread 0 accumulates SM clock cycles per phase (diagnostics in uclgpu_stats)
#define TIMER_START long long t_start_ = clock64();
#define TIMER_ADD(field)                                    \
    if (threadIdx.x == 0) {                                  \
        long long t_now_ = clock64();                        \
        s.st.field += t_now_ - t_start_;                     \
        t_start_ = t_now_;                                   \
    }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Sum over the block, returned to every thread (one barrier; scratch is ping-ponged).
__device__ __forceinline__ double block_sum(Smem &s, Blk &b, double v)
{
    v = warp_sum(v);
    double *red = s.red[b.phase & 1];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    BLOCK_SYNC();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < NWARPS; i++) t += red[i];
    b.phase++;
    return t;
}
__device__ __forceinline__ double block_min(Smem &s, Blk &b, double v)
{
    v = warp_min(v);
    double *red = s.red[b.phase & 1];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    BLOCK_SYNC();
    double t = red[0];
#pragma unroll
    for (int i = 1; i < NWARPS; i++) t = fmin(t, red[i]);
    b.phase++;
    return t;
}

// DVNORM dvode.f90:9021: sqrt(sum((v*w)^2)/N), N = NEQ (density and BULK/SURFACE included)
__device__ __forceinline__ double wrms_norm(Smem &s, Blk &b, const double *v, const double *w)
{
    double t = 0.0;
    if (threadIdx.x < NEQ) {
        t = v[threadIdx.x] * w[threadIdx.x];
        t = t * t;
    }
    return sqrt(block_sum(s, b, t) / (double)NEQ);
}

// =====================================================================================
// photoreactions.f90
// =====================================================================================
// splint photoreactions.f90:440-560 (bisection bracket + end clamps)
__device__ __forceinline__ double nr_splint(const double *xa, const double *ya, const double *y2a, int n,
                                            double x)
{
    int jlo = 0, jhi = n + 1;
    while (jhi - jlo != 1) {
        int jm = (jhi + jlo) >> 1;
        if (x > xa[jm - 1]) jlo = jm; else jhi = jm;
    }
    if (jlo == 0) { jlo = 1; jhi = 2; }
    if (jlo == n) { jlo = n - 1; jhi = n; }
    double h = xa[jhi - 1] - xa[jlo - 1];
    double a = (xa[jhi - 1] - x) / h;
    double bb = (x - xa[jlo - 1]) / h;
    return a * ya[jlo - 1] + bb * ya[jhi - 1] +
           ((a * a * a - a) * y2a[jlo - 1] + (bb * bb * bb - bb) * y2a[jhi - 1]) * (h * h) / 6.0;
}

// xlambda photoreactions.f90:225-243 (second derivatives tabulated at generation time)
__device__ __forceinline__ double xlambda_dev(double lambda)
{
    double v = lambda;
    if (lambda < net_lambda_grid[0]) v = net_lambda_grid[0];
    if (lambda > net_lambda_grid[29]) v = net_lambda_grid[29];
    double r = nr_splint(net_lambda_grid, net_xlambda_grid, net_xlambda_d2, 30, v);
    return r < 0.0 ? 0.0 : r;
}

// scatter photoreactions.f90:133-196
__device__ __forceinline__ double scatter_dev(double x1, double av)
{
    const double c[6] = {1.0, 2.006, -1.438, 7.364e-01, -5.076e-01, -5.920e-02};
    const double k1[6] = {7.514e-01, 8.490e-01, 1.013, 1.282, 2.005, 5.832};
    double tv = av / 1.086;
    double tl = tv * xlambda_dev(x1);
    double sc = 0.0;
    if (tl < 1.0) {
        double expo = k1[0] * tl;
        if (expo < 100.0) sc = c[0] * exp(-expo);
    } else {
#pragma unroll
        for (int i = 1; i < 6; i++) {
            double expo = k1[i] * tl;
            if (expo < 100.0) sc = sc + c[i] * exp(-expo);
        }
    }
    return sc;
}

// H2SelfShielding photoreactions.f90:91-127 with dopplerWidth = turbVel/(1000*1e-8), radWidth = 8e7
__device__ __forceinline__ double h2_self_shielding_dev(double nh2)
{
    const double doppler = 1.0 / (1000.0 * 1.0e-8), radw = 8.0e7;
    double taud = 0.5 * nh2 * (double)1.5e-2f * 1.0e-2 / doppler;
    double sj;
    if (taud == 0.0) sj = 1.0;
    else if (taud < 2.0) sj = exp(-0.6666667 * taud);
    else if (taud < 10.0) sj = 0.638 * pow(taud, -1.25);
    else if (taud < 100.0) sj = 0.505 * pow(taud, -1.15);
    else sj = 0.344 * pow(taud, -1.0667);
    double r = radw / (1.7724539 * doppler);
    double t = 3.02 * pow(r * 1.0e+03, -0.064);
    double u = sqrt(taud * r) / t;
    double sr = r / (t * sqrt(0.78539816 + u * u));
    return sj + sr;
}

// natural cubic spline second derivatives, photoreactions.f90:335-409 (n <= 8)
__device__ __forceinline__ void nr_spline_small(const double *x, const double *y, int n, double *y2)
{
    double u[8];
    y2[0] = 0.0;
    u[0] = 0.0;
    for (int i = 1; i < n - 1; i++) {
        double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
        double p = sig * y2[i - 1] + 2.0;
        y2[i] = (sig - 1.0) / p;
        u[i] = (6.0 * ((y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1])) /
                    (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p;
    }
    y2[n - 1] = (0.0 - 0.0 * u[n - 2]) / (0.0 * y2[n - 2] + 1.0);
    for (int k = n - 2; k >= 0; k--) y2[k] = y2[k] * y2[k + 1] + u[k];
}

// splint on a unit-spaced grid xa[k] = x0 + k (k = 0..n-1): same bracket rule and arithmetic as
// nr_splint, with the abscissae computed instead of loaded (they are exact small integers)
__device__ __forceinline__ double nr_splint_unit(double x0, const double *ya, const double *y2a, int n, double x)
{
    int jlo = 0;
    for (int k = 0; k < n; k++) jlo += (x > x0 + (double)k) ? 1 : 0; // = the bisection result on a sorted grid
    int jhi = jlo + 1;
    if (jlo == 0) { jlo = 1; jhi = 2; }
    if (jlo == n) { jlo = n - 1; jhi = n; }
    double xlo = x0 + (double)(jlo - 1), xhi = x0 + (double)(jhi - 1);
    double h = xhi - xlo;
    double a = (xhi - x) / h;
    double bb = (x - xlo) / h;
    return a * ya[jlo - 1] + bb * ya[jhi - 1] +
           ((a * a * a - a) * y2a[jlo - 1] + (bb * bb * bb - bb) * y2a[jhi - 1]) * (h * h) / 6.0;
}

// COPhotoDissRate photoreactions.f90:57-72,251-304 -- executed by ONE thread, on the critical path
// of every RHS evaluation (the other warps wait for it at the barrier after the plain flux phase):
// the shielding tables are in constant memory and both grids (log N(CO) = 12..18, log N(H2) = 18..23)
// are unit spaced, so no table walk touches L2 or local memory.
__device__ __noinline__ double co_photo_rate_dev(double nh2, double nco, double radfield, double av)
{
    double lognco = log10(nco + 1.0), lognh2 = log10(nh2 + 1.0);
    double lu = log10(fabs(nco) + 1.0), lw = log10(fabs(nh2) + 1.0);
    if (lognco < 12.0) lognco = 12.0;
    if (lognh2 < 18.0) lognh2 = 18.0;
    if (lognco > 18.0) lognco = 18.0;
    if (lognh2 > 23.0) lognh2 = 23.0;
    double yy[7], y2[7];
#pragma unroll
    for (int j = 0; j < 7; j++) yy[j] = nr_splint_unit(18.0, net_sco_rows + 6 * j, net_sco_d2 + 6 * j, 6, lognh2);
    // natural spline over the unit-spaced N(CO) grid (nr_spline_small with x[i] = 12 + i)
    {
        double u[7];
        y2[0] = 0.0;
        u[0] = 0.0;
#pragma unroll
        for (int i = 1; i < 6; i++) {
            const double xm = 12.0 + (double)(i - 1), x0 = 12.0 + (double)i, xp = 12.0 + (double)(i + 1);
            double sig = (x0 - xm) / (xp - xm);
            double p = sig * y2[i - 1] + 2.0;
            y2[i] = (sig - 1.0) / p;
            u[i] = (6.0 * ((yy[i + 1] - yy[i]) / (xp - x0) - (yy[i] - yy[i - 1]) / (x0 - xm)) / (xp - xm) - sig * u[i - 1]) / p;
        }
        y2[6] = (0.0 - 0.0 * u[5]) / (0.0 * y2[5] + 1.0);
#pragma unroll
        for (int k = 5; k >= 0; k--) y2[k] = y2[k] * y2[k + 1] + u[k];
    }
    // splint over N(CO): bracket computed, operands selected from registers
    double ssf;
    {
        int jlo = 0;
#pragma unroll
        for (int k = 0; k < 7; k++) jlo += (lognco > 12.0 + (double)k) ? 1 : 0;
        int jhi = jlo + 1;
        if (jlo == 0) { jlo = 1; jhi = 2; }
        if (jlo == 7) { jlo = 6; jhi = 7; }
        double ylo = yy[0], yhi = yy[1], dlo = y2[0], dhi = y2[1];
#pragma unroll
        for (int k = 1; k < 6; k++)
            if (jlo - 1 == k) { ylo = yy[k]; yhi = yy[k + 1]; dlo = y2[k]; dhi = y2[k + 1]; }
        double xlo = 12.0 + (double)(jlo - 1), xhi = 12.0 + (double)(jhi - 1);
        double h = xhi - xlo;
        double a = (xhi - lognco) / h;
        double bb = (lognco - xlo) / h;
        ssf = pow(10.0, a * ylo + bb * yhi + ((a * a * a - a) * dlo + (bb * bb * bb - bb) * dhi) * (h * h) / 6.0);
    }
    double lb = (5675.0 - 200.6 * lw) - (571.6 - 24.09 * lw) * lu + (18.22 - 0.7664 * lw) * (lu * lu);
    if (lb > 1076.1) lb = 1076.1;
    if (lb < 913.6) lb = 913.6;
    double sca = scatter_dev(lb, av);
    return (2.e-10) * (radfield / (double)1.7f) * ssf * sca;
}

// cIonizationRate photoreactions.f90:74-85
__device__ __forceinline__ double c_ionization_rate_dev(double alpha, double gamma, double gastemp, double nc,
                                                        double nh2, double av, double radfield)
{
    double tauc = gamma * av + 1.1e-17 * nc + (0.9 * pow(gastemp, 0.27) * pow(nh2 / 1.59e21, 0.45));
    return alpha * (radfield / (double)1.7f) * exp(-tauc);
}

// h2FormEfficiency surfacereactions.f90:63-107
__device__ __noinline__ double h2_form_efficiency_dev(double gastemp, double dusttemp)
{
    const double SIL_MU = 0.005, SIL_ES = 110.0, SIL_EH2 = 320.0, SIL_EHP = 450.0, SIL_EHC = 3.0e4,
                 SIL_NUH2 = 3.0e12, SIL_NUHC = 1.3e13, SIL_CS = 8.473e-22;
    const double GRA_MU = 0.005, GRA_ES = 260.0, GRA_EH2 = 520.0, GRA_EHP = 800.0, GRA_EHC = 3.0e4,
                 GRA_NUH2 = 3.0e12, GRA_NUHC = 1.3e13, GRA_CS = 7.908e-22;
    double thermal_velocity = 1.45e5 * sqrt(gastemp / 1.0e2);
    double sticking = 1.0 / (1.0 + 0.04 * sqrt(gastemp + dusttemp) + 0.2 * (gastemp / 1.0e2) +
                             0.08 * ((gastemp / 1.0e2) * (gastemp / 1.0e2)));
    double hflux = 1.0e-10;
    double f1, f2, eps, sq, sil, gra;
    f1 = SIL_MU * hflux / (2 * SIL_NUH2 * exp(-SIL_EH2 / dusttemp));
    sq = (1.0 + sqrt((SIL_EHC - SIL_ES) / (SIL_EHP - SIL_ES)));
    f2 = 1.0 * (sq * sq) / 4.0 * exp(-SIL_ES / dusttemp);
    eps = 1.0 / (1.0 + SIL_NUHC / (2 * hflux) * exp(-1.5 * SIL_EHC / dusttemp) * (sq * sq));
    sil = 1.0 / (1.0 + f1 + f2) * eps;
    f1 = GRA_MU * hflux / (2 * GRA_NUH2 * exp(-GRA_EH2 / dusttemp));
    sq = (1.0 + sqrt((GRA_EHC - GRA_ES) / (GRA_EHP - GRA_ES)));
    f2 = 1.0 * (sq * sq) / 4.0 * exp(-GRA_ES / dusttemp);
    eps = 1.0 / (1.0 + GRA_NUHC / (2 * hflux) * exp(-1.5 * GRA_EHC / dusttemp) * (sq * sq));
    gra = 1.0 / (1.0 + f1 + f2) * eps;
    return 0.5 * thermal_velocity * (SIL_CS * sil + GRA_CS * gra) * sticking;
}

// =====================================================================================
// calculateReactionRates, rates.f90:21-343.  s.abund holds abund(:,dstep); st holds physics.
// Three phases: unmasked per-reaction rates | freeze-out switches | masks + photo overrides.
// =====================================================================================
__device__ __forceinline__ double freeze_rate_dev(const Scalars &st, int r)
{
    // freezeOutRate rates.f90:354-367
    double fr = 1.0 + st.c_beta[r] * 16.71e-4 / (C_GRAIN_RADIUS * st.gastemp);
    if (st.p[UCL_P_FREEZEFACTOR] == 0.0 || st.dusttemp > C_MAX_GRAIN_TEMP) return 0.0;
    return fr * st.p[UCL_P_FREEZEFACTOR] * st.c_alpha[r] * st.thermal_vel * sqrt(st.gastemp / net_mass1[r]) * C_GCS;
}

__device__ __forceinline__ double diffusion_rate_dev(const Scalars &st, int r)
{
    // diffusionReactionRate surfacereactions.f90:158-211 (tunnelling term tabulated)
    int i1 = net_ia[r], i2 = net_ib[r];
    double td = st.dusttemp;
    double v1 = net_vdiff[i1], v2 = net_vdiff[i2], e1 = net_binding_energy[i1], e2 = net_binding_energy[i2];
    double diffuse = v1 * exp(-0.5 * e1 / td);
    diffuse = diffuse + (v2 * exp(-0.5 * e2 / td));
    double desorb = v1 * exp(-e1 / td);
    desorb = desorb + v2 * exp(-e2 / td);
    double reac = st.c_gama[r] / td;
    double tunnel = net_tunnel[r];
    if (reac > tunnel) reac = tunnel;
    reac = fmax(v1, v2) * exp(-reac);
    reac = reac / (reac + desorb + diffuse);
    return st.c_alpha[r] * reac * diffuse * NET_GDR / NET_NSITES;
}

__device__ __noinline__ void calc_rates(Smem &s)
{
    Scalars &st = s.st;
    const int tid = threadIdx.x;
    TIMER_START
    // ---- per-interval scalars (thread 0) ------------------------------------------------
    if (tid == 0) {
        const double *p = st.p;
        const bool desorb = p[UCL_P_DESORB] != 0.0;
        const bool mantle = st.safe_mantle > C_MIN_SURFACE_ABUND;
        st.thermal_vel = sqrt(8.0 * C_KBOLTZ / (C_PI * C_AMU));
        st.k_desoh2 = (desorb && p[UCL_P_H2DESORB] != 0.0 && mantle)
                          ? p[UCL_P_EPSILON] * h2_form_efficiency_dev(st.gastemp, st.dusttemp) : 0.0;
        // NB st.phi is the module variable clobbered by the GAR block (rates.f90:322-326)
        st.k_descr = (desorb && p[UCL_P_CRDESORB] != 0.0 && mantle)
                         ? 4.0 * C_PI * st.zeta * 1.64e-4 * (C_GSA)*st.phi : 0.0;
        if (desorb && p[UCL_P_UVDESORB] != 0.0 && mantle && st.zeta > 0) {
            double v = C_GCS * p[UCL_P_UV_YIELD] * 4.875e3 * st.zeta;
            st.k_deuvcr = v * (1 + (st.radfield / p[UCL_P_UVCREFF]) * (1.0 / st.zeta) * exp(-(double)1.8f * st.av));
        } else {
            st.k_deuvcr = 0.0;
        }
        {
            double tr = st.gastemp / 52.0;
            st.stick_h = 1.0 * (1.0 + 2.5 * tr) / pow(1.0 + tr, 2.5);
            tr = st.gastemp / 87.0;
            st.stick_h2 = 0.87 * (1.0 + 2.5 * tr) / pow(1.0 + tr, 2.5);
        }
        st.h2form_dust = h2_form_efficiency_dev(st.dusttemp, st.dusttemp);
        st.lh_on = (st.dusttemp < C_MAX_GRAIN_TEMP) && mantle;
        st.swap_off = (st.dusttemp > C_MAX_GRAIN_TEMP) || (st.safe_mantle < C_MIN_SURFACE_ABUND);
        st.scat_h2_pre = 5.18e-11 * (st.radfield / (double)1.7f) * scatter_dev(1000.0, st.av);
    }
    BLOCK_SYNC();
    const double gastemp = st.gastemp, dusttemp = st.dusttemp, zeta = st.zeta, av = st.av, radfield = st.radfield;
    const double ICE_GAS = (double)0.3f;
    // ---- phase 1: unmasked rates ---------------------------------------------------------
    for (int r = tid; r < NREAC; r += NT) {
        double k = 0.0;
        const double al = st.c_alpha[r], be = st.c_beta[r], ga = st.c_gama[r];
        switch (net_rtype[r]) {
        case 1: /* CRP :34-41 */
            k = (NET_CRP_LO != NET_CRP_HI) ? al * zeta : 0.0;
            if (st.p[UCL_P_IMPROVEDH2CRPDISSOCIATION] != 0.0 && r == NET_NR_H2_CRP) k = st.h2crprate;
            break;
        case 0: /* PHOTON :43-56 */
            if (NET_PHOTON_LO != NET_PHOTON_HI) {
                k = al * exp(-ga * av) * radfield / (double)1.7f;
                if (net_phase[r] == 2) k = k * ICE_GAS * pow((double)(1.0f - 0.007f), ((double)1.0f + (double)0.5f / st.blr));
                else if (net_phase[r] == 1) k = k * ICE_GAS;
            }
            break;
        case 2: /* CRPHOT :59-72 */
            if (NET_CRPHOT_LO != NET_CRPHOT_HI) {
                k = al * ga * 1.0 / (1.0 - st.p[UCL_P_OMEGA]) * zeta * pow(gastemp / 300, be);
                if (net_phase[r] == 2) k = k * ICE_GAS * pow((double)(1 - 0.007f), (1 + (double)0.5f / st.blr));
                else if (net_phase[r] == 1) k = k * ICE_GAS;
            }
            break;
        case 3: /* FREEZE :75-84 */
            if (NET_FREEZE_LO != NET_FREEZE_HI) {
                if (r == NET_NR_EFREEZE) k = freeze_rate_dev(st, NET_NR_HFREEZE);
                else k = freeze_rate_dev(st, r);
                if (r == NET_NR_H2FREEZE) k = st.stick_h2 * k;
                if (r == NET_NR_HFREEZE) k = st.stick_h * k;
            }
            break;
        case 6: /* DESOH2 :91-103 */
            k = (NET_DESOH2_LO != NET_DESOH2_HI && !(ga > st.p[UCL_P_EBMAXH2])) ? st.k_desoh2 : 0.0;
            break;
        case 7: /* DESCR :110-125 */
            k = (NET_DESCR_LO != NET_DESCR_HI && !(ga > st.p[UCL_P_EBMAXCR])) ? st.k_descr : 0.0;
            break;
        case 8: /* DEUVCR :132-148 */
            k = (NET_DEUVCR_LO != NET_DEUVCR_HI && !(ga > st.p[UCL_P_EBMAXUVCR])) ? st.k_deuvcr : 0.0;
            break;
        case 5: /* THERM :202-229 */
            if (NET_THERM_LO != NET_THERM_HI && st.p[UCL_P_THERMDESORB] != 0.0 && net_ia[r] >= 0)
                k = net_vdiff[net_ia[r]] * exp(-ga / dusttemp);
            break;
        case 12: /* LH :236-260 */
        case 13: /* LHDES */
            if (NET_LH_LO != NET_LH_HI && st.lh_on) {
                int is_des = net_rtype[r] == 13;
                int rl = is_des ? net_partner[r] : r; // the LH reaction whose diffusion rate is shared
                double base = diffusion_rate_dev(st, rl);
                int rd = is_des ? r : net_partner[r];
                double des = net_desfrac[rd] * base;
                if (net_phase[rd] == 2) des = 0.0; // bulk species cannot chemically desorb
                k = is_des ? des : base - des;
            }
            break;
        case 10: /* ER :264-279 (only when the type has more than one reaction) */
        case 11: /* ERDES */
            if (NET_ER_LO != NET_ER_HI) {
                int is_des = net_rtype[r] == 11;
                int rl = is_des ? net_partner[r] : r;
                double base = freeze_rate_dev(st, rl) * exp(-st.c_gama[rl] / dusttemp);
                int rd = is_des ? r : net_partner[r];
                double des = net_desfrac[rd] * base;
                if (net_phase[rd] == 2) des = 0.0;
                k = is_des ? des : base - des;
            }
            break;
        case 9: /* H2FORM :281-289 */
            k = 0.0;
            break;
        case 14: /* BULKSWAP surfacereactions.f90:135-150 */
            if (!st.swap_off && net_ia[r] >= 0) k = net_vdiff[net_ia[r]] * exp(-net_binding_energy[net_ia[r]] / dusttemp);
            break;
        case 15: /* SURFSWAP surfacereactions.f90:125-132 */
            k = st.swap_off ? 0.0 : 1.0;
            break;
        case 22: /* TWOBODY :295-299 */
            k = al * (pow(gastemp / (double)300.f, be)) * exp(-ga / gastemp);
            break;
        case 16: /* IONOPOL1 :301-305 */
            k = (NET_IONOPOL1_LO != NET_IONOPOL1_HI) ? al * be * (0.62 + 0.4767 * ga * sqrt(300.0 / gastemp)) : 0.0;
            break;
        case 17: /* IONOPOL2 :307-313 */
            k = (NET_IONOPOL2_LO != NET_IONOPOL2_HI)
                    ? al * be * (1.0 + 0.0967 * ga * sqrt(300.0 / gastemp) + ga * ga * (double)300.0f / ((double)10.526f * gastemp))
                    : 0.0;
            break;
        case 18: /* CRS :156-162 */
            k = (NET_CRS_LO != NET_CRS_HI) ? al * (be * (ga / 100) * ((double)8.6f * zeta * (double)1.3f)) : 0.0;
            break;
        case 20: /* EXRELAX :165-176 */
            k = (NET_EXRELAX_LO != NET_EXRELAX_HI && net_ia[r] >= 0) ? net_vdiff[net_ia[r]] : 0.0;
            break;
        case 19: /* EXSOLID :179-197 */
            if (NET_EXSOLID_LO != NET_EXSOLID_HI && net_ia[r] >= 0 && net_ib[r] >= 0)
                k = al * ((net_vdiff[net_ib[r]] + net_vdiff[net_ia[r]]) / (1.5e15 * 1.8e-8));
            break;
        case 21: /* GAR :316-332 -- uses the phi computed in THIS call */
            k = 0.0; /* filled in phase 3 once phi is known */
            break;
        default:
            k = 0.0;
        }
        if (r == NET_NR_H2FORM_CT) k = st.h2form_dust; /* rates.f90:282 */
        if (r == NET_NR_H2FORM_ER || r == NET_NR_H2FORM_ERDES) k = 0.0;
        s.rate[r] = k;
    }
    BLOCK_SYNC();
    // ---- phase 2: freeze-out switches rates.f90:106-107,127-128,150-151,223-225 -------------
    if (tid < NSURF) {
        int fp = net_freeze_partners[tid];
        double rf = s.rate[fp];
        double a = s.abund[net_re1[fp]];
        bool off = false;
        if (NET_DESOH2_LO != NET_DESOH2_HI) off |= (rf * a) < C_MIN_SURFACE_ABUND * s.rate[NET_DESOH2_LO + tid];
        if (NET_DESCR_LO != NET_DESCR_HI) off |= (rf * a * st.density) < C_MIN_SURFACE_ABUND * s.rate[NET_DESCR_LO + tid];
        if (NET_DEUVCR_LO != NET_DEUVCR_HI) off |= (rf * a * st.density) < C_MIN_SURFACE_ABUND * s.rate[NET_DEUVCR_LO + tid];
        if (NET_THERM_LO != NET_THERM_HI && st.p[UCL_P_THERMDESORB] != 0.0)
            off |= (rf * a * st.density) < C_MIN_SURFACE_ABUND * s.rate[NET_THERM_LO + tid];
        if (off) s.rate[fp] = 0.0; // partners are distinct: no thread reads another's rate[fp]
    }
    BLOCK_SYNC();
    // ---- phase 3: THERM mantle guard, GAR/phi, T-range masks, photo overrides ----------------
    double phi_new = radfield * exp(-(double)2.5f * av) * sqrt(gastemp) / (s.abund[NSPEC] * s.abund[NET_NELEC]);
    phi_new = fmin(fmax(phi_new, (double)1e2f), (double)1e6f);
    for (int r = tid; r < NREAC; r += NT) {
        double k = s.rate[r];
        int ty = net_rtype[r];
        if (ty == 5 && st.safe_mantle < C_MIN_SURFACE_ABUND) k = 0.0;
#if NET_NGAR > 0
        if (ty == 21 && NET_GAR_LO != NET_GAR_HI) {
            const double *g = net_gar_params + 7 * (r - NET_GAR_LO);
            k = (double)0.6f * st.c_alpha[r] * g[0] /
                ((double)1.f + g[1] * pow(phi_new, g[2]) *
                                   ((double)1.f + g[3] * pow(gastemp, g[4]) * pow(phi_new, -g[5] - g[6] * log(gastemp))));
        }
#endif
        if (!net_extrapolate[r] && gastemp < net_min_temps[r]) k = 0.0;
        if (!net_extrapolate[r] && gastemp > net_max_temps[r]) k = 0.0;
        if (r == NET_NR_H2_HV) k = st.scat_h2_pre * h2_self_shielding_dev(st.h2col);
        if (r == NET_NR_CO_HV) k = co_photo_rate_dev(st.h2col, st.cocol, radfield, av);
        if (r == NET_NR_C_HV)
            k = c_ionization_rate_dev(st.c_alpha[r], st.c_gama[r], gastemp, st.ccol, st.h2col, av, radfield);
        s.rate[r] = k;
    }
    BLOCK_SYNC();
    if (tid == 0) {
        st.phi = phi_new;
        st.last_temp = gastemp;
    }
    BLOCK_SYNC();
    TIMER_ADD(cyc_rates)
}

// densdot physics-core.f90:90-103
__device__ __forceinline__ double densdot_dev(const Scalars &st, double density)
{
    if (density < st.p[UCL_P_FINALDENS] && st.p[UCL_P_FREEFALL] != 0.0) {
        double id = st.p[UCL_P_INITIALDENS];
        double e = (double)0.33f;
        return st.p[UCL_P_FREEFALLFACTOR] * pow(pow(density, (double)4.f) / id, e) *
               pow(8.4e-30 * id * (pow(density / id, e) - (double)1.f), (double)0.5f);
    }
    return 0.0;
}
// d(densdot)/dn by differentiating the expression above
__device__ __forceinline__ double ddensdot_dev(const Scalars &st, double n)
{
    if (n < st.p[UCL_P_FINALDENS] && st.p[UCL_P_FREEFALL] != 0.0) {
        double id = st.p[UCL_P_INITIALDENS];
        double e = (double)0.33f;
        double a = pow(pow(n, 4.0) / id, e);          // ~ n^(4e)
        double u = 8.4e-30 * id * (pow(n / id, e) - 1.0);
        if (!(u > 0.0)) return 0.0;
        double b = sqrt(u);
        double da = a * 4.0 * e / n;
        double du = 8.4e-30 * id * e * pow(n / id, e) / n;
        return st.p[UCL_P_FREEFALLFACTOR] * (da * b + a * 0.5 * du / b);
    }
    return 0.0;
}

