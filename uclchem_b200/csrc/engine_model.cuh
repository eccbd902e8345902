// engine_model.cuh -- per-cell model driver on the device.
//
// Restates, for one cell per CTA, the reference's driver and physics hooks:
//   solveAbundances          wrap.f90:549-697
//   initializeChemistry      chemistry.f90:46-141
//   updateChemistry          chemistry.f90:145-239 (usepostprocess=.true. semantics)
//   integrateODESystem       chemistry.f90:241-292 (ISTATE policy)
//   coreInitialize/UpdatePhysics physics-core.f90:42-88
//   cloud.f90:20-65, hotcore.f90:30-107, cshock.f90:38-252, sputtering.f90:65-235
//   output row layout        io.f90:59-83
#pragma once
#include "engine_bdf.cuh"
#include "engine_collapse.cuh"

#define JSV_STRIDE ((NET_NVAL + 31) & ~31)

struct RunArgs {
    int kind;
    long long ncell;       // cells in the input arrays (stride of params)
    long long nrun;        // cells this launch integrates: queue positions 0..nrun-1, cell = order ? order[pos] : pos
    int compact;           // 1: results are indexed by queue position (compact), 0: by cell
    const double *params;  // [UCLGPU_NPARAM][ncell]
    const double *y0;      // [ncell][NEQ] or null; with y0_index: a table, cell c starts from row y0_index[c]
    const int *y0_index;   // [ncell] or null
    double *y_final;       // [ncell][NEQ]
    double *phys_final;    // [ncell][UCLGPU_NPHYS] or null
    int *flag;             // [ncell]
    uclgpu_stats *stats;   // [ncell] or null
    int timepoints;        // trajectories: rows = timepoints+1
    double *phys_traj, *chem_traj, *rates_traj, *tdiss;
    unsigned long long *counter; // work queue
    double *jsave;               // [gridDim.x][JSV_STRIDE] saved-Jacobian scratch
    double *trace;               // debug: [trace_cap][8] Newton-iteration records of cell 0 (or null)
    int trace_cap, dump_at;
    int warm_restart;            // experimental (UCLGPU_WARM=1): keep the BDF history across output times
    long long max_steps;         // uclgpu_opts.step_budget: abandon a cell (flag -5) beyond this many BDF steps; 0 = off
    const double *pp_grid;       // postprocess: [ncell][10][pp_ntime] tracer histories, or null
    int pp_ntime, pp_coldens;
    const double *coef;          // [3][NREAC] overridden alpha / beta / gamma tables, or null
    const int *order;            // processing order of the cells (most expensive first) or null
    double *dump;
};

// ---- physics hooks (thread 0) ---------------------------------------------------------------
__device__ __noinline__ void ionization_dependency_dev(Scalars &st)
{
    // physics-core.f90:121-157: zeta itself is never updated there; only h2CRPRate
    if (st.p[UCL_P_IMPROVEDH2CRPDISSOCIATION] != 0.0) {
        const double ckL[10] = {1.582911005330e7, -6.465722684896e6, 1.172189025424e6, -1.237950798073e5,
                                8.393404654312e3, -3.788811358130e2, 1.138688455029e1, -2.197136304567e-1,
                                2.469841278950e-3, -1.232393620924e-5};
        const double ckH[10] = {1.217227462831e7, -4.989649250304e6, 9.079152156645e5, -9.624890825395e4,
                                6.551161486120e3, -2.968976216187e2, 8.959037875226e0, -1.735757324445e-1,
                                1.959267277734e-3, -9.816996707980e-6};
        const bool L = st.p[UCL_P_IONMODEL] == 0.0;
        double sum = 0.0, lc = log10(st.coldens);
        for (int k = 0; k < 10; k++) sum += (L ? ckL[k] : ckH[k]) * pow(lc, (double)k);
        st.h2crprate = pow(10.0, sum) * st.zetascale;
    }
}

__device__ double coshinv_f_dev(float x)
{
    // single-precision constant expression log((1/x)+sqrt((1/x)**2-1)), cshock.f90:104,111
    float inv = 1.0f / x;
    float sq = (float)sqrt((double)(inv * inv - 1.0f));
    return (double)(float)log((double)(inv + sq));
}

__device__ __noinline__ int cshock_initialize_dev(Scalars &st)
{
    // cshock.f90:38-141
    double *p = st.p;
    st.vs = p[UCL_P_VS];
    st.timestep_factor = p[UCL_P_TIMESTEPFACTOR];
    st.min_postshock_temp = p[UCL_P_MINIMUMPOSTSHOCKTEMP];
    st.cs_drift_vel = 0.0;
    st.cs_zn0 = 0.0;
    st.cs_vn0 = 0.0;
    st.cloudsize = (p[UCL_P_ROUT] - p[UCL_P_RIN]) * C_PC;
    if (p[UCL_P_FREEFALL] != 0.0) p[UCL_P_FREEFALL] = 0.0;
    if (p[UCL_P_POINTS] > 1) return -1;
    const double id = p[UCL_P_INITIALDENS], vs = st.vs;
    st.density = id;
    st.current_time_old = 0.0;
    double max_temp;
    if (id > (double)316227.78f /* 10**5.5 in single precision */) {
        max_temp = ((double)2.91731f * vs * vs) - ((double)23.78974f * vs) + (double)225.204167337f;
    } else if (id > (double)31622.777f /* 10.0**4.5 */) {
        max_temp = ((double)3.38989f * vs * vs) + ((double)16.6519f * vs) + (double)96.569f;
        max_temp = (double)0.5f * max_temp;
    } else {
        max_temp = ((double)0.47258f * vs * vs) + ((double)40.44161f * vs) - (double)128.635455216f;
    }
    st.cs_max_temp = max_temp;
    st.cs_dlength = (double)12.0f * C_PC * vs / id;
    st.cs_dissipation_time = (st.cs_dlength * 1.0e-5 / vs) / C_SPY;
    st.cs_z2 = st.cs_dlength / coshinv_f_dev(0.01f);
    st.cs_z1 = st.cs_z2 / (double)4.5f;
    double zmax = st.cs_dlength / coshinv_f_dev(0.15f);
    st.cs_z3 = zmax / 6;
    st.cs_at = (1 / zmax) * pow((max_temp - p[UCL_P_INITIALTEMP]) * (exp(6.0) - (double)1.f), (double)(1.f / 6.f));
    double bm0 = p[UCL_P_BM0] * 1e-06;
    double va = bm0 / sqrt(4 * C_PI * C_MH);
    va = va / 1.e5;
    double v0 = (double)2.f, v01 = 0;
    int guard = 0;
    while (fabs(v0 - v01) >= (double)1e-6f && guard++ < 100000) {
        v01 = v0;
        double g1 = -(va * va * vs * vs) / 2;
        double g2 = v01 * v01 - v01 * vs - va * va / 2;
        v0 = sqrt(g1 / g2);
    }
    st.cs_v0 = v0;
    return 0;
}

__device__ __noinline__ void cshock_update_physics_dev(Scalars &st)
{
    // shst cshock.f90:225-252
    const double KM = 1.e5;
    double vs = st.vs, v0 = st.cs_v0;
    double vn1 = 1e30, vn = st.cs_vn0, zn = st.cs_zn0;
    int loop = 0;
    while (fabs(vn - vn1) >= (double)1.e-10f && loop < 100) {
        vn1 = vn;
        double f1 = vs - vn1, f0 = vs - st.cs_vn0;
        zn = st.cs_zn0 + (st.current_time - st.current_time_old) * KM * (f1 + f0) / 2;
        double xcos = zn / st.cs_z2;
        double ach = 0.5 * (exp(xcos) + exp(-xcos));
        vn = (vs - v0) - ((vs - v0) / ach);
        loop++;
    }
    double xcos = zn / st.cs_z1;
    double ach = 0.5 * (exp(xcos) + exp(-xcos));
    double vi = (vs - v0) - ((vs - v0) / ach);
    st.cs_drift_vel = vi - vn;
    st.cs_zn0 = zn;
    st.cs_vn0 = vn;
    st.cs_zn = zn;
    st.cs_vn = vn;
    // updatePhysics cshock.f90:159-208
    if (st.time_in_years > 0.0) st.density = st.p[UCL_P_INITIALDENS] * vs / (vs - vn);
    if (st.time_in_years > 0.0)
        st.gastemp = st.p[UCL_P_INITIALTEMP] + (pow(st.cs_at * zn, 6.0)) / (exp(zn / st.cs_z3) - 1);
    bool post = st.time_in_years > st.cs_dissipation_time;
    if (st.gastemp < st.min_postshock_temp && post) st.gastemp = st.min_postshock_temp;
    st.dusttemp = st.gastemp;
}

// ---- jshock.f90 (James et al. 2020 J-shock parameterisation), thread 0 ------------------------------------
__device__ __noinline__ int jshock_initialize_dev(Scalars &st)
{
    // jshock.f90:29-78
    double *p = st.p;
    st.vs = p[UCL_P_VS];
    st.cloudsize = (p[UCL_P_ROUT] - p[UCL_P_RIN]) * C_PC;
    if (p[UCL_P_FREEFALL] != 0.0) p[UCL_P_FREEFALL] = 0.0;
    if (p[UCL_P_POINTS] > 1) return -1;
    const double vs = st.vs, id = p[UCL_P_INITIALDENS];
    st.density = id;
    st.js_max_temp = (double)5e3f * ((vs / 10) * (vs / 10));
    st.current_time_old = 0.0;
    const double v2 = vs * vs;
    const double poly = (double)-2.058e-07f * (v2 * v2) + (double)3.844e-05f * (v2 * vs) - (double)0.002478f * v2 +
                        (double)0.06183f * vs - (double)0.4254f;
    st.js_vmin = pow(poly * poly, (double)0.5f);
    // mean free path / shock width: the literals of the product are single precision, pi is a double parameter
    const double inner = (double)(sqrtf(2.0f) * 1e3f) * (C_PI * (double)(2.4e-8f * 2.4e-8f));
    st.js_tshock = ((1.0 / inner) / 1e4) / (vs * 1e5);
    st.js_tcool = (1 / id) * 1e6 * (60 * 60 * 24 * 365);
    st.js_max_dens = vs * id * 1e2;
    st.js_t_lambda = log(st.js_max_temp / p[UCL_P_INITIALTEMP]);
    st.js_n_lambda = log(st.js_max_dens / id);
    st.js_v0 = 0.0;
    return 0;
}

__device__ __noinline__ void jshock_update_physics_dev(Scalars &st)
{
    // jshock.f90:102-135
    const double *p = st.p;
    const double ct = st.current_time, id = p[UCL_P_INITIALDENS];
    double v0 = st.vs * exp(log(st.js_vmin / st.vs) * (ct / (p[UCL_P_FINALTIME] * 60 * 60 * 24 * 365)));
    if (v0 < st.js_vmin) v0 = st.js_vmin;
    st.js_v0 = v0;
    double tn;
    if (ct <= st.js_tshock) {
        const double x = ct / st.js_tshock;
        tn = (x * x) * st.js_max_temp + p[UCL_P_INITIALTEMP];
        st.density = (x * x * x) * (4 * id);
        if (st.density < id) st.density = id;
    } else if (ct <= st.js_tcool) {
        tn = st.js_max_temp * exp(-st.js_t_lambda * (ct / st.js_tcool));
        st.density = (4 * id) * exp(st.js_n_lambda * (ct / st.js_tcool));
        if (tn <= 10) tn = 10;
        if (st.density > st.js_max_dens) st.density = st.js_max_dens;
    } else {
        tn = 10;
        st.density = st.js_max_dens;
    }
    st.gastemp = tn;
    st.dusttemp = tn;
}

__device__ __noinline__ int initialize_physics_dev(Scalars &st)
{
    double *p = st.p;
    // coreInitializePhysics physics-core.f90:42-73
    st.time_in_years = st.current_time / C_SPY;
    st.cloudsize = (p[UCL_P_ROUT] - p[UCL_P_RIN]) * C_PC;
    st.gastemp = p[UCL_P_INITIALTEMP];
    st.dusttemp = st.gastemp;
    st.density = p[UCL_P_INITIALDENS];
    st.current_time_old = 0.0;
    st.radfield = p[UCL_P_RADFIELD];
    st.zeta = p[UCL_P_ZETA];
    st.h2crprate = 0.0;
    if (p[UCL_P_IMPROVEDH2CRPDISSOCIATION] != 0.0 && p[UCL_P_COSMICRAYATTENUATION] == 0.0) return -1;
    st.coldens = (double)1.0f * st.cloudsize / (double)1.0f * p[UCL_P_INITIALDENS];
    st.av = p[UCL_P_BASEAV] + st.coldens / 1.6e21;
    st.zetascale = st.zeta;
    switch (st.kind) {
    case UCLGPU_CLOUD: // cloud.f90:20-31
        if (p[UCL_P_FREEFALL] != 0.0) st.density = (double)1.001f * p[UCL_P_INITIALDENS];
        return 0;
    case UCLGPU_HOTCORE: // hotcore.f90:30-50
        st.temp_indx = (int)p[UCL_P_TEMPINDX];
        st.max_temp = p[UCL_P_MAXTEMP];
        if (p[UCL_P_FREEFALL] != 0.0) st.density = (double)1.001f * p[UCL_P_INITIALDENS];
        if (st.temp_indx > 6 || st.temp_indx < 1) return -1;
        return 0;
    case UCLGPU_CSHOCK:
        return cshock_initialize_dev(st);
    case UCLGPU_COLLAPSE: // scalar part; the enclosed-mass quadrature follows on the whole CTA (run_cell)
        return collapse_initialize_t0(st);
    case UCLGPU_JSHOCK:
        return jshock_initialize_dev(st);
    case UCLGPU_POSTPROCESS: { // postprocess.f90:24-92
        if (!st.pp_grid || st.pp_ntime < 1) return -1;
        const double *g = st.pp_grid;
        const int n = st.pp_ntime;
        if (st.pp_coldens) st.cloudsize = (double)0.f; // shielding column densities supplied separately
        st.p[UCL_P_ENDATFINALDENSITY] = 0.0;
        st.p[UCL_P_FREEFALL] = 0.0;
        st.pp_tstep = 1;
        st.target_time = g[0];
        st.density = g[1 * n];
        st.gastemp = g[2 * n];
        st.dusttemp = g[3 * n];
        st.radfield = g[4 * n];
        st.zeta = g[5 * n];
        if (st.pp_coldens) {
            st.coldens = g[6 * n];
            st.av = (double)5.348e-22f * st.coldens;
        }
        st.p[UCL_P_FINALTIME] = g[n - 1] / C_SPY;
        return 0;
    }
    }
    return -1;
}

__device__ void update_target_time_dev(Scalars &st)
{
    const double t = st.time_in_years;
    switch (st.kind) {
    case UCLGPU_CLOUD: // cloud.f90:39-51
        if (t >= 1.0e6) st.target_time = (t + 1.0e5) * C_SPY;
        else if (t > 10.0) {
            // orderMagnitude = 10**FLOOR(LOG10(t)): built from exact powers of ten so that the
            // cadence does not depend on the last bit of the device log10/pow
            double om = 10.0;
            while (om * 10.0 <= t) om *= 10.0;
            st.target_time = ((floor(t / om) + 1.0) * om) * C_SPY;
        } else if (t > 0.0) st.target_time = 10 * t * C_SPY;
        else st.target_time = C_SPY * 1.0e-7;
        break;
    case UCLGPU_HOTCORE: // hotcore.f90:56-72
        if (t > 1.0e6) st.target_time = (t + 1.0e5) * C_SPY;
        else if (t > 1.0e5) st.target_time = (t + 1.0e4) * C_SPY;
        else if (t > 1.0e4) st.target_time = (t + 1000.0) * C_SPY;
        else if (t > 1000) st.target_time = (t + 100.0) * C_SPY;
        else if (t > 100) st.target_time = (t + 10.0) * C_SPY;
        else if (t > 0.0) st.target_time = (t * 10.0) * C_SPY;
        else st.target_time = C_SPY * 1.0e-7;
        break;
    case UCLGPU_CSHOCK: // cshock.f90:149-156
        if (t < 2.0 * st.cs_dissipation_time)
            st.target_time = (t + st.timestep_factor * st.cs_dissipation_time) * C_SPY;
        else st.target_time = ((double)1.1f * t) * C_SPY;
        break;
    case UCLGPU_COLLAPSE:
        collapse_target_time_dev(st);
        break;
    case UCLGPU_JSHOCK: // jshock.f90:85-97
        if (t > (double)1e6f) st.target_time = (t + (double)1e5f) * C_SPY;
        else if (t > 1.0e4) st.target_time = (t + 1000) * C_SPY;
        else if (t > 1.0e3) st.target_time = (t + (double)100.f) * C_SPY;
        else if (t * C_SPY < st.js_tshock) st.target_time = st.current_time + (double)0.05f * st.js_tshock;
        else st.target_time = (double)1.1f * st.current_time;
        break;
    case UCLGPU_POSTPROCESS: // postprocess.f90:100-106
        if (st.pp_tstep > st.pp_ntime) st.pp_tstep = st.pp_ntime; // guard: the Fortran would read past the array
        st.target_time = st.pp_grid[st.pp_tstep - 1] + (double)1.f * C_SPY;
        break;
    }
}

__device__ void update_physics_dev(Scalars &st)
{
    // coreUpdatePhysics physics-core.f90:75-88 (points = 1)
    st.coldens = st.cloudsize / (double)1.0f * st.density;
    st.av = st.p[UCL_P_BASEAV] + st.coldens / 1.6e21;
    st.dusttemp = st.gastemp;
    if (st.p[UCL_P_COSMICRAYATTENUATION] != 0.0) ionization_dependency_dev(st);
    switch (st.kind) {
    case UCLGPU_HOTCORE: { // hotcore.f90:78-90
        const double tempa[6] = {1.927e-1, 4.8560e-2, 7.8470e-3, 9.6966e-4, 1.706e-4, 4.74e-7};
        const float tempb[6] = {0.5339f, 0.6255f, 0.8395f, 1.085f, 1.289f, 1.98f};
        if (st.gastemp < st.max_temp) {
            double g = (st.cloudsize / (st.p[UCL_P_ROUT] * C_PC)) * (double)(1.0f / 1.0f);
            g = pow(g, -0.5);
            int k = st.temp_indx - 1;
            g = st.p[UCL_P_INITIALTEMP] + ((tempa[k] * pow(st.current_time / C_SPY, (double)tempb[k])) * g);
            if (g > st.max_temp) g = st.max_temp;
            st.gastemp = g;
        }
        st.dusttemp = st.gastemp;
        break;
    }
    case UCLGPU_CSHOCK:
        cshock_update_physics_dev(st);
        break;
    case UCLGPU_JSHOCK:
        jshock_update_physics_dev(st);
        break;
    case UCLGPU_POSTPROCESS: { // postprocess.f90:112-129
        const double *g = st.pp_grid;
        const int n = st.pp_ntime, k = st.pp_tstep - 1;
        st.target_time = g[k];
        st.density = g[1 * n + k];
        st.gastemp = g[2 * n + k];
        st.dusttemp = g[3 * n + k];
        st.radfield = g[4 * n + k];
        st.zeta = g[5 * n + k];
        if (st.pp_coldens) {
            st.coldens = g[6 * n + k];
            st.av = (double)5.348e-22f * st.coldens;
        }
        st.pp_tstep = st.pp_tstep + 1;
        break;
    }
    default:
        break;
    }
}

// ---- sputtering.f90 -----------------------------------------------------------------------------
struct Sput {
    double sconst, eta, epso;
};
__device__ __forceinline__ double ice_yield_integrand_dev(const Sput &q, double x, double pmass, double gastemp)
{
    const double ebind = (double)0.53f * 1.6e-12;
    double sv = q.sconst * sqrt(pmass);
    double eps = (x * x) * C_KBOLTZ * gastemp;
    eps = q.eta * eps / ebind;
    double d = eps - q.epso;
    double yield = 8.3e-4 * (d * d) / ((double)1.f + pow(eps / (double)30.f, (double)1.3333f));
    return yield * (x * x) * (exp(-((x - sv) * (x - sv))) - exp(-((x + sv) * (x + sv))));
}

// iceYieldRate sputtering.f90:117-149; the trapezoid stages are summed by the whole block
__device__ __noinline__ double ice_yield_rate_dev(Smem &s, Blk &b, Sput &q, double pmass, double pdens, double gastemp)
{
    const double ebind = (double)0.53f * 1.6e-12;
    const double target_mass = (double)18.0f * C_MH;
    q.eta = (double)4.f * (double)0.8f * pmass * target_mass * pow(pmass + target_mass, -2.0);
    q.epso = fmax((double)1.f, (double)4.f * q.eta);
    double sv = q.sconst * sqrt(pmass);
    double lower = sqrt(q.epso * ebind / (q.eta * C_KBOLTZ * gastemp));
    int i = 1;
    double upper = lower + (1e3 - lower) * 0.5;
    while (ice_yield_integrand_dev(q, upper, pmass, gastemp) < 1e-200 && (upper - lower) > 1.0e-3) {
        i++;
        upper = lower + (1e3 - lower) * pow(0.5, (double)i);
    }
    if (!((upper - lower) > 1e-4)) return 0.0;
    // trapezoidIntegrate / trapzd sputtering.f90:186-235
    const double tol = (double)1.e-3f;
    double val = 0.0, olds = (double)-1.e30f;
    for (int j = 1; j <= 25; j++) {
        if (j == 1) {
            val = (double)0.5f * (upper - lower) *
                  (ice_yield_integrand_dev(q, lower, pmass, gastemp) + ice_yield_integrand_dev(q, upper, pmass, gastemp));
        } else {
            long long it = 1LL << (j - 2);
            double tnm = (double)it;
            double del = (upper - lower) / tnm;
            double x0 = lower + (double)0.5f * del;
            double part = 0.0;
            for (long long k = threadIdx.x; k < it; k += NT) part += ice_yield_integrand_dev(q, x0 + (double)k * del, pmass, gastemp);
            double sum = block_sum(s, b, part);
            val = (double)0.5f * (val + (upper - lower) * sum / tnm);
        }
        if (fabs(val - olds) <= tol * fabs(olds)) break;
        olds = val;
    }
    double r = val / sv;
    r = r * 1.e-5 * 1.e-5 * sqrt(8.0 * C_KBOLTZ * gastemp * C_PI / pmass);
    return r * pdens;
}

// shock sublimation -> sputterIces (cshock.f90:211-219 with the drift velocity, jshock.f90:143-152 with the shock
// velocity v0), sputtering.f90:65-112
__device__ __noinline__ void shock_sublimation_dev(Smem &s, Blk &b, double shockvel)
{
    Scalars &st = s.st;
    const int tid = threadIdx.x;
    double t = 0.0;
    if (tid < NICE) t = s.abund[net_ice_list[tid]];
    double total = block_sum(s, b, t);
    if (total > 1e-25 && shockvel > 0) {
        const double gastemp = st.gastemp, density = st.density;
        Sput q;
        q.sconst = sqrt((shockvel * shockvel * 1.e5 * 1.e5) / (2.0 * gastemp * C_KBOLTZ));
        const int proj[6] = {NET_NH2, NET_NHE, NET_NC, NET_NO, NET_NSI, NET_NCO};
        double rate = 0.0;
        for (int k = 0; k < 6; k++)
            rate = rate + ice_yield_rate_dev(s, b, q, net_mass[proj[k]] * C_MH, density * s.abund[proj[k]], gastemp);
        rate = rate * (density / NET_GDR);
        double frac = rate * (st.current_time - st.current_time_old);
        frac = frac / total;
        if (frac > 1.0) frac = 1.0;
        if (frac < 0.0) frac = 0.0;
        const bool all = shockvel >= (double)19.0f;
        // many-one vector subscript on the left-hand side: the LAST store per gas species wins
        // (sputtering.f90:104-110).  Surface entries come first in iceList, bulk entries last.
        double moved = 0.0;
        int gas = -1;
        bool use = false;
        BLOCK_SYNC();
        if (tid < NICE) {
            int ice = net_ice_list[tid];
            gas = net_gas_ice_list[tid];
            use = all || !net_is_refractory_ice[tid];
            moved = s.abund[gas] + frac * s.abund[ice];
        }
        BLOCK_SYNC();
        if (tid < NICE / 2 && use) s.abund[gas] = moved; // surface partners
        BLOCK_SYNC();
        if (tid >= NICE / 2 && tid < NICE && use) s.abund[gas] = moved; // bulk partners overwrite
        BLOCK_SYNC();
        if (tid < NICE && use) {
            int ice = net_ice_list[tid];
            s.abund[ice] = s.abund[ice] - frac * s.abund[ice];
        }
    }
    BLOCK_SYNC();
    if (tid < NEQ && s.abund[tid] < 1.0e-50) s.abund[tid] = 0.0;
    BLOCK_SYNC();
}

// ---- chemistry -------------------------------------------------------------------------------------
// initializeChemistry chemistry.f90:57-105.  Absent elements carry index NSPEC (the density
// slot): they are written there and then overwritten.  Thread 0, after the fill.
__device__ __noinline__ void initialize_abundances_t0(Smem &s)
{
    const double *p = s.st.p;
    double *a = s.abund;
    a[NET_NO] = p[UCL_P_FO];
    a[NET_NN] = p[UCL_P_FN];
    a[NET_NMG] = p[UCL_P_FMG];
    a[NET_NP] = p[UCL_P_FP];
    a[NET_NF] = p[UCL_P_FF];
    a[NET_NNA] = p[UCL_P_FNA];
    a[NET_NLI] = p[UCL_P_FLI];
    a[NET_NPAH] = p[UCL_P_FPAH];
    a[NET_NSX] = p[UCL_P_FS];
    a[NET_NSIX] = p[UCL_P_FSI];
    a[NET_NCLX] = p[UCL_P_FCL];
    switch ((int)p[UCL_P_ION]) {
    case 0: a[NET_NC] = p[UCL_P_FC]; a[NET_NCX] = 1.e-10; break;
    case 1: a[NET_NC] = p[UCL_P_FC] * (double)0.5f; a[NET_NCX] = p[UCL_P_FC] * (double)0.5f; break;
    case 2: a[NET_NC] = 1.e-10; a[NET_NCX] = p[UCL_P_FC]; break;
    }
    a[NET_N18O] = p[UCL_P_F18O];
    a[NET_N15N] = p[UCL_P_F15N];
    a[NET_N13C] = p[UCL_P_F13C];
    a[NET_NELEC] = a[NET_NCX] + a[NET_NSIX] + a[NET_NSX] + a[NET_NCLX] + a[NET_NMGX];
    for (int i = 0; i < NEQ; i++) a[i] = a[i] * p[UCL_P_METALLICITY];
    a[NET_NH] = p[UCL_P_FH];
    a[NET_NH2] = (double)0.5f * ((double)1.0f - p[UCL_P_FH]);
    a[NET_ND] = p[UCL_P_FD];
    a[NET_NHE] = p[UCL_P_FHE];
    a[NEQ - 1] = s.st.density;
}

// per-try set-up of updateChemistry chemistry.f90:166-203
__device__ void chemistry_setup_dev(Smem &s)
{
    Scalars &st = s.st;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    BLOCK_SYNC();
    if (warp == 0) {
        double sb = 0.0;
        for (int k = lane; k < NSURF; k += 32) sb += s.abund[net_bulk_list[k]];
        sb = warp_sum(sb);
        if (lane == 0) s.abund[NET_IB] = sb;
    } else if (warp == 1) {
        double ss = 0.0;
        for (int k = lane; k < NSURF; k += 32) ss += s.abund[net_surface_list[k]];
        ss = warp_sum(ss);
        if (lane == 0) s.abund[NET_IS] = ss;
    }
    T0_BEGIN
    double *a = s.abund;
    if (st.p[UCL_P_FREEFALL] == 0.0) a[NEQ - 1] = st.density;
    double cs = st.cloudsize / (double)1.0f;
    st.h2col = 0.0 + (double)0.5f * a[NET_NH2] * st.density * cs;
    st.cocol = 0.0 + (double)0.5f * a[NET_NCO] * st.density * cs;
    st.ccol = 0.0 + (double)0.5f * a[NET_NC] * st.density * cs;
    if (st.pp_coldens) { // chemistry.f90:183-189: postprocessed tracers have column densities provided
        const int n = st.pp_ntime, k = st.pp_tstep - 1;
        st.h2col = st.pp_grid[7 * n + k];
        st.cocol = st.pp_grid[8 * n + k];
        st.ccol = st.pp_grid[6 * n + k] * a[NET_NC];
    }
    st.safe_mantle = fmax(1e-30, a[NET_IS]);
    st.safe_bulk = fmax(1e-30, a[NET_IB]);
    st.blr = fmin(1.0, NET_NSITES / (NET_GDR * st.safe_bulk));
    T0_END
    calc_rates(s);
}

// updateChemistry chemistry.f90:145-239.  Returns the successFlag (block-uniform).
__device__ int update_chemistry_dev(Smem &s, Blk &b)
{
    Scalars &st = s.st;
    const int tid = threadIdx.x;
    int loop = 0;
    BLOCK_SYNC();
    const double original_target = st.target_time;
    while (st.current_time < st.target_time && loop < 10) {
        chemistry_setup_dev(s);
        if (tid < NEQ) {
            double v = st.abstol_factor * s.abund[tid];
            if (v < st.p[UCL_P_ABSTOL_MIN]) v = st.p[UCL_P_ABSTOL_MIN];
            s.atol[tid] = v;
        }
        BLOCK_SYNC();
        const bool warm = st.use_tcrit && st.hist_valid;
        int istate = bdf_integrate(s, b, st.target_time, warm);
        if (istate == -3) return UCLGPU_INT_UNRECOVERABLE_ERROR;
        if (st.step_budget > 0 && st.nst + st.netf + st.ncfn > st.step_budget) return UCLGPU_INT_TOO_MANY_FAILS_ERROR;
        // integrateODESystem chemistry.f90:256-291
        if (st.p[UCL_P_ENFORCECHARGECONSERVATION] != 0.0) {
            double q = 0.0;
            if (tid < NSPEC && net_is_ion[tid]) q = s.abund[tid];
            q = block_sum(s, b, q);
            if (tid == 0) s.abund[NET_NELEC] = q;
        }
        T0_BEGIN
        st.hist_valid = (istate == 2) ? 1 : 0; // any failed call is followed by a cold restart (ISTATE=1)
        if (istate < 0) st.nfailcall++;
        switch (istate) {
        case -1: case -4: case -5: break; // the shortened target is overwritten just below (Q3)
        case -2: st.abstol_factor = st.abstol_factor * (double)10.0f; break;
        default: st.mxstep = 10000;
        }
        T0_END
        if (tid < NEQ && s.abund[tid] < C_MIN_ABUND) s.abund[tid] = C_MIN_ABUND;
        T0_BEGIN
        st.density = s.abund[NEQ - 1];
        st.target_time = original_target;
        T0_END
        loop++;
    }
    return 0;
}

__device__ __noinline__ void output_row_dev(Smem &s, const RunArgs &a, long long cell, int dtime)
{
    // io.f90:59-98; dtime is 1-based
    const Scalars &st = s.st;
    const int tid = threadIdx.x;
    const size_t row = (size_t)cell * (a.timepoints + 1) + (dtime - 1);
    if (a.phys_traj && tid == 0) {
        double *r = a.phys_traj + row * UCLGPU_NPHYS;
        r[0] = st.time_in_years; r[1] = st.density; r[2] = st.gastemp; r[3] = st.dusttemp;
        r[4] = st.av; r[5] = st.radfield; r[6] = st.zeta; r[7] = 1.0;
    }
    if (a.chem_traj)
        for (int i = tid; i < NSPEC; i += NT) a.chem_traj[row * NSPEC + i] = s.abund[i];
    if (a.rates_traj)
        for (int i = tid; i < NREAC; i += NT) a.rates_traj[row * NREAC + i] = s.rate[i];
}

// solveAbundances wrap.f90:549-697 for one cell
// (inputs are read at index `cell`, results are written at index `out`)
__device__ void run_cell(Smem &s, Blk &b, const RunArgs &a, long long cell, long long out)
{
    Scalars &st = s.st;
    const int tid = threadIdx.x;
    BLOCK_SYNC();
    if (tid < UCLGPU_NPARAM) st.p[tid] = a.params[(size_t)tid * a.ncell + cell];
    for (int i = tid; i < NREAC; i += NT) s.rate[i] = 0.0;
    T0_BEGIN
    st.kind = a.kind;
    st.current_time = 0.0;
    st.time_in_years = 0.0;
    st.phi = st.p[UCL_P_PHI];
    st.abstol_factor = st.p[UCL_P_ABSTOL_FACTOR];
    st.mxstep = (int)st.p[UCL_P_MXSTEP];
    st.step_budget = a.max_steps;
    st.pp_ntime = a.pp_ntime;
    st.pp_coldens = (a.kind == UCLGPU_POSTPROCESS) ? a.pp_coldens : 0;
    st.pp_grid = a.pp_grid ? a.pp_grid + (size_t)cell * 10 * a.pp_ntime : nullptr;
    st.pp_tstep = 1;
    st.c_alpha = a.coef ? a.coef : net_alpha;
    st.c_beta = a.coef ? a.coef + NREAC : net_beta;
    st.c_gama = a.coef ? a.coef + 2 * NREAC : net_gama;
    st.rtol = st.p[UCL_P_RELTOL];
    st.last_temp = 99.0e99;
    st.nst = st.nfe = st.nje = st.nlu = st.nni = st.ncfn = st.netf = st.nintervals = 0;
    st.nsing = st.nmaxcor = st.ndiverge = st.nfailcall = 0;
    st.hist_valid = 0;
    st.use_tcrit = (a.kind == UCLGPU_CLOUD && a.warm_restart) ? 1 : 0; // experimental, off by default
    st.cyc_rates = st.cyc_rhs = st.cyc_jac = st.cyc_factor = st.cyc_dense = st.cyc_solve = 0;
    st.cyc_total = clock64();
    st.flag = initialize_physics_dev(st);
    T0_END
    if (a.kind == UCLGPU_COLLAPSE && st.flag == 0) collapse_initialize_dev(s, b);
    int flag = 0;
    int dtime = 1;
    const bool want_traj = a.phys_traj || a.chem_traj || a.rates_traj;
    if (st.flag != 0) {
        flag = UCLGPU_PHYSICS_INIT_ERROR;
    } else {
        if (tid < NEQ) s.abund[tid] = C_MIN_ABUND;
        T0_BEGIN
        initialize_abundances_t0(s);
        T0_END
        if (a.y0 && tid < NSPEC) s.abund[tid] = a.y0[(size_t)(a.y0_index ? a.y0_index[cell] : cell) * NEQ + tid];
        BLOCK_SYNC();
        if (want_traj) {
            if (dtime > a.timepoints + 1) flag = UCLGPU_NOT_ENOUGH_TIMEPOINTS_ERROR;
            else output_row_dev(s, a, out, dtime);
        }
        for (;;) {
            BLOCK_SYNC();
            const double *p = st.p;
            bool go = flag == 0 && ((p[UCL_P_ENDATFINALDENSITY] != 0.0 && st.density < p[UCL_P_FINALDENS]) ||
                                    (p[UCL_P_ENDATFINALDENSITY] == 0.0 && st.time_in_years < p[UCL_P_FINALTIME]));
            if (!go) break;
            dtime++;
            T0_BEGIN
            st.current_time_old = st.current_time;
            st.time_in_years = st.current_time / C_SPY;
            update_target_time_dev(st);
            T0_END
            flag = update_chemistry_dev(s, b);
            if (flag < 0) break;
            T0_BEGIN
            st.nintervals++;
            st.time_in_years = st.target_time / C_SPY;
            update_physics_dev(st);
            T0_END
            if (st.kind == UCLGPU_COLLAPSE) collapse_update_physics_dev(s, b); // modelUpdatePhysics on the whole CTA
            if (st.kind == UCLGPU_CSHOCK) shock_sublimation_dev(s, b, st.cs_drift_vel);
            if (st.kind == UCLGPU_JSHOCK) shock_sublimation_dev(s, b, st.js_v0);
            if (want_traj) {
                if (dtime > a.timepoints + 1) flag = UCLGPU_NOT_ENOUGH_TIMEPOINTS_ERROR;
                else output_row_dev(s, a, out, dtime);
            }
        }
    }
    BLOCK_SYNC();
    for (int i = tid; i < NEQ; i += NT) a.y_final[(size_t)out * NEQ + i] = s.abund[i];
    if (tid == 0) {
        a.flag[out] = flag;
        if (a.phys_final) {
            double *r = a.phys_final + (size_t)out * UCLGPU_NPHYS;
            r[0] = st.time_in_years; r[1] = st.density; r[2] = st.gastemp; r[3] = st.dusttemp;
            r[4] = st.av; r[5] = st.radfield; r[6] = st.zeta; r[7] = 1.0;
        }
        if (a.tdiss) a.tdiss[out] = (st.kind == UCLGPU_CSHOCK) ? st.cs_dissipation_time : 0.0;
        if (a.stats) {
            uclgpu_stats &o = a.stats[out];
            o.nst = st.nst; o.nfe = st.nfe; o.nje = st.nje; o.nlu = st.nlu; o.nni = st.nni;
            o.ncfn = st.ncfn; o.netf = st.netf; o.nintervals = st.nintervals;
            o.nsing = st.nsing; o.nmaxcor = st.nmaxcor; o.ndiverge = st.ndiverge; o.nfailcall = st.nfailcall;
            o.cyc_rates = st.cyc_rates; o.cyc_rhs = st.cyc_rhs; o.cyc_jac = st.cyc_jac; o.cyc_factor = st.cyc_factor;
            o.cyc_dense = st.cyc_dense; o.cyc_solve = st.cyc_solve; o.cyc_total = clock64() - st.cyc_total;
            o.reserved = 0;
        }
    }
    BLOCK_SYNC();
}
