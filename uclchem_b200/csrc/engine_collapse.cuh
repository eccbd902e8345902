// engine_collapse.cuh -- the collapse model (collapse.f90, Priestley et al. 2018) for one cell.
//
//   initializePhysics collapse.f90:28-61   -> collapse_initialize_dev()
//   updateTargetTime  collapse.f90:63-75   -> collapse_target_time_dev()
//   updatePhysics     collapse.f90:79-106  -> collapse_update_physics_dev()
//   rhofit ... avfit  collapse.f90:178-383 -> col_*()
//
// A model is one parcel at r = rout (points = 1) that follows a fitted density profile rho(r, t) of an MHD
// collapse simulation.  The reference evaluates three trapezoid quadratures of that profile with 10^3-10^4
// panels on every output interval (column density to the cloud edge, enclosed mass, radius enclosing a mass);
// here the panels are dealt over the threads of the cell's CTA and reduced with block sums, and the
// "first panel whose running mass reaches M" search is a chunked scan.  Single-precision literals of the
// Fortran source are kept as (double)x.yf.
#pragma once
#include "engine_core.cuh"

#define CF(x) ((double)(x##f))
#define C_AU 2.063e5 /* constants.f90:12 */

struct ColFit {
    double rho0, r0, a;
};

__device__ __forceinline__ double col_unit_time_yr() { return pow(2 * C_PI * 6.67e-8 * 2.2e4 * C_MH, -0.5) / C_SPY; }
__device__ __forceinline__ double col_unit_radius_pc()
{
    return sqrt(1.38e-16 * 10 / 2 / C_MH) * pow(2 * C_PI * 6.67e-8 * 2.2e4 * C_MH, -0.5) / C_PC;
}

// density profile, collapse.f90:178-199
__device__ __forceinline__ double col_rhofit(int mode, double r, const ColFit &f)
{
    if (mode <= 2) return f.rho0 / (1 + pow(r * C_AU / f.r0, f.a));
    if (mode == 3) {
        const double x = r / col_unit_radius_pc() / f.r0;
        return 2.2e4 * f.rho0 / pow(1 + x * x, f.a);
    }
    return f.rho0 / (1 + pow(r / 7.5e-1 / f.r0, f.a));
}

// central density, radius parameter and slope of the profile at time t (years), collapse.f90:202-266
__device__ __noinline__ ColFit col_fit(const Scalars &st, double t)
{
    ColFit f;
    const double rem = st.col_max_time - t;
    switch (st.col_mode) {
    case 1:
        f.rho0 = pow(10.0, CF(61.8) * pow(rem, CF(-0.01)) - CF(49.4));
        f.r0 = pow(10.0, CF(-28.5) * pow(rem, CF(-0.01)) + CF(28.93));
        f.a = 2.4;
        break;
    case 2:
        f.rho0 = pow(10.0, CF(68.4) * pow(rem, CF(-0.01)) - CF(55.7));
        f.r0 = pow(10.0, CF(-39.0) * pow(rem, CF(-0.01)) + CF(38.7));
        f.a = CF(1.9) + CF(0.5) * exp(-t / CF(1e5));
        break;
    case 3: {
        const double tn = t / col_unit_time_yr();
        f.rho0 = pow(10.0, CF(3.54) * pow(CF(5.47) - tn, CF(-0.15)) - CF(2.73));
        f.r0 = pow(10.0, CF(-1.34) * pow(CF(5.47) - tn, CF(-0.15)) + CF(1.47));
        f.a = CF(2.0) - CF(0.5) * pow(tn / CF(5.47), 9.0);
        break;
    }
    default: {
        const double t6 = 1e-6 * t;
        f.rho0 = (t <= 6.0) ? 2.0e3 + 1.7e3 * (t / CF(6.0) - CF(1.0)) : pow(10.0, CF(5.3) * pow(CF(16.138) - t6, CF(-0.1)) - CF(1.0));
        f.r0 = pow(10.0, CF(-2.57) * pow(CF(16.138) - t6, CF(-0.1)) + CF(1.85));
        f.a = CF(2.4) - CF(0.2) * pow(t6 / CF(16.138), 40.0);
    }
    }
    return f;
}

// radial velocity of the filament / ambipolar fits at radius r (pc), cm/s; collapse.f90:269-382
__device__ __noinline__ double col_velocity(const Scalars &st, double r, double t)
{
    if (st.col_mode == 3) {
        const double tn = t / col_unit_time_yr();
        double rmin, vmin, av;
        if (tn == 0.0) { rmin = 7.2; vmin = 0.0; av = 0.4; }
        else {
            const double lt = log(tn);
            if (lt < 1.6) { rmin = CF(-1.149) * tn + CF(7.2); vmin = CF(0.0891) * tn; av = CF(0.0101) * tn + CF(0.4); }
            else if (lt < 1.674) { rmin = CF(-9.2) * lt + CF(16.25); vmin = CF(5.5) * lt - CF(8.37); av = CF(0.695) * lt - CF(0.663); }
            else { rmin = CF(-22.0) * lt + CF(37.65); vmin = CF(18.9) * lt - CF(30.8); av = CF(2.69) * lt - CF(4.0); }
        }
        const double nr = r / col_unit_radius_pc() - rmin;
        const double x = nr / rmin;
        const double v = (nr < 0.0) ? vmin * (x * x - 1) : vmin * (exp(-2.0 * av * nr) - 2 * exp(-av * nr));
        return sqrt(1.38e-16 * 10 / 2 / C_MH) * v;
    }
    const double t6 = 1e-6 * t;
    double rmin;
    if (t6 <= CF(10.2)) rmin = CF(-0.0039) * t6 + CF(0.49);
    else if (t6 <= CF(15.1)) rmin = CF(-0.0306) * (t6 - CF(10.2)) + CF(0.45);
    else rmin = CF(-0.282) * (t6 - CF(15.1)) + CF(0.3);
    const double vmin = CF(3.44) * pow(CF(16.138) - t6, CF(-0.35)) - CF(0.7);
    const double av = (t6 <= CF(10.2)) ? CF(0.143) * t6 : CF(0.217) * (t6 - CF(10.2)) + CF(1.46);
    const double rmid = CF(0.5), r75 = r / 7.5e-1, nr = r75 - rmin;
    double v;
    if (r75 < rmin) { const double x = nr / rmin; v = vmin * (x * x - 1); }
    else if (r75 <= rmid) v = (vmin - av) * pow(nr / (rmid - rmin), CF(0.3)) - vmin;
    else v = av / (CF(1.0) - rmid) * (r75 - rmid) - av;
    return 1e3 * v;
}

// mass of panel i (1-based) of the enclosed-mass quadrature with step dr, collapse.f90:129-131,147-148
__device__ __forceinline__ double col_mass_panel(int mode, int i, double dr, const ColFit &f)
{
    const double drho = 0.5 * (col_rhofit(mode, i * dr, f) + col_rhofit(mode, (i - 1) * dr, f));
    const double r = i * dr;
    return drho * dr * (r * r);
}

// thread 0: mode, time scale, end time, parcel radius (collapse.f90:34-51).  Returns -1 for a bad mode.
__device__ __noinline__ int collapse_initialize_t0(Scalars &st)
{
    double *p = st.p;
    st.col_mode = (int)p[UCL_P_COLLAPSE_MODE];
    st.col_max_time = 0.0;
    switch (st.col_mode) {
    case 1: st.col_max_time = 1.175e6; p[UCL_P_FINALTIME] = CF(0.97) * st.col_max_time; break;
    case 2: st.col_max_time = 1.855e5; p[UCL_P_FINALTIME] = CF(0.97) * st.col_max_time; break;
    case 3: case 4: break;
    default: return -1;
    }
    st.col_parcel_radius = 1 * p[UCL_P_ROUT] / (double)1.0f;
    const ColFit f = col_fit(st, st.time_in_years);
    st.density = col_rhofit(st.col_mode, p[UCL_P_RIN], f);
    return 0;
}

// whole CTA: mass inside the parcel's starting radius (findMassInRadius, collapse.f90:117-135), modes 1-2.
// Ends with a barrier.
__device__ __noinline__ void collapse_initialize_dev(Smem &s, Blk &b)
{
    Scalars &st = s.st;
    BLOCK_SYNC();
    if (st.col_mode > 2) return; // block-uniform
    const ColFit f = col_fit(st, st.time_in_years);
    const double dr = st.col_parcel_radius / 1000;
    double m = 0.0;
    for (int i = 1 + (int)threadIdx.x; i <= 1000; i += NT) m += col_mass_panel(st.col_mode, i, dr, f);
    m = block_sum(s, b, m);
    T0_BEGIN
    st.col_mass_in_radius = m;
    T0_END
}

__device__ void collapse_target_time_dev(Scalars &st)
{
    const double t = st.time_in_years;
    if (t > 10000) st.target_time = (t + CF(1000.0)) * C_SPY;
    else if (t > 1000) st.target_time = (t + CF(100.0)) * C_SPY;
    else if (t > 0.0) st.target_time = (t * 10) * C_SPY;
    else st.target_time = 3.16e7 * 10.e-8;
}

// whole CTA, after coreUpdatePhysics: column density to the cloud edge, Av, the parcel's new radius and its
// density (collapse.f90:79-106).  Ends with a barrier.
__device__ __noinline__ void collapse_update_physics_dev(Smem &s, Blk &b)
{
    Scalars &st = s.st;
    const int tid = threadIdx.x;
    BLOCK_SYNC();
    const int mode = st.col_mode;
    const double t = st.time_in_years;
    const ColFit f = col_fit(st, t); // every thread evaluates the three fits itself (a handful of pow calls)
    const double rin = st.p[UCL_P_RIN], rout = st.p[UCL_P_ROUT];
    // findcoldens collapse.f90:158-176: 10^4 trapezoid panels between rin and rout
    double cd = 0.0;
    {
        const double size = rout - rin, dr = size / 10000;
        if (size > 0.0)
            for (int i = 1 + tid; i <= 10000; i += NT) {
                const double r1 = rin + (i - 1) * dr, r2 = rin + i * dr;
                cd += 0.5 * (col_rhofit(mode, r2, f) + col_rhofit(mode, r1, f)) * dr * C_PC;
            }
    }
    cd = block_sum(s, b, cd);
    double radius = st.col_parcel_radius;
    if (mode <= 2) {
        // findNewRadius collapse.f90:138-155: first panel whose running mass reaches the enclosed mass.
        // Thread k sums the CH panels k*CH+1 .. (k+1)*CH; thread 0 walks the chunk sums, then the panels of the
        // chunk in which the running mass crosses (and goes on beyond 10^4 panels if it has not crossed yet).
        constexpr int CH = (10000 + NT - 1) / NT;
        const double dr = rout / 1.0e4;
        double part = 0.0;
        for (int j = 0; j < CH; j++) {
            const int i = tid * CH + 1 + j;
            if (i <= 10000) part += col_mass_panel(mode, i, dr, f);
        }
        double *scratch = s.flux; // dead outside rhs_eval
        scratch[tid] = part;
        BLOCK_SYNC();
        if (tid == 0) {
            const double target = st.col_mass_in_radius;
            double m1 = 0.0;
            int k = 0;
            while (k < NT && k * CH < 10000 && m1 + scratch[k] < target) m1 += scratch[k++];
            int i = k * CH + 1;
            radius = 0.0;
            while (m1 < target && i < 100000000) { // guard: a NaN profile would never cross
                m1 += col_mass_panel(mode, i, dr, f);
                radius = i * dr;
                i++;
            }
            if (i == 1) radius = 0.0; // target <= 0: the reference's loop body never runs
        }
    } else if (tid == 0) {
        const double dt = st.target_time - st.current_time; // zero once the interval has been integrated
        radius = radius + col_velocity(st, radius, t) * dt / C_PC;
    }
    T0_BEGIN
    st.coldens = cd;
    st.av = st.p[UCL_P_BASEAV] + cd / 1.6e21;
    st.col_parcel_radius = radius;
    st.density = col_rhofit(mode, radius, f);
    T0_END
}
