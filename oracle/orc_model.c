/*
 * ORACLE (test infrastructure) -- model driver: CPU restatement of
 *   wrap.f90:549-697   solveAbundances (time loop, hooks, output rows)
 *   chemistry.f90:46-141 initializeChemistry, :145-239 updateChemistry,
 *   chemistry.f90:241-292 integrateODESystem (ISTATE policy)
 *   physics-core.f90:42-88 core physics, :121-157 ionizationDependency
 *   cloud.f90, hotcore.f90:30-90 hooks; cshock hooks live in orc_cshock.c
 *   io.f90:59-83 output row layout
 * One orc_model holds what the reference keeps in module globals, so models
 * are independent and the grid runner can thread over them.
 */
#define _POSIX_C_SOURCE 200809L
#include "orc_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>

#define MIN_ABUND 1.0e-30
#define MAX_LOOPS 10

static orc_model *model_alloc(const orc_network *net, int kind, const double *params)
{
    orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
    int neq = net->nspec + 1;
    m->net = net;
    m->kind = kind;
    memcpy(m->p, params, sizeof(double) * UCLGPU_NPARAM);
    m->abund = (double *)calloc(neq, sizeof(double));
    m->rate = (double *)calloc(net->nreac, sizeof(double));
    m->vdiff = (double *)calloc(net->nice, sizeof(double));
    m->desfrac = (double *)calloc(net->nreac, sizeof(double));
    m->abstol = (double *)calloc(neq, sizeof(double));
    m->vode = vode_alloc(neq);
    return m;
}

static void model_free(orc_model *m)
{
    free(m->abund); free(m->rate); free(m->vdiff); free(m->desfrac); free(m->abstol);
    vode_free(m->vode);
    free(m);
}

/* ionizationDependency, physics-core.f90:121-157 (zeta itself is never updated there) */
static void ionization_dependency(orc_model *m)
{
    static const double ckLDiss[10] = {1.582911005330e7, -6.465722684896e6, 1.172189025424e6,
                                       -1.237950798073e5, 8.393404654312e3, -3.788811358130e2,
                                       1.138688455029e1, -2.197136304567e-1, 2.469841278950e-3,
                                       -1.232393620924e-5};
    static const double ckHDiss[10] = {1.217227462831e7, -4.989649250304e6, 9.079152156645e5,
                                       -9.624890825395e4, 6.551161486120e3, -2.968976216187e2,
                                       8.959037875226e0, -1.735757324445e-1, 1.959267277734e-3,
                                       -9.816996707980e-6};
    if (m->p[UCL_P_IMPROVEDH2CRPDISSOCIATION] != 0.0) {
        const double *ck = (m->p[UCL_P_IONMODEL] == 0.0) ? ckLDiss : ckHDiss;
        double sum = 0.0, lc = log10(m->coldens);
        for (int k = 0; k < 10; k++) sum += ck[k] * pow(lc, (double)k);
        m->h2crprate = pow(10.0, sum) * m->zetascale;
    }
}

/* coreInitializePhysics, physics-core.f90:42-73 */
static int core_initialize_physics(orc_model *m)
{
    const double *p = m->p;
    m->time_in_years = m->current_time / SECONDS_PER_YEAR;
    m->cloudsize = (p[UCL_P_ROUT] - p[UCL_P_RIN]) * PC;
    m->gastemp = p[UCL_P_INITIALTEMP];
    m->dusttemp = m->gastemp;
    m->density = p[UCL_P_INITIALDENS];
    m->current_time_old = 0.0;
    m->radfield = p[UCL_P_RADFIELD];
    m->zeta = p[UCL_P_ZETA];
    if (p[UCL_P_IMPROVEDH2CRPDISSOCIATION] != 0.0 && p[UCL_P_COSMICRAYATTENUATION] == 0.0) return -1;
    m->coldens = (double)1.0f * m->cloudsize / (double)1.0f * p[UCL_P_INITIALDENS];
    m->av = p[UCL_P_BASEAV] + m->coldens / 1.6e21;
    m->zetascale = m->zeta;
    return 0;
}

/* coreUpdatePhysics, physics-core.f90:75-88 (points = 1) */
static void core_update_physics(orc_model *m)
{
    m->coldens = m->cloudsize / (double)1.0f * m->density;
    m->av = m->p[UCL_P_BASEAV] + m->coldens / 1.6e21;
    m->dusttemp = m->gastemp;
    if (m->p[UCL_P_COSMICRAYATTENUATION] != 0.0) ionization_dependency(m);
}

/* hotcore.f90:18-27 */
static const double HC_TEMPA[6] = {1.927e-1, 4.8560e-2, 7.8470e-3, 9.6966e-4, 1.706e-4, 4.74e-7};
static const float HC_TEMPB[6] = {0.5339f, 0.6255f, 0.8395f, 1.085f, 1.289f, 1.98f};

static int model_initialize_physics(orc_model *m)
{
    const double *p = m->p;
    switch (m->kind) {
    case UCLGPU_CLOUD: /* cloud.f90:20-31 */
        m->cloudsize = (p[UCL_P_ROUT] - p[UCL_P_RIN]) * PC;
        if (p[UCL_P_FREEFALL] != 0.0) m->density = (double)1.001f * p[UCL_P_INITIALDENS];
        return 0;
    case UCLGPU_HOTCORE: /* hotcore.f90:30-50 */
        m->temp_indx = (int)p[UCL_P_TEMPINDX];
        m->max_temp = p[UCL_P_MAXTEMP];
        if (p[UCL_P_FREEFALL] != 0.0) m->density = (double)1.001f * p[UCL_P_INITIALDENS];
        if (m->temp_indx > 6 || m->temp_indx < 1) return -1;
        return 0;
    case UCLGPU_CSHOCK:
        return orc_cshock_initialize(m);
    case UCLGPU_COLLAPSE:
        return orc_collapse_initialize(m);
    case UCLGPU_JSHOCK:
        return orc_jshock_initialize(m);
    case UCLGPU_POSTPROCESS: { /* postprocess.f90:24-92 */
        if (!m->pp_grid || m->pp_ntime < 1) return -1;
        const double *g = m->pp_grid;
        const int n = m->pp_ntime;
        if (m->pp_coldens) m->cloudsize = (double)0.f; /* shielding column densities supplied separately */
        m->p[UCL_P_ENDATFINALDENSITY] = 0.0;
        m->p[UCL_P_FREEFALL] = 0.0;
        m->pp_tstep = 1;
        m->target_time = g[0];
        m->density = g[1 * n];
        m->gastemp = g[2 * n];
        m->dusttemp = g[3 * n];
        m->radfield = g[4 * n];
        m->zeta = g[5 * n];
        if (m->pp_coldens) {
            m->coldens = g[6 * n];
            m->av = (double)5.348e-22f * m->coldens;
        }
        m->p[UCL_P_FINALTIME] = g[n - 1] / SECONDS_PER_YEAR;
        return 0;
    }
    }
    return -1;
}

static void update_target_time(orc_model *m)
{
    double t = m->time_in_years;
    switch (m->kind) {
    case UCLGPU_CLOUD: /* cloud.f90:39-51 */
        if (t >= 1.0e6) {
            m->target_time = (t + 1.0e5) * SECONDS_PER_YEAR;
        } else if (t > 10.0) {
            double om = pow(10.0, floor(log10(t)));
            m->target_time = ((floor(t / om) + 1.0) * om) * SECONDS_PER_YEAR;
        } else if (t > 0.0) {
            m->target_time = 10 * t * SECONDS_PER_YEAR;
        } else {
            m->target_time = SECONDS_PER_YEAR * 1.0e-7;
        }
        break;
    case UCLGPU_HOTCORE: /* hotcore.f90:56-72 */
        if (t > 1.0e6)
            m->target_time = (t + 1.0e5) * SECONDS_PER_YEAR;
        else if (t > 1.0e5)
            m->target_time = (t + 1.0e4) * SECONDS_PER_YEAR;
        else if (t > 1.0e4)
            m->target_time = (t + 1000.0) * SECONDS_PER_YEAR;
        else if (t > 1000)
            m->target_time = (t + 100.0) * SECONDS_PER_YEAR;
        else if (t > 100)
            m->target_time = (t + 10.0) * SECONDS_PER_YEAR;
        else if (t > 0.0)
            m->target_time = (t * 10.0) * SECONDS_PER_YEAR;
        else
            m->target_time = SECONDS_PER_YEAR * 1.0e-7;
        break;
    case UCLGPU_CSHOCK:
        orc_cshock_update_target_time(m);
        break;
    case UCLGPU_COLLAPSE:
        orc_collapse_update_target_time(m);
        break;
    case UCLGPU_JSHOCK:
        orc_jshock_update_target_time(m);
        break;
    case UCLGPU_POSTPROCESS: /* postprocess.f90:100-106 (tstep past the end of the history: clamped, see run loop) */
        if (m->pp_tstep > m->pp_ntime) m->pp_tstep = m->pp_ntime; /* guard: the Fortran would read past the array */
        m->target_time = m->pp_grid[m->pp_tstep - 1] + (double)1.f * SECONDS_PER_YEAR;
        break;
    }
}

static void model_update_physics(orc_model *m)
{
    switch (m->kind) {
    case UCLGPU_CLOUD:
        break;
    case UCLGPU_HOTCORE: /* hotcore.f90:78-90 (dstep = points = 1) */
        if (m->gastemp < m->max_temp) {
            double g = (m->cloudsize / (m->p[UCL_P_ROUT] * PC)) * (double)(1.0f / 1.0f);
            g = pow(g, -0.5);
            int k = m->temp_indx - 1;
            g = m->p[UCL_P_INITIALTEMP] +
                ((HC_TEMPA[k] * pow(m->current_time / SECONDS_PER_YEAR, (double)HC_TEMPB[k])) * g);
            if (g > m->max_temp) g = m->max_temp;
            m->gastemp = g;
        }
        m->dusttemp = m->gastemp;
        break;
    case UCLGPU_CSHOCK:
        orc_cshock_update_physics(m);
        break;
    case UCLGPU_COLLAPSE:
        orc_collapse_update_physics(m);
        break;
    case UCLGPU_JSHOCK:
        orc_jshock_update_physics(m);
        break;
    case UCLGPU_POSTPROCESS: { /* postprocess.f90:112-129 */
        const double *g = m->pp_grid;
        const int n = m->pp_ntime, k = m->pp_tstep - 1;
        m->target_time = g[k];
        m->density = g[1 * n + k];
        m->gastemp = g[2 * n + k];
        m->dusttemp = g[3 * n + k];
        m->radfield = g[4 * n + k];
        m->zeta = g[5 * n + k];
        if (m->pp_coldens) {
            m->coldens = g[6 * n + k];
            m->av = (double)5.348e-22f * m->coldens;
        }
        m->pp_tstep = m->pp_tstep + 1;
        break;
    }
    }
}

static void sublimation(orc_model *m)
{
    /* cloud.f90:62-65 no-op; hotcore.f90:92-107 no-op for THREE_PHASE networks */
    if (m->kind == UCLGPU_CSHOCK) orc_cshock_sublimation(m);
    if (m->kind == UCLGPU_JSHOCK) orc_jshock_sublimation(m); /* jshock.f90:143-152 */
}

/* initializeChemistry, chemistry.f90:46-141.  Absent elements carry index nspec
 * (the density slot), are written there and then overwritten. */
static void initialize_chemistry(orc_model *m, int read_abunds)
{
    const orc_network *net = m->net;
    const int32_t *nm = net->named;
    const double *p = m->p;
    int neq = net->nspec + 1;
    double *a = m->abund;
    if (!read_abunds) {
        for (int i = 0; i < neq; i++) a[i] = MIN_ABUND;
        a[nm[I_NO]] = p[UCL_P_FO];
        a[nm[I_NN]] = p[UCL_P_FN];
        a[nm[I_NMG]] = p[UCL_P_FMG];
        a[nm[I_NP]] = p[UCL_P_FP];
        a[nm[I_NF]] = p[UCL_P_FF];
        a[nm[I_NNA]] = p[UCL_P_FNA];
        a[nm[I_NLI]] = p[UCL_P_FLI];
        a[nm[I_NPAH]] = p[UCL_P_FPAH];
        a[nm[I_NSX]] = p[UCL_P_FS];
        a[nm[I_NSIX]] = p[UCL_P_FSI];
        a[nm[I_NCLX]] = p[UCL_P_FCL];
        switch ((int)p[UCL_P_ION]) {
        case 0:
            a[nm[I_NC]] = p[UCL_P_FC];
            a[nm[I_NCX]] = 1.e-10;
            break;
        case 1:
            a[nm[I_NC]] = p[UCL_P_FC] * (double)0.5f;
            a[nm[I_NCX]] = p[UCL_P_FC] * (double)0.5f;
            break;
        case 2:
            a[nm[I_NC]] = 1.e-10;
            a[nm[I_NCX]] = p[UCL_P_FC];
            break;
        }
        a[nm[I_N18O]] = p[UCL_P_F18O];
        a[nm[I_N15N]] = p[UCL_P_F15N];
        a[nm[I_N13C]] = p[UCL_P_F13C];
        a[nm[I_NELEC]] = a[nm[I_NCX]] + a[nm[I_NSIX]] + a[nm[I_NSX]] + a[nm[I_NCLX]] + a[nm[I_NMGX]];
        for (int i = 0; i < neq; i++) a[i] = a[i] * p[UCL_P_METALLICITY];
        a[nm[I_NH]] = p[UCL_P_FH];
        a[nm[I_NH2]] = (double)0.5f * ((double)1.0f - p[UCL_P_FH]);
        a[nm[I_ND]] = p[UCL_P_FD];
        a[nm[I_NHE]] = p[UCL_P_FHE];
    }
    a[neq - 1] = m->density;
    orc_init_vdiff(m);
    orc_init_desfrac(m);
    for (int j = 0; j < net->nreac; j++) m->rate[j] = 0.0;
    m->last_temp = 99.0e99;
}

/* integrateODESystem, chemistry.f90:241-292. Returns successFlag. */
static int integrate_ode_system(orc_model *m)
{
    const orc_network *net = m->net;
    int neq = net->nspec + 1;
    for (int i = 0; i < neq; i++) {
        m->abstol[i] = m->abstol_factor * m->abund[i];
        if (m->abstol[i] < m->p[UCL_P_ABSTOL_MIN]) m->abstol[i] = m->p[UCL_P_ABSTOL_MIN];
    }
    vode_t *v = m->vode;
    int istate = vode_solve(v, orc_rhs, m, m->abund, &m->current_time, m->target_time, m->p[UCL_P_RELTOL],
                            m->abstol, m->mxstep);
    m->stats.nst += v->nst; m->stats.nfe += v->nfe; m->stats.nje += v->nje; m->stats.nlu += v->nlu;
    m->stats.nni += v->nni; m->stats.ncfn += v->ncfn; m->stats.netf += v->netf;
    if (orc_deadline_expired()) return ORC_FLAG_DEADLINE; /* benchmark guard, not reference behaviour */
    {
        const char *dbg = getenv("ORC_DEBUG"); /* "1": failed DVODE calls, "2": every call */
        if (dbg && (istate != 2 || dbg[0] == '2'))
            fprintf(stderr, "[oracle] ISTATE %d at t=%.6e yr (target %.6e yr) T=%.3f nst=%ld netf=%ld ncfn=%ld nje=%ld nlu=%ld\n",
                    istate, m->current_time / SECONDS_PER_YEAR, m->target_time / SECONDS_PER_YEAR, m->gastemp, v->nst,
                    v->netf, v->ncfn, v->nje, v->nlu);
    }
    switch (istate) {
    case -1:
    case -4:
    case -5:
        m->target_time = m->current_time + (m->target_time - m->current_time) * (double)0.1f;
        break;
    case -2:
        m->abstol_factor = m->abstol_factor * (double)10.0f;
        break;
    case -3:
        return UCLGPU_INT_UNRECOVERABLE_ERROR;
    default:
        m->mxstep = 10000;
    }
    if (m->p[UCL_P_ENFORCECHARGECONSERVATION] != 0.0) {
        double s = 0.0;
        for (int i = 0; i < net->nspec; i++)
            if (net->is_ion[i]) s += m->abund[i];
        m->abund[net->named[I_NELEC]] = s;
    }
    return 0;
}

/* per-try set-up of updateChemistry, chemistry.f90:166-203 */
static void chemistry_setup(orc_model *m)
{
    const orc_network *net = m->net;
    const int32_t *nm = net->named;
    int neq = net->nspec + 1;
    double *a = m->abund;
    if (m->p[UCL_P_FREEFALL] == 0.0) a[neq - 1] = m->density;
    double cs = m->cloudsize / (double)1.0f;
    m->h2col = 0.0 + (double)0.5f * a[nm[I_NH2]] * m->density * cs;
    m->cocol = 0.0 + (double)0.5f * a[nm[I_NCO]] * m->density * cs;
    m->ccol = 0.0 + (double)0.5f * a[nm[I_NC]] * m->density * cs;
    if (m->pp_coldens) { /* chemistry.f90:183-189: postprocessed tracers have column densities provided */
        const int n = m->pp_ntime, k = m->pp_tstep - 1;
        m->h2col = m->pp_grid[7 * n + k];
        m->cocol = m->pp_grid[8 * n + k];
        m->ccol = m->pp_grid[6 * n + k] * a[nm[I_NC]];
    }
    double sb = 0.0, ss = 0.0;
    for (int k = 0; k < net->nsurf; k++) sb += a[net->bulk_list[k]];
    for (int k = 0; k < net->nsurf; k++) ss += a[net->surface_list[k]];
    a[nm[I_NBULK]] = sb;
    a[nm[I_NSURFACE]] = ss;
    m->safe_mantle = fmax(1e-30, a[nm[I_NSURFACE]]);
    m->safe_bulk = fmax(1e-30, a[nm[I_NBULK]]);
    if (net->n_refractory > 0) {
        double s = 0.0;
        for (int k = 0; k < net->n_refractory; k++) s += a[net->refractory_list[k]];
        m->safe_bulk = m->safe_bulk - s;
    }
    m->blr = fmin(1.0, orc_num_sites_per_grain() / (orc_gas_dust_density_ratio() * m->safe_bulk));
    orc_calculate_reaction_rates(m);
}

/* updateChemistry, chemistry.f90:145-239 (usepostprocess=.true.: the target is
 * restored after every try and the too-many-fails checks never run, Q3) */
static int update_chemistry(orc_model *m)
{
    int neq = m->net->nspec + 1;
    int loop = 0;
    double original_target = m->target_time;
    while (m->current_time < m->target_time && loop < MAX_LOOPS) {
        chemistry_setup(m);
        int flag = integrate_ode_system(m);
        if (flag < 0) return flag;
        for (int i = 0; i < neq; i++)
            if (m->abund[i] < MIN_ABUND) m->abund[i] = MIN_ABUND;
        m->density = m->abund[neq - 1];
        loop++;
        m->target_time = original_target;
    }
    return 0;
}

static int output_row(orc_model *m, int dtime, int timepoints, double *phys, double *chem, double *rates)
{
    /* io.f90:59-98; dtime is 1-based */
    const orc_network *net = m->net;
    if (dtime > timepoints + 1) return UCLGPU_NOT_ENOUGH_TIMEPOINTS_ERROR;
    if (phys) {
        double *r = phys + (size_t)(dtime - 1) * UCLGPU_NPHYS;
        r[0] = m->time_in_years; r[1] = m->density; r[2] = m->gastemp; r[3] = m->dusttemp;
        r[4] = m->av; r[5] = m->radfield; r[6] = m->zeta; r[7] = 1.0;
    }
    if (chem) memcpy(chem + (size_t)(dtime - 1) * net->nspec, m->abund, sizeof(double) * net->nspec);
    if (rates) memcpy(rates + (size_t)(dtime - 1) * net->nreac, m->rate, sizeof(double) * net->nreac);
    return 0;
}

int orc_run_model(const orc_network *net, int kind, const double *params, const double *y0, double *y_final,
                  double *phys_final, int timepoints, double *phys_traj, double *chem_traj, double *rates_traj,
                  int *nrows, double *dissipation_time, orc_stats *stats)
{
    return orc_run_model_pp(net, kind, params, y0, y_final, phys_final, timepoints, phys_traj, chem_traj, rates_traj,
                            nrows, dissipation_time, stats, 0, NULL, 0);
}

/* ... and the postprocess model (wrap.f90:357-443): kind = UCLGPU_POSTPROCESS with a tracer history
 * pp_grid[10][pp_ntime] (rows: time in s, density, gas T, dust T, radfield, zeta, N_H, N_H2, N_CO, N_C) */
int orc_run_model_pp(const orc_network *net, int kind, const double *params, const double *y0, double *y_final,
                     double *phys_final, int timepoints, double *phys_traj, double *chem_traj, double *rates_traj,
                     int *nrows, double *dissipation_time, orc_stats *stats, int pp_ntime, const double *pp_grid,
                     int pp_coldens)
{
    orc_model *m = model_alloc(net, kind, params);
    m->pp_ntime = pp_ntime;
    m->pp_grid = pp_grid;
    m->pp_coldens = (kind == UCLGPU_POSTPROCESS) ? pp_coldens : 0;
    const double *p = m->p;
    int neq = net->nspec + 1;
    int flag = 0;
    int want_traj = (phys_traj || chem_traj || rates_traj);
    m->current_time = 0.0;
    m->time_in_years = 0.0;
    m->phi = p[UCL_P_PHI];
    m->abstol_factor = p[UCL_P_ABSTOL_FACTOR];
    m->mxstep = (int)p[UCL_P_MXSTEP];
    if (core_initialize_physics(m) != 0 || model_initialize_physics(m) != 0) {
        flag = UCLGPU_PHYSICS_INIT_ERROR;
        goto finish;
    }
    initialize_chemistry(m, 0);
    if (y0) {
        /* wrap.f90:636-640: abund(:nspec+1) = abundanceStart(:nspec+1); the density
         * slot is re-synchronised by updateChemistry (:166) or, in free fall, kept
         * from initializeChemistry. */
        for (int i = 0; i < net->nspec; i++) m->abund[i] = y0[i];
    }
    int dtime = 1;
    if (want_traj) {
        flag = output_row(m, dtime, timepoints, phys_traj, chem_traj, rates_traj);
        if (flag < 0) goto finish;
    }
    while (flag == 0 && ((p[UCL_P_ENDATFINALDENSITY] != 0.0 && m->density < p[UCL_P_FINALDENS]) ||
                         (p[UCL_P_ENDATFINALDENSITY] == 0.0 && m->time_in_years < p[UCL_P_FINALTIME]))) {
        dtime++;
        m->current_time_old = m->current_time;
        m->time_in_years = m->current_time / SECONDS_PER_YEAR;
        update_target_time(m);
        m->current_time = m->current_time_old;
        flag = update_chemistry(m);
        if (flag < 0) break;
        m->stats.nintervals++;
        m->time_in_years = m->target_time / SECONDS_PER_YEAR;
        core_update_physics(m);
        model_update_physics(m);
        sublimation(m);
        if (want_traj) flag = output_row(m, dtime, timepoints, phys_traj, chem_traj, rates_traj);
    }
    if (nrows) *nrows = dtime;
finish:
    if (y_final) memcpy(y_final, m->abund, sizeof(double) * neq);
    if (phys_final) {
        phys_final[0] = m->time_in_years; phys_final[1] = m->density; phys_final[2] = m->gastemp;
        phys_final[3] = m->dusttemp; phys_final[4] = m->av; phys_final[5] = m->radfield;
        phys_final[6] = m->zeta; phys_final[7] = 1.0;
    }
    if (dissipation_time) *dissipation_time = m->cs_dissipation_time;
    if (stats) *stats = m->stats;
    model_free(m);
    return flag;
}

typedef struct {
    const orc_network *net;
    int kind;
    int64_t ncell;
    const double *params, *y0;
    double *y_final, *phys_final;
    int32_t *flag;
    orc_stats *stats;
    int64_t *next; /* shared work counter */
    pthread_mutex_t *lock;
    double *cell_seconds; /* optional: wall time of each model (benchmark accounting) */
} grid_job;

static void *grid_worker(void *arg)
{
    grid_job *j = (grid_job *)arg;
    int neq = j->net->nspec + 1;
    for (;;) {
        pthread_mutex_lock(j->lock);
        int64_t c = (*j->next)++;
        pthread_mutex_unlock(j->lock);
        if (c >= j->ncell || orc_deadline_expired()) break; /* cells never started keep ORC_FLAG_DEADLINE */
        double p[UCLGPU_NPARAM];
        for (int k = 0; k < UCLGPU_NPARAM; k++) p[k] = j->params[(size_t)k * j->ncell + c];
        int nrows = 0;
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        j->flag[c] = orc_run_model(j->net, j->kind, p, j->y0 ? j->y0 + (size_t)c * neq : NULL,
                                   j->y_final + (size_t)c * neq,
                                   j->phys_final ? j->phys_final + (size_t)c * UCLGPU_NPHYS : NULL, 0, NULL,
                                   NULL, NULL, &nrows, NULL, j->stats ? j->stats + c : NULL);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        if (j->cell_seconds) j->cell_seconds[c] = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    }
    return NULL;
}

/* The reference runs grids as independent processes (scripts/grid.py:58-59); the
 * oracle uses one worker thread per requested core with a shared work counter. */
int orc_run_grid(const orc_network *net, int kind, int64_t ncell, const double *params, const double *y0,
                 double *y_final, double *phys_final, int32_t *flag, orc_stats *stats, int nthreads)
{
    return orc_run_grid_timed(net, kind, ncell, params, y0, y_final, phys_final, flag, stats, nthreads, NULL);
}

/* ... and with the wall time of every model (cells the deadline guard never let start keep -1) */
int orc_run_grid_timed(const orc_network *net, int kind, int64_t ncell, const double *params, const double *y0,
                       double *y_final, double *phys_final, int32_t *flag, orc_stats *stats, int nthreads,
                       double *cell_seconds)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    int64_t next = 0;
    pthread_mutex_t lock = PTHREAD_MUTEX_INITIALIZER;
    grid_job job = {net, kind, ncell, params, y0, y_final, phys_final, flag, stats, &next, &lock, cell_seconds};
    if (cell_seconds)
        for (int64_t c = 0; c < ncell; c++) cell_seconds[c] = -1.0;
    for (int64_t c = 0; c < ncell; c++) flag[c] = ORC_FLAG_DEADLINE;
    pthread_t th[256];
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, grid_worker, &job);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}

/* helper shared by get_rates/get_odes: the state wrap.f90:446-547 builds */
static orc_model *prepare_single(const orc_network *net, const double *params, const double *y_in)
{
    orc_model *m = model_alloc(net, UCLGPU_CLOUD, params);
    m->current_time = 0.0;
    m->phi = m->p[UCL_P_PHI];
    m->abstol_factor = m->p[UCL_P_ABSTOL_FACTOR];
    m->mxstep = (int)m->p[UCL_P_MXSTEP];
    core_initialize_physics(m);
    model_initialize_physics(m);
    initialize_chemistry(m, 0);
    for (int i = 0; i < net->nspec; i++) m->abund[i] = y_in[i];
    m->abund[net->nspec] = m->p[UCL_P_INITIALDENS];
    return m;
}

int orc_get_rates(const orc_network *net, const double *params, const double *y_in, double *rates_out)
{
    orc_model *m = prepare_single(net, params, y_in);
    chemistry_setup(m);
    memcpy(rates_out, m->rate, sizeof(double) * net->nreac);
    model_free(m);
    return 0;
}

/* F at exactly the given state (no pre-integration): the counterpart of the device probe uclgpu_probe_rhs */
int orc_probe_rhs(const orc_network *net, const double *params, const double *y_in, double *ydot_out)
{
    orc_model *m = prepare_single(net, params, y_in);
    chemistry_setup(m);
    orc_rhs(m, m->current_time, m->abund, ydot_out);
    model_free(m);
    return 0;
}

int orc_get_odes(const orc_network *net, const double *params, const double *y_in, double *ydot_out)
{
    /* wrap.f90:516-547 integrates 1e-7 s first (updateChemistry) and then calls F */
    orc_model *m = prepare_single(net, params, y_in);
    m->target_time = 1.0e-7;
    int flag = update_chemistry(m);
    orc_rhs(m, m->current_time, m->abund, ydot_out);
    model_free(m);
    return flag;
}

/* Experiment hook support (debug only, see orc_vode.c): the frozen rate coefficients of the model a DVODE callback
 * is running in (ctx of the RHS is the orc_model). */
const double *orc_ctx_rate(void *ctx) { return ((orc_model *)ctx)->rate; }
double orc_ctx_surfgrowth(void *ctx) { return ((orc_model *)ctx)->surfgrowth; }
