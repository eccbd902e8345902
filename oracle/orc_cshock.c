/*
 * ORACLE (test infrastructure) -- C-shock physics hooks and ice sputtering:
 *   cshock.f90:38-141 initializePhysics, :149-156 updateTargetTime,
 *   cshock.f90:159-208 updatePhysics, :211-219 sublimation, :225-252 shst
 *   sputtering.f90:65-235 sputterIces / iceYieldRate / trapezoid integration
 * No golden trajectory exists for this model in the reference: parity for the
 * C-shock path is pinned only through this self-validated restatement.
 */
#include "orc_internal.h"

#include <math.h>
#include <string.h>

static const double MH = 1.67262164e-24, KM = 1.e5;

/* single-precision constant expression log((1/x)+sqrt((1/x)**2-1)) as gfortran folds it */
static double coshinv_f(float x)
{
    float inv = 1.0f / x;
    float sq = (float)sqrt((double)(inv * inv - 1.0f));
    return (double)(float)log((double)(inv + sq));
}

int orc_cshock_initialize(orc_model *m)
{
    double *p = m->p;
    m->vs = p[UCL_P_VS];
    m->timestep_factor = p[UCL_P_TIMESTEPFACTOR];
    m->min_postshock_temp = p[UCL_P_MINIMUMPOSTSHOCKTEMP];
    m->cs_drift_vel = 0.0;
    m->cs_zn0 = 0.0;
    m->cs_vn0 = 0.0;
    m->cloudsize = (p[UCL_P_ROUT] - p[UCL_P_RIN]) * PC;
    if (p[UCL_P_FREEFALL] != 0.0) p[UCL_P_FREEFALL] = 0.0; /* cshock.f90:52-56 */
    if (p[UCL_P_POINTS] > 1) return -1;
    double id = p[UCL_P_INITIALDENS], vs = m->vs;
    m->density = id;
    m->current_time_old = 0.0;
    double max_temp;
    if (id > (double)powf(10.0f, 5.5f)) {
        max_temp = ((double)2.91731f * vs * vs) - ((double)23.78974f * vs) + (double)225.204167337f;
    } else if (id > (double)powf(10.0f, 4.5f)) {
        max_temp = ((double)3.38989f * vs * vs) + ((double)16.6519f * vs) + (double)96.569f;
        max_temp = (double)0.5f * max_temp;
    } else {
        max_temp = ((double)0.47258f * vs * vs) + ((double)40.44161f * vs) - (double)128.635455216f;
    }
    m->cs_max_temp = max_temp;
    double tsat = ((double)-15.38729f * vs * vs * vs) + ((double)2069.56962f * vs * vs) -
                  ((double)90272.826991f * vs) + (double)1686858.54278f;
    m->cs_tsat = tsat / id;
    m->cs_dlength = (double)12.0f * PC * vs / id;
    m->cs_dissipation_time = (m->cs_dlength * 1.0e-5 / vs) / SECONDS_PER_YEAR;
    double coshinv1 = coshinv_f(0.01f);
    m->cs_z2 = m->cs_dlength / coshinv1;
    m->cs_z1 = m->cs_z2 / (double)4.5f;
    double coshinv2 = coshinv_f(0.15f);
    double zmax = m->cs_dlength / coshinv2;
    m->cs_z3 = zmax / 6;
    double a1 = 6.0;
    m->cs_at = (1 / zmax) * pow((max_temp - p[UCL_P_INITIALTEMP]) * (exp(a1) - (double)1.f),
                                (double)(1.f / 6.f));
    /* bm0 = bm0*1D-06 (cshock.f90:126; each model starts from the default block) */
    double bm0 = p[UCL_P_BM0] * 1e-06;
    double va = bm0 / sqrt(4 * PI_F * MH);
    va = va / KM;
    double v0 = (double)2.f, v01 = 0;
    while (fabs(v0 - v01) >= (double)1e-6f) {
        v01 = v0;
        double g1 = -(va * va * vs * vs) / 2;
        double g2 = v01 * v01 - v01 * vs - va * va / 2;
        v0 = sqrt(g1 / g2);
    }
    m->cs_v0 = v0;
    return 0;
}

void orc_cshock_update_target_time(orc_model *m)
{
    if (m->time_in_years < 2.0 * m->cs_dissipation_time)
        m->target_time = (m->time_in_years + m->timestep_factor * m->cs_dissipation_time) * SECONDS_PER_YEAR;
    else
        m->target_time = ((double)1.1f * m->time_in_years) * SECONDS_PER_YEAR;
}

/* shst, cshock.f90:225-252 */
static void shst(orc_model *m)
{
    double vs = m->vs, v0 = m->cs_v0;
    double vn1 = 1e30, vn = m->cs_vn0, zn = m->cs_zn0;
    int loop = 0;
    while (fabs(vn - vn1) >= (double)1.e-10f && loop < 100) {
        vn1 = vn;
        double f1 = vs - vn1, f0 = vs - m->cs_vn0;
        zn = m->cs_zn0 + (m->current_time - m->current_time_old) * KM * (f1 + f0) / 2;
        double xcos = zn / m->cs_z2;
        double ach = 0.5 * (exp(xcos) + exp(-xcos));
        vn = (vs - v0) - ((vs - v0) / ach);
        loop++;
    }
    double xcos = zn / m->cs_z1;
    double ach = 0.5 * (exp(xcos) + exp(-xcos));
    double vi = (vs - v0) - ((vs - v0) / ach);
    m->cs_drift_vel = vi - vn;
    m->cs_zn0 = zn;
    m->cs_vn0 = vn;
    m->cs_zn = zn;
    m->cs_vn = vn;
}

void orc_cshock_update_physics(orc_model *m)
{
    const double bt = 6.0;
    shst(m);
    if (m->time_in_years > 0.0) m->density = m->p[UCL_P_INITIALDENS] * m->vs / (m->vs - m->cs_vn);
    if (m->time_in_years > 0.0) {
        double tn = m->p[UCL_P_INITIALTEMP] + (pow(m->cs_at * m->cs_zn, bt)) / (exp(m->cs_zn / m->cs_z3) - 1);
        m->gastemp = tn;
    }
    int post_shock = m->time_in_years > m->cs_dissipation_time;
    if (m->gastemp < m->min_postshock_temp && post_shock) m->gastemp = m->min_postshock_temp;
    m->dusttemp = m->gastemp;
}

/* --- sputtering.f90 ------------------------------------------------------- */
typedef struct {
    double sconst, eta, epso;
} sput_t;

static const double ICE_BINDING_ENERGY_F = 0.53f; /* 0.53*1.6d-12 with a default-real 0.53 */

static double ice_yield_integrand(const sput_t *s, double x, double pmass, double gastemp)
{
    const double yield_const = 8.3e-4;
    const double ebind = (double)(float)ICE_BINDING_ENERGY_F * 1.6e-12;
    double sv = s->sconst * sqrt(pmass);
    double eps = (x * x) * K_BOLTZ * gastemp;
    eps = s->eta * eps / ebind;
    double d = eps - s->epso;
    double yield = yield_const * (d * d) / ((double)1.f + pow(eps / (double)30.f, (double)1.3333f));
    return yield * (x * x) * (exp(-((x - sv) * (x - sv))) - exp(-((x + sv) * (x + sv))));
}

static double ice_yield_integral_limit(const sput_t *s, double xth, double pmass, double gastemp)
{
    int i = 1;
    double lim = xth + (1e3 - xth) * pow(0.5, i);
    while (ice_yield_integrand(s, lim, pmass, gastemp) < 1e-200 && (lim - xth) > 1.0e-3) {
        i++;
        lim = xth + (1e3 - xth) * pow(0.5, i);
    }
    return lim;
}

static void trapzd(const sput_t *s, double a, double b, double *sum_io, int n, double pmass, double gastemp)
{
    if (n == 1) {
        *sum_io = (double)0.5f * (b - a) * (ice_yield_integrand(s, a, pmass, gastemp) + ice_yield_integrand(s, b, pmass, gastemp));
    } else {
        long it = 1L << (n - 2);
        double tnm = (double)it;
        double del = (b - a) / tnm;
        double x = a + (double)0.5f * del;
        double sum = 0.0;
        for (long j = 0; j < it; j++) {
            sum = sum + ice_yield_integrand(s, x, pmass, gastemp);
            x = x + del;
        }
        *sum_io = (double)0.5f * (*sum_io + (b - a) * sum / tnm);
    }
}

static double trapezoid_integrate(const sput_t *s, double lo, double hi, double pmass, double gastemp)
{
    const double tol = (double)1.e-3f;
    double val = 0.0, olds = (double)-1.e30f;
    for (int j = 1; j <= 25; j++) {
        trapzd(s, lo, hi, &val, j, pmass, gastemp);
        if (fabs(val - olds) <= tol * fabs(olds)) return val;
        olds = val;
    }
    return val;
}

static double ice_yield_rate(sput_t *s, double pmass, double pdens, double gastemp)
{
    const double ebind = (double)(float)ICE_BINDING_ENERGY_F * 1.6e-12;
    const double target_mass = (double)18.0f * MH;
    const double eff = (double)0.8f;
    s->eta = (double)4.f * eff * pmass * target_mass * pow(pmass + target_mass, -2.0);
    s->epso = fmax((double)1.f, (double)4.f * s->eta);
    double sv = s->sconst * sqrt(pmass);
    double lower = sqrt(s->epso * ebind / (s->eta * K_BOLTZ * gastemp));
    double upper = ice_yield_integral_limit(s, lower, pmass, gastemp);
    double r;
    if ((upper - lower) > 1e-4) {
        r = trapezoid_integrate(s, lower, upper, pmass, gastemp) / sv;
        r = r * 1.e-5 * 1.e-5 * sqrt(8.0 * K_BOLTZ * gastemp * PI_F / pmass);
        r = r * pdens;
    } else {
        r = 0.0;
    }
    return r;
}

/* sputterIces, sputtering.f90:65-112 */
static void sputter_ices(orc_model *m, double shockvel, double gastemp, double density, double time_delta)
{
    const orc_network *net = m->net;
    const int32_t *nm = net->named;
    double *a = m->abund;
    sput_t s;
    s.sconst = (shockvel * shockvel * KM * KM) / (2.0 * gastemp * K_BOLTZ);
    s.sconst = sqrt(s.sconst);
    /* projectiles=(/nh2,nhe,nc,no,nsi,nco/) sputtering.f90:33 */
    int proj[6] = {nm[I_NH2], nm[I_NHE], nm[I_NC], nm[I_NO], nm[I_NSI], nm[I_NCO]};
    double rate = 0.0;
    for (int k = 0; k < 6; k++)
        rate = rate + ice_yield_rate(&s, net->mass[proj[k]] * MH, density * a[proj[k]], gastemp);
    double grain_number_density = density / orc_gas_dust_density_ratio();
    rate = rate * grain_number_density;
    double frac = rate * time_delta;
    double total = 0.0;
    for (int k = 0; k < net->nice; k++) total += a[net->ice_list[k]];
    frac = frac / total;
    if (frac > 1.0) frac = 1.0;
    if (frac < 0.0) frac = 0.0;
    /* shockVel >= VAPORIZE_SPEED (19 km/s) also sputters refractory species.
     * abund(gasIceList)=abund(gasIceList)+frac*abund(iceList) has a many-one vector
     * subscript on its left-hand side (each gas species is listed for its surface AND
     * its bulk partner): the right-hand side is evaluated from the old values into a
     * temporary and the elements are stored in order, so the LAST store (the bulk
     * partner) wins.  Reproduced as such. */
    int all = shockvel >= (double)19.0f;
    double tmp[4096];
    char use[4096];
    for (int k = 0; k < net->nice; k++) {
        int ice = net->ice_list[k], gas = net->gas_ice_list[k];
        use[k] = 1;
        if (!all && net->n_refractory > 0)
            for (int q = 0; q < net->n_refractory; q++)
                if (net->refractory_list[q] == ice) use[k] = 0;
        tmp[k] = a[gas] + frac * a[ice];
    }
    for (int k = 0; k < net->nice; k++)
        if (use[k]) a[net->gas_ice_list[k]] = tmp[k];
    for (int k = 0; k < net->nice; k++)
        if (use[k]) {
            int ice = net->ice_list[k];
            a[ice] = a[ice] - frac * a[ice];
        }
}

void orc_cshock_sublimation(orc_model *m)
{
    const orc_network *net = m->net;
    int neq = net->nspec + 1;
    double time_delta = m->current_time - m->current_time_old;
    double total = 0.0;
    for (int k = 0; k < net->nice; k++) total += m->abund[net->ice_list[k]];
    if (total > 1e-25 && m->cs_drift_vel > 0)
        sputter_ices(m, m->cs_drift_vel, m->gastemp, m->density, time_delta);
    for (int i = 0; i < neq; i++)
        if (m->abund[i] < 1.0e-50) m->abund[i] = 0.0;
}


/* ---- jshock_mod, jshock.f90 (James et al. 2020 J-shock parameterisation; single point) -------------------- */
#define JF(x) ((double)(x##f))

/* jshock.f90:29-78 */
int orc_jshock_initialize(orc_model *m)
{
    double *p = m->p;
    m->vs = p[UCL_P_VS];
    m->cloudsize = (p[UCL_P_ROUT] - p[UCL_P_RIN]) * PC;
    if (p[UCL_P_FREEFALL] != 0.0) p[UCL_P_FREEFALL] = 0.0;
    if (p[UCL_P_POINTS] > 1) return -1;
    const double vs = m->vs, id = p[UCL_P_INITIALDENS];
    m->density = id;
    m->js_max_temp = JF(5e3) * pow(vs / 10, 2.0);
    m->current_time_old = 0.0;
    double poly = JF(-2.058e-07) * pow(vs, 4.0) + JF(3.844e-05) * pow(vs, 3.0) - JF(0.002478) * pow(vs, 2.0) + JF(0.06183) * vs -
                  JF(0.4254);
    m->js_vmin = pow(pow(poly, 2.0), (double)0.5f);
    /* mfp = ((SQRT(2.0)*(1e3)*(pi*(2.4e-8)**2))**(-1))/1d4: SQRT(2.0), 1e3 and 2.4e-8 are single precision;
     * pi is the double parameter (single-precision literal value), so the product is double */
    double inner = (double)(sqrtf(2.0f) * 1e3f) * (PI_F * (double)(2.4e-8f * 2.4e-8f));
    m->js_tshock = ((1.0 / inner) / 1e4) / (vs * 1e5);
    m->js_tcool = (1 / id) * 1e6 * (60 * 60 * 24 * 365);
    m->js_max_dens = vs * id * 1e2;
    m->js_t_lambda = log(m->js_max_temp / p[UCL_P_INITIALTEMP]);
    m->js_n_lambda = log(m->js_max_dens / id);
    m->js_v0 = 0.0;
    return 0;
}

/* jshock.f90:85-97 */
void orc_jshock_update_target_time(orc_model *m)
{
    double t = m->time_in_years;
    if (t > JF(1e6))
        m->target_time = (t + JF(1e5)) * SECONDS_PER_YEAR;
    else if (t > 1.0e4)
        m->target_time = (t + 1000) * SECONDS_PER_YEAR;
    else if (t > 1.0e3)
        m->target_time = (t + JF(100.)) * SECONDS_PER_YEAR;
    else if (t * SECONDS_PER_YEAR < m->js_tshock)
        m->target_time = m->current_time + JF(0.05) * m->js_tshock;
    else
        m->target_time = JF(1.1) * m->current_time;
}

/* jshock.f90:102-135 */
void orc_jshock_update_physics(orc_model *m)
{
    const double *p = m->p;
    const double ct = m->current_time, id = p[UCL_P_INITIALDENS];
    double v0 = m->vs * exp(log(m->js_vmin / m->vs) * (ct / (p[UCL_P_FINALTIME] * 60 * 60 * 24 * 365)));
    if (v0 < m->js_vmin) v0 = m->js_vmin;
    m->js_v0 = v0;
    double tn;
    if (ct <= m->js_tshock) {
        tn = pow(ct / m->js_tshock, 2.0) * m->js_max_temp + p[UCL_P_INITIALTEMP];
        m->density = pow(ct / m->js_tshock, 3.0) * (4 * id);
        if (m->density < id) m->density = id;
    } else if (ct > m->js_tshock && ct <= m->js_tcool) {
        tn = m->js_max_temp * exp(-m->js_t_lambda * (ct / m->js_tcool));
        m->density = (4 * id) * exp(m->js_n_lambda * (ct / m->js_tcool));
        if (tn <= 10) tn = 10;
        if (m->density > m->js_max_dens) m->density = m->js_max_dens;
    } else {
        tn = 10;
        m->density = m->js_max_dens;
    }
    m->gastemp = tn;
    m->dusttemp = m->gastemp;
}

/* jshock.f90:143-152 */
void orc_jshock_sublimation(orc_model *m)
{
    const orc_network *net = m->net;
    int neq = net->nspec + 1;
    double time_delta = m->current_time - m->current_time_old;
    double total = 0.0;
    for (int k = 0; k < net->nice; k++) total += m->abund[net->ice_list[k]];
    if (total > 1e-25 && m->js_v0 > 0) sputter_ices(m, m->js_v0, m->gastemp, m->density, time_delta);
    for (int i = 0; i < neq; i++)
        if (m->abund[i] < 1.0e-50) m->abund[i] = 0.0;
}
