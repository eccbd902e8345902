/*
 * ORACLE (test infrastructure) -- CPU restatement of UCLCHEM's chemistry:
 *   photoreactions.f90 (H2/CO self shielding, C ionisation, dust scattering, NR splines)
 *   surfacereactions.f90 (H2 formation, diffusion, chemical desorption, swap rates)
 *   rates.f90:21-378 (calculateReactionRates, freezeOutRate, stickingCoefficient)
 *   odes.f90 GETYDOT, re-expressed from the MakeRates emitter rules
 *     (reaction.py:779-819, io_functions.py:533-661) as a table walk
 *   chemistry.f90:294-352 (F)
 * Single-precision literals of the Fortran source are reproduced with C float
 * literals promoted to double.
 */
#include "orc_internal.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- constants.f90:2-18, surfacereactions.f90:34-52 ----------------------- */
const double K_BOLTZ = 1.38065040e-16;
const double REDUCED_PLANCK = 1.054571628e-27;
const double AMU = 1.66053892e-24;
const double PI_F = (double)3.141592654f; /* PI = 3.141592654 is a default-real literal */
const double PC = 3.086e18;
const double SECONDS_PER_YEAR = 3.16e7;

static const double GAS_DUST_MASS_RATIO = 100.0, GRAIN_RADIUS = 1.e-5, GRAIN_DENSITY = 3.0;
static const double SURFACE_SITE_DENSITY = 1.5e15;
const double MAX_GRAIN_TEMP = 150.0, MIN_SURFACE_ABUND = 1.0e-20;
static const double CHEMICAL_BARRIER_THICKNESS = 1.40e-8;
static const double DIFFUSION_BIND_RATIO = 0.5;

double orc_thermal_vel(void) { return sqrt(8.0 * K_BOLTZ / (PI_F * AMU)); }
double orc_gas_dust_density_ratio(void)
{
    return (4.0 * PI_F * (GRAIN_RADIUS * GRAIN_RADIUS * GRAIN_RADIUS) * GRAIN_DENSITY * GAS_DUST_MASS_RATIO) /
           (3.0 * AMU);
}
double orc_num_sites_per_grain(void)
{
    return GRAIN_RADIUS * GRAIN_RADIUS * SURFACE_SITE_DENSITY * 4.0 * PI_F;
}
static double grain_crosssection_per_h(void) { return 0.5 * (7.908e-22 + 8.473e-22); }
static double grain_surfacearea_per_h(void) { return 4.0 * grain_crosssection_per_h(); }
static double vdiff_prefactor(void) { return 2.0 * K_BOLTZ * SURFACE_SITE_DENSITY / PI_F / PI_F / AMU; }

/* surfacereactions.f90:120-123 */
double orc_bulk_gain_from_mantle_buildup(void)
{
    return 0.5 * orc_gas_dust_density_ratio() / orc_num_sites_per_grain();
}

/* ======================= photoreactions.f90 ================================ */
static const double LAMBDA_GRID[30] = {910.0, 950.0, 1000.0, 1050.0, 1110.0, 1180.0, 1250.0, 1390.0,
                                       1490.0, 1600.0, 1700.0, 1800.0, 1900.0, 2000.0, 2100.0, 2190.0,
                                       2300.0, 2400.0, 2500.0, 2740.0, 3440.0, 4000.0, 4400.0, 5500.0,
                                       7000.0, 9000.0, 12500.0, 22000.0, 34000.0, 1.0e9};
static const double XLAMBDA_GRID[30] = {5.76, 5.18, 4.65, 4.16, 3.73, 3.40, 3.11, 2.74, 2.63, 2.62,
                                        2.54, 2.50, 2.58, 2.78, 3.01, 3.12, 2.86, 2.58, 2.35, 2.00,
                                        1.58, 1.42, 1.32, 1.00, 0.75, 0.48, 0.28, 0.12, 0.05, 0.00};
static const double NCO_GRID[8] = {12.0, 13.0, 14.0, 15.0, 16.0, 17.0, 18.0, 19.0};
static const double NH2_GRID[6] = {18.0, 19.0, 20.0, 21.0, 22.0, 23.0};
/* SCO_GRID(8,6) in storage (column-major) order == literal order, photoreactions.f90:33-39 */
static const double SCO_FLAT[48] = {
    0.000e+00, -1.408e-02, -1.099e-01, -4.400e-01, -1.154e+00, -1.888e+00, -2.760e+00, -4.001e+00,
    -8.539e-02, -1.015e-01, -2.104e-01, -5.608e-01, -1.272e+00, -1.973e+00, -2.818e+00, -4.055e+00,
    -1.451e-01, -1.612e-01, -2.708e-01, -6.273e-01, -1.355e+00, -2.057e+00, -2.902e+00, -4.122e+00,
    -4.559e-01, -4.666e-01, -5.432e-01, -8.665e-01, -1.602e+00, -2.303e+00, -3.146e+00, -4.421e+00,
    -1.303e+00, -1.312e+00, -1.367e+00, -1.676e+00, -2.305e+00, -3.034e+00, -3.758e+00, -5.077e+00,
    -3.883e+00, -3.888e+00, -3.936e+00, -4.197e+00, -4.739e+00, -5.165e+00, -5.441e+00, -6.446e+00};

/* spline, photoreactions.f90:335-409 (natural boundaries only: yp1=ypn=1e30) */
static void nr_spline(const double *x, const double *y, int n, double *y2)
{
    double u[100];
    y2[0] = 0.0;
    u[0] = 0.0;
    for (int i = 1; i < n - 1; i++) {
        double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
        double p = sig * y2[i - 1] + 2.0;
        y2[i] = (sig - 1.0) / p;
        u[i] = (6.0 * ((y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1])) /
                    (x[i + 1] - x[i - 1]) -
                sig * u[i - 1]) /
               p;
    }
    double qn = 0.0, un = 0.0;
    y2[n - 1] = (un - qn * u[n - 2]) / (qn * y2[n - 2] + 1.0);
    for (int k = n - 2; k >= 0; k--) y2[k] = y2[k] * y2[k + 1] + u[k];
}

/* splint, photoreactions.f90:440-560.  The hunt phase only chooses the search
 * start; the bracket it ends with is the bisection bracket xa(jlo) <= x < xa(jhi)
 * (ascending tables), with the end clamps of :544-554. */
static double nr_splint(const double *xa, const double *ya, const double *y2a, int n, double x)
{
    int jlo = 0, jhi = n + 1; /* 1-based */
    while (jhi - jlo != 1) {
        int jm = (jhi + jlo) / 2;
        if (x > xa[jm - 1])
            jlo = jm;
        else
            jhi = jm;
    }
    if (jlo == 0) { jlo = 1; jhi = 2; }
    if (jlo == n) { jlo = n - 1; jhi = n; }
    double h = xa[jhi - 1] - xa[jlo - 1];
    double a = (xa[jhi - 1] - x) / h;
    double b = (x - xa[jlo - 1]) / h;
    return a * ya[jlo - 1] + b * ya[jhi - 1] +
           ((a * a * a - a) * y2a[jlo - 1] + (b * b * b - b) * y2a[jhi - 1]) * (h * h) / 6.0;
}

/* xlambda, photoreactions.f90:225-243 */
static double xlambda(double lambda)
{
    double deriv[30];
    nr_spline(LAMBDA_GRID, XLAMBDA_GRID, 30, deriv);
    double v = lambda;
    if (lambda < LAMBDA_GRID[0]) v = LAMBDA_GRID[0];
    if (lambda > LAMBDA_GRID[29]) v = LAMBDA_GRID[29];
    double r = nr_splint(LAMBDA_GRID, XLAMBDA_GRID, deriv, 30, v);
    if (r < 0.0) r = 0.0;
    return r;
}

/* scatter, photoreactions.f90:133-196 */
static double scatter(double x1, double av)
{
    static const double c[6] = {1.0, 2.006, -1.438, 7.364e-01, -5.076e-01, -5.920e-02};
    static const double k1[6] = {7.514e-01, 8.490e-01, 1.013, 1.282, 2.005, 5.832};
    double tv = av / 1.086;
    double tl = tv * xlambda(x1);
    double sc = 0.0;
    if (tl < 1.0) {
        double expo = k1[0] * tl;
        if (expo < 100.0) sc = c[0] * exp(-expo);
    } else {
        for (int i = 1; i < 6; i++) {
            double expo = k1[i] * tl;
            if (expo < 100.0) sc = sc + c[i] * exp(-expo);
        }
    }
    return sc;
}

/* H2SelfShielding, photoreactions.f90:91-127 */
static double h2_self_shielding(double nh2, double doppler, double radwidth)
{
    const double FPARA = (double)0.5f, FOSC = 1.0e-2;
    double taud = FPARA * nh2 * (double)1.5e-2f * FOSC / doppler;
    double sj, sr;
    if (taud == 0.0)
        sj = 1.0;
    else if (taud < 2.0)
        sj = exp(-0.6666667 * taud);
    else if (taud < 10.0)
        sj = 0.638 * pow(taud, -1.25);
    else if (taud < 100.0)
        sj = 0.505 * pow(taud, -1.15);
    else
        sj = 0.344 * pow(taud, -1.0667);
    if (radwidth == 0.0) {
        sr = 0.0;
    } else {
        double r = radwidth / (1.7724539 * doppler);
        double t = 3.02 * pow(r * 1.0e+03, -0.064);
        double u = sqrt(taud * r) / t;
        sr = r / (t * sqrt(0.78539816 + u * u));
    }
    return sj + sr;
}

/* H2PhotoDissRate, photoreactions.f90:46-55 */
double orc_h2_photo_diss_rate(double nh2, double radfield, double av, double turbvel)
{
    const double base = 5.18e-11, xl = (double)1000.0f, radw = 8.0e7;
    double doppler = turbvel / (xl * 1.0e-8);
    return base * (radfield / (double)1.7f) * scatter(xl, av) * h2_self_shielding(nh2, doppler, radw);
}

/* COSelfShielding, photoreactions.f90:251-271 with the (8,6)->(7,6) mis-stride of
 * splie2/splin2 (:306-333, :411-438) reproduced. */
static double co_self_shielding(double nh2, double nco)
{
    const int m = 7, n = 6;
    double deriv[48];
    double ytmp[8], y2tmp[8], yytmp[8];
    for (int j = 0; j < m; j++) { /* splie2 */
        for (int k = 0; k < n; k++) ytmp[k] = SCO_FLAT[j + k * m];
        nr_spline(NH2_GRID, ytmp, n, y2tmp);
        for (int k = 0; k < n; k++) deriv[j + k * m] = y2tmp[k];
    }
    double lognco = log10(nco + 1.0);
    double lognh2 = log10(nh2 + 1.0);
    if (lognco < NCO_GRID[0]) lognco = NCO_GRID[0];
    if (lognh2 < NH2_GRID[0]) lognh2 = NH2_GRID[0];
    if (lognco > NCO_GRID[m - 1]) lognco = NCO_GRID[m - 1];
    if (lognh2 > NH2_GRID[n - 1]) lognh2 = NH2_GRID[n - 1];
    for (int j = 0; j < m; j++) { /* splin2 */
        for (int k = 0; k < n; k++) {
            ytmp[k] = SCO_FLAT[j + k * m];
            y2tmp[k] = deriv[j + k * m];
        }
        yytmp[j] = nr_splint(NH2_GRID, ytmp, y2tmp, n, lognh2);
    }
    nr_spline(NCO_GRID, yytmp, m, y2tmp);
    double r = nr_splint(NCO_GRID, yytmp, y2tmp, m, lognco);
    return pow(10.0, r);
}

/* lbar, photoreactions.f90:275-304 */
static double lbar(double u, double w)
{
    double lu = log10(fabs(u) + 1.0);
    double lw = log10(fabs(w) + 1.0);
    double lb = (5675.0 - 200.6 * lw) - (571.6 - 24.09 * lw) * lu + (18.22 - 0.7664 * lw) * (lu * lu);
    if (lb > 1076.1) lb = 1076.1;
    if (lb < 913.6) lb = 913.6;
    return lb;
}

/* COPhotoDissRate, photoreactions.f90:57-72 */
double orc_co_photo_diss_rate(double nh2, double nco, double radfield, double av)
{
    double ssf = co_self_shielding(nh2, nco);
    double lba = lbar(nco, nh2);
    double sca = scatter(lba, av);
    return (2.e-10) * (radfield / (double)1.7f) * ssf * sca;
}

/* cIonizationRate, photoreactions.f90:74-85 */
static double c_ionization_rate(double alpha, double gamma, double gastemp, double nc, double nh2, double av,
                                double radfield)
{
    double tauc = gamma * av + 1.1e-17 * nc + (0.9 * pow(gastemp, 0.27) * pow(nh2 / 1.59e21, 0.45));
    return alpha * (radfield / (double)1.7f) * exp(-tauc);
}

/* ======================= surfacereactions.f90 ============================== */
/* h2FormEfficiency, surfacereactions.f90:63-107 */
static double h2_form_efficiency(double gastemp, double dusttemp)
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

static int ice_pos(const orc_network *net, int species)
{
    int pos = -1;
    for (int i = 0; i < net->nice; i++)
        if (net->ice_list[i] == species) pos = i;
    return pos;
}

static int in_list(const int32_t *list, int n, int v)
{
    for (int i = 0; i < n; i++)
        if (list[i] == v) return 1;
    return 0;
}

/* diffusionReactionRate, surfacereactions.f90:158-211 */
static double diffusion_reaction_rate(const orc_model *m, int r, double dusttemp)
{
    const orc_network *net = m->net;
    int index1 = ice_pos(net, net->re[3 * r + 0]);
    int index2 = ice_pos(net, net->re[3 * r + 1]);
    const double *vd = m->vdiff, *eb = net->binding_energy;
    double diffuse = vd[index1] * exp(-DIFFUSION_BIND_RATIO * eb[index1] / dusttemp);
    diffuse = diffuse + (vd[index2] * exp(-DIFFUSION_BIND_RATIO * eb[index2] / dusttemp));
    double desorb = vd[index1] * exp(-eb[index1] / dusttemp);
    desorb = desorb + vd[index2] * exp(-eb[index2] / dusttemp);
    double reac = net->gama[r] / dusttemp;
    double rm = net->reduced_masses[r];
    if (rm == 0.0) {
        double m1 = net->mass[net->ice_list[index1]], m2 = net->mass[net->ice_list[index2]];
        rm = m1 * m2 / (m1 + m2);
    }
    double tunnel =
        2.0 * CHEMICAL_BARRIER_THICKNESS / REDUCED_PLANCK * sqrt(2.0 * AMU * rm * K_BOLTZ * net->gama[r]);
    if (reac > tunnel) reac = tunnel;
    reac = fmax(vd[index1], vd[index2]) * exp(-reac);
    reac = reac / (reac + desorb + diffuse);
    return net->alpha[r] * reac * diffuse * orc_gas_dust_density_ratio() / orc_num_sites_per_grain();
}

/* desorptionFraction, surfacereactions.f90:218-297 (index-space mix-up of Q12 kept) */
static double desorption_fraction(const orc_network *net, int r)
{
    const double EFFECTIVE_SURFACE_MASS = 120.0;
    int react1 = -1, react2 = -1, prod[4] = {-1, -1, -1, -1};
    const int32_t *re = net->re + 3 * r, *pr = net->pr + 4 * r;
    for (int i = 0; i < net->nice; i++) {
        int ice = net->ice_list[i], gas = net->gas_ice_list[i];
        if (ice == re[0]) react1 = i;
        if (gas == re[0]) react1 = i;
        if (ice == re[1]) react2 = i;
        if (gas == re[1]) react2 = i;
        for (int k = 0; k < 4; k++) {
            if (pr[k] >= 0 && ice == pr[k]) prod[k] = i;
        }
        for (int k = 0; k < 4; k++) {
            if (pr[k] >= 0 && gas == pr[k]) prod[k] = i;
        }
    }
    double max_be = 0.0, prod_enth = 0.0, eps_cd = 0.0;
    for (int k = 0; k < 4; k++) {
        if (prod[k] >= 0) {
            max_be = fmax(max_be, net->binding_energy[prod[k]]);
            prod_enth = prod_enth + net->formation_enthalpy[prod[k]];
            eps_cd = eps_cd + net->mass[prod[k]]; /* mass() indexed by ice position: reference quirk */
        }
    }
    double q = (eps_cd - EFFECTIVE_SURFACE_MASS) / (eps_cd + EFFECTIVE_SURFACE_MASS);
    eps_cd = q * q;
    double dh = net->formation_enthalpy[react1] + net->formation_enthalpy[react2] - prod_enth;
    dh = dh * 4.184e03 / (1.38054e-23 * 6.02214129e23);
    dh = dh + net->gama[r];
    if (dh == 0.00) dh = (double)1e-30f;
    int dof = net->atom_counts[prod[0]]; /* atomCounts() indexed by ice position: reference quirk */
    for (int k = 1; k < 4; k++)
        if (prod[k] >= 0 && net->atom_counts[prod[k]] > dof) dof = net->atom_counts[prod[k]];
    dof = 3 * dof;
    double frac = exp((-max_be * (double)(float)dof) / (eps_cd * dh));
    if (dh < 0.0) frac = 0.0;
    frac = frac / 10;
    int ngn = net->named[I_NGN], ngo = net->named[I_NGO], ngoh = net->named[I_NGOH], nh = net->named[I_NH];
    if (re[0] == ngn && re[1] == ngn) frac = (double)0.5f;
    if ((re[0] == ngo && re[1] == nh) || (re[0] == nh && re[1] == ngo)) frac = (double)0.3f;
    if ((re[0] == ngoh && re[1] == nh) || (re[0] == nh && re[1] == ngoh)) frac = (double)0.25f;
    return frac;
}

/* freezeOutRate for one reaction, rates.f90:354-367 */
static double freeze_out_rate(const orc_model *m, int r)
{
    const orc_network *net = m->net;
    double fr = 1.0 + net->beta[r] * 16.71e-4 / (GRAIN_RADIUS * m->gastemp);
    if (m->p[UCL_P_FREEZEFACTOR] == 0.0 || m->dusttemp > MAX_GRAIN_TEMP) return 0.0;
    return fr * m->p[UCL_P_FREEZEFACTOR] * net->alpha[r] * orc_thermal_vel() *
           sqrt(m->gastemp / net->mass[net->re[3 * r]]) * grain_crosssection_per_h();
}

/* stickingCoefficient, rates.f90:370-378 */
static double sticking_coefficient(double s0, double tcrit, double gastemp)
{
    double beta = 2.5;
    double tr = gastemp / tcrit;
    return s0 * (1.0 + beta * tr) / pow(1.0 + tr, beta);
}

/* vdiff set-up, chemistry.f90:108-112 */
void orc_init_vdiff(orc_model *m)
{
    const orc_network *net = m->net;
    for (int i = 0; i < net->nice; i++) {
        int j = net->ice_list[i];
        double v = vdiff_prefactor() * net->binding_energy[i] / net->mass[j];
        m->vdiff[i] = sqrt(v);
    }
}

#define RANGE(T) int lo = net->type_lo[T], hi = net->type_hi[T]
/* The reference guards every block with IF (idx1 .ne. idx2): absent types AND
 * types with exactly one reaction are skipped. */
#define PRESENT (lo >= 0 && lo != hi)

static void freeze_switch(orc_model *m, double *rate, int lo, int n_partner, int with_density)
{
    /* WHERE(rate(freezePartners)*abund(re1(freezePartners))[*density] < MIN_SURFACE_ABUND*rate(idx1:idx2))
     * rates.f90:106-107,127-128,150-151,223-224 -- mask evaluated before any assignment. */
    const orc_network *net = m->net;
    char mask[4096];
    for (int k = 0; k < n_partner; k++) {
        int fp = net->freeze_partners[k];
        double lhs = rate[fp] * m->abund[net->re[3 * fp]];
        if (with_density) lhs = lhs * m->density;
        mask[k] = lhs < MIN_SURFACE_ABUND * rate[lo + k];
    }
    for (int k = 0; k < n_partner; k++)
        if (mask[k]) rate[net->freeze_partners[k]] = 0.0;
}

/* calculateReactionRates, rates.f90:21-343 */
void orc_calculate_reaction_rates(orc_model *m)
{
    const orc_network *net = m->net;
    const double *p = m->p;
    double *rate = m->rate;
    const double *alpha = net->alpha, *beta = net->beta, *gama = net->gama;
    const double zeta = m->zeta, radfield = m->radfield, av = m->av;
    const double gastemp = m->gastemp, dusttemp = m->dusttemp;
    const double safe_mantle = m->safe_mantle;
    const int desorb = p[UCL_P_DESORB] != 0.0;
    const double ICE_GAS = (double)0.3f; /* ICE_GAS_PHOTO_CROSSSECTION_RATIO = 0.3 */

    { /* CRP :34-41 */
        RANGE(T_CRP);
        if (PRESENT)
            for (int j = lo; j <= hi; j++) rate[j] = alpha[j] * zeta;
        if (p[UCL_P_IMPROVEDH2CRPDISSOCIATION] != 0.0) rate[net->named[R_H2_CRP]] = m->h2crprate;
    }
    { /* PHOTON :43-56 */
        RANGE(T_PHOTON);
        if (PRESENT)
            for (int j = lo; j <= hi; j++) {
                rate[j] = alpha[j] * exp(-gama[j] * av) * radfield / (double)1.7f;
                int r1 = net->re[3 * j];
                if (in_list(net->bulk_list, net->nsurf, r1))
                    rate[j] = rate[j] * ICE_GAS * pow((double)(1.0f - 0.007f), ((double)1.0f + (double)0.5f / m->blr));
                else if (in_list(net->surface_list, net->nsurf, r1))
                    rate[j] = rate[j] * ICE_GAS;
            }
    }
    { /* CRPHOT :59-72 */
        RANGE(T_CRPHOT);
        if (PRESENT)
            for (int j = lo; j <= hi; j++) {
                rate[j] = alpha[j] * gama[j] * 1.0 / (1.0 - p[UCL_P_OMEGA]) * zeta * pow(gastemp / 300, beta[j]);
                int r1 = net->re[3 * j];
                if (in_list(net->bulk_list, net->nsurf, r1))
                    rate[j] = rate[j] * ICE_GAS * pow((double)(1 - 0.007f), (1 + (double)0.5f / m->blr));
                else if (in_list(net->surface_list, net->nsurf, r1))
                    rate[j] = rate[j] * ICE_GAS;
            }
    }
    { /* FREEZE :75-84 */
        RANGE(T_FREEZE);
        if (PRESENT) {
            for (int j = lo; j <= hi; j++) rate[j] = freeze_out_rate(m, j);
            int rh = net->named[R_HFREEZE], re_ = net->named[R_EFREEZE], rh2 = net->named[R_H2FREEZE];
            rate[re_] = rate[rh];
            rate[rh2] = sticking_coefficient(0.87, 87.0, gastemp) * rate[rh2];
            rate[rh] = sticking_coefficient(1.0, 52.0, gastemp) * rate[rh];
        }
    }
    { /* DESOH2 :91-108 */
        RANGE(T_DESOH2);
        if (PRESENT) {
            if (desorb && p[UCL_P_H2DESORB] != 0.0 && safe_mantle > MIN_SURFACE_ABUND) {
                double v = p[UCL_P_EPSILON] * h2_form_efficiency(gastemp, dusttemp);
                for (int j = lo; j <= hi; j++) rate[j] = (gama[j] > p[UCL_P_EBMAXH2]) ? 0.0 : v;
            } else {
                for (int j = lo; j <= hi; j++) rate[j] = 0.0;
            }
            freeze_switch(m, rate, lo, net->nsurf, 0);
        }
    }
    { /* DESCR :110-129.  `phi` here is the module variable that the GAR block at
         :322-326 overwrites on EVERY call (it is not guarded): from the second
         rates evaluation on, phi is the clamped G*sqrt(T)/n_e value. */
        RANGE(T_DESCR);
        if (PRESENT) {
            if (desorb && p[UCL_P_CRDESORB] != 0.0 && safe_mantle > MIN_SURFACE_ABUND) {
                double v = 4.0 * PI_F * zeta * 1.64e-4 * (grain_surfacearea_per_h()) * m->phi;
                for (int j = lo; j <= hi; j++) rate[j] = (gama[j] > p[UCL_P_EBMAXCR]) ? 0.0 : v;
            } else {
                for (int j = lo; j <= hi; j++) rate[j] = 0.0;
            }
            freeze_switch(m, rate, lo, net->nsurf, 1);
        }
    }
    { /* DEUVCR :132-152 */
        RANGE(T_DEUVCR);
        if (PRESENT) {
            if (desorb && p[UCL_P_UVDESORB] != 0.0 && safe_mantle > MIN_SURFACE_ABUND && zeta > 0) {
                double v = grain_crosssection_per_h() * p[UCL_P_UV_YIELD] * 4.875e3 * zeta;
                v = v * (1 + (radfield / p[UCL_P_UVCREFF]) * (1.0 / zeta) * exp(-(double)1.8f * av));
                for (int j = lo; j <= hi; j++) rate[j] = (gama[j] > p[UCL_P_EBMAXUVCR]) ? 0.0 : v;
            } else {
                for (int j = lo; j <= hi; j++) rate[j] = 0.0;
            }
            freeze_switch(m, rate, lo, net->nsurf, 1);
        }
    }
    { /* CRS :156-162 */
        RANGE(T_CRS);
        if (PRESENT)
            for (int j = lo; j <= hi; j++)
                rate[j] = alpha[j] * (beta[j] * (gama[j] / 100) * ((double)8.6f * zeta * (double)1.3f));
    }
    { /* EXRELAX :165-176 */
        RANGE(T_EXRELAX);
        if (PRESENT) {
            double va = 0.0;
            for (int j = lo; j <= hi; j++) {
                int pos = ice_pos(net, net->re[3 * j]);
                if (pos >= 0) va = m->vdiff[pos];
                rate[j] = va;
            }
        }
    }
    { /* EXSOLID :179-197 */
        RANGE(T_EXSOLID);
        if (PRESENT) {
            double va = 0.0, vb = 0.0;
            for (int j = lo; j <= hi; j++) {
                int p1 = ice_pos(net, net->re[3 * j]), p2 = ice_pos(net, net->re[3 * j + 1]);
                if (p1 >= 0) va = m->vdiff[p1];
                if (p2 >= 0) vb = m->vdiff[p2];
                rate[j] = (vb + va) / (SURFACE_SITE_DENSITY * 1.8e-8);
                rate[j] = alpha[j] * rate[j];
            }
        }
    }
    { /* THERM :202-229 */
        RANGE(T_THERM);
        if (PRESENT) {
            if (p[UCL_P_THERMDESORB] != 0.0) {
                for (int j = lo; j <= hi; j++) {
                    int pos = ice_pos(net, net->re[3 * j]);
                    if (pos >= 0) rate[j] = m->vdiff[pos] * exp(-gama[j] / dusttemp);
                }
                /* non-conformable WHERE: only the first size(freezePartners) THERM entries take part (Q11) */
                freeze_switch(m, rate, lo, net->nsurf, 1);
                if (safe_mantle < MIN_SURFACE_ABUND)
                    for (int j = lo; j <= hi; j++) rate[j] = 0.0;
            } else {
                for (int j = lo; j <= hi; j++) rate[j] = 0.0;
            }
        }
    }
    { /* LH / LHDES :236-260 */
        RANGE(T_LH);
        if (PRESENT) {
            int dlo = net->type_lo[T_LHDES], dhi = net->type_hi[T_LHDES];
            if (dusttemp < MAX_GRAIN_TEMP && safe_mantle > MIN_SURFACE_ABUND) {
                for (int j = lo; j <= hi; j++) rate[j] = diffusion_reaction_rate(m, j, dusttemp);
                for (int k = 0; k <= dhi - dlo; k++) rate[dlo + k] = rate[lo + k];
                for (int j = dlo; j <= dhi; j++) {
                    rate[j] = m->desfrac[j] * rate[j];
                    if (in_list(net->bulk_list, net->nsurf, net->re[3 * j])) rate[j] = 0.0;
                }
                for (int k = 0; k <= hi - lo; k++) rate[lo + k] = rate[lo + k] - rate[dlo + k];
            } else {
                for (int j = lo; j <= hi; j++) rate[j] = 0.0;
                for (int j = dlo; j <= dhi; j++) rate[j] = 0.0;
            }
        }
    }
    { /* ER / ERDES :264-279 (skipped by the guard when there is exactly one ER reaction) */
        RANGE(T_ER);
        if (PRESENT) {
            int dlo = net->type_lo[T_ERDES], dhi = net->type_hi[T_ERDES];
            for (int j = lo; j <= hi; j++) {
                rate[j] = freeze_out_rate(m, j);
                rate[j] = rate[j] * exp(-gama[j] / dusttemp);
            }
            for (int k = 0; k <= dhi - dlo; k++) rate[dlo + k] = rate[lo + k];
            for (int j = dlo; j <= dhi; j++) {
                rate[j] = m->desfrac[j] * rate[j];
                if (in_list(net->bulk_list, net->nsurf, net->re[3 * j])) rate[j] = 0.0;
            }
            for (int k = 0; k <= hi - lo; k++) rate[lo + k] = rate[lo + k] - rate[dlo + k];
        }
    }
    /* H2 formation :281-289 (PARAMETERIZE_H2FORM = .True.) */
    rate[net->named[R_H2FORM_CT]] = h2_form_efficiency(dusttemp, dusttemp);
    rate[net->named[R_H2FORM_ER]] = 0.0;
    rate[net->named[R_H2FORM_ERDES]] = 0.0;

    { /* bulkSurfaceExchangeReactions, surfacereactions.f90:109-150 */
        int off = (dusttemp > MAX_GRAIN_TEMP) || (safe_mantle < MIN_SURFACE_ABUND);
        int lo = net->type_lo[T_BULKSWAP], hi = net->type_hi[T_BULKSWAP];
        if (lo >= 0)
            for (int j = lo; j <= hi; j++) {
                if (off) {
                    rate[j] = 0.0;
                } else {
                    int pos = ice_pos(net, net->re[3 * j]);
                    if (pos >= 0) rate[j] = m->vdiff[pos] * exp(-net->binding_energy[pos] / dusttemp);
                }
            }
        lo = net->type_lo[T_SURFSWAP];
        hi = net->type_hi[T_SURFSWAP];
        if (lo >= 0)
            for (int j = lo; j <= hi; j++) rate[j] = off ? 0.0 : 1.0;
    }
    { /* TWOBODY :295-299 (recomputed only when T changed; masks below make that equivalent) */
        RANGE(T_TWOBODY);
        if (lo >= 0 && m->last_temp != gastemp)
            for (int j = lo; j <= hi; j++)
                rate[j] = alpha[j] * (pow(gastemp / (double)300.f, beta[j])) * exp(-gama[j] / gastemp);
    }
    { /* IONOPOL1 :301-305 */
        RANGE(T_IONOPOL1);
        if (PRESENT)
            for (int j = lo; j <= hi; j++)
                rate[j] = alpha[j] * beta[j] * (0.62 + 0.4767 * gama[j] * sqrt(300.0 / gastemp));
    }
    { /* IONOPOL2 :307-313 */
        RANGE(T_IONOPOL2);
        if (PRESENT)
            for (int j = lo; j <= hi; j++)
                rate[j] = alpha[j] * beta[j] *
                          (1.0 + 0.0967 * gama[j] * sqrt(300.0 / gastemp) +
                           gama[j] * gama[j] * (double)300.0f / ((double)10.526f * gastemp));
    }
    m->last_temp = gastemp;
    { /* GAR :316-332 -- the phi assignment is unconditional and leaks into DESCR */
        double phi = radfield * exp(-(double)2.5f * av) * sqrt(gastemp) /
                     (m->abund[net->nspec] * m->abund[net->named[I_NELEC]]);
        phi = fmin(fmax(phi, (double)1e2f), (double)1e6f);
        m->phi = phi;
        RANGE(T_GAR);
        if (PRESENT)
            for (int j = lo; j <= hi; j++) {
                const double *g = net->gar_params + 7 * (j - lo);
                rate[j] = (double)0.6f * alpha[j] * g[0] /
                          ((double)1.f + g[1] * pow(phi, g[2]) *
                                             ((double)1.f + g[3] * pow(gastemp, g[4]) *
                                                                pow(phi, -g[5] - g[6] * log(gastemp))));
            }
    }
    /* temperature-range masks :335-337 */
    for (int j = 0; j < net->nreac; j++) {
        if (!net->extrapolate[j] && gastemp < net->min_temps[j]) rate[j] = 0.0;
        if (!net->extrapolate[j] && gastemp > net->max_temps[j]) rate[j] = 0.0;
    }
    /* detailed photoreactions :340-342 */
    rate[net->named[R_H2_HV]] = orc_h2_photo_diss_rate(m->h2col, radfield, av, 1.0);
    rate[net->named[R_CO_HV]] = orc_co_photo_diss_rate(m->h2col, m->cocol, radfield, av);
    {
        int rc = net->named[R_C_HV];
        rate[rc] = c_ionization_rate(alpha[rc], gama[rc], gastemp, m->ccol, m->h2col, av, radfield);
    }
}

void orc_init_desfrac(orc_model *m)
{
    /* desorptionFraction does not depend on the physical state: tabulate once */
    const orc_network *net = m->net;
    for (int j = 0; j < net->nreac; j++) m->desfrac[j] = 0.0;
    int types[2] = {T_LHDES, T_ERDES};
    for (int t = 0; t < 2; t++) {
        int lo = net->type_lo[types[t]], hi = net->type_hi[types[t]];
        if (lo < 0) continue;
        for (int j = lo; j <= hi; j++) m->desfrac[j] = desorption_fraction(net, j);
    }
}

/* Experiment switch (debug only, tools/study_frozen_branch.py; 0 in every test and bench run = the reference's
 * rule): +1 / -1 force the mantle-growth / mantle-loss branch of the three-phase transfer regardless of the sign
 * of the uncorrected surface growth.  Not thread safe: single-model studies only. */
int g_orc_force_branch = 0;

/* GETYDOT, odes.f90:6-5181, as a walk over the MakeRates rules */
void orc_getydot(const orc_network *net, const double *rate, const double *y, double blr,
                 double surface_coverage, double safe_mantle, double safe_bulk, double dens, double *ydot,
                 double *surfgrowth_uncorrected)
{
    const int nspec = net->nspec, neq = nspec + 1;
    const int iB = net->named[I_NBULK], iS = net->named[I_NSURFACE];
    double *flux = (double *)malloc(sizeof(double) * net->nreac);
    double *loss = (double *)calloc(neq, sizeof(double));
    double *prod = (double *)calloc(neq, sizeof(double));
    (void)dens;
    /* totalSwap (odes.f90:10-...): sum of the BULKSWAP terms */
    double total_swap = 0.0;
    {
        int lo = net->type_lo[T_BULKSWAP], hi = net->type_hi[T_BULKSWAP];
        if (lo >= 0)
            for (int r = lo; r <= hi; r++) total_swap += rate[r] * y[net->re[3 * r]] * blr;
    }
    for (int r = 0; r < net->nreac; r++) {
        const int32_t *f = net->flux_factors + 5 * r;
        double v = rate[r];
        for (int k = 0; k < 5; k++) {
            int s = f[k];
            if (s < neq)
                v = v * y[s];
            else if (s == neq + 1)
                v = v * blr;
            else if (s == neq + 2)
                v = v / safe_mantle;
            else if (s == neq + 3)
                v = v * total_swap / safe_mantle;
            /* neq+0: constant one */
        }
        flux[r] = v;
    }
    for (int t = 0; t < net->n_loss; t++) loss[net->loss_species[t]] += flux[net->loss_reaction[t]];
    for (int t = 0; t < net->n_gain; t++) prod[net->gain_species[t]] += flux[net->gain_reaction[t]];
    for (int i = 0; i < nspec; i++) ydot[i] = prod[i] - loss[i];
    double sb = 0.0, ss = 0.0;
    for (int k = 0; k < net->nsurf; k++) sb += ydot[net->bulk_list[k]];
    ydot[iB] = sb;
    for (int k = 0; k < net->nsurf; k++) ss += ydot[net->surface_list[k]];
    ydot[iS] = ss;
    if (surfgrowth_uncorrected) *surfgrowth_uncorrected = ss;
    /* three-phase transfer, odes.f90:4815-5153 (species order: surface block, then bulk block) */
    int nrefr = net->n_refractory;
    if (g_orc_force_branch ? g_orc_force_branch < 0 : ydot[iS] < 0) {
        surface_coverage = fmin(1.0, safe_bulk / safe_mantle);
        for (int k = 0; k < net->nsurf; k++) {
            int s = net->surface_list[k], b = net->bulk_list[k];
            if (!(nrefr && in_list(net->refractory_list, nrefr, b)))
                ydot[s] = ydot[s] - ydot[iS] * surface_coverage * y[b] / safe_bulk;
        }
        for (int k = 0; k < net->nsurf; k++) {
            int b = net->bulk_list[k];
            if (!(nrefr && in_list(net->refractory_list, nrefr, b)))
                ydot[b] = ydot[b] + ydot[iS] * surface_coverage * y[b] / safe_bulk;
        }
    } else {
        for (int k = 0; k < net->nsurf; k++) {
            int s = net->surface_list[k];
            ydot[s] = ydot[s] - ydot[iS] * surface_coverage * y[s];
        }
        for (int k = 0; k < net->nsurf; k++) {
            int s = net->surface_list[k], b = net->bulk_list[k];
            ydot[b] = ydot[b] + ydot[iS] * surface_coverage * y[s];
        }
    }
    sb = 0.0;
    for (int k = 0; k < net->nsurf; k++) sb += ydot[net->bulk_list[k]];
    ydot[iB] = sb;
    ss = 0.0;
    for (int k = 0; k < net->nsurf; k++) ss += ydot[net->surface_list[k]];
    ydot[iS] = ss;
    free(flux);
    free(loss);
    free(prod);
}

/* densdot, physics-core.f90:90-103 */
double orc_densdot(const orc_model *m, double density)
{
    const double *p = m->p;
    if (density < p[UCL_P_FINALDENS] && p[UCL_P_FREEFALL] != 0.0) {
        double id = p[UCL_P_INITIALDENS];
        double e = (double)0.33f;
        return p[UCL_P_FREEFALLFACTOR] * pow(pow(density, (double)4.f) / id, e) *
               pow(8.4e-30 * id * (pow(density / id, e) - (double)1.f), (double)0.5f);
    }
    return 0.0;
}

/* F, chemistry.f90:294-352 */
void orc_rhs(void *ctx, double t, const double *y, double *ydot)
{
    orc_model *m = (orc_model *)ctx;
    const orc_network *net = m->net;
    const int neq = net->nspec + 1;
    (void)t;
    double d = y[neq - 1];
    for (int i = 0; i < neq; i++) ydot[i] = 0.0;
    /* points = 1: cloudSize/real(points) = cloudSize; *ColToCell = 0 */
    if (!m->pp_coldens) { /* chemistry.f90:310-319: column densities are fixed for postprocessing data */
        m->cocol = 0.0 + 0.5 * y[net->named[I_NCO]] * d * (m->cloudsize / (double)1.0f);
        m->h2col = 0.0 + 0.5 * y[net->named[I_NH2]] * d * (m->cloudsize / (double)1.0f);
        m->rate[net->named[R_H2_HV]] = orc_h2_photo_diss_rate(m->h2col, m->radfield, m->av, 1.0);
        m->rate[net->named[R_CO_HV]] = orc_co_photo_diss_rate(m->h2col, m->cocol, m->radfield, m->av);
    }
    m->safe_mantle = fmax(1e-30, y[net->named[I_NSURFACE]]);
    m->safe_bulk = fmax(1e-30, y[net->named[I_NBULK]]);
    m->blr = fmin(1.0, orc_num_sites_per_grain() / (orc_gas_dust_density_ratio() * m->safe_bulk));
    double cov = orc_bulk_gain_from_mantle_buildup();
    orc_getydot(net, m->rate, y, m->blr, cov, m->safe_mantle, m->safe_bulk, d, ydot, &m->surfgrowth);
    if (m->p[UCL_P_ENFORCECHARGECONSERVATION] != 0.0) {
        double s = 0.0;
        for (int i = 0; i < net->nspec; i++)
            if (net->is_ion[i]) s += ydot[i];
        ydot[net->named[I_NELEC]] = s;
    }
    ydot[neq - 1] = orc_densdot(m, y[neq - 1]);
}
