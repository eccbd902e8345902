/* ORACLE (test infrastructure): shared internals. */
#ifndef ORC_INTERNAL_H
#define ORC_INTERNAL_H

#include "../include/uclgpu.h" /* parameter column enum only (public product header) */
#include "orc_vode.h"
#include "uclchem_oracle.h"

typedef struct orc_model {
    const orc_network *net;
    double p[UCLGPU_NPARAM];
    int kind;
    /* physicscore module state (physics-core.f90:11-20) */
    double gastemp, dusttemp, density, av, coldens, cloudsize, radfield, zeta, zetascale, h2crprate;
    double time_in_years, current_time, target_time, current_time_old;
    /* chemistry module state (chemistry.f90:28-44, rates.f90:12-18, surfacereactions.f90:8-9,54) */
    double *abund; /* [neq] */
    double *rate;  /* [nreac] */
    double *vdiff; /* [nice] */
    double *desfrac; /* [nreac] tabulated desorptionFraction */
    double *abstol;
    double h2col, cocol, ccol;
    double safe_mantle, safe_bulk, blr, surfgrowth;
    double last_temp;
    double phi;           /* DEFAULTPARAMETERS phi, clobbered by rates.f90:322-326 */
    double abstol_factor; /* sticky x10 after ISTATE=-2 (chemistry.f90:269) */
    int mxstep;
    /* hotcore.f90 */
    int temp_indx;
    double max_temp;
    /* cshock.f90 module state */
    double vs, timestep_factor, min_postshock_temp, bm0;
    double cs_tout, cs_tsat, cs_dlength, cs_z1, cs_z2, cs_z3, cs_v0, cs_at, cs_z, cs_vn0, cs_zn0, cs_vi0,
        cs_dissipation_time, cs_max_temp, cs_drift_vel, cs_zn, cs_vn, cs_tn, cs_ts, cs_initial_dens_cs,
        cs_grain_number_density, cs_grain_radius_cgs;
    int cs_coflag;
    /* postprocess.f90 module state: tracer history [10][pp_ntime] = time (s), density, gas T, dust T, radfield,
     * zeta, N_H, N_H2, N_CO, N_C; the last four only with pp_coldens */
    int pp_ntime, pp_coldens, pp_tstep;
    const double *pp_grid;
    /* jshock.f90 module state */
    double js_max_temp, js_vmin, js_tshock, js_tcool, js_max_dens, js_t_lambda, js_n_lambda, js_v0;
    /* collapse.f90 module state */
    int collapse_mode;
    double col_max_time, col_parcel_radius, col_mass_in_radius;
    double sput_projectile_abund[6];
    vode_t *vode;
    orc_stats stats;
} orc_model;

extern const double K_BOLTZ, REDUCED_PLANCK, AMU, PI_F, PC, SECONDS_PER_YEAR, MAX_GRAIN_TEMP,
    MIN_SURFACE_ABUND;

double orc_thermal_vel(void);
double orc_gas_dust_density_ratio(void);
double orc_num_sites_per_grain(void);
double orc_bulk_gain_from_mantle_buildup(void);
void orc_init_vdiff(orc_model *m);
void orc_init_desfrac(orc_model *m);
void orc_calculate_reaction_rates(orc_model *m);
double orc_densdot(const orc_model *m, double density);
void orc_rhs(void *ctx, double t, const double *y, double *ydot);

/* orc_cshock.c */
int orc_cshock_initialize(orc_model *m);
void orc_cshock_update_target_time(orc_model *m);
void orc_cshock_update_physics(orc_model *m);
void orc_cshock_sublimation(orc_model *m);
int orc_jshock_initialize(orc_model *m);
void orc_jshock_update_target_time(orc_model *m);
void orc_jshock_update_physics(orc_model *m);
void orc_jshock_sublimation(orc_model *m);

/* orc_collapse.c */
int orc_collapse_initialize(orc_model *m);
void orc_collapse_update_target_time(orc_model *m);
void orc_collapse_update_physics(orc_model *m);

#endif
