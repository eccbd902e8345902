/*
 * ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's hot path (UCLCHEM, Fortran):
 * solveAbundances -> updateChemistry -> calculateReactionRates + DVODE_F90(F)
 * with METHOD_FLAG=22 (BDF, dense finite-difference Jacobian, LINPACK LU) and
 * a cold restart at every output time.  Each function cites the reference
 * file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (uclchem_b200/, include/uclgpu.h) never does.
 *
 * Parity status: PINNED against the reference's own golden trajectories
 * examples/example-output/{static,phase1,phase2}-full.dat (fixtures copied as
 * small numeric tables under tests/golden/, see tests/golden/README.md).  The
 * reference Fortran itself cannot be compiled in this image (no Fortran
 * compiler), so oracle/_ref does not exist.
 */
#ifndef UCLCHEM_ORACLE_H
#define UCLCHEM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NTYPES 23
/* order must match uclchem_b200.network.TYPE_NAMES */
enum orc_type {
    T_PHOTON = 0, T_CRP, T_CRPHOT, T_FREEZE, T_DESORB, T_THERM, T_DESOH2, T_DESCR,
    T_DEUVCR, T_H2FORM, T_ER, T_ERDES, T_LH, T_LHDES, T_BULKSWAP, T_SURFSWAP,
    T_IONOPOL1, T_IONOPOL2, T_CRS, T_EXSOLID, T_EXRELAX, T_GAR, T_TWOBODY
};

/* order must match oracle/oracle.py NAMED */
enum orc_named {
    I_NH = 0, I_NH2, I_NC, I_NCX, I_NO, I_NN, I_NMG, I_NMGX, I_NP, I_NF, I_NNA, I_NLI,
    I_NPAH, I_NSX, I_NSIX, I_NCLX, I_ND, I_NHE, I_N18O, I_N15N, I_N13C, I_NELEC, I_NCO,
    I_NBULK, I_NSURFACE, I_NGN, I_NGO, I_NGOH, I_NSI,
    R_H2FORM_CT, R_H2FORM_ER, R_H2FORM_ERDES, R_HFREEZE, R_EFREEZE, R_H2FREEZE,
    R_H2_HV, R_CO_HV, R_C_HV, R_H2_CRP,
    ORC_NNAMED
};

typedef struct {
    int32_t nspec, nreac, nice, nsurf, ngar, n_loss, n_gain, n_refractory;
    const double *mass;
    const int32_t *atom_counts;
    const int32_t *surface_list, *bulk_list, *ice_list, *gas_ice_list;
    const double *binding_energy, *formation_enthalpy;
    const int32_t *re;  /* [nreac][3] */
    const int32_t *pr;  /* [nreac][4] */
    const double *alpha, *beta, *gama, *min_temps, *max_temps, *reduced_masses;
    const int32_t *extrapolate, *rtype;
    const int32_t *flux_factors; /* [nreac][5] */
    const int32_t *loss_species, *loss_reaction, *gain_species, *gain_reaction;
    const int32_t *freeze_partners; /* [nsurf] */
    const double *gar_params;       /* [ngar][7] */
    const int32_t *type_lo, *type_hi; /* [ORC_NTYPES], -1 = absent */
    const int32_t *named;           /* [ORC_NNAMED] */
    const int32_t *refractory_list;
    const int32_t *is_ion;          /* [nspec] name contains '+' (chemistry.f90:116-121) */
} orc_network;

typedef struct {
    int64_t nst, nfe, nje, nlu, nni, ncfn, netf, nintervals;
} orc_stats;

/* calculateReactionRates at the model's initial state (wrap.f90:446-514 get_rates). */
int orc_get_rates(const orc_network *net, const double *params, const double *y_in /*[nspec]*/,
                  double *rates_out /*[nreac]*/);
/* F(y) at the model's initial physical state (wrap.f90:516-547 get_odes). */
int orc_get_odes(const orc_network *net, const double *params, const double *y_in /*[nspec]*/,
                 double *ydot_out /*[nspec+1]*/);
/* F(y) at exactly the given state (updateChemistry's set-up, chemistry.f90:166-203, then F): kernel-level
 * parity counterpart of uclgpu_probe_rhs */
int orc_probe_rhs(const orc_network *net, const double *params, const double *y_in /*[nspec]*/,
                  double *ydot_out /*[nspec+1]*/);
/* bare GETYDOT (odes.f90:6) for RHS pinning */
void orc_getydot(const orc_network *net, const double *rate, const double *y, double blr,
                 double surface_coverage, double safe_mantle, double safe_bulk, double dens,
                 double *ydot, double *surfgrowth_uncorrected);
/* photoreactions.f90 helpers for unit pinning */
double orc_h2_photo_diss_rate(double nh2, double radfield, double av, double turbvel);
double orc_co_photo_diss_rate(double nh2, double nco, double radfield, double av);

/* One full model (solveAbundances, wrap.f90:549-697).
 * kind: 0 cloud, 1 hot_core, 2 cshock.  y0: [nspec] or NULL.
 * traj arrays may be NULL; layout [timepoints+1][8|nspec|nreac] row-major.
 * Returns the successFlag (constants.f90:22-27).                               */
int orc_run_model(const orc_network *net, int kind, const double *params, const double *y0,
                  double *y_final /*[nspec+1]*/, double *phys_final /*[8]*/, int timepoints,
                  double *phys_traj, double *chem_traj, double *rates_traj, int *nrows,
                  double *dissipation_time, orc_stats *stats);

int orc_run_model_pp(const orc_network *net, int kind, const double *params, const double *y0, double *y_final,
                     double *phys_final, int timepoints, double *phys_traj, double *chem_traj, double *rates_traj,
                     int *nrows, double *dissipation_time, orc_stats *stats, int pp_ntime,
                     const double *pp_grid /*[10][pp_ntime]*/, int pp_coldens);

/* ncell independent models, params [nparam][ncell]; OpenMP over cells. */
int orc_run_grid(const orc_network *net, int kind, int64_t ncell, const double *params,
                 const double *y0 /*[ncell][nspec+1] or NULL*/, double *y_final /*[ncell][nspec+1]*/,
                 double *phys_final /*[ncell][8]*/, int32_t *flag, orc_stats *stats, int nthreads);
int orc_run_grid_timed(const orc_network *net, int kind, int64_t ncell, const double *params, const double *y0,
                       double *y_final, double *phys_final, int32_t *flag, orc_stats *stats, int nthreads,
                       double *cell_seconds /*[ncell] wall time per model, or NULL*/);

#ifdef __cplusplus
}
#endif
#endif
