/* ORACLE (test infrastructure): DVODE MF=22 restatement, see orc_vode.c */
#ifndef ORC_VODE_H
#define ORC_VODE_H

typedef void (*vode_rhs)(void *ctx, double t, const double *y, double *ydot);

typedef struct {
    int n;
    double *yh, *ewt, *savf, *acor, *y, *ftem, *wm, *jsv;
    int *ipvt;
    double tau[14], el[14], tq[6];
    double h, hu, hscal, hnew, tn, rc, prl1, rl1, eta, etamax, crate, drc, acnrm, conp, uround,
        ccmxj, hmxi, hmin, rtol;
    const double *atol;
    int nq, l, lmax, nqwait, newq, newh, jstart, kflag, jcur, icf, ipup, nslp, nslj, msbj, nqu;
    long nst, nfe, nje, nlu, nni, ncfn, netf;
    vode_rhs f;
    void *ctx;
    void *trace; /* FILE* when ORC_TRACE is set (debug) */
} vode_t;

vode_t *vode_alloc(int n);
void vode_free(vode_t *s);
/* Integrate from *t to tout with a cold start (ISTATE=1, ITASK=1). Returns DVODE's ISTATE:
 * 2 success, 1 tout==t, -1 mxstep, -2 too much accuracy, -3 illegal input, -4 error-test
 * failures, -5 convergence failures, -6 EWT<=0.  On failure y,t hold the last good step. */
int vode_solve(vode_t *s, vode_rhs f, void *ctx, double *y, double *t, double tout, double rtol,
               const double *atol, int mxstep);

/* Wall-clock guard for bounded benchmark samples (NOT part of the reference algorithm): after
 * orc_set_deadline(seconds) every model still running `seconds` from now stops with
 * ORC_FLAG_DEADLINE.  0 switches the guard off (default). */
/* Experiment hook (debug only): see orc_vode.c */
typedef int (*orc_la_setup)(void *ctx, const double *y, double gamma, int fresh);
typedef void (*orc_la_solve)(void *ctx, double *b);
void orc_set_linalg_hook(orc_la_setup setup, orc_la_solve solve);
typedef void (*orc_pass_hook)(void *ctx, int phase);
void orc_set_pass_hook(orc_pass_hook h);

#define ORC_FLAG_DEADLINE (-98)
void orc_set_deadline(double seconds_from_now);
int orc_deadline_expired(void);

#endif
