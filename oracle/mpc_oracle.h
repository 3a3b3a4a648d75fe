/*
 * mpc_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's MPC hot path
 *   MPC::solve            /root/reference/src/control/MPC.cpp:183-325
 *   FG_eval::operator()   /root/reference/src/control/MPC.cpp:50-154
 *   CppAD::ipopt::solve + Ipopt + MUMPS (third party, NOT under /root/reference;
 *   Ipopt >= 3.12.7 per install_Ipopt_CppAD.md:10,22; CppAD unpinned)
 *
 * PARITY UNPINNED: the reference ships no numeric golden vectors for this path and
 * Ipopt/CppAD/MUMPS are absent from this image, so the interior-point iteration below is a
 * restatement of Ipopt's published algorithm (Waechter & Biegler 2006 + the 3.12 defaults listed
 * in SURVEY.md App. B.2) with a dense Bunch-Kaufman LDL^T in place of MUMPS.  It is pinned only
 * by (i) an independent SciPy solve of the same NLP, (ii) KKT certificates, (iii) the reference's
 * own FG_eval text compiled against an AD shim (oracle/ref_shim -> oracle/_ref), see DESIGN.md.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * use anything in this directory.
 */
#ifndef MPC_ORACLE_H
#define MPC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NMAX 64          /* maximum horizon length */
#define ORC_NCOEF 5          /* polynomial coefficients, low -> high order, zero padded */
#define ORC_NTAB 16

/* Config::* statics after Config::load's unit conversions (Config.cpp:31-87), SI units. */
typedef struct orc_config {
  int N;                     /* Config::N */
  int n_steers;              /* Config::steers.size() */
  int n_steer_speeds;        /* Config::steerSpeeds.size() */
  int max_iter;              /* Ipopt max_iter (default 3000) */
  double dt, Lf;
  double cte_panic, epsi_panic;
  double max_speed, max_steering, max_accel, max_decel;
  double weights[12];        /* Config::weights, indices Config.h:14-61 */
  double steers[ORC_NTAB];
  double steer_speeds[ORC_NTAB];
  double tol;                /* Ipopt tol (default 1e-8) */
  double mu_init;            /* 0.1 */
  int max_soc;               /* Ipopt max_soc (default 4) */
  int obj_scaling;           /* 1 = gradient-based nlp scaling (Ipopt default) */
  int watchdog_trigger;      /* Ipopt watchdog_shortened_iter_trigger (default 10, 0 = off) */
  int filter_reset_trigger;  /* Ipopt filter_reset_trigger (default 5) */
  double tiny_step_tol;      /* Ipopt tiny_step_tol (default 10 eps, 0 = off) */
} orc_config;

/* One problem: MPC::solve's explicit and hidden inputs (MPC.cpp:213-218, 229-232, RoadGeometry). */
typedef struct orc_problem {
  double state[6];           /* x, y, psi, v, cte, epsi */
  double coeffs[ORC_NCOEF];  /* roadGeometry.polynomial, low -> high */
  double yaw_lo, yaw_hi;     /* Config::yawLow / yawHigh */
} orc_problem;

/* CppAD::ipopt::solve_result::status_type integers (printed by MPC.cpp:301). */
enum {
  ORC_NOT_DEFINED = 0, ORC_SUCCESS = 1, ORC_MAXITER_EXCEEDED = 2, ORC_STOP_AT_TINY_STEP = 3,
  ORC_STOP_AT_ACCEPTABLE_POINT = 4, ORC_LOCAL_INFEASIBILITY = 5, ORC_USER_REQUESTED_STOP = 6,
  ORC_FEASIBLE_POINT_FOUND = 7, ORC_DIVERGING_ITERATES = 8, ORC_RESTORATION_FAILURE = 9,
  ORC_ERROR_IN_STEP_COMPUTATION = 10, ORC_INVALID_NUMBER_DETECTED = 11,
  ORC_TOO_FEW_DEGREES_OF_FREEDOM = 12, ORC_INTERNAL_ERROR = 13, ORC_UNKNOWN = 14
};

typedef struct orc_result {
  int status;                /* enum above */
  int iters;                 /* interior-point iterations */
  int n_regularized;         /* iterations that needed inertia correction */
  int n_soc;                 /* accepted second-order-correction steps */
  int n_backtrack;           /* total line-search halvings */
  double obj;                /* unscaled objective at the solution (solution.obj_value) */
  double result[9];          /* MPC.cpp:322-324: x1,y1,psi1,v1,cte1,epsi1,delta0,a0,cost */
  double kkt_error;          /* final scaled E_0 */
  double z[8 * ORC_NMAX];    /* solution.x in the reference's layout (MPC.cpp:189-196) */
  double lambda[6 * ORC_NMAX];
  double zl[8 * ORC_NMAX], zu[8 * ORC_NMAX];
  int n_resto, n_resto_iter, n_watchdog, n_tiny, n_filter_reset, n_filter_max;   /* see orc_ipm_stats */
} orc_result;

void orc_config_defaults(orc_config *cfg);   /* solver knobs only (Ipopt 3.12 defaults) */

/* A generic equality-constrained, bound-constrained NLP (dense derivatives) for the interior-point
 * core.  Used twice: by the analytic restatement of FG_eval below, and by oracle/ref_shim, where
 * f, g and their derivatives come from the reference's OWN FG_eval text through an AD tape. */
typedef struct orc_nlp {
  int n, m;
  void *user;
  double (*f)(void *user, const double *x);
  void (*grad)(void *user, const double *x, double *g);                 /* n */
  void (*g)(void *user, const double *x, double *c);                    /* m */
  void (*jac)(void *user, const double *x, double *J);                  /* m x n row-major */
  void (*hess)(void *user, const double *x, double sigma, const double *lam, double *H);  /* n x n */
} orc_nlp;
typedef struct orc_ipm_stats {
  int status, iters, n_regularized, n_soc, n_backtrack;
  double obj, kkt_error;
  int n_resto;               /* calls of the feasibility restoration phase */
  int n_resto_iter;          /* iterations spent inside it (included in iters) */
  int n_watchdog;            /* watchdog activations */
  int n_tiny;                /* tiny steps taken */
  int n_filter_reset;        /* filter resets by the heuristic */
  int n_filter_max;          /* largest number of filter entries held at once */
} orc_ipm_stats;
/* cfg supplies only tol, max_iter, mu_init, max_soc, obj_scaling.  Returns 0, or -1 if gl != gu. */
int orc_ipm_solve(const orc_nlp *nlp, const orc_config *cfg, const double *xi, const double *xl,
                  const double *xu, const double *gl, const double *gu, double *x_out, double *lam_out,
                  double *zl_out, double *zu_out, orc_ipm_stats *stats);

/* NLP pieces, reference layout; all UNSCALED.  hess is n x n dense row-major, jac m x n. */
double orc_eval_f(const orc_config *cfg, const orc_problem *p, const double *z);
void orc_eval_grad(const orc_config *cfg, const orc_problem *p, const double *z, double *grad);
void orc_eval_g(const orc_config *cfg, const orc_problem *p, const double *z, double *g);
void orc_eval_jac(const orc_config *cfg, const orc_problem *p, const double *z, double *jac);
void orc_eval_hess(const orc_config *cfg, const orc_problem *p, const double *z, double sigma,
                   const double *lambda, double *hess);
void orc_bounds(const orc_config *cfg, const orc_problem *p, double *xl, double *xu,
                double *gl, double *gu, double *xi);
/* frozen-branch constants decided at the start point: arrays of N */
void orc_frozen(const orc_config *cfg, const orc_problem *p, double *wc, double *we,
                double *vref, double *nvw);

int orc_solve(const orc_config *cfg, const orc_problem *p, orc_result *out);
/* B problems over n_threads host threads (static contiguous chunks); returns 0 */
int orc_solve_batch(const orc_config *cfg, const orc_problem *p, int B, orc_result *out,
                    int n_threads);
/* compact batch: result9 [B][9], status/iters [B]; no per-problem orc_result kept */
int orc_solve_batch_compact(const orc_config *cfg, const orc_problem *p, int B, double *result9,
                            double *traj_x, double *traj_y, int *status, int *iters,
                            int n_threads);

#ifdef __cplusplus
}
#endif
#endif
