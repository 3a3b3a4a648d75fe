/*
 * mpc_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See mpc_oracle.h.
 *
 * Part 1: the NLP of FG_eval (MPC.cpp:50-154) with CppAD's record-once semantics: every `if` on
 *         an AD value is decided at the start point xi (MPC.cpp:207-218) and frozen.
 * Part 2: dense symmetric-indefinite LDL^T (Bunch-Kaufman) standing in for MUMPS (MPC.cpp:175).
 * Part 3: Ipopt's primal-dual interior-point filter line-search iteration (restated from the
 *         published algorithm; defaults of Ipopt 3.12).
 */
#include "mpc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* Part 1: NLP                                                                                */
/* ------------------------------------------------------------------------------------------ */

/* variable offsets, MPC.cpp:56-63 / 189-196 */
#define IX(i) (i)
#define IY(i) (N + (i))
#define IPSI(i) (2 * N + (i))
#define IV(i) (3 * N + (i))
#define IC(i) (4 * N + (i))
#define IE(i) (5 * N + (i))
#define ID(i) (6 * N + (i))
#define IA(i) (7 * N - 1 + (i))

#define MACH_EPS_ 2.220446049250313e-16
void orc_config_defaults(orc_config *cfg) {
  cfg->max_iter = 3000;
  cfg->tol = 1e-8;
  cfg->mu_init = 0.1;
  cfg->max_soc = 4;
  cfg->obj_scaling = 1;
  cfg->watchdog_trigger = 10;
  cfg->filter_reset_trigger = 5;
  cfg->tiny_step_tol = 10.0 * MACH_EPS_;
}

/* Vehicle::computeSpeedTarget(AD, double), Vehicle.cpp:50-64 (AD flavour: '<' select, no fmin) */
static double speed_target(const orc_config *cfg, double angle, double max) {
  double y = fabs(angle);
  double back = cfg->steer_speeds[cfg->n_steer_speeds - 1];
  for (int i = 0; i < cfg->n_steers; i++) {
    if (y <= cfg->steers[i]) {
      if (cfg->n_steer_speeds > i) return cfg->steer_speeds[i] < max ? cfg->steer_speeds[i] : max;
      return back < max ? back : max;
    }
  }
  return back < max ? back : max;
}

/* Branch outcomes of MPC.cpp:72,79,87,89 at xi (zeros except stage 0 = state).  The actuator
 * branches MPC.cpp:98,103,111 compare zeros at xi and are therefore always false: the terms
 * WEIGHT_A, WEIGHT_DECEL_LOW_V and WEIGHT_DA never enter the recorded objective. */
void orc_frozen(const orc_config *cfg, const orc_problem *p, double *wc, double *we, double *vref,
                double *nvw) {
  const double *w = cfg->weights;
  for (int i = 0; i < cfg->N; i++) {
    double cte = i ? 0.0 : p->state[4], epsi = i ? 0.0 : p->state[5];
    double psi = i ? 0.0 : p->state[2], v = i ? 0.0 : p->state[3];
    wc[i] = (fabs(cte) < cfg->cte_panic) ? w[0] : w[11];
    we[i] = (fabs(epsi) > cfg->epsi_panic) ? w[10] : w[1];
    vref[i] = speed_target(cfg, psi, cfg->max_speed);
    nvw[i] = (v < 0) ? w[9] : 0.0;
  }
}

/* polyeval / polyder, utils.h:28-47, plus second and third derivative */
static void poly4(const double *c, double x, double *f, double *f1, double *f2, double *f3) {
  double r = 0;
  for (int i = ORC_NCOEF - 1; i >= 0; i--) r = r * x + c[i];
  *f = r;
  r = 0;
  for (int i = ORC_NCOEF - 1; i >= 1; i--) r = r * x + i * c[i];
  *f1 = r;
  r = 0;
  for (int i = ORC_NCOEF - 1; i >= 2; i--) r = r * x + (double)(i * (i - 1)) * c[i];
  *f2 = r;
  r = 0;
  for (int i = ORC_NCOEF - 1; i >= 3; i--) r = r * x + (double)(i * (i - 1) * (i - 2)) * c[i];
  *f3 = r;
}

double orc_eval_f(const orc_config *cfg, const orc_problem *p, const double *z) {
  int N = cfg->N;
  double wc[ORC_NMAX], we[ORC_NMAX], vref[ORC_NMAX], nvw[ORC_NMAX];
  const double *w = cfg->weights;
  orc_frozen(cfg, p, wc, we, vref, nvw);
  double f = 0;
  for (int i = 0; i < N; i++) { /* MPC.cpp:71-92 */
    f += z[IC(i)] * z[IC(i)] * wc[i];
    f += z[IE(i)] * z[IE(i)] * we[i];
    double dv = z[IV(i)] - vref[i];
    f += dv * dv * w[2];
    f += z[IV(i)] * z[IV(i)] * nvw[i];
  }
  for (int i = 0; i < N - 1; i++) f += z[ID(i)] * z[ID(i)] * w[3]; /* MPC.cpp:95-96 */
  for (int i = 0; i < N - 2; i++) {                                /* MPC.cpp:109-110 */
    double dd = z[ID(i + 1)] - z[ID(i)];
    f += dd * dd * w[4];
  }
  return f;
}

void orc_eval_grad(const orc_config *cfg, const orc_problem *p, const double *z, double *g) {
  int N = cfg->N, n = 8 * N - 2;
  double wc[ORC_NMAX], we[ORC_NMAX], vref[ORC_NMAX], nvw[ORC_NMAX];
  const double *w = cfg->weights;
  orc_frozen(cfg, p, wc, we, vref, nvw);
  memset(g, 0, sizeof(double) * n);
  for (int i = 0; i < N; i++) {
    g[IC(i)] = 2 * wc[i] * z[IC(i)];
    g[IE(i)] = 2 * we[i] * z[IE(i)];
    g[IV(i)] = 2 * w[2] * (z[IV(i)] - vref[i]) + 2 * nvw[i] * z[IV(i)];
  }
  for (int i = 0; i < N - 1; i++) g[ID(i)] = 2 * w[3] * z[ID(i)];
  for (int i = 0; i < N - 2; i++) {
    double dd = z[ID(i + 1)] - z[ID(i)];
    g[ID(i + 1)] += 2 * w[4] * dd;
    g[ID(i)] -= 2 * w[4] * dd;
  }
}

/* fg[1..] of FG_eval, MPC.cpp:116-153 */
void orc_eval_g(const orc_config *cfg, const orc_problem *p, const double *z, double *g) {
  int N = cfg->N;
  double dt = cfg->dt, Lf = cfg->Lf;
  g[IX(0)] = z[IX(0)];
  g[IY(0)] = z[IY(0)];
  g[IPSI(0)] = z[IPSI(0)];
  g[IV(0)] = z[IV(0)];
  g[IC(0)] = z[IC(0)];
  g[IE(0)] = z[IE(0)];
  for (int i = 1; i < N; i++) {
    double x0 = z[IX(i - 1)], y0 = z[IY(i - 1)], psi0 = z[IPSI(i - 1)], v0 = z[IV(i - 1)];
    double epsi0 = z[IE(i - 1)], delta0 = z[ID(i - 1)], a0 = z[IA(i - 1)];
    double vdt = v0 * dt;
    double psi = psi0 + delta0 * vdt / Lf;
    double f, f1, f2, f3;
    poly4(p->coeffs, x0, &f, &f1, &f2, &f3);
    g[IX(i)] = z[IX(i)] - (x0 + cos(psi0) * vdt);
    g[IY(i)] = z[IY(i)] - (y0 + sin(psi0) * vdt);
    g[IPSI(i)] = z[IPSI(i)] - psi;
    g[IV(i)] = z[IV(i)] - (v0 + a0 * dt);
    g[IC(i)] = z[IC(i)] - ((f - y0) + sin(epsi0) * vdt);
    g[IE(i)] = z[IE(i)] - (psi - atan(f1));
  }
}

/* dense Jacobian (m x n row-major); analytic entries of SURVEY.md App. A.4 */
void orc_eval_jac(const orc_config *cfg, const orc_problem *p, const double *z, double *J) {
  int N = cfg->N, n = 8 * N - 2, m = 6 * N;
  double dt = cfg->dt, Lf = cfg->Lf;
  memset(J, 0, sizeof(double) * (size_t)m * n);
#define JJ(r, c) J[(size_t)(r) * n + (c)]
  JJ(IX(0), IX(0)) = 1;
  JJ(IY(0), IY(0)) = 1;
  JJ(IPSI(0), IPSI(0)) = 1;
  JJ(IV(0), IV(0)) = 1;
  JJ(IC(0), IC(0)) = 1;
  JJ(IE(0), IE(0)) = 1;
  for (int i = 1; i < N; i++) {
    int j = i - 1;
    double x0 = z[IX(j)], psi0 = z[IPSI(j)], v0 = z[IV(j)], epsi0 = z[IE(j)], delta0 = z[ID(j)];
    double f, f1, f2, f3;
    poly4(p->coeffs, x0, &f, &f1, &f2, &f3);
    JJ(IX(i), IX(i)) = 1;
    JJ(IX(i), IX(j)) = -1;
    JJ(IX(i), IPSI(j)) = dt * v0 * sin(psi0);
    JJ(IX(i), IV(j)) = -dt * cos(psi0);
    JJ(IY(i), IY(i)) = 1;
    JJ(IY(i), IY(j)) = -1;
    JJ(IY(i), IPSI(j)) = -dt * v0 * cos(psi0);
    JJ(IY(i), IV(j)) = -dt * sin(psi0);
    JJ(IPSI(i), IPSI(i)) = 1;
    JJ(IPSI(i), IPSI(j)) = -1;
    JJ(IPSI(i), IV(j)) = -delta0 * dt / Lf;
    JJ(IPSI(i), ID(j)) = -dt * v0 / Lf;
    JJ(IV(i), IV(i)) = 1;
    JJ(IV(i), IV(j)) = -1;
    JJ(IV(i), IA(j)) = -dt;
    JJ(IC(i), IC(i)) = 1;
    JJ(IC(i), IX(j)) = -f1;
    JJ(IC(i), IY(j)) = 1;
    JJ(IC(i), IV(j)) = -dt * sin(epsi0);
    JJ(IC(i), IE(j)) = -dt * v0 * cos(epsi0);
    JJ(IE(i), IE(i)) = 1;
    JJ(IE(i), IX(j)) = f2 / (1 + f1 * f1);
    JJ(IE(i), IPSI(j)) = -1;
    JJ(IE(i), IV(j)) = -delta0 * dt / Lf;
    JJ(IE(i), ID(j)) = -dt * v0 / Lf;
  }
#undef JJ
}

/* dense Hessian of sigma*f + lambda^T g (n x n row-major, full symmetric) */
void orc_eval_hess(const orc_config *cfg, const orc_problem *p, const double *z, double sigma,
                   const double *lam, double *H) {
  int N = cfg->N, n = 8 * N - 2;
  double dt = cfg->dt, Lf = cfg->Lf;
  double wc[ORC_NMAX], we[ORC_NMAX], vref[ORC_NMAX], nvw[ORC_NMAX];
  const double *w = cfg->weights;
  orc_frozen(cfg, p, wc, we, vref, nvw);
  memset(H, 0, sizeof(double) * (size_t)n * n);
#define HH(r, c) H[(size_t)(r) * n + (c)]
#define HS(r, c, val)     \
  do {                    \
    double v__ = (val);   \
    HH(r, c) += v__;      \
    if ((r) != (c)) HH(c, r) += v__; \
  } while (0)
  for (int i = 0; i < N; i++) {
    HS(IC(i), IC(i), sigma * 2 * wc[i]);
    HS(IE(i), IE(i), sigma * 2 * we[i]);
    HS(IV(i), IV(i), sigma * 2 * (w[2] + nvw[i]));
  }
  for (int i = 0; i < N - 1; i++) HS(ID(i), ID(i), sigma * 2 * w[3]);
  for (int i = 0; i < N - 2; i++) {
    HS(ID(i), ID(i), sigma * 2 * w[4]);
    HS(ID(i + 1), ID(i + 1), sigma * 2 * w[4]);
    HS(ID(i), ID(i + 1), -sigma * 2 * w[4]);
  }
  for (int i = 1; i < N; i++) {
    int j = i - 1;
    double x0 = z[IX(j)], psi0 = z[IPSI(j)], v0 = z[IV(j)], epsi0 = z[IE(j)];
    double f, f1, f2, f3;
    poly4(p->coeffs, x0, &f, &f1, &f2, &f3);
    double q = 1 + f1 * f1;
    double lx = lam[IX(i)], ly = lam[IY(i)], lp = lam[IPSI(i)], lc = lam[IC(i)], le = lam[IE(i)];
    /* row x  */
    HS(IPSI(j), IPSI(j), lx * dt * v0 * cos(psi0));
    HS(IPSI(j), IV(j), lx * dt * sin(psi0));
    /* row y */
    HS(IPSI(j), IPSI(j), ly * dt * v0 * sin(psi0));
    HS(IPSI(j), IV(j), -ly * dt * cos(psi0));
    /* row psi */
    HS(IV(j), ID(j), -lp * dt / Lf);
    /* row cte */
    HS(IX(j), IX(j), -lc * f2);
    HS(IV(j), IE(j), -lc * dt * cos(epsi0));
    HS(IE(j), IE(j), lc * dt * v0 * sin(epsi0));
    /* row epsi */
    HS(IV(j), ID(j), -le * dt / Lf);
    HS(IX(j), IX(j), le * (f3 * q - 2 * f1 * f2 * f2) / (q * q));
  }
#undef HS
#undef HH
}

/* MPC.cpp:204-281: start point, variable bounds, constraint bounds */
void orc_bounds(const orc_config *cfg, const orc_problem *p, double *xl, double *xu, double *gl,
                double *gu, double *xi) {
  int N = cfg->N, n = 8 * N - 2, m = 6 * N;
  for (int i = 0; i < n; i++) xi[i] = 0;
  xi[IX(0)] = p->state[0];
  xi[IY(0)] = p->state[1];
  xi[IPSI(0)] = p->state[2];
  xi[IV(0)] = p->state[3];
  xi[IC(0)] = p->state[4];
  xi[IE(0)] = p->state[5];
  for (int i = 0; i < N; i++) {
    xl[IX(i)] = xl[IY(i)] = xl[IC(i)] = xl[IE(i)] = -1.0e19;
    xu[IX(i)] = xu[IY(i)] = xu[IC(i)] = xu[IE(i)] = 1.0e19;
    xl[IPSI(i)] = p->yaw_lo;
    xu[IPSI(i)] = p->yaw_hi;
    xl[IV(i)] = -cfg->max_speed;
    xu[IV(i)] = cfg->max_speed;
  }
  for (int i = 0; i < N - 1; i++) {
    xl[ID(i)] = -cfg->max_steering;
    xu[ID(i)] = cfg->max_steering;
    xl[IA(i)] = cfg->max_decel;
    xu[IA(i)] = cfg->max_accel;
  }
  for (int i = 0; i < m; i++) gl[i] = gu[i] = 0;
  gl[IX(0)] = gu[IX(0)] = p->state[0];
  gl[IY(0)] = gu[IY(0)] = p->state[1];
  gl[IPSI(0)] = gu[IPSI(0)] = p->state[2];
  gl[IV(0)] = gu[IV(0)] = p->state[3];
  gl[IC(0)] = gu[IC(0)] = p->state[4];
  gl[IE(0)] = gu[IE(0)] = p->state[5];
}

/* ------------------------------------------------------------------------------------------ */
/* Part 2: dense Bunch-Kaufman LDL^T with inertia (stand-in for MUMPS)                         */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  int n;
  double *A;   /* working matrix, full storage n x n */
  double *L;   /* unit lower factor */
  double *D;   /* block diagonal: D[3k], D[3k+1] (sub), D[3k+2] unused */
  int *perm;   /* perm[k] = original index at position k */
  int *blk;    /* blk[k] = 1: 1x1 pivot at k; 2: first row of a 2x2 pivot; 0: second row */
  double *tmp;
  int n_neg, n_zero;
} ldl_t;

static ldl_t *ldl_new(int n) {
  ldl_t *f = (ldl_t *)calloc(1, sizeof(ldl_t));
  f->n = n;
  f->A = (double *)malloc(sizeof(double) * (size_t)n * n);
  f->L = (double *)malloc(sizeof(double) * (size_t)n * n);
  f->D = (double *)malloc(sizeof(double) * 3 * (size_t)n);
  f->perm = (int *)malloc(sizeof(int) * n);
  f->blk = (int *)malloc(sizeof(int) * n);
  f->tmp = (double *)malloc(sizeof(double) * n);
  return f;
}
static void ldl_free(ldl_t *f) {
  free(f->A); free(f->L); free(f->D); free(f->perm); free(f->blk); free(f->tmp); free(f);
}

static void ldl_swap(ldl_t *f, int a, int b) {
  if (a == b) return;
  int n = f->n;
  double *A = f->A, *L = f->L;
  for (int j = 0; j < n; j++) { double t = A[(size_t)a * n + j]; A[(size_t)a * n + j] = A[(size_t)b * n + j]; A[(size_t)b * n + j] = t; }
  for (int i = 0; i < n; i++) { double t = A[(size_t)i * n + a]; A[(size_t)i * n + a] = A[(size_t)i * n + b]; A[(size_t)i * n + b] = t; }
  for (int j = 0; j < n; j++) { double t = L[(size_t)a * n + j]; L[(size_t)a * n + j] = L[(size_t)b * n + j]; L[(size_t)b * n + j] = t; }
  int t = f->perm[a]; f->perm[a] = f->perm[b]; f->perm[b] = t;
}

/* factor the symmetric matrix in f->A (destroyed) */
static void ldl_factor(ldl_t *f) {
  int n = f->n;
  double *A = f->A, *L = f->L;
  const double alpha = (1.0 + sqrt(17.0)) / 8.0;
  memset(L, 0, sizeof(double) * (size_t)n * n);
  for (int i = 0; i < n; i++) { f->perm[i] = i; f->blk[i] = 1; }
  f->n_neg = f->n_zero = 0;
#define AA(i, j) A[(size_t)(i) * n + (j)]
#define LL(i, j) L[(size_t)(i) * n + (j)]
  int k = 0;
  while (k < n) {
    int kstep = 1, kp = k;
    double absakk = fabs(AA(k, k));
    int imax = k;
    double colmax = 0;
    for (int i = k + 1; i < n; i++)
      if (fabs(AA(i, k)) > colmax) { colmax = fabs(AA(i, k)); imax = i; }
    if (fmax(absakk, colmax) == 0.0) {
      /* zero column: singular pivot */
      f->n_zero++;
      f->D[3 * k] = 0.0;
      LL(k, k) = 1;
      f->blk[k] = 1;
      k++;
      continue;
    }
    if (absakk >= alpha * colmax) {
      kp = k;
    } else {
      double rowmax = 0;
      for (int j = k; j < n; j++)
        if (j != imax && fabs(AA(imax, j)) > rowmax) rowmax = fabs(AA(imax, j));
      if (absakk >= alpha * colmax * (colmax / rowmax)) kp = k;
      else if (fabs(AA(imax, imax)) >= alpha * rowmax) kp = imax;
      else { kp = imax; kstep = 2; }
    }
    if (kstep == 1) {
      ldl_swap(f, k, kp);
      double d = AA(k, k);
      f->D[3 * k] = d;
      f->blk[k] = 1;
      if (d < 0) f->n_neg++;
      LL(k, k) = 1;
      double r = 1.0 / d;
      for (int i = k + 1; i < n; i++) LL(i, k) = AA(i, k) * r;
      for (int i = k + 1; i < n; i++) {
        double lik = LL(i, k);
        if (lik == 0.0) continue;
        for (int j = k + 1; j < n; j++) AA(i, j) -= lik * AA(k, j);
      }
      k += 1;
    } else {
      ldl_swap(f, k + 1, kp);
      double d11 = AA(k, k), d21 = AA(k + 1, k), d22 = AA(k + 1, k + 1);
      f->D[3 * k] = d11;
      f->D[3 * k + 1] = d21;
      f->D[3 * (k + 1)] = d22;
      f->blk[k] = 2;
      f->blk[k + 1] = 0;
      double det = d11 * d22 - d21 * d21;
      /* BK 2x2 pivots are indefinite: one positive, one negative eigenvalue */
      if (det < 0) f->n_neg += 1;
      else if (det == 0) f->n_zero += 1;
      else if (d11 < 0) f->n_neg += 2;
      LL(k, k) = 1;
      LL(k + 1, k + 1) = 1;
      for (int i = k + 2; i < n; i++) {
        double a1 = AA(i, k), a2 = AA(i, k + 1);
        LL(i, k) = (a1 * d22 - a2 * d21) / det;
        LL(i, k + 1) = (a2 * d11 - a1 * d21) / det;
      }
      for (int i = k + 2; i < n; i++) {
        double l1 = LL(i, k), l2 = LL(i, k + 1);
        if (l1 == 0.0 && l2 == 0.0) continue;
        for (int j = k + 2; j < n; j++) AA(i, j) -= l1 * AA(k, j) + l2 * AA(k + 1, j);
      }
      k += 2;
    }
  }
#undef AA
#undef LL
}

/* solve (P^T L D L^T P) x = b, in place */
static void ldl_solve(const ldl_t *f, double *b) {
  int n = f->n;
  const double *L = f->L;
  double *y = f->tmp;
  for (int i = 0; i < n; i++) y[i] = b[f->perm[i]];
  for (int i = 0; i < n; i++) {
    double s = y[i];
    for (int j = 0; j < i; j++) s -= L[(size_t)i * n + j] * y[j];
    y[i] = s;
  }
  for (int k = 0; k < n;) {
    if (f->blk[k] == 2) {
      double d11 = f->D[3 * k], d21 = f->D[3 * k + 1], d22 = f->D[3 * (k + 1)];
      double det = d11 * d22 - d21 * d21;
      double y1 = y[k], y2 = y[k + 1];
      y[k] = (y1 * d22 - y2 * d21) / det;
      y[k + 1] = (y2 * d11 - y1 * d21) / det;
      k += 2;
    } else {
      y[k] = y[k] / f->D[3 * k];
      k += 1;
    }
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = y[i];
    for (int j = i + 1; j < n; j++) s -= L[(size_t)j * n + i] * y[j];
    y[i] = s;
  }
  for (int i = 0; i < n; i++) b[f->perm[i]] = y[i];
}

/* ------------------------------------------------------------------------------------------ */
/* Part 3: interior-point iteration (Ipopt 3.12 defaults)                                     */
/*                                                                                            */
/* Restated FROM MEMORY of the published algorithm (Waechter & Biegler 2006) and of Ipopt     */
/* 3.12's sources -- none of which is under /root/reference or in this image:                 */
/*   IpIpoptAlg / IpMonotoneMuUpdate / IpOptErrorConvCheck   main loop, mu rule, termination  */
/*   IpBacktrackingLineSearch   backtracking, watchdog, tiny steps, soft restoration          */
/*   IpFilterLSAcceptor         filter tests, second-order correction, filter reset heuristic */
/*   IpPDPerturbationHandler    inertia correction schedule                                   */
/*   IpRestoMinC_1Nrm, IpRestoIpoptNLP, IpRestoIterateInitializer, IpRestoConvCheck,          */
/*   IpRestoFilterConvCheck     feasibility restoration phase                                 */
/* Deliberate simplifications, each marked below: the nested restoration solve has no         */
/* watchdog of its own; the soft restoration step and RestoreAcceptablePoint are not restated.*/
/* ------------------------------------------------------------------------------------------ */

/* Ipopt constants (IpoptAlg defaults) */
#define KAPPA_EPS 10.0          /* barrier_tol_factor */
#define KAPPA_MU 0.2            /* mu_linear_decrease_factor */
#define THETA_MU 1.5            /* mu_superlinear_decrease_power */
#define TAU_MIN 0.99
#define KAPPA_1 0.01            /* bound_push */
#define KAPPA_2 0.01            /* bound_frac */
#define BOUND_RELAX 1e-8
#define S_MAX 100.0
#define KAPPA_SIGMA 1e10
#define GAMMA_THETA 1e-5
#define GAMMA_PHI 1e-8
#define DELTA_LS 1.0
#define S_THETA 1.1
#define S_PHI 2.3
#define ETA_PHI 1e-8
#define KAPPA_SOC 0.99
#define ALPHA_MIN_FRAC 0.05
#define ALPHA_RED 0.5
#define OBJ_MAX_INC 5.0
#define DW_FIRST 1e-4           /* first_hessian_perturbation */
#define DW_MIN 1e-20
#define DW_MAX 1e20             /* max_hessian_perturbation */
#define DW_INC_FIRST 100.0      /* perturb_inc_fact_first */
#define DW_INC 8.0              /* perturb_inc_fact */
#define DW_DEC (1.0 / 3.0)      /* perturb_dec_fact */
#define DUAL_INF_TOL 1.0
#define CONSTR_VIOL_TOL 1e-4
#define COMPL_INF_TOL 1e-4
#define ACCEPT_TOL 1e-6
#define ACCEPT_ITER 15
#define ACCEPT_DUAL_INF_TOL 1e10
#define ACCEPT_CONSTR_VIOL_TOL 1e-2
#define ACCEPT_COMPL_INF_TOL 1e-2
#define CONSTR_MULT_INIT_MAX 1e3
#define NLP_INF 1e19
#define FILTER_MAX 128
#define MACH_EPS 2.220446049250313e-16
#define MAX_FILTER_RESETS 5
#define WATCHDOG_TRIAL_MAX 3    /* watchdog_trial_iter_max */
#define TINY_STEP_Y_TOL 1e-2
#define RESTO_RHO 1000.0        /* resto_penalty_parameter */
#define RESTO_KAPPA 0.9         /* required_infeasibility_reduction */
#define RESTO_THETA_MAX_FACT 1e8
#define BOUND_MULT_RESET_THRESHOLD 1e3

typedef struct {
  const orc_config *cfg;     /* solver options only (tol, max_iter, mu_init, max_soc, obj_scaling) */
  const orc_nlp *nlp;
  int n, m;
  double sf;           /* objective scaling factor */
  double *cs;          /* constraint scaling (m) */
  double *xl, *xu;     /* relaxed bounds (n), +-inf as +-HUGE_VAL */
  double *xl0, *xu0;   /* original bounds */
  int *has_l, *has_u;
  double *gl;
  /* iterate */
  double *x, *lam, *zl, *zu;
  /* work */
  double *grad, *c, *J, *H, *rhs, *dx, *dlam, *dzl, *dzu, *xt, *ct, *csoc;
  ldl_t *kkt;
  int nzl, nzu;
  int trace;
} ipm_t;

static double nlp_f(const ipm_t *s, const double *x) { return s->sf * s->nlp->f(s->nlp->user, x); }
static void nlp_c(const ipm_t *s, const double *x, double *c) {
  s->nlp->g(s->nlp->user, x, c);
  for (int j = 0; j < s->m; j++) c[j] = (c[j] - s->gl[j]) * s->cs[j];
}
static void nlp_grad(const ipm_t *s, const double *x, double *g) {
  s->nlp->grad(s->nlp->user, x, g);
  for (int i = 0; i < s->n; i++) g[i] *= s->sf;
}
static void nlp_jac(const ipm_t *s, const double *x, double *J) {
  s->nlp->jac(s->nlp->user, x, J);
  for (int j = 0; j < s->m; j++)
    if (s->cs[j] != 1.0)
      for (int i = 0; i < s->n; i++) J[(size_t)j * s->n + i] *= s->cs[j];
}

static double log_slacks(const ipm_t *s, const double *x) {
  double sl = 0;
  for (int i = 0; i < s->n; i++) {
    if (s->has_l[i]) sl += log(x[i] - s->xl[i]);
    if (s->has_u[i]) sl += log(s->xu[i] - x[i]);
  }
  return sl;
}
static double barrier_phi(const ipm_t *s, const double *x, double mu) { return nlp_f(s, x) - mu * log_slacks(s, x); }

static double norm1(const double *v, int n) { double s = 0; for (int i = 0; i < n; i++) s += fabs(v[i]); return s; }
static double norminf(const double *v, int n) { double s = 0; for (int i = 0; i < n; i++) if (fabs(v[i]) > s) s = fabs(v[i]); return s; }

/* assemble and factor [[H + Sigma + dw I, J^T], [J, -(dc I + Dc)]]; returns 1 if inertia is (n, m, 0).
 * zl/zu: the bound multipliers in Sigma; Dc: extra (2,2) diagonal (restoration phase) or NULL */
static int kkt_factor(ipm_t *s, const double *x, const double *zl, const double *zu, double dw, double dc,
                      const double *Dc) {
  int n = s->n, m = s->m, d = n + m;
  double *A = s->kkt->A;
  memset(A, 0, sizeof(double) * (size_t)d * d);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) A[(size_t)i * d + j] = s->H[(size_t)i * n + j];
    double sig = 0;
    if (s->has_l[i]) sig += zl[i] / (x[i] - s->xl[i]);
    if (s->has_u[i]) sig += zu[i] / (s->xu[i] - x[i]);
    A[(size_t)i * d + i] += sig + dw;
  }
  for (int j = 0; j < m; j++) {
    for (int i = 0; i < n; i++) {
      double v = s->J[(size_t)j * n + i];
      A[(size_t)(n + j) * d + i] = v;
      A[(size_t)i * d + (n + j)] = v;
    }
    A[(size_t)(n + j) * d + (n + j)] = -(dc + (Dc ? Dc[j] : 0.0));
  }
  ldl_factor(s->kkt);
  return (s->kkt->n_zero == 0 && s->kkt->n_neg == m);
}

/* Ipopt's inertia correction schedule (IpPDPerturbationHandler, delta_c never needed: J has full row rank).
 * Returns the delta_w used, or -1 if none up to max_hessian_perturbation gives inertia (n, m, 0). */
static double factor_with_inertia_correction(ipm_t *s, const double *x, const double *zl, const double *zu,
                                             double *dw_last, const double *SigN, const double *SigP, double *Dc,
                                             orc_ipm_stats *out) {
  double dw = 0.0;
  int first_try = 1;
  for (;;) {
    if (Dc)   /* restoration phase: the slacks n, p are primal variables and get delta_w too */
      for (int j = 0; j < s->m; j++) Dc[j] = 1.0 / (SigN[j] + dw) + 1.0 / (SigP[j] + dw);
    if (kkt_factor(s, x, zl, zu, dw, 0.0, Dc)) break;
    if (first_try) {
      out->n_regularized++;
      dw = (*dw_last == 0.0) ? DW_FIRST : fmax(DW_MIN, *dw_last * DW_DEC);
      first_try = 0;
    } else {
      dw = (*dw_last == 0.0) ? dw * DW_INC_FIRST : dw * DW_INC;
    }
    if (dw > DW_MAX) return -1.0;
  }
  if (dw > 0.0) *dw_last = dw;
  return dw;
}

/* primal-dual errors at the current iterate (grad, c, J must be current) */
static void errors(const ipm_t *s, double mu, double *dual_inf, double *cviol, double *compl_) {
  int n = s->n, m = s->m;
  double di = 0;
  for (int i = 0; i < n; i++) {
    double r = s->grad[i];
    for (int j = 0; j < m; j++) r += s->J[(size_t)j * n + i] * s->lam[j];
    if (s->has_l[i]) r -= s->zl[i];
    if (s->has_u[i]) r += s->zu[i];
    if (fabs(r) > di) di = fabs(r);
  }
  *dual_inf = di;
  *cviol = norminf(s->c, m);
  double cp = 0;
  for (int i = 0; i < n; i++) {
    if (s->has_l[i]) { double v = fabs((s->x[i] - s->xl[i]) * s->zl[i] - mu); if (v > cp) cp = v; }
    if (s->has_u[i]) { double v = fabs((s->xu[i] - s->x[i]) * s->zu[i] - mu); if (v > cp) cp = v; }
  }
  *compl_ = cp;
}

static void err_scaling(const ipm_t *s, double *sd, double *sc) {
  double zsum = 0;
  for (int i = 0; i < s->n; i++) {
    if (s->has_l[i]) zsum += fabs(s->zl[i]);
    if (s->has_u[i]) zsum += fabs(s->zu[i]);
  }
  double lsum = norm1(s->lam, s->m);
  int nz = s->nzl + s->nzu;
  *sd = fmax(S_MAX, (lsum + zsum) / (double)(s->m + nz)) / S_MAX;
  *sc = (nz > 0) ? fmax(S_MAX, zsum / (double)nz) / S_MAX : 1.0;
}

static double frac_to_bound_primal(const ipm_t *s, const double *x, const double *dx, double tau) {
  double a = 1.0;
  for (int i = 0; i < s->n; i++) {
    if (s->has_l[i] && dx[i] < 0) { double v = -tau * (x[i] - s->xl[i]) / dx[i]; if (v < a) a = v; }
    if (s->has_u[i] && dx[i] > 0) { double v = tau * (s->xu[i] - x[i]) / dx[i]; if (v < a) a = v; }
  }
  return a;
}
static double frac_to_bound_dual(const ipm_t *s, const double *zl, const double *zu, const double *dzl,
                                 const double *dzu, double tau) {
  double a = 1.0;
  for (int i = 0; i < s->n; i++) {
    if (s->has_l[i] && dzl[i] < 0) { double v = -tau * zl[i] / dzl[i]; if (v < a) a = v; }
    if (s->has_u[i] && dzu[i] < 0) { double v = -tau * zu[i] / dzu[i]; if (v < a) a = v; }
  }
  return a;
}
/* dz from dx (the eliminated rows of the primal-dual system) */
static void bound_mult_step(const ipm_t *s, double mu, const double *x, const double *zl, const double *zu,
                            const double *dx, double *dzl, double *dzu) {
  for (int i = 0; i < s->n; i++) {
    dzl[i] = s->has_l[i] ? mu / (x[i] - s->xl[i]) - zl[i] - zl[i] / (x[i] - s->xl[i]) * dx[i] : 0.0;
    dzu[i] = s->has_u[i] ? mu / (s->xu[i] - x[i]) - zu[i] + zu[i] / (s->xu[i] - x[i]) * dx[i] : 0.0;
  }
}
/* z += alpha_z dz followed by Ipopt's kappa_sigma safeguard (IpoptAlgorithm::correct_bound_multiplier) */
static void bound_mult_update(const ipm_t *s, double mu, const double *x, double *zl, double *zu,
                              const double *dzl, const double *dzu, double alpha_z) {
  for (int i = 0; i < s->n; i++) {
    if (s->has_l[i]) {
      double zz = zl[i] + alpha_z * dzl[i], sl = x[i] - s->xl[i];
      zl[i] = fmax(fmin(zz, KAPPA_SIGMA * mu / sl), mu / (KAPPA_SIGMA * sl));
    }
    if (s->has_u[i]) {
      double zz = zu[i] + alpha_z * dzu[i], su = s->xu[i] - x[i];
      zu[i] = fmax(fmin(zz, KAPPA_SIGMA * mu / su), mu / (KAPPA_SIGMA * su));
    }
  }
}

typedef struct { double theta, phi; } filt_t;

/* the line search's persistent state (IpBacktrackingLineSearch + IpFilterLSAcceptor members) */
typedef struct {
  double theta, phi, gbd;            /* reference_theta_, reference_barr_, reference_gradBarrTDelta_ */
  double theta_min, theta_max;
  filt_t F[FILTER_MAX];
  int nF, nF_max;
  int n_filter_resets, succ_filter_rej, last_rej_filter;
  int reset_trigger;                 /* filter_reset_trigger */
} ls_t;

static int cmp_le(double lhs, double rhs, double basis) { /* Ipopt Compare_le */
  return lhs - rhs <= 10.0 * MACH_EPS * fabs(basis);
}

static int is_ftype(const ls_t *L, double alpha) {
  return (L->gbd < 0.0 && alpha * pow(-L->gbd, S_PHI) > DELTA_LS * pow(L->theta, S_THETA));
}
static int armijo(const ls_t *L, double alpha, double phi_t) {
  return cmp_le(phi_t - L->phi, ETA_PHI * alpha * L->gbd, L->phi);
}
static int filter_ok(const ls_t *L, double theta_t, double phi_t) {
  for (int k = 0; k < L->nF; k++)
    if (!(cmp_le(theta_t, L->F[k].theta, L->F[k].theta) || cmp_le(phi_t, L->F[k].phi, L->F[k].phi)))
      return 0;
  return 1;
}
/* FilterLSAcceptor::IsAcceptableToCurrentIterate */
static int acceptable_to_current(const ls_t *L, double theta_t, double phi_t, int from_resto) {
  if (!from_resto && phi_t > L->phi) {   /* obj_max_inc: the barrier objective must not explode */
    double basval = 1.0;
    if (fabs(L->phi) > 10.0) basval = log10(fabs(L->phi));
    if (log10(phi_t - L->phi) > OBJ_MAX_INC + basval) return 0;
  }
  return cmp_le(theta_t, (1.0 - GAMMA_THETA) * L->theta, L->theta) ||
         cmp_le(phi_t - L->phi, -GAMMA_PHI * L->theta, L->phi);
}
static void filter_clear(ls_t *L) { L->nF = 0; L->succ_filter_rej = 0; L->last_rej_filter = 0; }
/* FilterLSAcceptor::CheckAcceptabilityOfTrialPoint, including the filter reset heuristic */
static int ls_accept(ls_t *L, double alpha, double theta_t, double phi_t) {
  if (!(theta_t == theta_t) || !(phi_t == phi_t) || isinf(phi_t)) return 0;
  if (L->theta_max > 0 && theta_t > L->theta_max) return 0;
  int ok;
  if (alpha > 0 && is_ftype(L, alpha) && L->theta <= L->theta_min) ok = armijo(L, alpha, phi_t);
  else ok = acceptable_to_current(L, theta_t, phi_t, 0);
  if (!ok) { L->last_rej_filter = 0; return 0; }
  if (!filter_ok(L, theta_t, phi_t)) { L->last_rej_filter = 1; return 0; }
  if (L->n_filter_resets < MAX_FILTER_RESETS) {
    if (L->last_rej_filter) {
      if (++L->succ_filter_rej >= L->reset_trigger) { L->n_filter_resets++; filter_clear(L); }
    } else {
      L->succ_filter_rej = 0;
    }
    L->last_rej_filter = 0;
  }
  return 1;
}
static void filter_add(ls_t *L, double theta, double phi) {
  /* drop dominated entries */
  int k = 0;
  for (int j = 0; j < L->nF; j++)
    if (!(L->F[j].theta >= theta && L->F[j].phi >= phi)) L->F[k++] = L->F[j];
  L->nF = k;
  if (L->nF < FILTER_MAX) { L->F[L->nF].theta = theta; L->F[L->nF].phi = phi; L->nF++; }
  if (L->nF > L->nF_max) L->nF_max = L->nF;
}
static void filter_augment(ls_t *L) { filter_add(L, (1.0 - GAMMA_THETA) * L->theta, L->phi - GAMMA_PHI * L->theta); }
/* FilterLSAcceptor::CalculateAlphaMin */
static double alpha_min_of(const ls_t *L) {
  double a = GAMMA_THETA;
  if (L->gbd < 0) {
    a = fmin(GAMMA_THETA, GAMMA_PHI * L->theta / (-L->gbd));
    if (L->theta <= L->theta_min) a = fmin(a, DELTA_LS * pow(L->theta, S_THETA) / pow(-L->gbd, S_PHI));
  }
  return ALPHA_MIN_FRAC * a;
}

/* solve the (factored) KKT system for rhs_x = -(grad phi_mu + J^T lam), rhs_c = -cvec */
static void kkt_solve_dir(ipm_t *s, double mu, const double *cvec, double *dx, double *dlam) {
  int n = s->n, m = s->m;
  double *rhs = s->rhs;
  for (int i = 0; i < n; i++) {
    double r = s->grad[i];
    for (int j = 0; j < m; j++) r += s->J[(size_t)j * n + i] * s->lam[j];
    if (s->has_l[i]) r -= mu / (s->x[i] - s->xl[i]);
    if (s->has_u[i]) r += mu / (s->xu[i] - s->x[i]);
    rhs[i] = -r;
  }
  for (int j = 0; j < m; j++) rhs[n + j] = -cvec[j];
  ldl_solve(s->kkt, rhs);
  memcpy(dx, rhs, sizeof(double) * n);
  memcpy(dlam, rhs + n, sizeof(double) * m);
}

/* ---- feasibility restoration phase (IpRestoMinC_1Nrm::PerformRestoration) ----------------------
 *   min  rho sum(n + p) + eta/2 ||D_R (x - x_R)||^2   s.t.  c(x) + n - p = 0,  xl <= x <= xu,  n, p >= 0
 * with eta = sqrt(mu), D_R = diag(1 / max(1, |x_R|)), solved by the same interior-point iteration (monotone mu,
 * inertia correction, filter line search with second-order correction) from x_R, until the point is acceptable
 * to the ORIGINAL problem's filter and reduces its infeasibility to RESTO_KAPPA (IpRestoFilterConvCheck).
 * The slacks n, p and their multipliers are eliminated from the Newton system, which leaves the original KKT
 * structure with a diagonal -(Sigma_n^-1 + Sigma_p^-1) in the (2,2) block.
 * SIMPLIFICATION: the nested solve has no watchdog of its own.
 * Returns 0 when the original algorithm can go on from the new s->x (multipliers reset as Ipopt does), else the
 * status to stop with.  *iter counts every restoration iteration (Ipopt numbers them with the outer ones). */
static int restoration(ipm_t *s, ls_t *LSo, double mu_o, double tau_o, int *iter, orc_ipm_stats *out) {
  const int n = s->n, m = s->m;
  const double rho = RESTO_RHO;
  double *buf = (double *)calloc((size_t)(9 * n + 20 * m), sizeof(double));
  double *xR = buf, *dr2 = xR + n, *zl = dr2 + n, *zu = zl + n, *dzl = zu + n, *dzu = dzl + n, *xs = dzu + n, *xt = xs + n;
  double *nn = xt + n, *pp = nn + m, *zn = pp + m, *zp = zn + m, *y = zp + m, *dn = y + m, *dp = dn + m, *dy = dp + m;
  double *SigN = dy + m, *SigP = SigN + m, *Dc = SigP + m, *rc = Dc + m, *csoc = rc + m, *rct = csoc + m;
  double *nt = rct + m, *pt = nt + m, *dy_soc = pt + m, *dn_soc = dy_soc + m, *dp_soc = dn_soc + m;
  double *dx_soc = dp_soc + m;
  int status = -1;
  out->n_resto++;

  memcpy(xR, s->x, sizeof(double) * n);
  for (int i = 0; i < n; i++) { double d = 1.0 / fmax(1.0, fabs(xR[i])); dr2[i] = d * d; }
  double mu = fmax(mu_o, norminf(s->c, m));          /* resto_mu = max(curr_mu, ||c||_inf) */
  double tau = fmax(TAU_MIN, 1.0 - mu);
  double tol_r = s->cfg->tol;
  /* RestoIterateInitializer: n, p from the barrier-optimal split of c, z = mu / slack, z_x capped at rho, y = 0 */
  for (int j = 0; j < m; j++) {
    double cj = s->c[j], a = mu / (2.0 * rho) - 0.5 * cj, b = cj * mu / (2.0 * rho);
    nn[j] = a + sqrt(a * a + b);
    pp[j] = cj + nn[j];
    zn[j] = mu / nn[j];
    zp[j] = mu / pp[j];
    y[j] = 0.0;
  }
  for (int i = 0; i < n; i++) { zl[i] = fmin(rho, s->zl[i]); zu[i] = fmin(rho, s->zu[i]); }
  memcpy(xs, s->x, sizeof(double) * n);   /* xs: the restoration iterate; s->x stays x_R until the end */

  /* original-problem quantities at x_R for the convergence test */
  const double theta_R = norm1(s->c, m), infpr_R = norminf(s->c, m);
  const double orig_inf_pr_max = fmax(RESTO_KAPPA * infpr_R, fmin(s->cfg->tol, CONSTR_VIOL_TOL));

  ls_t LS;
  memset(&LS, 0, sizeof(LS));
  LS.theta_max = RESTO_THETA_MAX_FACT * 1.0;   /* max(1, theta_0) with theta_0 = ||c + n - p||_1 = 0 at the start */
  LS.theta_min = 1e-4 * 1.0;
  LS.reset_trigger = s->cfg->filter_reset_trigger;
  double dw_last = 0.0;
  int accept_cnt = 0, first = 1;
  const int nz = s->nzl + s->nzu + 2 * m;
  /* grad, c, J of the original problem at xs live in s->grad (unused here), s->c, s->J */

  for (;;) {
    const double eta = sqrt(mu);
    for (int j = 0; j < m; j++) rc[j] = s->c[j] + nn[j] - pp[j];
    /* ---- optimality errors of the restoration problem */
    double dinf = 0, cviol = norminf(rc, m), cp0 = 0, cpm = 0, zsum = 0, ysum = norm1(y, m);
    for (int i = 0; i < n; i++) {
      double r = eta * dr2[i] * (xs[i] - xR[i]);
      for (int j = 0; j < m; j++) r += s->J[(size_t)j * n + i] * y[j];
      if (s->has_l[i]) { r -= zl[i]; double v = (xs[i] - s->xl[i]) * zl[i]; cp0 = fmax(cp0, fabs(v)); cpm = fmax(cpm, fabs(v - mu)); zsum += fabs(zl[i]); }
      if (s->has_u[i]) { r += zu[i]; double v = (s->xu[i] - xs[i]) * zu[i]; cp0 = fmax(cp0, fabs(v)); cpm = fmax(cpm, fabs(v - mu)); zsum += fabs(zu[i]); }
      dinf = fmax(dinf, fabs(r));
    }
    for (int j = 0; j < m; j++) {
      dinf = fmax(dinf, fmax(fabs(rho + y[j] - zn[j]), fabs(rho - y[j] - zp[j])));
      double v = nn[j] * zn[j], w = pp[j] * zp[j];
      cp0 = fmax(cp0, fmax(fabs(v), fabs(w)));
      cpm = fmax(cpm, fmax(fabs(v - mu), fabs(w - mu)));
      zsum += fabs(zn[j]) + fabs(zp[j]);
    }
    const double sd = fmax(S_MAX, (ysum + zsum) / (double)(m + nz)) / S_MAX, sc = fmax(S_MAX, zsum / (double)nz) / S_MAX;
    const double E0 = fmax(dinf / sd, fmax(cviol, cp0 / sc));

    /* ---- IpRestoConvCheck / IpRestoFilterConvCheck: is xs good enough for the original problem? */
    if (!first) {
      const double theta_t = norm1(s->c, m), infpr_t = norminf(s->c, m);
      int conv = 0;
      if (RESTO_KAPPA * theta_R < theta_t) conv = 0;
      else if (infpr_t > orig_inf_pr_max) conv = 0;
      else {
        const double phi_t = barrier_phi(s, xs, mu_o);
        conv = filter_ok(LSo, theta_t, phi_t) && acceptable_to_current(LSo, theta_t, phi_t, 1);
      }
      if (conv) { status = 0; break; }
      /* is the restoration problem itself solved?  then the original one is locally infeasible (or the filter
       * blocks a feasible point) */
      int solved = (E0 <= tol_r && dinf <= DUAL_INF_TOL && cviol <= CONSTR_VIOL_TOL && cp0 <= COMPL_INF_TOL);
      if (!solved) {
        if (E0 <= ACCEPT_TOL && dinf <= ACCEPT_DUAL_INF_TOL && cviol <= ACCEPT_CONSTR_VIOL_TOL && cp0 <= ACCEPT_COMPL_INF_TOL) {
          if (++accept_cnt >= ACCEPT_ITER) solved = 1;
        } else accept_cnt = 0;
      }
      if (solved) {
        if (infpr_t <= 1e2 * s->cfg->tol) {
          if (tol_r > 1e-1 * s->cfg->tol) { tol_r *= 1e-2; accept_cnt = 0; }   /* tighten once and go on */
          else { status = ORC_RESTORATION_FAILURE; break; }   /* converged to a feasible point the filter rejects */
        } else { status = ORC_LOCAL_INFEASIBILITY; break; }
      }
      if (!(E0 == E0)) { status = ORC_INVALID_NUMBER_DETECTED; break; }
    }
    if (*iter >= s->cfg->max_iter) { status = ORC_MAXITER_EXCEEDED; break; }

    /* ---- monotone barrier update (skipped in the first restoration iteration: first_iter_resto_) */
    if (!first) {
      for (;;) {
        const double Emu = fmax(dinf / sd, fmax(cviol, cpm / sc));
        if (!(Emu <= KAPPA_EPS * mu)) break;
        const double mu_min = fmin(tol_r, COMPL_INF_TOL) / (KAPPA_EPS + 1.0);
        const double new_mu = fmax(mu_min, fmin(KAPPA_MU * mu, pow(mu, THETA_MU)));
        if (new_mu == mu) break;
        mu = new_mu;
        tau = fmax(TAU_MIN, 1.0 - mu);
        filter_clear(&LS);
        /* complementarity and (through eta) the dual infeasibility depend on mu */
        cpm = 0;
        for (int i = 0; i < n; i++) {
          if (s->has_l[i]) cpm = fmax(cpm, fabs((xs[i] - s->xl[i]) * zl[i] - mu));
          if (s->has_u[i]) cpm = fmax(cpm, fabs((s->xu[i] - xs[i]) * zu[i] - mu));
        }
        for (int j = 0; j < m; j++) cpm = fmax(cpm, fmax(fabs(nn[j] * zn[j] - mu), fabs(pp[j] * zp[j] - mu)));
        const double eta2 = sqrt(mu);
        dinf = 0;
        for (int i = 0; i < n; i++) {
          double r = eta2 * dr2[i] * (xs[i] - xR[i]);
          for (int j = 0; j < m; j++) r += s->J[(size_t)j * n + i] * y[j];
          if (s->has_l[i]) r -= zl[i];
          if (s->has_u[i]) r += zu[i];
          dinf = fmax(dinf, fabs(r));
        }
        for (int j = 0; j < m; j++) dinf = fmax(dinf, fmax(fabs(rho + y[j] - zn[j]), fabs(rho - y[j] - zp[j])));
      }
    }
    first = 0;
    const double etam = sqrt(mu);

    /* ---- search direction: W = sum_j y_j Hess c_j + eta D_R^2 (no objective term) */
    {
      double *ls = s->ct;
      for (int j = 0; j < m; j++) ls[j] = y[j] * s->cs[j];
      s->nlp->hess(s->nlp->user, xs, 0.0, ls, s->H);
      for (int i = 0; i < n; i++) s->H[(size_t)i * n + i] += etam * dr2[i];
    }
    for (int j = 0; j < m; j++) { SigN[j] = zn[j] / nn[j]; SigP[j] = zp[j] / pp[j]; }
    const double dw = factor_with_inertia_correction(s, xs, zl, zu, &dw_last, SigN, SigP, Dc, out);
    if (dw < 0) { status = ORC_ERROR_IN_STEP_COMPUTATION; break; }
    double *rhs = s->rhs;
#define RESTO_SOLVE(CV, DX, DY, DN, DP)                                                              \
    do {                                                                                             \
      for (int i = 0; i < n; i++) {                                                                  \
        double r = etam * dr2[i] * (xs[i] - xR[i]);                                                  \
        for (int j = 0; j < m; j++) r += s->J[(size_t)j * n + i] * y[j];                             \
        if (s->has_l[i]) r -= mu / (xs[i] - s->xl[i]);                                               \
        if (s->has_u[i]) r += mu / (s->xu[i] - xs[i]);                                               \
        rhs[i] = -r;                                                                                 \
      }                                                                                              \
      for (int j = 0; j < m; j++) {                                                                  \
        const double rn = rho + y[j] - mu / nn[j], rp = rho - y[j] - mu / pp[j];                     \
        rhs[n + j] = -(CV)[j] + rn / (SigN[j] + dw) - rp / (SigP[j] + dw);                           \
      }                                                                                              \
      ldl_solve(s->kkt, rhs);                                                                        \
      memcpy((DX), rhs, sizeof(double) * n);                                                         \
      memcpy((DY), rhs + n, sizeof(double) * m);                                                     \
      for (int j = 0; j < m; j++) {                                                                  \
        const double rn = rho + y[j] - mu / nn[j], rp = rho - y[j] - mu / pp[j];                     \
        (DN)[j] = (-rn - (DY)[j]) / (SigN[j] + dw);                                                  \
        (DP)[j] = (-rp + (DY)[j]) / (SigP[j] + dw);                                                  \
      }                                                                                              \
    } while (0)
    RESTO_SOLVE(rc, s->dx, dy, dn, dp);

    /* ---- fraction to the boundary over x, n, p */
#define RESTO_ALPHA_MAX(DX, DN, DP, A)                                                               \
    do {                                                                                             \
      (A) = frac_to_bound_primal(s, xs, (DX), tau);                                                  \
      for (int j = 0; j < m; j++) {                                                                  \
        if ((DN)[j] < 0) { double v = -tau * nn[j] / (DN)[j]; if (v < (A)) (A) = v; }                \
        if ((DP)[j] < 0) { double v = -tau * pp[j] / (DP)[j]; if (v < (A)) (A) = v; }                \
      }                                                                                              \
    } while (0)
    double alpha_max;
    RESTO_ALPHA_MAX(s->dx, dn, dp, alpha_max);

    /* ---- filter line search on (theta_R, phi_R) */
    double slog = log_slacks(s, xs), fR = 0;
    for (int j = 0; j < m; j++) { slog += log(nn[j]) + log(pp[j]); fR += rho * (nn[j] + pp[j]); }
    for (int i = 0; i < n; i++) fR += 0.5 * etam * dr2[i] * (xs[i] - xR[i]) * (xs[i] - xR[i]);
    LS.theta = norm1(rc, m);
    LS.phi = fR - mu * slog;
    LS.gbd = 0;
    for (int i = 0; i < n; i++) {
      double g = etam * dr2[i] * (xs[i] - xR[i]);
      if (s->has_l[i]) g -= mu / (xs[i] - s->xl[i]);
      if (s->has_u[i]) g += mu / (s->xu[i] - xs[i]);
      LS.gbd += g * s->dx[i];
    }
    for (int j = 0; j < m; j++) LS.gbd += (rho - mu / nn[j]) * dn[j] + (rho - mu / pp[j]) * dp[j];
    const double alpha_min = alpha_min_of(&LS);
#define RESTO_TRIAL(A, DX, DN, DP, TH, PH)                                                           \
    do {                                                                                             \
      for (int i = 0; i < n; i++) xt[i] = xs[i] + (A) * (DX)[i];                                     \
      nlp_c(s, xt, rct);                                                                             \
      double sl_ = log_slacks(s, xt), f_ = 0;                                                        \
      for (int j = 0; j < m; j++) {                                                                  \
        nt[j] = nn[j] + (A) * (DN)[j]; pt[j] = pp[j] + (A) * (DP)[j];                                \
        rct[j] += nt[j] - pt[j];                                                                     \
        sl_ += log(nt[j]) + log(pt[j]); f_ += rho * (nt[j] + pt[j]);                                 \
      }                                                                                              \
      for (int i = 0; i < n; i++) f_ += 0.5 * etam * dr2[i] * (xt[i] - xR[i]) * (xt[i] - xR[i]);     \
      (TH) = norm1(rct, m); (PH) = f_ - mu * sl_;                                                    \
    } while (0)
    double alpha = alpha_max, alpha_test = alpha_max;
    double *dx_use = s->dx, *dy_use = dy, *dn_use = dn, *dp_use = dp;
    int accepted = 0, ntrial = 0;
    double phi_acc = 0;
    while (!accepted) {
      double theta_t, phi_t;
      RESTO_TRIAL(alpha, s->dx, dn, dp, theta_t, phi_t);
      alpha_test = alpha;
      if (ls_accept(&LS, alpha_test, theta_t, phi_t)) { accepted = 1; phi_acc = phi_t; break; }
      if (ntrial == 0 && s->cfg->max_soc > 0 && theta_t >= LS.theta) {
        int cnt = 0;
        double theta_soc_old = 0, theta_trial = theta_t, alpha_soc = alpha;
        memcpy(csoc, rc, sizeof(double) * m);
        while (cnt < s->cfg->max_soc && !accepted && (cnt == 0 || theta_trial <= KAPPA_SOC * theta_soc_old)) {
          theta_soc_old = theta_trial;
          for (int j = 0; j < m; j++) csoc[j] = alpha_soc * csoc[j] + rct[j];
          RESTO_SOLVE(csoc, dx_soc, dy_soc, dn_soc, dp_soc);
          RESTO_ALPHA_MAX(dx_soc, dn_soc, dp_soc, alpha_soc);
          double phi_soc;
          RESTO_TRIAL(alpha_soc, dx_soc, dn_soc, dp_soc, theta_trial, phi_soc);
          if (ls_accept(&LS, alpha_test, theta_trial, phi_soc)) {
            accepted = 1; alpha = alpha_soc; phi_acc = phi_soc;
            dx_use = dx_soc; dy_use = dy_soc; dn_use = dn_soc; dp_use = dp_soc;
            out->n_soc++;
          } else cnt++;
        }
        if (accepted) break;
      }
      alpha *= ALPHA_RED;
      ntrial++;
      out->n_backtrack++;
      if (!(alpha > alpha_min)) break;   /* theta_R = 0 gives alpha_min = 0: `alpha < alpha_min` would never end */
    }
    if (!accepted) { status = ORC_RESTORATION_FAILURE; break; }   /* no restoration inside the restoration */
    if (!is_ftype(&LS, alpha_test) || !armijo(&LS, alpha_test, phi_acc)) filter_augment(&LS);

    /* ---- accept: primal step alpha, y with alpha, bound multipliers with their own fraction to the boundary */
    {
      double az = 1.0;
      bound_mult_step(s, mu, xs, zl, zu, dx_use, dzl, dzu);
      az = frac_to_bound_dual(s, zl, zu, dzl, dzu, tau);
      for (int j = 0; j < m; j++) {
        const double dzn = mu / nn[j] - zn[j] - SigN[j] * dn_use[j], dzp = mu / pp[j] - zp[j] - SigP[j] * dp_use[j];
        if (dzn < 0) { double v = -tau * zn[j] / dzn; if (v < az) az = v; }
        if (dzp < 0) { double v = -tau * zp[j] / dzp; if (v < az) az = v; }
      }
      for (int j = 0; j < m; j++) {
        const double dzn = mu / nn[j] - zn[j] - SigN[j] * dn_use[j], dzp = mu / pp[j] - zp[j] - SigP[j] * dp_use[j];
        const double nnew = nn[j] + alpha * dn_use[j], pnew = pp[j] + alpha * dp_use[j];
        zn[j] = fmax(fmin(zn[j] + az * dzn, KAPPA_SIGMA * mu / nnew), mu / (KAPPA_SIGMA * nnew));
        zp[j] = fmax(fmin(zp[j] + az * dzp, KAPPA_SIGMA * mu / pnew), mu / (KAPPA_SIGMA * pnew));
        nn[j] = nnew; pp[j] = pnew;
        y[j] += alpha * dy_use[j];
      }
      for (int i = 0; i < n; i++) xs[i] += alpha * dx_use[i];
      bound_mult_update(s, mu, xs, zl, zu, dzl, dzu, az);
    }
    nlp_c(s, xs, s->c);
    nlp_jac(s, xs, s->J);
    (*iter)++;
    out->n_resto_iter++;
    if (s->trace)
      fprintf(stderr, "  %4dr th_o %.3e  thR %.3e  lg(mu) %5.1f  dw %.1e  a %.2e  nt %d\n", *iter, norm1(s->c, m), LS.theta,
              log10(mu), dw, alpha, ntrial);
  }
#undef RESTO_SOLVE
#undef RESTO_ALPHA_MAX
#undef RESTO_TRIAL

  if (status == 0) {
    /* back to the original problem (PerformRestoration): the bound multipliers take one primal-dual "step" from
     * x_R to the new point, are reset to 1 if that leaves them above bound_mult_reset_threshold; the constraint
     * multipliers are reset to 0 (constr_mult_reset_threshold = 0) */
    double zmax = 0;
    for (int i = 0; i < n; i++) {
      dzl[i] = 0; dzu[i] = 0;
      if (s->has_l[i]) { const double sR = xR[i] - s->xl[i], sN = xs[i] - s->xl[i]; dzl[i] = (s->zl[i] * (sR - sN) + mu_o) / sR - s->zl[i]; }
      if (s->has_u[i]) { const double sR = s->xu[i] - xR[i], sN = s->xu[i] - xs[i]; dzu[i] = (s->zu[i] * (sR - sN) + mu_o) / sR - s->zu[i]; }
    }
    const double az = frac_to_bound_dual(s, s->zl, s->zu, dzl, dzu, tau_o);
    for (int i = 0; i < n; i++) {
      if (s->has_l[i]) { s->zl[i] += az * dzl[i]; zmax = fmax(zmax, fabs(s->zl[i])); }
      if (s->has_u[i]) { s->zu[i] += az * dzu[i]; zmax = fmax(zmax, fabs(s->zu[i])); }
    }
    if (zmax > BOUND_MULT_RESET_THRESHOLD)
      for (int i = 0; i < n; i++) { if (s->has_l[i]) s->zl[i] = 1.0; if (s->has_u[i]) s->zu[i] = 1.0; }
    for (int j = 0; j < m; j++) s->lam[j] = 0.0;
    memcpy(s->x, xs, sizeof(double) * n);
    (*iter)++;   /* the call itself is an iteration of the outer algorithm */
  } else {
    /* failure: the outer algorithm stops with its own iterate x_R; restore its function values */
    nlp_c(s, s->x, s->c);
    nlp_jac(s, s->x, s->J);
  }
  free(buf);
  return status;
}

/* Generic core: minimise f(x) s.t. g(x) = gl (= gu), xl <= x <= xu from the start point xi.
 * Bounds beyond +-1e19 are "no bound".  Outputs x (clipped to the original bounds), lambda, zl, zu
 * (unscaled) and the run statistics. */
int orc_ipm_solve(const orc_nlp *nlp, const orc_config *cfg, const double *xi_in, const double *xl_in,
                  const double *xu_in, const double *gl_in, const double *gu_in, double *x_out,
                  double *lam_out, double *zl_out, double *zu_out, orc_ipm_stats *out) {
  ipm_t S;
  ipm_t *s = &S;
  memset(s, 0, sizeof(S));
  int n = nlp->n, m = nlp->m;
  s->cfg = cfg; s->nlp = nlp; s->n = n; s->m = m;
  s->trace = getenv("ORC_TRACE") != NULL;
  size_t nd = sizeof(double);
  double *pool = (double *)calloc((size_t)(50 * n + 20 * m + (size_t)m * n + (size_t)n * n + (n + m)), nd);
  double *q = pool;
#define TAKE(cnt) (q += (cnt), q - (cnt))
  s->cs = TAKE(m); s->xl = TAKE(n); s->xu = TAKE(n); s->xl0 = TAKE(n); s->xu0 = TAKE(n); s->gl = TAKE(m);
  s->x = TAKE(n); s->lam = TAKE(m); s->zl = TAKE(n); s->zu = TAKE(n);
  s->grad = TAKE(n); s->c = TAKE(m); s->J = TAKE((size_t)m * n); s->H = TAKE((size_t)n * n);
  s->rhs = TAKE(n + m); s->dx = TAKE(n); s->dlam = TAKE(m); s->dzl = TAKE(n); s->dzu = TAKE(n);
  s->xt = TAKE(n); s->ct = TAKE(m); s->csoc = TAKE(m);
  double *xi = TAKE(n), *dx_soc = TAKE(n), *dlam_soc = TAKE(m);
  /* watchdog backup: iterate and search direction at the point where it was started */
  double *wd_x = TAKE(n), *wd_lam = TAKE(m), *wd_zl = TAKE(n), *wd_zu = TAKE(n);
  double *wd_dx = TAKE(n), *wd_dlam = TAKE(m), *wd_dzl = TAKE(n), *wd_dzu = TAKE(n);
#undef TAKE
  s->has_l = (int *)calloc(2 * (size_t)n, sizeof(int));
  s->has_u = s->has_l + n;
  s->kkt = ldl_new(n + m);

  memset(out, 0, sizeof(*out));
  int status = ORC_NOT_DEFINED;

  memcpy(xi, xi_in, nd * n);
  memcpy(s->xl0, xl_in, nd * n);
  memcpy(s->xu0, xu_in, nd * n);
  memcpy(s->gl, gl_in, nd * m);
  for (int j = 0; j < m; j++)
    if (gl_in[j] != gu_in[j]) { ldl_free(s->kkt); free(s->has_l); free(pool); return -1; }  /* equalities only */

  /* gradient-based NLP scaling at the user start point (nlp_scaling_max_gradient 100) */
  s->sf = 1.0;
  for (int j = 0; j < m; j++) s->cs[j] = 1.0;
  if (cfg->obj_scaling) {
    nlp->grad(nlp->user, xi, s->grad);
    double gmax = norminf(s->grad, n);
    if (gmax > 100.0) s->sf = fmax(100.0 / gmax, 1e-8);
    nlp->jac(nlp->user, xi, s->J);
    for (int j = 0; j < m; j++) {
      double rmax = norminf(s->J + (size_t)j * n, n);
      if (rmax > 100.0) s->cs[j] = fmax(100.0 / rmax, 1e-8);
    }
  }

  /* bounds: +-1e19 is "no bound" (nlp_{lower,upper}_bound_inf); relax by bound_relax_factor */
  for (int i = 0; i < n; i++) {
    s->has_l[i] = s->xl0[i] > -NLP_INF;
    s->has_u[i] = s->xu0[i] < NLP_INF;
    s->xl[i] = s->has_l[i] ? s->xl0[i] - fmin(CONSTR_VIOL_TOL, BOUND_RELAX * fmax(1.0, fabs(s->xl0[i]))) : -HUGE_VAL;
    s->xu[i] = s->has_u[i] ? s->xu0[i] + fmin(CONSTR_VIOL_TOL, BOUND_RELAX * fmax(1.0, fabs(s->xu0[i]))) : HUGE_VAL;
    s->nzl += s->has_l[i];
    s->nzu += s->has_u[i];
  }

  /* start point pushed into the interior (bound_push / bound_frac) */
  for (int i = 0; i < n; i++) {
    double v = xi[i];
    if (s->has_l[i] && s->has_u[i]) {
      double span = s->xu[i] - s->xl[i];
      double pl = fmin(KAPPA_1 * fmax(1.0, fabs(s->xl[i])), KAPPA_2 * span);
      double pu = fmin(KAPPA_1 * fmax(1.0, fabs(s->xu[i])), KAPPA_2 * span);
      if (v < s->xl[i] + pl) v = s->xl[i] + pl;
      if (v > s->xu[i] - pu) v = s->xu[i] - pu;
    } else if (s->has_l[i]) {
      double pl = KAPPA_1 * fmax(1.0, fabs(s->xl[i]));
      if (v < s->xl[i] + pl) v = s->xl[i] + pl;
    } else if (s->has_u[i]) {
      double pu = KAPPA_1 * fmax(1.0, fabs(s->xu[i]));
      if (v > s->xu[i] - pu) v = s->xu[i] - pu;
    }
    s->x[i] = v;
    s->zl[i] = s->has_l[i] ? 1.0 : 0.0;   /* bound_mult_init_val */
    s->zu[i] = s->has_u[i] ? 1.0 : 0.0;
  }

  double mu = cfg->mu_init;
  double tau = fmax(TAU_MIN, 1.0 - mu);

  nlp_grad(s, s->x, s->grad);
  nlp_c(s, s->x, s->c);
  nlp_jac(s, s->x, s->J);

  /* least-squares multiplier estimate: [[I, J^T],[J, 0]] [r; lam] = [-(grad - zl + zu); 0] */
  {
    int d = n + m;
    double *A = s->kkt->A;
    memset(A, 0, sizeof(double) * (size_t)d * d);
    for (int i = 0; i < n; i++) A[(size_t)i * d + i] = 1.0;
    for (int j = 0; j < m; j++)
      for (int i = 0; i < n; i++) {
        double v = s->J[(size_t)j * n + i];
        A[(size_t)(n + j) * d + i] = v;
        A[(size_t)i * d + (n + j)] = v;
      }
    ldl_factor(s->kkt);
    for (int i = 0; i < n; i++) s->rhs[i] = -(s->grad[i] - s->zl[i] + s->zu[i]);
    for (int j = 0; j < m; j++) s->rhs[n + j] = 0;
    ldl_solve(s->kkt, s->rhs);
    double lmax = norminf(s->rhs + n, m);
    int bad = !(lmax <= CONSTR_MULT_INIT_MAX);
    for (int j = 0; j < m; j++) s->lam[j] = bad ? 0.0 : s->rhs[n + j];
  }

  ls_t LS;
  memset(&LS, 0, sizeof(LS));
  LS.reset_trigger = cfg->filter_reset_trigger;
  {
    double th0 = norm1(s->c, m);
    LS.theta_max = 1e4 * fmax(1.0, th0);
    LS.theta_min = 1e-4 * fmax(1.0, th0);
  }
  /* BacktrackingLineSearch members */
  int in_watchdog = 0, wd_short = 0, wd_trial = 0, tiny_last = 0, tiny_flag = 0;
  double wd_alpha_test = 0, wd_theta = 0, wd_phi = 0, wd_gbd = 0;

  double dw_last = 0.0;
  int iter = 0, accept_cnt = 0;
  double E0 = 0;

  for (;;) {
    /* ---- convergence check */
    double dinf, cviol, compl0, compl_mu, sd, sc;
    int acceptable_now = 0;
    errors(s, 0.0, &dinf, &cviol, &compl0);
    err_scaling(s, &sd, &sc);
    E0 = fmax(dinf / sd, fmax(cviol, compl0 / sc));
    {
      /* unscaled quantities (constraint scaling is 1 except in pathological windows) */
      double dinf_u = dinf / s->sf, compl_u = compl0 / s->sf, cviol_u = 0;
      for (int j = 0; j < m; j++) { double v = fabs(s->c[j] / s->cs[j]); if (v > cviol_u) cviol_u = v; }
      if (E0 <= cfg->tol && dinf_u <= DUAL_INF_TOL && cviol_u <= CONSTR_VIOL_TOL && compl_u <= COMPL_INF_TOL) {
        status = ORC_SUCCESS;
        break;
      }
      if (E0 <= ACCEPT_TOL && dinf_u <= ACCEPT_DUAL_INF_TOL && cviol_u <= ACCEPT_CONSTR_VIOL_TOL &&
          compl_u <= ACCEPT_COMPL_INF_TOL) {
        acceptable_now = 1;
        if (++accept_cnt >= ACCEPT_ITER) { status = ORC_STOP_AT_ACCEPTABLE_POINT; break; }
      } else {
        accept_cnt = 0;
      }
    }
    if (!(E0 == E0)) { status = ORC_INVALID_NUMBER_DETECTED; break; }
    if (iter >= cfg->max_iter) { status = ORC_MAXITER_EXCEEDED; break; }

    /* ---- monotone barrier update (mu_allow_fast_monotone_decrease); a repeated tiny step forces a decrease */
    {
      int stop = 0;
      for (;;) {
        errors(s, mu, &dinf, &cviol, &compl_mu);
        double Emu = fmax(dinf / sd, fmax(cviol, compl_mu / sc));
        if (!(Emu <= KAPPA_EPS * mu) && !tiny_flag) break;
        double mu_min = fmin(cfg->tol, COMPL_INF_TOL) / (KAPPA_EPS + 1.0);
        double new_mu = fmax(mu_min, fmin(KAPPA_MU * mu, pow(mu, THETA_MU)));
        if (new_mu == mu) { if (tiny_flag) stop = 1; break; }
        mu = new_mu;
        tau = fmax(TAU_MIN, 1.0 - mu);
        tiny_flag = 0;
        /* BacktrackingLineSearch::Reset */
        filter_clear(&LS);
        in_watchdog = 0; wd_short = 0; tiny_last = 0;
      }
      tiny_flag = 0;
      if (stop) { status = ORC_STOP_AT_TINY_STEP; break; }
    }

    /* ---- search direction with inertia correction */
    { /* Hessian of sf*f + sum_j lam_j * cs_j * g_j (scaled multipliers on scaled constraints) */
      double *ls = s->ct;
      for (int j = 0; j < m; j++) ls[j] = s->lam[j] * s->cs[j];
      nlp->hess(nlp->user, s->x, s->sf, ls, s->H);
    }
    double dw = factor_with_inertia_correction(s, s->x, s->zl, s->zu, &dw_last, NULL, NULL, NULL, out);
    if (dw < 0) { status = ORC_ERROR_IN_STEP_COMPUTATION; break; }
    kkt_solve_dir(s, mu, s->c, s->dx, s->dlam);
    bound_mult_step(s, mu, s->x, s->zl, s->zu, s->dx, s->dzl, s->dzu);

    /* ---- line search (BacktrackingLineSearch::FindAcceptableTrialPoint) */
    if (!in_watchdog) {   /* InitThisLineSearch: reference values of the current iterate */
      LS.theta = norm1(s->c, m);
      LS.phi = barrier_phi(s, s->x, mu);
      LS.gbd = 0;
      for (int i = 0; i < n; i++) {
        double gphi = s->grad[i];
        if (s->has_l[i]) gphi -= mu / (s->x[i] - s->xl[i]);
        if (s->has_u[i]) gphi += mu / (s->xu[i] - s->x[i]);
        LS.gbd += gphi * s->dx[i];
      }
    }
    /* DetectTinyStep */
    int tiny = 0;
    {
      double mx = 0;
      for (int i = 0; i < n; i++) { double v = fabs(s->dx[i]) / (1.0 + fabs(s->x[i])); if (v > mx) mx = v; }
      tiny = cfg->tiny_step_tol > 0 && (mx <= cfg->tiny_step_tol) && (norm1(s->c, m) <= 1e-4);
    }
#define STOP_WATCHDOG()                                                                              \
    do {                                                                                             \
      memcpy(s->x, wd_x, nd * n); memcpy(s->lam, wd_lam, nd * m); memcpy(s->zl, wd_zl, nd * n); memcpy(s->zu, wd_zu, nd * n); \
      memcpy(s->dx, wd_dx, nd * n); memcpy(s->dlam, wd_dlam, nd * m); memcpy(s->dzl, wd_dzl, nd * n); memcpy(s->dzu, wd_dzu, nd * n); \
      nlp_grad(s, s->x, s->grad); nlp_c(s, s->x, s->c); nlp_jac(s, s->x, s->J);                       \
      LS.theta = wd_theta; LS.phi = wd_phi; LS.gbd = wd_gbd;                                         \
      in_watchdog = 0; wd_short = 0;                                                                 \
    } while (0)
    if (in_watchdog && tiny) { STOP_WATCHDOG(); tiny = 0; }
    if (cfg->watchdog_trigger > 0 && !in_watchdog && !tiny && wd_short >= cfg->watchdog_trigger) {   /* StartWatchDog */
      in_watchdog = 1; wd_trial = 0;
      memcpy(wd_x, s->x, nd * n); memcpy(wd_lam, s->lam, nd * m); memcpy(wd_zl, s->zl, nd * n); memcpy(wd_zu, s->zu, nd * n);
      memcpy(wd_dx, s->dx, nd * n); memcpy(wd_dlam, s->dlam, nd * m); memcpy(wd_dzl, s->dzl, nd * n); memcpy(wd_dzu, s->dzu, nd * n);
      wd_alpha_test = frac_to_bound_primal(s, s->x, s->dx, tau);
      wd_theta = LS.theta; wd_phi = LS.phi; wd_gbd = LS.gbd;
      out->n_watchdog++;
    }

    int accepted = 0, n_steps = 0, via_resto = 0, update_filter = 0;
    double alpha = 0, alpha_test = 0, phi_acc = 0;
    const double *dx_use = s->dx, *dlam_use = s->dlam;
    char tag = ' ';

    if (tiny) {
      /* a numerically insignificant step is taken without any test; twice in a row (with a small dual step)
       * it forces the barrier parameter down, and ends the run if that is no longer possible */
      alpha = frac_to_bound_primal(s, s->x, s->dx, tau);
      for (int i = 0; i < n; i++) s->xt[i] = s->x[i] + alpha * s->dx[i];
      if (tiny_last && norminf(s->dlam, m) < TINY_STEP_Y_TOL) tiny_flag = 1;
      tiny_last = 1;
      accepted = 1; tag = 't';
      out->n_tiny++;
    } else {
      tiny_last = 0;
      {
        int skip_first = 0;
        for (;;) {
          /* DoBacktrackingLineSearch */
          const double alpha_max = frac_to_bound_primal(s, s->x, s->dx, tau);
          const double alpha_min = in_watchdog ? alpha_max : alpha_min_of(&LS);
          alpha = alpha_max;
          alpha_test = in_watchdog ? wd_alpha_test : alpha;
          if (skip_first) alpha *= ALPHA_RED;
          n_steps = 0; accepted = 0;
          dx_use = s->dx; dlam_use = s->dlam;
          while (alpha > alpha_min || n_steps == 0) {
            for (int i = 0; i < n; i++) s->xt[i] = s->x[i] + alpha * s->dx[i];
            nlp_c(s, s->xt, s->ct);
            double theta_t = norm1(s->ct, m);
            double phi_t = barrier_phi(s, s->xt, mu);
            if (!in_watchdog) alpha_test = alpha;
            if (ls_accept(&LS, alpha_test, theta_t, phi_t)) { accepted = 1; phi_acc = phi_t; break; }
            if (in_watchdog) break;
            /* second-order correction, only for the first trial step and if theta did not decrease */
            if (alpha == alpha_max && cfg->max_soc > 0 && LS.theta <= theta_t) {
              int cnt = 0;
              double theta_soc_old = 0, theta_trial = theta_t, alpha_soc = alpha;
              memcpy(s->csoc, s->c, nd * m);
              while (cnt < cfg->max_soc && !accepted && (cnt == 0 || theta_trial <= KAPPA_SOC * theta_soc_old)) {
                theta_soc_old = theta_trial;
                for (int j = 0; j < m; j++) s->csoc[j] = alpha_soc * s->csoc[j] + s->ct[j];
                kkt_solve_dir(s, mu, s->csoc, dx_soc, dlam_soc);
                alpha_soc = frac_to_bound_primal(s, s->x, dx_soc, tau);
                for (int i = 0; i < n; i++) s->xt[i] = s->x[i] + alpha_soc * dx_soc[i];
                nlp_c(s, s->xt, s->ct);
                theta_trial = norm1(s->ct, m);
                double phi_soc = barrier_phi(s, s->xt, mu);
                if (ls_accept(&LS, alpha_test, theta_trial, phi_soc)) {
                  accepted = 1; phi_acc = phi_soc;
                  alpha = alpha_soc;
                  dx_use = dx_soc;
                  dlam_use = dlam_soc;
                  out->n_soc++;
                } else {
                  cnt++;
                }
              }
              if (accepted) break;
            }
            alpha *= ALPHA_RED;
            n_steps++;
            out->n_backtrack++;
          }
          if (!in_watchdog) { update_filter = accepted; tag = accepted ? 'f' : ' '; break; }
          if (accepted) { in_watchdog = 0; update_filter = 1; tag = 'W'; break; }
          if (++wd_trial > WATCHDOG_TRIAL_MAX) {
            STOP_WATCHDOG();          /* back to the stored iterate and direction, backtrack from alpha_max / 2 */
            skip_first = 1;
            continue;
          }
          accepted = 1; tag = 'w';    /* take the full step untested */
          break;
        }
      }
    }

    if (!accepted) {
      /* the step size fell below alpha_min: restoration phase.  SIMPLIFICATION: Ipopt first tries a "soft
       * restoration" step (the same direction with one step length for primal and dual variables, accepted if the
       * primal-dual system error drops by 1e-4); that is not restated -- on every problem of the horizon grid that
       * reaches this point the soft step was rejected when it was still tried here. */
      filter_augment(&LS);    /* PrepareRestoPhaseStart */
      {
        if (acceptable_now) { status = ORC_STOP_AT_ACCEPTABLE_POINT; break; }   /* "restoration phase called at acceptable point" */
        if (norm1(s->c, m) <= 1e-2 * cfg->tol) { status = ORC_RESTORATION_FAILURE; break; }   /* "... at almost feasible point" (RestoreAcceptablePoint not restated) */
        int rs = restoration(s, &LS, mu, tau, &iter, out);
        if (rs != 0) { status = rs; break; }
        via_resto = 1; accepted = 1; tag = 'R';
        wd_short = 0;
      }
    }

    /* ---- accept the trial point */
    if (via_resto) {
      /* restoration() has set x, lam and z */
    } else {
      if (update_filter && (!is_ftype(&LS, alpha_test) || !armijo(&LS, alpha_test, phi_acc))) filter_augment(&LS);
      double alpha_z;
      if (dx_use != s->dx) bound_mult_step(s, mu, s->x, s->zl, s->zu, dx_use, s->dzl, s->dzu);   /* dz follows the accepted (SOC) dx */
      alpha_z = frac_to_bound_dual(s, s->zl, s->zu, s->dzl, s->dzu, tau);
      for (int i = 0; i < n; i++) s->x[i] = s->xt[i];
      for (int j = 0; j < m; j++) s->lam[j] += alpha * dlam_use[j];
      bound_mult_update(s, mu, s->x, s->zl, s->zu, s->dzl, s->dzu, alpha_z);
      if (tag != 't') { if (n_steps == 0) wd_short = 0; else wd_short++; }
      iter++;
    }
    nlp_grad(s, s->x, s->grad);
    nlp_c(s, s->x, s->c);
    nlp_jac(s, s->x, s->J);
    if (s->trace)
      fprintf(stderr, "%4d  f %.8e  th %.3e  lg(mu) %5.1f  dw %.1e  a %.2e %c  ls %d  wd %d nF %d\n", iter, nlp_f(s, s->x) / s->sf,
              norm1(s->c, m), log10(mu), dw, alpha, tag, n_steps, wd_short, LS.nF);
  }
#undef STOP_WATCHDOG

  /* ---- finalize: honor_original_bounds, unscale */
  for (int i = 0; i < n; i++) {
    double v = s->x[i];
    if (s->has_l[i] && v < s->xl0[i]) v = s->xl0[i];
    if (s->has_u[i] && v > s->xu0[i]) v = s->xu0[i];
    x_out[i] = v;
    if (zl_out) zl_out[i] = s->zl[i] / s->sf;
    if (zu_out) zu_out[i] = s->zu[i] / s->sf;
  }
  if (lam_out)
    for (int j = 0; j < m; j++) lam_out[j] = s->lam[j] * s->cs[j] / s->sf;
  out->status = status;
  out->n_filter_reset = LS.n_filter_resets;
  out->n_filter_max = LS.nF_max;
  out->iters = iter;
  out->kkt_error = E0;
  out->obj = nlp->f(nlp->user, x_out);

  ldl_free(s->kkt);
  free(s->has_l);
  free(pool);
  return 0;
}

/* The analytic restatement of FG_eval as an orc_nlp */
typedef struct { const orc_config *cfg; const orc_problem *prob; } mpc_user;
static double mpc_f(void *u, const double *x) { mpc_user *q = (mpc_user *)u; return orc_eval_f(q->cfg, q->prob, x); }
static void mpc_grad(void *u, const double *x, double *g) { mpc_user *q = (mpc_user *)u; orc_eval_grad(q->cfg, q->prob, x, g); }
static void mpc_g(void *u, const double *x, double *c) { mpc_user *q = (mpc_user *)u; orc_eval_g(q->cfg, q->prob, x, c); }
static void mpc_jac(void *u, const double *x, double *J) { mpc_user *q = (mpc_user *)u; orc_eval_jac(q->cfg, q->prob, x, J); }
static void mpc_hess(void *u, const double *x, double sigma, const double *lam, double *H) {
  mpc_user *q = (mpc_user *)u;
  orc_eval_hess(q->cfg, q->prob, x, sigma, lam, H);
}

int orc_solve(const orc_config *cfg, const orc_problem *prob, orc_result *out) {
  int N = cfg->N, n = 8 * N - 2, m = 6 * N;
  mpc_user user = {cfg, prob};
  orc_nlp nlp = {n, m, &user, mpc_f, mpc_grad, mpc_g, mpc_jac, mpc_hess};
  double *buf = (double *)calloc((size_t)(3 * n + 2 * m), sizeof(double));
  double *xl = buf, *xu = xl + n, *xi = xu + n, *gl = xi + n, *gu = gl + m;
  orc_bounds(cfg, prob, xl, xu, gl, gu, xi);
  memset(out, 0, sizeof(*out));
  orc_ipm_stats st;
  int rc = orc_ipm_solve(&nlp, cfg, xi, xl, xu, gl, gu, out->z, out->lambda, out->zl, out->zu, &st);
  free(buf);
  if (rc) return rc;
  out->status = st.status; out->iters = st.iters; out->n_regularized = st.n_regularized;
  out->n_soc = st.n_soc; out->n_backtrack = st.n_backtrack; out->obj = st.obj; out->kkt_error = st.kkt_error;
  out->n_resto = st.n_resto; out->n_resto_iter = st.n_resto_iter; out->n_watchdog = st.n_watchdog;
  out->n_tiny = st.n_tiny; out->n_filter_reset = st.n_filter_reset; out->n_filter_max = st.n_filter_max;
  /* MPC.cpp:322-324 */
  out->result[0] = out->z[IX(1)];
  out->result[1] = out->z[IY(1)];
  out->result[2] = out->z[IPSI(1)];
  out->result[3] = out->z[IV(1)];
  out->result[4] = out->z[IC(1)];
  out->result[5] = out->z[IE(1)];
  out->result[6] = out->z[ID(0)];
  out->result[7] = out->z[IA(0)];
  out->result[8] = out->obj;
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* batch drivers (one solve per host core, static chunks): the CPU baseline                    */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  const orc_config *cfg;
  const orc_problem *p;
  orc_result *out;
  double *result9, *traj_x, *traj_y;
  int *status, *iters;
  int B, lo, hi;
} job_t;

static void *job_run(void *arg) {
  job_t *j = (job_t *)arg;
  int N = j->cfg->N;
  orc_result tmp;
  for (int b = j->lo; b < j->hi; b++) {
    orc_result *r = j->out ? &j->out[b] : &tmp;
    orc_solve(j->cfg, &j->p[b], r);
    if (j->result9) memcpy(j->result9 + 9 * (size_t)b, r->result, 9 * sizeof(double));
    if (j->status) j->status[b] = r->status;
    if (j->iters) j->iters[b] = r->iters;
    if (j->traj_x) for (int i = 0; i < N; i++) j->traj_x[(size_t)b * N + i] = r->z[IX(i)];
    if (j->traj_y) for (int i = 0; i < N; i++) j->traj_y[(size_t)b * N + i] = r->z[IY(i)];
  }
  return 0;
}

static int run_jobs(job_t *proto, int B, int nt) {
  if (nt < 1) nt = 1;
  if (nt > 256) nt = 256;
  pthread_t th[256];
  job_t jobs[256];
  for (int t = 0; t < nt; t++) {
    jobs[t] = *proto;
    jobs[t].lo = (int)((long long)B * t / nt);
    jobs[t].hi = (int)((long long)B * (t + 1) / nt);
    pthread_create(&th[t], 0, job_run, &jobs[t]);
  }
  for (int t = 0; t < nt; t++) pthread_join(th[t], 0);
  return 0;
}

int orc_solve_batch(const orc_config *cfg, const orc_problem *p, int B, orc_result *out, int nt) {
  job_t j;
  memset(&j, 0, sizeof(j));
  j.cfg = cfg; j.p = p; j.out = out; j.B = B;
  return run_jobs(&j, B, nt);
}

int orc_solve_batch_compact(const orc_config *cfg, const orc_problem *p, int B, double *result9,
                            double *traj_x, double *traj_y, int *status, int *iters, int nt) {
  job_t j;
  memset(&j, 0, sizeof(j));
  j.cfg = cfg; j.p = p; j.result9 = result9; j.traj_x = traj_x; j.traj_y = traj_y;
  j.status = status; j.iters = iters; j.B = B;
  return run_jobs(&j, B, nt);
}
