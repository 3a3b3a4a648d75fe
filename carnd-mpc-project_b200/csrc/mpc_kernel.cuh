// mpc_kernel.cuh -- the sm_100a interior-point kernel: one MPC problem per group of G lanes.
//
// Replaces, for a whole batch at once, what the reference does once per telemetry message in
//   MPC::solve -> CppAD::ipopt::solve -> Ipopt + MUMPS   (/root/reference/src/control/MPC.cpp:183-325)
// * FG_eval's AD tape (MPC.cpp:50-154) becomes hand-derived bicycle-model residuals, Jacobian
//   entries and Lagrangian-Hessian entries, one horizon stage per lane, in registers.
// * MUMPS' sparse LDL^T of the (14N-2)x(14N-2) KKT matrix (MPC.cpp:175) becomes a stage-wise
//   Riccati recursion over the block-banded KKT system: the 7x7 cost-to-go matrix lives one row
//   per lane, stage data are broadcast from shared memory, the 2x2 control pivots give the
//   inertia test that Ipopt takes from the linear solver.
// * Ipopt's primal-dual filter line-search iteration (monotone mu, fraction to the boundary,
//   second-order correction, inertia correction, Ipopt 3.12 defaults) runs unchanged in spirit,
//   with all reductions done by warp shuffles.
// FP64 throughout (the parity contract is 1e-4 abs / 1e-6 rel against an FP64 Ipopt solve); no
// tensor cores: the work is a chain of 7x9 structured products, not a dense contraction.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace mpcb200 {

// ---- Ipopt 3.12 defaults in force in the reference (only print_level, linear_solver and
// max_cpu_time are overridden at MPC.cpp:163-178)
#define K_KAPPA_EPS 10.0
#define K_KAPPA_MU 0.2
#define K_THETA_MU 1.5
#define K_TAU_MIN 0.99
#define K_KAPPA_1 0.01
#define K_KAPPA_2 0.01
#define K_BOUND_RELAX 1e-8
#define K_S_MAX 100.0
#define K_KAPPA_SIGMA 1e10
#define K_GAMMA_THETA 1e-5
#define K_GAMMA_PHI 1e-8
#define K_S_THETA 1.1
#define K_S_PHI 2.3
#define K_ETA_PHI 1e-8
#define K_KAPPA_SOC 0.99
#define K_MAX_SOC 4
#define K_ALPHA_MIN_FRAC 0.05
#define K_DW_FIRST 1e-4
#define K_DW_MIN 1e-20
#define K_DW_MAX 1e20
#define K_DW_INC_FIRST 100.0
#define K_DW_INC 8.0
#define K_DW_DEC (1.0 / 3.0)
#define K_DUAL_INF_TOL 1.0
#define K_CONSTR_VIOL_TOL 1e-4
#define K_COMPL_INF_TOL 1e-4
#define K_ACCEPT_TOL 1e-6
#define K_ACCEPT_ITER 15
#define K_ACCEPT_CONSTR_VIOL_TOL 1e-2
#define K_ACCEPT_COMPL_INF_TOL 1e-2
#define K_CONSTR_MULT_INIT_MAX 1e3
#define K_NLP_INF 1e19
#define K_EPS 2.220446049250313e-16
#define K_NFILT 8

// per-stage fields kept in shared memory, layout SD[field][stage]
enum {
  F_A13 = 0, F_A14, F_A23, F_A24, F_A34, F_B3, F_A51, F_A54, F_A56, F_A61,                // dF/d(s,u)
  F_QXX, F_QYY, F_QPP, F_QPV, F_QVV, F_QVE, F_QCC, F_QEE, F_SVD, F_RDD, F_RAA,            // W + Sigma
  F_GP, F_GV, F_GC, F_GE, F_GDP, F_GD, F_GA,                                              // grad phi_mu
  F_D0, F_D1, F_D2, F_D3, F_D4, F_D5,                                                     // -c_{i+1}
  NFIELD
};
enum { PC_DT = 0, PC_DTLF, PC_SF, PC_CW, PC_C0, PC_S0 = PC_C0 + 5, PC_LO = PC_S0 + 6, PC_HI = PC_LO + 4,
       PC_LO0 = PC_HI + 4, PC_HI0 = PC_LO0 + 4, PC_WC2 = PC_HI0 + 4, PC_WE2, PC_WV2, PC_VREF, PC_WD2,
       PC_WC2_0, PC_WE2_0, PC_VREF_0, PC_NV2_0, PC_W, PC_SIZE = PC_W + 12 };
enum { NKK = 16, TS_LD = 10, TS_SIZE = 7 * TS_LD, MV_SIZE = 10, RB_SIZE = 4 };

struct KParams {
  int B, Nmax, max_iter, n_steers, n_steer_speeds;
  double dt, Lf, cte_panic, epsi_panic, max_speed, max_steering, max_accel, max_decel, tol;
  double weights[12];
  double steers[16];
  double steer_speeds[16];
  const double *state, *coeffs, *yaw_lo, *yaw_hi, *weights_pp;
  const int *N_pp;
  const double *dt_pp;
  double *result, *traj_x, *traj_y, *full;
  int *status, *iters;
  int *counter;
  int ws_stride;   // doubles of shared memory per problem
  // migration of long-running problems from the lane kernel to the coop kernel (mpc_lane_kernel.cuh)
  double *ckpt;      // [ckpt_cap] records of lane_ckpt_doubles(NS) doubles, or NULL
  int *ckpt_count;   // records written (may run past ckpt_cap: the surplus problems simply stay where they are)
  int *ckpt_next;    // next record the coop kernel takes
  int ckpt_cap, handoff_iter;
  // tail packing (mpc_lane_kernel.cuh): once the queue is empty, a warp with at most park_lanes problems left parks
  // them too; a resume launch reads the records of the previous launch (ckpt_in, *ckpt_in_count of them) packed 32 to
  // a warp and parks into the other buffer
  int park_lanes;
  const double *ckpt_in;
  const int *ckpt_in_count;
  // a chain of launches (main, resume..., final): counters [2j] = records parked by launch j, [2j + 1] = cursor of
  // launch j over its input; the two record buffers; this launch's position, the final launch's, and the number of
  // records below which a resume launch leaves them to the next launch
  int *chain_counts;
  double *chain_buf0, *chain_buf1;
  int chain_pos, chain_last, resume_min;
  const int *perm;   // order in which the work queue hands out the problems (ragged batches: longest horizon first), or NULL
  // optional multiplier outputs (solution.lambda / zl / zu of CppAD::ipopt::solve_result), unscaled
  double *dual_lam, *dual_zl, *dual_zu;
};

__host__ __device__ inline int workspace_doubles(int Nmax) {
  return PC_SIZE + NFIELD * Nmax + NKK * Nmax + TS_SIZE + MV_SIZE + RB_SIZE + 2 * K_NFILT;
}

// ------------------------------------------------------------------------------------------------
// group (G lanes) collectives: xor butterflies give every lane the bit-identical result, which
// keeps all control flow of a problem uniform across its lanes
// ------------------------------------------------------------------------------------------------
template <int G> __device__ __forceinline__ double gsum(double v, unsigned m) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o, G);
  return v;
}
__device__ __forceinline__ double nanmax(double a, double b) { return (a > b || a != a) ? a : b; }
template <int G> __device__ __forceinline__ double gmax(double v, unsigned m) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(m, v, o, G));
  return v;
}
template <int G> __device__ __forceinline__ double gmin(double v, unsigned m) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(m, v, o, G));
  return v;
}

// per-lane iterate of one horizon stage
struct Stage {
  double s[6];    // x, y, psi, v, cte, epsi
  double u[2];    // delta, a            (stages 0..N-2)
  double lam[6];  // multiplier of the constraint that defines s_i
  double zl[4], zu[4];  // bound multipliers of psi, v, delta, a
};
// transcendental part of a stage evaluated at a point (reused by the derivative build)
struct Trig {
  double sp, cp, se, ce, f1, f2, f3;
};

// Vehicle::computeSpeedTarget(AD<double>, double), Vehicle.cpp:50-64
__device__ inline double speed_target(const KParams &P, double angle, double mx) {
  double y = fabs(angle);
  double back = P.steer_speeds[P.n_steer_speeds - 1];
  for (int i = 0; i < P.n_steers; i++) {
    if (y <= P.steers[i]) {
      if (P.n_steer_speeds > i) return P.steer_speeds[i] < mx ? P.steer_speeds[i] : mx;
      return back < mx ? back : mx;
    }
  }
  return back < mx ? back : mx;
}

// F(s_i, u_i): right-hand sides of MPC.cpp:144-152; polyeval/polyder as utils.h:28-47
__device__ __forceinline__ void eval_point(const double *PC, const double *s, const double *u, Trig &t,
                                           double *F) {
  const double dt = PC[PC_DT], dtLf = PC[PC_DTLF];
  sincos(s[2], &t.sp, &t.cp);
  sincos(s[5], &t.se, &t.ce);
  const double c0 = PC[PC_C0], c1 = PC[PC_C0 + 1], c2 = PC[PC_C0 + 2], c3 = PC[PC_C0 + 3], c4 = PC[PC_C0 + 4];
  const double x = s[0];
  double f = (((c4 * x + c3) * x + c2) * x + c1) * x + c0;
  t.f1 = ((4.0 * c4 * x + 3.0 * c3) * x + 2.0 * c2) * x + c1;
  t.f2 = (12.0 * c4 * x + 6.0 * c3) * x + 2.0 * c2;
  t.f3 = 24.0 * c4 * x + 6.0 * c3;
  const double vdt = s[3] * dt;
  F[0] = s[0] + t.cp * vdt;
  F[1] = s[1] + t.sp * vdt;
  F[2] = s[2] + u[0] * s[3] * dtLf;
  F[3] = s[3] + u[1] * dt;
  F[4] = (f - s[1]) + t.se * vdt;
  F[5] = F[2] - atan(t.f1);
}

// out = A_i^T * ln (6 state rows) and B_i^T * ln (2 control rows), sparse pattern of App. A.4
struct Lin {
  double a13, a14, a23, a24, a34, b3, a51, a54, a56, a61;
};
__device__ __forceinline__ void At_apply(const Lin &L, double dt, const double *ln, double *os, double *ou) {
  os[0] = ln[0] + L.a51 * ln[4] + L.a61 * ln[5];
  os[1] = ln[1] - ln[4];
  os[2] = L.a13 * ln[0] + L.a23 * ln[1] + ln[2] + ln[5];
  os[3] = L.a14 * ln[0] + L.a24 * ln[1] + L.a34 * (ln[2] + ln[5]) + ln[3] + L.a54 * ln[4];
  os[4] = 0.0;
  os[5] = L.a56 * ln[4];
  ou[0] = L.b3 * (ln[2] + ln[5]);
  ou[1] = dt * ln[3];
}

// The whole per-problem solver state that is uniform over the group's lanes
struct Uni {
  double mu, tau, theta_min, theta_max, dw_last;
};

template <int G>
struct Solver {
  const KParams &P;
  double *PC, *SD, *KK, *TS, *MV, *RB, *FLT;
  const int g;         // lane within the group == horizon stage owned by this lane
  const unsigned gm;   // member mask of the group
  int N, NS;
  bool act, hasu;
  Stage z;
  Trig tg;             // trig/poly at the current iterate
  double cn[6];        // c_{i+1}(x) = s_{i+1} - F(s_i,u_i), owned by lane i
  double c0[6];        // c_0 = s_0 - state (only meaningful on lane 0)
  double ds[6], du[2], lnew[6];   // search direction: primal step and NEW multipliers
  double fx, lsum, theta;          // scaled objective, sum of log slacks, ||c||_1 at the iterate
  int nfilt;

  __device__ Solver(const KParams &P_, double *W, int g_, unsigned gm_) : P(P_), g(g_), gm(gm_) {
    NS = P.Nmax;
    PC = W;
    SD = PC + PC_SIZE;
    KK = SD + NFIELD * NS;
    TS = KK + NKK * NS;
    MV = TS + TS_SIZE;
    RB = MV + MV_SIZE;
    FLT = RB + RB_SIZE;
  }
  __device__ __forceinline__ void sync() { __syncwarp(gm); }
  __device__ __forceinline__ double down(double v) { return __shfl_down_sync(gm, v, 1, G); }
  __device__ __forceinline__ double up(double v) { return __shfl_up_sync(gm, v, 1, G); }
  __device__ __forceinline__ double bcast(double v, int src) { return __shfl_sync(gm, v, src, G); }

  __device__ __forceinline__ double var4(const double *s, const double *u, int k) const {
    return k == 0 ? s[2] : (k == 1 ? s[3] : (k == 2 ? u[0] : u[1]));
  }
  __device__ __forceinline__ bool valid4(int k) const { return k < 2 ? act : hasu; }
  __device__ __forceinline__ double wc2() const { return g == 0 ? PC[PC_WC2_0] : PC[PC_WC2]; }
  __device__ __forceinline__ double we2() const { return g == 0 ? PC[PC_WE2_0] : PC[PC_WE2]; }
  __device__ __forceinline__ double vref() const { return g == 0 ? PC[PC_VREF_0] : PC[PC_VREF]; }
  __device__ __forceinline__ double nv2() const { return g == 0 ? PC[PC_NV2_0] : 0.0; }

  // residuals, scaled objective and log-barrier sum at a point (s,u given per lane)
  __device__ void point_metrics(const double *s, const double *u, Trig &t, double *cnext, double *c0_,
                                double &f_out, double &lsum_out, double &theta_out) {
    double F[6];
    if (hasu) {
      eval_point(PC, s, u, t, F);
    } else {
      t.sp = t.cp = t.se = t.ce = t.f1 = t.f2 = t.f3 = 0.0;
#pragma unroll
      for (int k = 0; k < 6; k++) F[k] = 0.0;
    }
    double th = 0.0;
#pragma unroll
    for (int k = 0; k < 6; k++) {
      double sn = down(s[k]);
      cnext[k] = hasu ? sn - F[k] : 0.0;
      c0_[k] = (g == 0) ? s[k] - PC[PC_S0 + k] : 0.0;
      th += fabs(cnext[k]) + fabs(c0_[k]);
    }
    double fl = 0.0, prod = 1.0;
    if (act) {
      double dv = s[3] - vref();
      fl = 0.5 * (wc2() * s[4] * s[4] + we2() * s[5] * s[5] + PC[PC_WV2] * dv * dv + nv2() * s[3] * s[3]);
      prod = (s[2] - PC[PC_LO]) * (PC[PC_HI] - s[2]) * (s[3] - PC[PC_LO + 1]) * (PC[PC_HI + 1] - s[3]);
    }
    double dprev = up(u[0]);
    if (hasu) {
      fl += 0.5 * PC[PC_WD2] * u[0] * u[0];
      if (g >= 1) {
        double dd = u[0] - dprev;
        fl += 0.5 * PC[PC_CW] * dd * dd;
      }
      prod *= (u[0] - PC[PC_LO + 2]) * (PC[PC_HI + 2] - u[0]) * (u[1] - PC[PC_LO + 3]) * (PC[PC_HI + 3] - u[1]);
    }
    double ll = act ? log(prod) : 0.0;
    f_out = gsum<G>(fl, gm);
    lsum_out = gsum<G>(ll, gm);
    theta_out = gsum<G>(th, gm);
  }

  __device__ __forceinline__ void load_lin(Lin &L) const {
    L.a13 = SD[F_A13 * NS + g]; L.a14 = SD[F_A14 * NS + g]; L.a23 = SD[F_A23 * NS + g];
    L.a24 = SD[F_A24 * NS + g]; L.a34 = SD[F_A34 * NS + g]; L.b3 = SD[F_B3 * NS + g];
    L.a51 = SD[F_A51 * NS + g]; L.a54 = SD[F_A54 * NS + g]; L.a56 = SD[F_A56 * NS + g];
    L.a61 = SD[F_A61 * NS + g];
  }

  // Build this lane's stage fields in shared memory: dF/d(s,u), Hessian of the Lagrangian plus
  // barrier Sigma (no delta_w: the Riccati adds it), gradient of the barrier objective.
  // ls_mode: the least-squares multiplier system [[I, J^T],[J, 0]] (Hessian = I, no barrier).
  __device__ void build_fields(double mu, bool ls_mode) {
    const double dt = PC[PC_DT], dtLf = PC[PC_DTLF];
    double ln[6];
#pragma unroll
    for (int k = 0; k < 6; k++) ln[k] = down(z.lam[k]);
    const double dprev = up(z.u[0]);
    const bool cpl = hasu && g >= 1;
    double sig[4], gb[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      double xv = var4(z.s, z.u, k);
      if (valid4(k)) {
        double il = 1.0 / (xv - PC[PC_LO + k]), iu = 1.0 / (PC[PC_HI + k] - xv);
        sig[k] = z.zl[k] * il + z.zu[k] * iu;
        gb[k] = ls_mode ? (z.zu[k] - z.zl[k]) : mu * (iu - il);
      } else {
        sig[k] = 0.0;
        gb[k] = 0.0;
      }
    }
    if (!act) return;
    const double v = z.s[3], vdt = v * dt;
    double *o = SD + g;
    if (hasu) {
      const double q = 1.0 + tg.f1 * tg.f1, iq = 1.0 / q;
      o[F_A13 * NS] = -vdt * tg.sp;
      o[F_A14 * NS] = dt * tg.cp;
      o[F_A23 * NS] = vdt * tg.cp;
      o[F_A24 * NS] = dt * tg.sp;
      o[F_A34 * NS] = z.u[0] * dtLf;
      o[F_B3 * NS] = v * dtLf;
      o[F_A51 * NS] = tg.f1;
      o[F_A54 * NS] = dt * tg.se;
      o[F_A56 * NS] = vdt * tg.ce;
      o[F_A61 * NS] = -tg.f2 * iq;
      if (!ls_mode) {
        o[F_QXX * NS] = -ln[4] * tg.f2 + ln[5] * (tg.f3 * q - 2.0 * tg.f1 * tg.f2 * tg.f2) * iq * iq;
        o[F_QPP * NS] = (ln[0] * tg.cp + ln[1] * tg.sp) * vdt + sig[0];
        o[F_QPV * NS] = (ln[0] * tg.sp - ln[1] * tg.cp) * dt;
        o[F_QVE * NS] = -ln[4] * tg.ce * dt;
        o[F_QEE * NS] = ln[4] * tg.se * vdt + we2();
        o[F_SVD * NS] = -(ln[2] + ln[5]) * dtLf;
        o[F_RDD * NS] = PC[PC_WD2] + (cpl ? PC[PC_CW] : 0.0) + sig[2];
        o[F_RAA * NS] = sig[3];
      }
      const double dd = z.u[0] - dprev;
      o[F_GDP * NS] = cpl ? -PC[PC_CW] * dd : 0.0;
      o[F_GD * NS] = PC[PC_WD2] * z.u[0] + (cpl ? PC[PC_CW] * dd : 0.0) + gb[2];
      o[F_GA * NS] = gb[3];
    } else if (!ls_mode) {
      o[F_QXX * NS] = 0.0;
      o[F_QPP * NS] = sig[0];
      o[F_QPV * NS] = 0.0;
      o[F_QVE * NS] = 0.0;
      o[F_QEE * NS] = we2();
    }
    if (!ls_mode) {
      o[F_QYY * NS] = 0.0;
      o[F_QVV * NS] = PC[PC_WV2] + nv2() + sig[1];
      o[F_QCC * NS] = wc2();
    } else {
      o[F_QXX * NS] = 1.0; o[F_QYY * NS] = 1.0; o[F_QPP * NS] = 1.0; o[F_QPV * NS] = 0.0;
      o[F_QVV * NS] = 1.0; o[F_QVE * NS] = 0.0; o[F_QCC * NS] = 1.0; o[F_QEE * NS] = 1.0;
      o[F_SVD * NS] = 0.0; o[F_RDD * NS] = 1.0; o[F_RAA * NS] = 1.0;
    }
    o[F_GP * NS] = gb[0];
    o[F_GV * NS] = PC[PC_WV2] * (v - vref()) + nv2() * v + gb[1];
    o[F_GC * NS] = wc2() * z.s[4];
    o[F_GE * NS] = we2() * z.s[5];
  }

  // write the constraint part of the KKT right-hand side: d_i = -c_{i+1}
  __device__ __forceinline__ void store_d(const double *cnext) {
    if (hasu) {
#pragma unroll
      for (int k = 0; k < 6; k++) SD[(F_D0 + k) * NS + g] = -cnext[k];
    }
  }

  // Backward Riccati sweep over the stages.  Lane r < 7 carries row r of the 7x7 cost-to-go P of the
  // augmented state (ds, ddelta_prev); lane 9 carries the linear term.  cw = delta-rate coupling
  // Hessian (2*sf*w4), dw = Ipopt's delta_w.  Returns whether every 2x2 control pivot was positive
  // definite, i.e. whether the KKT matrix has inertia (n, m, 0).
  __device__ bool riccati(double dw, double cw) {
    double Pn[7], pn;
    {
      const int t = N - 1;
#pragma unroll
      for (int c = 0; c < 7; c++) Pn[c] = 0.0;
      pn = 0.0;
      if (g == 0) Pn[0] = SD[F_QXX * NS + t] + dw;
      if (g == 1) Pn[1] = SD[F_QYY * NS + t] + dw;
      if (g == 2) { Pn[2] = SD[F_QPP * NS + t] + dw; pn = SD[F_GP * NS + t]; }
      if (g == 3) { Pn[3] = SD[F_QVV * NS + t] + dw; pn = SD[F_GV * NS + t]; }
      if (g == 4) { Pn[4] = SD[F_QCC * NS + t] + dw; pn = SD[F_GC * NS + t]; }
      if (g == 5) { Pn[5] = SD[F_QEE * NS + t] + dw; pn = SD[F_GE * NS + t]; }
    }
    bool ok = true;
    const double dt = PC[PC_DT];
    for (int i = N - 2; i >= 0; i--) {
      const double a13 = SD[F_A13 * NS + i], a14 = SD[F_A14 * NS + i], a23 = SD[F_A23 * NS + i];
      const double a24 = SD[F_A24 * NS + i], a34 = SD[F_A34 * NS + i], b3 = SD[F_B3 * NS + i];
      const double a51 = SD[F_A51 * NS + i], a54 = SD[F_A54 * NS + i], a56 = SD[F_A56 * NS + i];
      const double a61 = SD[F_A61 * NS + i];
      // T = P+ * Atilde (row r), plus the vector column v+ = P+ d + p+
      {
        double vv = pn;
#pragma unroll
        for (int k = 0; k < 6; k++) vv += Pn[k] * SD[(F_D0 + k) * NS + i];
        const double p25 = Pn[2] + Pn[5];
        if (g < 7) {
          double *tr = TS + g * TS_LD;
          tr[0] = Pn[0] + a51 * Pn[4] + a61 * Pn[5];
          tr[1] = Pn[1] - Pn[4];
          tr[2] = a13 * Pn[0] + a23 * Pn[1] + p25;
          tr[3] = a14 * Pn[0] + a24 * Pn[1] + a34 * p25 + Pn[3] + a54 * Pn[4];
          tr[4] = 0.0;
          tr[5] = a56 * Pn[4];
          tr[6] = 0.0;
          tr[7] = b3 * p25 + Pn[6];
          tr[8] = dt * Pn[3];
          tr[9] = vv;
        }
      }
      sync();
      // lane b < 10 takes column b of T and forms row b of M = Atilde^T T (+ H), b = 9: the vector
      double M[9];
      {
        const int b = g < 9 ? g : 9;
        double Tc[7];
#pragma unroll
        for (int k = 0; k < 7; k++) Tc[k] = TS[k * TS_LD + b];
        const double t25 = Tc[2] + Tc[5];
        M[0] = Tc[0] + a51 * Tc[4] + a61 * Tc[5];
        M[1] = Tc[1] - Tc[4];
        M[2] = a13 * Tc[0] + a23 * Tc[1] + t25;
        M[3] = a14 * Tc[0] + a24 * Tc[1] + a34 * t25 + Tc[3] + a54 * Tc[4];
        M[4] = 0.0;
        M[5] = a56 * Tc[4];
        M[6] = 0.0;
        M[7] = b3 * t25 + Tc[6];
        M[8] = dt * Tc[3];
      }
      const bool cpl = i >= 1;
      switch (g) {
        case 0: M[0] += SD[F_QXX * NS + i] + dw; break;
        case 1: M[1] += SD[F_QYY * NS + i] + dw; break;
        case 2: M[2] += SD[F_QPP * NS + i] + dw; M[3] += SD[F_QPV * NS + i]; break;
        case 3: M[2] += SD[F_QPV * NS + i]; M[3] += SD[F_QVV * NS + i] + dw; M[5] += SD[F_QVE * NS + i];
                M[7] += SD[F_SVD * NS + i]; break;
        case 4: M[4] += SD[F_QCC * NS + i] + dw; break;
        case 5: M[3] += SD[F_QVE * NS + i]; M[5] += SD[F_QEE * NS + i] + dw; break;
        case 6: if (cpl) { M[6] += cw; M[7] -= cw; } break;
        case 7: M[3] += SD[F_SVD * NS + i]; if (cpl) M[6] -= cw; M[7] += SD[F_RDD * NS + i] + dw;
                RB[0] = M[7]; RB[1] = M[8]; break;
        case 8: M[8] += SD[F_RAA * NS + i] + dw; RB[2] = M[8]; break;
        case 9: M[2] += SD[F_GP * NS + i]; M[3] += SD[F_GV * NS + i]; M[4] += SD[F_GC * NS + i];
                M[5] += SD[F_GE * NS + i]; M[6] += SD[F_GDP * NS + i]; M[7] += SD[F_GD * NS + i];
                M[8] += SD[F_GA * NS + i];
#pragma unroll
                for (int k = 0; k < 7; k++) MV[k] = M[k];
                break;
        default: break;
      }
      sync();
      const double r11 = RB[0], r12 = RB[1], r22 = RB[2];
      const double det = r11 * r22 - r12 * r12;
      ok = ok && (r11 > 0.0) && (det > 0.0);
      const double idet = 1.0 / det;
      // gains: column c of K (lanes 0..6) and the feed-forward k (lane 9) share one formula
      const double K0 = -(r22 * M[7] - r12 * M[8]) * idet;
      const double K1 = -(r11 * M[8] - r12 * M[7]) * idet;
      if (g < 7) { KK[g * NS + i] = K0; KK[(7 + g) * NS + i] = K1; }
      if (g == 9) { KK[14 * NS + i] = K0; KK[15 * NS + i] = K1; }
      sync();
      // Schur complement: P[r][c] = M[r][c] + M[r][7] K0[c] + M[r][8] K1[c]
      if (g < 7) {
#pragma unroll
        for (int c = 0; c < 7; c++) Pn[c] = M[c] + M[7] * KK[c * NS + i] + M[8] * KK[(7 + c) * NS + i];
        pn = MV[g] + M[7] * KK[14 * NS + i] + M[8] * KK[15 * NS + i];
      }
    }
    return ok;
  }

  // Forward sweep (every lane runs the recursion; lane i keeps stage i's step), then the backward
  // costate recursion for the new multipliers.  t0 = ds_0.
  __device__ void forward_and_costate(const double *t0_lane0, double dw, double cw) {
    const double dt = PC[PC_DT];
    double t[7];
#pragma unroll
    for (int k = 0; k < 6; k++) t[k] = bcast(t0_lane0[k], 0);
    t[6] = 0.0;
    double dup = 0.0;   // d delta_{i-1}
    du[0] = du[1] = 0.0;
    for (int i = 0; i < N - 1; i++) {
      double u0 = KK[14 * NS + i], u1 = KK[15 * NS + i];
#pragma unroll
      for (int c = 0; c < 7; c++) {
        u0 += KK[c * NS + i] * t[c];
        u1 += KK[(7 + c) * NS + i] * t[c];
      }
      if (g == i) {
#pragma unroll
        for (int k = 0; k < 6; k++) ds[k] = t[k];
        du[0] = u0;
        du[1] = u1;
        dup = t[6];
      }
      const double a13 = SD[F_A13 * NS + i], a14 = SD[F_A14 * NS + i], a23 = SD[F_A23 * NS + i];
      const double a24 = SD[F_A24 * NS + i], a34 = SD[F_A34 * NS + i], b3 = SD[F_B3 * NS + i];
      const double a51 = SD[F_A51 * NS + i], a54 = SD[F_A54 * NS + i], a56 = SD[F_A56 * NS + i];
      const double a61 = SD[F_A61 * NS + i];
      double n0 = t[0] + a13 * t[2] + a14 * t[3] + SD[(F_D0 + 0) * NS + i];
      double n1 = t[1] + a23 * t[2] + a24 * t[3] + SD[(F_D0 + 1) * NS + i];
      double n2 = t[2] + a34 * t[3] + b3 * u0 + SD[(F_D0 + 2) * NS + i];
      double n3 = t[3] + dt * u1 + SD[(F_D0 + 3) * NS + i];
      double n4 = a51 * t[0] - t[1] + a54 * t[3] + a56 * t[5] + SD[(F_D0 + 4) * NS + i];
      double n5 = a61 * t[0] + t[2] + a34 * t[3] + b3 * u0 + SD[(F_D0 + 5) * NS + i];
      t[0] = n0; t[1] = n1; t[2] = n2; t[3] = n3; t[4] = n4; t[5] = n5; t[6] = u0;
    }
    if (g == N - 1) {
#pragma unroll
      for (int k = 0; k < 6; k++) ds[k] = t[k];
      dup = t[6];
    }
    if (!act) {
#pragma unroll
      for (int k = 0; k < 6; k++) ds[k] = 0.0;
    }
    // h = (H + Sigma + dw I) * step + grad, state rows of this lane's stage
    double h[6];
    Lin L;
    if (act) {
      const double *o = SD + g;
      const double qpv = o[F_QPV * NS], qve = o[F_QVE * NS];
      h[0] = (o[F_QXX * NS] + dw) * ds[0];
      h[1] = (o[F_QYY * NS] + dw) * ds[1];
      h[2] = (o[F_QPP * NS] + dw) * ds[2] + qpv * ds[3] + o[F_GP * NS];
      h[3] = qpv * ds[2] + (o[F_QVV * NS] + dw) * ds[3] + qve * ds[5] + o[F_GV * NS];
      h[4] = (o[F_QCC * NS] + dw) * ds[4] + o[F_GC * NS];
      h[5] = qve * ds[3] + (o[F_QEE * NS] + dw) * ds[5] + o[F_GE * NS];
      if (hasu) {
        h[3] += o[F_SVD * NS] * du[0];
        load_lin(L);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 6; k++) h[k] = 0.0;
    }
    (void)dup; (void)cw;
#pragma unroll
    for (int k = 0; k < 6; k++) lnew[k] = -h[k];
    for (int i = N - 2; i >= 0; i--) {
      double nx[6];
#pragma unroll
      for (int k = 0; k < 6; k++) nx[k] = bcast(lnew[k], i + 1);
      if (g == i) {
        double os[6], ou[2];
        At_apply(L, dt, nx, os, ou);
#pragma unroll
        for (int k = 0; k < 6; k++) lnew[k] = os[k] - h[k];
      }
    }
  }

  // factor + solve for the Newton direction with Ipopt's inertia correction; false if it gave up
  __device__ bool direction(Uni &U, const double *cnext, const double *c0_, double &dw_used) {
    store_d(cnext);
    sync();
    const double cw = PC[PC_CW];
    double dw = 0.0;
    bool ok = riccati(0.0, cw);
    if (!ok) {
      bool first = true;
      for (;;) {
        if (first) {
          dw = (U.dw_last == 0.0) ? K_DW_FIRST : fmax(K_DW_MIN, U.dw_last * K_DW_DEC);
          first = false;
        } else {
          dw = (U.dw_last == 0.0) ? dw * K_DW_INC_FIRST : dw * K_DW_INC;
        }
        if (dw > K_DW_MAX) break;
        sync();
        ok = riccati(dw, cw);
        if (ok) break;
      }
      if (!ok) return false;
      U.dw_last = dw;
    }
    dw_used = dw;
    double t0[6];
#pragma unroll
    for (int k = 0; k < 6; k++) t0[k] = -c0_[k];
    forward_and_costate(t0, dw, cw);
    return true;
  }

  // primal-dual error terms at the iterate: ||grad L||_inf, ||c||_inf, sums for s_d / s_c
  __device__ void kkt_errors(double &dinf, double &cviol, double &lam1, double &z1) {
    double ln[6];
#pragma unroll
    for (int k = 0; k < 6; k++) ln[k] = down(z.lam[k]);
    const double dprev = up(z.u[0]), dnext = down(z.u[0]);
    double r = 0.0, cv = 0.0, l1 = 0.0, zz = 0.0;
    if (act) {
      double os[6] = {0, 0, 0, 0, 0, 0}, ou[2] = {0, 0};
      if (hasu) {
        Lin L;
        load_lin(L);
        At_apply(L, PC[PC_DT], ln, os, ou);
      }
      const double v = z.s[3];
      double gs[6];
      gs[0] = 0.0; gs[1] = 0.0;
      gs[2] = -z.zl[0] + z.zu[0];
      gs[3] = PC[PC_WV2] * (v - vref()) + nv2() * v - z.zl[1] + z.zu[1];
      gs[4] = wc2() * z.s[4];
      gs[5] = we2() * z.s[5];
#pragma unroll
      for (int k = 0; k < 6; k++) {
        r = nanmax(r, fabs(gs[k] + z.lam[k] - os[k]));
        l1 += fabs(z.lam[k]);
        cv = nanmax(cv, nanmax(fabs(cn[k]), fabs(c0[k])));
      }
      zz = fabs(z.zl[0]) + fabs(z.zu[0]) + fabs(z.zl[1]) + fabs(z.zu[1]);
      if (hasu) {
        double gd = PC[PC_WD2] * z.u[0];
        if (g >= 1) gd += PC[PC_CW] * (z.u[0] - dprev);
        if (g <= N - 3) gd -= PC[PC_CW] * (dnext - z.u[0]);
        r = nanmax(r, fabs(gd - ou[0] - z.zl[2] + z.zu[2]));
        r = nanmax(r, fabs(-ou[1] - z.zl[3] + z.zu[3]));
        zz += fabs(z.zl[2]) + fabs(z.zu[2]) + fabs(z.zl[3]) + fabs(z.zu[3]);
      }
    }
    dinf = gmax<G>(r, gm);
    cviol = gmax<G>(cv, gm);
    lam1 = gsum<G>(l1, gm);
    z1 = gsum<G>(zz, gm);
  }
  __device__ double compl_err(double mu) {
    double cp = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (valid4(k)) {
        double xv = var4(z.s, z.u, k);
        cp = nanmax(cp, fabs((xv - PC[PC_LO + k]) * z.zl[k] - mu));
        cp = nanmax(cp, fabs((PC[PC_HI + k] - xv) * z.zu[k] - mu));
      }
    }
    return gmax<G>(cp, gm);
  }

  // largest alpha in (0,1] keeping x + alpha dx inside the (relaxed) bounds by the fraction tau
  __device__ double frac_to_bound(double tau) {
    double a = 1.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (valid4(k)) {
        double xv = var4(z.s, z.u, k), dx = var4(ds, du, k);
        if (dx < 0.0) a = fmin(a, -tau * (xv - PC[PC_LO + k]) / dx);
        if (dx > 0.0) a = fmin(a, tau * (PC[PC_HI + k] - xv) / dx);
      }
    }
    return gmin<G>(a, gm);
  }
  __device__ void dual_steps(double mu, double *dzl, double *dzu) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (valid4(k)) {
        double xv = var4(z.s, z.u, k), dx = var4(ds, du, k);
        double il = 1.0 / (xv - PC[PC_LO + k]), iu = 1.0 / (PC[PC_HI + k] - xv);
        dzl[k] = mu * il - z.zl[k] - z.zl[k] * il * dx;
        dzu[k] = mu * iu - z.zu[k] + z.zu[k] * iu * dx;
      } else {
        dzl[k] = dzu[k] = 0.0;
      }
    }
  }
  __device__ double frac_to_bound_dual(double tau, const double *dzl, const double *dzu) {
    double a = 1.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (valid4(k)) {
        if (dzl[k] < 0.0) a = fmin(a, -tau * z.zl[k] / dzl[k]);
        if (dzu[k] < 0.0) a = fmin(a, -tau * z.zu[k] / dzu[k]);
      }
    }
    return gmin<G>(a, gm);
  }
  // grad(phi_mu)^T dx
  __device__ double grad_barrier_dot(double mu) {
    double acc = 0.0;
    if (act) {
      const double v = z.s[3];
      acc = (PC[PC_WV2] * (v - vref()) + nv2() * v) * ds[3] + wc2() * z.s[4] * ds[4] + we2() * z.s[5] * ds[5];
    }
    const double dprev = up(z.u[0]), dnext = down(z.u[0]);
    if (hasu) {
      double gd = PC[PC_WD2] * z.u[0];
      if (g >= 1) gd += PC[PC_CW] * (z.u[0] - dprev);
      if (g <= N - 3) gd -= PC[PC_CW] * (dnext - z.u[0]);
      acc += gd * du[0];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (valid4(k)) {
        double xv = var4(z.s, z.u, k), dx = var4(ds, du, k);
        acc += mu * (1.0 / (PC[PC_HI + k] - xv) - 1.0 / (xv - PC[PC_LO + k])) * dx;
      }
    }
    return gsum<G>(acc, gm);
  }

  // ---- filter (Ipopt's FilterLSAcceptor), entries in shared memory, evaluated by every lane
  __device__ __forceinline__ static bool cmp_le(double lhs, double rhs, double basis) {
    return lhs - rhs <= 10.0 * K_EPS * fabs(basis);
  }
  struct LS {
    double theta, phi, gbd;
  };
  __device__ bool is_ftype(const LS &l, double alpha) const {
    return l.gbd < 0.0 && alpha * pow(-l.gbd, K_S_PHI) > pow(l.theta, K_S_THETA);
  }
  __device__ bool armijo(const LS &l, double alpha, double phi_t) const {
    return cmp_le(phi_t - l.phi, K_ETA_PHI * alpha * l.gbd, l.phi);
  }
  __device__ bool ls_accept(const LS &l, const Uni &U, double alpha, double theta_t, double phi_t) const {
    if (!(theta_t == theta_t) || !(phi_t == phi_t) || isinf(phi_t)) return false;
    if (theta_t > U.theta_max) return false;
    bool ok;
    if (alpha > 0.0 && is_ftype(l, alpha) && l.theta <= U.theta_min) {
      ok = armijo(l, alpha, phi_t);
    } else {
      ok = cmp_le(theta_t, (1.0 - K_GAMMA_THETA) * l.theta, l.theta) ||
           cmp_le(phi_t - l.phi, -K_GAMMA_PHI * l.theta, l.phi);
    }
    if (!ok) return false;
    for (int k = 0; k < nfilt; k++) {
      double ft = FLT[2 * k], fp = FLT[2 * k + 1];
      if (!(cmp_le(theta_t, ft, ft) || cmp_le(phi_t, fp, fp))) return false;
    }
    return true;
  }
  __device__ void filter_add(double theta, double phi) {
    // every lane computes the same compaction; lane 0 writes it back
    double nt[K_NFILT], np[K_NFILT];
    int k = 0;
    for (int j = 0; j < nfilt; j++) {
      double ft = FLT[2 * j], fp = FLT[2 * j + 1];
      if (!(ft >= theta && fp >= phi)) { nt[k] = ft; np[k] = fp; k++; }
    }
    if (k == K_NFILT) {   // full: drop the oldest entry
      for (int j = 1; j < k; j++) { nt[j - 1] = nt[j]; np[j - 1] = np[j]; }
      k--;
    }
    nt[k] = theta; np[k] = phi; k++;
    sync();
    if (g == 0)
      for (int j = 0; j < k; j++) { FLT[2 * j] = nt[j]; FLT[2 * j + 1] = np[j]; }
    nfilt = k;
    sync();
  }

  // ================================================================================================
  __device__ void solve(int b) {
    const int B = P.B;
    // ---- problem inputs -> shared constants (coalescing is irrelevant here: 104 B per problem)
    N = P.N_pp ? P.N_pp[b] : P.Nmax;
    if (N > P.Nmax) N = P.Nmax;
    if (N < 2) N = 2;
    act = g < N;
    hasu = g < N - 1;
    sync();
    if (g < 6) PC[PC_S0 + g] = P.state[(size_t)g * B + b];
    else if (g < 11) PC[PC_C0 + (g - 6)] = P.coeffs[(size_t)(g - 6) * B + b];
    if (g < 12) PC[PC_W + g] = P.weights_pp ? P.weights_pp[(size_t)g * B + b] : P.weights[g];
    const double ylo = P.yaw_lo[b], yhi = P.yaw_hi[b];
    const double dt = P.dt_pp ? P.dt_pp[b] : P.dt;
    sync();
    {
      // frozen branches of FG_eval at the start point (MPC.cpp:72,79,87,89) and Ipopt's
      // gradient-based objective scaling (nlp_scaling_max_gradient = 100)
      const double *w = PC + PC_W;
      const double cte0 = PC[PC_S0 + 4], epsi0 = PC[PC_S0 + 5], psi0 = PC[PC_S0 + 2], v0 = PC[PC_S0 + 3];
      const double wc0 = fabs(cte0) < P.cte_panic ? w[0] : w[11];
      const double wcN = 0.0 < P.cte_panic ? w[0] : w[11];
      const double we0 = fabs(epsi0) > P.epsi_panic ? w[10] : w[1];
      const double weN = 0.0 > P.epsi_panic ? w[10] : w[1];
      const double vr0 = speed_target(P, psi0, P.max_speed), vrN = speed_target(P, 0.0, P.max_speed);
      const double nvw0 = v0 < 0.0 ? w[9] : 0.0;
      double gmx = fmax(fabs(2.0 * wc0 * cte0), fabs(2.0 * we0 * epsi0));
      gmx = fmax(gmx, fabs(2.0 * w[2] * (v0 - vr0) + 2.0 * nvw0 * v0));
      gmx = fmax(gmx, fabs(2.0 * w[2] * vrN));
      const double sf = gmx > 100.0 ? fmax(100.0 / gmx, 1e-8) : 1.0;
      const double lo0[4] = {ylo, -P.max_speed, -P.max_steering, P.max_decel};
      const double hi0[4] = {yhi, P.max_speed, P.max_steering, P.max_accel};
      if (g == 0) {
        PC[PC_DT] = dt;
        PC[PC_DTLF] = dt / P.Lf;
        PC[PC_SF] = sf;
        PC[PC_CW] = 2.0 * sf * w[4];
        PC[PC_WC2] = 2.0 * sf * wcN; PC[PC_WE2] = 2.0 * sf * weN; PC[PC_WV2] = 2.0 * sf * w[2];
        PC[PC_VREF] = vrN; PC[PC_WD2] = 2.0 * sf * w[3];
        PC[PC_WC2_0] = 2.0 * sf * wc0; PC[PC_WE2_0] = 2.0 * sf * we0; PC[PC_VREF_0] = vr0;
        PC[PC_NV2_0] = 2.0 * sf * nvw0;
        for (int k = 0; k < 4; k++) {
          PC[PC_LO0 + k] = lo0[k];
          PC[PC_HI0 + k] = hi0[k];
          PC[PC_LO + k] = lo0[k] - fmin(K_CONSTR_VIOL_TOL, K_BOUND_RELAX * fmax(1.0, fabs(lo0[k])));
          PC[PC_HI + k] = hi0[k] + fmin(K_CONSTR_VIOL_TOL, K_BOUND_RELAX * fmax(1.0, fabs(hi0[k])));
        }
      }
    }
    sync();
    // ---- start point (MPC.cpp:207-218) pushed into the interior, multipliers
#pragma unroll
    for (int k = 0; k < 6; k++) { z.s[k] = (g == 0) ? PC[PC_S0 + k] : 0.0; z.lam[k] = 0.0; }
    z.u[0] = z.u[1] = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double lo = PC[PC_LO + k], hi = PC[PC_HI + k], span = hi - lo;
      const double pl = fmin(K_KAPPA_1 * fmax(1.0, fabs(lo)), K_KAPPA_2 * span);
      const double pu = fmin(K_KAPPA_1 * fmax(1.0, fabs(hi)), K_KAPPA_2 * span);
      double v = var4(z.s, z.u, k);
      if (v < lo + pl) v = lo + pl;
      if (v > hi - pu) v = hi - pu;
      if (valid4(k)) {
        if (k == 0) z.s[2] = v; else if (k == 1) z.s[3] = v; else if (k == 2) z.u[0] = v; else z.u[1] = v;
      }
      z.zl[k] = z.zu[k] = valid4(k) ? 1.0 : 0.0;
    }
    Uni U;
    U.mu = 0.1;
    U.tau = fmax(K_TAU_MIN, 1.0 - U.mu);
    U.dw_last = 0.0;
    nfilt = 0;
    point_metrics(z.s, z.u, tg, cn, c0, fx, lsum, theta);
    U.theta_max = 1e4 * fmax(1.0, theta);
    U.theta_min = 1e-4 * fmax(1.0, theta);

    // ---- least-squares multipliers: [[I, J^T],[J, 0]] [r; lam] = [-(grad f - zl + zu); 0]
    {
      build_fields(0.0, true);
      double zero6[6] = {0, 0, 0, 0, 0, 0};
      if (hasu) {
#pragma unroll
        for (int k = 0; k < 6; k++) SD[(F_D0 + k) * NS + g] = 0.0;
      }
      sync();
      riccati(0.0, 0.0);
      forward_and_costate(zero6, 0.0, 0.0);
      double lm = 0.0;
#pragma unroll
      for (int k = 0; k < 6; k++) lm = nanmax(lm, act ? fabs(lnew[k]) : 0.0);
      lm = gmax<G>(lm, gm);
      const bool bad = !(lm <= K_CONSTR_MULT_INIT_MAX);
#pragma unroll
      for (int k = 0; k < 6; k++) z.lam[k] = (bad || !act) ? 0.0 : lnew[k];
      sync();
    }

    int status = 0, iter = 0, accept_cnt = 0;
    const double sf = PC[PC_SF];
    const int nz = 4 * N + 4 * (N - 1), m = 6 * N;
    double E0 = 0.0;
    for (;;) {
      // fields at the current iterate carry A (needed by the error evaluation); mu-dependent parts
      // are rebuilt after the barrier update
      build_fields(U.mu, false);
      sync();
      double dinf, cviol, lam1, z1;
      kkt_errors(dinf, cviol, lam1, z1);
      const double sd = fmax(K_S_MAX, (lam1 + z1) / (double)(m + nz)) / K_S_MAX;
      const double sc = fmax(K_S_MAX, z1 / (double)nz) / K_S_MAX;
      const double compl0 = compl_err(0.0);
      E0 = nanmax(dinf / sd, nanmax(cviol, compl0 / sc));
      {
        const double dinf_u = dinf / sf, compl_u = compl0 / sf;
        if (E0 <= P.tol && dinf_u <= K_DUAL_INF_TOL && cviol <= K_CONSTR_VIOL_TOL && compl_u <= K_COMPL_INF_TOL) {
          status = 1;
          break;
        }
        if (E0 <= K_ACCEPT_TOL && cviol <= K_ACCEPT_CONSTR_VIOL_TOL && compl_u <= K_ACCEPT_COMPL_INF_TOL) {
          if (++accept_cnt >= K_ACCEPT_ITER) { status = 4; break; }
        } else {
          accept_cnt = 0;
        }
      }
      if (!(E0 == E0)) { status = 11; break; }
      if (iter >= P.max_iter) { status = 2; break; }

      // ---- monotone barrier update with fast decrease
      bool mu_changed = false;
      for (;;) {
        const double cm = compl_err(U.mu);
        const double Emu = nanmax(dinf / sd, nanmax(cviol, cm / sc));
        if (!(Emu <= K_KAPPA_EPS * U.mu)) break;
        const double mu_min = fmin(P.tol, K_COMPL_INF_TOL) / (K_KAPPA_EPS + 1.0);
        const double new_mu = fmax(mu_min, fmin(K_KAPPA_MU * U.mu, pow(U.mu, K_THETA_MU)));
        if (new_mu == U.mu) break;
        U.mu = new_mu;
        U.tau = fmax(K_TAU_MIN, 1.0 - U.mu);
        nfilt = 0;
        mu_changed = true;
      }
      if (mu_changed) {
        sync();
        build_fields(U.mu, false);
      }

      // ---- search direction
      double dw_used;
      if (!direction(U, cn, c0, dw_used)) { status = 10; break; }
      double dzl[4], dzu[4];
      dual_steps(U.mu, dzl, dzu);
      const double alpha_max = frac_to_bound(U.tau);
      double alpha_z = frac_to_bound_dual(U.tau, dzl, dzu);

      // ---- filter line search
      LS ls;
      ls.theta = theta;
      ls.phi = fx - U.mu * lsum;
      ls.gbd = grad_barrier_dot(U.mu);
      double alpha_min = K_GAMMA_THETA;
      if (ls.gbd < 0.0) {
        alpha_min = fmin(K_GAMMA_THETA, K_GAMMA_PHI * ls.theta / (-ls.gbd));
        if (ls.theta <= U.theta_min)
          alpha_min = fmin(alpha_min, pow(ls.theta, K_S_THETA) / pow(-ls.gbd, K_S_PHI));
      }
      alpha_min *= K_ALPHA_MIN_FRAC;

      double alpha = alpha_max, alpha_test = alpha_max;
      bool accepted = false;
      int ntrial = 0;
      double st[6], ut[2], cnt_[6], c0t[6], ft, lt, tht;
      Trig tt;
      while (!accepted) {
#pragma unroll
        for (int k = 0; k < 6; k++) st[k] = z.s[k] + alpha * ds[k];
        ut[0] = z.u[0] + alpha * du[0];
        ut[1] = z.u[1] + alpha * du[1];
        point_metrics(st, ut, tt, cnt_, c0t, ft, lt, tht);
        alpha_test = alpha;
        if (ls_accept(ls, U, alpha_test, tht, ft - U.mu * lt)) { accepted = true; break; }
        if (ntrial == 0 && tht >= ls.theta) {
          // second-order correction (Ipopt max_soc = 4): same matrix, corrected constraint rhs
          int cnt = 0;
          double theta_soc_old = 0.0, theta_trial = tht, alpha_soc = alpha;
          double cs[6], cs0[6];
#pragma unroll
          for (int k = 0; k < 6; k++) { cs[k] = cn[k]; cs0[k] = c0[k]; }
          bool tried = false;
          while (cnt < K_MAX_SOC && !accepted && (cnt == 0 || theta_trial <= K_KAPPA_SOC * theta_soc_old)) {
            theta_soc_old = theta_trial;
#pragma unroll
            for (int k = 0; k < 6; k++) { cs[k] = alpha_soc * cs[k] + cnt_[k]; cs0[k] = alpha_soc * cs0[k] + c0t[k]; }
            sync();
            store_d(cs);
            sync();
            riccati(dw_used, PC[PC_CW]);
            double t0[6];
#pragma unroll
            for (int k = 0; k < 6; k++) t0[k] = -cs0[k];
            forward_and_costate(t0, dw_used, PC[PC_CW]);
            tried = true;
            alpha_soc = frac_to_bound(U.tau);
#pragma unroll
            for (int k = 0; k < 6; k++) st[k] = z.s[k] + alpha_soc * ds[k];
            ut[0] = z.u[0] + alpha_soc * du[0];
            ut[1] = z.u[1] + alpha_soc * du[1];
            point_metrics(st, ut, tt, cnt_, c0t, ft, lt, theta_trial);
            if (ls_accept(ls, U, alpha_test, theta_trial, ft - U.mu * lt)) {
              accepted = true;
              alpha = alpha_soc;
              tht = theta_trial;
              dual_steps(U.mu, dzl, dzu);
              alpha_z = frac_to_bound_dual(U.tau, dzl, dzu);
            } else {
              cnt++;
            }
          }
          if (accepted) break;
          if (tried) {
            // restore the uncorrected direction for the backtracking steps
            sync();
            store_d(cn);
            sync();
            riccati(dw_used, PC[PC_CW]);
            double t0[6];
#pragma unroll
            for (int k = 0; k < 6; k++) t0[k] = -c0[k];
            forward_and_costate(t0, dw_used, PC[PC_CW]);
          }
        }
        alpha *= 0.5;
        ntrial++;
        if (alpha < alpha_min) break;
      }
      if (!accepted) { status = 9; break; }   // Ipopt would enter restoration here

      if (!is_ftype(ls, alpha_test) || !armijo(ls, alpha_test, ft - U.mu * lt))
        filter_add((1.0 - K_GAMMA_THETA) * ls.theta, ls.phi - K_GAMMA_PHI * ls.theta);

      // ---- accept the trial point
#pragma unroll
      for (int k = 0; k < 6; k++) {
        z.s[k] = st[k];
        z.lam[k] += alpha * (lnew[k] - z.lam[k]);
        cn[k] = cnt_[k];
        c0[k] = c0t[k];
      }
      z.u[0] = ut[0];
      z.u[1] = ut[1];
      tg = tt;
      fx = ft; lsum = lt; theta = tht;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (valid4(k)) {
          const double xv = var4(z.s, z.u, k);
          const double sl = xv - PC[PC_LO + k], su = PC[PC_HI + k] - xv;
          double a = z.zl[k] + alpha_z * dzl[k];
          z.zl[k] = fmax(fmin(a, K_KAPPA_SIGMA * U.mu / sl), U.mu / (K_KAPPA_SIGMA * sl));
          a = z.zu[k] + alpha_z * dzu[k];
          z.zu[k] = fmax(fmin(a, K_KAPPA_SIGMA * U.mu / su), U.mu / (K_KAPPA_SIGMA * su));
        }
      }
      iter++;
      sync();
    }

    // ---- finalize: honor_original_bounds, unscaled objective, outputs of MPC.cpp:306-324
    {
      double v;
      v = z.s[2]; v = fmax(v, PC[PC_LO0]); v = fmin(v, PC[PC_HI0]); z.s[2] = v;
      v = z.s[3]; v = fmax(v, PC[PC_LO0 + 1]); v = fmin(v, PC[PC_HI0 + 1]); z.s[3] = v;
      v = z.u[0]; v = fmax(v, PC[PC_LO0 + 2]); v = fmin(v, PC[PC_HI0 + 2]); z.u[0] = v;
      v = z.u[1]; v = fmax(v, PC[PC_LO0 + 3]); v = fmin(v, PC[PC_HI0 + 3]); z.u[1] = v;
    }
    double fl = 0.0;
    if (act) {
      const double dv = z.s[3] - vref();
      fl = 0.5 * (wc2() * z.s[4] * z.s[4] + we2() * z.s[5] * z.s[5] + PC[PC_WV2] * dv * dv + nv2() * z.s[3] * z.s[3]);
    }
    const double dprev = up(z.u[0]);
    if (hasu) {
      fl += 0.5 * PC[PC_WD2] * z.u[0] * z.u[0];
      if (g >= 1) { const double dd = z.u[0] - dprev; fl += 0.5 * PC[PC_CW] * dd * dd; }
    }
    const double cost = gsum<G>(fl, gm) / sf;
    if (g == 1) {
#pragma unroll
      for (int k = 0; k < 6; k++) P.result[(size_t)k * B + b] = z.s[k];
    }
    if (g == 0) {
      P.result[(size_t)6 * B + b] = z.u[0];
      P.result[(size_t)7 * B + b] = z.u[1];
      P.result[(size_t)8 * B + b] = cost;
      if (P.status) P.status[b] = status;
      if (P.iters) P.iters[b] = iter;
    }
    if (act) {
      if (P.traj_x) P.traj_x[(size_t)g * B + b] = z.s[0];
      if (P.traj_y) P.traj_y[(size_t)g * B + b] = z.s[1];
      if (P.full) {
        const int Nf = P.Nmax;
#pragma unroll
        for (int k = 0; k < 6; k++) P.full[(size_t)(k * Nf + g) * B + b] = z.s[k];
        if (hasu) {
          P.full[(size_t)(6 * Nf + g) * B + b] = z.u[0];
          P.full[(size_t)(7 * Nf - 1 + g) * B + b] = z.u[1];
        }
      }
    }
    sync();
  }
};

// Persistent grid: each group of G lanes pulls the next problem index from a global counter, so
// the spread of interior-point iteration counts (10 typical, 30+ when a yaw bound is nearly
// active) does not leave lanes idle behind a static problem->warp map.
template <int G>
__global__ void __launch_bounds__(128) mpc_ipm_kernel(const KParams P) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  const int g = lane % G;
  const unsigned gm = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - g));
  double *W = smem + (size_t)(threadIdx.x / G) * P.ws_stride;
  Solver<G> S(P, W, g, gm);
  for (;;) {
    int b = 0;
    if (g == 0) b = atomicAdd(P.counter, 1);
    b = __shfl_sync(gm, b, 0, G);
    if (b >= P.B) break;
    S.solve(b);
  }
}

}  // namespace mpcb200
