// mpc_run_logic.h -- the arithmetic of one control step around the solve, shared by the host entry
// points (mpc_host.cpp) and the device kernels of the closed-loop rollout (mpc_rollout.cuh):
//   Vehicle::globalToVehicle            /root/reference/src/model/Vehicle.cpp:105-114
//   RoadGeometry::fit                   src/model/RoadGeometry.cpp:18-39   (adaptive order 2..maxFitOrder-1)
//   polyfit (Vandermonde + Householder QR least squares)   src/utils/utils.cpp:10-29
//   RoadGeometry::computeOrientationChange / orientation   src/model/RoadGeometry.cpp:41-61
//   Vehicle::computeYawChangeSpeedLimit / computeSpeedTarget(double)   src/model/Vehicle.cpp:34-48,66-79
//   Vehicle::computeThrottle, Vehicle::move                src/model/Vehicle.cpp:81-103,145-168
//   MPC::run pre/post-processing        src/control/MPC.cpp:327-356,361-381
// Pure functions of their arguments (the reference writes Config::yawLow/yawHigh, MPC.cpp:345-352; here they
// are outputs).
#pragma once
#include <math.h>
#include "../../include/mpc_b200.h"

#ifdef __CUDACC__
#define MPC_HD __host__ __device__ inline
#else
#define MPC_HD inline
#endif

namespace mpcrun {

MPC_HD double polyval(const double *c, int n, double x) {   // utils.h:28-34
  double r = 0.0;
  for (int i = n - 1; i >= 0; i--) r = r * x + c[i];
  return r;
}
MPC_HD double polyder(const double *c, int n, double x) {   // utils.h:41-47
  double r = 0.0;
  for (int i = n - 1; i >= 1; i--) r = r * x + i * c[i];
  return r;
}
MPC_HD double normalize_angle(double a) {                   // utils.h:86-92
  const double pi = 3.14159265358979323846;
  while (a >= pi) a -= 2. * pi;
  while (a < -pi) a += 2. * pi;
  return a;
}
MPC_HD double orientation(const double *c, int n, double px, double dir) {   // RoadGeometry.cpp:41-47
  double psi = atan(polyder(c, n, px));
  if (dir < 0) psi = normalize_angle(psi + 3.14159265358979323846);
  return psi;
}
MPC_HD double table_limit(const double *keys, int nk, const double *vals, int nv, double angle, double mx) {
  const double y = fabs(angle);
  if (nv < 1) return mx;   // no table: no limit (callers reject such configs; this keeps the read in bounds)
  for (int i = 0; i < nk; i++)
    if (y <= keys[i]) return fmin(nv > i ? vals[i] : vals[nv - 1], mx);
  return fmin(vals[nv - 1], mx);
}

// least squares min ||A c - y|| for the m x n Vandermonde matrix by Householder QR (m <= 16, n <= 5)
MPC_HD bool polyfit(const double *x, const double *y, int m, int order, double *c) {
  const int n = order + 1;
  if (m > MPC_MAX_WAYPOINTS || n > MPC_NCOEF || order < 1 || order > m - 1) return false;
  double A[MPC_MAX_WAYPOINTS][MPC_NCOEF], b[MPC_MAX_WAYPOINTS];
  for (int j = 0; j < m; j++) {
    A[j][0] = 1.0;
    for (int i = 0; i < order; i++) A[j][i + 1] = A[j][i] * x[j];
    b[j] = y[j];
  }
  for (int k = 0; k < n; k++) {
    double nrm = 0.0;
    for (int j = k; j < m; j++) nrm += A[j][k] * A[j][k];
    nrm = sqrt(nrm);
    if (nrm == 0.0) return false;
    const double alpha = A[k][k] > 0 ? -nrm : nrm;
    double v[MPC_MAX_WAYPOINTS];
    for (int j = k; j < m; j++) v[j] = A[j][k];
    v[k] -= alpha;
    double vv = 0.0;
    for (int j = k; j < m; j++) vv += v[j] * v[j];
    if (vv > 0.0) {
      for (int col = k; col < n; col++) {
        double s = 0.0;
        for (int j = k; j < m; j++) s += v[j] * A[j][col];
        s = 2.0 * s / vv;
        for (int j = k; j < m; j++) A[j][col] -= s * v[j];
      }
      double s = 0.0;
      for (int j = k; j < m; j++) s += v[j] * b[j];
      s = 2.0 * s / vv;
      for (int j = k; j < m; j++) b[j] -= s * v[j];
    }
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int j = i + 1; j < n; j++) s -= A[i][j] * c[j];
    c[i] = s / A[i][i];
  }
  return true;
}

// MPC::run up to the solve (MPC.cpp:327-356).  ptsx/ptsy: global waypoints, transformed in place.
MPC_HD int run_prepare(const mpc_config *cfg, const double *pose, double steering, double *ptsx, double *ptsy,
                       int npts, double *state, double *coeffs, double *yaw_lo, double *yaw_hi, mpc_run_aux *aux) {
  if (npts < 3 || npts > MPC_MAX_WAYPOINTS) return MPC_EINVAL;
  const double cs = cos(pose[2]), sn = sin(pose[2]);
  for (int i = 0; i < npts; i++) {       // in place, like MPC.cpp:329
    const double vx = ptsx[i] - pose[0], vy = ptsy[i] - pose[1];
    ptsx[i] = vx * cs + vy * sn;
    ptsy[i] = vy * cs - vx * sn;
  }
  // adaptive-order fit: order 2, 3, ... while the squared error is above maxFitError and order < maxFitOrder
  int order = 2, ncoef = 0;
  double c[MPC_NCOEF], err = 0.0;
  do {
    if (order > npts - 1 || order + 1 > MPC_NCOEF) break;
    if (!polyfit(ptsx, ptsy, npts, order, c)) return MPC_EINVAL;
    ncoef = order + 1;
    order++;
    err = 0.0;
    for (int i = 0; i < npts; i++) { const double d = ptsy[i] - polyval(c, ncoef, ptsx[i]); err += d * d; }
  } while (err > cfg->max_fit_error && order < cfg->max_fit_order);
  if (ncoef == 0) return MPC_EINVAL;
  for (int i = 0; i < MPC_NCOEF; i++) coeffs[i] = i < ncoef ? c[i] : 0.0;
  const double cte = polyval(c, ncoef, 0.0);                   // MPC.cpp:334
  const double epsi = -atan(c[1]);                             // MPC.cpp:336
  const double xl = ptsx[npts - 1], xf = ptsx[0];
  const double dir = xl - 0.0;
  const double myc = (orientation(c, ncoef, xl, dir) - orientation(c, ncoef, 0.0, dir)) * (xl - xf) / xl;   // MPC.cpp:339
  const double max_speed = table_limit(cfg->yaw_changes, cfg->n_yaw_changes, cfg->yaw_change_speeds,
                                       cfg->n_yaw_change_speeds, myc, cfg->max_speed);           // MPC.cpp:340
  const double target = table_limit(cfg->steers, cfg->n_steers, cfg->steer_speeds, cfg->n_steer_speeds,
                                    steering, max_speed);                                        // MPC.cpp:342
  if (myc < 0) { *yaw_lo = myc; *yaw_hi = 0.1; } else { *yaw_lo = -0.1; *yaw_hi = myc; }         // MPC.cpp:345-352
  state[0] = 0; state[1] = 0; state[2] = 0; state[3] = pose[3]; state[4] = cte; state[5] = epsi; // MPC.cpp:355-356
  aux->max_yaw_change = myc; aux->max_speed = max_speed; aux->target_speed = target; aux->fit_error = err;
  aux->fit_order = ncoef - 1;
  return MPC_OK;
}

// MPC::run after the solve (MPC.cpp:361-381)
MPC_HD void run_finish(const mpc_config *cfg, const mpc_run_aux *aux, double v, const double *r, double *out8) {
  double steer = r[6];
  if (fabs(aux->max_yaw_change) > cfg->steer_adjust_thresh) steer += cfg->steer_adjust_ratio * aux->max_yaw_change;  // MPC.cpp:364-366
  const double accel = fmin(r[7], aux->target_speed - v);                                        // MPC.cpp:369
  double sv = steer / cfg->max_steering;                                                         // MPC.cpp:371
  sv = sv < -1.0 ? -1.0 : (sv > 1.0 ? 1.0 : sv);
  out8[0] = r[0]; out8[1] = r[1]; out8[2] = r[2]; out8[3] = r[3]; out8[4] = sv; out8[5] = accel;
  out8[6] = r[4]; out8[7] = r[5];                                                                // MPC.cpp:381
}

// Vehicle::computeThrottle, Vehicle.cpp:81-103 (keep = target / Config::maxSpeed)
MPC_HD double compute_throttle(double accel, double target, double max_accel, double max_decel, double max_speed) {
  const double keep = target / max_speed;
  if (accel >= 0) {
    if (accel < 0.001) return keep;
    return fmin(1.0, keep + (1 - keep) * accel / max_accel);
  }
  if (accel <= -15) return -1;
  if (accel < -10) return -0.95 - (1 - 0.95) * accel / max_decel;
  if (accel < -5) return -0.9 - (1 - 0.9) * accel / max_decel;
  return -0.85 - (1 - 0.85) * accel / max_decel;
}

// Vehicle::move, Vehicle.cpp:145-168: position uses the OLD heading; the speed clamp there is dead code
// (the clamped member is overwritten by v + a*dt right after), so there is none here either.
MPC_HD void vehicle_move(double *x, double *y, double *psi, double *v, double steering, double accel, double length, double dt) {
  const double dist = *v * dt;
  const double dpsi = steering * dist / length;
  const double px = *x + dist * cos(*psi), py = *y + dist * sin(*psi);
  *x = px; *y = py; *psi = *psi + dpsi; *v = *v + accel * dt;
}

}  // namespace mpcrun
