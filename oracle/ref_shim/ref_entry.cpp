// ref_entry.cpp -- C entry points around the reference's OWN, UNMODIFIED controller sources
// (test infrastructure).  Compiled together with
//   /root/reference/src/control/MPC.cpp, src/model/{Vehicle,RoadGeometry}.cpp, src/utils/{utils,Config}.cpp
// against the CppAD / Ipopt stand-ins in include/ (see include/cppad/ipopt/solve.hpp).  What runs here
// is the reference's MPC::run / MPC::solve / FG_eval / Config::load text; only CppAD's tape and Ipopt
// are stand-ins.  Output: oracle/_ref/libmpc_ref.so (never committed).
#include <cstring>
#include <string>
#include <vector>
#include <cppad/cppad.hpp>
#include <cppad/ipopt/solve.hpp>
#include "Eigen/Core"

// Test access to MPC::roadGeometry (private by default in `class MPC`) so that solve() can be called on
// given coefficients, the way src/test.cpp:85 calls it after run() has set them.  Everything MPC.h
// includes is included first, so the macro only touches the `class MPC {` line; the reference
// sources themselves are compiled untouched (the class key is not part of the ABI).
#include "model/RoadGeometry.h"
#include "model/Vehicle.h"
#define class struct
#include "control/MPC.h"
#undef class
#include "utils/Config.h"

extern "C" {

// Config::load(fileName), src/utils/Config.cpp:31-87
int ref_config_load(const char *path) {
  try { Config::load(std::string(path)); } catch (...) { return -1; }
  return 0;
}

// the statics the hot path reads, for checking the product's own loader (include/mpc_b200.h mpc_config)
void ref_config_get(double *scal /*[16]*/, double *weights /*[12]*/, double *steers /*[16]*/, double *steer_speeds /*[16]*/,
                    double *yaw_changes /*[16]*/, double *yaw_change_speeds /*[16]*/, int *counts /*[4]*/) {
  scal[0] = (double)Config::N; scal[1] = Config::dt; scal[2] = Config::Lf; scal[3] = Config::ctePanic;
  scal[4] = Config::epsiPanic; scal[5] = Config::maxSpeed; scal[6] = Config::maxSteering;
  scal[7] = Config::maxAcceleration; scal[8] = Config::maxDeceleration; scal[9] = (double)Config::maxFitOrder;
  scal[10] = Config::maxFitError; scal[11] = (double)Config::latency; scal[12] = Config::lookahead;
  scal[13] = Config::ipoptTimeout; scal[14] = Config::steerAdjustmentThresh; scal[15] = Config::steerAdjustmentRatio;
  for (size_t i = 0; i < 12 && i < Config::weights.size(); i++) weights[i] = Config::weights[i];
  counts[0] = (int)Config::steers.size(); counts[1] = (int)Config::steerSpeeds.size();
  counts[2] = (int)Config::yawChanges.size(); counts[3] = (int)Config::yawChangeSpeeds.size();
  for (int i = 0; i < counts[0] && i < 16; i++) steers[i] = Config::steers[i];
  for (int i = 0; i < counts[1] && i < 16; i++) steer_speeds[i] = Config::steerSpeeds[i];
  for (int i = 0; i < counts[2] && i < 16; i++) yaw_changes[i] = Config::yawChanges[i];
  for (int i = 0; i < counts[3] && i < 16; i++) yaw_change_speeds[i] = Config::yawChangeSpeeds[i];
}
void ref_config_set_weights(const double *w, int n) { Config::weights.assign(w, w + n); }
void ref_config_set_horizon(int N, double dt) { Config::N = (size_t)N; Config::dt = dt; }

// MPC::solve (src/control/MPC.cpp:183-325) on explicit coefficients / yaw bounds
int ref_solve(const double *state, const double *coeffs, int ncoef, double yaw_lo, double yaw_hi,
              double *result9, double *traj_x, double *traj_y, int *status, int *iters) {
  MPC mpc;
  Eigen::VectorXd poly(ncoef);
  for (int i = 0; i < ncoef; i++) poly[i] = coeffs[i];
  mpc.roadGeometry.getPolynomial() = poly;
  Config::yawLow = yaw_lo;
  Config::yawHigh = yaw_hi;
  Eigen::VectorXd st(6);
  for (int i = 0; i < 6; i++) st[i] = state[i];
  std::vector<double> tx, ty;
  std::vector<double> r = mpc.solve(st, 40, &tx, &ty);
  for (int i = 0; i < 9; i++) result9[i] = r[i];
  for (size_t i = 0; i < tx.size(); i++) { if (traj_x) traj_x[i] = tx[i]; if (traj_y) traj_y[i] = ty[i]; }
  if (status) *status = CppAD::ipopt::last_stats().status;
  if (iters) *iters = CppAD::ipopt::last_stats().iters;
  return 0;
}

// MPC::run (src/control/MPC.cpp:327-382) from a pose and waypoints, as src/test.cpp:64-67 and
// src/mpc_main.cpp:155-169 call it.  ptsx/ptsy are transformed in place like the reference does.
int ref_run(double px, double py, double psi, double v, double steering, double accel, double *ptsx, double *ptsy,
            int npts, double *result8, double *traj_x, double *traj_y, double *coeffs5, int *ncoef, double *yaw_lo,
            double *yaw_hi, int *status, int *iters) {
  MPC mpc;
  Vehicle vehicle;
  vehicle.setLength(Config::Lf);
  vehicle.update(px, py, psi, v, steering, accel);
  std::vector<double> x(ptsx, ptsx + npts), y(ptsy, ptsy + npts), tx, ty;
  std::vector<double> r = mpc.run(vehicle, x, y, &tx, &ty);
  for (int i = 0; i < 8; i++) result8[i] = r[i];
  for (int i = 0; i < npts; i++) { ptsx[i] = x[i]; ptsy[i] = y[i]; }
  for (size_t i = 0; i < tx.size(); i++) { if (traj_x) traj_x[i] = tx[i]; if (traj_y) traj_y[i] = ty[i]; }
  Eigen::VectorXd &poly = mpc.roadGeometry.getPolynomial();
  *ncoef = (int)poly.size();
  for (int i = 0; i < 5; i++) coeffs5[i] = i < poly.size() ? poly[i] : 0.0;
  *yaw_lo = Config::yawLow; *yaw_hi = Config::yawHigh;
  if (status) *status = CppAD::ipopt::last_stats().status;
  if (iters) *iters = CppAD::ipopt::last_stats().iters;
  return 0;
}

// Vehicle::move / computeThrottle (src/model/Vehicle.cpp:81-103,145-168) for the closed-loop plant
void ref_vehicle_move(double *pose6 /* x,y,psi,v,steering,accel */, double length, double dt) {
  Vehicle vh;
  vh.setLength(length);
  vh.update(pose6[0], pose6[1], pose6[2], pose6[3], pose6[4], pose6[5]);
  vh.move(dt);
  pose6[0] = vh.getX(); pose6[1] = vh.getY(); pose6[2] = vh.getOrientation(); pose6[3] = vh.getVelocity();
}
double ref_compute_throttle(double accel, double target, double max_accel, double max_decel) {
  Vehicle vh;
  return vh.computeThrottle(accel, target, max_accel, max_decel);
}
int ref_tape_size(void) { return CppAD::ipopt::last_stats().tape_size; }

}  // extern "C"
