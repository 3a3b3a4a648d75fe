// Drop-in replacement for the reference's src/control/MPC.cpp: same class (src/control/MPC.h is kept
// as it is), the solve goes to the B200 library through the C-ABI instead of CppAD + Ipopt + MUMPS.
// Build it in the reference tree INSTEAD of the original file and link libmpc_b200.so; the rest of
// the tree (mpc_main.cpp, test.cpp, Vehicle, RoadGeometry, Config) is untouched.
//
// Replaces: FG_eval (MPC.cpp:15-154), MPC::MPC (:160-179), MPC::solve (:183-325), MPC::run (:327-382).
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include "MPC.h"
#include "../utils/Config.h"
#include "mpc_b200.hpp"

namespace {
// Config:: statics (already converted by Config::load, Config.cpp:39-86) -> the library's POD
mpc_config config_from_statics() {
  mpc_config c;
  mpc_config_defaults(&c);
  c.N = (int)Config::N; c.dt = Config::dt; c.Lf = Config::Lf;
  c.cte_panic = Config::ctePanic; c.epsi_panic = Config::epsiPanic;
  c.max_speed = Config::maxSpeed; c.max_steering = Config::maxSteering;
  c.max_accel = Config::maxAcceleration; c.max_decel = Config::maxDeceleration;
  c.max_fit_order = Config::maxFitOrder; c.max_fit_error = Config::maxFitError;
  c.latency_ms = (int)Config::latency; c.lookahead = Config::lookahead; c.ipopt_timeout = Config::ipoptTimeout;
  c.steer_adjust_thresh = Config::steerAdjustmentThresh; c.steer_adjust_ratio = Config::steerAdjustmentRatio;
  for (size_t i = 0; i < MPC_NWEIGHTS && i < Config::weights.size(); i++) c.weights[i] = Config::weights[i];
  auto tab = [](const std::vector<double> &v, double *out, int *n) {
    *n = (int)(v.size() < MPC_NTAB ? v.size() : MPC_NTAB);
    for (int i = 0; i < *n; i++) out[i] = v[i];
  };
  tab(Config::steers, c.steers, &c.n_steers);
  tab(Config::steerSpeeds, c.steer_speeds, &c.n_steer_speeds);
  tab(Config::yawChanges, c.yaw_changes, &c.n_yaw_changes);
  tab(Config::yawChangeSpeeds, c.yaw_change_speeds, &c.n_yaw_change_speeds);
  return c;
}
// One solver object (device workspace, stream, pinned staging block) for the life of the process.  The reference
// builds a new MPC -- and with it a new tape and a new Ipopt problem -- for every telemetry message
// (mpc_main.cpp:99) and reads the Config statics afresh on every call; here the handle is kept and only told about
// a configuration when the statics have changed since the last call (Config::load, or the -speed / -latency
// overrides of mpc_main.cpp:244-246).  Creating and destroying a handle per call costs four cudaMalloc, a
// cudaMallocHost and a stream, far more than the ~200 us solve.
mpcb200::MPC &solver_for_statics() {
  static std::unique_ptr<mpcb200::MPC> impl;
  static mpc_config last;
  const mpc_config c = config_from_statics();   // starts from mpc_config_defaults, which zeroes the padding too
  if (!impl) {
    impl.reset(new mpcb200::MPC(c, 0));
    last = c;
  } else if (std::memcmp(&last, &c, sizeof(c)) != 0) {
    impl->setConfig(c);
    last = c;
  }
  return *impl;
}
}  // namespace

MPC::MPC() {}
MPC::~MPC() {}

vector<double> MPC::solve(VectorXd &state, double target_velocity, vector<double> *x_trajectory,
                          vector<double> *y_trajectory, double dir) {
  mpcb200::MPC &impl = solver_for_statics();
  VectorXd &poly = roadGeometry.getPolynomial();
  impl.setRoad(poly.data(), (int)poly.size(), Config::yawLow, Config::yawHigh);
  return impl.solve(state, target_velocity, x_trajectory, y_trajectory, dir);
}

vector<double> MPC::run(Vehicle &vehicle, vector<double> &ptsx, vector<double> &ptsy,
                        std::vector<double> *x_trajectory, std::vector<double> *y_trajectory) {
  mpcb200::MPC &impl = solver_for_statics();
  mpcb200::VehiclePose pose = {vehicle.getX(), vehicle.getY(), vehicle.getOrientation(), vehicle.getVelocity(),
                               vehicle.getSteering(), vehicle.getAcceleration()};
  vector<double> out = impl.run(pose, ptsx, ptsy, x_trajectory, y_trajectory);
  // leave behind what the reference's run() leaves behind, for a following solve() (src/test.cpp:85)
  this->vehicle = vehicle;
  VectorXd poly(impl.lastRun().fit_order + 1);
  for (int i = 0; i < poly.size(); i++) poly[i] = impl.polynomial()[i];
  roadGeometry.getPolynomial() = poly;
  Config::yawLow = impl.yawLow();
  Config::yawHigh = impl.yawHigh();
  return out;
}
