// mpc_host.cpp -- host entry points of the control-step logic around the solve (MPC::run,
// /root/reference/src/control/MPC.cpp:327-382); the arithmetic lives in mpc_run_logic.h so that the
// device kernels of the closed-loop rollout share it.
#include "mpc_run_logic.h"

extern "C" int mpc_run_prepare(const mpc_config *cfg, const double *pose, double steering, double *ptsx,
                               double *ptsy, int npts, double *state, double *coeffs, double *yaw_lo,
                               double *yaw_hi, mpc_run_aux *aux) {
  if (!cfg || !pose || !ptsx || !ptsy || !state || !coeffs || !yaw_lo || !yaw_hi || !aux) return MPC_EINVAL;
  // an empty speed-limit table is undefined behaviour in the reference (.back() of an empty vector, Vehicle.cpp:66-79)
  if (cfg->n_yaw_changes < 0 || cfg->n_yaw_changes > MPC_NTAB || cfg->n_yaw_change_speeds < 1 ||
      cfg->n_yaw_change_speeds > MPC_NTAB || cfg->n_steers < 0 || cfg->n_steers > MPC_NTAB || cfg->n_steer_speeds < 1 ||
      cfg->n_steer_speeds > MPC_NTAB || cfg->max_fit_order < 3 || cfg->max_fit_order > MPC_NCOEF)
    return MPC_EINVAL;
  return mpcrun::run_prepare(cfg, pose, steering, ptsx, ptsy, npts, state, coeffs, yaw_lo, yaw_hi, aux);
}

extern "C" int mpc_run_finish(const mpc_config *cfg, const mpc_run_aux *aux, double v, const double *r,
                              double *out8) {
  if (!cfg || !aux || !r || !out8) return MPC_EINVAL;
  mpcrun::run_finish(cfg, aux, v, r, out8);
  return MPC_OK;
}

extern "C" double mpc_compute_throttle(const mpc_config *cfg, double accel, double target) {
  return mpcrun::compute_throttle(accel, target, cfg->max_accel, cfg->max_decel, cfg->max_speed);
}

extern "C" void mpc_vehicle_move(double *pose4, double steering, double accel, double length, double dt) {
  mpcrun::vehicle_move(&pose4[0], &pose4[1], &pose4[2], &pose4[3], steering, accel, length, dt);
}
