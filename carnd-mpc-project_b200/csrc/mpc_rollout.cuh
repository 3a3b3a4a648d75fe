// mpc_rollout.cuh -- kernels either side of the solve for whole control steps on the device:
//  * mpc_run_pre_kernel / mpc_run_post_kernel: MPC::run (/root/reference/src/control/MPC.cpp:327-382) for a
//    batch of vehicles -- frame transform, adaptive-order Householder-QR fit, cte/epsi, yaw bounds and
//    speed tables in front of the solve; steering adjustment, acceleration clamp and normalisation behind it;
//  * mpc_loop_pre_kernel / mpc_loop_post_kernel: one step of the closed loop of src/mpc_main.cpp:113-214 with
//    the simulator replaced by the reference's own kinematic plant (Vehicle::move, Vehicle.cpp:145-168):
//    waypoint window, latency compensation, run(), throttle map, actuation delay, plant step.
// One thread per vehicle: a 6-point fit of order <= 4 is a few hundred flops.
#pragma once
#include <cuda_runtime.h>
#include "mpc_lane_kernel.cuh"
#include "mpc_run_logic.h"

namespace mpcb200 {

struct RunBatch {
  int B, npts;
  const double *pose;       // [4][B] x, y, psi, v (global frame)
  const double *steering;   // [B] or NULL (0)
  const double *ptsx, *ptsy;   // [npts][B] global waypoints
  double *ptsx_v, *ptsy_v;     // [npts][B] or NULL: vehicle-frame waypoints (MPC.cpp:329 transforms in place)
  double *state, *coeffs, *yaw_lo, *yaw_hi;   // NLP inputs, [6][B], [5][B], [B], [B]
  double *aux;              // [4][B]: max_yaw_change, target_speed, v, fit_order
};

__global__ void __launch_bounds__(128) mpc_run_pre_kernel(const mpc_config cfg, const RunBatch R) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= R.B) return;
  const size_t B = (size_t)R.B;
  double pose[4], px[MPC_MAX_WAYPOINTS], py[MPC_MAX_WAYPOINTS];
  for (int k = 0; k < 4; k++) pose[k] = R.pose[k * B + b];
  for (int i = 0; i < R.npts; i++) { px[i] = R.ptsx[i * B + b]; py[i] = R.ptsy[i * B + b]; }
  double st[6] = {0, 0, 0, 0, 0, 0}, co[MPC_NCOEF] = {0, 0, 0, 0, 0}, lo = 0.0, hi = 0.0;
  mpc_run_aux aux = {0, 0, 0, 0, 0};
  const int rc = mpcrun::run_prepare(&cfg, pose, R.steering ? R.steering[b] : 0.0, px, py, R.npts, st, co, &lo, &hi, &aux);
  if (rc != MPC_OK) { lo = 0.0; hi = -1.0; }   // empty yaw interval: the solve reports a non-success status
  for (int k = 0; k < 6; k++) R.state[k * B + b] = st[k];
  for (int k = 0; k < MPC_NCOEF; k++) R.coeffs[k * B + b] = co[k];
  R.yaw_lo[b] = lo; R.yaw_hi[b] = hi;
  R.aux[0 * B + b] = aux.max_yaw_change; R.aux[1 * B + b] = aux.target_speed; R.aux[2 * B + b] = pose[3];
  R.aux[3 * B + b] = (double)aux.fit_order;
  if (R.ptsx_v)
    for (int i = 0; i < R.npts; i++) { R.ptsx_v[i * B + b] = px[i]; R.ptsy_v[i * B + b] = py[i]; }
}

__global__ void __launch_bounds__(128) mpc_run_post_kernel(const mpc_config cfg, int Bn, const double *aux,
                                                           const double *result, double *out8) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= Bn) return;
  const size_t B = (size_t)Bn;
  mpc_run_aux a = {aux[0 * B + b], 0.0, aux[1 * B + b], 0.0, 0};
  double r[9], o[8];
  for (int k = 0; k < 9; k++) r[k] = result[k * B + b];
  mpcrun::run_finish(&cfg, &a, aux[2 * B + b], r, o);
  for (int k = 0; k < 8; k++) out8[k * B + b] = o[k];
}

// ---- closed loop -------------------------------------------------------------------------------------
struct LoopArgs {
  int V, n_track, step;
  const double *track_x, *track_y;   // [n_track] closed centre line (lake_track_waypoints.csv)
  double *veh;       // [6][V] x, y, psi, v, steering angle (rad, as telemetry would report it), last computed throttle
  int *seg;          // [V] first waypoint of the current 6-point window
  double *pending;   // [2][V] command in flight (delta, throttle) when the actuators lag one step
  double dt_ctrl, tau_solve;
  double *state, *coeffs, *yaw_lo, *yaw_hi, *aux, *result;
  const int *status, *iters;
  double *rec;       // [T][8][V] or NULL: cte, epsi, v, steer in [-1,1], throttle, cost, status, iterations
};

// telemetry -> NLP inputs (mpc_main.cpp:113-169), vehicle b
__device__ __forceinline__ void loop_pre(const mpc_config &cfg, const LoopArgs &A, int b) {
  const size_t V = (size_t)A.V;
  double x = A.veh[0 * V + b], y = A.veh[1 * V + b], psi = A.veh[2 * V + b], v = A.veh[3 * V + b];
  const double steer = A.veh[4 * V + b], thr = A.veh[5 * V + b];
  // the simulator sends the next waypoints starting at the last one behind the car: advance the window
  // while its second point is already behind (vehicle-frame x <= 0)
  int seg = A.seg[b];
  const double cs = cos(psi), sn = sin(psi);
  for (int guard = 0; guard < 8; guard++) {
    const int j = (seg + 1) % A.n_track;
    if ((A.track_x[j] - x) * cs + (A.track_y[j] - y) * sn > 0.0) break;
    seg = j;
  }
  A.seg[b] = seg;
  double px[6], py[6];
  for (int i = 0; i < 6; i++) { const int j = (seg + i) % A.n_track; px[i] = A.track_x[j]; py[i] = A.track_y[j]; }
  psi = mpcrun::normalize_angle(psi);                                  // mpc_main.cpp:127
  const double accel_est = (thr - v / 50.0) * 6;                      // mpc_main.cpp:156
  if (cfg.latency_ms)                                                  // mpc_main.cpp:157-159, fixed solve-time estimate
    mpcrun::vehicle_move(&x, &y, &psi, &v, steer, accel_est, cfg.Lf, cfg.lookahead + A.tau_solve);
  const double pose[4] = {x, y, psi, v};
  double st[6] = {0, 0, 0, 0, 0, 0}, co[MPC_NCOEF] = {0, 0, 0, 0, 0}, lo = 0.0, hi = 0.0;
  mpc_run_aux aux = {0, 0, 0, 0, 0};
  const int rc = mpcrun::run_prepare(&cfg, pose, steer, px, py, 6, st, co, &lo, &hi, &aux);
  if (rc != MPC_OK) { lo = 0.0; hi = -1.0; }
  for (int k = 0; k < 6; k++) A.state[k * V + b] = st[k];
  for (int k = 0; k < MPC_NCOEF; k++) A.coeffs[k * V + b] = co[k];
  A.yaw_lo[b] = lo; A.yaw_hi[b] = hi;
  A.aux[0 * V + b] = aux.max_yaw_change; A.aux[1 * V + b] = aux.target_speed; A.aux[2 * V + b] = v;
  A.aux[3 * V + b] = (double)aux.fit_order;
}
__global__ void __launch_bounds__(128) mpc_loop_pre_kernel(const mpc_config cfg, const LoopArgs A) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < A.V) loop_pre(cfg, A, b);
}

// solve result -> actuators -> plant step (mpc_main.cpp:171-214; Vehicle::move as the simulator), vehicle b, step k
__device__ __forceinline__ void loop_post(const mpc_config &cfg, const LoopArgs &A, int b, int step) {
  const size_t V = (size_t)A.V;
  mpc_run_aux a = {A.aux[0 * V + b], 0.0, A.aux[1 * V + b], 0.0, 0};
  double r[9], o[8];
  for (int k = 0; k < 9; k++) r[k] = A.result[k * V + b];
  mpcrun::run_finish(&cfg, &a, A.aux[2 * V + b], r, o);
  const double accel = o[5];                                                             // mpc_main.cpp:171
  const double throttle = mpcrun::compute_throttle(accel, o[3], cfg.max_accel, cfg.max_decel, cfg.max_speed);   // :174
  const double delta_cmd = o[4] * cfg.max_steering;   // steering angle in the controller's sign convention
  double d_apply = delta_cmd, t_apply = throttle;
  if (cfg.latency_ms) {   // the command reaches the actuators one control interval later (mpc_main.cpp:210-214)
    d_apply = A.pending[0 * V + b]; t_apply = A.pending[1 * V + b];
    A.pending[0 * V + b] = delta_cmd; A.pending[1 * V + b] = throttle;
  }
  double x = A.veh[0 * V + b], y = A.veh[1 * V + b], psi = A.veh[2 * V + b], v = A.veh[3 * V + b];
  const double a_plant = (t_apply - v / 50.0) * 6;   // the throttle -> acceleration map the controller itself assumes
  mpcrun::vehicle_move(&x, &y, &psi, &v, d_apply, a_plant, cfg.Lf, A.dt_ctrl);
  A.veh[0 * V + b] = x; A.veh[1 * V + b] = y; A.veh[2 * V + b] = psi; A.veh[3 * V + b] = v;
  A.veh[4 * V + b] = d_apply; A.veh[5 * V + b] = throttle;
  if (A.rec) {
    double *q = A.rec + (size_t)step * 8 * V + b;
    q[0 * V] = A.state[4 * V + b]; q[1 * V] = A.state[5 * V + b]; q[2 * V] = A.aux[2 * V + b]; q[3 * V] = o[4];
    q[4 * V] = throttle; q[5 * V] = r[8]; q[6 * V] = (double)A.status[b]; q[7 * V] = (double)A.iters[b];
  }
}
__global__ void __launch_bounds__(128) mpc_loop_post_kernel(const mpc_config cfg, const LoopArgs A) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < A.V) loop_post(cfg, A, b, A.step);
}

// The whole closed loop in ONE launch: a lane group owns a vehicle for all T control steps -- message
// handling and MPC::run's pre-processing by the group's first lane, the solve by the group exactly as in
// mpc_coop_kernel (same Lane code, so the same bits as the launch-per-step path), actuation and plant step by the
// first lane again -- then takes the next vehicle.  Vehicles never wait for one another: a control step costs a
// vehicle its own ~10 trips, not the slowest vehicle's, and there are no launch seams.  The per-vehicle problem
// slots (state, coeffs, yaw bounds, result) are the same global arrays the launch-per-step path uses.
template <int NS>
__global__ void __launch_bounds__(128, 2) mpc_rollout_kernel(const KParams P, const mpc_config cfg, const LoopArgs A, int T) {
  extern __shared__ double coop_smem[];
  const int G = Lane<NS, true>::NS_GROUP;
  const int lane = threadIdx.x & 31;
  Lane<NS, true> Z;
  Z.g0 = lane % G; Z.gstep = G;
  Z.gm = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - Z.g0));
  Z.ST = reinterpret_cast<double (*)[ST_ROW_SH]>(coop_smem + (size_t)(threadIdx.x / G) * NS * ST_ROW_SH);
  Z.lh_stale = false; Z.no_handoff = true;
  const int group = (int)((blockIdx.x * blockDim.x + threadIdx.x) / G), n_groups = (int)(gridDim.x * blockDim.x / G);
  Z.scratch = P.scratch + (size_t)group * P.scratch_stride;
  for (int b = group; b < A.V; b += n_groups) {
#pragma unroll 1
    for (int k = 0; k < T; k++) {
      if (Z.g0 == 0) loop_pre(cfg, A, b);
      Z.gsync();
      Z.init(P, b);
      Z.run_to_completion(P);
      if (Z.g0 == 0) {
        Z.write_outputs(P);
        __threadfence_block();
        loop_post(cfg, A, b, k);
      }
      Z.gsync();
    }
  }
}

}  // namespace mpcb200
