// A C++ caller of include/mpc_b200.h on every GPU of the box, without Python: mpc_create_multi /
// mpc_solve_batch_multi against one handle on device 0 (mpc_solve_batch_host).  Problems: the pose of the
// reference's offline harness (src/test.cpp:45-50) moved and slowed at random, through mpc_run_prepare.
// usage: test_multi_gpu config.json B n_handles [pinned]   (handles are dealt round-robin over the visible devices; "pinned":
// the caller's arrays are page-locked with cudaHostRegister, so the copies of the devices overlap)
// prints "devices <d> handles <n> B <B> mismatches <k> ok_frac <f> ms_multi <t> ms_single <t>"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "mpc_b200.h"

int main(int argc, char **argv) {
  if (argc < 4) { fprintf(stderr, "usage\n"); return 2; }
  mpc_config cfg;
  if (mpc_config_load_json(argv[1], &cfg) != MPC_OK) { fprintf(stderr, "config load failed\n"); return 3; }
  const int B = atoi(argv[2]), nh = atoi(argv[3]), N = cfg.N;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { fprintf(stderr, "no device\n"); return 4; }
  std::vector<int> devs(nh);
  for (int g = 0; g < nh; g++) devs[g] = g % ndev;
  const double ptsx0[6] = {-134.97, -145.1165, -158.3417, -164.3164, -169.3365, -175.4917};
  const double ptsy0[6] = {18.404, 4.339378, -17.42898, -30.18062, -42.84062, -66.52898};
  std::vector<double> state(6 * (size_t)B), coeffs(5 * (size_t)B), ylo(B), yhi(B);
  unsigned long long s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; };
  for (int b = 0; b < B; b++) {
    double px[6], py[6], st[6], co[5], lo, hi;
    memcpy(px, ptsx0, sizeof(px)); memcpy(py, ptsy0, sizeof(py));
    const double pose[4] = {-146.7283 + 2.0 * (rnd() - 0.5), 1.660802 + 2.0 * (rnd() - 0.5), 4.125825 + 0.2 * (rnd() - 0.5), 5.0 + 40.0 * rnd()};
    mpc_run_aux aux;
    if (mpc_run_prepare(&cfg, pose, 0.0, px, py, 6, st, co, &lo, &hi, &aux) != MPC_OK) { fprintf(stderr, "prepare failed\n"); return 5; }
    for (int k = 0; k < 6; k++) state[(size_t)k * B + b] = st[k];
    for (int k = 0; k < 5; k++) coeffs[(size_t)k * B + b] = co[k];
    ylo[b] = lo; yhi[b] = hi;
  }
  auto outputs = [&](std::vector<double> &res, std::vector<double> &tx, std::vector<double> &ty, std::vector<int> &stt, std::vector<int> &it) {
    res.assign(9 * (size_t)B, 0.0); tx.assign((size_t)N * B, 0.0); ty.assign((size_t)N * B, 0.0); stt.assign(B, 0); it.assign(B, 0);
  };
  std::vector<double> r1, x1, y1, r2, x2, y2;
  std::vector<int> s1, i1, s2, i2;
  outputs(r1, x1, y1, s1, i1); outputs(r2, x2, y2, s2, i2);
  if (argc > 4 && !strcmp(argv[4], "pinned")) {
    auto pin = [](void *p, size_t bytes) { if (cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) { fprintf(stderr, "cudaHostRegister failed\n"); exit(10); } };
    pin(state.data(), state.size() * 8); pin(coeffs.data(), coeffs.size() * 8); pin(ylo.data(), ylo.size() * 8); pin(yhi.data(), yhi.size() * 8);
    pin(r1.data(), r1.size() * 8); pin(x1.data(), x1.size() * 8); pin(y1.data(), y1.size() * 8); pin(s1.data(), s1.size() * 4); pin(i1.data(), i1.size() * 4);
    pin(r2.data(), r2.size() * 8); pin(x2.data(), x2.size() * 8); pin(y2.data(), y2.size() * 8); pin(s2.data(), s2.size() * 4); pin(i2.data(), i2.size() * 4);
  }
  mpc_handle *h = nullptr;
  mpc_multi *m = nullptr;
  if (mpc_create(&cfg, 0, &h) != MPC_OK) { fprintf(stderr, "create failed: %s\n", mpc_last_error()); return 6; }
  if (mpc_create_multi(&cfg, devs.data(), nh, &m) != MPC_OK) { fprintf(stderr, "create_multi failed: %s\n", mpc_last_error()); return 7; }
  double ms1 = 0, ms2 = 0;
  for (int rep = 0; rep < 3; rep++) {
    auto t0 = std::chrono::steady_clock::now();
    int rc = mpc_solve_batch_host(h, B, state.data(), coeffs.data(), ylo.data(), yhi.data(), 0, 0, 0, r1.data(), x1.data(), y1.data(), 0, s1.data(), i1.data());
    auto t1 = std::chrono::steady_clock::now();
    if (rc) { fprintf(stderr, "solve failed %d: %s\n", rc, mpc_last_error()); return 8; }
    rc = mpc_solve_batch_multi(m, B, state.data(), coeffs.data(), ylo.data(), yhi.data(), 0, 0, 0, r2.data(), x2.data(), y2.data(), 0, s2.data(), i2.data());
    auto t2 = std::chrono::steady_clock::now();
    if (rc) { fprintf(stderr, "solve_multi failed %d: %s\n", rc, mpc_last_error()); return 9; }
    ms1 = std::chrono::duration<double, std::milli>(t1 - t0).count();
    ms2 = std::chrono::duration<double, std::milli>(t2 - t1).count();
  }
  long long bad = 0, ok = 0;
  for (size_t k = 0; k < r1.size(); k++) bad += memcmp(&r1[k], &r2[k], sizeof(double)) != 0;
  for (size_t k = 0; k < x1.size(); k++) bad += (memcmp(&x1[k], &x2[k], 8) != 0) + (memcmp(&y1[k], &y2[k], 8) != 0);
  for (int b = 0; b < B; b++) { bad += (s1[b] != s2[b]) + (i1[b] != i2[b]); ok += s2[b] == MPC_STATUS_SUCCESS; }
  printf("devices %d handles %d B %d mismatches %lld ok_frac %.4f ms_multi %.3f ms_single %.3f\n", ndev, mpc_multi_device_count(m), B, bad,
         (double)ok / B, ms2, ms1);
  mpc_destroy_multi(m);
  mpc_destroy(h);
  return bad == 0 ? 0 : 1;
}
