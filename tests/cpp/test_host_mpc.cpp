// The scenario of the reference's offline harness (src/test.cpp:45-111): run() once from the fixture
// pose, then 25 x solve() feeding the step-1 state back -- through the C++ host class over the C-ABI.
// usage: test_host_mpc config.json x y psi v  px0 py0 ... (6 waypoints)   -> one line of numbers per call
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "mpc_b200.hpp"

int main(int argc, char **argv) {
  if (argc < 6 + 12) { fprintf(stderr, "usage\n"); return 2; }
  mpc_config cfg;
  if (mpc_config_load_json(argv[1], &cfg) != MPC_OK) { fprintf(stderr, "config load failed\n"); return 3; }
  mpcb200::VehiclePose v = {atof(argv[2]), atof(argv[3]), atof(argv[4]), atof(argv[5]), 0.0, 0.0};
  std::vector<double> px, py;
  for (int i = 0; i < 6; i++) { px.push_back(atof(argv[6 + 2 * i])); py.push_back(atof(argv[7 + 2 * i])); }
  try {
    mpcb200::MPC mpc(cfg, 0);
    std::vector<double> tx, ty;
    std::vector<double> r = mpc.run(v, px, py, &tx, &ty);
    printf("run");
    for (double x : r) printf(" %.17g", x);
    printf(" | %d %d %zu\n", mpc.lastStatus(), mpc.lastIters(), tx.size());
    std::vector<double> state = {r[0], r[1], r[2], r[3], r[6], r[7]};
    for (int k = 0; k < 25; k++) {
      std::vector<double> s = mpc.solve(state, 40);
      printf("solve");
      for (double x : s) printf(" %.17g", x);
      printf(" | %d %d\n", mpc.lastStatus(), mpc.lastIters());
      for (int i = 0; i < 6; i++) state[i] = s[i];
    }
  } catch (const std::exception &e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
