// Driver for the drop-in check: the calls of the reference's src/test.cpp:13-111 (Config::load,
// Vehicle::update, MPC::run, 25 x MPC::solve) written against the REFERENCE'S headers, linked with
// integration/reference_tree/src/control/MPC.cpp instead of the reference's MPC.cpp.  (test.cpp itself
// cannot be built here: it includes matplotlibcpp.h, which needs python2.7.)
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "control/MPC.h"
#include "utils/Config.h"

int main(int argc, char **argv) {
  if (argc < 6 + 12) return 2;
  MPC mpc;
  Config::load(argv[1]);
  Vehicle vehicle;
  vehicle.setLength(Config::Lf);
  vehicle.update(atof(argv[2]), atof(argv[3]), atof(argv[4]), atof(argv[5]), 0, 0);
  std::vector<double> px, py;
  for (int i = 0; i < 6; i++) { px.push_back(atof(argv[6 + 2 * i])); py.push_back(atof(argv[7 + 2 * i])); }
  std::vector<double> vars = mpc.run(vehicle, px, py);
  printf("run");
  for (double x : vars) printf(" %.17g", x);
  printf("\n");
  Eigen::VectorXd state(6);
  state << vars[0], vars[1], vars[2], vars[3], vars[6], vars[7];
  for (int k = 0; k < 25; k++) {
    std::vector<double> s = mpc.solve(state, 40);
    printf("solve");
    for (double x : s) printf(" %.17g", x);
    printf("\n");
    state << s[0], s[1], s[2], s[3], s[4], s[5];
  }
  // optional: latency of MPC::solve through the reference's own class interface (what mpc_main.cpp pays per
  // telemetry message for the solve), "latency <reps>" after the scenario arguments
  if (argc >= 6 + 12 + 2 && std::string(argv[18]) == "latency") {
    const int reps = atoi(argv[19]);
    std::vector<double> us;
    state << vars[0], vars[1], vars[2], vars[3], vars[6], vars[7];
    for (int k = 0; k < reps + 50; k++) {
      const auto t0 = std::chrono::steady_clock::now();
      std::vector<double> s = mpc.solve(state, 40);
      const auto t1 = std::chrono::steady_clock::now();
      if (k >= 50) us.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
      if (k % 25 == 24) state << vars[0], vars[1], vars[2], vars[3], vars[6], vars[7];
      else state << s[0], s[1], s[2], s[3], s[4], s[5];
    }
    std::sort(us.begin(), us.end());
    if (!us.empty()) printf("latency_us p50 %.2f p99 %.2f reps %d\n", us[us.size() / 2], us[(size_t)(us.size() * 0.99)], (int)us.size());
  }
  return 0;
}
