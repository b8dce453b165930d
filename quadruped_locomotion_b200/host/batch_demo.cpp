// batch_demo.cpp - a planned motion through the StateBatch mirror: a trot-like plan sampled every 10 ms
// (BatchExecutor.cpp:69-83), previewed in one GPU call.  Prints one JSON object: number of samples, number of
// stances, the first and last foot position of LF, the smallest friction margin along the plan and the sum of the
// vertical contact forces at three samples (must equal the weight: the preview distributes the gravity compensation).
#include <cmath>
#include <cstdio>

#include "qlb_batch_adapter.hpp"

using namespace qlb_host;

int main() {
  try {
    auto device = std::make_shared<Device>(QLB_MODEL_QUADRUPED_MODEL, 0);
    StateBatch batch;
    const int N = 400;   // 4 s of plan
    for (int k = 0; k < N; k++) {
      const double t = 0.01 * (k + 1);
      State s;
      const double sway = 0.02 * std::sin(2.0 * M_PI * t);
      s.setPoseBaseToWorld({0.05 * t, sway, 0.45}, {1.0, 0.0, 0.0, 0.0});
      s.setBaseStateFromFeedback({0.05, 0.0, 0.0}, {0.0, 0.0, 0.0});
      JointPositions q = {0.0, 0.7, -1.4, 0.0, -0.7, 1.4, 0.0, 0.7, -1.4, 0.0, -0.7, 1.4};
      const int phase = (k / 50) % 4;   // 0: all four, 1: LF+RH swing, 2: all four, 3: RF+LH swing
      for (int l = 0; l < 4; l++) {
        const bool swing = (phase == 1 && (l == 0 || l == 2)) || (phase == 3 && (l == 1 || l == 3));
        s.setSupportLeg(static_cast<LimbEnum>(l), !swing);
        if (swing) q[3 * l + 2] += (l == 0 || l == 2) ? -0.3 : 0.3;   // lift the foot
      }
      s.setCurrentLimbJoints(q);
      batch.addState(t, s);
    }
    StateBatchComputer computer(device);
    if (!computer.computeAll(batch)) return 3;
    computer.computeEndEffectorTrajectories(batch);
    computer.computeStances(batch);
    const auto ee = batch.getEndEffectorPositions();
    double min_margin = 1e300;
    for (const auto& kv : batch.getFrictionMargins()) min_margin = std::fmin(min_margin, kv.second);
    std::printf("{\"samples\": %zu, \"stances\": %zu, \"lf_first\": [%.12g, %.12g, %.12g], \"lf_last\": [%.12g, %.12g, %.12g], \"min_margin\": %.12g, \"fz\": [",
                batch.getStates().size(), batch.getStances().size(), ee[0].begin()->second[0], ee[0].begin()->second[1], ee[0].begin()->second[2],
                ee[0].rbegin()->second[0], ee[0].rbegin()->second[1], ee[0].rbegin()->second[2], min_margin);
    int printed = 0;
    for (const double t : {0.25, 0.75, 1.75}) {
      const auto& f = batch.getContactForces().at(batch.getContactForces().upper_bound(t - 1e-9)->first);
      std::printf("%s%.12g", printed++ ? ", " : "", -(f[0][2] + f[1][2] + f[2][2] + f[3][2]));
    }
    std::printf("]}\n");
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
