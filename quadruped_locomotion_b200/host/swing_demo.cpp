// swing_demo.cpp - all twelve joint torques of one controller tick through the C++ adapter classes: the stance
// legs from VirtualModelController::compute(), the swing legs from MyRobotSolver::update(limb), merged in the
// shared State the way RosBalanceController::update does (ros_balance_controller.cpp:384-447,472-603).
// stdin (doubles): q[12] qd[12] qdd[12] mask kp[3] kd[3] foot_target_position[12] foot_target_velocity[12].
// Prints {"efforts": [12], "swing": [[3] per leg]}; tests/test_host_adapter.py compares with the oracle.
#include <cstdio>
#include <memory>

#include "qlb_adapter.hpp"

using namespace qlb_host;

int main() {
  double v[12 + 12 + 12 + 1 + 3 + 3 + 12 + 12];
  for (double& x : v)
    if (std::scanf("%lf", &x) != 1) { std::fprintf(stderr, "bad input\n"); return 2; }
  const double* p = v;
  JointPositions q; for (int i = 0; i < 12; i++) q[i] = *p++;
  const double* qd = p; p += 12;
  const double* qdd = p; p += 12;
  const int mask = static_cast<int>(*p++);
  Vector3 kp, kd;
  for (int i = 0; i < 3; i++) kp[i] = *p++;
  for (int i = 0; i < 3; i++) kd[i] = *p++;
  const double* pt = p; p += 12;
  const double* vt = p;
  try {
    auto device = std::make_shared<Device>(QLB_MODEL_QUADRUPED_MODEL, 0);
    auto state = std::make_shared<State>();
    auto cfd = std::make_shared<ContactForceDistribution>(device, state);
    VirtualModelController vmc(device, state, cfd);
    MyRobotSolver swing(device, state);
    if (!cfd->loadParameters() || !vmc.loadParameters() || !swing.loadLimbModelFromURDF()) return 3;
    swing.setGains(kp, kd);
    for (int l = 0; l < 4; l++) state->setSupportLeg(static_cast<LimbEnum>(l), (mask >> l) & 1);
    state->setCurrentLimbJoints(q);
    state->setPoseBaseToWorld({{0, 0, 0.45}}, {{1, 0, 0, 0}});
    state->setTargetPoseBaseToWorld({{0, 0, 0.45}}, {{1, 0, 0, 0}});
    if (!vmc.compute()) return 4;   // stance legs: gravity compensation distributed over the support legs
    std::printf("{\"swing\": [");
    bool firstl = true;
    for (int l = 0; l < 4; l++) {
      const LimbEnum limb = static_cast<LimbEnum>(l);
      if (state->isSupportLeg(limb)) continue;
      swing.setJointVelocityAndAcceleration(limb, {{qd[3 * l], qd[3 * l + 1], qd[3 * l + 2]}}, {{qdd[3 * l], qdd[3 * l + 1], qdd[3 * l + 2]}});
      swing.setDesiredPositionAndVelocity(limb, {{pt[3 * l], pt[3 * l + 1], pt[3 * l + 2]}}, {{vt[3 * l], vt[3 * l + 1], vt[3 * l + 2]}});
      if (!swing.update(limb)) return 5;
      std::printf("%s[%d, %.17g, %.17g, %.17g]", firstl ? "" : ", ", l, swing.getVecTauAct()[0], swing.getVecTauAct()[1], swing.getVecTauAct()[2]);
      firstl = false;
    }
    std::printf("], \"efforts\": [");
    for (int i = 0; i < 12; i++) std::printf("%s%.17g", i ? ", " : "", state->getAllJointEfforts()[i]);
    std::printf("]}\n");
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
