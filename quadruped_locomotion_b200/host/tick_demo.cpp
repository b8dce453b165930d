// tick_demo.cpp - one controller tick (batch of 1) through the C++ adapter classes, the way
// RosBalanceController::update drives the reference (ros_balance_controller.cpp:384-447):
// set feedback + targets in the shared State, call VirtualModelController::compute(), read efforts.
// Scenario on stdin (all doubles): q[12] quat[4] pos[3] linvel[3] angvel[3] tquat[4] tpos[3] tlinvel[3]
// tangvel[3] mask mu ; also a direct computeForceDistribution(F, T) call with wrench[6].
// Prints one JSON object; tests/test_host_adapter.py compares it with the CPU oracle.
#include <cstdio>
#include <memory>

#include "qlb_adapter.hpp"

using namespace qlb_host;

int main() {
  double v[12 + 4 + 3 + 3 + 3 + 4 + 3 + 3 + 3 + 2 + 6];
  for (double& x : v)
    if (std::scanf("%lf", &x) != 1) { std::fprintf(stderr, "bad input\n"); return 2; }
  const double* p = v;
  JointPositions q; for (int i = 0; i < 12; i++) q[i] = *p++;
  Quaternion quat; for (int i = 0; i < 4; i++) quat[i] = *p++;
  Position pos; for (int i = 0; i < 3; i++) pos[i] = *p++;
  LinearVelocity lv; for (int i = 0; i < 3; i++) lv[i] = *p++;
  LocalAngularVelocity av; for (int i = 0; i < 3; i++) av[i] = *p++;
  Quaternion tquat; for (int i = 0; i < 4; i++) tquat[i] = *p++;
  Position tpos; for (int i = 0; i < 3; i++) tpos[i] = *p++;
  LinearVelocity tlv; for (int i = 0; i < 3; i++) tlv[i] = *p++;
  LocalAngularVelocity tav; for (int i = 0; i < 3; i++) tav[i] = *p++;
  const int mask = static_cast<int>(*p++);
  const double mu = *p++;
  Force F; Torque T;
  for (int i = 0; i < 3; i++) F[i] = *p++;
  for (int i = 0; i < 3; i++) T[i] = *p++;

  try {
    auto device = std::make_shared<Device>(QLB_MODEL_QUADRUPED_MODEL, 0);
    auto state = std::make_shared<State>();
    auto cfd = std::make_shared<ContactForceDistribution>(device, state);
    VirtualModelController vmc(device, state, cfd);
    cfd->setFrictionCoefficient(mu);
    if (!cfd->loadParameters() || !vmc.loadParameters()) { std::fprintf(stderr, "loadParameters failed\n"); return 3; }
    for (int l = 0; l < 4; l++) state->setSupportLeg(static_cast<LimbEnum>(l), (mask >> l) & 1);
    state->setCurrentLimbJoints(q);
    state->setPoseBaseToWorld(pos, quat);
    state->setBaseStateFromFeedback(lv, av);
    state->setTargetPoseBaseToWorld(tpos, tquat);
    state->setTargetBaseTwist(tlv, tav);

    // (1) wrench mode: ContactForceDistribution::computeForceDistribution(F, T)
    const bool ok1 = cfd->computeForceDistribution(F, T);
    Force nf; Torque nt;
    cfd->getNetForceAndTorqueOnBase(nf, nt);
    std::printf("{\"cfd_ok\": %s, \"cfd_contact_force\": [", ok1 ? "true" : "false");
    for (int l = 0; l < 4; l++)
      for (int a = 0; a < 3; a++)
        std::printf("%s%.17g", (l || a) ? ", " : "", cfd->getLegInfo(static_cast<LimbEnum>(l)).desiredContactForce_[a]);
    std::printf("], \"cfd_efforts\": [");
    for (int i = 0; i < 12; i++) std::printf("%s%.17g", i ? ", " : "", state->getAllJointEfforts()[i]);
    std::printf("], \"cfd_net\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g], \"cfd_flags\": %u, ", nf[0], nf[1], nf[2], nt[0], nt[1],
                nt[2], cfd->getLastFlags());

    // (2) state mode: VirtualModelController::compute()
    const bool ok2 = vmc.compute();
    std::printf("\"vmc_ok\": %s, \"vmc_wrench\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g], \"vmc_efforts\": [", ok2 ? "true" : "false",
                vmc.getDesiredVirtualForceInBaseFrame()[0], vmc.getDesiredVirtualForceInBaseFrame()[1],
                vmc.getDesiredVirtualForceInBaseFrame()[2], vmc.getDesiredVirtualTorqueInBaseFrame()[0],
                vmc.getDesiredVirtualTorqueInBaseFrame()[1], vmc.getDesiredVirtualTorqueInBaseFrame()[2]);
    for (int i = 0; i < 12; i++) std::printf("%s%.17g", i ? ", " : "", state->getAllJointEfforts()[i]);
    std::printf("]}\n");
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
