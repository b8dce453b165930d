// params_demo.cpp - parameter loading through the adapter classes, the way the reference's controller does it
// (ContactForceDistribution::loadParameters / VirtualModelController::loadParameters reading the ROS parameter
// server, fed from a controller_gains.yaml): a parameter file changes the distribution, a missing key makes
// loadParameters() fail and computeForceDistribution() refuse to run.
// Prints: "loaded <0|1> <0|1>", "fz <LF> <RF> <RH> <LH>" (stance forces with the file's weights),
//         "missing <0|1> <key>", "refused <0|1>".
#include <cstdio>

#include "qlb_adapter.hpp"

using namespace qlb_host;

int main() {
  try {
    auto device = std::make_shared<Device>(QLB_MODEL_QUADRUPED_MODEL, 0);
    auto state = std::make_shared<State>();
    auto cfd = std::make_shared<ContactForceDistribution>(device, state);
    VirtualModelController vmc(device, state, cfd);
    device->parameters().loadYaml(
        "balance_controller:\n"
        "  contact_force_distribution:\n"
        "    weights:\n"
        "      regularizer:\n"
        "        value: 0.01   # a hundred times the default: forces are spread more evenly\n"
        "    constraints:\n"
        "      friction_coefficient: 0.3\n"
        "      minimal_normal_force: 25\n"
        "  virtual_model_controller:\n"
        "    vertical:\n"
        "      kp: 12000\n");
    const bool ok1 = cfd->loadParameters(), ok2 = vmc.loadParameters();
    std::printf("loaded %d %d\n", ok1 ? 1 : 0, ok2 ? 1 : 0);
    state->setCurrentLimbJoints({0.0, 0.7, -1.4, 0.0, -0.7, 1.4, 0.0, 0.7, -1.4, 0.0, -0.7, 1.4});
    state->setPoseBaseToWorld({0, 0, 0.45}, {1, 0, 0, 0});
    if (!cfd->computeForceDistribution({60.0, 0.0, 499.8}, {0.0, 0.0, 0.0})) return 3;
    std::printf("fz %.12g %.12g %.12g %.12g\n", -cfd->getLegInfo(LimbEnum::LF_LEG).desiredContactForce_[2],
                -cfd->getLegInfo(LimbEnum::RF_LEG).desiredContactForce_[2], -cfd->getLegInfo(LimbEnum::RH_LEG).desiredContactForce_[2],
                -cfd->getLegInfo(LimbEnum::LH_LEG).desiredContactForce_[2]);
    std::printf("params %.6g %.6g %.6g %.6g\n", cfd->getGroundForceWeight(), cfd->getMinimalNormalGroundForce(),
                cfd->getFrictionCoefficient(LimbEnum::RH_LEG), device->params().kp_translation[2]);
    device->parameters().deleteParam("/balance_controller/contact_force_distribution/weights/torque/pitch");
    const bool ok3 = cfd->loadParameters();
    std::printf("missing %d %s\n", ok3 ? 0 : 1, device->missingParameter().c_str());
    std::printf("refused %d\n", cfd->computeForceDistribution({0.0, 0.0, 499.8}, {0.0, 0.0, 0.0}) ? 0 : 1);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
