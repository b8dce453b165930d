// qlb_batch_adapter.hpp - mirror of the reference's batch-of-states classes over the one-call plan preview
// (qlb_preview_plan_host):
//   free_gait::StateBatch          (free_gait_core/include/free_gait_core/executor/StateBatch.hpp:19-65)
//   free_gait::StateBatchComputer  (free_gait_core/src/executor/StateBatchComputer.cpp:20-132)
// The reference fills a StateBatch with one State per 10 ms of a planned motion (BatchExecutor::processInThread,
// BatchExecutor.cpp:69-83) and then walks it state by state on the CPU, calling the kinematics for every limb of
// every state.  Here the whole batch goes to the GPU in one call; besides the end-effector trajectories and stances of
// the reference it also yields what the balance controller would command along the plan - contact forces, joint
// torques, friction margins - which is the preview SURVEY 8f rank 2 asks for.
#pragma once

#include <map>
#include <memory>
#include <stdexcept>
#include <vector>

#include "qlb_adapter.hpp"

namespace qlb_host {

using Stance = std::map<LimbEnum, Position>;   // free_gait::Stance: support foot positions in the world frame

class StateBatch {
 public:
  // StateBatch.hpp:30-53
  const std::map<double, State>& getStates() const { return states_; }
  void addState(const double time, const State& state) { states_[time] = state; }
  bool isValidTime(const double time) const { return !states_.empty() && time >= getStartTime() && time <= getEndTime(); }
  double getStartTime() const { return states_.begin()->first; }
  double getEndTime() const { return states_.rbegin()->first; }
  const State& getState(const double time) const {   // the state at or just before `time`
    auto it = states_.upper_bound(time);
    if (it == states_.begin()) throw std::out_of_range("StateBatch::getState: time before the batch");
    return std::prev(it)->second;
  }
  void clear() {
    states_.clear(); endEffectorPositions_.clear(); stances_.clear(); contactForces_.clear(); jointTorques_.clear();
    frictionMargins_.clear(); normalForceSlacks_.clear(); flags_.clear();
  }
  std::vector<std::map<double, Position>> getEndEffectorPositions() const { return endEffectorPositions_; }
  std::map<double, Stance> getStances() const { return stances_; }
  // the preview (not in the reference): per time sample
  const std::map<double, std::array<Force, 4>>& getContactForces() const { return contactForces_; }   // desiredContactForce_ = -grf
  const std::map<double, JointEfforts>& getJointTorques() const { return jointTorques_; }
  const std::map<double, double>& getFrictionMargins() const { return frictionMargins_; }
  const std::map<double, double>& getNormalForceSlacks() const { return normalForceSlacks_; }
  const std::map<double, uint32_t>& getFlags() const { return flags_; }

  friend class StateBatchComputer;

 private:
  std::map<double, State> states_;
  std::vector<std::map<double, Position>> endEffectorPositions_;
  std::map<double, Stance> stances_;
  std::map<double, std::array<Force, 4>> contactForces_;
  std::map<double, JointEfforts> jointTorques_;
  std::map<double, double> frictionMargins_, normalForceSlacks_;
  std::map<double, uint32_t> flags_;
};

class StateBatchComputer {
 public:
  explicit StateBatchComputer(std::shared_ptr<Device> device) : device_(std::move(device)) {}

  // One GPU call for the whole batch; fills everything the compute*() functions below hand out.
  bool computeAll(StateBatch& batch) {
    const size_t B = batch.states_.size();
    records_.resize(B);
    preview_.resize(B);
    size_t i = 0;
    for (const auto& kv : batch.states_) {
      const State& s = kv.second;
      qlb_robot_state_record& r = records_[i++];
      // the planned state: target pose / twist where the plan sets them (the executor's State carries both)
      for (int a = 0; a < 3; a++) {
        r.base_position[a] = s.getPositionWorldToBaseInWorldFrame()[a];
        r.base_linear_velocity[a] = s.getLinearVelocityBaseInWorldFrame()[a];
        r.base_angular_velocity[a] = s.getAngularVelocityBaseInBaseFrame()[a];
      }
      const Quaternion& q = s.getOrientationBaseToWorld();   // (w, x, y, z) -> message order x, y, z, w
      r.base_orientation_xyzw[0] = q[1]; r.base_orientation_xyzw[1] = q[2]; r.base_orientation_xyzw[2] = q[3]; r.base_orientation_xyzw[3] = q[0];
      for (int j = 0; j < 12; j++) r.joint_position[j] = s.getJointPositionFeedback()[j];
      for (int l = 0; l < 4; l++) {
        const LimbEnum limb = static_cast<LimbEnum>(l);
        for (int a = 0; a < 3; a++) r.surface_normal[3 * l + a] = s.getSurfaceNormal(limb)[a];
        r.support_leg[l] = s.isSupportLeg(limb) ? 1 : 0;
        r.reserved[l] = 0;
      }
    }
    const double mu = device_->params().friction_default;
    const double mus[4] = {mu, mu, mu, mu};
    if (qlb_preview_plan_host(device_->ctx(), B, records_.data(), mus, preview_.data()) != QLB_OK) return false;
    batch.contactForces_.clear(); batch.jointTorques_.clear(); batch.frictionMargins_.clear(); batch.normalForceSlacks_.clear();
    batch.flags_.clear();
    i = 0;
    for (const auto& kv : batch.states_) {
      const qlb_preview_record& p = preview_[i++];
      std::array<Force, 4> f;
      JointEfforts tau;
      for (int l = 0; l < 4; l++) f[l] = {-p.grf[3 * l], -p.grf[3 * l + 1], -p.grf[3 * l + 2]};
      for (int j = 0; j < 12; j++) tau[j] = p.tau[j];
      batch.contactForces_[kv.first] = f;
      batch.jointTorques_[kv.first] = tau;
      batch.frictionMargins_[kv.first] = p.friction_margin;
      batch.normalForceSlacks_[kv.first] = p.min_normal_slack;
      batch.flags_[kv.first] = p.flags;
    }
    computed_for_ = B;
    return true;
  }

  // StateBatchComputer::computeEndEffectorTrajectories (StateBatchComputer.cpp:64-77)
  void computeEndEffectorTrajectories(StateBatch& batch) {
    if (computed_for_ != batch.states_.size() && !computeAll(batch)) throw std::runtime_error("preview failed");
    batch.endEffectorPositions_.assign(4, {});
    size_t i = 0;
    for (const auto& kv : batch.states_) {
      const qlb_preview_record& p = preview_[i++];
      for (int l = 0; l < 4; l++) batch.endEffectorPositions_[l][kv.first] = {p.feet_world[3 * l], p.feet_world[3 * l + 1], p.feet_world[3 * l + 2]};
    }
  }

  // StateBatchComputer::computeStances (StateBatchComputer.cpp:79-113): a new stance whenever a support flag changes
  void computeStances(StateBatch& batch) {
    if (computed_for_ != batch.states_.size() && !computeAll(batch)) throw std::runtime_error("preview failed");
    batch.stances_.clear();
    const State* previous = nullptr;
    size_t i = 0;
    for (const auto& kv : batch.states_) {
      const qlb_preview_record& p = preview_[i++];
      bool changed = previous == nullptr;
      for (int l = 0; l < 4 && !changed; l++)
        changed = previous->isSupportLeg(static_cast<LimbEnum>(l)) != kv.second.isSupportLeg(static_cast<LimbEnum>(l));
      if (changed) {
        Stance stance;
        for (int l = 0; l < 4; l++) {
          const LimbEnum limb = static_cast<LimbEnum>(l);
          if (kv.second.isSupportLeg(limb)) stance[limb] = {p.feet_world[3 * l], p.feet_world[3 * l + 1], p.feet_world[3 * l + 2]};
        }
        batch.stances_[kv.first] = stance;
      }
      previous = &kv.second;
    }
  }

 private:
  std::shared_ptr<Device> device_;
  std::vector<qlb_robot_state_record> records_;
  std::vector<qlb_preview_record> preview_;
  size_t computed_for_ = 0;
};

}  // namespace qlb_host
