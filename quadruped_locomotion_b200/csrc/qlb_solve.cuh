// qlb_solve.cuh - what the fused kernels share: the argument block of one solve call, solver constants and
// the rotation-error map of the virtual model controller.
//
// Reference path being replaced (per state): ContactForceDistribution::computeForceDistribution
// (balance_controller/src/contact_force_distribution/ContactForceDistribution.cpp:99-136) with
// QuadrupedKinematics FK / Jacobian / gravity (quadruped_model/src/quadrupedkinematics.cpp:143-278,
// 485-552) underneath and, in state mode, VirtualModelController::compute
// (balance_controller/src/motion_control/VirtualModelController.cpp:89-268) in front.
// The kernels are in qlb_solve_quad.cuh.  (The earlier half-warp-per-QP kernel with the 12x12 system in
// shared memory, 1.7e8 QP/s, is in the history of this file; profiles/r1_v1..v7 are its ncu records.)
#pragma once

#include "qlb_device.cuh"

namespace qlb {

constexpr int kModePolish = 0, kModeIpm = 1, kModeDone = 2;
constexpr int kPolishPasses = 10;  // repair passes per polish attempt after the interior point
constexpr double kNeighbourhood = 1e-3;

template <typename T>
struct SolveArgsT {
  unsigned long long B;
  const T* q;
  const T* quat;
  const T* wrench;
  const uint8_t* mask;
  const T* mu;       // may be null
  const T* normals;  // may be null
  // state mode
  const T* pose;
  const T* twist;
  const T* tpose;
  const T* ttwist;
  T* grf;
  T* tau;
  uint32_t* flags;
  T* netwrench;      // may be null
  T* wrench_out;     // may be null (state mode)
  unsigned long long* counter;  // work counter, zeroed before launch
  unsigned long long* counter2; // work counter of the second pass (leg-per-lane kernels)
  unsigned* list;               // indices of the states left for the second pass, or null
  unsigned* list_pat;           // per listed state: its first repaired pattern, 5 bits per leg
  unsigned* list_count;         // their number
  unsigned long long* counter3; // work counter of the third pass
  unsigned* list2;              // states left for the interior-point pass
  unsigned* list2_count;
  const DeviceModelT<T>* model;
  const DeviceParamsT<T>* params;
  const DeviceParamsT<double>* params64;  // the same parameters in FP64 (solver core of the mixed FP32 variant)
};
using SolveArgs = SolveArgsT<double>;

// kindr logarithmic map of (q_t^-1 * q): see VirtualModelController.cpp:120,124
template <typename T>
__device__ __forceinline__ void quat_rel_log(const T* qt, const T* q, T (&v)[3]) {
  // rel = conj(qt) * q
  const T aw = qt[0], ax = -qt[1], ay = -qt[2], az = -qt[3];
  T w = aw * q[0] - ax * q[1] - ay * q[2] - az * q[3];
  T x = aw * q[1] + ax * q[0] + ay * q[3] - az * q[2];
  T y = aw * q[2] - ax * q[3] + ay * q[0] + az * q[1];
  T z = aw * q[3] + ax * q[2] - ay * q[1] + az * q[0];
  if (w < T(0.0)) { w = -w; x = -x; y = -y; z = -z; }
  const T n = sqrt(x * x + y * y + z * z);
  T k = T(2.0);
  if (n >= T(1e-12)) k = T(2.0) * atan2(n, w) / n;
  v[0] = k * x; v[1] = k * y; v[2] = k * z;
}

}  // namespace qlb
