// qlb_solve.cuh - the fused kernel: state -> leg kinematics -> QP assembly -> interior-point solve with
// active-set polish -> joint torques.  Nothing between the input state and the outputs touches HBM.
//
// Reference path being replaced (per state): ContactForceDistribution::computeForceDistribution
// (balance_controller/src/contact_force_distribution/ContactForceDistribution.cpp:99-136) with
// QuadrupedKinematics FK / Jacobian / gravity (quadruped_model/src/quadrupedkinematics.cpp:143-278,
// 485-552) underneath and, in state mode, VirtualModelController::compute
// (balance_controller/src/motion_control/VirtualModelController.cpp:89-268) in front.
#pragma once

#include "qlb_device.cuh"

namespace qlb {

constexpr int kBatch = 4;         // QPs staged per warp batch (32 bytes = one sector per component row)
#ifndef QLB_MIN_CTAS
#define QLB_MIN_CTAS 3
#endif
constexpr int kWarpsPerCta = 4;
constexpr int kThreads = 32 * kWarpsPerCta;

// input staging rows
constexpr int kRowQ = 0, kRowQuat = 12, kRowWrench = 16, kRowMu = 22, kRowNormal = 26, kInRows = 38;
// state mode adds: pose(7) twist(6) target pose(7) target twist(6) -> quat/wrench rows are derived
constexpr int kRowPose = 38, kRowTwist = 45, kRowTPose = 51, kRowTTwist = 58, kInRowsState = 64;
// output staging rows
constexpr int kRowGrf = 0, kRowTau = 12, kRowNet = 24, kRowWout = 30, kOutRows = 36;

constexpr int kModePolish = 0, kModeIpm = 1, kModeDone = 2;
constexpr int kPdasFirst = 1;      // active-set passes tried right after the unconstrained solve
constexpr int kPolishPasses = 10;  // passes per polish attempt
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
  int vec_ok;             // all row pointers 16-byte aligned and B even
};
using SolveArgs = SolveArgsT<double>;

template <int ROWS>
struct alignas(16) WarpSmem {
  double in[ROWS][kBatch];
  double out[kOutRows][kBatch];
  double atl[2][kVars][6];   // per group: wrench-map column of every slot, a_l = [e; r x e]
  double pc[2][kVars][6];    // per group: left factor columns P_l of the 6x6 system N = S^-1 + sum_l P_l C_l'
  double cc[2][kVars][6];    // per group: right factor columns C_l (polish: reduced columns; IPM: a_l)
  double fix[2][4][6];       // per group, per leg: contribution of a pinned normal force
  double tail[7][32];        // per lane: slot direction e (3), Jacobian column (3), gravity torque
  double bw[2][6];           // per group: the wrench b
  double t6[2][8];           // per group: solution of the 6x6 system (S-weighted wrench residual)
  double xb[2][2][16];       // exchange buffer of the factorisation (double buffered, one row per group)
  double vb[2][16];          // per group: vector exchange (right-hand sides)
  double hs[6 * kPitch];     // the round's 6x6 matrix / its Cholesky factor, row r in lane 16*grp + r
  uint32_t flags[kBatch];
  uint8_t mask[kBatch];
};

template <int ROWS>
struct alignas(16) CtaSmem {
  WarpSmem<ROWS> w[kWarpsPerCta];
  DeviceParams prm;
};

// ---------------------------------------------------------------- staging (coalesced 16-byte accesses)
template <int ROWS>
__device__ __forceinline__ void stage_in(double (*dst)[kBatch], const double* __restrict__ src, int rows,
                                         unsigned long long B, unsigned long long b0, int nvalid, int lane,
                                         bool vec) {
  if (vec && nvalid == kBatch) {
    const int chunks = rows * (kBatch / 2);
    for (int ch = lane; ch < chunks; ch += 32) {
      const int r = ch / (kBatch / 2), o = ch % (kBatch / 2);
      const double2 v = __ldg(reinterpret_cast<const double2*>(src + (size_t)r * B + b0) + o);
      *reinterpret_cast<double2*>(&dst[r][2 * o]) = v;
    }
  } else {
    for (int e = lane; e < rows * kBatch; e += 32) {
      const int r = e / kBatch, o = e % kBatch;
      dst[r][o] = (o < nvalid) ? __ldg(src + (size_t)r * B + b0 + o) : 0.0;
    }
  }
}
__device__ __forceinline__ void stage_out(double* __restrict__ dst, const double (*src)[kBatch], int rows,
                                          unsigned long long B, unsigned long long b0, int nvalid, int lane,
                                          bool vec) {
  if (vec && nvalid == kBatch) {
    const int chunks = rows * (kBatch / 2);
    for (int ch = lane; ch < chunks; ch += 32) {
      const int r = ch / (kBatch / 2), o = ch % (kBatch / 2);
      reinterpret_cast<double2*>(dst + (size_t)r * B + b0)[o] = *reinterpret_cast<const double2*>(&src[r][2 * o]);
    }
  } else {
    for (int e = lane; e < rows * kBatch; e += 32) {
      const int r = e / kBatch, o = e % kBatch;
      if (o < nvalid) dst[(size_t)r * B + b0 + o] = src[r][o];
    }
  }
}

// The five rows of a leg are spread over its three lanes: the lane of the normal component (c = 0)
// owns row 0 (y_n >= F_min), the lane of y_1 owns rows 1, 2 (mu y_n +- y_1 >= 0), the lane of y_2 rows
// 3, 4.  Every lane therefore carries at most two rows, "A" and "B" (B is void on the normal lane).
//
// D~ x restricted to this lane's rows: xn = the leg's normal component, x = this lane's component
__device__ __forceinline__ void rows_apply(double xn, double x, double mu, int c, double& eA, double& eB) {
  const double m = mu * xn;
  eA = (c == 0) ? xn : m + x;
  eB = m - x;
}
// (D~' v)_lane: the normal lane needs the sums vA + vB of its two leg-mates (full-warp shuffles)
__device__ __forceinline__ double dt_lane(double vA, double vB, double mu, int c, int l0) {
  const double sv = vA + vB;
  const double S = gshfl(sv, l0 + 1) + gshfl(sv, l0 + 2);
  return (c == 0) ? fma(mu, S, vA) : vA - vB;
}
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) v += __shfl_xor_sync(kFull, v, o, kGroup);
  return v;
}
__device__ __forceinline__ float group_min(float v) {
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) v = fminf(v, __shfl_xor_sync(kFull, v, o, kGroup));
  return v;
}

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

// ---------------------------------------------------------------- one QP per group
// Every lane of the warp calls this (two QPs side by side).  qi = index of this group's QP in the batch.
template <int MODE, int ROWS>
__device__ __forceinline__ void solve_group(WarpSmem<ROWS>& ws, const DeviceModel& mdl, const DeviceParams& prm,
                                            const int qi, const int lane, const bool have_mu,
                                            const bool have_normals, const bool want_net) {
  const int grp = lane >> 4, gl = lane & 15;
  const bool var_lane = gl < kVars;
  const int leg = var_lane ? gl / 3 : 3;
  const int c = var_lane ? gl - 3 * leg : 0;
  const int l0 = 3 * leg;

  // ---------------- inputs
  const unsigned mask = ws.mask[qi] & 0xFu;
  const bool alive = var_lane && ((mask >> leg) & 1u);
  const int ns = __popc(mask);
  const double qv = ws.in[kRowQ + l0 + c][qi];
  double quat[4], b[6];
  bool bad = !isfinite(qv);
  if (MODE == 1) {
    // virtual model controller prologue (VirtualModelController.cpp:104-268)
    double pose[7], tw[6], tp[7], tt[6];
#pragma unroll
    for (int r = 0; r < 7; r++) { pose[r] = ws.in[kRowPose + r][qi]; tp[r] = ws.in[kRowTPose + r][qi]; bad |= !isfinite(pose[r]) || !isfinite(tp[r]); }
#pragma unroll
    for (int r = 0; r < 6; r++) { tw[r] = ws.in[kRowTwist + r][qi]; tt[r] = ws.in[kRowTTwist + r][qi]; bad |= !isfinite(tw[r]) || !isfinite(tt[r]); }
#pragma unroll
    for (int r = 0; r < 4; r++) quat[r] = pose[3 + r];
    const double w = quat[0], x = quat[1], y = quat[2], z = quat[3];
    double R[9];
    R[0] = w * w + x * x - y * y - z * z; R[1] = 2.0 * (x * y - w * z); R[2] = 2.0 * (x * z + w * y);
    R[3] = 2.0 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2.0 * (y * z - w * x);
    R[6] = 2.0 * (x * z - w * y); R[7] = 2.0 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
    double ep[3], ev[3], ew[3], eR[3];
#pragma unroll
    for (int a = 0; a < 3; a++) { ep[a] = tp[a] - pose[a]; ev[a] = tt[a] - tw[a]; ew[a] = tt[3 + a] - tw[3 + a]; }
    quat_rel_log(tp + 3, quat, eR);
#pragma unroll
    for (int a = 0; a < 3; a++) eR[a] = -eR[a];
    // gravity compensation (VMC.cpp:162-188): g_b = R^T (0,0,-g)
    double gb[3], Fg[3], Tg[3], ft[3];
#pragma unroll
    for (int a = 0; a < 3; a++) { gb[a] = -prm.gravity * R[6 + a]; ft[a] = -prm.grav_pct * prm.torso_mass * gb[a]; Fg[a] = ft[a]; }
    Tg[0] = prm.com[1] * ft[2] - prm.com[2] * ft[1];
    Tg[1] = prm.com[2] * ft[0] - prm.com[0] * ft[2];
    Tg[2] = prm.com[0] * ft[1] - prm.com[1] * ft[0];
#pragma unroll
    for (int l = 0; l < 4; l++) {
      double fl[3], r[3];
#pragma unroll
      for (int a = 0; a < 3; a++) { fl[a] = -prm.grav_pct * prm.leg_mass[l] * gb[a]; Fg[a] += fl[a]; r[a] = prm.leg_pos[l][a] - prm.com[a]; }
      Tg[0] += r[1] * fl[2] - r[2] * fl[1];
      Tg[1] += r[2] * fl[0] - r[0] * fl[2];
      Tg[2] += r[0] * fl[1] - r[1] * fl[0];
    }
    const double ffx = tt[0], ffy = tt[1];
    const double gfz = prm.kp_t[2] * ep[2], gdz = prm.kd_t[2] * ev[2];
    const double dwv[3] = {prm.kd_r[0] * ew[0], prm.kd_r[1] * ew[1], prm.kd_r[2] * ew[2]};
    const double fwz = prm.kff_r[2] * tt[5];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      // R^T v, component a = column a of R dotted with v
      const double epb = R[a] * ep[0] + R[3 + a] * ep[1] + R[6 + a] * ep[2];
      const double evb = R[a] * ev[0] + R[3 + a] * ev[1] + R[6 + a] * ev[2];
      const double ffb = R[a] * ffx + R[3 + a] * ffy;
      b[a] = prm.kp_t[a] * epb + prm.kd_t[a] * evb + prm.kff_t[a] * ffb + Fg[a] + R[6 + a] * gfz + R[6 + a] * gdz;
      const double dwb = R[a] * dwv[0] + R[3 + a] * dwv[1] + R[6 + a] * dwv[2];
      b[3 + a] = prm.kp_r[a] * eR[a] + dwb + R[6 + a] * fwz + Tg[a];
    }
    if (gl < 6) ws.out[kRowWout + gl][qi] = b[gl];
  } else {
#pragma unroll
    for (int r = 0; r < 4; r++) { quat[r] = ws.in[kRowQuat + r][qi]; bad |= !isfinite(quat[r]); }
#pragma unroll
    for (int r = 0; r < 6; r++) { b[r] = ws.in[kRowWrench + r][qi]; bad |= !isfinite(b[r]); }
  }
  const double mu = have_mu ? ws.in[kRowMu + leg][qi] : prm.mu_default;
  double nw[3] = {0.0, 0.0, 1.0};
  if (have_normals) {
#pragma unroll
    for (int a = 0; a < 3; a++) nw[a] = ws.in[kRowNormal + l0 + a][qi];
  }

  // ---------------- base rotation, friction frame (CFD.cpp:223,237,286-309), gravity in base frame (:518-519)
  double nb[3], t1[3], t2[3], gb[3];
  {
    const double w = quat[0], x = quat[1], y = quat[2], z = quat[3];
    double R[9];
    R[0] = w * w + x * x - y * y - z * z; R[1] = 2.0 * (x * y - w * z); R[2] = 2.0 * (x * z + w * y);
    R[3] = 2.0 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2.0 * (y * z - w * x);
    R[6] = 2.0 * (x * z - w * y); R[7] = 2.0 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      nb[a] = R[a] * nw[0] + R[3 + a] * nw[1] + R[6 + a] * nw[2];  // R^T n_world
      gb[a] = -prm.gravity * R[6 + a];
    }
    const double ey[3] = {R[3], R[4], R[5]};  // R^T e_y
    t1[0] = nb[1] * ey[2] - nb[2] * ey[1];
    t1[1] = nb[2] * ey[0] - nb[0] * ey[2];
    t1[2] = nb[0] * ey[1] - nb[1] * ey[0];
    double rn = fast_rsqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
#pragma unroll
    for (int a = 0; a < 3; a++) t1[a] *= rn;
    t2[0] = nb[1] * t1[2] - nb[2] * t1[1];
    t2[1] = nb[2] * t1[0] - nb[0] * t1[2];
    t2[2] = nb[0] * t1[1] - nb[1] * t1[0];
    rn = fast_rsqrt(t2[0] * t2[0] + t2[1] * t2[1] + t2[2] * t2[2]);
#pragma unroll
    for (int a = 0; a < 3; a++) t2[a] *= rn;
    if (alive) {
      bad |= !isfinite(mu);
#pragma unroll
      for (int a = 0; a < 3; a++) bad |= !isfinite(nb[a]) || !isfinite(t1[a]) || !isfinite(t2[a]);
    }
  }
  // group-wide verdict on the inputs
  const unsigned badbits = (__ballot_sync(kFull, bad) >> (16 * grp)) & 0xFFFFu;

  // ---------------- leg forward kinematics, Jacobian column c, gravity torque c (QK.cpp:143-278,485-552)
  double foot[3], jcol[3], gtau;
  {
    double sn, cs;
    sincos_small(qv, &sn, &cs);
    double R[9], p[3], zc[3] = {0, 0, 0}, pjc[3] = {0, 0, 0}, mcs[3] = {0, 0, 0};
#pragma unroll
    for (int e = 0; e < 9; e++) R[e] = mdl.rot[leg][0][e];
#pragma unroll
    for (int a = 0; a < 3; a++) p[a] = mdl.xyz[leg][0][a];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (j > 0) {
        const double* xj = mdl.xyz[leg][j];
#pragma unroll
        for (int a = 0; a < 3; a++) p[a] += R[3 * a] * xj[0] + R[3 * a + 1] * xj[1] + R[3 * a + 2] * xj[2];
        if (j < 3) {
          const double* Rj = mdl.rot[leg][j];
          double T[9];
#pragma unroll
          for (int r = 0; r < 3; r++)
#pragma unroll
            for (int s = 0; s < 3; s++) T[3 * r + s] = R[3 * r] * Rj[s] + R[3 * r + 1] * Rj[3 + s] + R[3 * r + 2] * Rj[6 + s];
#pragma unroll
          for (int e = 0; e < 9; e++) R[e] = T[e];
        }
      }
      if (j < 3) {
        if (c == j) {
#pragma unroll
          for (int a = 0; a < 3; a++) { zc[a] = R[3 * a + 2]; pjc[a] = p[a]; }
        }
        const double cj = gshfl(cs, l0 + j), sj = gshfl(sn, l0 + j);
#pragma unroll
        for (int r = 0; r < 3; r++) {  // R = R * Rz(q_j)
          const double a0 = R[3 * r], a1 = R[3 * r + 1];
          R[3 * r] = cj * a0 + sj * a1;
          R[3 * r + 1] = cj * a1 - sj * a0;
        }
      }
      const double* cm = mdl.com[leg][j];
      const double mj = (j >= c) ? mdl.mass[leg][j] : 0.0;
#pragma unroll
      for (int a = 0; a < 3; a++)
        mcs[a] += mj * (p[a] + R[3 * a] * cm[0] + R[3 * a + 1] * cm[1] + R[3 * a + 2] * cm[2]);
    }
#pragma unroll
    for (int a = 0; a < 3; a++) foot[a] = p[a];
    const double dv[3] = {foot[0] - pjc[0], foot[1] - pjc[1], foot[2] - pjc[2]};
    jcol[0] = zc[1] * dv[2] - zc[2] * dv[1];
    jcol[1] = zc[2] * dv[0] - zc[0] * dv[2];
    jcol[2] = zc[0] * dv[1] - zc[1] * dv[0];
    const double ms = mdl.msuf[leg][c];
    const double arm[3] = {mcs[0] - ms * pjc[0], mcs[1] - ms * pjc[1], mcs[2] - ms * pjc[2]};
    gtau = -(zc[0] * (arm[1] * gb[2] - arm[2] * gb[1]) + zc[1] * (arm[2] * gb[0] - arm[0] * gb[2]) +
             zc[2] * (arm[0] * gb[1] - arm[1] * gb[0]));
  }

  // ---------------- wrench-map column of this slot: [e; r x e], e = column c of Q_leg (CFD.cpp:186-200)
  double ev[3];
#pragma unroll
  for (int a = 0; a < 3; a++) ev[a] = (c == 0) ? nb[a] : (c == 1 ? t1[a] : t2[a]);
  double at_raw[6];
  at_raw[0] = ev[0]; at_raw[1] = ev[1]; at_raw[2] = ev[2];
  at_raw[3] = foot[1] * ev[2] - foot[2] * ev[1];
  at_raw[4] = foot[2] * ev[0] - foot[0] * ev[2];
  at_raw[5] = foot[0] * ev[1] - foot[1] * ev[0];
  if (var_lane) {
#pragma unroll
    for (int r = 0; r < 6; r++) ws.atl[grp][gl][r] = alive ? at_raw[r] : 0.0;
  }
  __syncwarp();

  // ---------------- solver state
  // The 12x12 systems of this QP all have the form  K + A~' S A~  with K block diagonal (3x3 per leg) and
  // A~ the 6 x 12 wrench map, so they are solved through the 6x6 "dual" system
  //     (S^-1 + A~ K^-1 A~') t = A~ K^-1 r,      x = K^-1 (r - A~' t)
  // (push-through / Woodbury): one 6x6 Cholesky per round instead of a 12x12 one.  For the polish rounds
  // t is the S-weighted wrench residual S(b - A~ y), from which the gradient follows without a mat-vec.
  // Everything that contains a shuffle is executed by the whole warp with the full mask; a group that
  // is not in the corresponding mode just computes values it never commits.
  if (var_lane) {
#pragma unroll
    for (int a = 0; a < 3; a++) { ws.tail[a][lane] = ev[a]; ws.tail[3 + a][lane] = jcol[a]; }
    ws.tail[6][lane] = gtau;
  }
  if (gl < 6) ws.bw[grp][gl] = b[gl];
  // strictly feasible interior-point start for this leg: push c0 along the normal
  const double c0 = fmax(fmax(2.0 * prm.fmin, (b[0] * nb[0] + b[1] * nb[1] + b[2] * nb[2]) * (ns > 0 ? 1.0 / ns : 0.0)),
                         prm.fmin + 1.0);
  const double* const at = ws.atl[grp][var_lane ? gl : 0];  // own wrench-map column (kept in shared memory)
  double gt = 0.0;  // g~ of this slot = -a_l . (S b)
#pragma unroll
  for (int r = 0; r < 6; r++) gt = fma(-(alive ? at_raw[r] : 0.0) * prm.S[r], b[r], gt);
  const float gscale = fmaxf(1.f, group_max(fabsf((float)gt)));
  __syncwarp();

  double* const hs = ws.hs;
  const int rr = gl % 6, hh = (gl / 6) % 2;  // this lane computes entries (rr, 3hh .. 3hh+2) of the 6x6 matrix
  const double sinv = 1.0 / prm.S[rr];
  const bool rowB = alive && c > 0;  // this lane owns a second constraint row
  double y = 0.0;                    // interior-point iterate / final solution of this slot
  double rd = 0.0;                   // dual residual of this slot, kept up to date incrementally
  double sA = 1.0, sB = 1.0, lamA = 0.0, lamB = 0.0, rpA = 0.0, rpB = 0.0;  // slack, multiplier, primal residual
  int pat = 0;  // active pattern of this lane's rows: normal lane 1 = pinned at F_min; tangential lanes
                // -1 = row A active (y_c = -mu y_n), +1 = row B active (y_c = +mu y_n)
  int mode = kModePolish, it = 0, pass = 0, status = 0;
  bool first = true, converged = false, want_polish = false;
  double alpha_prev = 1.0;
  if (badbits != 0u) { mode = kModeDone; status = 4; }
  else if (ns == 0) { mode = kModeDone; status = 1; }
  const float rm = ns > 0 ? 1.f / (5.f * ns) : 0.f;

  // ---------------- rounds: one 6x6 factorisation + one or two substitutions, shared by both groups
  int rounds = 0;
#pragma unroll 1
  for (;;) {
    if (__all_sync(kFull, mode == kModeDone)) break;
    if (++rounds > 200 && mode != kModeDone) { mode = kModeDone; status = 2; }  // hard stop, never reached in practice
    const bool pol_round = (mode == kModePolish), ipm_round = (mode == kModeIpm);
    const bool any_pol = __any_sync(kFull, pol_round), any_ipm = __any_sync(kFull, ipm_round);
    double rhs = 0.0, rsA = 1.0, rsB = 1.0, k0 = 0.0, k1 = 0.0, k2 = 0.0, iwd = 0.0, thA = 0.0, thB = 0.0;

    // ---- B1. polish: reduced columns C_l of the pattern, P_l = C_l / wd_l, pinned contributions
    if (any_pol) {
      const int p1 = gshfl(pat, l0 + 1), p2 = gshfl(pat, l0 + 2);
      if (pol_round && var_lane) {
        const bool free_slot = alive && (pat == 0);
        const double* a1 = ws.atl[grp][l0 + 1];
        const double* a2 = ws.atl[grp][l0 + 2];
        const double q1 = p1 * mu, q2 = p2 * mu;
        const double wd = (c == 0) ? prm.W * fma(mu * mu, (double)(p1 * p1 + p2 * p2), 1.0) : prm.W;
        iwd = free_slot ? 1.0 / wd : 0.0;
        const double f = (c == 0 && pat != 0 && alive) ? prm.fmin : 0.0;
#pragma unroll
        for (int r = 0; r < 6; r++) {
          const double cn = (c == 0) ? fma(q2, a2[r], fma(q1, a1[r], at[r])) : at[r];
          const double cr = free_slot ? cn : 0.0;
          ws.cc[grp][gl][r] = cr;
          ws.pc[grp][gl][r] = cr * iwd;
          if (c == 0) ws.fix[grp][leg][r] = f * cn;
        }
      }
    }
    // ---- B2. interior point: K = w I + D~' diag(lam/s) D~ per leg (3x3 arrow matrix), its inverse,
    //          P_l = A~ K^-1 e_l, C_l = a_l, predictor right-hand side
    if (any_ipm) {
      rsA = fast_rcp(sA); rsB = fast_rcp(sB);
      thA = lamA * rsA; thB = lamB * rsB;
      const double T = thA + thB, Dl = thA - thB;
      const double T0 = gshfl(T, l0), T1 = gshfl(T, l0 + 1), D1 = gshfl(Dl, l0 + 1), T2 = gshfl(T, l0 + 2), D2 = gshfl(Dl, l0 + 2);
      const double dtv = dt_lane(fma(-thA, rpA, lamA), fma(-thB, rpB, lamB), mu, c, l0);
      // arrow matrix [[a, b1, b2], [b1, d1, 0], [b2, 0, d2]] and its inverse via the Schur complement of a
      const double d1 = prm.W + T1, d2 = prm.W + T2, b1 = mu * D1, b2 = mu * D2;
      const double a = prm.W + fma(mu * mu, T1 + T2, T0);
      const double id1 = fast_rcp(d1), id2 = fast_rcp(d2);
      const double e1 = b1 * id1, e2 = b2 * id2;
      const double isg = fast_rcp(a - b1 * e1 - b2 * e2);
      if (c == 0) { k0 = isg; k1 = -e1 * isg; k2 = -e2 * isg; }
      else if (c == 1) { k0 = -e1 * isg; k1 = fma(e1 * e1, isg, id1); k2 = e1 * e2 * isg; }
      else { k0 = -e2 * isg; k1 = e1 * e2 * isg; k2 = fma(e2 * e2, isg, id2); }
      if (ipm_round && var_lane) {
        const double* an = ws.atl[grp][l0];
        const double* a1 = ws.atl[grp][l0 + 1];
        const double* a2 = ws.atl[grp][l0 + 2];
#pragma unroll
        for (int r = 0; r < 6; r++) {
          ws.pc[grp][gl][r] = fma(k2, a2[r], fma(k1, a1[r], k0 * an[r]));
          ws.cc[grp][gl][r] = at[r];
        }
        rhs = -rd - dtv;
        ws.vb[grp][gl] = rhs;
      }
    }
    __syncwarp();

    // ---- N. the 6x6 system: N = S^-1 + sum_l P_l C_l' (three entries per lane), right-hand side
    //         polish: b - pinned contributions;  interior point: sum_l P_l r_l
    double rhs6 = 0.0;
    if (var_lane && mode != kModeDone) {
      double n0 = (rr == 3 * hh) ? sinv : 0.0, n1 = (rr == 3 * hh + 1) ? sinv : 0.0, n2 = (rr == 3 * hh + 2) ? sinv : 0.0;
#pragma unroll 2
      for (int l = 0; l < kVars; l++) {
        const double pl = ws.pc[grp][l][rr];
        const double* cl = &ws.cc[grp][l][3 * hh];
        n0 = fma(pl, cl[0], n0);
        n1 = fma(pl, cl[1], n1);
        n2 = fma(pl, cl[2], n2);
        if (ipm_round) rhs6 = fma(pl, ws.vb[grp][l], rhs6);
      }
      hs[(3 * hh) * kPitch + 16 * grp + rr] = n0;
      hs[(3 * hh + 1) * kPitch + 16 * grp + rr] = n1;
      hs[(3 * hh + 2) * kPitch + 16 * grp + rr] = n2;
      if (pol_round)
        rhs6 = ws.bw[grp][rr] - ((ws.fix[grp][0][rr] + ws.fix[grp][1][rr]) + (ws.fix[grp][2][rr] + ws.fix[grp][3][rr]));
    } else if (gl < 6) {
#pragma unroll
      for (int j = 0; j < 6; j++) hs[j * kPitch + lane] = (j == gl) ? 1.0 : 0.0;
    }
    __syncwarp();

    // ---- C. factorise (forward substitution of the first right-hand side fused in), back-substitute
    double rdiag, zf = (gl < 6) ? rhs6 : 0.0;
    const bool pd = reg_cholesky6_fwd(hs, ws.xb, rdiag, zf, grp, gl, lane);
    if (!pd && mode != kModeDone) { mode = kModeDone; status = 4; y = 0.0; }

    double sol = 0.0, rcA = 0.0, rcB = 0.0;
#pragma unroll 1
    for (int ph = 0; ph < (any_ipm ? 2 : 1); ph++) {
      if (ph == 1) {
        // corrector: new right-hand side of the 6x6 system, full substitution
        if (var_lane) ws.vb[grp][gl] = rhs;
        __syncwarp();
        double r6 = 0.0;
        if (gl < 6) {
#pragma unroll 2
          for (int l = 0; l < kVars; l++) r6 = fma(ws.pc[grp][l][gl], ws.vb[grp][l], r6);
        }
        zf = smem_forward<6>(hs, ws.xb, rdiag, r6, grp, gl, lane);
      }
      const double t = smem_backward<6>(hs, ws.xb, rdiag, zf, grp, gl);
      if (gl < 6) ws.t6[grp][gl] = t;
      __syncwarp();
      // a_l . t and c_l . t (c_l = own column of the right factor; = a_l in interior-point rounds)
      double att = 0.0, ctt = 0.0;
      {
        const double* const cl = ws.cc[grp][var_lane ? gl : 0];
#pragma unroll
        for (int r = 0; r < 6; r++) { const double tr = ws.t6[grp][r]; att = fma(at[r], tr, att); ctt = fma(cl[r], tr, ctt); }
      }
      if (ph == 0) sol = ctt * iwd;       // polish: z_l = c_l . t / wd_l
      if (ph == 0 && pol_round) rhs = att;  // keep a_l . t for the gradient
      if (!any_ipm) break;
      // interior point: x = K^-1 (r - A~' t)
      const double wl = rhs - att;
      const double x = fma(k2, gshfl(wl, l0 + 2), fma(k1, gshfl(wl, l0 + 1), k0 * gshfl(wl, l0)));
      // direction of this lane's rows: ds = D~ dy - rp, dl = -(rc + lam ds)/s
      // (phase 0: rc = s lam, phase 1: rc = s lam + dsa dla - sigma mu)
      double deA, deB;
      rows_apply(gshfl(x, l0), x, mu, c, deA, deB);
      if (!alive) deA = 0.0;
      if (!rowB) deB = 0.0;
      const double dsA = deA - rpA, dsB = deB - rpB;
      const double dlA = alive ? -fma(lamA, dsA, (ph == 0) ? sA * lamA : rcA) * rsA : 0.0;
      const double dlB = rowB ? -fma(lamB, dsB, (ph == 0) ? sB * lamB : rcB) * rsB : 0.0;
      float ratio = 0.f;
      if (ipm_round && alive) {
        ratio = fmaxf(-(float)dsA * (float)rsA, -(float)dlA * rcp_approx((float)lamA));
        if (rowB) ratio = fmaxf(ratio, fmaxf(-(float)dsB * (float)rsB, -(float)dlB * rcp_approx((float)lamB)));
        ratio = fmaxf(ratio, 0.f);
      }
      ratio = group_max(ratio);
      if (ph == 0) {
        // affine step length, centring parameter, corrector right-hand side
        const float pa = ipm_round ? (float)(sA * lamA) + (float)(sB * lamB) : 0.f;
        const double ala = (ratio > 1.f) ? 1.0 / (double)ratio : 1.0;
        const float pb = ipm_round ? (float)(fma(ala, dsA, sA) * fma(ala, dlA, lamA)) + (float)(fma(ala, dsB, sB) * fma(ala, dlB, lamB)) : 0.f;
        const float mu_c = group_sum(pa) * rm;
        const float mua = group_sum(pb) * rm;
        const float q3 = (ipm_round && mu_c > 0.f) ? mua / mu_c : 0.f;
        float sigma = q3 * q3 * q3;
        if (alpha_prev < 0.1 && sigma < 0.5f) sigma = 0.5f;  // short step last time: re-centre
        const double sigmu = (double)sigma * (double)mu_c;
        rcA = alive ? fma(dsA, dlA, sA * lamA) - sigmu : 0.0;
        rcB = rowB ? fma(dsB, dlB, sB * lamB) - sigmu : 0.0;
        const double dtv = dt_lane((rcA - lamA * rpA) * rsA, (rcB - lamB * rpB) * rsB, mu, c, l0);
        if (ipm_round) rhs = var_lane ? -rd - dtv : 0.0;
      } else {
        double al = (ratio > 0.995f) ? 0.995 / (double)ratio : 1.0;
        // stay inside the neighbourhood min_i s_i lam_i >= gamma * mu (both groups loop together)
#pragma unroll 1
        for (int tries = 0; tries < 20; tries++) {
          const float prA = alive ? (float)(fma(al, dsA, sA) * fma(al, dlA, lamA)) : 0.f;
          const float prB = rowB ? (float)(fma(al, dsB, sB) * fma(al, dlB, lamB)) : 0.f;
          const float ps = group_sum(prA + prB) * rm;
          const float pm = group_min(fminf(alive ? prA : 3e38f, rowB ? prB : 3e38f));
          const bool ok = !ipm_round || (pm >= (float)kNeighbourhood * ps && pm > 0.f);
          if (__all_sync(kFull, ok)) break;
          if (!ok) al *= 0.7;
        }
        // take the step; the residuals follow without a mat-vec:
        //   G~ dy = rhs - D~' diag(lam/s) D~ dy   =>   rd += al (rhs - D~'(theta .* de + dl)),   rp *= (1 - al)
        const double dtw = dt_lane(fma(thA, deA, dlA), fma(thB, deB, dlB), mu, c, l0);
        if (ipm_round) {
          sA = fma(al, dsA, sA); lamA = fma(al, dlA, lamA); rpA *= (1.0 - al);
          sB = fma(al, dsB, sB); lamB = fma(al, dlB, lamB); rpB *= (1.0 - al);
          rd = fma(al, rhs - dtw, rd);
          y = fma(al, x, y);
          alpha_prev = al;
          it++;
        }
        // convergence test and decision to polish, on the new iterate
        const float mu_n = group_sum(ipm_round ? (float)(sA * lamA) + (float)(sB * lamB) : 0.f) * rm;
        const float nrd = group_max((var_lane && ipm_round) ? fabsf((float)rd) : 0.f);
        const float nrp = group_max(ipm_round ? fmaxf(fabsf((float)rpA), fabsf((float)rpB)) : 0.f);
        const float scale = fmaxf(1.f, group_max(fabsf((float)y)));
        if (ipm_round) {
          const float tolf = (float)prm.tol * scale;
          converged = (mu_n <= tolf) && (nrp <= tolf) && (nrd <= 100.f * tolf);
          const bool out_of_iters = it >= prm.max_iter;
          want_polish = converged || out_of_iters || (it >= 2 && mu_n <= 1e-3f * scale);
          if (out_of_iters && !converged) status = 2;
        }
      }
      __syncwarp();
    }

    // ---- M. polish: recover y, gradient, multipliers, slacks; verify the KKT signs; repair the pattern.
    if (any_pol) {
      // y of the polished point: pinned / tied components follow from the leg's normal component
      double yp = (c == 0) ? ((pat != 0) ? prm.fmin : sol) : sol;
      const double ynp = gshfl(yp, l0);
      if (c != 0 && pat != 0) yp = pat * mu * ynp;
      if (!alive) yp = 0.0;
      // gradient of the objective in contact coordinates: G~ y + g~ = w y - a_l . S(b - A~ y) = w y - a_l . t
      const double gam = fma(prm.W, yp, -rhs);
      // multipliers and slacks of this lane's rows
      double eA, eB;
      rows_apply(ynp, yp, mu, c, eA, eB);
      if (c == 0) eA -= prm.fmin;
      double uA = 0.0, uB = 0.0;
      if (c != 0) { uA = (pat == -1) ? gam : 0.0; uB = (pat == 1) ? -gam : 0.0; }
      const double su = uA + uB;
      const double U = gshfl(su, l0 + 1) + gshfl(su, l0 + 2);
      if (c == 0) uA = (pat != 0) ? gam - mu * U : 0.0;
      const bool actA = (c == 0) ? (pat != 0) : (pat == -1), actB = (pat == 1);
      const float scale = fmaxf(1.f, group_max(fabsf((float)yp)));
      const double tol_u = 1e-13 * (double)gscale, tol_s = 1e-10 * (double)scale;
      // worst violation of this lane: a negative multiplier (drop the row) before a negative slack (add it)
      int fix = 0;  // 0 none, 1 drop, 2 add row A, 3 add row B
      double key = 0.0;
      if (alive && pol_round) {
        if (actA && uA < -tol_u) { fix = 1; key = uA * 1e6; }
        else if (actB && uB < -tol_u) { fix = 1; key = uB * 1e6; }
        else {
          const double vA2 = actA ? 0.0 : eA, vB2 = (rowB && !actB) ? eB : 0.0;
          if (vA2 < -tol_s && vA2 <= vB2) { fix = 2; key = vA2; }
          else if (vB2 < -tol_s) { fix = 3; key = vB2; }
        }
      }
      const unsigned viol = (__ballot_sync(kFull, fix != 0) >> (16 * grp)) & 0xFFFu;
      // globally worst lane (used after the first passes, prevents cycling)
      const float keyf = (float)key;
      const float best = group_min(keyf);
      const unsigned tie = (__ballot_sync(kFull, fix != 0 && keyf == best) >> (16 * grp)) & 0xFFFu;
      bool start_ipm = false;
      if (pol_round && mode == kModePolish) {
        if (viol == 0u) {
          y = yp;
          mode = kModeDone;
          if (status == 2) status = 0;  // iteration limit hit but the polish verified the optimum
        } else {
          pass++;
          const bool give_up = first ? (pass > kPdasFirst) : (pass >= kPolishPasses);
          if (!give_up) {
            // repair: every violating lane moves during the first passes, then only the worst one
            const bool mine = (fix != 0) && (pass <= 2 || (tie != 0u && (__ffs(tie) - 1) == gl));
            if (mine) pat = (fix == 1) ? 0 : ((c == 0) ? 1 : (fix == 2 ? -1 : 1));
          } else if (first) {
            first = false;
            mode = kModeIpm;
            start_ipm = true;
            pat = 0;
          } else if (converged || status == 2) {
            // interior-point iterate is final but the polish could not certify an active set
            mode = kModeDone;
            if (status == 0) status = 3;
          } else {
            mode = kModeIpm;  // keep iterating, try again after the next iteration
          }
        }
      }
      // ---- a group starts its interior-point iteration: strictly feasible point y0 (every stance leg
      //      pushes c0 along its normal), multipliers centred at the gradient scale, residuals of the start
      if (__any_sync(kFull, start_ipm)) {
        const double y0 = (alive && c == 0) ? c0 : 0.0;
        if (var_lane) {
#pragma unroll
          for (int r = 0; r < 6; r++) ws.pc[grp][gl][r] = at[r] * y0;
        }
        __syncwarp();
        if (gl < 6) {
          double ay = 0.0;
#pragma unroll 2
          for (int l = 0; l < kVars; l++) ay += ws.pc[grp][l][gl];
          ws.t6[grp][gl] = prm.S[gl] * (ws.bw[grp][gl] - ay);   // S (b - A~ y0)
        }
        __syncwarp();
        double g0 = prm.W * y0;
#pragma unroll
        for (int r = 0; r < 6; r++) g0 = fma(-at[r], ws.t6[grp][r], g0);
        const double gmax = (double)fmaxf(1.f, group_max(var_lane ? fabsf((float)g0) : 0.f));
        double e0A, e0B;
        rows_apply(c0, 0.0, mu, c, e0A, e0B);
        if (c == 0) e0A -= prm.fmin;
        const double s0A = fmax(e0A, 1e-3 * c0), s0B = fmax(e0B, 1e-3 * c0);  // mu <= 0 would make friction rows non-positive
        const double l0A = alive ? gmax * fast_rcp(s0A) : 0.0, l0B = rowB ? gmax * fast_rcp(s0B) : 0.0;
        const double dtl = dt_lane(l0A, l0B, mu, c, l0);
        if (start_ipm) {
          sA = alive ? s0A : 1.0; lamA = l0A; rpA = alive ? s0A - e0A : 0.0;
          sB = rowB ? s0B : 1.0;  lamB = l0B; rpB = rowB ? s0B - e0B : 0.0;
          y = y0;
          rd = var_lane ? g0 - dtl : 0.0;
        }
        __syncwarp();
      }
    }
    // an interior-point group that asked for a polish switches now (its iterate stays untouched)
    if (mode == kModeIpm && want_polish) {
      want_polish = false;
      mode = kModePolish;
      pass = 0;
      pat = 0;
      if (alive) pat = (c == 0) ? (lamA > sA ? 1 : 0) : ((lamA > sA) ? -1 : ((lamB > sB) ? 1 : 0));
    }
  }

  // ---------------- outputs: forces in base frame, torques, net wrench, flags
  const bool solved = (status == 0 || status == 2 || status == 3);
  if (!solved) { y = 0.0; pat = 0; }
  // f_leg = y_n n + y_1 t1 + y_2 t2 : sum of the three lanes' contributions y * e
  double f[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const double w = (alive && solved) ? y * ws.tail[a][lane] : 0.0;
    f[a] = (gshfl(w, l0) + gshfl(w, l0 + 1)) + gshfl(w, l0 + 2);
  }
  const double fx = (c == 0) ? f[0] : (c == 1 ? f[1] : f[2]);
  if (var_lane) {
    // tau_c = column c of J dotted with (-f) + G_c (CFD.cpp:535-559); swing legs report zero
    const double tq = (alive && solved)
                          ? ws.tail[6][lane] - (ws.tail[3][lane] * f[0] + ws.tail[4][lane] * f[1] + ws.tail[5][lane] * f[2])
                          : 0.0;
    ws.out[kRowGrf + gl][qi] = fx;
    ws.out[kRowTau + gl][qi] = tq;
  }
  if (want_net) {
    // A x = sum over slots of column * y (CFD.cpp:614-625)
    if (var_lane) {
#pragma unroll
      for (int r = 0; r < 6; r++) ws.pc[grp][gl][r] = solved ? ws.atl[grp][gl][r] * y : 0.0;
    }
    __syncwarp();
    if (gl < 6) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < kVars; j++) acc += ws.pc[grp][j][gl];
      ws.out[kRowNet + gl][qi] = acc;
    }
    __syncwarp();
  }
  {
    unsigned bits = 0u;
    if (alive && solved && pat != 0) {
      const int row = (c == 0) ? 0 : (2 * c - 1 + (pat == 1 ? 1 : 0));  // reference row order: F_min, +t1, -t1, +t2, -t2
      bits = 1u << (4 + 5 * leg + row);
    }
    // OR over the group
    bits |= __shfl_xor_sync(kFull, bits, 1, kGroup);
    bits |= __shfl_xor_sync(kFull, bits, 2, kGroup);
    bits |= __shfl_xor_sync(kFull, bits, 4, kGroup);
    bits |= __shfl_xor_sync(kFull, bits, 8, kGroup);
    if (gl == 0) {
      const unsigned itc = it > 31 ? 31u : (unsigned)it;
      ws.flags[qi] = mask | bits | ((unsigned)status << 24) | (itc << 27);
    }
  }
}

// ---------------------------------------------------------------- kernel
template <int MODE>
__global__ void __launch_bounds__(kThreads, QLB_MIN_CTAS) qlb_solve_kernel(const SolveArgs a) {
  constexpr int ROWS = (MODE == 1) ? kInRowsState : kInRows;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CtaSmem<ROWS>& sm = *reinterpret_cast<CtaSmem<ROWS>*>(smem_raw);
  {
    const double* src = reinterpret_cast<const double*>(a.params);
    double* dst = reinterpret_cast<double*>(&sm.prm);
    for (int i = threadIdx.x; i < (int)(sizeof(DeviceParams) / 8); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#ifdef QLB_PIN_IDS
  asm volatile("" : "+r"(lane), "+r"(warp));  // keep the ids in registers instead of re-deriving them from S2R
#endif
  WarpSmem<ROWS>& ws = sm.w[warp];
  const unsigned long long nbatch = (a.B + kBatch - 1) / kBatch;
  const bool vec = a.vec_ok != 0;
  const bool have_mu = a.mu != nullptr, have_normals = a.normals != nullptr, want_net = a.netwrench != nullptr;

  for (;;) {
    unsigned long long bi = 0;
    if (lane == 0) bi = atomicAdd(a.counter, 1ull);
    bi = __shfl_sync(kFull, bi, 0);
    if (bi >= nbatch) break;
    const unsigned long long b0 = bi * kBatch;
    const int nvalid = (int)((a.B - b0 < (unsigned long long)kBatch) ? (a.B - b0) : kBatch);

    stage_in<ROWS>(&ws.in[kRowQ], a.q, 12, a.B, b0, nvalid, lane, vec);
    if (MODE == 1) {
      stage_in<ROWS>(&ws.in[kRowPose], a.pose, 7, a.B, b0, nvalid, lane, vec);
      stage_in<ROWS>(&ws.in[kRowTwist], a.twist, 6, a.B, b0, nvalid, lane, vec);
      stage_in<ROWS>(&ws.in[kRowTPose], a.tpose, 7, a.B, b0, nvalid, lane, vec);
      stage_in<ROWS>(&ws.in[kRowTTwist], a.ttwist, 6, a.B, b0, nvalid, lane, vec);
    } else {
      stage_in<ROWS>(&ws.in[kRowQuat], a.quat, 4, a.B, b0, nvalid, lane, vec);
      stage_in<ROWS>(&ws.in[kRowWrench], a.wrench, 6, a.B, b0, nvalid, lane, vec);
    }
    if (have_mu) stage_in<ROWS>(&ws.in[kRowMu], a.mu, 4, a.B, b0, nvalid, lane, vec);
    if (have_normals) stage_in<ROWS>(&ws.in[kRowNormal], a.normals, 12, a.B, b0, nvalid, lane, vec);
    if (lane < kBatch) ws.mask[lane] = (lane < nvalid) ? a.mask[b0 + lane] : (uint8_t)0;
    __syncwarp();

#pragma unroll 1
    for (int pair = 0; 2 * pair < nvalid; pair++)
      solve_group<MODE, ROWS>(ws, *a.model, sm.prm, 2 * pair + (lane >> 4), lane, have_mu, have_normals, want_net);
    __syncwarp();

    stage_out(a.grf, &ws.out[kRowGrf], 12, a.B, b0, nvalid, lane, vec);
    stage_out(a.tau, &ws.out[kRowTau], 12, a.B, b0, nvalid, lane, vec);
    if (want_net) stage_out(a.netwrench, &ws.out[kRowNet], 6, a.B, b0, nvalid, lane, vec);
    if (MODE == 1 && a.wrench_out) stage_out(a.wrench_out, &ws.out[kRowWout], 6, a.B, b0, nvalid, lane, vec);
    if (lane < nvalid) a.flags[b0 + lane] = ws.flags[lane];
    __syncwarp();
  }
}

}  // namespace qlb
