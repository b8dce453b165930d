// qlb_device.cuh - device-side building blocks of the fused contact-force-distribution kernel.
//
// Mapping: one QP per HALF-WARP ("group" of 16 lanes, two QPs per warp).  Lane gl < 12 of a group owns
// variable gl = 3*leg + c of the 12-slot problem and row gl of every 12x12 matrix; lanes 12..15 carry
// zeros.  All linear algebra is register-resident; rows are exchanged with width-16 warp shuffles.
//
// The problem is solved in per-leg contact coordinates y = (y_n, y_1, y_2) = Q_leg^T f_leg with
// Q_leg = [n t1 t2] the friction-pyramid frame the reference builds in
// ContactForceDistribution.cpp:286-309.  In these coordinates the five rows of a leg
// (ContactForceDistribution.cpp:241-247,315-325) are
//     y_n >= F_min,  mu y_n + y_1 >= 0,  mu y_n - y_1 >= 0,  mu y_n + y_2 >= 0,  mu y_n - y_2 >= 0
// so the constraint matrix never has to be stored.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace qlb {

constexpr int kGroup = 16;        // lanes per QP
constexpr int kVars = 12;         // variable slots (4 legs x 3)
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------- shuffles / reductions in a group
// Always full-mask: callers keep the whole warp converged around every exchange (a partial, run-time
// mask makes nvcc emit WARPSYNC + a convergence barrier per shuffle - measured 30x slower).
__device__ __forceinline__ double gshfl(double v, int src) { return __shfl_sync(kFull, v, src, kGroup); }
__device__ __forceinline__ float gshfl(float v, int src) { return __shfl_sync(kFull, v, src, kGroup); }
__device__ __forceinline__ int gshfl(int v, int src) { return __shfl_sync(kFull, v, src, kGroup); }

// per-leg value (replicated on the three lanes of a leg) -> sum / max / min over the four legs
__device__ __forceinline__ double leg_sum(double v) {
  return (gshfl(v, 0) + gshfl(v, 3)) + (gshfl(v, 6) + gshfl(v, 9));
}
__device__ __forceinline__ float leg_max(float v) {
  return fmaxf(fmaxf(gshfl(v, 0), gshfl(v, 3)), fmaxf(gshfl(v, 6), gshfl(v, 9)));
}
__device__ __forceinline__ double leg_min(double v) {
  return fmin(fmin(gshfl(v, 0), gshfl(v, 3)), fmin(gshfl(v, 6), gshfl(v, 9)));
}
__device__ __forceinline__ float leg_sum(float v) {
  return (gshfl(v, 0) + gshfl(v, 3)) + (gshfl(v, 6) + gshfl(v, 9));
}
__device__ __forceinline__ float leg_min(float v) {
  return fminf(fminf(gshfl(v, 0), gshfl(v, 3)), fminf(gshfl(v, 6), gshfl(v, 9)));
}
// max over the 16 lanes of a group (butterfly)
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o, kGroup));
  return v;
}

// single MUFU.RCP (no denormal / range slow path)
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// 1/x to ~1e-13 relative: FP32 seed + two Newton steps in FP64 (x must be in FP32 normal range)
__device__ __forceinline__ double fast_rcp(double x) {
  double r = (double)rcp_approx((float)x);
  r = r * fma(-x, r, 2.0);
  return r * fma(-x, r, 2.0);
}

__device__ __forceinline__ float fast_rcp(float x) { return rcp_approx(x); }

// 1/sqrt(x) to FP64 rounding: FP32 seed (one MUFU.RSQ) + two Newton steps in FP64.  x must be a
// normal FP32-range number (pivots of the KKT matrices are 1e-4 .. 1e16); x <= 0 gives NaN.
__device__ __forceinline__ double fast_rsqrt(double x) {
  float xf = (float)x, rf;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(xf));
  double r = (double)rf;
  const double hx = 0.5 * x;
  r = r * fma(-hx * r, r, 1.5);
  return r * fma(-hx * r, r, 1.5);
}

// FP32 twin: MUFU.RSQ + one Newton step
__device__ __forceinline__ float fast_rsqrt(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * fmaf(-0.5f * x * r, r, 1.5f);
}

// sin and cos for joint angles (|x| up to ~1e4 rad; URDF limits are +-3 rad): two-constant Cody-Waite
// reduction by pi/2 and the fdlibm kernel polynomials.  No slow path, no local memory.
__device__ __forceinline__ void sincos_small(double x, double* sn, double* cs) {
  const double kd = rint(x * 6.36619772367581382433e-01);  // x * 2/pi
  const int k = (int)kd;
  double r = fma(-kd, 1.57079632673412561417e+00, x);
  r = fma(-kd, 6.07710050650619224932e-11, r);
  const double z = r * r;
  double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  ps = fma(z, ps, 2.75573137070700676789e-06);
  ps = fma(z, ps, -1.98412698298579493134e-04);
  ps = fma(z, ps, 8.33333333332248946124e-03);
  ps = fma(z, ps, -1.66666666666666324348e-01);
  const double sr = fma(z * r, ps, r);
  double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  pc = fma(z, pc, -2.75573143513906633035e-07);
  pc = fma(z, pc, 2.48015872894767294178e-05);
  pc = fma(z, pc, -1.38888888888741095749e-03);
  pc = fma(z, pc, 4.16666666666666019037e-02);
  const double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
  const double s0 = (k & 1) ? cr : sr;
  const double c0 = (k & 1) ? sr : cr;
  *sn = (k & 2) ? -s0 : s0;
  *cs = ((k + 1) & 2) ? -c0 : c0;
}

__device__ __forceinline__ void sincos_small(float x, float* sn, float* cs) { sincosf(x, sn, cs); }

// ---------------------------------------------------------------- 12x12 Cholesky through shuffles
// In : H[j] = entry (gl, j) of a symmetric positive definite matrix (full row; lanes >= 12: zeros).
// Out: H[j<gl] = L[gl][j], H[gl] = L[gl][gl], H[j>gl] = L[j][gl] (row gl of L^T), rdiag = 1/L[gl][gl].
// Right-looking; at step k the scaled column k is broadcast entry by entry and every lane below
// updates its whole trailing row, so the trailing matrix stays symmetric and lane k can keep the
// broadcast values as its row of L^T (needed by the backward substitution).
// Returns false (group-uniform) when a pivot is not positive.
__device__ __forceinline__ bool group_cholesky(double (&H)[kVars], double& rdiag, const int gl) {
  bool ok = true;
  rdiag = 1.0;
#pragma unroll
  for (int k = 0; k < kVars; k++) {
    const double dkk = gshfl(H[k], k);
    ok = ok && (dkk > 0.0);
    const double rinv = fast_rsqrt(dkk);
    if (gl >= k) H[k] *= rinv;
    if (gl == k) rdiag = rinv;
    const double a = (gl > k) ? -H[k] : 0.0;  // rows above k are finished: multiplier 0 leaves them alone
#pragma unroll
    for (int j = k + 1; j < kVars; j++) {
      const double v = gshfl(H[k], j);  // L[j][k]
      H[j] = fma(a, v, H[j]);
      if (gl == k) H[j] *= rinv;        // row k of L^T = own trailing row scaled (the trailing matrix is symmetric)
    }
  }
  return ok;
}

// Same factorisation with the column broadcast going through shared memory (one 8-byte store per lane,
// one __syncwarp and a handful of 16-byte broadcast loads per step instead of 2 x (11-k) shuffles) and
// the forward substitution L z = rhs fused in.  xb = this warp's exchange buffer [2][2][16] (double
// buffered by step parity, one row per group); lt = this group's 12 x 13 transposition buffer.
// In: acc = rhs[gl].  Out: acc = z[gl].
__device__ __forceinline__ bool group_cholesky_fwd(double (&H)[kVars], double& rdiag, double& acc,
                                                   double (*xb)[2][16], double* lt, const int grp, const int gl) {
  bool ok = true;
  rdiag = 1.0;
#pragma unroll
  for (int k = 0; k < kVars; k++) {
    double* buf = xb[k & 1][grp];
    if (gl >= k && gl < kVars) buf[gl] = H[k];
    if (gl == k) buf[12] = acc;
    __syncwarp();
    double bcol[kVars];
#pragma unroll
    for (int p2 = 0; p2 < kVars / 2; p2++) {
      if (2 * p2 + 1 >= k) {  // 16-byte broadcast loads of the part of the column that is still needed
        const double2 t = *reinterpret_cast<const double2*>(buf + 2 * p2);
        bcol[2 * p2] = t.x;
        bcol[2 * p2 + 1] = t.y;
      }
    }
    const double dkk = bcol[k];
    const double rk = buf[12];
    ok = ok && (dkk > 0.0);
    const double rinv = fast_rsqrt(dkk);
    const double zk = rk * rinv;
    const double a = (gl > k) ? -H[k] * (rinv * rinv) : 0.0;  // rows above k are finished
    if (gl >= k) H[k] *= rinv;                                // L[i][k]
    if (gl == k) { rdiag = rinv; acc = zk; }
    if (gl > k) acc = fma(-H[k], zk, acc);
#pragma unroll
    for (int j = k + 1; j < kVars; j++) {
      const double bj = bcol[j];           // raw H[j][k]
      H[j] = fma(a, bj, H[j]);             // H[i][j] -= H[i][k] H[j][k] / d
    }
  }
  // rows of L^T for the backward substitution: transpose L through shared memory
  // (row pitch 13 doubles: both the row writes and the column reads are bank-conflict free)
#pragma unroll
  for (int j = 0; j < kVars; j++)
    if (gl < kVars) lt[gl * 13 + j] = H[j];
  __syncwarp();
#pragma unroll
  for (int j = 1; j < kVars; j++)
    if (gl < j) H[j] = lt[j * 13 + gl];
  __syncwarp();
  return ok;
}

// forward substitution L z = b (shuffles): lane gl passes b[gl], receives z[gl]
__device__ __forceinline__ double group_forward(const double (&H)[kVars], const double rdiag, const double b,
                                                const int gl) {
  double acc = b;
#pragma unroll
  for (int j = 0; j < kVars; j++) {
    const double zj = gshfl(acc * rdiag, j);
    if (gl > j) acc = fma(-H[j], zj, acc);
  }
  return acc * rdiag;
}
// backward substitution L^T x = z (shuffles)
__device__ __forceinline__ double group_backward(const double (&H)[kVars], const double rdiag, const double z,
                                                 const int gl) {
  double acc = z;
#pragma unroll
  for (int j = kVars - 1; j >= 0; j--) {
    const double xj = gshfl(acc * rdiag, j);
    if (gl < j) acc = fma(-H[j], xj, acc);
  }
  return acc * rdiag;
}

// ---------------------------------------------------------------- rolled variants, matrix in shared memory
// The unrolled register-resident routines above are ~1.5k instructions per round; with a dozen warps
// per SM at different program counters the kernel then streams its instructions from L2 (ncu:
// stall_no_instruction > 50 % of all samples, profiles/r1_v4_ncu_summary.txt).  These versions keep row
// gl of the matrix in shared memory, hs[j * kPitch + lane] = entry (gl, j), and are plain loops: a few
// dozen instructions that stay in the instruction cache.  kPitch = 33 makes both the row accesses
// (lane varies) and the transposed accesses of the backward substitution bank-conflict free.
constexpr int kPitch = 33;

// Cholesky with the forward substitution of one right-hand side fused in.
// In: acc = rhs[gl].  Out: hs = L (entry (gl, j), j <= gl), rdiag = 1/L[gl][gl], acc = z[gl] with L z = rhs.
template <int N>
__device__ __forceinline__ bool smem_cholesky_fwd(double* __restrict__ hs, double (*xb)[2][16], double& rdiag,
                                                  double& acc, const int grp, const int gl, const int lane) {
  bool ok = true;
  rdiag = 1.0;
#pragma unroll 1
  for (int k = 0; k < N; k++) {
    double* buf = xb[k & 1][grp];
    const bool row = gl < N;                   // lanes beyond the matrix carry zeros
    const double hk = row ? hs[k * kPitch + lane] : 0.0;
    if (gl >= k && row) buf[gl] = hk;          // publish the raw column k
    if (gl == k) buf[12] = acc;                // and the pivot row's right-hand side
    __syncwarp();
    const double dkk = buf[k];
    const double rk = buf[12];
    ok = ok && (dkk > 0.0);
    const double rinv = fast_rsqrt(dkk);
    const double zk = rk * rinv;
    const double lik = hk * rinv;
    const double a = (gl > k) ? -lik * rinv : 0.0;  // finished rows: multiplier 0
    if (gl >= k && row) hs[k * kPitch + lane] = lik;
    if (gl == k) { rdiag = rinv; acc = zk; }
    if (gl > k) acc = fma(-lik, zk, acc);
    if (row) {
#pragma unroll 4
      for (int j = k + 1; j < N; j++) hs[j * kPitch + lane] = fma(a, buf[j], hs[j * kPitch + lane]);
    }
  }
  return ok;
}

// 6x6 variant with the row in registers and the step loop unrolled (15 multiply-adds in all, ~170
// instructions): the matrix is read from hs once, the factor written back for the substitutions.
__device__ __forceinline__ bool reg_cholesky6_fwd(double* __restrict__ hs, double (*xb)[2][16], double& rdiag,
                                                  double& acc, const int grp, const int gl, const int lane) {
  const bool row = gl < 6;
  double h[6];
#pragma unroll
  for (int j = 0; j < 6; j++) h[j] = row ? hs[j * kPitch + lane] : 0.0;
  bool ok = true;
  rdiag = 1.0;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double* buf = xb[k & 1][grp];
    if (gl >= k && row) buf[gl] = h[k];  // publish the raw column k
    if (gl == k) buf[6] = acc;           // and the pivot row's right-hand side
    __syncwarp();
    double bcol[8];
#pragma unroll
    for (int p2 = 0; p2 < 4; p2++) {
      if (2 * p2 + 1 >= k) {             // 16-byte broadcast loads of what is still needed (pair 3 = rhs slot)
        const double2 t = *reinterpret_cast<const double2*>(buf + 2 * p2);
        bcol[2 * p2] = t.x;
        bcol[2 * p2 + 1] = t.y;
      }
    }
    const double dkk = bcol[k];
    ok = ok && (dkk > 0.0);
    const double rinv = fast_rsqrt(dkk);
    const double zk = bcol[6] * rinv;
    const double lik = h[k] * rinv;
    const double a = (gl > k) ? -lik * rinv : 0.0;  // finished rows: multiplier 0
    if (gl >= k) h[k] = lik;
    if (gl == k) { rdiag = rinv; acc = zk; }
    if (gl > k) acc = fma(-lik, zk, acc);
#pragma unroll
    for (int j = k + 1; j < 6; j++) h[j] = fma(a, bcol[j], h[j]);
  }
  if (row) {
#pragma unroll
    for (int j = 0; j < 6; j++) hs[j * kPitch + lane] = h[j];
  }
  __syncwarp();
  return ok;
}

// forward substitution L z = b: lane gl passes b[gl], receives z[gl]
template <int N>
__device__ __forceinline__ double smem_forward(const double* __restrict__ hs, double (*xb)[2][16], const double rdiag,
                                               const double b, const int grp, const int gl, const int lane) {
  double acc = b, z = 0.0;
#pragma unroll 1
  for (int j = 0; j < N; j++) {
    double* buf = xb[j & 1][grp];
    if (gl == j) buf[13] = acc * rdiag;
    __syncwarp();
    const double zj = buf[13];
    if (gl == j) z = zj;
    if (gl > j && gl < N) acc = fma(-hs[j * kPitch + lane], zj, acc);
  }
  return z;
}

// backward substitution L^T x = z; L[j][gl] is read transposed from row j's lane
template <int N>
__device__ __forceinline__ double smem_backward(const double* __restrict__ hs, double (*xb)[2][16], const double rdiag,
                                                const double z, const int grp, const int gl) {
  double acc = z, x = 0.0;
#pragma unroll 1
  for (int j = N - 1; j >= 0; j--) {
    double* buf = xb[j & 1][grp];
    if (gl == j) buf[13] = acc * rdiag;
    __syncwarp();
    const double xj = buf[13];
    if (gl == j) x = xj;
    if (gl < j) acc = fma(-hs[gl * kPitch + 16 * grp + j], xj, acc);  // gl < j < N: in range
  }
  return x;
}

// ---------------------------------------------------------------- model / parameters in device memory
// T = double for the FP64 entry points, float for their _f32 twins (the context keeps both copies).
template <typename T>
struct DeviceModelT {
  T rot[4][4][9];   // [leg][joint] rotation of <origin rpy>, row-major
  T xyz[4][4][3];   // [leg][joint] <origin xyz>
  T mass[4][4];     // link masses
  T com[4][4][3];   // link COM in link frame (the foot link's is pre-rotated by rot[leg][3])
  T msuf[4][4];     // suffix sums of mass: msuf[leg][c] = sum_{l>=c} mass[leg][l]
};
using DeviceModel = DeviceModelT<double>;

template <typename T>
struct DeviceParamsT {
  T S[6];
  T W, fmin, mu_default, gravity;
  T tol;
  int max_iter;
  int pad;
  // virtual model controller
  T kp_t[3], kd_t[3], kff_t[3], kp_r[3], kd_r[3], kff_r[3];
  T torso_mass, leg_mass[4], leg_pos[4][3], com[3], grav_pct;
};
using DeviceParams = DeviceParamsT<double>;

}  // namespace qlb
