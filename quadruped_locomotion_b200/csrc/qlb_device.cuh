// qlb_device.cuh - device-side building blocks of the fused contact-force-distribution kernels: fast
// reciprocal / reciprocal square root / sincos, and the model and parameter blocks.
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

constexpr unsigned kFull = 0xffffffffu;

// single MUFU.RCP (no denormal / range slow path)
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// 1/x to ~1e-13 relative: the FP64 seed instruction (MUFU.RCP64H, works on the high word: no conversion to FP32 and
// back - two quarter-rate instructions less in every dependent chain, 2.6 % of the fused kernel's time - and no FP32
// range to respect) + two Newton steps.  Measured on B200 over 2^24 arguments in 1e-6 .. 1e18 (tools/seed_accuracy.cu):
// seed 9.9e-7 relative (1/sqrt: 9.2e-7), after the two steps 2.2e-16 (1/sqrt: 3.2e-16).
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = r * fma(-x, r, 2.0);
  return r * fma(-x, r, 2.0);
}

__device__ __forceinline__ float fast_rcp(float x) { return rcp_approx(x); }
// 1/x to FP64 rounding (three Newton steps): for pivots whose error would be amplified by cancellation
__device__ __forceinline__ double full_rcp(double x) {
  double r = fast_rcp(x);
  return r * fma(-x, r, 2.0);
}
__device__ __forceinline__ float full_rcp(float x) { return rcp_approx(x); }

// (A single third-order step - r (1 + e/2 + 3 e^2/8), e = 1 - x r^2, five operations instead of seven, and the same for
// the reciprocal - is as accurate (2.7e-16 / 2.2e-16) and was measured twice: with the FP32 seeds the shorter code made
// the register allocator spill two values in the fused kernel, 0.510 ms against 0.490 ms; with the FP64 seeds nothing
// spills and it is 0.3 % faster, within the noise.  Kept as two Newton steps.)
// 1/sqrt(x) to FP64 rounding: FP64 seed (one MUFU.RSQ64H) + two Newton steps.  Any normal FP64 x > 0 (pivots of
// the KKT matrices are 1e-4 .. 1e16); x <= 0 gives NaN.
__device__ __forceinline__ double fast_rsqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
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
  // nearest multiple of pi/2 by the add-magic-constant trick: the low word of x 2/pi + 1.5 2^52 is the integer k
  // (no FP64 round / convert instructions, which run at a fraction of the FMA rate)
  const double km = fma(x, 6.36619772367581382433e-01, 6755399441055744.0);
  const int k = __double2loint(km);
  const double kd = km - 6755399441055744.0;
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
  // quadrant signs: flip the sign bit in the high word (one logic operation each)
  *sn = __hiloint2double(__double2hiint(s0) ^ ((k & 2) << 30), __double2loint(s0));
  *cs = __hiloint2double(__double2hiint(c0) ^ (((k + 1) & 2) << 30), __double2loint(c0));
}

// FP32 twin (|x| <= 1e6, the bound the callers check): three-constant Cody-Waite reduction, minimax polynomials on
// [-pi/4, pi/4]; ~1 ulp for joint-range arguments.  No slow path, no local memory (sincosf carries a Payne-Hanek
// branch with a stack frame).
__device__ __forceinline__ void sincos_small(float x, float* sn, float* cs) {
  const float km = fmaf(x, 6.36619772e-01f, 12582912.0f);   // 1.5 2^23: the low bits of the sum hold the integer k
  const int k = __float_as_int(km);
  const float kf = km - 12582912.0f;
  float r = fmaf(-kf, 1.57079601e+00f, x);
  r = fmaf(-kf, 3.13916473e-07f, r);
  r = fmaf(-kf, 5.39030253e-15f, r);
  const float z = r * r;
  float ps = fmaf(z, -1.95152959e-04f, 8.33216087e-03f);
  ps = fmaf(z, ps, -1.66666546e-01f);
  const float sr = fmaf(z * r, ps, r);
  float pc = fmaf(z, 2.44331571e-05f, -1.38873163e-03f);
  pc = fmaf(z, pc, 4.16666457e-02f);
  const float cr = fmaf(z * z, pc, fmaf(-0.5f, z, 1.0f));
  const float s0 = (k & 1) ? cr : sr;
  const float c0 = (k & 1) ? sr : cr;
  *sn = __int_as_float(__float_as_int(s0) ^ ((k & 2) << 30));
  *cs = __int_as_float(__float_as_int(c0) ^ (((k + 1) & 2) << 30));
}

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
