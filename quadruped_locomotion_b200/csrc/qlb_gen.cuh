// qlb_gen.cuh - device-side generator of the synthetic robot states: the same counter-based
// streams as quadruped_locomotion_b200/synth.py, BIT-IDENTICAL to it.  Every value is a pure function of
// (seed, stream, instance index); the elementary functions (sin / cos / log) are built from IEEE-exact operations
// only and every multiply and add is rounded separately (__dmul_rn / __dadd_rn are never contracted into an FMA),
// in the operation order of the numpy code.  One thread per state, coalesced SoA stores.
//
// This is what lets a Monte-Carlo sweep (BASELINE config C5, the consumer shape of
// free_gait_core/src/executor/BatchExecutor.cpp:40-83) run without staging its inputs from the host.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace qlb {

struct GenArgs {
  unsigned long long B, start, seed;
  int config;        // 1, 2, 3 (also C4), 5
  double* q; double* quat; double* wrench; uint8_t* mask; double* mu; double* normals;          // FP64 outputs (any may be null)
  float* q32; float* quat32; float* wrench32; float* mu32; float* normals32;                      // or their FP32 twins
};

__device__ __forceinline__ unsigned long long gen_mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned long long gen_bits(unsigned long long seed, unsigned stream, unsigned long long idx) {
  const unsigned long long key = gen_mix64(seed + (unsigned long long)stream * 0xD1B54A32D192ED03ull);
  return gen_mix64(key ^ gen_mix64(idx));
}
__device__ __forceinline__ double gen_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double gen_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double gen_sub(double a, double b) { return __dsub_rn(a, b); }

// lo + (hi - lo) * u,  u = top 53 bits / 2^53
__device__ __forceinline__ double gen_uniform(unsigned long long seed, unsigned stream, unsigned long long idx, double lo, double span) {
  const double u = gen_mul((double)(gen_bits(seed, stream, idx) >> 11), 1.0 / 9007199254740992.0);
  return gen_add(lo, gen_mul(span, u));
}

__device__ __forceinline__ void gen_sincos(double x, double* sn, double* cs) {
  const double kd = rint(gen_mul(x, 6.36619772367581382433e-01));
  const long long k = (long long)kd;
  const double r = gen_sub(gen_sub(x, gen_mul(kd, 1.57079632673412561417e+00)), gen_mul(kd, 6.07710050650619224932e-11));
  const double z = gen_mul(r, r);
  double ps = 1.58969099521155010221e-10;
  ps = gen_add(gen_mul(ps, z), -2.50507602534068634195e-08);
  ps = gen_add(gen_mul(ps, z), 2.75573137070700676789e-06);
  ps = gen_add(gen_mul(ps, z), -1.98412698298579493134e-04);
  ps = gen_add(gen_mul(ps, z), 8.33333333332248946124e-03);
  ps = gen_add(gen_mul(ps, z), -1.66666666666666324348e-01);
  const double sr = gen_add(r, gen_mul(gen_mul(z, r), ps));
  double pc = -1.13596475577881948265e-11;
  pc = gen_add(gen_mul(pc, z), 2.08757232129817482790e-09);
  pc = gen_add(gen_mul(pc, z), -2.75573143513906633035e-07);
  pc = gen_add(gen_mul(pc, z), 2.48015872894767294178e-05);
  pc = gen_add(gen_mul(pc, z), -1.38888888888741095749e-03);
  pc = gen_add(gen_mul(pc, z), 4.16666666666666019037e-02);
  const double cr = gen_add(gen_sub(1.0, gen_mul(0.5, z)), gen_mul(gen_mul(z, z), pc));
  const bool odd = (k & 1) != 0;
  const double s0 = odd ? cr : sr, c0 = odd ? sr : cr;
  *sn = (k & 2) ? -s0 : s0;
  *cs = ((k + 1) & 2) ? -c0 : c0;
}

// log x, 0 < x <= 1
__device__ __forceinline__ double gen_log(double x) {
  int e;
  double m = frexp(x, &e);
  const bool small = m < 0.70710678118654752440;
  if (small) { m = gen_mul(m, 2.0); e -= 1; }
  const double ed = (double)e;
  const double s = __ddiv_rn(gen_sub(m, 1.0), gen_add(m, 1.0));
  const double z = gen_mul(s, s);
  double p = 1.0 / 21.0;
  p = gen_add(gen_mul(p, z), 1.0 / 19.0);
  p = gen_add(gen_mul(p, z), 1.0 / 17.0);
  p = gen_add(gen_mul(p, z), 1.0 / 15.0);
  p = gen_add(gen_mul(p, z), 1.0 / 13.0);
  p = gen_add(gen_mul(p, z), 1.0 / 11.0);
  p = gen_add(gen_mul(p, z), 1.0 / 9.0);
  p = gen_add(gen_mul(p, z), 1.0 / 7.0);
  p = gen_add(gen_mul(p, z), 1.0 / 5.0);
  p = gen_add(gen_mul(p, z), 1.0 / 3.0);
  p = gen_add(gen_mul(p, z), 1.0 / 1.0);
  return gen_add(gen_mul(ed, 6.93147180369123816490e-01), gen_add(gen_mul(ed, 1.90821492927058770002e-10), gen_mul(gen_mul(2.0, s), p)));
}

// Box-Muller: (sigma * sqrt(-2 log(1 - u1))) * cos(2 pi u2)
__device__ __forceinline__ double gen_normal(unsigned long long seed, unsigned stream, unsigned long long idx, double sigma) {
  const double u1 = gen_uniform(seed, 2u * stream + 1000u, idx, 0.0, 1.0);
  const double u2 = gen_uniform(seed, 2u * stream + 1001u, idx, 0.0, 1.0);
  double sn, cs;
  gen_sincos(gen_mul(2.0 * 3.141592653589793, u2), &sn, &cs);
  return gen_mul(gen_mul(sigma, __dsqrt_rn(gen_mul(-2.0, gen_log(gen_sub(1.0, u1))))), cs);
}

template <typename T>
__device__ __forceinline__ void gen_store(double* p64, T* p32, unsigned long long B, int row, unsigned long long i, double v) {
  if (p64) p64[(size_t)row * B + i] = v;
  if (p32) p32[(size_t)row * B + i] = (T)v;
}

__global__ void __launch_bounds__(256) qlb_generate_kernel(const GenArgs a) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B) return;
  const unsigned long long B = a.B, idx = a.start + i, seed = a.seed;
  const int cfg = a.config;
  const double pi = 3.141592653589793;
  double q[12], quat[4], w[6], mu[4];
  unsigned mask;
  if (cfg == 1) {
    const double q1[12] = {0.0, 0.7, -1.4, 0.0, -0.7, 1.4, 0.0, 0.7, -1.4, 0.0, -0.7, 1.4};
#pragma unroll
    for (int j = 0; j < 12; j++) q[j] = q1[j];
    quat[0] = 1.0; quat[1] = quat[2] = quat[3] = 0.0;
#pragma unroll
    for (int r = 0; r < 6; r++) w[r] = 0.0;
    w[2] = 51.0 * 9.8;
#pragma unroll
    for (int l = 0; l < 4; l++) mu[l] = 0.6;
    mask = 0xFu;
  } else {
    const unsigned long long nominal = (cfg == 5) ? (idx >> 10) : idx;   // C5: 2^10 perturbations per nominal state
#pragma unroll
    for (int leg = 0; leg < 4; leg++) {
      const double s = (leg == 0 || leg == 2) ? 1.0 : -1.0;   // knee sign LF, RF, RH, LH
      q[3 * leg + 0] = gen_uniform(seed, 10 + 3 * leg, nominal, -0.25, 0.25 - (-0.25));
      q[3 * leg + 1] = gen_mul(s, gen_add(0.7, gen_uniform(seed, 11 + 3 * leg, nominal, -0.3, 0.3 - (-0.3))));
      q[3 * leg + 2] = gen_mul(-s, gen_add(1.4, gen_uniform(seed, 12 + 3 * leg, nominal, -0.4, 0.4 - (-0.4))));
    }
    double yaw = gen_uniform(seed, 1, nominal, -pi, pi - (-pi));
    double pitch = gen_uniform(seed, 2, nominal, -0.25, 0.25 - (-0.25));
    double roll = gen_uniform(seed, 3, nominal, -0.25, 0.25 - (-0.25));
    double scale = 1.0;
    if (cfg == 5) {
      yaw = gen_add(yaw, gen_normal(seed, 40, idx, 0.3));
      pitch = gen_add(pitch, gen_normal(seed, 41, idx, 0.1));
      roll = gen_add(roll, gen_normal(seed, 42, idx, 0.1));
#pragma unroll
      for (int l = 0; l < 4; l++) mu[l] = gen_uniform(seed, 50 + l, idx, 0.2, 1.0 - 0.2);
      scale = gen_uniform(seed, 60, idx, 0.8, 1.2 - 0.8);
    } else {
#pragma unroll
      for (int l = 0; l < 4; l++) mu[l] = 0.6;
    }
    double sy, cy, sp, cp, sr, cr;
    gen_sincos(gen_mul(0.5, yaw), &sy, &cy);
    gen_sincos(gen_mul(0.5, pitch), &sp, &cp);
    gen_sincos(gen_mul(0.5, roll), &sr, &cr);
    const double qw = gen_add(gen_mul(gen_mul(cr, cp), cy), gen_mul(gen_mul(sr, sp), sy));
    const double qx = gen_sub(gen_mul(gen_mul(sr, cp), cy), gen_mul(gen_mul(cr, sp), sy));
    const double qy = gen_add(gen_mul(gen_mul(cr, sp), cy), gen_mul(gen_mul(sr, cp), sy));
    const double qz = gen_sub(gen_mul(gen_mul(cr, cp), sy), gen_mul(gen_mul(sr, sp), cy));
    quat[0] = qw; quat[1] = qx; quat[2] = qy; quat[3] = qz;
    // third row of the rotation matrix base->world
    const double r20 = gen_mul(2.0, gen_sub(gen_mul(qx, qz), gen_mul(qw, qy)));
    const double r21 = gen_mul(2.0, gen_add(gen_mul(qy, qz), gen_mul(qw, qx)));
    const double r22 = gen_add(gen_sub(gen_sub(gen_mul(qw, qw), gen_mul(qx, qx)), gen_mul(qy, qy)), gen_mul(qz, qz));
    double weight = 51.0 * 9.8;
    if (cfg == 5) weight = gen_mul(weight, scale);
    const double r2[3] = {r20, r21, r22};
#pragma unroll
    for (int c = 0; c < 3; c++) {
      w[c] = gen_add(gen_mul(r2[c], weight), gen_normal(seed, 30 + c, nominal, 60.0));
      w[3 + c] = gen_normal(seed, 33 + c, nominal, 25.0);
    }
    if (cfg == 2) {
      mask = (idx % 2ull == 0ull) ? ((1u << 1) | (1u << 3)) : ((1u << 0) | (1u << 2));
    } else {
      const double sel = gen_uniform(seed, 4, nominal, 0.0, 1.0);
      const long long pick = (long long)(gen_bits(seed, 5, nominal) >> 40);
      const unsigned diag = (pick % 2 == 0) ? ((1u << 1) | (1u << 3)) : ((1u << 0) | (1u << 2));
      const unsigned three = 0xFu & ~(1u << (unsigned)(pick % 4));
      mask = (sel < 0.60) ? 0xFu : ((sel < 0.85) ? diag : three);
    }
  }
#pragma unroll
  for (int j = 0; j < 12; j++) gen_store(a.q, a.q32, B, j, i, q[j]);
#pragma unroll
  for (int r = 0; r < 4; r++) gen_store(a.quat, a.quat32, B, r, i, quat[r]);
#pragma unroll
  for (int r = 0; r < 6; r++) gen_store(a.wrench, a.wrench32, B, r, i, w[r]);
#pragma unroll
  for (int l = 0; l < 4; l++) gen_store(a.mu, a.mu32, B, l, i, mu[l]);
#pragma unroll
  for (int j = 0; j < 12; j++) gen_store(a.normals, a.normals32, B, j, i, (j % 3 == 2) ? 1.0 : 0.0);
  if (a.mask) a.mask[i] = (uint8_t)mask;
}

}  // namespace qlb
