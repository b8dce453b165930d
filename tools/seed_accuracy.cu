// Accuracy of the FP64 seed instructions (MUFU.RCP64H / MUFU.RSQ64H) and of the refinements built on them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o quadruped_locomotion_b200/variants/seed_accuracy tools/seed_accuracy.cu
// Prints the maximum relative error over 2^24 arguments spread over 1e-6 .. 1e18 (the range of the KKT pivots and beyond).
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ double seed_rcp(double x) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
__device__ __forceinline__ double seed_rsq(double x) { double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }

__global__ void k(double* out, int n) {
  double m[6] = {0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double u = (i + 0.5) / n;
    const double x = exp2(-20.0 + 80.0 * u) * (1.0 + 0.61803398875 * ((i * 2654435761u) >> 8) / 16777216.0);
    const double rc = 1.0 / x, rs = 1.0 / sqrt(x);
    double r = seed_rcp(x);
    m[0] = fmax(m[0], fabs(r - rc) / rc);
    double a = r * fma(-x, r, 2.0); a = a * fma(-x, a, 2.0);
    m[1] = fmax(m[1], fabs(a - rc) / rc);
    const double e = fma(-x, r, 1.0);
    m[2] = fmax(m[2], fabs(fma(r * e, e + 1.0, r) - rc) / rc);
    r = seed_rsq(x);
    m[3] = fmax(m[3], fabs(r - rs) / rs);
    const double hx = 0.5 * x;
    a = r * fma(-hx * r, r, 1.5); a = a * fma(-hx * a, a, 1.5);
    m[4] = fmax(m[4], fabs(a - rs) / rs);
    const double e2 = fma(-x * r, r, 1.0);
    m[5] = fmax(m[5], fabs(fma(r * e2, fma(0.375, e2, 0.5), r) - rs) / rs);
  }
  for (int j = 0; j < 6; j++) {
    double v = m[j];
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned long long*>(out + j), (unsigned long long)__double_as_longlong(v));
  }
}

int main() {
  double* d; cudaMalloc(&d, 6 * sizeof(double)); cudaMemset(d, 0, 6 * sizeof(double));
  k<<<592, 256>>>(d, 1 << 24);
  double h[6]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  printf("{\"rcp_seed\": %.3e, \"rcp_two_newton\": %.3e, \"rcp_third_order\": %.3e, \"rsqrt_seed\": %.3e, \"rsqrt_two_newton\": %.3e, \"rsqrt_third_order\": %.3e}\n",
         h[0], h[1], h[2], h[3], h[4], h[5]);
  return cudaGetLastError() != cudaSuccess;
}
