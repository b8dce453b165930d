// qlb_records.cuh - array-of-structs face of the wrench-mode solve: qlb_wrench_record[] in, qlb_result_record[] out.
// A host caller then moves ONE contiguous block per direction over PCIe (large 1-D copies run at 96 GB/s in both
// directions together on this box, the pitched 2-D copies of SoA column ranges at 73 GB/s); the transposition
// between records and the solver's SoA arrays happens on the device, where it is HBM-bound byte movement:
// a CTA moves 128 records through a padded shared-memory tile, 8-byte coalesced on both sides.
#pragma once

#include "qlb.h"

namespace qlb {

constexpr int kRecTile = 128;
constexpr int kInWords = (int)(sizeof(qlb_wrench_record) / 8);    // 27
constexpr int kOutWords = (int)(sizeof(qlb_result_record) / 8);   // 31
static_assert(sizeof(qlb_wrench_record) == 216 && sizeof(qlb_result_record) == 248, "record layouts are part of the ABI");

// records -> q[12][B], quat[4][B], wrench[6][B], mu[4][B], mask[B]
__global__ void __launch_bounds__(kRecTile) qlb_unpack_records_kernel(unsigned long long B, const qlb_wrench_record* __restrict__ rec,
                                                                      double* __restrict__ q, double* __restrict__ quat,
                                                                      double* __restrict__ wrench, double* __restrict__ mu,
                                                                      uint8_t* __restrict__ mask) {
  __shared__ double tile[kRecTile * (kInWords + 1)];   // row stride 28 words: a column walk hits 32 different banks
  const unsigned long long base = (unsigned long long)blockIdx.x * kRecTile;
  const unsigned long long left = B - base;
  const int n = left < (unsigned long long)kRecTile ? (int)left : kRecTile;
  const double* src = reinterpret_cast<const double*>(rec + base);
  for (int v = threadIdx.x; v < n * kInWords; v += kRecTile) {
    const int r = v / kInWords, c = v - r * kInWords;
    tile[r * (kInWords + 1) + c] = __ldg(src + v);
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t >= n) return;
  const double* r = tile + t * (kInWords + 1);
  const unsigned long long i = base + t;
#pragma unroll
  for (int a = 0; a < 12; a++) q[(size_t)a * B + i] = r[a];
#pragma unroll
  for (int a = 0; a < 4; a++) quat[(size_t)a * B + i] = r[12 + a];
#pragma unroll
  for (int a = 0; a < 6; a++) wrench[(size_t)a * B + i] = r[16 + a];
#pragma unroll
  for (int a = 0; a < 4; a++) mu[(size_t)a * B + i] = r[22 + a];
  mask[i] = (uint8_t)((unsigned long long)__double_as_longlong(r[26]) & 0xFull);
}

// grf[12][B], tau[12][B], netwrench[6][B], flags[B] -> records
__global__ void __launch_bounds__(kRecTile) qlb_pack_results_kernel(unsigned long long B, const double* __restrict__ grf,
                                                                    const double* __restrict__ tau, const double* __restrict__ net,
                                                                    const uint32_t* __restrict__ flags, qlb_result_record* __restrict__ rec) {
  __shared__ double tile[kRecTile * (kOutWords + 2)];   // row stride 33 words
  const unsigned long long base = (unsigned long long)blockIdx.x * kRecTile;
  const unsigned long long left = B - base;
  const int n = left < (unsigned long long)kRecTile ? (int)left : kRecTile;
  const int t = threadIdx.x;
  if (t < n) {
    double* r = tile + t * (kOutWords + 2);
    const unsigned long long i = base + t;
#pragma unroll
    for (int a = 0; a < 12; a++) r[a] = grf[(size_t)a * B + i];
#pragma unroll
    for (int a = 0; a < 12; a++) r[12 + a] = tau[(size_t)a * B + i];
#pragma unroll
    for (int a = 0; a < 6; a++) r[24 + a] = net[(size_t)a * B + i];
    r[30] = __longlong_as_double((long long)(unsigned long long)flags[i]);   // flags in the low word, reserved word zero
  }
  __syncthreads();
  double* dst = reinterpret_cast<double*>(rec + base);
  for (int v = threadIdx.x; v < n * kOutWords; v += kRecTile) {
    const int r = v / kOutWords, c = v - r * kOutWords;
    dst[v] = tile[r * (kOutWords + 2) + c];
  }
}

// preview of a planned motion: SoA results -> qlb_preview_record[]
constexpr int kPrevWords = (int)(sizeof(qlb_preview_record) / 8);   // 51
constexpr int kPrevTile = 64;   // 64 records x 53 words = 27 KB of shared memory
static_assert(sizeof(qlb_preview_record) == 408, "record layout is part of the ABI");
__global__ void __launch_bounds__(kPrevTile) qlb_pack_preview_kernel(unsigned long long B, const double* __restrict__ feet,
                                                                    const double* __restrict__ grf, const double* __restrict__ tau,
                                                                    const double* __restrict__ net, const double* __restrict__ wrench,
                                                                    const double* __restrict__ margin, const double* __restrict__ minn,
                                                                    const uint32_t* __restrict__ flags, qlb_preview_record* __restrict__ rec) {
  __shared__ double tile[kPrevTile * (kPrevWords + 2)];   // row stride 53 words
  const unsigned long long base = (unsigned long long)blockIdx.x * kPrevTile;
  const unsigned long long left = B - base;
  const int n = left < (unsigned long long)kPrevTile ? (int)left : kPrevTile;
  const int t = threadIdx.x;
  if (t < n) {
    double* r = tile + t * (kPrevWords + 2);
    const unsigned long long i = base + t;
#pragma unroll
    for (int a = 0; a < 12; a++) { r[a] = feet[(size_t)a * B + i]; r[12 + a] = grf[(size_t)a * B + i]; r[24 + a] = tau[(size_t)a * B + i]; }
#pragma unroll
    for (int a = 0; a < 6; a++) { r[36 + a] = net[(size_t)a * B + i]; r[42 + a] = wrench[(size_t)a * B + i]; }
    r[48] = margin[i]; r[49] = minn[i];
    r[50] = __longlong_as_double((long long)(unsigned long long)flags[i]);
  }
  __syncthreads();
  double* dst = reinterpret_cast<double*>(rec + base);
  for (int v = threadIdx.x; v < n * kPrevWords; v += kPrevTile) {
    const int r = v / kPrevWords, c = v - r * kPrevWords;
    dst[v] = tile[r * (kPrevWords + 2) + c];
  }
}

// one value per leg for every state of a batch (the friction coefficient of a preview)
__global__ void __launch_bounds__(256) qlb_fill_rows_kernel(unsigned long long B, int rows, double* __restrict__ dst, double v0, double v1,
                                                            double v2, double v3) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const double v[4] = {v0, v1, v2, v3};
  for (int r = 0; r < rows; r++) dst[(size_t)r * B + i] = v[r & 3];
}

}  // namespace qlb
