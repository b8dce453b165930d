// nccl_demo.cpp - the multi-GPU shape of the path from a C++ host: one context per GPU, every GPU
// solves its own contiguous slice of a generated batch (no inter-GPU traffic on the solve path), then ONE
// collective: qlb_stats_allreduce over an NCCL communicator.  One process, one thread per GPU.
//   nccl_demo [states_per_gpu]      prints the reduced statistics; exit code 0 when every rank agrees
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "qlb.h"
#include "qlb_models.h"

int main(int argc, char** argv) {
  const size_t B = argc > 1 ? (size_t)atoll(argv[1]) : 65536;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { std::fprintf(stderr, "no CUDA device\n"); return 2; }
  std::vector<int> devs(ndev);
  for (int i = 0; i < ndev; i++) devs[i] = i;
  std::vector<ncclComm_t> comms(ndev);
  if (ncclCommInitAll(comms.data(), ndev, devs.data()) != ncclSuccess) { std::fprintf(stderr, "ncclCommInitAll failed\n"); return 2; }
  std::vector<qlb_stats> stats(ndev);
  std::vector<int> rc(ndev, 0);
  std::vector<std::thread> th;
  for (int r = 0; r < ndev; r++) {
    th.emplace_back([&, r]() {
      cudaSetDevice(r);
      qlb_context* ctx = nullptr;
      if ((rc[r] = qlb_create(&ctx, QLB_MODEL_QUADRUPED_MODEL, nullptr, r, 0)) != QLB_OK) return;
      double *q, *quat, *wrench, *mu, *grf, *tau, *net;
      uint8_t* mask; uint32_t* flags;
      cudaMalloc(&q, 12 * B * 8); cudaMalloc(&quat, 4 * B * 8); cudaMalloc(&wrench, 6 * B * 8); cudaMalloc(&mu, 4 * B * 8);
      cudaMalloc(&grf, 12 * B * 8); cudaMalloc(&tau, 12 * B * 8); cudaMalloc(&net, 6 * B * 8);
      cudaMalloc(&mask, B); cudaMalloc(&flags, B * 4);
      // rank r owns states [r B, (r + 1) B) of the C3 stream
      rc[r] = qlb_generate_states(ctx, 3, B, (uint64_t)r * B, 0, q, quat, wrench, mask, mu, nullptr, nullptr);
      if (rc[r] == QLB_OK) rc[r] = qlb_solve_wrench(ctx, B, q, quat, wrench, mask, mu, nullptr, grf, tau, flags, net, nullptr);
      if (rc[r] == QLB_OK) rc[r] = qlb_batch_stats(ctx, B, flags, wrench, net, &stats[r], nullptr);
      if (rc[r] == QLB_OK) rc[r] = qlb_stats_allreduce(ctx, comms[r], &stats[r], nullptr);
      cudaFree(q); cudaFree(quat); cudaFree(wrench); cudaFree(mu); cudaFree(grf); cudaFree(tau); cudaFree(net); cudaFree(mask); cudaFree(flags);
      qlb_destroy(ctx);
    });
  }
  for (auto& t : th) t.join();
  for (int r = 0; r < ndev; r++) ncclCommDestroy(comms[r]);
  bool ok = true;
  for (int r = 0; r < ndev; r++) ok = ok && rc[r] == QLB_OK && stats[r].count == stats[0].count && stats[r].sum_wrench_err == stats[0].sum_wrench_err;
  ok = ok && stats[0].count == (double)(B * ndev);
  std::printf("gpus %d states %.0f ok %.0f mean_rounds %.4f mean_wrench_err %.6f max_wrench_err %.6f %s\n", ndev, stats[0].count,
              stats[0].count_status[0], stats[0].sum_iterations / stats[0].count, stats[0].sum_wrench_err / stats[0].count,
              stats[0].max_wrench_err, ok ? "AGREE" : "MISMATCH");
  return ok ? 0 : 1;
}
