// threads_demo.cpp - one context shared by several host threads: each thread solves its own slice of a batch through the
// host entry point (qlb_solve_wrench_host), all at the same time; the results must be the bits of a serial run.
//   threads_demo [threads] [states_per_thread]     prints "threads T states N SAME|DIFFERENT"; exit code 0 when same
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "qlb.h"
#include "qlb_models.h"

int main(int argc, char** argv) {
  const int T = argc > 1 ? atoi(argv[1]) : 4;
  const size_t n = argc > 2 ? (size_t)atoll(argv[2]) : 20000;
  const size_t B = n * T;
  qlb_context* ctx = nullptr;
  if (qlb_create(&ctx, QLB_MODEL_QUADRUPED_MODEL, nullptr, 0, 0) != QLB_OK) { std::fprintf(stderr, "qlb_create failed\n"); return 2; }
  // the C3 stream, generated on the device and brought to the host once
  double *dq, *dquat, *dw, *dmu; uint8_t* dmask;
  cudaMalloc(&dq, 12 * B * 8); cudaMalloc(&dquat, 4 * B * 8); cudaMalloc(&dw, 6 * B * 8); cudaMalloc(&dmu, 4 * B * 8); cudaMalloc(&dmask, B);
  if (qlb_generate_states(ctx, 3, B, 0, 0, dq, dquat, dw, dmask, dmu, nullptr, nullptr) != QLB_OK) return 2;
  std::vector<double> q(12 * B), quat(4 * B), w(6 * B), mu(4 * B);
  std::vector<uint8_t> mask(B);
  cudaMemcpy(q.data(), dq, 12 * B * 8, cudaMemcpyDeviceToHost); cudaMemcpy(quat.data(), dquat, 4 * B * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(w.data(), dw, 6 * B * 8, cudaMemcpyDeviceToHost); cudaMemcpy(mu.data(), dmu, 4 * B * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(mask.data(), dmask, B, cudaMemcpyDeviceToHost);
  // slice t as its own SoA block [rows][n]
  auto slice = [&](const std::vector<double>& a, int rows, int t) {
    std::vector<double> s((size_t)rows * n);
    for (int r = 0; r < rows; r++) std::memcpy(&s[(size_t)r * n], &a[(size_t)r * B + (size_t)t * n], n * 8);
    return s;
  };
  struct Out { std::vector<double> grf, tau, net; std::vector<uint32_t> flags; int rc; };
  auto run = [&](int t, Out& o) {
    const auto sq = slice(q, 12, t), sqt = slice(quat, 4, t), sw = slice(w, 6, t), smu = slice(mu, 4, t);
    o.grf.assign(12 * n, 0.0); o.tau.assign(12 * n, 0.0); o.net.assign(6 * n, 0.0); o.flags.assign(n, 0u);
    o.rc = qlb_solve_wrench_host(ctx, n, sq.data(), sqt.data(), sw.data(), mask.data() + (size_t)t * n, smu.data(), nullptr,
                                 o.grf.data(), o.tau.data(), o.flags.data(), o.net.data());
  };
  std::vector<Out> par(T), ser(T);
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++) th.emplace_back([&, t]() { run(t, par[t]); });
  for (auto& x : th) x.join();
  for (int t = 0; t < T; t++) run(t, ser[t]);
  bool same = true;
  for (int t = 0; t < T; t++)
    same = same && par[t].rc == QLB_OK && ser[t].rc == QLB_OK && par[t].flags == ser[t].flags &&
           !std::memcmp(par[t].grf.data(), ser[t].grf.data(), 12 * n * 8) && !std::memcmp(par[t].tau.data(), ser[t].tau.data(), 12 * n * 8) &&
           !std::memcmp(par[t].net.data(), ser[t].net.data(), 6 * n * 8);
  size_t ok = 0;
  for (int t = 0; t < T; t++) for (size_t i = 0; i < n; i++) ok += ((par[t].flags[i] >> 24) & 7u) == 0u;
  std::printf("threads %d states %zu ok %zu %s\n", T, B, ok, same ? "SAME" : "DIFFERENT");
  cudaFree(dq); cudaFree(dquat); cudaFree(dw); cudaFree(dmu); cudaFree(dmask);
  qlb_destroy(ctx);
  return same ? 0 : 1;
}
