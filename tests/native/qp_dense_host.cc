// Host build of the generic QP algorithm of csrc/qlb_qp_dense.cuh (one "lane"): TEST INFRASTRUCTURE.  The CPU test
// suite checks the algorithm against the oracle before the same source runs as a warp-cooperative CUDA kernel.
// Built by tests/test_qp_dense.py with g++; nothing in the product loads it.
#include "qlb_qp_dense.cuh"

extern "C" void qp_dense_host(unsigned long long B, int n, int m, int p, const double* G, const double* g0, const double* CE,
                              const double* ce0, const double* CI, const double* ci0, double* x, double* cost, uint32_t* status,
                              uint32_t* active) {
  qlb::QpDenseArgs a;
  a.B = B; a.n = n; a.m = m; a.p = p; a.G = G; a.g0 = g0; a.CE = CE; a.ce0 = ce0; a.CI = CI; a.ci0 = ci0;
  a.x = x; a.cost = cost; a.status = status; a.active = active;
  static qlb::QpWork ws;
  const qlb::QpOneLane par;
  for (unsigned long long b = 0; b < B; b++) qlb::qp_dense_solve(par, a, b, ws);
}
