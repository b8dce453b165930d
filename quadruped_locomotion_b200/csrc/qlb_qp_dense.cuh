// qlb_qp_dense.cuh - generic small dense QP, one WARP per problem, workspace in shared memory:
//     min 1/2 x'Gx + g0'x   s.t.   CE' x + ce0 = 0   (p columns),   CI' x + ci0 >= 0   (m columns)
// in the argument convention of the reference's in-repo backend quadprogpp::solve_quadprog
// (qp_solver/include/qp_solver/QuadProg++.h:8-30), so that qp_solver::QuadraticProblemSolver::minimize
// (qp_solver/src/quadraticproblemsolver.cpp:65-97) can be backed by it.  n <= 12, m <= 24, p <= 12.
// This is the path of the pose-optimisation callers (3..6 variables, qp_solver/src/sequencequadraticproblemsolver.cpp);
// the contact-force QP has its own fused kernel.
//
// Method (this repository's own formulation, shared with the contact-force kernels: everything goes through the
// Gram matrix of the active constraint normals, which is re-factorised whenever the working set changes):
//   * whiten: G = L L', w = L' x, h = L^-1 g0, m_j = L^-1 c_j.  The problem becomes
//         min 1/2 |w|^2 + h'w   s.t.   m_j' w + c0_j (=, >=) 0,
//     whose unconstrained minimiser is w = -h.  Lanes whiten the constraint columns in parallel (one forward
//     substitution each).
//   * dual active-set iteration in the whitened space.  For a working set W with normals M_W: S = M_W' M_W (built by
//     the lanes entry by entry, Cholesky-factorised in shared memory).  To bring a violated constraint p in:
//         r = S^-1 M_W' m_p,   z = m_p - M_W r   (the part of m_p orthogonal to the working set),
//     full step t2 = -slack_p / z'z along z; the multipliers of W move by -t r, so the step is cut at
//     t1 = min u_k / r_k over the inequalities of W with r_k > 0 and that constraint leaves W.  z = 0 (m_p depends on
//     W): a pure multiplier step, or - if no multiplier can give way - the problem is infeasible.
//   Equalities enter first and never leave; an equality whose normal is zero (or depends on earlier ones) with a
//   consistent right-hand side is skipped - the reference's callers pass an all-zero column.
//   Every iteration increases the dual objective, so the method is finite; the violated constraint chosen is the one
//   with the most negative slack, which is also what the reference does, so optimum AND working set agree with it
//   on non-degenerate problems (tests/test_qp_dense.py compares against the reference's own solver).
//
// The same source compiles for the host (one "lane", tests/native/qp_dense_host.cc): the algorithm is checked on
// the CPU against the oracle before it ever runs on a GPU.  That host build is test infrastructure; the library
// exports only the CUDA path.
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define QLB_QP_HD __host__ __device__ __forceinline__
#else
#define QLB_QP_HD inline
#endif
#include <math.h>

namespace qlb {

constexpr int kQpMaxN = 12, kQpMaxM = 24, kQpMaxP = 12, kQpMaxK = kQpMaxM + kQpMaxP;

struct QpDenseArgs {
  unsigned long long B;
  int n, m, p;
  const double* G;    // [n*n][B]
  const double* g0;   // [n][B]
  const double* CE;   // [n*p][B]  element (i, j) of the n x p matrix at (i*p + j)
  const double* ce0;  // [p][B]
  const double* CI;   // [n*m][B]  element (i, j) of the n x m matrix at (i*m + j)
  const double* ci0;  // [m][B]
  double* x;          // [n][B]
  double* cost;       // [B] or null
  uint32_t* status;   // [B]: 0 ok, 1 infeasible, 2 G not positive definite / non-finite input, 3 iteration limit
  uint32_t* active;   // [B] or null: bit i set = inequality i in the final working set
};

// Workspace of one problem (shared memory on the device).
struct QpWork {
  double L[kQpMaxN * kQpMaxN];    // Cholesky factor of G (lower, row-major), then reused
  double M[kQpMaxK * kQpMaxN];    // whitened normals, constraint-major: m_j at M + j * kQpMaxN
  double c0[kQpMaxK];
  double h[kQpMaxN], w[kQpMaxN], z[kQpMaxN], d[kQpMaxN];
  double S[kQpMaxN * kQpMaxN];    // Gram matrix of the working set / its Cholesky factor
  double r[kQpMaxN], u[kQpMaxN];  // direction in the multipliers; multipliers, aligned with W
  double red[32];
  int redi[32];
  int W[kQpMaxN];                 // working set: indices into the combined list (equalities 0..p-1, inequalities p..p+m-1)
  int q;                          // its size
  int flag;
};

// Execution policy: how many lanes work on one problem and how they meet.
struct QpOneLane {
  QLB_QP_HD int lane() const { return 0; }
  QLB_QP_HD int lanes() const { return 1; }
  QLB_QP_HD void sync() const {}
};
#ifdef __CUDACC__
struct QpWarp {
  __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31u); }
  __device__ __forceinline__ int lanes() const { return 32; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
};
#endif

QLB_QP_HD bool qp_finite(double v) { return v - v == 0.0; }

// In-place Cholesky of the leading k x k block of a row-major matrix with row stride ld (lower triangle); serial
// (k <= 12).  Returns false when a pivot is not positive.
QLB_QP_HD bool qp_chol(double* A, int k, int ld) {
  for (int j = 0; j < k; j++) {
    double dgl = A[j * ld + j];
    for (int t = 0; t < j; t++) dgl -= A[j * ld + t] * A[j * ld + t];
    if (!(dgl > 0.0) || !qp_finite(dgl)) return false;
    const double l = sqrt(dgl);
    A[j * ld + j] = l;
    for (int i = j + 1; i < k; i++) {
      double s = A[i * ld + j];
      for (int t = 0; t < j; t++) s -= A[i * ld + t] * A[j * ld + t];
      A[i * ld + j] = s / l;
    }
  }
  return true;
}

// r and z for bringing constraint `pj` into the working set (see the header).  All lanes; leaves ws.d = m_p,
// ws.r (q entries), ws.z; returns z'z.
template <class Par>
QLB_QP_HD double qp_direction(const Par& par, QpWork& ws, int n, int pj) {
  const int lane = par.lane(), nl = par.lanes();
  const int q = ws.q;
  const double* mp = ws.M + pj * kQpMaxN;
  // Gram matrix of the working set (lower triangle) and the right-hand side M_W' m_p
  for (int e = lane; e < q * q; e += nl) {
    const int a = e / q, b = e - a * q;
    if (b <= a) {
      const double* ma = ws.M + ws.W[a] * kQpMaxN;
      const double* mb = ws.M + ws.W[b] * kQpMaxN;
      double s = 0.0;
      for (int i = 0; i < n; i++) s += ma[i] * mb[i];
      ws.S[a * kQpMaxN + b] = s;
    }
  }
  for (int a = lane; a < q; a += nl) {
    const double* ma = ws.M + ws.W[a] * kQpMaxN;
    double s = 0.0;
    for (int i = 0; i < n; i++) s += ma[i] * mp[i];
    ws.r[a] = s;
  }
  par.sync();
  if (lane == 0) {
    // the working set is kept independent, so the factorisation succeeds; a failed pivot (rounding on a nearly
    // dependent set) is reported through ws.flag and ends the solve with the iteration-limit status
    if (q > 0 && !qp_chol(ws.S, q, kQpMaxN)) ws.flag = 1;
    for (int a = 0; a < q; a++) {
      double s = ws.r[a];
      for (int t = 0; t < a; t++) s -= ws.S[a * kQpMaxN + t] * ws.r[t];
      ws.r[a] = s / ws.S[a * kQpMaxN + a];
    }
    for (int a = q - 1; a >= 0; a--) {
      double s = ws.r[a];
      for (int t = a + 1; t < q; t++) s -= ws.S[t * kQpMaxN + a] * ws.r[t];
      ws.r[a] = s / ws.S[a * kQpMaxN + a];
    }
  }
  par.sync();
  for (int i = lane; i < n; i += nl) {
    double s = mp[i];
    for (int a = 0; a < q; a++) s -= ws.r[a] * ws.M[ws.W[a] * kQpMaxN + i];
    ws.z[i] = s;
    ws.d[i] = mp[i];
  }
  par.sync();
  double zz = 0.0;
  for (int i = 0; i < n; i++) zz += ws.z[i] * ws.z[i];
  return zz;
}

// Solve problem `b` of the batch.  Every lane of the policy runs this function with the same arguments; ws is the
// problem's workspace.
template <class Par>
QLB_QP_HD void qp_dense_solve(const Par& par, const QpDenseArgs& a, unsigned long long b, QpWork& ws) {
  const int lane = par.lane(), nl = par.lanes();
  const int n = a.n, m = a.m, p = a.p, K = a.m + a.p;
  const unsigned long long B = a.B;
  // ---- inputs: G, the constraint columns (combined list: equalities first), offsets
  bool fin = true;
  for (int e = lane; e < n * n; e += nl) {
    const double v = a.G[(unsigned long long)e * B + b];
    ws.L[(e / n) * kQpMaxN + (e % n)] = v;
    fin = fin && qp_finite(v);
  }
  for (int e = lane; e < n * K; e += nl) {
    const int j = e / n, i = e - j * n;   // component i of constraint j
    const double v = (j < p) ? a.CE[(unsigned long long)(i * p + j) * B + b] : a.CI[(unsigned long long)(i * m + (j - p)) * B + b];
    ws.M[j * kQpMaxN + i] = v;
    fin = fin && qp_finite(v);
  }
  for (int j = lane; j < K; j += nl) {
    const double v = (j < p) ? a.ce0[(unsigned long long)j * B + b] : a.ci0[(unsigned long long)(j - p) * B + b];
    ws.c0[j] = v;
    fin = fin && qp_finite(v);
  }
  for (int i = lane; i < n; i += nl) {
    const double v = a.g0[(unsigned long long)i * B + b];
    ws.h[i] = v;
    fin = fin && qp_finite(v);
  }
  ws.red[lane] = fin ? 0.0 : 1.0;
  if (lane == 0) { ws.q = 0; ws.flag = 0; }
  par.sync();
  bool bad = false;
  for (int l = 0; l < nl; l++) bad = bad || ws.red[l] != 0.0;
  par.sync();
  // ---- G = L L'
  if (lane == 0) ws.redi[0] = (!bad && qp_chol(ws.L, n, kQpMaxN)) ? 1 : 0;
  par.sync();
  int status = 0;
  if (ws.redi[0] == 0) status = 2;
  par.sync();
  if (status == 0) {
    // ---- whiten: m_j = L^-1 c_j (one forward substitution per lane), h = L^-1 g0, w = -h
    for (int j = lane; j <= K; j += nl) {
      double* v = (j < K) ? (ws.M + j * kQpMaxN) : ws.h;
      for (int i = 0; i < n; i++) {
        double s = v[i];
        for (int t = 0; t < i; t++) s -= ws.L[i * kQpMaxN + t] * v[t];
        v[i] = s / ws.L[i * kQpMaxN + i];
      }
    }
    par.sync();
    for (int i = lane; i < n; i += nl) ws.w[i] = -ws.h[i];
    par.sync();
    // ---- equalities first; they never leave the working set
    for (int e = 0; e < p && status == 0; e++) {
      const double zz = qp_direction(par, ws, n, e);
      double dd = 0.0, sl = ws.c0[e], sc = fabs(ws.c0[e]);
      for (int i = 0; i < n; i++) { dd += ws.d[i] * ws.d[i]; sl += ws.d[i] * ws.w[i]; sc += fabs(ws.d[i] * ws.w[i]); }
      par.sync();
      if (ws.flag != 0) { status = 3; break; }
      if (!(zz > 1e-22 * dd) || dd == 0.0) {
        // zero or dependent normal: consistent -> the column is absent; else no point satisfies the equalities
        if (fabs(sl) > 1e-9 * (sc + 1e-300) && fabs(sl) > 1e-12) status = 1;
        continue;
      }
      const double t = -sl / zz;
      if (lane == 0) {
        for (int k = 0; k < ws.q; k++) ws.u[k] -= t * ws.r[k];
        ws.W[ws.q] = e; ws.u[ws.q] = t; ws.q = ws.q + 1;
      }
      for (int i = lane; i < n; i += nl) ws.w[i] += t * ws.z[i];
      par.sync();
    }
    // ---- inequalities
    int iter = 0;
    const int max_iter = 12 * (K + 2);
    while (status == 0) {
      // most violated inactive inequality (lowest index on ties)
      double best = 0.0;
      int bj = -1;
      for (int j = p + lane; j < K; j += nl) {
        bool inW = false;
        for (int k = 0; k < ws.q; k++) inW = inW || ws.W[k] == j;
        if (inW) continue;
        const double* mj = ws.M + j * kQpMaxN;
        double sl = ws.c0[j], sc = fabs(ws.c0[j]);
        for (int i = 0; i < n; i++) { sl += mj[i] * ws.w[i]; sc += fabs(mj[i] * ws.w[i]); }
        if (sl < -1e-12 * (sc + 1e-300) && sl < best) { best = sl; bj = j; }
      }
      ws.red[lane] = best; ws.redi[lane] = bj;
      par.sync();
      best = 0.0; bj = -1;
      for (int l = 0; l < nl; l++)
        if (ws.redi[l] >= 0 && (ws.red[l] < best || (ws.red[l] == best && ws.redi[l] < bj))) { best = ws.red[l]; bj = ws.redi[l]; }
      par.sync();
      if (bj < 0) break;   // optimal
      double up = 0.0;     // multiplier of the entering constraint
      for (;;) {
        if (++iter > max_iter) { status = 3; break; }
        const double zz = qp_direction(par, ws, n, bj);
        if (ws.flag != 0) { status = 3; break; }
        double dd = 0.0, sl = ws.c0[bj];
        for (int i = 0; i < n; i++) { dd += ws.d[i] * ws.d[i]; sl += ws.d[i] * ws.w[i]; }
        // the multiplier step: which inequality of W gives way first?
        double t1 = 1e300;
        int kd = -1;
        for (int k = 0; k < ws.q; k++)
          if (ws.W[k] >= p && ws.r[k] > 0.0) {
            const double t = ws.u[k] / ws.r[k];
            if (t < t1) { t1 = t; kd = k; }
          }
        const bool dependent = !(zz > 1e-22 * dd);
        if (dependent && kd < 0) { status = 1; break; }   // nothing can give way: infeasible
        const double t2 = dependent ? 1e300 : -sl / zz;
        const bool full = t2 <= t1;
        const double t = full ? t2 : t1;
        par.sync();   // every lane has read u, r, W before lane 0 changes them
        if (lane == 0) {
          for (int k = 0; k < ws.q; k++) { ws.u[k] -= t * ws.r[k]; if (ws.W[k] >= p && ws.u[k] < 0.0) ws.u[k] = 0.0; }
          if (full) {
            ws.W[ws.q] = bj; ws.u[ws.q] = up + t; ws.q = ws.q + 1;
          } else {
            for (int k = kd; k + 1 < ws.q; k++) { ws.W[k] = ws.W[k + 1]; ws.u[k] = ws.u[k + 1]; }
            ws.q = ws.q - 1;
          }
        }
        if (!dependent)
          for (int i = lane; i < n; i += nl) ws.w[i] += t * ws.z[i];
        up += t;
        par.sync();
        if (full) break;
      }
    }
  }
  // ---- outputs: x = L^-T w, cost, status, working set
  par.sync();
  if (lane == 0) {
    double cost = 0.0;
    if (status == 0 || status == 3) {
      for (int i = 0; i < n; i++) cost += ws.w[i] * (0.5 * ws.w[i] + ws.h[i]);
      for (int i = n - 1; i >= 0; i--) {
        double s = ws.w[i];
        for (int t = i + 1; t < n; t++) s -= ws.L[t * kQpMaxN + i] * ws.z[t];
        ws.z[i] = s / ws.L[i * kQpMaxN + i];   // z reused for x
      }
    }
    unsigned act = 0u;
    if (status == 0)
      for (int k = 0; k < ws.q; k++)
        if (ws.W[k] >= p) act |= 1u << (ws.W[k] - p);
    const bool have = (status == 0 || status == 3);
    for (int i = 0; i < n; i++) a.x[(unsigned long long)i * B + b] = have ? ws.z[i] : 0.0;
    if (a.cost) a.cost[b] = have ? cost : (double)INFINITY;   // no solution: +inf like the reference (QuadProg++.cc:340-344)
    a.status[b] = (uint32_t)status;
    if (a.active) a.active[b] = act;
  }
  par.sync();
}

#ifdef __CUDACC__
constexpr int kQpWarpsPerCta = 4;
__global__ void __launch_bounds__(32 * kQpWarpsPerCta) qlb_qp_dense_kernel(const QpDenseArgs a) {
  __shared__ QpWork work[kQpWarpsPerCta];
  const int warp = threadIdx.x >> 5;
  const QpWarp par;
  for (unsigned long long b = (unsigned long long)blockIdx.x * kQpWarpsPerCta + warp; b < a.B; b += (unsigned long long)gridDim.x * kQpWarpsPerCta)
    qp_dense_solve(par, a, b, work[warp]);
}
#endif

}  // namespace qlb
