// qlb_qp_dense.cuh - generic small dense QP, one thread per problem:
//     min 1/2 x'Gx + g0'x   s.t.   CE' x + ce0 = 0   (p columns),   CI' x + ci0 >= 0   (m columns)
// in exactly the argument convention of the reference's in-repo backend quadprogpp::solve_quadprog
// (qp_solver/include/qp_solver/QuadProg++.h:8-30, qp_solver/src/QuadProg++.cc:52-446), so that
// qp_solver::QuadraticProblemSolver::minimize (qp_solver/src/quadraticproblemsolver.cpp:65-97) can be
// backed by it.  Dual active-set method of Goldfarb and Idnani with the same pivoting rules and
// tolerances as the reference, hence the same optimum and working set.  n <= 12, m <= 24, p <= 12.
// This is the path for the pose-optimisation style callers (3..6 variables, a handful per tick); the
// contact-force QP has its own fused kernel (qlb_solve.cuh).
#pragma once

#include <cfloat>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace qlb {

constexpr int kQpMaxN = 12, kQpMaxM = 24, kQpMaxP = 12;

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

struct GiWork {
  double L[kQpMaxN][kQpMaxN];
  double J[kQpMaxN][kQpMaxN];
  double R[kQpMaxN][kQpMaxN];
};

__device__ inline double gi_hypot(double a, double b) {
  const double a1 = fabs(a), b1 = fabs(b);
  if (a1 > b1) { const double t = b1 / a1; return a1 * sqrt(1.0 + t * t); }
  if (b1 > a1) { const double t = a1 / b1; return b1 * sqrt(1.0 + t * t); }
  return a1 * sqrt(2.0);
}
__device__ inline bool gi_givens(double a, double b, double& c, double& s, double& h) {
  const double hh = gi_hypot(a, b);
  if (fabs(hh) < DBL_EPSILON) return false;
  c = a / hh; s = b / hh;
  if (c < 0.0) { c = -c; s = -s; h = -hh; } else { h = hh; }
  return true;
}
__device__ inline void gi_direction(const GiWork& w, int n, const double* np, int iq, double* d, double* z, double* r) {
  for (int i = 0; i < n; i++) {
    double s = 0.0;
    for (int j = 0; j < n; j++) s += w.J[j][i] * np[j];
    d[i] = s;
  }
  for (int i = 0; i < n; i++) {
    double s = 0.0;
    for (int j = iq; j < n; j++) s += w.J[i][j] * d[j];
    z[i] = s;
  }
  for (int i = iq - 1; i >= 0; i--) {
    double s = 0.0;
    for (int j = i + 1; j < iq; j++) s += w.R[i][j] * r[j];
    r[i] = (d[i] - s) / w.R[i][i];
  }
}
__device__ inline bool gi_add(GiWork& w, int n, double* d, int& iq, double& rnorm) {
  for (int j = n - 1; j >= iq + 1; j--) {
    double c, s, h;
    if (!gi_givens(d[j - 1], d[j], c, s, h)) continue;
    d[j] = 0.0;
    d[j - 1] = h;
    const double xny = s / (1.0 + c);
    for (int k = 0; k < n; k++) {
      const double t1 = w.J[k][j - 1], t2 = w.J[k][j];
      w.J[k][j - 1] = t1 * c + t2 * s;
      w.J[k][j] = xny * (t1 + w.J[k][j - 1]) - t2;
    }
  }
  iq++;
  for (int i = 0; i < iq; i++) w.R[i][iq - 1] = d[i];
  if (fabs(d[iq - 1]) <= DBL_EPSILON * rnorm) return false;
  rnorm = fmax(rnorm, fabs(d[iq - 1]));
  return true;
}
__device__ inline void gi_delete(GiWork& w, int n, int* A, double* u, int p, int& iq, int l) {
  int qq = -1;
  for (int i = p; i < iq; i++)
    if (A[i] == l) { qq = i; break; }
  if (qq < 0) return;
  for (int i = qq; i < iq - 1; i++) {
    A[i] = A[i + 1];
    u[i] = u[i + 1];
    for (int j = 0; j < n; j++) w.R[j][i] = w.R[j][i + 1];
  }
  A[iq - 1] = A[iq];
  u[iq - 1] = u[iq];
  A[iq] = 0;
  u[iq] = 0.0;
  for (int j = 0; j < iq; j++) w.R[j][iq - 1] = 0.0;
  iq--;
  if (iq == 0) return;
  for (int j = qq; j < iq; j++) {
    double c, s, h;
    if (!gi_givens(w.R[j][j], w.R[j + 1][j], c, s, h)) continue;
    w.R[j + 1][j] = 0.0;
    w.R[j][j] = h;
    const double xny = s / (1.0 + c);
    for (int k = j + 1; k < iq; k++) {
      const double t1 = w.R[j][k], t2 = w.R[j + 1][k];
      w.R[j][k] = t1 * c + t2 * s;
      w.R[j + 1][k] = xny * (t1 + w.R[j][k]) - t2;
    }
    for (int k = 0; k < n; k++) {
      const double t1 = w.J[k][j], t2 = w.J[k][j + 1];
      w.J[k][j] = t1 * c + t2 * s;
      w.J[k][j + 1] = xny * (w.J[k][j] + t1) - t2;
    }
  }
}

__global__ void __launch_bounds__(64) qlb_qp_dense_kernel(const QpDenseArgs a) {
  const unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const int n = a.n, m = a.m;
  const size_t B = a.B;
  GiWork w;
  double g0[kQpMaxN], x[kQpMaxN], z[kQpMaxN], d[kQpMaxN], np[kQpMaxN], x_old[kQpMaxN];
  double s[kQpMaxM + kQpMaxP], r[kQpMaxM + kQpMaxP], u[kQpMaxM + kQpMaxP + 1], u_old[kQpMaxM + kQpMaxP + 1];
  int A[kQpMaxM + kQpMaxP + 1], A_old[kQpMaxM + kQpMaxP + 1], iai[kQpMaxM + kQpMaxP];
  bool iaexcl[kQpMaxM + kQpMaxP];
  int eqcol[kQpMaxP];
  unsigned st = 0;
  bool finite = true;

  // equality columns that are identically zero carry no constraint (the reference's callers pass one,
  // qp_solver/src/pose_optimization/PoseOptimizationQP.cpp:106-112); with a non-zero offset they are infeasible
  int p = 0;
  for (int j = 0; j < a.p; j++) {
    bool zero = true;
    for (int i = 0; i < n; i++) zero = zero && (a.CE[(size_t)(i * a.p + j) * B + b] == 0.0);
    const double c0 = a.ce0[(size_t)j * B + b];
    finite = finite && isfinite(c0);
    if (zero) { if (c0 != 0.0) st = 1; }
    else eqcol[p++] = j;
  }
  double c1 = 0.0, c2 = 0.0;
  for (int i = 0; i < n; i++) {
    g0[i] = a.g0[(size_t)i * B + b];
    finite = finite && isfinite(g0[i]);
    for (int j = 0; j < n; j++) {
      w.L[i][j] = a.G[(size_t)(i * n + j) * B + b];
      finite = finite && isfinite(w.L[i][j]);
      w.R[i][j] = 0.0;
    }
    c1 += w.L[i][i];
  }
  // Cholesky G = L L' (QuadProg++.cc:672-712)
  bool pd = finite;
  for (int i = 0; i < n && pd; i++) {
    for (int j = i; j < n; j++) {
      double sum = w.L[i][j];
      for (int k = i - 1; k >= 0; k--) sum -= w.L[i][k] * w.L[j][k];
      if (i == j) {
        if (!(sum > 0.0)) { pd = false; break; }
        w.L[i][i] = sqrt(sum);
      } else {
        w.L[j][i] = sum / w.L[i][i];
      }
    }
    for (int k = i + 1; k < n; k++) w.L[i][k] = w.L[k][i];
  }
  double fval = CUDART_INF;
  unsigned actbits = 0;
  int iq = 0;
  if (!pd) st = 2;
  if (st == 0) {
    // J = L^-T, c2 = trace(J)
    for (int i = 0; i < n; i++) {
      for (int k = 0; k < n; k++) {  // forward elimination of e_i
        double v = (k == i) ? 1.0 : 0.0;
        for (int j = 0; j < k; j++) v -= w.L[k][j] * z[j];
        z[k] = v / w.L[k][k];
      }
      for (int j = 0; j < n; j++) w.J[i][j] = z[j];
      c2 += z[i];
    }
    // x = -G^-1 g0
    for (int k = 0; k < n; k++) {
      double v = g0[k];
      for (int j = 0; j < k; j++) v -= w.L[k][j] * z[j];
      z[k] = v / w.L[k][k];
    }
    for (int k = n - 1; k >= 0; k--) {
      double v = z[k];
      for (int j = k + 1; j < n; j++) v -= w.L[k][j] * x[j];
      x[k] = v / w.L[k][k];
    }
    fval = 0.0;
    for (int i = 0; i < n; i++) { x[i] = -x[i]; fval += g0[i] * x[i]; }
    fval *= 0.5;
    for (int i = 0; i <= m + p; i++) { u[i] = 0.0; A[i] = 0; }
    double rnorm = 1.0;
    // equality constraints (QuadProg++.cc:178-210)
    for (int i = 0; i < p; i++) {
      const int col = eqcol[i];
      for (int j = 0; j < n; j++) np[j] = a.CE[(size_t)(j * a.p + col) * B + b];
      gi_direction(w, n, np, iq, d, z, r);
      double zz = 0.0, znp = 0.0, npx = 0.0;
      for (int k = 0; k < n; k++) { zz += z[k] * z[k]; znp += z[k] * np[k]; npx += np[k] * x[k]; }
      double t2 = 0.0;
      if (fabs(zz) > DBL_EPSILON) t2 = (-npx - a.ce0[(size_t)col * B + b]) / znp;
      for (int k = 0; k < n; k++) x[k] += t2 * z[k];
      u[iq] = t2;
      for (int k = 0; k < iq; k++) u[k] -= t2 * r[k];
      fval += 0.5 * (t2 * t2) * znp;
      A[i] = -i - 1;
      gi_add(w, n, d, iq, rnorm);
    }
    for (int i = 0; i < m; i++) iai[i] = i;
    int phase = 0, ip = 0, l = 0, iter = 0;
    double ss = 0.0, t1, t2, t;
    for (;;) {
      if (phase == 0) {  // QuadProg++.cc:216-262
        if (++iter > 200) { st = 3; break; }
        for (int i = p; i < iq; i++) iai[A[i]] = -1;
        ss = 0.0;
        ip = 0;
        double psi = 0.0;
        for (int i = 0; i < m; i++) {
          iaexcl[i] = true;
          double sum = 0.0;
          for (int j = 0; j < n; j++) sum += a.CI[(size_t)(j * m + i) * B + b] * x[j];
          sum += a.ci0[(size_t)i * B + b];
          s[i] = sum;
          psi += fmin(0.0, sum);
        }
        if (fabs(psi) <= m * DBL_EPSILON * c1 * c2 * 100.0) break;
        for (int i = 0; i < iq; i++) { u_old[i] = u[i]; A_old[i] = A[i]; }
        for (int i = 0; i < n; i++) x_old[i] = x[i];
        phase = 1;
      }
      if (phase == 1) {  // QuadProg++.cc:264-288
        for (int i = 0; i < m; i++)
          if (s[i] < ss && iai[i] != -1 && iaexcl[i]) { ss = s[i]; ip = i; }
        if (ss >= 0.0) break;
        for (int i = 0; i < n; i++) np[i] = a.CI[(size_t)(i * m + ip) * B + b];
        u[iq] = 0.0;
        A[iq] = ip;
        phase = 2;
      }
      gi_direction(w, n, np, iq, d, z, r);  // QuadProg++.cc:290-338
      l = 0;
      t1 = CUDART_INF;
      for (int k = p; k < iq; k++)
        if (r[k] > 0.0 && u[k] / r[k] < t1) { t1 = u[k] / r[k]; l = A[k]; }
      double zz = 0.0, znp = 0.0;
      for (int k = 0; k < n; k++) { zz += z[k] * z[k]; znp += z[k] * np[k]; }
      if (fabs(zz) > DBL_EPSILON) {
        t2 = -s[ip] / znp;
        if (t2 < 0) t2 = CUDART_INF;
      } else {
        t2 = CUDART_INF;
      }
      t = fmin(t1, t2);
      if (t >= CUDART_INF) { st = 1; fval = CUDART_INF; break; }
      if (t2 >= CUDART_INF) {  // dual step
        for (int k = 0; k < iq; k++) u[k] -= t * r[k];
        u[iq] += t;
        iai[l] = l;
        gi_delete(w, n, A, u, p, iq, l);
        continue;
      }
      for (int k = 0; k < n; k++) x[k] += t * z[k];
      fval += t * znp * (0.5 * t + u[iq]);
      for (int k = 0; k < iq; k++) u[k] -= t * r[k];
      u[iq] += t;
      if (fabs(t - t2) < DBL_EPSILON) {  // full step
        if (!gi_add(w, n, d, iq, rnorm)) {
          iaexcl[ip] = false;
          gi_delete(w, n, A, u, p, iq, ip);
          for (int i = 0; i < m; i++) iai[i] = i;
          for (int i = p; i < iq; i++) { A[i] = A_old[i]; u[i] = u_old[i]; iai[A[i]] = -1; }
          for (int i = 0; i < n; i++) x[i] = x_old[i];
          phase = 1;
        } else {
          iai[ip] = -1;
          phase = 0;
        }
        continue;
      }
      iai[l] = l;  // partial step
      gi_delete(w, n, A, u, p, iq, l);
      double sum = 0.0;
      for (int k = 0; k < n; k++) sum += a.CI[(size_t)(k * m + ip) * B + b] * x[k];
      s[ip] = sum + a.ci0[(size_t)ip * B + b];
      phase = 2;
    }
    if (st == 0)
      for (int i = p; i < iq; i++) actbits |= 1u << A[i];
  }
  for (int i = 0; i < n; i++) a.x[(size_t)i * B + b] = (st == 0 || st == 3) ? x[i] : 0.0;
  if (a.cost) a.cost[b] = (st == 0 || st == 3) ? fval : CUDART_INF;
  a.status[b] = st;
  if (a.active) a.active[b] = actbits;
}

}  // namespace qlb
