// ref_shim.cc - C entry points around the REFERENCE's own vendored solver, compiled in place from
// /root/reference/qp_solver/src/{QuadProg++.cc,Array.cc} into oracle/_ref/ (never copied into this
// repo).  TEST INFRASTRUCTURE ONLY: validates the oracle restatement and serves as the CPU baseline
// ("kind": "reference").  Always called with zero equality columns: the fork removed the
// dependent-equality guard (QuadProg++.cc:203-209), so an all-zero CE column corrupts its answer.
#include <cmath>
#include <exception>
#include <limits>

#include "qp_solver/QuadProg++.h"

extern "C" {

// min 1/2 x'Gx + g0'x  s.t.  D x >= d   (G n x n row-major, D m x n row-major)
double qref_solve_quadprog(int n, int m, const double* G, const double* g0, const double* D,
                           const double* d, double* x) {
  try {
    quadprogpp::Matrix<double> Gm(n, n), CE(n, 0), CI(n, m);
    quadprogpp::Vector<double> g(n), ce0(0), ci0(m), xv(n);
    for (int i = 0; i < n; i++) {
      g[i] = g0[i];
      for (int j = 0; j < n; j++) Gm[i][j] = G[i * n + j];
    }
    for (int i = 0; i < m; i++) {
      ci0[i] = -d[i];
      for (int j = 0; j < n; j++) CI[j][i] = D[i * n + j];
    }
    const double f = quadprogpp::solve_quadprog(Gm, g, CE, ce0, CI, ci0, xv);
    for (int i = 0; i < n; i++) x[i] = xv[i];
    return f;
  } catch (const std::exception&) {
    return std::numeric_limits<double>::quiet_NaN();
  }
}

// general form with equality columns: CE n x p row-major, CE' x + ce0 = 0
double qref_solve_quadprog_eq(int n, int m, int p, const double* G, const double* g0,
                              const double* CEa, const double* ce0a, const double* D,
                              const double* d, double* x) {
  try {
    quadprogpp::Matrix<double> Gm(n, n), CE(n, p), CI(n, m);
    quadprogpp::Vector<double> g(n), ce0(p), ci0(m), xv(n);
    for (int i = 0; i < n; i++) {
      g[i] = g0[i];
      for (int j = 0; j < n; j++) Gm[i][j] = G[i * n + j];
      for (int j = 0; j < p; j++) CE[i][j] = CEa[i * p + j];
    }
    for (int j = 0; j < p; j++) ce0[j] = ce0a[j];
    for (int i = 0; i < m; i++) {
      ci0[i] = -d[i];
      for (int j = 0; j < n; j++) CI[j][i] = D[i * n + j];
    }
    const double f = quadprogpp::solve_quadprog(Gm, g, CE, ce0, CI, ci0, xv);
    for (int i = 0; i < n; i++) x[i] = xv[i];
    return f;
  } catch (const std::exception&) {
    return std::numeric_limits<double>::quiet_NaN();
  }
}

}  // extern "C"
