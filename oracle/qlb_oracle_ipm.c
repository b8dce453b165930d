/*
 * qlb_oracle_ipm.c - CPU ORACLE (TEST INFRASTRUCTURE ONLY): dense primal-dual interior point.
 *
 * The live reference path hands the QP to OOQP (Mehrotra/Gondzio predictor-corrector, MA27) through
 * ooqpei::QuadraticProblemFormulation::solve (ContactForceDistribution.cpp:490); neither library is
 * vendored.  This file is the "OOQP-style" CPU stand-in used beside the GPU in bench.py and a second
 * opinion for the Goldfarb-Idnani oracle: the same algorithm family the GPU kernel runs
 *   1. unconstrained minimiser; finished if it is feasible
 *   2. Mehrotra predictor-corrector iterations on  min 1/2 x'Gx + g'x, Dx - s = d, s >= 0
 *   3. active-set polish: rows with lambda_i > s_i are made equalities, the KKT system is solved
 *      exactly and the signs of multipliers / slacks are verified (repaired a few times if needed)
 * but written for a general dense D.
 */
#include <float.h>
#include <math.h>
#include <string.h>

#include "qlb_oracle.h"

#define NMAX QO_MAX_N
#define MMAX QO_MAX_M

static int chol(int n, double* H) { /* lower Cholesky in place, row-major n x n */
  for (int j = 0; j < n; j++) {
    double dj = H[j * n + j];
    for (int k = 0; k < j; k++) dj -= H[j * n + k] * H[j * n + k];
    if (!(dj > 0.0)) return 0;
    dj = sqrt(dj);
    H[j * n + j] = dj;
    for (int i = j + 1; i < n; i++) {
      double s = H[i * n + j];
      for (int k = 0; k < j; k++) s -= H[i * n + k] * H[j * n + k];
      H[i * n + j] = s / dj;
    }
  }
  return 1;
}
static void chol_solve(int n, const double* L, double* b) {
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * b[k];
    b[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * b[k];
    b[i] = s / L[i * n + i];
  }
}

/* exact solve of  min 1/2 x'Gx + g'x  s.t.  D_A x = d_A  by the range-space method.
 * LG = Cholesky factor of G.  Returns 0 if D_A is rank deficient. */
static int kkt_solve(int n, int m, const double* LG, const double* g0, const double* D, const double* d,
                     const int* act, double* x, double* u) {
  int idx[NMAX], k = 0;
  for (int i = 0; i < m; i++) {
    u[i] = 0.0;
    if (act[i]) { if (k == n) return 0; idx[k++] = i; }
  }
  double xg[NMAX];
  for (int i = 0; i < n; i++) xg[i] = -g0[i];
  chol_solve(n, LG, xg); /* unconstrained minimiser */
  if (k == 0) { memcpy(x, xg, n * sizeof(double)); return 1; }
  double Y[NMAX * NMAX], M[NMAX * NMAX], rhs[NMAX];
  for (int a = 0; a < k; a++) { /* Y_a = G^-1 D_a' */
    for (int j = 0; j < n; j++) Y[a * n + j] = D[idx[a] * n + j];
    chol_solve(n, LG, Y + a * n);
  }
  for (int a = 0; a < k; a++) {
    for (int b = 0; b < k; b++) {
      double s = 0.0;
      for (int j = 0; j < n; j++) s += D[idx[a] * n + j] * Y[b * n + j];
      M[a * k + b] = s;
    }
    double s = d[idx[a]];
    for (int j = 0; j < n; j++) s -= D[idx[a] * n + j] * xg[j];
    rhs[a] = s;
  }
  if (!chol(k, M)) return 0;
  chol_solve(k, M, rhs);
  for (int j = 0; j < n; j++) {
    double s = xg[j];
    for (int a = 0; a < k; a++) s += Y[a * n + j] * rhs[a];
    x[j] = s;
  }
  for (int a = 0; a < k; a++) u[idx[a]] = rhs[a];
  return 1;
}

/* polish from an active-set guess; repairs the guess with primal-dual active-set steps */
static int polish(int n, int m, const double* LG, const double* g0, const double* D, const double* d,
                  int* act, double* x, double* u) {
  double scale = 1.0;
  for (int it = 0; it < 12; it++) {
    double xt[NMAX], ut[MMAX];
    if (!kkt_solve(n, m, LG, g0, D, d, act, xt, ut)) return 0;
    scale = 1.0;
    for (int j = 0; j < n; j++) if (fabs(xt[j]) > scale) scale = fabs(xt[j]);
    int changed = 0, worst = -1;
    double wv = 0.0;
    /* one change per pass: drop the most negative multiplier, else add the most violated row */
    for (int i = 0; i < m; i++)
      if (act[i] && ut[i] < wv) { wv = ut[i]; worst = i; }
    if (worst >= 0) { act[worst] = 0; changed = 1; }
    else {
      wv = -1e-12 * scale;
      for (int i = 0; i < m; i++) {
        if (act[i]) continue;
        double s = -d[i];
        for (int j = 0; j < n; j++) s += D[i * n + j] * xt[j];
        if (s < wv) { wv = s; worst = i; }
      }
      if (worst >= 0) { act[worst] = 1; changed = 1; }
    }
    if (!changed) {
      memcpy(x, xt, n * sizeof(double));
      memcpy(u, ut, m * sizeof(double));
      return 1;
    }
  }
  return 0;
}

int qo_ipm(int n, int m, const double* G, const double* g0, const double* D, const double* d,
           double tol, int max_iter, const double* x0, double* x, int* active, double* u, int* iterations) {
  double LG[NMAX * NMAX], H[NMAX * NMAX];
  double s[MMAX], lam[MMAX], rd[NMAX], rp[MMAX], rhs[NMAX], dx[NMAX], ds[MMAX], dl[MMAX];
  double dxa[NMAX], dsa[MMAX], dla[MMAX];
  *iterations = 0;
  for (int i = 0; i < m; i++) { active[i] = 0; u[i] = 0.0; }
  memcpy(LG, G, n * n * sizeof(double));
  if (!chol(n, LG)) return 4;
  /* 1. unconstrained minimiser */
  for (int i = 0; i < n; i++) x[i] = -g0[i];
  chol_solve(n, LG, x);
  double smin = INFINITY, scale = 1.0;
  for (int j = 0; j < n; j++) if (fabs(x[j]) > scale) scale = fabs(x[j]);
  for (int i = 0; i < m; i++) {
    double v = -d[i];
    for (int j = 0; j < n; j++) v += D[i * n + j] * x[j];
    s[i] = v;
    if (v < smin) smin = v;
  }
  if (smin >= 0.0) return 0;
  /* 2. start.  With a strictly feasible x0 (the contact-force QP always has one: every stance leg
   * pushing along its normal): s = D x0 - d > 0, multipliers centred at the scale of the gradient.
   * Without one: keep the unconstrained minimiser, push the slacks inside, centre the multipliers. */
  int feasible_start = 0;
  if (x0) {
    feasible_start = 1;
    for (int i = 0; i < m; i++) {
      double v = -d[i];
      for (int j = 0; j < n; j++) v += D[i * n + j] * x0[j];
      s[i] = v;
      if (!(v > 0.0)) feasible_start = 0;
    }
  }
  if (feasible_start) {
    double gmax = 1.0;
    for (int i = 0; i < n; i++) {
      double v = g0[i];
      for (int j = 0; j < n; j++) v += G[i * n + j] * x0[j];
      if (fabs(v) > gmax) gmax = fabs(v);
    }
    memcpy(x, x0, n * sizeof(double));
    for (int i = 0; i < m; i++) lam[i] = gmax / s[i];
  } else {
    for (int i = 0; i < m; i++) {
      double v = -d[i];
      for (int j = 0; j < n; j++) v += D[i * n + j] * x[j];
      s[i] = v;
    }
    const double shift = fmax(-1.5 * smin, 1e-2 * scale);
    double mu0 = 0.0;
    for (int i = 0; i < m; i++) { s[i] = fmax(s[i], 0.0) + shift; }
    for (int i = 0; i < m; i++) mu0 += s[i];
    mu0 /= m;
    for (int i = 0; i < m; i++) lam[i] = mu0 / s[i];
  }
  double alpha_prev = 1.0;
  int status = 2;
  for (int it = 0; it < max_iter; it++) {
    double mu = 0.0, nrd = 0.0, nrp = 0.0;
    for (int i = 0; i < m; i++) mu += s[i] * lam[i];
    mu /= m;
    for (int i = 0; i < n; i++) {
      double v = g0[i];
      for (int j = 0; j < n; j++) v += G[i * n + j] * x[j];
      for (int k = 0; k < m; k++) v -= D[k * n + i] * lam[k];
      rd[i] = v;
      if (fabs(v) > nrd) nrd = fabs(v);
    }
    for (int i = 0; i < m; i++) {
      double v = s[i] + d[i];
      for (int j = 0; j < n; j++) v -= D[i * n + j] * x[j];
      rp[i] = v; /* s - (Dx - d) */
      if (fabs(v) > nrp) nrp = fabs(v);
    }
    scale = 1.0;
    for (int j = 0; j < n; j++) if (fabs(x[j]) > scale) scale = fabs(x[j]);
    if (mu <= tol * scale && nrp <= tol * scale && nrd <= tol * scale * 100.0) { status = 0; break; }
    /* try to finish early: guess the active set, polish, verify */
    if (it >= 2 && mu <= 1e-3 * scale) {
      int guess[MMAX]; double xt[NMAX], ut[MMAX];
      for (int i = 0; i < m; i++) guess[i] = lam[i] > s[i];
      if (polish(n, m, LG, g0, D, d, guess, xt, ut)) {
        memcpy(x, xt, n * sizeof(double)); memcpy(u, ut, m * sizeof(double));
        for (int i = 0; i < m; i++) active[i] = guess[i];
        *iterations = it;
        return 0;
      }
    }
    *iterations = it + 1;
    /* H = G + D' diag(lam/s) D */
    memcpy(H, G, n * n * sizeof(double));
    for (int k = 0; k < m; k++) {
      const double th = lam[k] / s[k];
      for (int i = 0; i < n; i++) {
        const double di = D[k * n + i] * th;
        if (di == 0.0) continue;
        for (int j = 0; j < n; j++) H[i * n + j] += di * D[k * n + j];
      }
    }
    if (!chol(n, H)) { status = 4; break; }
    /* predictor: rc = s*lam */
    for (int i = 0; i < n; i++) {
      double v = -rd[i];
      for (int k = 0; k < m; k++) v -= D[k * n + i] * ((s[k] * lam[k] - lam[k] * rp[k]) / s[k]);
      rhs[i] = v;
    }
    memcpy(dxa, rhs, n * sizeof(double));
    chol_solve(n, H, dxa);
    double alpha = 1.0;
    for (int k = 0; k < m; k++) {
      double v = -rp[k];
      for (int j = 0; j < n; j++) v += D[k * n + j] * dxa[j];
      dsa[k] = v;
      dla[k] = -(s[k] * lam[k] + lam[k] * v) / s[k];
      if (dsa[k] < 0.0) alpha = fmin(alpha, -s[k] / dsa[k]);
      if (dla[k] < 0.0) alpha = fmin(alpha, -lam[k] / dla[k]);
    }
    double mua = 0.0;
    for (int k = 0; k < m; k++) mua += (s[k] + alpha * dsa[k]) * (lam[k] + alpha * dla[k]);
    mua /= m;
    double sigma = pow(mua / mu, 3.0);
    if (alpha_prev < 0.1 && sigma < 0.5) sigma = 0.5; /* short step last time: re-centre */
    /* corrector */
    for (int i = 0; i < n; i++) {
      double v = -rd[i];
      for (int k = 0; k < m; k++) {
        const double rc = s[k] * lam[k] + dsa[k] * dla[k] - sigma * mu;
        v -= D[k * n + i] * ((rc - lam[k] * rp[k]) / s[k]);
      }
      rhs[i] = v;
    }
    memcpy(dx, rhs, n * sizeof(double));
    chol_solve(n, H, dx);
    alpha = 1.0;
    double amax = INFINITY;
    for (int k = 0; k < m; k++) {
      double v = -rp[k];
      for (int j = 0; j < n; j++) v += D[k * n + j] * dx[j];
      ds[k] = v;
      const double rc = s[k] * lam[k] + dsa[k] * dla[k] - sigma * mu;
      dl[k] = -(rc + lam[k] * v) / s[k];
      if (ds[k] < 0.0) amax = fmin(amax, -s[k] / ds[k]);
      if (dl[k] < 0.0) amax = fmin(amax, -lam[k] / dl[k]);
    }
    alpha = fmin(1.0, 0.995 * amax);
    for (int tries = 0; tries < 20; tries++) { /* stay in the neighbourhood min s_i lam_i >= 1e-3 mu */
      double ps = 0.0, pm = INFINITY;
      for (int k = 0; k < m; k++) {
        const double pr = (s[k] + alpha * ds[k]) * (lam[k] + alpha * dl[k]);
        ps += pr;
        if (pr < pm) pm = pr;
      }
      if (pm >= 1e-3 * ps / m && pm > 0.0) break;
      alpha *= 0.7;
    }
    alpha_prev = alpha;
    for (int j = 0; j < n; j++) x[j] += alpha * dx[j];
    for (int k = 0; k < m; k++) { s[k] += alpha * ds[k]; lam[k] += alpha * dl[k]; }
  }
  /* 3. final polish */
  int guess[MMAX]; double xt[NMAX], ut[MMAX];
  for (int i = 0; i < m; i++) guess[i] = lam[i] > s[i];
  if (polish(n, m, LG, g0, D, d, guess, xt, ut)) {
    memcpy(x, xt, n * sizeof(double)); memcpy(u, ut, m * sizeof(double));
    for (int i = 0; i < m; i++) active[i] = guess[i];
    return status == 4 ? 4 : 0;
  }
  for (int i = 0; i < m; i++) { active[i] = guess[i]; u[i] = lam[i]; }
  return status == 0 ? 3 : status;
}
