/*
 * qlb_oracle.c - CPU ORACLE.  TEST INFRASTRUCTURE ONLY (see qlb_oracle.h).
 *
 * Every function cites the reference lines it restates.  Paths are relative to the reference root;
 * CFD.cpp = balance_controller/src/contact_force_distribution/ContactForceDistribution.cpp,
 * QK.cpp  = quadruped_model/src/quadrupedkinematics.cpp,
 * VMC.cpp = balance_controller/src/motion_control/VirtualModelController.cpp.
 */
#include "qlb_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ small 3-vector helpers */
static void m3_mul(const double A[9], const double B[9], double C[9]) {
  double T[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(C, T, sizeof T);
}
static void m3_vec(const double A[9], const double v[3], double o[3]) {
  double t0 = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
  double t1 = A[3] * v[0] + A[4] * v[1] + A[5] * v[2];
  double t2 = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
  o[0] = t0; o[1] = t1; o[2] = t2;
}
static void m3t_vec(const double A[9], const double v[3], double o[3]) {
  double t0 = A[0] * v[0] + A[3] * v[1] + A[6] * v[2];
  double t1 = A[1] * v[0] + A[4] * v[1] + A[7] * v[2];
  double t2 = A[2] * v[0] + A[5] * v[1] + A[8] * v[2];
  o[0] = t0; o[1] = t1; o[2] = t2;
}
static void cross3(const double a[3], const double b[3], double o[3]) {
  double t0 = a[1] * b[2] - a[2] * b[1];
  double t1 = a[2] * b[0] - a[0] * b[2];
  double t2 = a[0] * b[1] - a[1] * b[0];
  o[0] = t0; o[1] = t1; o[2] = t2;
}
static double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

void qo_default_params(qo_params* p) {
  /* balance_controller/config/controller_gains.yaml:27-41 */
  const double S[6] = {1.0, 5.0, 1.0, 10.0, 10.0, 5.0};
  memcpy(p->S, S, sizeof S);
  p->W = 0.0001;
  p->fmin = 10.0;
  p->gravity = 9.8; /* CFD.cpp:518 */
}

/* kindr RotationQuaternion(w,x,y,z).rotate(v) == R v with the standard matrix of a unit Hamilton
 * quaternion; getOrientationBaseToWorld() maps base coordinates to world
 * (quadruped_state.cpp:127). */
void qo_quat_to_rot(const double q[4], double R[9]) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2.0 * (x * y - w * z);         R[2] = 2.0 * (x * z + w * y);
  R[3] = 2.0 * (x * y + w * z);         R[4] = w * w - x * x + y * y - z * z; R[5] = 2.0 * (y * z - w * x);
  R[6] = 2.0 * (x * z - w * y);         R[7] = 2.0 * (y * z + w * x);         R[8] = w * w - x * x - y * y + z * z;
}

/* URDF <origin rpy>: urdfdom turns (roll,pitch,yaw) into a quaternion, kdl_parser hands that to
 * KDL::Rotation::Quaternion; R = Rz(yaw) Ry(pitch) Rx(roll)  (SURVEY Appendix D). */
void qo_rpy_to_rot(const double rpy[3], double R[9]) {
  const double phi = 0.5 * rpy[0], the = 0.5 * rpy[1], psi = 0.5 * rpy[2];
  double x = sin(phi) * cos(the) * cos(psi) - cos(phi) * sin(the) * sin(psi);
  double y = cos(phi) * sin(the) * cos(psi) + sin(phi) * cos(the) * sin(psi);
  double z = cos(phi) * cos(the) * sin(psi) - sin(phi) * sin(the) * cos(psi);
  double w = cos(phi) * cos(the) * cos(psi) + sin(phi) * sin(the) * sin(psi);
  const double nrm = sqrt(x * x + y * y + z * z + w * w);
  x /= nrm; y /= nrm; z /= nrm; w /= nrm;
  const double qq[4] = {w, x, y, z};
  qo_quat_to_rot(qq, R);
}

/* One leg: foot position (QK.cpp:143-212 -> KDL ChainFkSolverPos_recursive), the translational
 * Jacobian = top three rows of KDL's geometric Jacobian (QK.cpp:214-278, quadruped_state.cpp:321-326)
 * and the gravity torques of KDL ChainDynParam::JntToGravity (QK.cpp:485-552):
 *   frame_k = frame_{k-1} * T(xyz_k, rpy_k) * Rz(q_k),   column k of J = z_k x (p_foot - p_k),
 *   G_k = - z_k . sum_{l>=k} (c_l - p_k) x (m_l g). */
void qo_leg_kinematics(const qo_leg_model* leg, const double q[3], const double grav[3],
                       double foot[3], double jac[9], double gtau[3]) {
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
  double zax[3][3], pj[3][3], com[4][3];
  for (int k = 0; k < 4; k++) {
    double t[3], Rk[9];
    m3_vec(R, leg->joint_xyz[k], t);
    for (int a = 0; a < 3; a++) p[a] += t[a];
    qo_rpy_to_rot(leg->joint_rpy[k], Rk);
    m3_mul(R, Rk, R);
    if (k < 3) {
      for (int a = 0; a < 3; a++) { zax[k][a] = R[3 * a + 2]; pj[k][a] = p[a]; }
      const double c = cos(q[k]), s = sin(q[k]);
      const double Rz[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
      m3_mul(R, Rz, R);
    }
    m3_vec(R, leg->link_com[k], t);
    for (int a = 0; a < 3; a++) com[k][a] = p[a] + t[a];
  }
  for (int a = 0; a < 3; a++) foot[a] = p[a];
  for (int k = 0; k < 3; k++) {
    double dv[3], col[3];
    for (int a = 0; a < 3; a++) dv[a] = foot[a] - pj[k][a];
    cross3(zax[k], dv, col);
    for (int a = 0; a < 3; a++) jac[3 * a + k] = col[a];
    double acc[3] = {0, 0, 0};
    for (int l = k; l < 4; l++) {
      double arm[3], w[3], mom[3];
      for (int a = 0; a < 3; a++) { arm[a] = com[l][a] - pj[k][a]; w[a] = leg->link_mass[l] * grav[a]; }
      cross3(arm, w, mom);
      for (int a = 0; a < 3; a++) acc[a] += mom[a];
    }
    gtau[k] = -dot3(zax[k], acc);
  }
}

/* ------------------------------------------------------------------ QP assembly (CFD.cpp:138-336) */
void qo_assemble(const qo_leg_model legs[4], const qo_params* prm, const double q[12],
                 const double quat[4], const double wrench[6], unsigned mask, const double* mu,
                 double mu_default, const double* normals_world, qo_qp* out) {
  memset(out, 0, sizeof *out);
  double Rbw[9];
  qo_quat_to_rot(quat, Rbw);
  /* gravity in base frame: getOrientationBaseToWorld().inverseRotate((0,0,-9.8)), CFD.cpp:518-519 */
  const double gw[3] = {0.0, 0.0, -prm->gravity};
  double gb[3];
  m3t_vec(Rbw, gw, gb);
  int bad = 0;
  for (int c = 0; c < 12; c++) bad |= !isfinite(q[c]);
  for (int c = 0; c < 4; c++) bad |= !isfinite(quat[c]);
  for (int c = 0; c < 6; c++) bad |= !isfinite(wrench[c]);
  for (int l = 0; l < 4; l++)
    qo_leg_kinematics(&legs[l], q + 3 * l, gb, out->foot + 3 * l, out->jac + 9 * l, out->gtau + 3 * l);

  /* prepareLegLoading, CFD.cpp:138-166: stance legs packed in LF,RF,RH,LH order */
  int ns = 0;
  for (int l = 0; l < 4; l++)
    if (mask & (1u << l)) out->leg_of_slot[ns++] = l;
  const int n = 3 * ns, m = 5 * ns;
  out->ns = ns; out->n = n; out->m = m;
  for (int c = 0; c < 6; c++) out->b[c] = wrench[c];
  if (ns == 0) { out->bad_input = bad; return; }

  /* prepareOptimization, CFD.cpp:168-206: A = [I ... I; [r_0]x ... ], W = w I */
  double* A = out->A;
  for (int k = 0; k < ns; k++) {
    const double* r = out->foot + 3 * out->leg_of_slot[k];
    for (int a = 0; a < 3; a++) A[a * n + 3 * k + a] = 1.0;
    A[3 * n + 3 * k + 1] = -r[2]; A[3 * n + 3 * k + 2] = r[1];
    A[4 * n + 3 * k + 0] = r[2];  A[4 * n + 3 * k + 2] = -r[0];
    A[5 * n + 3 * k + 0] = -r[1]; A[5 * n + 3 * k + 1] = r[0];
  }
  /* ooqpei::QuadraticProblemFormulation::solve: Q = A'SA + W, c = -A'Sb (CFD.cpp:388-407,490) */
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) {
      double s = 0.0;
      for (int r = 0; r < 6; r++) s += A[r * n + i] * prm->S[r] * A[r * n + j];
      out->G[i * n + j] = s + (i == j ? prm->W : 0.0);
    }
    double s = 0.0;
    for (int r = 0; r < 6; r++) s += A[r * n + i] * prm->S[r] * wrench[r];
    out->g0[i] = -s;
  }
  /* addMinimalForceConstraints (CFD.cpp:210-252) rows 0..ns-1; addFrictionConstraints
   * (CFD.cpp:254-336) rows ns+4k .. ns+4k+3 */
  const double ey[3] = {0.0, 1.0, 0.0};
  double eyb[3];
  m3t_vec(Rbw, ey, eyb); /* orientationControlToBase.rotate(UnitY), CFD.cpp:301-302 */
  for (int k = 0; k < ns; k++) {
    const int l = out->leg_of_slot[k];
    const double nwd[3] = {0.0, 0.0, 1.0};
    const double* nw = normals_world ? normals_world + 3 * l : nwd;
    const double mk = mu ? mu[l] : mu_default;
    double nb[3], t1[3], t2[3];
    m3t_vec(Rbw, nw, nb); /* orientationWorldToBase.rotate(normal), CFD.cpp:237,286 */
    cross3(nb, eyb, t1);
    double nr = sqrt(dot3(t1, t1));
    for (int a = 0; a < 3; a++) t1[a] /= nr; /* .normalized(), no zero check (CFD.cpp:303) */
    cross3(nb, t1, t2);
    nr = sqrt(dot3(t2, t2));
    for (int a = 0; a < 3; a++) t2[a] /= nr; /* CFD.cpp:309 */
    for (int a = 0; a < 3; a++) bad |= !isfinite(t1[a]) || !isfinite(t2[a]) || !isfinite(nb[a]);
    bad |= !isfinite(mk);
    double* row = out->D + k * n + 3 * k;
    for (int a = 0; a < 3; a++) row[a] = nb[a];
    out->d[k] = prm->fmin;
    for (int r = 0; r < 4; r++) {
      const double* t = (r < 2) ? t1 : t2;
      const double sg = (r & 1) ? -1.0 : 1.0;
      row = out->D + (ns + 4 * k + r) * n + 3 * k;
      for (int a = 0; a < 3; a++) row[a] = mk * nb[a] + sg * t[a]; /* CFD.cpp:315-325 */
      out->d[ns + 4 * k + r] = 0.0;
    }
  }
  out->bad_input = bad;
}

/* ------------------------------------------------------------------ Goldfarb-Idnani
 * Restates quadprogpp::solve_quadprog (qp_solver/src/QuadProg++.cc:52-446) and its helpers
 * (:448-760): same preprocessing (Cholesky of G, J = L^-T, cond estimate c1*c2), same choice of the
 * most violated constraint, same step-length rules and tolerances, same Givens updates of J and R.
 * Differences: flat fixed-size arrays, inequality rows given as D x >= d (CI = D', ci0 = -d), and
 * the final working set / multipliers are returned. */
typedef struct gi_work {
  int n;
  double L[QO_MAX_N][QO_MAX_N];
  double J[QO_MAX_N][QO_MAX_N];
  double R[QO_MAX_N][QO_MAX_N];
  double rnorm;
} gi_work;

static double gi_hypot(double a, double b) { /* QuadProg++.cc:640-656 */
  const double a1 = fabs(a), b1 = fabs(b);
  if (a1 > b1) { const double t = b1 / a1; return a1 * sqrt(1.0 + t * t); }
  if (b1 > a1) { const double t = a1 / b1; return b1 * sqrt(1.0 + t * t); }
  return a1 * sqrt(2.0);
}

/* rotation that maps (a, b) to (+-h, 0); returns 0 when h is numerically zero */
static int gi_givens(double a, double b, double* cc, double* ss, double* hh) {
  const double h = gi_hypot(a, b);
  if (fabs(h) < DBL_EPSILON) return 0;
  double c = a / h, s = b / h;
  if (c < 0.0) { c = -c; s = -s; *hh = -h; } else { *hh = h; }
  *cc = c; *ss = s;
  return 1;
}

static int gi_cholesky(gi_work* w, const double* G) { /* QuadProg++.cc:672-712 */
  const int n = w->n;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) w->L[i][j] = G[i * n + j];
  for (int i = 0; i < n; i++) {
    for (int j = i; j < n; j++) {
      double sum = w->L[i][j];
      for (int k = i - 1; k >= 0; k--) sum -= w->L[i][k] * w->L[j][k];
      if (i == j) {
        if (sum <= 0.0) return 0;
        w->L[i][i] = sqrt(sum);
      } else {
        w->L[j][i] = sum / w->L[i][i];
      }
    }
    for (int k = i + 1; k < n; k++) w->L[i][k] = w->L[k][i];
  }
  return 1;
}
static void gi_forward(const gi_work* w, const double* b, double* y) { /* L y = b */
  const int n = w->n;
  y[0] = b[0] / w->L[0][0];
  for (int i = 1; i < n; i++) {
    y[i] = b[i];
    for (int j = 0; j < i; j++) y[i] -= w->L[i][j] * y[j];
    y[i] = y[i] / w->L[i][i];
  }
}
static void gi_backward(const gi_work* w, const double* y, double* x) { /* L' x = y */
  const int n = w->n;
  x[n - 1] = y[n - 1] / w->L[n - 1][n - 1];
  for (int i = n - 2; i >= 0; i--) {
    x[i] = y[i];
    for (int j = i + 1; j < n; j++) x[i] -= w->L[i][j] * x[j];
    x[i] = x[i] / w->L[i][i];
  }
}
/* dvec = J' np ; z = J[:, iq:] dvec[iq:] ; r = R[0:iq,0:iq]^-1 dvec[0:iq]  (QuadProg++.cc:448-493) */
static void gi_direction(const gi_work* w, const double* np, int iq, double* dvec, double* z, double* r) {
  const int n = w->n;
  for (int i = 0; i < n; i++) {
    double s = 0.0;
    for (int j = 0; j < n; j++) s += w->J[j][i] * np[j];
    dvec[i] = s;
  }
  for (int i = 0; i < n; i++) {
    z[i] = 0.0;
    for (int j = iq; j < n; j++) z[i] += w->J[i][j] * dvec[j];
  }
  for (int i = iq - 1; i >= 0; i--) {
    double s = 0.0;
    for (int j = i + 1; j < iq; j++) s += w->R[i][j] * r[j];
    r[i] = (dvec[i] - s) / w->R[i][i];
  }
}
static int gi_add(gi_work* w, double* dvec, int* iq) { /* QuadProg++.cc:495-567 */
  const int n = w->n;
  for (int j = n - 1; j >= *iq + 1; j--) {
    double c, s, h;
    if (!gi_givens(dvec[j - 1], dvec[j], &c, &s, &h)) continue;
    dvec[j] = 0.0;
    dvec[j - 1] = h;
    const double xny = s / (1.0 + c);
    for (int k = 0; k < n; k++) {
      const double t1 = w->J[k][j - 1], t2 = w->J[k][j];
      w->J[k][j - 1] = t1 * c + t2 * s;
      w->J[k][j] = xny * (t1 + w->J[k][j - 1]) - t2;
    }
  }
  (*iq)++;
  for (int i = 0; i < *iq; i++) w->R[i][*iq - 1] = dvec[i];
  if (fabs(dvec[*iq - 1]) <= DBL_EPSILON * w->rnorm) return 0;
  if (fabs(dvec[*iq - 1]) > w->rnorm) w->rnorm = fabs(dvec[*iq - 1]);
  return 1;
}
static void gi_delete(gi_work* w, int* A, double* u, int p, int* iq, int l) { /* QuadProg++.cc:569-638 */
  const int n = w->n;
  int qq = -1;
  for (int i = p; i < *iq; i++)
    if (A[i] == l) { qq = i; break; }
  if (qq < 0) return; /* the reference throws std::invalid_argument here */
  for (int i = qq; i < *iq - 1; i++) {
    A[i] = A[i + 1];
    u[i] = u[i + 1];
    for (int j = 0; j < n; j++) w->R[j][i] = w->R[j][i + 1];
  }
  A[*iq - 1] = A[*iq];
  u[*iq - 1] = u[*iq];
  A[*iq] = 0;
  u[*iq] = 0.0;
  for (int j = 0; j < *iq; j++) w->R[j][*iq - 1] = 0.0;
  (*iq)--;
  if (*iq == 0) return;
  for (int j = qq; j < *iq; j++) {
    double c, s, h;
    if (!gi_givens(w->R[j][j], w->R[j + 1][j], &c, &s, &h)) continue;
    w->R[j + 1][j] = 0.0;
    w->R[j][j] = h;
    const double xny = s / (1.0 + c);
    for (int k = j + 1; k < *iq; k++) {
      const double t1 = w->R[j][k], t2 = w->R[j + 1][k];
      w->R[j][k] = t1 * c + t2 * s;
      w->R[j + 1][k] = xny * (t1 + w->R[j][k]) - t2;
    }
    for (int k = 0; k < n; k++) {
      const double t1 = w->J[k][j], t2 = w->J[k][j + 1];
      w->J[k][j] = t1 * c + t2 * s;
      w->J[k][j + 1] = xny * (w->J[k][j] + t1) - t2;
    }
  }
}

double qo_goldfarb_idnani(int n, int m, int p, const double* G, const double* g0, const double* CE,
                          const double* ce0, const double* D, const double* d, double* x,
                          int* active, double* u_out, int* iterations) {
  gi_work w;
  double s[QO_MAX_M + QO_MAX_P], z[QO_MAX_N], r[QO_MAX_M + QO_MAX_P], dvec[QO_MAX_N], np[QO_MAX_N];
  double u[QO_MAX_M + QO_MAX_P + 1], x_old[QO_MAX_N], u_old[QO_MAX_M + QO_MAX_P + 1];
  int A[QO_MAX_M + QO_MAX_P + 1], A_old[QO_MAX_M + QO_MAX_P + 1], iai[QO_MAX_M + QO_MAX_P];
  int iaexcl[QO_MAX_M + QO_MAX_P];
  const double inf = INFINITY;
  int iq = 0, iter = 0, ip = 0;
  double fval, ss = 0.0;

  if (active) for (int i = 0; i < m; i++) active[i] = 0;
  if (u_out) for (int i = 0; i < m; i++) u_out[i] = 0.0;
  if (iterations) *iterations = 0;
  memset(&w, 0, sizeof w);
  memset(u, 0, sizeof u);
  memset(A, 0, sizeof A);
  w.n = n;
  w.rnorm = 1.0;

  double c1 = 0.0, c2 = 0.0;
  for (int i = 0; i < n; i++) c1 += G[i * n + i];
  if (!gi_cholesky(&w, G)) return NAN; /* the reference throws std::logic_error (QuadProg++.cc:690-700) */
  for (int i = 0; i < n; i++) {
    double e[QO_MAX_N] = {0}, col[QO_MAX_N];
    e[i] = 1.0;
    gi_forward(&w, e, col);
    for (int j = 0; j < n; j++) w.J[i][j] = col[j];
    c2 += col[i];
  }
  { /* unconstrained minimiser x = -G^-1 g0 (QuadProg++.cc:164-172) */
    double y[QO_MAX_N];
    gi_forward(&w, g0, y);
    gi_backward(&w, y, x);
    for (int i = 0; i < n; i++) x[i] = -x[i];
    fval = 0.0;
    for (int i = 0; i < n; i++) fval += g0[i] * x[i];
    fval *= 0.5;
  }
  /* equality constraints enter the working set first (QuadProg++.cc:178-210) */
  for (int i = 0; i < p; i++) {
    for (int j = 0; j < n; j++) np[j] = CE[j * p + i];
    gi_direction(&w, np, iq, dvec, z, r);
    double zz = 0.0, znp = 0.0, npx = 0.0;
    for (int k = 0; k < n; k++) { zz += z[k] * z[k]; znp += z[k] * np[k]; npx += np[k] * x[k]; }
    double t2 = 0.0;
    if (fabs(zz) > DBL_EPSILON) t2 = (-npx - ce0[i]) / znp;
    for (int k = 0; k < n; k++) x[k] += t2 * z[k];
    u[iq] = t2;
    for (int k = 0; k < iq; k++) u[k] -= t2 * r[k];
    fval += 0.5 * (t2 * t2) * znp;
    A[i] = -i - 1;
    gi_add(&w, dvec, &iq); /* the fork dropped the linear-dependence guard (QuadProg++.cc:203-209) */
  }
  for (int i = 0; i < m; i++) iai[i] = i;

  enum { OUTER, PICK, DIRECTION } phase = OUTER;
  double t1, t2, t;
  int l = 0;
  for (;;) {
    if (phase == OUTER) { /* label l1, QuadProg++.cc:216-262 */
      iter++;
      if (iter > 1000) { fval = inf; break; }
      for (int i = p; i < iq; i++) iai[A[i]] = -1;
      ss = 0.0;
      ip = 0;
      double psi = 0.0;
      for (int i = 0; i < m; i++) {
        iaexcl[i] = 1;
        double sum = 0.0;
        for (int j = 0; j < n; j++) sum += D[i * n + j] * x[j];
        sum += -d[i];
        s[i] = sum;
        psi += (sum < 0.0) ? sum : 0.0;
      }
      if (fabs(psi) <= m * DBL_EPSILON * c1 * c2 * 100.0) break;
      for (int i = 0; i < iq; i++) { u_old[i] = u[i]; A_old[i] = A[i]; }
      for (int i = 0; i < n; i++) x_old[i] = x[i];
      phase = PICK;
    }
    if (phase == PICK) { /* label l2, QuadProg++.cc:264-288 (ss and ip persist across re-entries) */
      for (int i = 0; i < m; i++)
        if (s[i] < ss && iai[i] != -1 && iaexcl[i]) { ss = s[i]; ip = i; }
      if (ss >= 0.0) break;
      for (int i = 0; i < n; i++) np[i] = D[ip * n + i];
      u[iq] = 0.0;
      A[iq] = ip;
      phase = DIRECTION;
    }
    /* label l2a, QuadProg++.cc:290-338 */
    gi_direction(&w, np, iq, dvec, z, r);
    l = 0;
    t1 = inf;
    for (int k = p; k < iq; k++)
      if (r[k] > 0.0 && u[k] / r[k] < t1) { t1 = u[k] / r[k]; l = A[k]; }
    double zz = 0.0, znp = 0.0;
    for (int k = 0; k < n; k++) { zz += z[k] * z[k]; znp += z[k] * np[k]; }
    if (fabs(zz) > DBL_EPSILON) {
      t2 = -s[ip] / znp;
      if (t2 < 0) t2 = inf;
    } else {
      t2 = inf;
    }
    t = (t1 < t2) ? t1 : t2;
    if (t >= inf) { fval = inf; break; } /* infeasible, QuadProg++.cc:340-345 */
    if (t2 >= inf) { /* dual step only, QuadProg++.cc:347-363 */
      for (int k = 0; k < iq; k++) u[k] -= t * r[k];
      u[iq] += t;
      iai[l] = l;
      gi_delete(&w, A, u, p, &iq, l);
      continue; /* DIRECTION again */
    }
    for (int k = 0; k < n; k++) x[k] += t * z[k]; /* primal + dual step, QuadProg++.cc:367-385 */
    fval += t * znp * (0.5 * t + u[iq]);
    for (int k = 0; k < iq; k++) u[k] -= t * r[k];
    u[iq] += t;
    if (fabs(t - t2) < DBL_EPSILON) { /* full step, QuadProg++.cc:387-424 */
      if (!gi_add(&w, dvec, &iq)) {
        iaexcl[ip] = 0;
        gi_delete(&w, A, u, p, &iq, ip);
        for (int i = 0; i < m; i++) iai[i] = i;
        for (int i = p; i < iq; i++) { A[i] = A_old[i]; u[i] = u_old[i]; iai[A[i]] = -1; }
        for (int i = 0; i < n; i++) x[i] = x_old[i];
        phase = PICK;
      } else {
        iai[ip] = -1;
        phase = OUTER;
      }
      continue;
    }
    /* partial step: drop constraint l, refresh s[ip] (QuadProg++.cc:426-444) */
    iai[l] = l;
    gi_delete(&w, A, u, p, &iq, l);
    double sum = 0.0;
    for (int k = 0; k < n; k++) sum += D[ip * n + k] * x[k];
    s[ip] = sum - d[ip];
    phase = DIRECTION;
  }
  if (iterations) *iterations = iter;
  if (isfinite(fval)) {
    for (int i = p; i < iq; i++) {
      if (active) active[A[i]] = 1;
      if (u_out) u_out[A[i]] = u[i];
    }
  }
  return fval;
}

/* ------------------------------------------------------------------ output mapping
 * x -> per-leg ground-reaction forces; desiredContactForce_ = -x (CFD.cpp:496-511);
 * tau = J'(-x) + G(q) for stance legs (CFD.cpp:516-578), swing legs reported as zero;
 * net wrench = A x (CFD.cpp:614-625).  Flags word as in include/qlb.h. */
void qo_finish(const qo_qp* qp, const double* x, const int* active, int status, int iterations,
               double grf[12], double tau[12], double netwrench[6], uint32_t* flags) {
  for (int c = 0; c < 12; c++) { grf[c] = 0.0; tau[c] = 0.0; }
  for (int c = 0; c < 6; c++) netwrench[c] = 0.0;
  uint32_t fl = 0;
  const int n = qp->n, ns = qp->ns;
  const int solved = (status == 0 || status == 2 || status == 3);
  for (int k = 0; k < ns; k++) {
    const int l = qp->leg_of_slot[k];
    fl |= 1u << l;
    if (!solved) continue;
    for (int a = 0; a < 3; a++) grf[3 * l + a] = x[3 * k + a];
    const double* J = qp->jac + 9 * l;
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int a = 0; a < 3; a++) s += J[3 * a + j] * (-x[3 * k + a]);
      tau[3 * l + j] = s + qp->gtau[3 * l + j];
    }
    if (active) {
      if (active[k]) fl |= 1u << (4 + 5 * l);
      for (int r = 0; r < 4; r++)
        if (active[ns + 4 * k + r]) fl |= 1u << (4 + 5 * l + 1 + r);
    }
  }
  if (solved)
    for (int r = 0; r < 6; r++) {
      double s = 0.0;
      for (int j = 0; j < n; j++) s += qp->A[r * n + j] * x[j];
      netwrench[r] = s;
    }
  fl |= ((uint32_t)status & 7u) << 24;
  fl |= (uint32_t)(iterations > 31 ? 31 : (iterations < 0 ? 0 : iterations)) << 27;
  *flags = fl;
}

/* non-degeneracy margin of a solution: how far the closest row is from flipping between active and
 * inactive.  Inactive rows: slack relative to the force scale.  Active rows: multiplier relative to
 * W * force scale - releasing a row with multiplier u moves x by up to u / lambda_min(G) = u / W, so
 * this is the same "relative change of x" unit. */
static double qo_margin(const qo_qp* qp, const double* x, const int* active, const double* u, double W) {
  const int n = qp->n, m = qp->m;
  double scale = 1.0, best = INFINITY;
  for (int j = 0; j < n; j++) if (fabs(x[j]) > scale) scale = fabs(x[j]);
  for (int i = 0; i < m; i++) {
    double s = -qp->d[i];
    for (int j = 0; j < n; j++) s += qp->D[i * n + j] * x[j];
    const double v = active[i] ? fabs(u[i]) / W : fabs(s);
    if (v < best) best = v;
  }
  return best / scale;
}

/* classify rows of an externally solved QP (the reference solver does not export its working set) */
static void qo_classify(const qo_qp* qp, const double* x, int* active, double* u) {
  const int n = qp->n, m = qp->m;
  double scale = 1.0;
  for (int j = 0; j < n; j++) if (fabs(x[j]) > scale) scale = fabs(x[j]);
  for (int i = 0; i < m; i++) {
    double s = -qp->d[i];
    for (int j = 0; j < n; j++) s += qp->D[i * n + j] * x[j];
    active[i] = fabs(s) <= 1e-9 * scale;
    u[i] = 0.0;
  }
}

static int qo_solve_one(const qo_qp* qp, int solver, qo_external_solver ext, int nsolves, double* x,
                        int* active, double* u, int* iters) {
  const int n = qp->n, m = qp->m;
  int status = 0;
  *iters = 0;
  if (solver == QO_SOLVER_GI) {
    const double f = qo_goldfarb_idnani(n, m, 0, qp->G, qp->g0, NULL, NULL, qp->D, qp->d, x, active, u, iters);
    if (isinf(f) && f > 0.0) status = 5;       /* the solver's "infeasible" return (QuadProg++.cc:340-344) */
    else if (!isfinite(f)) status = 4;
    if (status == 0 && nsolves == 2) {
      /* addDesiredLegLoadConstraints: second solve with C = I, c = x1 (CFD.cpp:369-381,120) */
      double CE[QO_MAX_N * QO_MAX_N] = {0}, ce0[QO_MAX_N], x2[QO_MAX_N], u2[QO_MAX_M];
      int act2[QO_MAX_M], it2;
      for (int i = 0; i < n; i++) { CE[i * n + i] = 1.0; ce0[i] = -x[i]; }
      const double f2 = qo_goldfarb_idnani(n, m, n, qp->G, qp->g0, CE, ce0, qp->D, qp->d, x2, act2, u2, &it2);
      if (isfinite(f2)) memcpy(x, x2, n * sizeof(double));
    }
  } else if (solver == QO_SOLVER_IPM) {
    /* strictly feasible start: every stance leg pushes c along its normal (rows 0..ns-1 of D) */
    double x0[QO_MAX_N], fn = 0.0;
    for (int a = 0; a < 3; a++) fn += qp->b[a] * qp->D[a];
    const double c = fmax(2.0 * qp->d[0], fn / qp->ns);
    for (int j = 0; j < n; j++) {
      x0[j] = 0.0;
      for (int k = 0; k < qp->ns; k++) x0[j] += c * qp->D[k * n + j];
    }
    status = qo_ipm(n, m, qp->G, qp->g0, qp->D, qp->d, 1e-9, 40, x0, x, active, u, iters);
    if (nsolves == 2) {
      double x2[QO_MAX_N], u2[QO_MAX_M]; int a2[QO_MAX_M], it2;
      qo_ipm(n, m, qp->G, qp->g0, qp->D, qp->d, 1e-9, 40, x0, x2, a2, u2, &it2);
    }
  } else {
    for (int rep = 0; rep < nsolves; rep++) {
      const double f = ext(n, m, qp->G, qp->g0, qp->D, qp->d, x);
      if (isinf(f) && f > 0.0) status = 5;
      else if (!isfinite(f)) status = 4;
    }
    if (status == 0) qo_classify(qp, x, active, u);
  }
  return status;
}

int qo_solve_wrench_batch(const qo_leg_model legs[4], const qo_params* prm, long B, const double* q,
                          const double* quat, const double* wrench, const uint8_t* stance_mask,
                          const double* mu, double mu_default, const double* normals, int solver,
                          qo_external_solver ext, int nsolves, int threads, double* grf, double* tau,
                          uint32_t* flags, double* netwrench, double* margin) {
  if (solver == QO_SOLVER_EXTERNAL && !ext) return -1;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#else
  (void)threads;
#endif
#pragma omp parallel for schedule(static)
  for (long i = 0; i < B; i++) {
    double qi[12], qu[4], wr[6], mui[4], nrm[12];
    for (int c = 0; c < 12; c++) qi[c] = q[c * B + i];
    for (int c = 0; c < 4; c++) qu[c] = quat[c * B + i];
    for (int c = 0; c < 6; c++) wr[c] = wrench[c * B + i];
    if (mu) for (int c = 0; c < 4; c++) mui[c] = mu[c * B + i];
    if (normals) for (int c = 0; c < 12; c++) nrm[c] = normals[c * B + i];
    qo_qp qp;
    qo_assemble(legs, prm, qi, qu, wr, stance_mask[i] & 0xF, mu ? mui : NULL, mu_default,
                normals ? nrm : NULL, &qp);
    double x[QO_MAX_N] = {0}, u[QO_MAX_M] = {0};
    int active[QO_MAX_M] = {0}, iters = 0, status;
    if (qp.bad_input) status = 4;
    else if (qp.ns == 0) status = 1;
    else status = qo_solve_one(&qp, solver, ext, nsolves, x, active, u, &iters);
    double g[12], t[12], nw[6];
    uint32_t fl;
    qo_finish(&qp, x, active, status, solver == QO_SOLVER_IPM ? iters : 0, g, t, nw, &fl);
    for (int c = 0; c < 12; c++) { grf[c * B + i] = g[c]; tau[c * B + i] = t[c]; }
    if (netwrench) for (int c = 0; c < 6; c++) netwrench[c * B + i] = nw[c];
    flags[i] = fl;
    if (margin) margin[i] = (status == 0 && qp.ns > 0) ? qo_margin(&qp, x, active, u, prm->W) : INFINITY;
  }
  return 0;
}
