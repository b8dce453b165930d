/*
 * qlb_oracle_vmc.c - CPU ORACLE (TEST INFRASTRUCTURE ONLY): the virtual-model controller wrench.
 * Restates VirtualModelController::computeError / computeGravityCompensation / computeVirtualForce /
 * computeVirtualTorque (balance_controller/src/motion_control/VirtualModelController.cpp:104-268)
 * with the kindr conventions of SURVEY.md Appendix D (kindr is not vendored: "parity unpinned").
 */
#include <math.h>
#include <string.h>

#include "qlb_oracle.h"

void qo_default_vmc_params(qo_vmc_params* p) {
  /* balance_controller/config/controller_gains.yaml:3-26 (heading, lateral, vertical / roll, pitch, yaw) */
  const double kp_t[3] = {5000, 5000, 10000}, kd_t[3] = {5000, 4000, 5000}, kff_t[3] = {10, 10, 100};
  const double kp_r[3] = {10000, 10000, 4000}, kd_r[3] = {1000, 1000, 1000}, kff_r[3] = {0.2, 0.2, 1000};
  /* quadruped_state.cpp:83-97 (LF, RF, RH, LH) */
  const double legpos[4][3] = {{0.42, 0.075, 0.0}, {0.42, -0.075, 0.0}, {-0.42, -0.075, 0.0}, {-0.42, 0.075, 0.0}};
  memcpy(p->kp_t, kp_t, sizeof kp_t); memcpy(p->kd_t, kd_t, sizeof kd_t); memcpy(p->kff_t, kff_t, sizeof kff_t);
  memcpy(p->kp_r, kp_r, sizeof kp_r); memcpy(p->kd_r, kd_r, sizeof kd_r); memcpy(p->kff_r, kff_r, sizeof kff_r);
  p->torso_mass = 27.0; /* quadruped_state.cpp:28 */
  for (int l = 0; l < 4; l++) { p->leg_mass[l] = 6.0; memcpy(p->leg_base_position[l], legpos[l], sizeof legpos[l]); }
  p->com[0] = p->com[1] = p->com[2] = 0.0;
  p->gravity_pct = 1.0; /* VirtualModelController.cpp:57 */
  p->gravity = 9.8;
}

static void rot_t_vec(const double R[9], const double v[3], double o[3]) { /* R' v */
  const double a = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  const double b = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  const double c = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  o[0] = a; o[1] = b; o[2] = c;
}
static void cross(const double a[3], const double b[3], double o[3]) {
  const double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
/* Hamilton product a*b, (w,x,y,z) */
static void qmul(const double a[4], const double b[4], double o[4]) {
  o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}
/* kindr logarithmic map of a unit quaternion: rotation vector, angle in [0, pi] */
static void qlog(const double qin[4], double v[3]) {
  double q[4] = {qin[0], qin[1], qin[2], qin[3]};
  if (q[0] < 0.0) for (int i = 0; i < 4; i++) q[i] = -q[i];
  const double n = sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < 1e-12) { v[0] = 2.0 * q[1]; v[1] = 2.0 * q[2]; v[2] = 2.0 * q[3]; return; }
  const double k = 2.0 * atan2(n, q[0]) / n;
  v[0] = k * q[1]; v[1] = k * q[2]; v[2] = k * q[3];
}

void qo_vmc_wrench(const qo_vmc_params* p, const double pose[7], const double twist[6],
                   const double tpose[7], const double ttwist[6], double wrench[6]) {
  double Rbw[9];
  qo_quat_to_rot(pose + 3, Rbw);
  /* computeError, VMC.cpp:104-160 */
  double ep[3], ev[3], ew[3], eR[3];
  for (int a = 0; a < 3; a++) { ep[a] = tpose[a] - pose[a]; ev[a] = ttwist[a] - twist[a]; ew[a] = ttwist[3 + a] - twist[3 + a]; }
  {
    /* orientationError_ = -(q*^-1).boxMinus(q^-1) = -log(q*^-1 q)  (VMC.cpp:120,124) */
    const double qt_inv[4] = {tpose[3], -tpose[4], -tpose[5], -tpose[6]};
    double rel[4], lv[3];
    qmul(qt_inv, pose + 3, rel);
    qlog(rel, lv);
    for (int a = 0; a < 3; a++) eR[a] = -lv[a];
  }
  /* computeGravityCompensation, VMC.cpp:162-188 */
  const double gw[3] = {0.0, 0.0, -p->gravity};
  double gb[3], Fg[3], Tg[3], ft[3];
  rot_t_vec(Rbw, gw, gb);
  for (int a = 0; a < 3; a++) { ft[a] = -p->gravity_pct * p->torso_mass * gb[a]; Fg[a] = ft[a]; }
  cross(p->com, ft, Tg);
  for (int l = 0; l < 4; l++) {
    double fl[3], r[3], t[3];
    for (int a = 0; a < 3; a++) { fl[a] = -p->gravity_pct * p->leg_mass[l] * gb[a]; Fg[a] += fl[a]; r[a] = p->leg_base_position[l][a] - p->com[a]; }
    cross(r, fl, t);
    for (int a = 0; a < 3; a++) Tg[a] += t[a];
  }
  /* computeVirtualForce, VMC.cpp:191-239 (orientationWorldToControl evaluates to identity, :206) */
  double epb[3], evb[3], ffb[3], gfb[3], gdb[3];
  const double ff[3] = {ttwist[0], ttwist[1], 0.0};
  const double gfw[3] = {0.0, 0.0, p->kp_t[2] * ep[2]};
  const double gdw[3] = {0.0, 0.0, p->kd_t[2] * ev[2]};
  rot_t_vec(Rbw, ep, epb); rot_t_vec(Rbw, ev, evb); rot_t_vec(Rbw, ff, ffb);
  rot_t_vec(Rbw, gfw, gfb); rot_t_vec(Rbw, gdw, gdb);
  for (int a = 0; a < 3; a++)
    wrench[a] = p->kp_t[a] * epb[a] + p->kd_t[a] * evb[a] + p->kff_t[a] * ffb[a] + Fg[a] + gfb[a] + gdb[a];
  /* computeVirtualTorque, VMC.cpp:242-268 */
  double dw[3], fw[3], dwb[3], fwb[3];
  for (int a = 0; a < 3; a++) dw[a] = p->kd_r[a] * ew[a];
  fw[0] = 0.0; fw[1] = 0.0; fw[2] = p->kff_r[2] * ttwist[5];
  rot_t_vec(Rbw, dw, dwb); rot_t_vec(Rbw, fw, fwb);
  for (int a = 0; a < 3; a++) wrench[3 + a] = p->kp_r[a] * eR[a] + dwb[a] + fwb[a] + Tg[a];
}
