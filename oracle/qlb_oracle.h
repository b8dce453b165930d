/*
 * qlb_oracle.h - CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's contact-force-distribution path, used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the CHECKER.
 * Nothing under quadruped_locomotion_b200/ or include/ may include, link or call this.
 *
 * Parity status: the reference's own tests hold no golden vectors for this path
 * (SURVEY.md section 4), and the libraries it calls (OOQP+MA27, orocos-KDL, kindr, Eigen) are not
 * vendored.  What pins this oracle:
 *   - the QP solver is cross-checked against the reference's own vendored QuadProg++
 *     (qp_solver/src/QuadProg++.cc) compiled in place into oracle/_ref/ (oracle/build_oracle.py);
 *   - the known-answer vectors of SURVEY.md Appendix C/D (tests/golden/) that were produced with
 *     that solver and two independent kinematics restatements;
 *   - KKT self-certification of every solution (stationarity, feasibility, complementarity).
 * OOQP itself cannot run here; "parity with OOQP" means parity with the exact optimum it approximates.
 */
#ifndef QLB_ORACLE_H
#define QLB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QO_MAX_N 12
#define QO_MAX_M 24
#define QO_MAX_P 12

/* Same memory layout as qlb_leg_model (include/qlb.h), declared again so that the oracle does not
 * depend on product headers. */
typedef struct qo_leg_model {
  double joint_xyz[4][3];
  double joint_rpy[4][3];
  double link_mass[4];
  double link_com[4][3];
} qo_leg_model;

typedef struct qo_params {
  double S[6];     /* virtualForceWeights_ */
  double W;        /* groundForceWeight_ */
  double fmin;     /* minimalNormalGroundForce_ */
  double gravity;  /* 9.8 */
} qo_params;

/* A packed QP exactly as ContactForceDistribution assembles it (stance legs packed in LF,RF,RH,LH
 * order):  min 1/2 x'Gx + g0'x  s.t.  D x >= d.  Row-major. */
typedef struct qo_qp {
  int ns, n, m;
  int leg_of_slot[4];           /* stance slot -> leg index */
  double G[QO_MAX_N * QO_MAX_N];
  double g0[QO_MAX_N];
  double D[QO_MAX_M * QO_MAX_N];
  double d[QO_MAX_M];
  double A[6 * QO_MAX_N];       /* the 6 x n wrench map */
  double b[6];
  double foot[12], jac[36], gtau[12]; /* per leg (all four), base frame */
  int bad_input;
} qo_qp;

enum { QO_SOLVER_GI = 0, QO_SOLVER_IPM = 1, QO_SOLVER_EXTERNAL = 2 };

/* external solver hook: the reference's own QuadProg++ from oracle/_ref (CI = D', ci0 = -d) */
typedef double (*qo_external_solver)(int n, int m, const double* G, const double* g0,
                                     const double* D, const double* d, double* x);

void qo_default_params(qo_params* p);
void qo_quat_to_rot(const double quat_wxyz[4], double R_bw[9]);
void qo_rpy_to_rot(const double rpy[3], double R[9]);
void qo_leg_kinematics(const qo_leg_model* leg, const double q[3], const double grav_base[3],
                       double foot[3], double jac[9], double gtau[3]);

/* Goldfarb-Idnani dual active-set method, generic small dense form:
 *   min 1/2 x'Gx + g0'x  s.t.  CE' x + ce0 = 0 (p columns),  D x >= d  (m rows).
 * CE is n x p row-major.  Returns the optimal cost, or +inf when infeasible.
 * active[m] gets 1 for rows in the final working set; u[m] their multipliers. */
double qo_goldfarb_idnani(int n, int m, int p, const double* G, const double* g0, const double* CE,
                          const double* ce0, const double* D, const double* d, double* x,
                          int* active, double* u, int* iterations);

/* Dense Mehrotra predictor-corrector interior point + exact active-set polish: the CPU mirror of the
 * GPU algorithm and the "OOQP-style" stand-in for timing.  x0 = optional strictly feasible start
 * (NULL: start from the unconstrained minimiser).  Returns 0 ok, 2 max-iter, 3 unverified. */
int qo_ipm(int n, int m, const double* G, const double* g0, const double* D, const double* d,
           double tol, int max_iter, const double* x0, double* x, int* active, double* u, int* iterations);

/* Assemble one state.  mu may be NULL (-> mu_default); normals_world may be NULL (-> (0,0,1)). */
void qo_assemble(const qo_leg_model legs[4], const qo_params* prm, const double q[12],
                 const double quat_wxyz[4], const double wrench[6], unsigned stance_mask,
                 const double* mu, double mu_default, const double* normals_world, qo_qp* out);

/* Map a packed solution back to per-leg outputs (grf, tau, netwrench) and the flags word. */
void qo_finish(const qo_qp* qp, const double* x, const int* active, int status, int iterations,
               double grf[12], double tau[12], double netwrench[6], uint32_t* flags);

/* Whole pipeline over a batch (SoA, component-major like the C ABI).  solver = QO_SOLVER_*;
 * nsolves = 1, or 2 to mimic the reference's double solve (CFD.cpp:367 then :120; the second with
 * the equality rows C = I, c = x1).  threads <= 0 -> all cores.  margin[B] (optional) receives the
 * non-degeneracy margin min(|slack|, |multiplier|) over rows, relative to the force scale. */
int qo_solve_wrench_batch(const qo_leg_model legs[4], const qo_params* prm, long B, const double* q,
                          const double* quat, const double* wrench, const uint8_t* stance_mask,
                          const double* mu, double mu_default, const double* normals, int solver,
                          qo_external_solver ext, int nsolves, int threads, double* grf, double* tau,
                          uint32_t* flags, double* netwrench, double* margin);

/* Virtual-model controller wrench (VirtualModelController.cpp:104-268). */
typedef struct qo_vmc_params {
  double kp_t[3], kd_t[3], kff_t[3], kp_r[3], kd_r[3], kff_r[3];
  double torso_mass, leg_mass[4], leg_base_position[4][3], com[3], gravity_pct, gravity;
} qo_vmc_params;
void qo_default_vmc_params(qo_vmc_params* p);
void qo_vmc_wrench(const qo_vmc_params* p, const double base_pose[7], const double base_twist[6],
                   const double target_pose[7], const double target_twist[6], double wrench[6]);

#ifdef __cplusplus
}
#endif
#endif
