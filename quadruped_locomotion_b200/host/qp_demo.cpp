// qp_demo.cpp - the reference's pose-optimisation known-answer test restated on the adapter classes:
// qp_solver/test/PoseOptimizationQpTest.cpp:21-52 (quadrupedSymmetricUnconstrained): nominal stance at
// z = -0.4, feet at z = -0.1  ->  P = 2 A'A = 8 I, q = -2 A'b = (0, 0, -2.4), support-polygon rows
// G x <= h, a zero equality column (PoseOptimizationQP.cpp:92-112).  Expected base position (0, 0, 0.3).
// Prints "x y z status active" per case, then one line for the sequential-QP loop
// (sequencequadraticproblemsolver.cpp:18-102) on a small nonlinear problem with a known answer:
// min (x-2)^2 + (y-1)^2  s.t.  x^2 + y^2 <= 1  ->  (2, 1) / sqrt(5); prints "x y iterations status".
#include <cstdio>

#include "qlb_qp_adapter.hpp"

using namespace qlb_host;

namespace {
struct Point2 {
  Vector p{0.0, 0.0};
  int getLocalSize() const { return 2; }
  Vector getParams() const { return p; }
  void plus(Vector& out, const Vector& in, const Vector& dp) const { out = {in[0] + dp[0], in[1] + dp[1]}; }
  void setParams(const Vector& v) { p = v; }
};
struct Objective2 : qp_solver::QuadraticObjectiveFunction {
  void getLocalHessian(Matrix& H, const Point2&) const { H = Matrix(2, 2); H(0, 0) = 2.0; H(1, 1) = 2.0; }
  void getLocalGradient(Vector& G, const Point2& x) const { G = {2.0 * (x.p[0] - 2.0), 2.0 * (x.p[1] - 1.0)}; }
  void computeValue(double& c, const Point2& x) const { c = (x.p[0] - 2.0) * (x.p[0] - 2.0) + (x.p[1] - 1.0) * (x.p[1] - 1.0); }
};
struct Disc2 : qp_solver::LinearFunctionConstraints {
  void getLocalInequalityConstraintJacobian(Matrix& A, const Point2& x) const { A = Matrix(1, 2); A(0, 0) = 2.0 * x.p[0]; A(0, 1) = 2.0 * x.p[1]; }
  void getInequalityConstraintMaxValues(Vector& b) const { b = {1.0}; }
  void getInequalityConstraintValues(Vector& b, const Point2& x) const { b = {x.p[0] * x.p[0] + x.p[1] * x.p[1]}; }
};
}  // namespace

int main() {
  try {
    auto device = std::make_shared<Device>(QLB_MODEL_QUADRUPED_MODEL, 0);
    qp_solver::QuadraticProblemSolver solver(device);
    for (int variant = 0; variant < 2; variant++) {
      Matrix P(3, 3);
      for (int i = 0; i < 3; i++) P(i, i) = 8.0;
      // variant 1: the nominal stance is shifted by 1.5 m in x, so the optimum leaves the support polygon
      Vector q = {variant ? -12.0 : 0.0, 0.0, -2.4};
      Matrix G(4, 3);  // support rectangle |x| <= 1, |y| <= 0.5, no constraint on z
      G(0, 0) = 1.0; G(1, 0) = -1.0; G(2, 1) = 1.0; G(3, 1) = -1.0;
      Vector h = {1.0, 1.0, 0.5, 0.5};
      Matrix Aeq(3, 1);
      Vector beq = {0.0};
      qp_solver::QuadraticObjectiveFunction cost;
      qp_solver::LinearFunctionConstraints cons;
      cost.setGlobalHessian(P);
      cost.setLinearTerm(q);
      cons.setGlobalInequalityConstraintJacobian(G);
      cons.setInequalityConstraintMaxValues(h);
      cons.setGlobalEqualityConstraintJacobian(Aeq.setZero());
      cons.setEqualityConstraintMaxValues(beq);
      Vector params(3, 0.0);
      if (!solver.minimize(cost, cons, params)) return 4;
      std::printf("%.17g %.17g %.17g %d %u\n", params[0], params[1], params[2], solver.status(), solver.activeSet());
    }
    {
      auto qp = std::make_shared<qp_solver::QuadraticProblemSolver>(device);
      sqp_solver::SequenceQuadraticProblemSolver<Objective2, Disc2, Point2> sqp(qp, 1e-10, 50);
      Objective2 f; Disc2 c; Point2 x;
      sqp.minimize(f, c, x);
      std::printf("%.17g %.17g %d %d\n", x.p[0], x.p[1], sqp.iterations(), sqp.status());
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
