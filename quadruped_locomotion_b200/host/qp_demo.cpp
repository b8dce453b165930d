// qp_demo.cpp - the reference's pose-optimisation known-answer test restated on the adapter classes:
// qp_solver/test/PoseOptimizationQpTest.cpp:21-52 (quadrupedSymmetricUnconstrained): nominal stance at
// z = -0.4, feet at z = -0.1  ->  P = 2 A'A = 8 I, q = -2 A'b = (0, 0, -2.4), support-polygon rows
// G x <= h, a zero equality column (PoseOptimizationQP.cpp:92-112).  Expected base position (0, 0, 0.3).
// Prints "x y z status" per case.
#include <cstdio>

#include "qlb_qp_adapter.hpp"

using namespace qlb_host;

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
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
