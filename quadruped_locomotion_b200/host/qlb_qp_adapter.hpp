// qlb_qp_adapter.hpp - mirror of the reference's qp_solver interface classes
// (qp_solver/include/qp_solver/quadraticproblemsolver.h:47-148, qp_solver/src/quadraticproblemsolver.cpp:
// 65-97,133-207) with the GPU generic-QP entry (qlb_qp_dense_host) behind minimize().
//
// Conventions kept from the reference:
//   - the user passes inequality constraints as  A x <= b : setGlobalInequalityConstraintJacobian stores
//     CI = -A' (quadraticproblemsolver.cpp:162-174) and setInequalityConstraintMaxValues stores ci0 = b;
//   - setGlobalEqualityConstraintJacobian(Aeq) stores Aeq as the n x p matrix CE verbatim (:175-186),
//     setEqualityConstraintMaxValues(beq) stores ce0 = beq (:198-207), i.e. CE' x + ce0 = 0;
//   - callers pass a zero n x 1 equality column (PoseOptimizationQP.cpp:106-112); the GPU entry treats
//     it as absent instead of relying on the fork's removed dependency guard (QuadProg++.cc:203-209);
//   - minimize() returns true (the reference always does, :96); the extra status() says what happened.
// Eigen is not available in this build environment: Matrix below is a minimal row-major stand-in for
// Eigen::MatrixXd / VectorXd (rows(), cols(), operator()(i,j)).
#pragma once

#include <cmath>
#include <cstdint>
#include <memory>
#include <vector>

#include "qlb.h"
#include "qlb_adapter.hpp"

namespace qlb_host {

struct Matrix {
  int r = 0, c = 0;
  std::vector<double> v;
  Matrix() = default;
  Matrix(int rows, int cols) : r(rows), c(cols), v(static_cast<size_t>(rows) * cols, 0.0) {}
  int rows() const { return r; }
  int cols() const { return c; }
  double& operator()(int i, int j) { return v[static_cast<size_t>(i) * c + j]; }
  double operator()(int i, int j) const { return v[static_cast<size_t>(i) * c + j]; }
  Matrix& setZero() { for (double& x : v) x = 0.0; return *this; }
};
using Vector = std::vector<double>;

namespace qp_solver {

class QuadraticObjectiveFunction {
 public:
  bool setGlobalHessian(const Matrix& hessian) { G_ = hessian; return true; }
  bool setLinearTerm(const Vector& jacobian) { g0_ = jacobian; return true; }
  Matrix G_;
  Vector g0_;
};

class LinearFunctionConstraints {
 public:
  bool setGlobalInequalityConstraintJacobian(const Matrix& A) {  // A x <= b  ->  CI = -A' (n x m)
    CI_ = Matrix(A.cols(), A.rows());
    for (int i = 0; i < A.rows(); i++)
      for (int j = 0; j < A.cols(); j++) CI_(j, i) = -A(i, j);
    return true;
  }
  bool setInequalityConstraintMaxValues(const Vector& b) { ci0_ = b; return true; }
  bool setGlobalEqualityConstraintJacobian(const Matrix& Aeq) { CE_ = Aeq; return true; }  // n x p, verbatim
  bool setEqualityConstraintMaxValues(const Vector& beq) { ce0_ = beq; return true; }
  Matrix CI_, CE_;
  Vector ci0_, ce0_;
};

class QuadraticProblemSolver {
 public:
  using parameters = Vector;
  explicit QuadraticProblemSolver(std::shared_ptr<Device> device) : device_(std::move(device)) {}

  // qp_solver::QuadraticProblemSolver::minimize (quadraticproblemsolver.cpp:65-97)
  bool minimize(const QuadraticObjectiveFunction& function, const LinearFunctionConstraints& constraints, parameters& params) {
    const int n = static_cast<int>(params.size());
    const int m = constraints.CI_.cols(), p = constraints.CE_.cols();
    status_ = 2;
    if (n < 1 || function.G_.rows() != n || function.G_.cols() != n || static_cast<int>(function.g0_.size()) != n) return true;
    if ((m > 0 && (constraints.CI_.rows() != n || static_cast<int>(constraints.ci0_.size()) != m)) ||
        (p > 0 && (constraints.CE_.rows() != n || static_cast<int>(constraints.ce0_.size()) != p)))
      return true;
    Vector x(n, 0.0);
    uint32_t st = 0, act = 0;
    double cost = 0.0;
    const int rc = qlb_qp_dense_host(device_->ctx(), 1, n, m, p, function.G_.v.data(), function.g0_.data(),
                                     p ? constraints.CE_.v.data() : nullptr, p ? constraints.ce0_.data() : nullptr,
                                     m ? constraints.CI_.v.data() : nullptr, m ? constraints.ci0_.data() : nullptr, x.data(),
                                     &cost, &st, &act);
    status_ = (rc == QLB_OK) ? static_cast<int>(st) : -1;
    if (rc == QLB_OK && (st == 0 || st == 3)) params = x;  // the reference copies x out unconditionally
    cost_ = cost;
    active_ = act;
    return true;
  }
  int status() const { return status_; }   // 0 ok, 1 infeasible, 2 bad problem, 3 iteration limit, -1 API error
  double cost() const { return cost_; }
  uint32_t activeSet() const { return active_; }

 private:
  std::shared_ptr<Device> device_;
  int status_ = 0;
  double cost_ = 0.0;
  uint32_t active_ = 0;
};

}  // namespace qp_solver

// Mirror of sqp_solver::SequenceQuadraticProblemSolver::minimize (qp_solver/src/sequencequadraticproblemsolver.cpp:
// 18-102): linearise the problem at the current parameters, solve the QP for the step dp on the GPU entry, apply it with
// the parameterisation's plus(), stop when |dp| < tolerance or after max_iteration steps.  The problem types are
// template parameters with the reference's method names:
//   Objective   : public qp_solver::QuadraticObjectiveFunction + getLocalHessian(H, params), getLocalGradient(G, params),
//                 computeValue(cost, params)
//   Constraints : public qp_solver::LinearFunctionConstraints + getLocalInequalityConstraintJacobian(A, params),
//                 getInequalityConstraintMaxValues(b_max), getInequalityConstraintValues(b, params)
//   Params      : getLocalSize(), getParams(), plus(result, p, dp), setParams(result)
//                 (the reference's PoseParameterization rebuilds its Pose from `result`, :71)
// Like the reference, every QP carries a zero equality column (:28-29,45-46), which the GPU entry treats as absent.
namespace sqp_solver {

template <class Objective, class Constraints, class Params>
class SequenceQuadraticProblemSolver {
 public:
  SequenceQuadraticProblemSolver(std::shared_ptr<qp_solver::QuadraticProblemSolver> quadratic_solver, double tolerance, int max_iteration)
      : quadratic_solver_(std::move(quadratic_solver)), tolerance_(tolerance), max_iteration_(max_iteration) {}

  bool minimize(Objective& objective, Constraints& constraints, Params& params) {
    const int n = params.getLocalSize();
    Matrix H, A, Aeq(n, 1);
    Vector G, b, b_max, beq(1, 0.0);
    auto linearise = [&]() {
      objective.getLocalHessian(H, params);
      objective.getLocalGradient(G, params);
      constraints.getLocalInequalityConstraintJacobian(A, params);
      constraints.getInequalityConstraintMaxValues(b_max);
      constraints.getInequalityConstraintValues(b, params);
      for (size_t i = 0; i < b.size(); i++) b[i] = b_max[i] - b[i];
      objective.setGlobalHessian(H);
      objective.setLinearTerm(G);
      constraints.setGlobalInequalityConstraintJacobian(A);
      constraints.setInequalityConstraintMaxValues(b);
      constraints.setGlobalEqualityConstraintJacobian(Aeq.setZero());
      constraints.setEqualityConstraintMaxValues(beq);
    };
    linearise();
    Vector result = params.getParams();
    iterations_ = 0;
    status_ = 0;
    while (iterations_ < max_iteration_) {
      iterations_++;
      Vector dp(n, 0.0);
      quadratic_solver_->minimize(objective, constraints, dp);
      status_ = quadratic_solver_->status();
      if (status_ != 0) break;               // the reference carries on with whatever the solver left in dp
      params.plus(result, result, dp);
      params.setParams(result);
      objective.computeValue(cost_, params);
      double norm2 = 0.0;
      for (double v : dp) norm2 += v * v;
      step_norm_ = std::sqrt(norm2);
      if (step_norm_ < tolerance_) break;
      linearise();
    }
    return true;   // the reference always returns true (:101)
  }
  int iterations() const { return iterations_; }
  int status() const { return status_; }
  double cost() const { return cost_; }
  double lastStepNorm() const { return step_norm_; }

 private:
  std::shared_ptr<qp_solver::QuadraticProblemSolver> quadratic_solver_;
  double tolerance_;
  int max_iteration_;
  int iterations_ = 0, status_ = 0;
  double cost_ = 0.0, step_norm_ = 0.0;
};

}  // namespace sqp_solver
}  // namespace qlb_host
