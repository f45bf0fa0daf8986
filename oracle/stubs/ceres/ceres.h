// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <ceres/ceres.h> (Ceres 1.13/1.14, un-vendored dependency of the reference,
// CMakeLists.txt:32 `find_package(Ceres)`).  Problem / CostFunction / LossFunction / LocalParameterization carry Ceres' public
// interface so that /root/reference/src/lidarOptimization.cpp and src/odomEstimationClass.cpp compile unmodified; Solve() runs the
// trust-region Levenberg-Marquardt restatement of oracle/ceres_lm.cpp and calls the REFERENCE's own
// EdgeAnalyticCostFunction / SurfNormAnalyticCostFunction / PoseSE3Parameterization through their virtual functions.
// Supported shape = what the path builds: ONE parameter block of global size 7 / local size 6, residual blocks of dimension 1.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <set>
#include <vector>
#include "../../floam_oracle.h"

namespace ceres {

class CostFunction {
 public:
  CostFunction() : num_residuals_(0) {}
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  const std::vector<int>& parameter_block_sizes() const { return parameter_block_sizes_; }
  int num_residuals() const { return num_residuals_; }
 protected:
  std::vector<int>* mutable_parameter_block_sizes() { return &parameter_block_sizes_; }
  void set_num_residuals(int n) { num_residuals_ = n; }
 private:
  std::vector<int> parameter_block_sizes_;
  int num_residuals_;
};
template <int kNumResiduals, int N0 = 0, int N1 = 0, int N2 = 0>
class SizedCostFunction : public CostFunction {
 public:
  SizedCostFunction() {
    set_num_residuals(kNumResiduals);
    if (N0) mutable_parameter_block_sizes()->push_back(N0);
    if (N1) mutable_parameter_block_sizes()->push_back(N1);
    if (N2) mutable_parameter_block_sizes()->push_back(N2);
  }
  virtual ~SizedCostFunction() {}
};

class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};
class HuberLoss : public LossFunction {  // ceres/loss_function.cc
 public:
  explicit HuberLoss(double a) : a_(a), b_(a * a) {}
  virtual void Evaluate(double s, double rho[3]) const {
    if (s > b_) {
      const double r = sqrt(s);
      rho[0] = 2.0 * a_ * r - b_;
      rho[1] = std::max(std::numeric_limits<double>::min(), a_ / r);
      rho[2] = -rho[1] / (2.0 * s);
    } else {
      rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
    }
  }
 private:
  const double a_, b_;
};
class CauchyLoss : public LossFunction {
 public:
  explicit CauchyLoss(double a) : b_(a * a), c_(1 / b_) {}
  virtual void Evaluate(double s, double rho[3]) const {
    const double sum = 1.0 + s * c_;
    const double inv = 1.0 / sum;
    rho[0] = b_ * log(sum);
    rho[1] = std::max(std::numeric_limits<double>::min(), inv);
    rho[2] = -c_ * (inv * inv);
  }
 private:
  const double b_, c_;
};

class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
  virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
  virtual bool ComputeJacobian(const double* x, double* jacobian) const = 0;
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
};

enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };

class Problem {
 public:
  struct Options {};
  Problem() : x_(NULL), parameterization_(NULL) {}
  explicit Problem(const Options&) : x_(NULL), parameterization_(NULL) {}
  ~Problem() {  // default ownership: the problem deletes cost functions, loss functions (once each) and parameterizations
    std::set<LossFunction*> losses;
    for (size_t i = 0; i < blocks_.size(); ++i) { delete blocks_[i].cost; if (blocks_[i].loss) losses.insert(blocks_[i].loss); }
    for (std::set<LossFunction*>::iterator it = losses.begin(); it != losses.end(); ++it) delete *it;
    delete parameterization_;
  }
  void AddParameterBlock(double* values, int size, LocalParameterization* local_parameterization) {
    if (x_ != NULL || size != 7 || local_parameterization->GlobalSize() != 7 || local_parameterization->LocalSize() != 6) unsupported("AddParameterBlock");
    x_ = values;
    parameterization_ = local_parameterization;
  }
  void AddResidualBlock(CostFunction* cost_function, LossFunction* loss_function, double* x0) {
    if (x0 != x_ || cost_function->num_residuals() != 1 || cost_function->parameter_block_sizes().size() != 1 || cost_function->parameter_block_sizes()[0] != 7)
      unsupported("AddResidualBlock");
    Block b; b.cost = cost_function; b.loss = loss_function;
    blocks_.push_back(b);
  }
  int NumResidualBlocks() const { return (int)blocks_.size(); }

  struct Block { CostFunction* cost; LossFunction* loss; };
  double* x_;
  LocalParameterization* parameterization_;
  std::vector<Block> blocks_;
 private:
  static void unsupported(const char* what) { std::fprintf(stderr, "ceres stand-in: %s outside the supported shape (one 7/6 block, 1-d residuals)\n", what); std::abort(); }
  Problem(const Problem&);
  void operator=(const Problem&);
};

struct Solver {
  struct Options {
    Options() : linear_solver_type(DENSE_QR), max_num_iterations(50), minimizer_progress_to_stdout(false), check_gradients(false),
                gradient_check_relative_precision(1e-8), num_threads(1) {}
    LinearSolverType linear_solver_type;
    int max_num_iterations;
    bool minimizer_progress_to_stdout;
    bool check_gradients;
    double gradient_check_relative_precision;
    int num_threads;
  };
  struct Summary {
    Summary() : initial_cost(0), final_cost(0), num_successful_steps(0), num_unsuccessful_steps(0) {}
    double initial_cost, final_cost;
    int num_successful_steps, num_unsuccessful_steps;
    fo::LmSummary lm;  // stand-in only: what the trust-region loop did
    std::string BriefReport() const { return "ceres stand-in"; }
  };
};

namespace floam_stub {
// every Solve() appends its summary here when recording is on (read back through oracle/ref_capi.cpp)
inline std::vector<fo::LmSummary>*& solve_log() { static std::vector<fo::LmSummary>* v = NULL; return v; }

struct Program : fo::LmProgram {  // ProgramEvaluator + ResidualBlock::Evaluate + Corrector over the problem's blocks
  Problem* p;
  explicit Program(Problem* problem) : p(problem) {}
  size_t num_residuals() const { return p->blocks_.size(); }
  bool evaluate(const double x[7], double* cost, std::vector<double>* residuals, std::vector<double>* jacobian, double gradient[6]) {
    const size_t C = p->blocks_.size();
    double total = 0.0;
    if (residuals) residuals->resize(C);
    if (jacobian) jacobian->resize(C * 6);
    if (gradient) for (int j = 0; j < 6; ++j) gradient[j] = 0.0;
    double P[7 * 6];  // ParameterBlock::UpdateLocalParameterizationJacobian: row-major GlobalSize x LocalSize, once per state
    if (jacobian && !p->parameterization_->ComputeJacobian(x, P)) return false;
    for (size_t i = 0; i < C; ++i) {
      double r = std::numeric_limits<double>::quiet_NaN(), jac7[7];
      for (int j = 0; j < 7; ++j) jac7[j] = std::numeric_limits<double>::quiet_NaN();  // InvalidateEvaluation
      double const* params[1] = {x};
      double* jacs[1] = {jac7};
      if (!p->blocks_[i].cost->Evaluate(params, &r, jacobian ? jacs : NULL)) return false;
      bool ok = std::isfinite(r);  // IsEvaluationValid
      if (jacobian) for (int j = 0; j < 7; ++j) ok = ok && std::isfinite(jac7[j]);
      if (!ok) return false;
      double jl[6];
      if (jacobian) {  // local Jacobian = global (1x7) * P (7x6): MatrixMatrixMultiply, inner index ascending
        for (int j = 0; j < 6; ++j) {
          double s = 0.0;
          for (int k = 0; k < 7; ++k) s += jac7[k] * P[k * 6 + j];
          jl[j] = s;
        }
      }
      const double squared_norm = r * r;
      const LossFunction* loss = p->blocks_[i].loss;
      if (loss == NULL) {
        total += 0.5 * squared_norm;
      } else {
        double rho[3];
        loss->Evaluate(squared_norm, rho);
        total += 0.5 * rho[0];
        const double sqrt_rho1 = sqrt(rho[1]);  // Corrector
        if (squared_norm == 0.0 || rho[2] <= 0.0) {
          if (jacobian) for (int j = 0; j < 6; ++j) jl[j] *= sqrt_rho1;
          r *= sqrt_rho1;
        } else {
          const double D = 1.0 + 2.0 * squared_norm * rho[2] / rho[1];
          const double alpha = 1.0 - sqrt(D);
          const double residual_scaling = sqrt_rho1 / (1 - alpha);
          const double alpha_sq_norm = alpha / squared_norm;
          if (jacobian) {
            double rtj[6];
            for (int j = 0; j < 6; ++j) rtj[j] = r * jl[j];
            for (int j = 0; j < 6; ++j) jl[j] = sqrt_rho1 * (jl[j] - alpha_sq_norm * r * rtj[j]);
          }
          r *= residual_scaling;
        }
      }
      if (residuals) (*residuals)[i] = r;
      if (jacobian) for (int j = 0; j < 6; ++j) (*jacobian)[i * 6 + j] = jl[j];
      if (gradient && jacobian) for (int j = 0; j < 6; ++j) gradient[j] += jl[j] * r;
    }
    *cost = total;
    return true;
  }
  void plus(const double x[7], const double delta[6], double out[7]) { p->parameterization_->Plus(x, delta, out); }
};
}  // namespace floam_stub

inline void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary) {
  if (options.linear_solver_type != DENSE_QR || options.check_gradients) { std::fprintf(stderr, "ceres stand-in: only DENSE_QR without gradient checking\n"); std::abort(); }
  floam_stub::Program program(problem);
  fo::LmSummary lm;
  fo::trust_region_lm(program, problem->x_, &lm, options.max_num_iterations);
  if (summary) {
    summary->lm = lm;
    summary->initial_cost = lm.initial_cost; summary->final_cost = lm.final_cost;
    summary->num_successful_steps = lm.accepted; summary->num_unsuccessful_steps = lm.iterations - lm.accepted;
  }
  if (floam_stub::solve_log()) floam_stub::solve_log()->push_back(lm);
}

}  // namespace ceres
