// ORACLE — TEST INFRASTRUCTURE ONLY (see types.h).  The cost functions / parameterisation are pinned the same way; the trust-region loop restates Ceres (un-vendored, absent here): unpinned.
// Restates src/lidarOptimization.cpp:12-152 (EdgeAnalyticCostFunction, SurfNormAnalyticCostFunction,
// PoseSE3Parameterization, getTransformFromSe3) and the Ceres 1.13/1.14 trust-region Levenberg-Marquardt loop
// with DENSE_QR as configured at src/odomEstimationClass.cpp:83-108 (un-vendored; SURVEY.md Appendix A.5:
// TrustRegionMinimizer, LevenbergMarquardtStrategy, DenseQRSolver, ResidualBlock::Evaluate, Corrector, HuberLoss).
#include "floam_oracle.h"
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <limits>

namespace fo {

// src/lidarOptimization.cpp:101-137
void getTransformFromSe3(const double se3[6], Quat& q, Vec3& t) {
  Vec3 omega{se3[0], se3[1], se3[2]};
  Vec3 upsilon{se3[3], se3[4], se3[5]};
  Mat3 Omega = {{{0, -omega.z, omega.y}, {omega.z, 0, -omega.x}, {-omega.y, omega.x, 0}}};
  double theta = norm(omega);
  double half_theta = 0.5 * theta;
  double imag_factor;
  double real_factor = std::cos(half_theta);
  if (theta < 1e-10) {
    double theta_sq = theta * theta;
    double theta_po4 = theta_sq * theta_sq;
    imag_factor = 0.5 - 0.0208333 * theta_sq + 0.000260417 * theta_po4;
  } else {
    double sin_half_theta = std::sin(half_theta);
    imag_factor = sin_half_theta / theta;
  }
  q = Quat{imag_factor * omega.x, imag_factor * omega.y, imag_factor * omega.z, real_factor};
  Mat3 J;
  if (theta < 1e-10) {
    J = quat_to_matrix(q);
  } else {
    Mat3 Omega2 = mat3_mul(Omega, Omega);
    double c1 = (1 - std::cos(theta)) / (theta * theta);
    double c2 = (theta - std::sin(theta)) / (std::pow(theta, 3));
    Mat3 I = mat3_identity();
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) J.m[i][j] = I.m[i][j] + c1 * Omega.m[i][j] + c2 * Omega2.m[i][j];
  }
  t = mat3_apply(J, upsilon);
}

// PoseSE3Parameterization::Plus, src/lidarOptimization.cpp:77-91
void se3_plus(const double x[7], const double delta[6], double out[7]) {
  Quat delta_q;
  Vec3 delta_t;
  getTransformFromSe3(delta, delta_q, delta_t);
  Quat quater{x[0], x[1], x[2], x[3]};
  Vec3 trans{x[4], x[5], x[6]};
  Quat qp = quat_mul(delta_q, quater);
  Vec3 tp = quat_rotate(delta_q, trans) + delta_t;
  out[0] = qp.x; out[1] = qp.y; out[2] = qp.z; out[3] = qp.w;
  out[4] = tp.x; out[5] = tp.y; out[6] = tp.z;
}

// EdgeAnalyticCostFunction::Evaluate :12-43 and SurfNormAnalyticCostFunction::Evaluate :51-74
bool evaluate_residual(const Residual& rb, const double x[7], double* r, double* jac7) {
  Quat q{x[0], x[1], x[2], x[3]};
  Vec3 t{x[4], x[5], x[6]};
  if (rb.kind == 0) {
    Vec3 lp = quat_rotate(q, rb.curr_point) + t;
    Vec3 nu = cross(lp - rb.a, lp - rb.b);
    Vec3 de = rb.a - rb.b;
    double de_norm = norm(de);
    *r = norm(nu) / de_norm;
    if (jac7) {
      // J = -nu^T/|nu| * skew(de) * [-skew(lp) | I] / |de|
      double nun = norm(nu);
      Vec3 w{-nu.x / nun, -nu.y / nun, -nu.z / nun};
      // row vector w^T * skew(de) = (de x w)^T ... computed as the explicit product to keep Eigen's order
      Mat3 skew_de = {{{0, -de.z, de.y}, {de.z, 0, -de.x}, {-de.y, de.x, 0}}};
      double wS[3];
      for (int j = 0; j < 3; ++j) wS[j] = w.x * skew_de.m[0][j] + w.y * skew_de.m[1][j] + w.z * skew_de.m[2][j];
      Mat3 nskew_lp = {{{0, lp.z, -lp.y}, {-lp.z, 0, lp.x}, {lp.y, -lp.x, 0}}};  // -skew(lp)
      for (int j = 0; j < 3; ++j)
        jac7[j] = (wS[0] * nskew_lp.m[0][j] + wS[1] * nskew_lp.m[1][j] + wS[2] * nskew_lp.m[2][j]) / de_norm;
      for (int j = 0; j < 3; ++j) jac7[3 + j] = wS[j] / de_norm;
      jac7[6] = 0.0;
    }
  } else {
    Vec3 point_w = quat_rotate(q, rb.curr_point) + t;
    *r = dot(rb.a, point_w) + rb.b.x;
    if (jac7) {
      Mat3 nskew = {{{0, point_w.z, -point_w.y}, {-point_w.z, 0, point_w.x}, {point_w.y, -point_w.x, 0}}};
      for (int j = 0; j < 3; ++j) jac7[j] = rb.a.x * nskew.m[0][j] + rb.a.y * nskew.m[1][j] + rb.a.z * nskew.m[2][j];
      jac7[3] = rb.a.x; jac7[4] = rb.a.y; jac7[5] = rb.a.z;
      jac7[6] = 0.0;
    }
  }
  bool ok = std::isfinite(*r);
  if (jac7) for (int j = 0; j < 7; ++j) ok = ok && std::isfinite(jac7[j]);
  return ok;
}

namespace {

// ceres loss functions: rho[0..2] at s
void loss_evaluate(LossKind kind, double s, double rho[3]) {
  if (kind == LOSS_HUBER) {  // ceres::HuberLoss(0.1), src/odomEstimationClass.cpp:86
    const double a = 0.1, b = a * a;
    if (s > b) {
      const double r = std::sqrt(s);
      rho[0] = 2.0 * a * r - b;
      rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
      rho[2] = -rho[1] / (2.0 * s);
    } else {
      rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
    }
  } else {  // ceres::CauchyLoss(0.2): never reachable in the reference (Q1); opt-in mode only
    const double a = 0.2, b = a * a, c = 1.0 / b;
    const double sum = 1.0 + s * c;
    const double inv = 1.0 / sum;
    rho[0] = b * std::log(sum);
    rho[1] = std::max(std::numeric_limits<double>::min(), inv);
    rho[2] = -c * (inv * inv);
  }
}

// ProgramEvaluator::Evaluate for this problem: residuals, cost, local Jacobian (C x 6, row-major), gradient.
bool evaluate_program(const std::vector<Residual>& blocks, LossKind loss, const double x[7], double* cost,
                      std::vector<double>* residuals, std::vector<double>* jacobian, double gradient[6]) {
  const size_t C = blocks.size();
  double total = 0.0;
  if (residuals) residuals->resize(C);
  if (jacobian) jacobian->resize(C * 6);
  if (gradient) for (int j = 0; j < 6; ++j) gradient[j] = 0.0;
  for (size_t i = 0; i < C; ++i) {
    double r, jac7[7];
    if (!evaluate_residual(blocks[i], x, &r, jacobian ? jac7 : nullptr)) return false;
    // local Jacobian = global(1x7) * ComputeJacobian(7x6 = [I6;0]) (src/lidarOptimization.cpp:93-100)
    double jl[6];
    if (jacobian) for (int j = 0; j < 6; ++j) jl[j] = jac7[j];
    double squared_norm = r * r;
    if (loss == LOSS_TRIVIAL) {
      total += 0.5 * squared_norm;
    } else {
      double rho[3];
      loss_evaluate(loss, squared_norm, rho);
      total += 0.5 * rho[0];
      // Corrector: sq_norm == 0 or rho[2] <= 0  ->  residual and Jacobian scaled by sqrt(rho[1])
      const double sqrt_rho1 = std::sqrt(rho[1]);
      if (squared_norm == 0.0 || rho[2] <= 0.0) {
        if (jacobian) for (int j = 0; j < 6; ++j) jl[j] *= sqrt_rho1;
        r *= sqrt_rho1;
      } else {
        // general corrector (rho'' > 0): not reachable with Huber/Cauchy; kept for completeness
        const double D = 1.0 + 2.0 * squared_norm * rho[2] / rho[1];
        const double alpha = 1.0 - std::sqrt(D);
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

struct BlockProgram : LmProgram {  // the restated cost functions + loss (evaluate_program above)
  const std::vector<Residual>& blocks;
  LossKind loss;
  BlockProgram(const std::vector<Residual>& b, LossKind l) : blocks(b), loss(l) {}
  size_t num_residuals() const override { return blocks.size(); }
  bool evaluate(const double x[7], double* cost, std::vector<double>* residuals, std::vector<double>* jacobian, double gradient[6]) override {
    return evaluate_program(blocks, loss, x, cost, residuals, jacobian, gradient);
  }
  void plus(const double x[7], const double delta[6], double out[7]) override { se3_plus(x, delta, out); }
};

double norm7(const double a[7]) {
  double s = 0;
  for (int i = 0; i < 7; ++i) s += a[i] * a[i];
  return std::sqrt(s);
}

}  // namespace

void ceres_solve_pose(const std::vector<Residual>& blocks, LossKind loss, double x_io[7], LmSummary* summary, int max_num_iterations) {
  BlockProgram program(blocks, loss);
  trust_region_lm(program, x_io, summary, max_num_iterations);
}

// TrustRegionMinimizer + LevenbergMarquardtStrategy + DenseQRSolver (Ceres 1.13/1.14) for ONE parameter block of global size 7
// and local size 6.  Shared by the restated problem above and by oracle/stubs/ceres/ceres.h, where the program calls the
// reference's own cost functions and parameterization through their virtual interfaces.
void trust_region_lm(LmProgram& program, double x_io[7], LmSummary* summary, int max_num_iterations) {
  LmSummary local;
  LmSummary& S = summary ? *summary : local;
  S = LmSummary();
  std::memset(S.H0, 0, sizeof(S.H0));
  std::memset(S.g0, 0, sizeof(S.g0));
  const size_t C = program.num_residuals();
  if (C == 0) { S.termination = 5; return; }

  // Solver::Options defaults in force (Appendix A.5)
  double radius = 1e4;
  const double max_radius = 1e16, min_radius = 1e-32, min_relative_decrease = 1e-3;
  const double min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
  const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  double decrease_factor = 2.0;
  bool reuse_diagonal = false;

  double x[7];
  std::memcpy(x, x_io, sizeof(x));
  double x_norm = norm7(x);
  double x_cost, gradient[6], scale[6], diagonal[6];
  std::vector<double> residuals, jacobian;  // jacobian holds J_s = J*diag(scale) after evaluate_gradient_and_jacobian
  double gradient_max_norm = 0.0;

  int iteration = 0;
  auto evaluate_gradient_and_jacobian = [&]() -> bool {  // TrustRegionMinimizer::EvaluateGradientAndJacobian
    if (!program.evaluate(x, &x_cost, &residuals, &jacobian, gradient)) return false;
    if (iteration == 0) {
      for (int j = 0; j < 6; ++j) {
        double s = 0;
        for (size_t i = 0; i < C; ++i) s += jacobian[i * 6 + j] * jacobian[i * 6 + j];
        scale[j] = 1.0 / (1.0 + std::sqrt(s));
      }
      for (int a = 0; a < 6; ++a) {
        S.g0[a] = gradient[a];
        for (int b = 0; b < 6; ++b) {
          double s = 0;
          for (size_t i = 0; i < C; ++i) s += jacobian[i * 6 + a] * jacobian[i * 6 + b];
          S.H0[a * 6 + b] = s;
        }
      }
    }
    for (size_t i = 0; i < C; ++i)
      for (int j = 0; j < 6; ++j) jacobian[i * 6 + j] *= scale[j];
    // gradient_max_norm = |x - Plus(x, -g)|_inf
    double ng[6], proj[7];
    for (int j = 0; j < 6; ++j) ng[j] = -gradient[j];
    program.plus(x, ng, proj);
    gradient_max_norm = 0.0;
    for (int j = 0; j < 7; ++j) gradient_max_norm = std::max(gradient_max_norm, std::fabs(x[j] - proj[j]));
    return true;
  };

  // IterationZero
  if (!evaluate_gradient_and_jacobian()) { S.termination = 4; return; }
  S.initial_cost = x_cost;
  bool last_step_successful = true;

  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (last_step_successful) std::memcpy(x_io, x, sizeof(x));  // monotonic steps: every accepted x is the minimum so far
    if (iteration >= max_num_iterations) { S.termination = 0; break; }
    if (last_step_successful && gradient_max_norm <= gradient_tolerance) { S.termination = 3; break; }
    if (radius <= min_radius) { S.termination = 6; break; }
    ++iteration;
    S.iterations = iteration;
    last_step_successful = false;

    // LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal) {
      for (int j = 0; j < 6; ++j) {
        double s = 0;
        for (size_t i = 0; i < C; ++i) s += jacobian[i * 6 + j] * jacobian[i * 6 + j];
        diagonal[j] = std::min(std::max(s, min_lm_diagonal), max_lm_diagonal);
      }
    }
    double lm_diagonal[6];
    for (int j = 0; j < 6; ++j) lm_diagonal[j] = std::sqrt(diagonal[j] / radius);
    // DenseQRSolver: [J_s; diag(D)] y = [r; 0]  via Householder QR ; step = -y
    const int rows = (int)C + 6;
    std::vector<double> A((size_t)rows * 6, 0.0), rhs(rows, 0.0);
    for (size_t i = 0; i < C; ++i) {
      for (int j = 0; j < 6; ++j) A[(size_t)j * rows + i] = jacobian[i * 6 + j];
      rhs[i] = residuals[i];
    }
    for (int j = 0; j < 6; ++j) A[(size_t)j * rows + C + j] = lm_diagonal[j];
    double step[6];
    householder_qr_solve(A.data(), rhs.data(), rows, 6, step);
    bool step_finite = true;
    for (int j = 0; j < 6; ++j) { step[j] = -step[j]; step_finite = step_finite && std::isfinite(step[j]); }
    reuse_diagonal = true;

    // model_cost_change = -(J_s step)^T (r + J_s step / 2)
    double model_cost_change = 0.0;
    if (step_finite) {
      for (size_t i = 0; i < C; ++i) {
        double m = 0;
        for (int j = 0; j < 6; ++j) m += jacobian[i * 6 + j] * step[j];
        model_cost_change += m * (residuals[i] + m / 2.0);
      }
      model_cost_change = -model_cost_change;
    }
    if (getenv("FO_LM_DEBUG")) fprintf(stderr, "it %d radius %g step %g %g %g %g %g %g finite %d model %g\n", iteration, radius, step[0], step[1], step[2], step[3], step[4], step[5], (int)step_finite, model_cost_change);
    if (!step_finite || !(model_cost_change > 0.0)) {  // HandleInvalidStep
      radius *= 0.5;
      reuse_diagonal = true;
      continue;  // (max_num_consecutive_invalid_steps=5 cannot be reached within 4 iterations)
    }
    double delta[6];
    for (int j = 0; j < 6; ++j) delta[j] = step[j] * scale[j];

    // ComputeCandidatePointAndEvaluateCost
    double candidate_x[7], candidate_cost;
    program.plus(x, delta, candidate_x);
    if (!program.evaluate(candidate_x, &candidate_cost, nullptr, nullptr, nullptr))
      candidate_cost = std::numeric_limits<double>::max();

    if (getenv("FO_LM_DEBUG")) fprintf(stderr, "it %d radius %g step %g %g %g %g %g %g delta %g %g %g %g %g %g model %g cost %g cand %g\n", iteration, radius, step[0], step[1], step[2], step[3], step[4], step[5], delta[0], delta[1], delta[2], delta[3], delta[4], delta[5], model_cost_change, x_cost, candidate_cost);
    // ParameterToleranceReached
    double diff[7];
    for (int j = 0; j < 7; ++j) diff[j] = x[j] - candidate_x[j];
    const double step_norm = norm7(diff);
    if (step_norm <= parameter_tolerance * (x_norm + parameter_tolerance)) { S.termination = 1; break; }
    // FunctionToleranceReached
    const double cost_change = x_cost - candidate_cost;
    if (std::fabs(cost_change) <= function_tolerance * x_cost) { S.termination = 2; break; }
    // IsStepSuccessful
    const double relative_decrease = cost_change / model_cost_change;
    if (relative_decrease > min_relative_decrease) {  // HandleSuccessfulStep
      std::memcpy(x, candidate_x, sizeof(x));
      x_norm = norm7(x);
      if (!evaluate_gradient_and_jacobian()) { S.termination = 4; break; }
      last_step_successful = true;
      S.accepted++;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3));
      radius = std::min(max_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
    } else {  // HandleUnsuccessfulStep
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
    }
  }
  S.final_cost = x_cost;
}

}  // namespace fo
