// ORACLE — TEST INFRASTRUCTURE ONLY (see types.h).  Restates Eigen 3.3 routines (un-vendored, absent here): parity unpinned for these library internals.
//
// Restatement of the Eigen 3.3.x routines the reference calls on the odometry path.
// Eigen is not vendored in /root/reference and is absent from this image; the algorithms below follow the
// published Eigen 3.3.4 sources (Geometry/Quaternion.h, Geometry/AngleAxis.h, Geometry/Transform.h,
// Eigenvalues/SelfAdjointEigenSolver.h, Eigenvalues/Tridiagonalization.h, Jacobi/Jacobi.h,
// Householder/Householder.h, QR/ColPivHouseholderQR.h, QR/HouseholderQR.h).  Call sites in the reference:
//   src/odomEstimationClass.cpp:62,70-71,115 (Isometry3d algebra, Quaterniond(Matrix3d), toRotationMatrix)
//   src/odomEstimationClass.cpp:175-179      (SelfAdjointEigenSolver<Matrix3d>)
//   src/odomEstimationClass.cpp:220-222      (Matrix<double,5,3>::colPivHouseholderQr().solve)
//   src/odomEstimationClass.cpp:329-331      (Isometry inverse, AngleAxisd(Matrix3d).angle())
#pragma once
#include "types.h"
#include <algorithm>
#include <cfloat>
#include <limits>
#include <utility>
#include <vector>

namespace fo {

// ---------------- Mat3 / Iso3 ----------------
inline Mat3 mat3_identity() { return {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}; }
inline Mat3 mat3_mul(const Mat3& a, const Mat3& b) {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}
inline Mat3 mat3_transpose(const Mat3& a) {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
  return r;
}
inline Vec3 mat3_apply(const Mat3& a, Vec3 v) {
  return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
          a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
inline Iso3 iso_identity() { return {mat3_identity(), {0, 0, 0}}; }
// Transform<double,3,Isometry>::operator* : R = R1 R2, t = R1 t2 + t1
inline Iso3 iso_mul(const Iso3& a, const Iso3& b) { return {mat3_mul(a.R, b.R), mat3_apply(a.R, b.t) + a.t}; }
// Transform::inverse(Isometry): R^T, -R^T t
inline Iso3 iso_inverse(const Iso3& a) {
  Mat3 rt = mat3_transpose(a.R);
  Vec3 t = mat3_apply(rt, a.t);
  return {rt, {-t.x, -t.y, -t.z}};
}

// ---------------- Quaternion (Eigen::Quaterniond semantics; coefficient order x,y,z,w) ----------------
inline Quat quat_mul(const Quat& a, const Quat& b) {  // internal::quat_product
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
inline Quat quat_inverse(const Quat& q) {  // QuaternionBase::inverse
  double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  if (n2 > 0) return {-q.x / n2, -q.y / n2, -q.z / n2, q.w / n2};
  return {0, 0, 0, 0};
}
inline Vec3 quat_rotate(const Quat& q, Vec3 v) {  // QuaternionBase::_transformVector
  Vec3 qv{q.x, q.y, q.z};
  Vec3 uv = cross(qv, v);
  uv = uv + uv;
  return v + q.w * uv + cross(qv, uv);
}
inline Mat3 quat_to_matrix(const Quat& q) {  // QuaternionBase::toRotationMatrix (no normalisation)
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  Mat3 r;
  r.m[0][0] = 1 - (tyy + tzz); r.m[0][1] = txy - twz;       r.m[0][2] = txz + twy;
  r.m[1][0] = txy + twz;       r.m[1][1] = 1 - (txx + tzz); r.m[1][2] = tyz - twx;
  r.m[2][0] = txz - twy;       r.m[2][1] = tyz + twx;       r.m[2][2] = 1 - (txx + tyy);
  return r;
}
inline Quat quat_from_matrix(const Mat3& mat) {  // internal::quaternionbase_assign_impl<Other,3,3>
  double c[4];                                   // x,y,z,w
  double t = mat.m[0][0] + mat.m[1][1] + mat.m[2][2];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    c[3] = 0.5 * t;
    t = 0.5 / t;
    c[0] = (mat.m[2][1] - mat.m[1][2]) * t;
    c[1] = (mat.m[0][2] - mat.m[2][0]) * t;
    c[2] = (mat.m[1][0] - mat.m[0][1]) * t;
  } else {
    int i = 0;
    if (mat.m[1][1] > mat.m[0][0]) i = 1;
    if (mat.m[2][2] > mat.m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(mat.m[i][i] - mat.m[j][j] - mat.m[k][k] + 1.0);
    c[i] = 0.5 * t;
    t = 0.5 / t;
    c[3] = (mat.m[k][j] - mat.m[j][k]) * t;
    c[j] = (mat.m[j][i] + mat.m[i][j]) * t;
    c[k] = (mat.m[k][i] + mat.m[i][k]) * t;
  }
  return {c[0], c[1], c[2], c[3]};
}
// AngleAxisd(Matrix3d).angle()  (AngleAxis::operator=(QuaternionBase))
inline double rotation_angle(const Mat3& R) {
  Quat q = quat_from_matrix(R);
  double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  if (n != 0.0) return 2.0 * std::atan2(n, std::fabs(q.w));
  return 0.0;
}
// Eigen::AngleAxisd(angle, axis) -> quaternion, and the product order of src/lidar.cpp:8-16
inline Quat quat_from_axis_angle(double angle, Vec3 axis) {
  double s = std::sin(0.5 * angle), c = std::cos(0.5 * angle);
  return {s * axis.x, s * axis.y, s * axis.z, c};
}
inline Quat euler2Quaternion(double roll, double pitch, double yaw) {  // src/lidar.cpp:8-16 (roll*yaw*pitch, degrees)
  Quat r = quat_from_axis_angle(roll * M_PI / 180.0, {1, 0, 0});
  Quat p = quat_from_axis_angle(pitch * M_PI / 180.0, {0, 1, 0});
  Quat y = quat_from_axis_angle(yaw * M_PI / 180.0, {0, 0, 1});
  return quat_mul(quat_mul(r, y), p);
}

// ---------------- SelfAdjointEigenSolver<Matrix3d> (general iterative path, not computeDirect) ----------------
struct Eigen3 {
  double values[3];      // ascending
  double vectors[3][3];  // vectors[row][col], column i <-> values[i]
  bool ok;
};

namespace detail {
struct Givens {
  double c, s;
};
inline Givens make_givens(double p, double q) {  // JacobiRotation::makeGivens (real)
  Givens g;
  if (q == 0.0) {
    g.c = p < 0.0 ? -1.0 : 1.0;
    g.s = 0.0;
  } else if (p == 0.0) {
    g.c = 0.0;
    g.s = q < 0.0 ? 1.0 : -1.0;
  } else if (std::fabs(p) > std::fabs(q)) {
    double t = q / p;
    double u = std::sqrt(1.0 + t * t);
    if (p < 0.0) u = -u;
    g.c = 1.0 / u;
    g.s = -t * g.c;
  } else {
    double t = p / q;
    double u = std::sqrt(1.0 + t * t);
    if (q < 0.0) u = -u;
    g.s = -1.0 / u;
    g.c = -t * g.s;
  }
  return g;
}
}  // namespace detail

inline Eigen3 self_adjoint_eigen3(const Mat3& A) {
  Eigen3 out;
  double diag[3], sub[2];
  double Q[3][3];
  // lower triangle, scaled to [-1,1]
  double m00 = A.m[0][0], m10 = A.m[1][0], m11 = A.m[1][1], m20 = A.m[2][0], m21 = A.m[2][1], m22 = A.m[2][2];
  double scale = std::max({std::fabs(m00), std::fabs(m10), std::fabs(m11), std::fabs(m20), std::fabs(m21), std::fabs(m22)});
  if (scale == 0.0) scale = 1.0;
  m00 /= scale; m10 /= scale; m11 /= scale; m20 /= scale; m21 /= scale; m22 /= scale;
  // tridiagonalization_inplace_selector<MatrixType,3,false>
  const double tol = std::numeric_limits<double>::min();
  diag[0] = m00;
  double v1norm2 = m20 * m20;
  if (v1norm2 <= tol) {
    diag[1] = m11; diag[2] = m22; sub[0] = m10; sub[1] = m21;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Q[i][j] = (i == j) ? 1.0 : 0.0;
  } else {
    double beta = std::sqrt(m10 * m10 + v1norm2);
    double invBeta = 1.0 / beta;
    double m01 = m10 * invBeta;
    double m02 = m20 * invBeta;
    double q = 2.0 * m01 * m21 + m02 * (m22 - m11);
    diag[1] = m11 + m02 * q;
    diag[2] = m22 - m02 * q;
    sub[0] = beta;
    sub[1] = m21 - m01 * q;
    double Qi[3][3] = {{1, 0, 0}, {0, m01, m02}, {0, m02, -m01}};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Q[i][j] = Qi[i][j];
  }
  // computeFromTridiagonal_impl
  const int n = 3, maxIterations = 30;
  int end = n - 1, start = 0, iter = 0;
  const double considerAsZero = std::numeric_limits<double>::min();
  const double precision = 2.0 * std::numeric_limits<double>::epsilon();
  while (end > 0) {
    for (int i = start; i < end; ++i)
      if (std::fabs(sub[i]) <= (std::fabs(diag[i]) + std::fabs(diag[i + 1])) * precision || std::fabs(sub[i]) <= considerAsZero)
        sub[i] = 0.0;
    while (end > 0 && sub[end - 1] == 0.0) end--;
    if (end <= 0) break;
    iter++;
    if (iter > maxIterations * n) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0.0) start--;
    // tridiagonal_qr_step
    double td = (diag[end - 1] - diag[end]) * 0.5;
    double e = sub[end - 1];
    double mu = diag[end];
    if (td == 0.0) {
      mu -= std::fabs(e);
    } else {
      double e2 = e * e;
      double h = std::hypot(td, e);
      if (e2 == 0.0) mu -= (e / (td + (td > 0.0 ? 1.0 : -1.0))) * (e / h);
      else mu -= e2 / (td + (td > 0.0 ? h : -h));
    }
    double x = diag[start] - mu;
    double z = sub[start];
    for (int k = start; k < end; ++k) {
      detail::Givens rot = detail::make_givens(x, z);
      double sdk = rot.s * diag[k] + rot.c * sub[k];
      double dkp1 = rot.s * sub[k] + rot.c * diag[k + 1];
      diag[k] = rot.c * (rot.c * diag[k] - rot.s * sub[k]) - rot.s * (rot.c * sub[k] - rot.s * diag[k + 1]);
      diag[k + 1] = rot.s * sdk + rot.c * dkp1;
      sub[k] = rot.c * sdk - rot.s * dkp1;
      if (k > start) sub[k - 1] = rot.c * sub[k - 1] - rot.s * z;
      x = sub[k];
      if (k < end - 1) {
        z = -rot.s * sub[k + 1];
        sub[k + 1] = rot.c * sub[k + 1];
      }
      // Q = Q * G  (applyOnTheRight(k,k+1,rot))
      for (int i = 0; i < 3; ++i) {
        double xi = Q[i][k], yi = Q[i][k + 1];
        Q[i][k] = rot.c * xi - rot.s * yi;
        Q[i][k + 1] = rot.s * xi + rot.c * yi;
      }
    }
  }
  out.ok = iter <= maxIterations * n;
  // sort ascending (selection by minCoeff)
  if (out.ok) {
    for (int i = 0; i < n - 1; ++i) {
      int k = 0;
      double mn = diag[i];
      for (int j = 1; j < n - i; ++j)
        if (diag[i + j] < mn) { mn = diag[i + j]; k = j; }
      if (k > 0) {
        std::swap(diag[i], diag[k + i]);
        for (int r = 0; r < 3; ++r) std::swap(Q[r][i], Q[r][k + i]);
      }
    }
  }
  for (int i = 0; i < 3; ++i) {
    out.values[i] = diag[i] * scale;
    for (int j = 0; j < 3; ++j) out.vectors[i][j] = Q[i][j];
  }
  return out;
}

// ---------------- Householder primitives (column-major storage: a[col*lda + row]) ----------------
namespace detail {
// MatrixBase::makeHouseholder on v[0..n): returns tau, beta; essential stored in v[1..n)
inline void make_householder_inplace(double* v, int n, double& tau, double& beta) {
  double tailSqNorm = 0.0;
  for (int i = 1; i < n; ++i) tailSqNorm += v[i] * v[i];
  double c0 = v[0];
  const double tol = std::numeric_limits<double>::min();
  if (n == 1 || tailSqNorm <= tol) {
    tau = 0.0;
    beta = c0;
    for (int i = 1; i < n; ++i) v[i] = 0.0;
  } else {
    beta = std::sqrt(c0 * c0 + tailSqNorm);
    if (c0 >= 0.0) beta = -beta;
    for (int i = 1; i < n; ++i) v[i] = v[i] / (c0 - beta);
    tau = (beta - c0) / beta;
  }
}
// applyHouseholderOnTheLeft to the column x[0..n) with essential ess[0..n-1)
inline void apply_householder_left(double* x, int n, const double* ess, double tau) {
  if (n == 1) {
    x[0] *= (1.0 - tau);
  } else if (tau != 0.0) {
    double tmp = 0.0;
    for (int i = 1; i < n; ++i) tmp += ess[i - 1] * x[i];
    tmp += x[0];
    x[0] -= tau * tmp;
    for (int i = 1; i < n; ++i) x[i] -= tau * ess[i - 1] * tmp;
  }
}
}  // namespace detail

// Matrix<double,5,3>::colPivHouseholderQr().solve(b)   (rows x 3, rows>=3)
inline void colpiv_qr_solve_nx3(const double* A_rowmajor, const double* b, int rows, double x_out[3]) {
  const int cols = 3, size = 3;
  std::vector<double> qr((size_t)rows * cols);  // column-major
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) qr[(size_t)c * rows + r] = A_rowmajor[r * cols + c];
  double hCoeffs[3], normsUpdated[3], normsDirect[3];
  int transp[3];
  for (int k = 0; k < cols; ++k) {
    double s = 0;
    for (int r = 0; r < rows; ++r) s += qr[(size_t)k * rows + r] * qr[(size_t)k * rows + r];
    normsDirect[k] = normsUpdated[k] = std::sqrt(s);
  }
  const double eps = std::numeric_limits<double>::epsilon();
  double maxn = std::max({normsUpdated[0], normsUpdated[1], normsUpdated[2]});
  const double threshold_helper = (maxn * eps) * (maxn * eps) / double(rows);
  const double norm_downdate_threshold = std::sqrt(eps);
  int nonzero_pivots = size;
  for (int k = 0; k < size; ++k) {
    int biggest = k;
    double bn = normsUpdated[k];
    for (int j = k + 1; j < cols; ++j)
      if (normsUpdated[j] > bn) { bn = normsUpdated[j]; biggest = j; }
    double biggest_sq = bn * bn;
    if (nonzero_pivots == size && biggest_sq < threshold_helper * double(rows - k)) nonzero_pivots = k;
    transp[k] = biggest;
    if (k != biggest) {
      for (int r = 0; r < rows; ++r) std::swap(qr[(size_t)k * rows + r], qr[(size_t)biggest * rows + r]);
      std::swap(normsUpdated[k], normsUpdated[biggest]);
      std::swap(normsDirect[k], normsDirect[biggest]);
    }
    double beta;
    detail::make_householder_inplace(&qr[(size_t)k * rows + k], rows - k, hCoeffs[k], beta);
    qr[(size_t)k * rows + k] = beta;
    for (int j = k + 1; j < cols; ++j)
      detail::apply_householder_left(&qr[(size_t)j * rows + k], rows - k, &qr[(size_t)k * rows + k + 1], hCoeffs[k]);
    for (int j = k + 1; j < cols; ++j) {
      if (normsUpdated[j] != 0.0) {
        double temp = std::fabs(qr[(size_t)j * rows + k]) / normsUpdated[j];
        temp = (1.0 + temp) * (1.0 - temp);
        temp = temp < 0.0 ? 0.0 : temp;
        double ratio = normsUpdated[j] / normsDirect[j];
        double temp2 = temp * ratio * ratio;
        if (temp2 <= norm_downdate_threshold) {
          double s = 0;
          for (int r = k + 1; r < rows; ++r) s += qr[(size_t)j * rows + r] * qr[(size_t)j * rows + r];
          normsDirect[j] = std::sqrt(s);
          normsUpdated[j] = normsDirect[j];
        } else {
          normsUpdated[j] *= std::sqrt(temp);
        }
      }
    }
  }
  // permutation indices: P = T0 T1 T2 applied to identity (PermutationMatrix from transpositions)
  int perm[3] = {0, 1, 2};
  for (int k = 0; k < size; ++k) std::swap(perm[k], perm[transp[k]]);
  // solve
  x_out[0] = x_out[1] = x_out[2] = 0.0;
  if (nonzero_pivots == 0) return;
  std::vector<double> c(b, b + rows);
  for (int k = 0; k < nonzero_pivots; ++k)
    detail::apply_householder_left(&c[k], rows - k, &qr[(size_t)k * rows + k + 1], hCoeffs[k]);
  for (int i = nonzero_pivots - 1; i >= 0; --i) {  // upper-triangular back substitution
    double s = c[i];
    for (int j = i + 1; j < nonzero_pivots; ++j) s -= qr[(size_t)j * rows + i] * c[j];
    c[i] = s / qr[(size_t)i * rows + i];
  }
  for (int i = 0; i < nonzero_pivots; ++i) x_out[perm[i]] = c[i];
}

// A.householderQr().solve(b) for a tall rows x cols matrix (column-major, overwritten); used by the DENSE_QR restatement.
inline void householder_qr_solve(double* A_colmajor, double* b, int rows, int cols, double* x_out) {
  std::vector<double> h(cols);
  for (int k = 0; k < cols; ++k) {
    double beta;
    detail::make_householder_inplace(&A_colmajor[(size_t)k * rows + k], rows - k, h[k], beta);
    A_colmajor[(size_t)k * rows + k] = beta;
    for (int j = k + 1; j < cols; ++j)
      detail::apply_householder_left(&A_colmajor[(size_t)j * rows + k], rows - k, &A_colmajor[(size_t)k * rows + k + 1], h[k]);
  }
  for (int k = 0; k < cols; ++k) detail::apply_householder_left(&b[k], rows - k, &A_colmajor[(size_t)k * rows + k + 1], h[k]);
  for (int i = cols - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < cols; ++j) s -= A_colmajor[(size_t)j * rows + i] * x_out[j];
    x_out[i] = s / A_colmajor[(size_t)i * rows + i];
  }
}

}  // namespace fo
