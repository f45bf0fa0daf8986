// Small fixed-size double-precision routines used by the association and solve kernels (device + host callable).
// They follow the same published algorithms the reference gets from Eigen/Ceres; results agree
// with the CPU oracle to rounding (GPU code may contract a*b+c into FMA here — only the float distance / curvature / voxel
// arithmetic in other files has to be bit-exact).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace floam {
namespace m {

#define FM_HD __host__ __device__ __forceinline__

struct V3 { double x, y, z; };
FM_HD V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
FM_HD V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
FM_HD V3 scale(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
FM_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
FM_HD V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
FM_HD double norm(V3 a) { return sqrt(dot(a, a)); }

// Eigen QuaternionBase::_transformVector, q = (x,y,z,w) not normalised
FM_HD V3 quat_rotate(const double* q, V3 v) {
  V3 qv{q[0], q[1], q[2]};
  V3 uv = cross(qv, v);
  uv = add(uv, uv);
  return add(add(v, scale(q[3], uv)), cross(qv, uv));
}
FM_HD void quat_mul(const double* a, const double* b, double* o) {  // (x,y,z,w)
  const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  const double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  const double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  const double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
FM_HD void quat_to_matrix(const double* q, double* R) {  // toRotationMatrix, row-major
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
FM_HD void quat_from_matrix(const double* R, double* q) {  // Quaterniond(Matrix3d), row-major R
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 3 + i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
    q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
  }
}
// T = [R|t] as 12 doubles. Isometry product / inverse (Eigen Transform<double,3,Isometry>)
FM_HD void iso_mul(const double* A, const double* B, double* O) {
  double r[12];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) r[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
    r[9 + i] = A[i * 3 + 0] * B[9] + A[i * 3 + 1] * B[10] + A[i * 3 + 2] * B[11] + A[9 + i];
  }
  for (int i = 0; i < 12; ++i) O[i] = r[i];
}
FM_HD void iso_inverse(const double* A, double* O) {
  double r[12];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r[i * 3 + j] = A[j * 3 + i];
  for (int i = 0; i < 3; ++i) r[9 + i] = -(r[i * 3 + 0] * A[9] + r[i * 3 + 1] * A[10] + r[i * 3 + 2] * A[11]);
  for (int i = 0; i < 12; ++i) O[i] = r[i];
}
FM_HD double rotation_angle(const double* R) {  // AngleAxisd(R).angle()
  double q[4];
  quat_from_matrix(R, q);
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  return n != 0.0 ? 2.0 * atan2(n, fabs(q[3])) : 0.0;
}

// getTransformFromSe3 + PoseSE3Parameterization::Plus (reference src/lidarOptimization.cpp:77-137)
FM_HD void se3_plus(const double* x, const double* delta, double* out) {
  const V3 omega{delta[0], delta[1], delta[2]}, upsilon{delta[3], delta[4], delta[5]};
  const double theta = norm(omega);
  const double half_theta = 0.5 * theta;
  double imag_factor;
  const double real_factor = cos(half_theta);
  if (theta < 1e-10) {
    const double theta_sq = theta * theta, theta_po4 = theta_sq * theta_sq;
    imag_factor = 0.5 - 0.0208333 * theta_sq + 0.000260417 * theta_po4;
  } else {
    imag_factor = sin(half_theta) / theta;
  }
  const double dq[4] = {imag_factor * omega.x, imag_factor * omega.y, imag_factor * omega.z, real_factor};
  V3 dt;
  if (theta < 1e-10) {
    double R[9];
    quat_to_matrix(dq, R);
    dt = {R[0] * upsilon.x + R[1] * upsilon.y + R[2] * upsilon.z, R[3] * upsilon.x + R[4] * upsilon.y + R[5] * upsilon.z,
          R[6] * upsilon.x + R[7] * upsilon.y + R[8] * upsilon.z};
  } else {
    // J = I + c1 * Omega + c2 * Omega^2 ; J*u = u + c1 (omega x u) + c2 (omega x (omega x u))
    const double c1 = (1 - cos(theta)) / (theta * theta);
    const double c2 = (theta - sin(theta)) / (theta * theta * theta);
    const V3 wu = cross(omega, upsilon);
    const V3 wwu = cross(omega, wu);
    dt = add(upsilon, add(scale(c1, wu), scale(c2, wwu)));
  }
  quat_mul(dq, x, out);
  const V3 tp = add(quat_rotate(dq, V3{x[4], x[5], x[6]}), dt);
  out[4] = tp.x; out[5] = tp.y; out[6] = tp.z;
}

// SelfAdjointEigenSolver<Matrix3d>: tridiagonalisation + implicit symmetric QR (Eigen 3.3). A = lower triangle
// (a00,a10,a11,a20,a21,a22). Returns eigenvalues ascending in vals and the eigenvector of the largest in vmax.
FM_HD void givens(double p, double q, double& c, double& s) {
  if (q == 0.0) { c = p < 0.0 ? -1.0 : 1.0; s = 0.0; }
  else if (p == 0.0) { c = 0.0; s = q < 0.0 ? 1.0 : -1.0; }
  else if (fabs(p) > fabs(q)) { const double t = q / p; double u = sqrt(1.0 + t * t); if (p < 0.0) u = -u; c = 1.0 / u; s = -t * c; }
  else { const double t = p / q; double u = sqrt(1.0 + t * t); if (q < 0.0) u = -u; s = -1.0 / u; c = -t * s; }
}
FM_HD bool eigen3_sym(double m00, double m10, double m11, double m20, double m21, double m22, double* vals, double* vmax) {
  double scale_ = fmax(fmax(fmax(fabs(m00), fabs(m10)), fmax(fabs(m11), fabs(m20))), fmax(fabs(m21), fabs(m22)));
  if (scale_ == 0.0) scale_ = 1.0;
  m00 /= scale_; m10 /= scale_; m11 /= scale_; m20 /= scale_; m21 /= scale_; m22 /= scale_;
  double diag[3], sub[2], Q[9];
  const double tiny = 2.2250738585072014e-308;
  diag[0] = m00;
  const double v1norm2 = m20 * m20;
  if (v1norm2 <= tiny) {
    diag[1] = m11; diag[2] = m22; sub[0] = m10; sub[1] = m21;
    Q[0] = 1; Q[1] = 0; Q[2] = 0; Q[3] = 0; Q[4] = 1; Q[5] = 0; Q[6] = 0; Q[7] = 0; Q[8] = 1;
  } else {
    const double beta = sqrt(m10 * m10 + v1norm2);
    const double invBeta = 1.0 / beta;
    const double m01 = m10 * invBeta, m02 = m20 * invBeta;
    const double q = 2.0 * m01 * m21 + m02 * (m22 - m11);
    diag[1] = m11 + m02 * q;
    diag[2] = m22 - m02 * q;
    sub[0] = beta;
    sub[1] = m21 - m01 * q;
    Q[0] = 1; Q[1] = 0; Q[2] = 0; Q[3] = 0; Q[4] = m01; Q[5] = m02; Q[6] = 0; Q[7] = m02; Q[8] = -m01;
  }
  int end = 2, start = 0, iter = 0;
  const double precision = 2.0 * 2.220446049250313e-16;
  while (end > 0) {
    for (int i = start; i < end; ++i)
      if (fabs(sub[i]) <= (fabs(diag[i]) + fabs(diag[i + 1])) * precision || fabs(sub[i]) <= tiny) sub[i] = 0.0;
    while (end > 0 && sub[end - 1] == 0.0) end--;
    if (end <= 0) break;
    iter++;
    if (iter > 90) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0.0) start--;
    const double td = (diag[end - 1] - diag[end]) * 0.5;
    const double e = sub[end - 1];
    double mu = diag[end];
    if (td == 0.0) {
      mu -= fabs(e);
    } else {
      const double e2 = e * e;
      const double h = hypot(td, e);
      if (e2 == 0.0) mu -= (e / (td + (td > 0.0 ? 1.0 : -1.0))) * (e / h);
      else mu -= e2 / (td + (td > 0.0 ? h : -h));
    }
    double x = diag[start] - mu;
    double z = sub[start];
    for (int k = start; k < end; ++k) {
      double c, s;
      givens(x, z, c, s);
      const double sdk = s * diag[k] + c * sub[k];
      const double dkp1 = s * sub[k] + c * diag[k + 1];
      diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
      diag[k + 1] = s * sdk + c * dkp1;
      sub[k] = c * sdk - s * dkp1;
      if (k > start) sub[k - 1] = c * sub[k - 1] - s * z;
      x = sub[k];
      if (k < end - 1) { z = -s * sub[k + 1]; sub[k + 1] = c * sub[k + 1]; }
      for (int i = 0; i < 3; ++i) {
        const double xi = Q[i * 3 + k], yi = Q[i * 3 + k + 1];
        Q[i * 3 + k] = c * xi - s * yi;
        Q[i * 3 + k + 1] = s * xi + c * yi;
      }
    }
  }
  // ascending selection sort with matching eigenvector columns
  for (int i = 0; i < 2; ++i) {
    int k = i;
    for (int j = i + 1; j < 3; ++j) if (diag[j] < diag[k]) k = j;
    if (k != i) {
      const double t = diag[i]; diag[i] = diag[k]; diag[k] = t;
      for (int r = 0; r < 3; ++r) { const double u = Q[r * 3 + i]; Q[r * 3 + i] = Q[r * 3 + k]; Q[r * 3 + k] = u; }
    }
  }
  vals[0] = diag[0] * scale_; vals[1] = diag[1] * scale_; vals[2] = diag[2] * scale_;
  vmax[0] = Q[2]; vmax[1] = Q[5]; vmax[2] = Q[8];
  return iter <= 90;
}

// Matrix<double,5,3>::colPivHouseholderQr().solve(b): A column-major a[c*5+r], overwritten. (Eigen 3.3 ColPivHouseholderQR)
FM_HD void householder_make(double* v, int n, double& tau, double& beta) {
  double tail = 0.0;
  for (int i = 1; i < n; ++i) tail += v[i] * v[i];
  const double c0 = v[0];
  if (n == 1 || tail <= 2.2250738585072014e-308) {
    tau = 0.0; beta = c0;
    for (int i = 1; i < n; ++i) v[i] = 0.0;
  } else {
    beta = sqrt(c0 * c0 + tail);
    if (c0 >= 0.0) beta = -beta;
    for (int i = 1; i < n; ++i) v[i] = v[i] / (c0 - beta);
    tau = (beta - c0) / beta;
  }
}
FM_HD void householder_apply(double* x, int n, const double* ess, double tau) {
  if (n == 1) { x[0] *= (1.0 - tau); return; }
  if (tau == 0.0) return;
  double tmp = 0.0;
  for (int i = 1; i < n; ++i) tmp += ess[i - 1] * x[i];
  tmp += x[0];
  x[0] -= tau * tmp;
  for (int i = 1; i < n; ++i) x[i] -= tau * ess[i - 1] * tmp;
}
FM_HD void colpiv_qr_solve_5x3(double* a, const double* b, double* x_out) {
  const int rows = 5, cols = 3;
  double h[3], nu[3], nd[3];
  int transp[3];
  const double eps = 2.220446049250313e-16;
  for (int k = 0; k < cols; ++k) {
    double s = 0;
    for (int r = 0; r < rows; ++r) s += a[k * 5 + r] * a[k * 5 + r];
    nd[k] = nu[k] = sqrt(s);
  }
  const double maxn = fmax(nu[0], fmax(nu[1], nu[2]));
  const double threshold_helper = (maxn * eps) * (maxn * eps) / (double)rows;
  const double downdate = sqrt(eps);
  int nonzero = cols;
  for (int k = 0; k < cols; ++k) {
    int big = k;
    double bn = nu[k];
    for (int j = k + 1; j < cols; ++j) if (nu[j] > bn) { bn = nu[j]; big = j; }
    if (nonzero == cols && bn * bn < threshold_helper * (double)(rows - k)) nonzero = k;
    transp[k] = big;
    if (k != big) {
      for (int r = 0; r < rows; ++r) { const double t = a[k * 5 + r]; a[k * 5 + r] = a[big * 5 + r]; a[big * 5 + r] = t; }
      double t = nu[k]; nu[k] = nu[big]; nu[big] = t;
      t = nd[k]; nd[k] = nd[big]; nd[big] = t;
    }
    double beta;
    householder_make(&a[k * 5 + k], rows - k, h[k], beta);
    a[k * 5 + k] = beta;
    for (int j = k + 1; j < cols; ++j) householder_apply(&a[j * 5 + k], rows - k, &a[k * 5 + k + 1], h[k]);
    for (int j = k + 1; j < cols; ++j) {
      if (nu[j] != 0.0) {
        double temp = fabs(a[j * 5 + k]) / nu[j];
        temp = (1.0 + temp) * (1.0 - temp);
        temp = temp < 0.0 ? 0.0 : temp;
        const double ratio = nu[j] / nd[j];
        const double temp2 = temp * ratio * ratio;
        if (temp2 <= downdate) {
          double s = 0;
          for (int r = k + 1; r < rows; ++r) s += a[j * 5 + r] * a[j * 5 + r];
          nd[j] = sqrt(s);
          nu[j] = nd[j];
        } else {
          nu[j] *= sqrt(temp);
        }
      }
    }
  }
  int perm[3] = {0, 1, 2};
  for (int k = 0; k < cols; ++k) { const int t = perm[k]; perm[k] = perm[transp[k]]; perm[transp[k]] = t; }
  x_out[0] = x_out[1] = x_out[2] = 0.0;
  if (nonzero == 0) return;
  double c[5];
  for (int r = 0; r < rows; ++r) c[r] = b[r];
  for (int k = 0; k < nonzero; ++k) householder_apply(&c[k], rows - k, &a[k * 5 + k + 1], h[k]);
  for (int i = nonzero - 1; i >= 0; --i) {
    double s = c[i];
    for (int j = i + 1; j < nonzero; ++j) s -= a[j * 5 + i] * c[j];
    c[i] = s / a[i * 5 + i];
  }
  for (int i = 0; i < nonzero; ++i) x_out[perm[i]] = c[i];
}

// Solve the symmetric positive definite 6x6 system A y = b (A full row-major 36). Cholesky; returns false when not SPD / non-finite.
FM_HD bool cholesky6_solve(const double* A, const double* b, double* y) {
  // L L^T = A with one reciprocal per column and fused multiply-adds (this solve replaces Ceres' Householder QR of the stacked
  // Jacobian; it is not part of the bit-exact arithmetic, only of the 1e-4 pose tolerance)
  double L[36], inv[6];
  for (int j = 0; j < 6; ++j) {
    double s = A[j * 6 + j];
    for (int k = 0; k < j; ++k) s = fma(-L[j * 6 + k], L[j * 6 + k], s);
    if (!(s > 0.0)) return false;
    const double d = sqrt(s);
    L[j * 6 + j] = d;
    inv[j] = 1.0 / d;
    for (int i = j + 1; i < 6; ++i) {
      double t = A[i * 6 + j];
      for (int k = 0; k < j; ++k) t = fma(-L[i * 6 + k], L[j * 6 + k], t);
      L[i * 6 + j] = t * inv[j];
    }
  }
  double z[6];
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s = fma(-L[i * 6 + k], z[k], s);
    z[i] = s * inv[i];
  }
  for (int i = 5; i >= 0; --i) {
    double s = z[i];
    for (int k = i + 1; k < 6; ++k) s = fma(-L[k * 6 + i], y[k], s);
    y[i] = s * inv[i];
  }
  for (int i = 0; i < 6; ++i) if (!isfinite(y[i])) return false;
  return true;
}

#undef FM_HD
}  // namespace m
}  // namespace floam
