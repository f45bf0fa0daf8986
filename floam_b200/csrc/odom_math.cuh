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
// Eigen 3.3 QuaternionBase::slerp(t, other), coefficients (x,y,z,w): the interpolation dmapping::ImuHandler computes tSlerp for and then
// drops (src/dataHandler.cpp:48-50,61-62); used by the opt-in FLOAM_FIX_IMU_SLERP mode only
FM_HD void quat_slerp(double t, const double* a, const double* b, double* o) {
  const double one = 1.0 - 2.220446049250313e-16;
  const double d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
  const double absD = fabs(d);
  double scale0, scale1;
  if (absD >= one) {
    scale0 = 1.0 - t; scale1 = t;
  } else {
    const double theta = acos(absD), sinTheta = sin(theta);
    scale0 = sin((1.0 - t) * theta) / sinTheta;
    scale1 = sin(t * theta) / sinTheta;
  }
  if (d < 0.0) scale1 = -scale1;
  for (int k = 0; k < 4; ++k) o[k] = scale0 * a[k] + scale1 * b[k];
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
  // one sincos of the half angle serves all four trigonometric values (double-angle identities); one reciprocal of theta serves
  // the three divisions. Same functions as the reference's getTransformFromSe3, evaluated with fewer long-latency operations.
  double sh, ch;
  sincos(half_theta, &sh, &ch);
  double imag_factor;
  const double real_factor = ch;
  V3 dt;
  if (theta < 1e-10) {
    const double theta_sq = theta * theta, theta_po4 = theta_sq * theta_sq;
    imag_factor = 0.5 - 0.0208333 * theta_sq + 0.000260417 * theta_po4;
    const double dq0[4] = {imag_factor * omega.x, imag_factor * omega.y, imag_factor * omega.z, real_factor};
    double R[9];
    quat_to_matrix(dq0, R);
    dt = {R[0] * upsilon.x + R[1] * upsilon.y + R[2] * upsilon.z, R[3] * upsilon.x + R[4] * upsilon.y + R[5] * upsilon.z,
          R[6] * upsilon.x + R[7] * upsilon.y + R[8] * upsilon.z};
  } else {
    const double inv_theta = 1.0 / theta;
    imag_factor = sh * inv_theta;
    const double sin_theta = 2.0 * sh * ch, one_minus_cos = 2.0 * sh * sh;
    // J = I + c1 * Omega + c2 * Omega^2 ; J*u = u + c1 (omega x u) + c2 (omega x (omega x u))
    const double c1 = one_minus_cos * inv_theta * inv_theta;
    const double c2 = (theta - sin_theta) * inv_theta * inv_theta * inv_theta;
    const V3 wu = cross(omega, upsilon);
    const V3 wwu = cross(omega, wu);
    dt = add(upsilon, add(scale(c1, wu), scale(c2, wwu)));
  }
  const double dq[4] = {imag_factor * omega.x, imag_factor * omega.y, imag_factor * omega.z, real_factor};
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
// Written with scalars and fully unrolled, statically indexed steps so that everything stays in registers on the device (the 3x3
// QL loop only ever rotates at k = 0 and/or k = 1); the arithmetic and its order are those of the general algorithm.
FM_HD bool eigen3_sym(double m00, double m10, double m11, double m20, double m21, double m22, double* vals, double* vmax) {
  double scale_ = fmax(fmax(fmax(fabs(m00), fabs(m10)), fmax(fabs(m11), fabs(m20))), fmax(fabs(m21), fabs(m22)));
  if (scale_ == 0.0) scale_ = 1.0;
  m00 /= scale_; m10 /= scale_; m11 /= scale_; m20 /= scale_; m21 /= scale_; m22 /= scale_;
  double d0, d1, d2, s0, s1;
  double q00 = 1, q01 = 0, q02 = 0, q10 = 0, q11 = 1, q12 = 0, q20 = 0, q21 = 0, q22 = 1;
  const double tiny = 2.2250738585072014e-308;
  d0 = m00;
  const double v1norm2 = m20 * m20;
  if (v1norm2 <= tiny) {
    d1 = m11; d2 = m22; s0 = m10; s1 = m21;
  } else {
    const double beta = sqrt(m10 * m10 + v1norm2);
    const double invBeta = 1.0 / beta;
    const double m01 = m10 * invBeta, m02 = m20 * invBeta;
    const double q = 2.0 * m01 * m21 + m02 * (m22 - m11);
    d1 = m11 + m02 * q;
    d2 = m22 - m02 * q;
    s0 = beta;
    s1 = m21 - m01 * q;
    q11 = m01; q12 = m02; q21 = m02; q22 = -m01;
  }
  int end = 2, start = 0, iter = 0;
  const double precision = 2.0 * 2.220446049250313e-16;
  while (end > 0) {
    if (start <= 0 && 0 < end && (fabs(s0) <= (fabs(d0) + fabs(d1)) * precision || fabs(s0) <= tiny)) s0 = 0.0;
    if (start <= 1 && 1 < end && (fabs(s1) <= (fabs(d1) + fabs(d2)) * precision || fabs(s1) <= tiny)) s1 = 0.0;
    if (end == 2 && s1 == 0.0) end = 1;
    if (end == 1 && s0 == 0.0) end = 0;
    if (end <= 0) break;
    iter++;
    if (iter > 90) break;
    start = end - 1;
    if (start == 1 && s0 != 0.0) start = 0;
    const double td = (end == 2 ? (d1 - d2) : (d0 - d1)) * 0.5;
    const double e = end == 2 ? s1 : s0;
    double mu = end == 2 ? d2 : d1;
    if (td == 0.0) {
      mu -= fabs(e);
    } else {
      const double e2 = e * e;
      const double h = hypot(td, e);
      if (e2 == 0.0) mu -= (e / (td + (td > 0.0 ? 1.0 : -1.0))) * (e / h);
      else mu -= e2 / (td + (td > 0.0 ? h : -h));
    }
    double x = (start == 0 ? d0 : d1) - mu;
    double z = start == 0 ? s0 : s1;
    if (start == 0) {  // k = 0
      double c, s;
      givens(x, z, c, s);
      const double sdk = s * d0 + c * s0;
      const double dkp1 = s * s0 + c * d1;
      d0 = c * (c * d0 - s * s0) - s * (c * s0 - s * d1);
      d1 = s * sdk + c * dkp1;
      s0 = c * sdk - s * dkp1;
      x = s0;
      if (end == 2) { z = -s * s1; s1 = c * s1; }
      double xi, yi;
      xi = q00; yi = q01; q00 = c * xi - s * yi; q01 = s * xi + c * yi;
      xi = q10; yi = q11; q10 = c * xi - s * yi; q11 = s * xi + c * yi;
      xi = q20; yi = q21; q20 = c * xi - s * yi; q21 = s * xi + c * yi;
    }
    if (end == 2) {    // k = 1
      double c, s;
      givens(x, z, c, s);
      const double sdk = s * d1 + c * s1;
      const double dkp1 = s * s1 + c * d2;
      d1 = c * (c * d1 - s * s1) - s * (c * s1 - s * d2);
      d2 = s * sdk + c * dkp1;
      s1 = c * sdk - s * dkp1;
      if (start == 0) s0 = c * s0 - s * z;
      double xi, yi;
      xi = q01; yi = q02; q01 = c * xi - s * yi; q02 = s * xi + c * yi;
      xi = q11; yi = q12; q11 = c * xi - s * yi; q12 = s * xi + c * yi;
      xi = q21; yi = q22; q21 = c * xi - s * yi; q22 = s * xi + c * yi;
    }
  }
  // ascending selection sort with matching eigenvector columns
#define FM_SWAPD(a, b) { const double t_ = a; a = b; b = t_; }
  {
    int k = 0;
    if (d1 < d0) k = 1;
    if (d2 < (k == 1 ? d1 : d0)) k = 2;
    if (k == 1) { FM_SWAPD(d0, d1); FM_SWAPD(q00, q01); FM_SWAPD(q10, q11); FM_SWAPD(q20, q21); }
    else if (k == 2) { FM_SWAPD(d0, d2); FM_SWAPD(q00, q02); FM_SWAPD(q10, q12); FM_SWAPD(q20, q22); }
    if (d2 < d1) { FM_SWAPD(d1, d2); FM_SWAPD(q01, q02); FM_SWAPD(q11, q12); FM_SWAPD(q21, q22); }
  }
#undef FM_SWAPD
  vals[0] = d0 * scale_; vals[1] = d1 * scale_; vals[2] = d2 * scale_;
  vmax[0] = q02; vmax[1] = q12; vmax[2] = q22;
  return iter <= 90;
}

// Matrix<double,5,3>::colPivHouseholderQr().solve(b): A column-major a[c*5+r], overwritten. (Eigen 3.3 ColPivHouseholderQR)
// Template sizes + full unrolling keep every array index static (registers on the device); pivot swaps are predicated.
template <int N>
FM_HD void householder_make(double* v, double& tau, double& beta) {
  double tail = 0.0;
#pragma unroll
  for (int i = 1; i < N; ++i) tail += v[i] * v[i];
  const double c0 = v[0];
  if (N == 1 || tail <= 2.2250738585072014e-308) {
    tau = 0.0; beta = c0;
#pragma unroll
    for (int i = 1; i < N; ++i) v[i] = 0.0;
  } else {
    beta = sqrt(c0 * c0 + tail);
    if (c0 >= 0.0) beta = -beta;
#pragma unroll
    for (int i = 1; i < N; ++i) v[i] = v[i] / (c0 - beta);
    tau = (beta - c0) / beta;
  }
}
template <int N>
FM_HD void householder_apply(double* x, const double* ess, double tau) {
  if (N == 1) { x[0] *= (1.0 - tau); return; }
  if (tau == 0.0) return;
  double tmp = 0.0;
#pragma unroll
  for (int i = 1; i < N; ++i) tmp += ess[i - 1] * x[i];
  tmp += x[0];
  x[0] -= tau * tmp;
#pragma unroll
  for (int i = 1; i < N; ++i) x[i] -= tau * ess[i - 1] * tmp;
}
template <int K>
FM_HD void colpiv_step(double* a, double* h, double* nu, double* nd, int* transp, int& nonzero, double threshold_helper, double downdate) {
  constexpr int rows = 5, cols = 3;
  int big = K;
  double bn = nu[K];
#pragma unroll
  for (int j = K + 1; j < cols; ++j) if (nu[j] > bn) { bn = nu[j]; big = j; }
  if (nonzero == cols && bn * bn < threshold_helper * (double)(rows - K)) nonzero = K;
  transp[K] = big;
#pragma unroll
  for (int j = K + 1; j < cols; ++j) {
    if (big == j) {
#pragma unroll
      for (int r = 0; r < rows; ++r) { const double t = a[K * 5 + r]; a[K * 5 + r] = a[j * 5 + r]; a[j * 5 + r] = t; }
      double t = nu[K]; nu[K] = nu[j]; nu[j] = t;
      t = nd[K]; nd[K] = nd[j]; nd[j] = t;
    }
  }
  double beta;
  householder_make<rows - K>(&a[K * 5 + K], h[K], beta);
  a[K * 5 + K] = beta;
#pragma unroll
  for (int j = K + 1; j < cols; ++j) householder_apply<rows - K>(&a[j * 5 + K], &a[K * 5 + K + 1], h[K]);
#pragma unroll
  for (int j = K + 1; j < cols; ++j) {
    if (nu[j] != 0.0) {
      double temp = fabs(a[j * 5 + K]) / nu[j];
      temp = (1.0 + temp) * (1.0 - temp);
      temp = temp < 0.0 ? 0.0 : temp;
      const double ratio = nu[j] / nd[j];
      const double temp2 = temp * ratio * ratio;
      if (temp2 <= downdate) {
        double s = 0;
#pragma unroll
        for (int r = K + 1; r < rows; ++r) s += a[j * 5 + r] * a[j * 5 + r];
        nd[j] = sqrt(s);
        nu[j] = nd[j];
      } else {
        nu[j] *= sqrt(temp);
      }
    }
  }
}
FM_HD void colpiv_qr_solve_5x3(double* a, const double* b, double* x_out) {
  constexpr int rows = 5, cols = 3;
  double h[3], nu[3], nd[3];
  int transp[3];
  const double eps = 2.220446049250313e-16;
#pragma unroll
  for (int k = 0; k < cols; ++k) {
    double s = 0;
#pragma unroll
    for (int r = 0; r < rows; ++r) s += a[k * 5 + r] * a[k * 5 + r];
    nd[k] = nu[k] = sqrt(s);
  }
  const double maxn = fmax(nu[0], fmax(nu[1], nu[2]));
  const double threshold_helper = (maxn * eps) * (maxn * eps) / (double)rows;
  const double downdate = sqrt(eps);
  int nonzero = cols;
  colpiv_step<0>(a, h, nu, nd, transp, nonzero, threshold_helper, downdate);
  colpiv_step<1>(a, h, nu, nd, transp, nonzero, threshold_helper, downdate);
  colpiv_step<2>(a, h, nu, nd, transp, nonzero, threshold_helper, downdate);
  // column permutation: perm = identity with the recorded transpositions applied in order
  int p0 = 0, p1 = 1, p2 = 2;
  if (transp[0] == 1) { const int t = p0; p0 = p1; p1 = t; } else if (transp[0] == 2) { const int t = p0; p0 = p2; p2 = t; }
  if (transp[1] == 2) { const int t = p1; p1 = p2; p2 = t; }
  x_out[0] = x_out[1] = x_out[2] = 0.0;
  if (nonzero == 0) return;
  double c[5];
#pragma unroll
  for (int r = 0; r < rows; ++r) c[r] = b[r];
  if (0 < nonzero) householder_apply<5>(&c[0], &a[0 * 5 + 1], h[0]);
  if (1 < nonzero) householder_apply<4>(&c[1], &a[1 * 5 + 2], h[1]);
  if (2 < nonzero) householder_apply<3>(&c[2], &a[2 * 5 + 3], h[2]);
  // back substitution on the leading nonzero x nonzero block
  if (nonzero == 3) {
    c[2] = c[2] / a[2 * 5 + 2];
    c[1] = (c[1] - a[2 * 5 + 1] * c[2]) / a[1 * 5 + 1];
    c[0] = ((c[0] - a[1 * 5 + 0] * c[1]) - a[2 * 5 + 0] * c[2]) / a[0 * 5 + 0];
  } else if (nonzero == 2) {
    c[1] = c[1] / a[1 * 5 + 1];
    c[0] = (c[0] - a[1 * 5 + 0] * c[1]) / a[0 * 5 + 0];
  } else {
    c[0] = c[0] / a[0 * 5 + 0];
  }
  const int pk[3] = {p0, p1, p2};
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (i < nonzero) {
      if (pk[i] == 0) x_out[0] = c[i]; else if (pk[i] == 1) x_out[1] = c[i]; else x_out[2] = c[i];
    }
}

// Solve the symmetric positive definite 6x6 system A y = b (A full row-major 36). Cholesky; returns false when not SPD / non-finite.
FM_HD bool cholesky6_solve(const double* A, const double* b, double* y) {
  // L L^T = A with one reciprocal per column and fused multiply-adds (this solve replaces Ceres' Householder QR of the stacked
  // Jacobian; it is not part of the bit-exact arithmetic, only of the 1e-4 pose tolerance)
  double L[36], inv[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double s = A[j * 6 + j];
#pragma unroll
    for (int k = 0; k < j; ++k) s = fma(-L[j * 6 + k], L[j * 6 + k], s);
    if (!(s > 0.0)) return false;
    inv[j] = rsqrt(s);          // one reciprocal square root instead of sqrt + divide on the serial critical path
    L[j * 6 + j] = s * inv[j];
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double t = A[i * 6 + j];
#pragma unroll
      for (int k = 0; k < j; ++k) t = fma(-L[i * 6 + k], L[j * 6 + k], t);
      L[i * 6 + j] = t * inv[j];
    }
  }
  double z[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s = fma(-L[i * 6 + k], z[k], s);
    z[i] = s * inv[i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double s = z[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) s = fma(-L[k * 6 + i], y[k], s);
    y[i] = s * inv[i];
  }
  for (int i = 0; i < 6; ++i) if (!isfinite(y[i])) return false;
  return true;
}

#undef FM_HD
}  // namespace m
}  // namespace floam
