// Scalar transform algebra shared by host (C-ABI helpers) and device (solve epilogue).
// One source for both, so Python helpers, the device loop and the parity tests run the same
// arithmetic.  Formulas follow the reference term by term (file:line cited per function).
#pragma once
#include <math.h>
#include "ica_common.cuh"

namespace ica {

// src/transformation.py:15-32
__host__ __device__ inline int nparams_of(int ttype) {
  switch (ttype) {
    case TRANSLATION: return 2;
    case EUCLIDEAN:   return 3;
    case SIMILARITY:  return 4;
    case AFFINITY:    return 6;
    case HOMOGRAPHY:  return 8;
    default:          return -1;
  }
}

__host__ __device__ inline int moment_degree_of(int ttype) {
  return ttype == HOMOGRAPHY ? 4 : (ttype == TRANSLATION ? 0 : 2);
}

// src/transformation.py:188-236 (row-major 3x3)
__host__ __device__ inline void params2matrix(const double* p, int ttype, double* m) {
  m[0] = 1; m[1] = 0; m[2] = 0; m[3] = 0; m[4] = 1; m[5] = 0; m[6] = 0; m[7] = 0; m[8] = 1;
  switch (ttype) {
    case TRANSLATION: m[2] = p[0]; m[5] = p[1]; break;
    case EUCLIDEAN: {
      double c = cos(p[2]), s = sin(p[2]);
      m[0] = c; m[1] = -s; m[2] = p[0]; m[3] = s; m[4] = c; m[5] = p[1];
    } break;
    case SIMILARITY:
      m[0] = 1 + p[2]; m[1] = -p[3]; m[2] = p[0]; m[3] = p[3]; m[4] = 1 + p[2]; m[5] = p[1];
      break;
    case AFFINITY:
      m[0] = 1 + p[2]; m[1] = p[3]; m[2] = p[0]; m[3] = p[4]; m[4] = 1 + p[5]; m[5] = p[1];
      break;
    case HOMOGRAPHY:
      m[0] = 1 + p[0]; m[1] = p[1]; m[2] = p[2]; m[3] = p[3]; m[4] = 1 + p[4]; m[5] = p[5];
      m[6] = p[6]; m[7] = p[7];
      break;
    default: break;
  }
}

// Matrix used by the warp: identity when every |p_i| < 1e-10 (src/bicubic_interpolation.py:173-175)
__host__ __device__ inline void warp_matrix(const double* p, int ttype, double* m) {
  int n = nparams_of(ttype);
  bool ident = true;
  for (int i = 0; i < n; ++i) ident = ident && (fabs(p[i]) < 1e-10);
  if (ident) { m[0] = 1; m[1] = 0; m[2] = 0; m[3] = 0; m[4] = 1; m[5] = 0; m[6] = 0; m[7] = 0; m[8] = 1; }
  else params2matrix(p, ttype, m);
}

// src/transformation.py:36-141: p <- params(M(p) M(dp)^-1), closed forms AS WRITTEN in the
// reference.  AFFINITY p[1] carries `d*d*ep` (tr.py:106) and HOMOGRAPHY p[4] lacks `-a + c*g`
// (tr.py:136): neither is exact matrix composition, both are required for trajectory parity
// (SURVEY.md Q9).
__host__ __device__ inline void update_transform(double* p, const double* dp, int ttype) {
  switch (ttype) {
    case TRANSLATION:
      p[0] -= dp[0]; p[1] -= dp[1];
      break;
    case EUCLIDEAN: {
      double a = cos(dp[2]), b = sin(dp[2]), c = dp[0], d = dp[1];
      double ap = cos(p[2]), bp = sin(p[2]), cp = p[0], dq = p[1];
      double cost = a * ap + b * bp;
      double sint = a * bp - b * ap;
      p[0] = cp - bp * (b * c - a * d) - ap * (a * c + b * d);
      p[1] = dq - bp * (a * c + b * d) + ap * (b * c - a * d);
      p[2] = atan2(sint, cost);
    } break;
    case SIMILARITY: {
      double a = dp[2], b = dp[3], c = dp[0], d = dp[1];
      double det = 2 * a + a * a + b * b + 1;
      if (det * det > 1e-10) {
        double ap = p[2], bp = p[3], cp = p[0], dq = p[1];
        p[0] = cp - bp * (-d - a * d + b * c) / det + (ap + 1) * (-c - a * c - b * d) / det;
        p[1] = dq + bp * (-c - a * c - b * d) / det + (ap + 1) * (-d - a * d + b * c) / det;
        p[2] = b * bp / det + (a + 1) * (ap + 1) / det - 1;
        p[3] = -b * (ap + 1) / det + bp * (a + 1) / det;
      }
    } break;
    case AFFINITY: {
      double a = dp[2], b = dp[3], c = dp[0], d = dp[4], e = dp[5], f = dp[1];
      double det = a - b * d + e + a * e + 1;
      if (det * det > 1e-10) {
        double ap = p[2], bp = p[3], cp = p[0], dq = p[4], ep = p[5], fp = p[1];
        p[0] = cp + (-f * bp - a * f * bp + c * d * bp) / det + (ap + 1) * (-c + b * f - c * e) / det;
        p[1] = fp + dq * (-c + b * f - c * e) / det +
               (-f + c * d - a * f - f * ep - a * f * ep + d * d * ep) / det;
        p[2] = ((1 + ap) * (1 + e) - d * bp) / det - 1;
        p[3] = (bp + a * bp - b - b * ap) / det;
        p[4] = (dq * (1 + e) - d - d * ep) / det;
        p[5] = (a + ep + a * ep + 1 - b * dq) / det - 1;
      }
    } break;
    case HOMOGRAPHY: {
      double a = dp[0], b = dp[1], c = dp[2], d = dp[3], e = dp[4], f = dp[5], g = dp[6], h = dp[7];
      double ap = p[0], bp = p[1], cp = p[2], dq = p[3], ep = p[4], fp = p[5], gp = p[6], hp = p[7];
      double det = f * hp + a * f * hp - c * d * hp + gp * (c - b * f + c * e) - a + b * d - e - a * e - 1;
      if (det * det > 1e-10) {
        p[0] = ((d * bp - f * g * bp) + cp * (g - d * h + g * e) + (ap + 1) * (f * h - e - 1)) / det - 1;
        p[1] = (h * cp + a * h * cp - b * g * cp - bp - a * bp + c * g * bp + b - c * h + b * ap - c * h * ap) / det;
        p[2] = (f * bp + a * f * bp - c * d * bp + (ap + 1) * (c - b * f + c * e) + cp * (-a + b * d - e - a * e - 1)) / det;
        p[3] = (fp * (g - d * h + g * e) + d - f * g + d * ep - f * g * ep + dq * (f * h - e - 1)) / det;
        p[4] = (b * dq - c * h * dq + h * fp + a * h * fp - b * g * fp - ep - a * ep + c * g * ep - 1) / det - 1;
        p[5] = (dq * (c - b * f + c * e) + f + a * f - c * d + f * ep + a * f * ep - c * d * ep + fp * (-a + b * d - e - a * e - 1)) / det;
        p[6] = (d * hp - f * g * hp + g - d * h + g * e + gp * (f * h - e - 1)) / det;
        p[7] = (h + a * h - b * g + b * gp - c * h * gp - hp - a * hp + c * g * hp) / det;
      }
    } break;
    default: break;
  }
}

// src/zoom.py:62-125: nu = max(nxx/nx, nyy/ny)
__host__ __device__ inline void zoom_in_parameters(const double* p, int ttype, double nx, double ny,
                                                   double nxx, double nyy, double* out) {
  double fx = nxx / nx, fy = nyy / ny;
  double nu = fx > fy ? fx : fy;
  int n = nparams_of(ttype);
  for (int i = 0; i < n; ++i) out[i] = p[i];
  if (ttype == HOMOGRAPHY) {
    out[2] = p[2] * nu; out[5] = p[5] * nu; out[6] = p[6] / nu; out[7] = p[7] / nu;
  } else {
    out[0] = p[0] * nu; out[1] = p[1] * nu;
  }
}

// src/derivatives.py:110-130: inverse by LU with partial pivoting (what np.linalg.inv does
// through LAPACK getrf/getri); zero matrix when a pivot is exactly zero (LinAlgError branch).
__host__ __device__ inline void inverse_hessian(const double* H, int n, double* Hinv) {
  double a[ICA_MAX_PARAMS][2 * ICA_MAX_PARAMS];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) { a[i][j] = H[i * n + j]; a[i][n + j] = (i == j) ? 1.0 : 0.0; }
  bool singular = false;
  for (int k = 0; k < n && !singular; ++k) {
    int piv = k; double best = fabs(a[k][k]);
    for (int i = k + 1; i < n; ++i) { double v = fabs(a[i][k]); if (v > best) { best = v; piv = i; } }
    if (!(best > 0.0)) { singular = true; break; }
    if (piv != k) for (int j = 0; j < 2 * n; ++j) { double t = a[k][j]; a[k][j] = a[piv][j]; a[piv][j] = t; }
    double inv = 1.0 / a[k][k];
    for (int j = 0; j < 2 * n; ++j) a[k][j] *= inv;
    for (int i = 0; i < n; ++i) {
      if (i == k) continue;
      double f = a[i][k];
      if (f != 0.0) for (int j = 0; j < 2 * n; ++j) a[i][j] -= f * a[k][j];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Hinv[i * n + j] = singular ? 0.0 : a[i][n + j];
}

// ---- Jacobian as monomials (src/derivatives.py:31-68) ----------------------------------------
// J[k] = coef * x^a * y^b (coef 0 = entry absent); jx = d x'/d p_k, jy = d y'/d p_k.
struct Mono { signed char coef, a, b; };

__host__ __device__ inline void jacobian_monomials(int ttype, Mono* jx, Mono* jy) {
  const Mono Z = {0, 0, 0}, ONE = {1, 0, 0}, X = {1, 1, 0}, Y = {1, 0, 1}, NY = {-1, 0, 1};
  const Mono NXX = {-1, 2, 0}, NXY = {-1, 1, 1}, NYY = {-1, 0, 2};
  for (int k = 0; k < ICA_MAX_PARAMS; ++k) { jx[k] = Z; jy[k] = Z; }
  switch (ttype) {
    case TRANSLATION: jx[0] = ONE; jy[1] = ONE; break;
    case EUCLIDEAN:   jx[0] = ONE; jx[2] = NY; jy[1] = ONE; jy[2] = X; break;
    case SIMILARITY:  jx[0] = ONE; jx[2] = X; jx[3] = NY; jy[1] = ONE; jy[2] = Y; jy[3] = X; break;
    case AFFINITY:    jx[0] = ONE; jx[2] = X; jx[3] = Y; jy[1] = ONE; jy[4] = X; jy[5] = Y; break;
    case HOMOGRAPHY:
      jx[0] = X; jx[1] = Y; jx[2] = ONE; jx[6] = NXX; jx[7] = NXY;
      jy[3] = X; jy[4] = Y; jy[5] = ONE; jy[6] = NXY; jy[7] = NYY;
      break;
    default: break;
  }
}

// Assemble H (n x n) and b (n) from the moment sums.  mom[k * kYPow + bpow]: k < 3*(dh+1) are
// MH[ij][a] (k = ij*(dh+1)+a, ij: 0=xx 1=xy 2=yy), then MB[i][a] (k = 3*(dh+1) + i*(dh/2+1) + a).
__host__ __device__ inline void assemble_system(const double* mom, int dh, int ttype, double* H, double* bvec) {
  Mono jx[ICA_MAX_PARAMS], jy[ICA_MAX_PARAMS];
  jacobian_monomials(ttype, jx, jy);
  const int n = nparams_of(ttype);
  const int hw = dh + 1, bw = dh / 2 + 1, boff = 3 * hw;
  for (int k = 0; k < n; ++k) {
    for (int l = 0; l < n; ++l) {
      double s = 0.0;
      // xx
      if (jx[k].coef && jx[l].coef)
        s += (double)(jx[k].coef * jx[l].coef) * mom[(0 * hw + jx[k].a + jx[l].a) * kYPow + jx[k].b + jx[l].b];
      // xy (both cross terms)
      if (jx[k].coef && jy[l].coef)
        s += (double)(jx[k].coef * jy[l].coef) * mom[(1 * hw + jx[k].a + jy[l].a) * kYPow + jx[k].b + jy[l].b];
      if (jy[k].coef && jx[l].coef)
        s += (double)(jy[k].coef * jx[l].coef) * mom[(1 * hw + jy[k].a + jx[l].a) * kYPow + jy[k].b + jx[l].b];
      // yy
      if (jy[k].coef && jy[l].coef)
        s += (double)(jy[k].coef * jy[l].coef) * mom[(2 * hw + jy[k].a + jy[l].a) * kYPow + jy[k].b + jy[l].b];
      H[k * n + l] = s;
    }
    double s = 0.0;
    if (jx[k].coef) s += (double)jx[k].coef * mom[(boff + 0 * bw + jx[k].a) * kYPow + jx[k].b];
    if (jy[k].coef) s += (double)jy[k].coef * mom[(boff + 1 * bw + jy[k].a) * kYPow + jy[k].b];
    bvec[k] = s;
  }
}

}  // namespace ica
