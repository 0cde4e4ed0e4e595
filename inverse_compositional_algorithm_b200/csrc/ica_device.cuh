// Device helpers shared by the per-iteration kernel and the stand-alone warp kernel.
#pragma once
#include "ica_common.cuh"

namespace ica {

// Warp coefficients in DISPLACEMENT form, fp32.  With M = warp_matrix(p) (M22 == 1):
//   x' - x = ((M00-1) x + M01 y + M02 - x (M20 x + M21 y)) / (1 + M20 x + M21 y)
// The differences M00-1, M11-1 are formed in fp64 before the conversion, so the per-pixel fp32
// arithmetic only ever touches quantities of the size of the displacement (tens of pixels,
// ulp ~1e-6 px) instead of absolute coordinates (ulp 6e-5 px at x~1000, 5e-4 at x~8000).
struct WarpCoef { float d00, m01, m02, m10, d11, m12, m20, m21; };

__host__ __device__ inline WarpCoef make_warp_coef(const double* m) {
  // normalise by M22 (always 1 for matrices from params2matrix; general matrices for ica_warp)
  double s = 1.0 / m[8];
  WarpCoef c;
  c.d00 = (float)(m[0] * s - 1.0); c.m01 = (float)(m[1] * s); c.m02 = (float)(m[2] * s);
  c.m10 = (float)(m[3] * s); c.d11 = (float)(m[4] * s - 1.0); c.m12 = (float)(m[5] * s);
  c.m20 = (float)(m[6] * s); c.m21 = (float)(m[7] * s);
  return c;
}

// 1/z: hardware approximation (MUFU.RCP, <= 1 ulp) refined by one Newton step -- 3 instructions instead of the ~11 of
// the correctly rounded __frcp_rn; the result is within 1 ulp, far inside the 1e-3 tie band handled below.
__device__ __forceinline__ float fast_rcp(float z) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z));
  const float e = fmaf(-z, r, 1.0f);
  return fmaf(r, e, r);
}
__device__ __forceinline__ float approx_rcp(float z) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z));
  return r;
}

// Projects integer pixel (x, y): integer tap origin (cx, cy) = floor(x'), floor(y') and the
// fractions.  Returns false when the coordinates are unusable (z <= 0, non-finite).
// Pixels whose projected coordinate falls within ~2.5e-4 of an integer (where fp32 rounding could
// pick the other tap set, or flip the NaN footprint) are re-evaluated in fp64 with exactly the
// operation order of skimage's _transform_projective -- (M0*x + M1*y) + M2, same for z, then
// the quotient -- so tap selection agrees with the reference bit for bit.  m64 is the 3x3
// matrix (row-major, fp64, typically in shared memory); the branch is taken by ~0.1% of pixels.
__device__ __forceinline__ bool project_px(const WarpCoef& k, const double* m64, int x, int y, int& cx,
                                           int& cy, float& tx, float& ty) {
  float fx = (float)x, fy = (float)y;
  float zm1 = fmaf(k.m20, fx, k.m21 * fy);
  float nx_ = fmaf(k.d00, fx, fmaf(k.m01, fy, k.m02)) - fx * zm1;
  float ny_ = fmaf(k.m10, fx, fmaf(k.d11, fy, k.m12)) - fy * zm1;
  float z = 1.0f + zm1;
  float rz = fast_rcp(z);
  float dx = nx_ * rz, dy = ny_ * rz;
  bool ok = (z > 0.0f) && (fabsf(dx) < 1.0e6f) && (fabsf(dy) < 1.0e6f);  // false for NaN too
  dx = ok ? dx : 0.0f; dy = ok ? dy : 0.0f;
  float flx = floorf(dx), fly = floorf(dy);
  tx = dx - flx; ty = dy - fly;
  cx = x + (int)flx; cy = y + (int)fly;
  // tie band: a comfortable multiple of the fp32 error of dx, dy (a few ulps of the displacement: 2.5e-4 px covers
  // displacements of ~100 px, the second term takes over for the huge ones); ~0.1% of the pixels take the branch
  const float kTie = fmaf(1.0e-6f, fabsf(dx) + fabsf(dy), 2.5e-4f);
  if (ok && (fminf(tx, ty) < kTie || fmaxf(tx, ty) > 1.0f - kTie)) {
    const double xd = (double)x, yd = (double)y;
    const double xx = __dadd_rn(__dadd_rn(__dmul_rn(m64[0], xd), __dmul_rn(m64[1], yd)), m64[2]);
    const double yy = __dadd_rn(__dadd_rn(__dmul_rn(m64[3], xd), __dmul_rn(m64[4], yd)), m64[5]);
    const double zz = __dadd_rn(__dadd_rn(__dmul_rn(m64[6], xd), __dmul_rn(m64[7], yd)), m64[8]);
    const double c = __ddiv_rn(xx, zz), r = __ddiv_rn(yy, zz);
    const double fc = floor(c), fr = floor(r);
    if (fabs(c) < 1.0e9 && fabs(r) < 1.0e9) {
      cx = (int)fc; cy = (int)fr;
      tx = (float)(c - fc); ty = (float)(r - fr);
    }
  }
  return ok;
}

// Catmull-Rom / Keys(a=-1/2) weights; algebraically the polynomial of skimage's
// cubic_interpolation (= src/bicubic_interpolation.py:39-41) grouped by sample.
__device__ __forceinline__ void keys_weights(float t, float& w0, float& w1, float& w2, float& w3) {
  float t2 = t * t;
  w0 = t * fmaf(t, fmaf(-0.5f, t, 1.0f), -0.5f);
  w1 = fmaf(t2, fmaf(1.5f, t, -2.5f), 1.0f);
  w2 = t * fmaf(t, fmaf(-1.5f, t, 2.0f), 0.5f);
  w3 = t2 * fmaf(0.5f, t, -0.5f);
}

// The same weights for two coordinates at once in packed fp32 (lane's two pixels): identical arithmetic per half.
__device__ __forceinline__ void keys_weights2(float2 t, float2& w0, float2& w1, float2& w2, float2& w3) {
  const float2 one = make_float2(1.0f, 1.0f), half = make_float2(0.5f, 0.5f), mhalf = make_float2(-0.5f, -0.5f);
  const float2 t2 = __fmul2_rn(t, t);
  w0 = __fmul2_rn(t, __ffma2_rn(t, __ffma2_rn(mhalf, t, one), mhalf));
  w1 = __ffma2_rn(t2, __ffma2_rn(make_float2(1.5f, 1.5f), t, make_float2(-2.5f, -2.5f)), one);
  w2 = __fmul2_rn(t, __ffma2_rn(t, __ffma2_rn(make_float2(-1.5f, -1.5f), t, make_float2(2.0f, 2.0f)), half));
  w3 = __fmul2_rn(t2, __ffma2_rn(half, t, mhalf));
}

// rho'(t2) of src/image_optimisation.py:17-53 (TRUNCATED_QUADRATIC element-wise)
__device__ __forceinline__ float rho_prime(float t2, float lambda2, int rtype) {
  switch (rtype) {
    case TRUNCATED_QUADRATIC: return t2 < lambda2 ? 1.0f : 0.0f;
    case GERMAN_MCCLURE: { float d = lambda2 + t2; return lambda2 * approx_rcp(d * d); }   // weights: 1 ulp is plenty
    case LORENTZIAN: return approx_rcp(lambda2 + t2);
    case CHARBONNIER: return rsqrtf(t2 + lambda2);
    default: return 1.0f;
  }
}

// Bicubic sample of one channel from global memory with the NaN-footprint rule of
// skimage.transform.warp(order=3, mode='constant', cval=nan): any tap outside the
// image poisons the pixel.  Generic (slow) path.
template <int C>
__device__ __forceinline__ float sample_global(const float* __restrict__ img, int pitch, int nx,
                                               int ny, int cx, int cy, int ch, const float* wx,
                                               const float* wy) {
  if (cx < 1 || cy < 1 || cx + 2 > nx - 1 || cy + 2 > ny - 1) return __int_as_float(0x7fc00000);
  const float* base = img + (long long)(cy - 1) * pitch + (cx - 1) * C + ch;
  float acc = 0.0f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const float* r = base + (long long)a * pitch;
    float v = wx[0] * __ldg(r) + wx[1] * __ldg(r + C) + wx[2] * __ldg(r + 2 * C) + wx[3] * __ldg(r + 3 * C);
    acc = fmaf(wy[a], v, acc);
  }
  return acc;
}

}  // namespace ica
