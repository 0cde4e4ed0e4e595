// Helper API of the reference on materialised arrays: the building blocks notebooks and tests import
// next to the three drivers.  The drivers themselves never materialise these arrays (K2 recomputes everything per
// pixel); these entry points exist so that a caller of the reference's helpers finds them here too, computed on the
// GPU in float64 like the reference (src/image_optimisation.py, src/derivatives.py, src/transformation.py).
// Reductions are deterministic: per-block partial sums in fixed slots, summed in order by one final block.
#include <stdint.h>
#include <math.h>
#include <algorithm>
#include "ica_common.cuh"

namespace ica {
namespace {

constexpr int kRedThreads = 256;

__device__ __forceinline__ double zero_if_nonfinite(double v) { return isfinite(v) ? v : 0.0; }

// rho'(t2) in float64, element-wise (src/image_optimisation.py:17-53; TRUNCATED_QUADRATIC element-wise)
__device__ __forceinline__ double rhop_f64(double t2, double lambda2, int type) {
  switch (type) {
    case TRUNCATED_QUADRATIC: return t2 < lambda2 ? 1.0 : 0.0;
    case GERMAN_MCCLURE: return lambda2 / ((lambda2 + t2) * (lambda2 + t2));
    case LORENTZIAN: return 1.0 / (lambda2 + t2);
    case CHARBONNIER: return 1.0 / sqrt(t2 + lambda2);
    default: return 1.0;
  }
}

__global__ void rhop_kernel(const double* __restrict__ t2, long long n, double lambda2, int type, double* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = rhop_f64(t2[i], lambda2, type);
}

// io.robust_error_function (io.py:56-79): t2 = sum_c DI_c^2 with non-finite DI -> 0; rho'(t2), non-finite -> 0
__global__ void robust_error_kernel(const double* __restrict__ DI, long long npix, int C, double lambda2, int type,
                                    double* __restrict__ rho) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    double t2 = 0.0;
    for (int c = 0; c < C; ++c) { const double d = zero_if_nonfinite(DI[i * C + c]); t2 += d * d; }
    rho[i] = isfinite(t2) ? rhop_f64(t2, lambda2, type) : 0.0;
  }
}

// io.steepest_descent_images (io.py:158-194): DIJ[y,x,c,k] = Ix[y,x,c] J[y,x,k] + Iy[y,x,c] J[y,x,k+n]
__global__ void steepest_descent_kernel(const double* __restrict__ Ix, const double* __restrict__ Iy,
                                        const double* __restrict__ J, long long npix, int C, int n, double* __restrict__ DIJ) {
  const long long total = npix * C * n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % n);
    const long long pc = e / n;
    const long long px = pc / C;
    DIJ[e] = Ix[pc] * J[px * 2 * n + k] + Iy[pc] * J[px * 2 * n + k + n];
  }
}

// Per-block partial sums of  sum_px w(px) sum_c DIJ[px,c,k] * X  with X = DIJ[px,c,m] (Hessian, nout = n*n) or
// X = DI[px,c] (independent vector, nout = n); non-finite factors are zero-filled element-wise (de.py:84, io.py:98-99).
__global__ void __launch_bounds__(kRedThreads) dij_reduce_kernel(
    const double* __restrict__ DIJ, const double* __restrict__ DI, const double* __restrict__ rho, long long npix, int C,
    int n, double* __restrict__ partials /* [gridDim.x][nout] */) {
  const int nout = DI ? n : n * n;
  double acc[ICA_MAX_PARAMS * ICA_MAX_PARAMS];
  for (int o = 0; o < nout; ++o) acc[o] = 0.0;
  for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < npix; px += (long long)gridDim.x * blockDim.x) {
    const double w = rho ? rho[px] : 1.0;
    for (int c = 0; c < C; ++c) {
      double g[ICA_MAX_PARAMS];
      for (int k = 0; k < n; ++k) g[k] = zero_if_nonfinite(DIJ[(px * C + c) * n + k]);
      if (DI) {
        const double d = zero_if_nonfinite(DI[px * C + c]) * w;
        for (int k = 0; k < n; ++k) acc[k] += g[k] * d;
      } else {
        for (int k = 0; k < n; ++k) {
          const double gw = g[k] * w;
          for (int m = 0; m < n; ++m) acc[k * n + m] += gw * g[m];
        }
      }
    }
  }
  __shared__ double sred[kRedThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 0; o < nout; ++o) {
    double v = acc[o];
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0) sred[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w2 = 0; w2 < kRedThreads / 32; ++w2) t += sred[w2];
      partials[(long long)blockIdx.x * nout + o] = t;
    }
    __syncthreads();
  }
}

__global__ void final_sum_kernel(const double* __restrict__ partials, int nblocks, int nout, double* __restrict__ out) {
  const int o = threadIdx.x;
  if (o >= nout) return;
  double t = 0.0;
  for (int b = 0; b < nblocks; ++b) t += partials[(long long)b * nout + o];
  out[o] = t;
}

// tr.transform_image (tr.py:266-318) = skimage.transform.warp(image, inverse_map, order=1, mode='constant', cval=0,
// clip=True, preserve_range=True): bilinear, a tap outside the image reads 0.  matrix maps output (col,row) -> input.
__global__ void warp_bilinear_kernel(const double* __restrict__ img, int nx, int ny, int C, const double* __restrict__ m,
                                     double* __restrict__ out, double* __restrict__ mm /* [4]: in lo, in hi, out lo, out hi */) {
  const long long npix = (long long)nx * ny;
  double ilo = 1e300, ihi = -1e300, olo = 1e300, ohi = -1e300;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % nx), y = (int)(i / nx);
    const double z = m[6] * x + m[7] * y + m[8];
    const double c = (m[0] * x + m[1] * y + m[2]) / z, r = (m[3] * x + m[4] * y + m[5]) / z;
    const bool fin = isfinite(c) && isfinite(r) && fabs(c) < 1e9 && fabs(r) < 1e9;
    const double cs = fin ? c : -1e9, rs = fin ? r : -1e9;
    const double c0 = floor(cs), r0 = floor(rs);
    const double xc = cs - c0, xr = rs - r0;
    const long long ic = (long long)c0, ir = (long long)r0;
    for (int ch = 0; ch < C; ++ch) {
      auto tap = [&](long long rr, long long cc) -> double {
        return (rr >= 0 && rr < ny && cc >= 0 && cc < nx) ? img[(rr * nx + cc) * C + ch] : 0.0;
      };
      const double top = (1.0 - xc) * tap(ir, ic) + xc * tap(ir, ic + 1);
      const double bot = (1.0 - xc) * tap(ir + 1, ic) + xc * tap(ir + 1, ic + 1);
      const double v = (1.0 - xr) * top + xr * bot;
      out[i * C + ch] = v;
      const double iv = img[i * C + ch];
      if (iv == iv) { ilo = fmin(ilo, iv); ihi = fmax(ihi, iv); }
      if (v == v) { olo = fmin(olo, v); ohi = fmax(ohi, v); }
    }
  }
  // min/max through ordered atomics on the bit patterns of non-negative / negative doubles is overkill here:
  // one atomic per thread on a tiny grid
  auto amin = [](double* a, double v) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(a);
    unsigned long long old = *p;
    while (__longlong_as_double((long long)old) > v) {
      const unsigned long long prev = atomicCAS(p, old, (unsigned long long)__double_as_longlong(v));
      if (prev == old) break;
      old = prev;
    }
  };
  auto amax = [](double* a, double v) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(a);
    unsigned long long old = *p;
    while (__longlong_as_double((long long)old) < v) {
      const unsigned long long prev = atomicCAS(p, old, (unsigned long long)__double_as_longlong(v));
      if (prev == old) break;
      old = prev;
    }
  };
  if (ilo <= ihi) { amin(mm + 0, ilo); amax(mm + 1, ihi); }
  if (olo <= ohi) { amin(mm + 2, olo); amax(mm + 3, ohi); }
}

// skimage's _clip_warp_output: clip to the input's [min,max]; when cval (0) lies outside that range but inside the
// output's range, pixels that are exactly cval keep it
__global__ void clip_like_skimage_kernel(double* __restrict__ out, long long n, const double* __restrict__ mm) {
  const double lo = mm[0], hi = mm[1];
  const double cval = 0.0;
  const bool keep_cval = !(lo <= cval && cval <= hi) && (mm[2] <= cval && cval <= mm[3]);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = out[i];
    if (v != v || (keep_cval && v == cval)) continue;
    out[i] = fmin(fmax(v, lo), hi);
  }
}

// bi.bicubic_interpolation_image (bicubic_interpolation.py:121-152 with :66-118 and tr.project :144-186): the
// IPOL-style warp.  NaN / 0 when the projected point is closer than delta to the border; otherwise Catmull-Rom on
// int(x)-s .. int(x)+2s (s = sign, int() truncates), indices clamped to the image, fraction from the clamped centre.
__device__ __forceinline__ double keys_f64(double v0, double v1, double v2, double v3, double x) {
  return v1 + 0.5 * x * (v2 - v0 + x * (2.0 * v0 - 5.0 * v1 + 4.0 * v2 - v3 + x * (3.0 * (v1 - v2) + v3 - v0)));
}
__global__ void warp_ipol_kernel(const double* __restrict__ img, int nx, int ny, int C, const double* __restrict__ p, int np_,
                                 int nan_out, int delta, double* __restrict__ out) {
  const long long npix = (long long)nx * ny;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const double xj = (double)(i % nx), yi = (double)(i / nx);
    double x, y;
    if (np_ == 2) { x = xj + p[0]; y = yi + p[1]; }
    else if (np_ == 3) { const double c = cos(p[2]), s = sin(p[2]); x = c * xj - s * yi + p[0]; y = s * xj + c * yi + p[1]; }
    else if (np_ == 4) { x = (1 + p[2]) * xj - p[3] * yi + p[0]; y = p[3] * xj + (1 + p[2]) * yi + p[1]; }
    else if (np_ == 6) { x = (1 + p[2]) * xj + p[3] * yi + p[0]; y = p[4] * xj + (1 + p[5]) * yi + p[1]; }
    else { const double d = p[6] * xj + p[7] * yi + 1; x = ((1 + p[0]) * xj + p[1] * yi + p[2]) / d; y = (p[3] * xj + (1 + p[4]) * yi + p[5]) / d; }
    const bool outside = (x < delta) || (x > nx - 1 - delta) || (y < delta) || (y > ny - 1 - delta) || !(x == x) || !(y == y);
    if (outside) {
      const double v = nan_out ? __longlong_as_double(0x7ff8000000000000ll) : 0.0;
      for (int k = 0; k < C; ++k) out[i * C + k] = v;
      continue;
    }
    const int sx = x < 0 ? -1 : 1, sy = y < 0 ? -1 : 1;
    const long long ix = (long long)x, iy = (long long)y;          // truncation toward zero, like int()
    auto cl = [](long long v, int n) -> int { return (int)(v < 0 ? 0 : (v >= n ? n - 1 : v)); };
    const int cx[4] = {cl(ix - sx, nx), cl(ix, nx), cl(ix + sx, nx), cl(ix + 2 * sx, nx)};
    const int cy[4] = {cl(iy - sy, ny), cl(iy, ny), cl(iy + sy, ny), cl(iy + 2 * sy, ny)};
    const double fx = x - cx[1], fy = y - cy[1];
    for (int k = 0; k < C; ++k) {
      double v[4];
      for (int a = 0; a < 4; ++a)
        v[a] = keys_f64(img[((long long)cy[0] * nx + cx[a]) * C + k], img[((long long)cy[1] * nx + cx[a]) * C + k],
                        img[((long long)cy[2] * nx + cx[a]) * C + k], img[((long long)cy[3] * nx + cx[a]) * C + k], fy);
      out[i * C + k] = keys_f64(v[0], v[1], v[2], v[3], fx);
    }
  }
}

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  template <typename T> T* as() { return static_cast<T*>(p); }
};

int require_gpu_() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) {
    cudaGetLastError();
    set_error("no CUDA device is visible; libica_b200 has no CPU fallback");
    return ICA_ERR_NO_DEVICE;
  }
  return ICA_OK;
}

int grid_for(long long n) { return (int)std::min<long long>((n + kRedThreads - 1) / kRedThreads, 148 * 8); }

#define H_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { set_error("%s failed: %s", #expr, cudaGetErrorString(e__)); return ICA_ERR_CUDA; } } while (0)

}  // namespace
}  // namespace ica

using namespace ica;

extern "C" {

int ica_rhop_host(const double* t2, int64_t count, double lambda_, int32_t robust_type, double* out) {
  if (!t2 || !out || count < 0) { set_error("NULL argument"); return ICA_ERR_INVALID; }
  if (robust_type < QUADRATIC || robust_type > CHARBONNIER) { set_error("Unknown type for robust error function"); return ICA_ERR_INVALID; }
  if (int rc = require_gpu_()) return rc;
  if (count == 0) return ICA_OK;
  DevBuf a, b;
  H_CUDA(a.alloc(count * sizeof(double))); H_CUDA(b.alloc(count * sizeof(double)));
  H_CUDA(cudaMemcpy(a.p, t2, count * sizeof(double), cudaMemcpyHostToDevice));
  rhop_kernel<<<grid_for(count), kRedThreads>>>(a.as<double>(), count, lambda_ * lambda_, robust_type, b.as<double>());
  H_CUDA(cudaGetLastError());
  H_CUDA(cudaMemcpy(out, b.p, count * sizeof(double), cudaMemcpyDeviceToHost));
  return ICA_OK;
}

int ica_robust_error_host(const double* DI, int32_t height, int32_t width, int32_t channels, double lambda_,
                          int32_t robust_type, double* rho_out) {
  if (!DI || !rho_out || height < 1 || width < 1 || channels < 1) { set_error("invalid argument"); return ICA_ERR_INVALID; }
  if (robust_type < QUADRATIC || robust_type > CHARBONNIER) { set_error("Unknown type for robust error function"); return ICA_ERR_INVALID; }
  if (int rc = require_gpu_()) return rc;
  const long long npix = (long long)height * width;
  DevBuf a, b;
  H_CUDA(a.alloc(npix * channels * sizeof(double))); H_CUDA(b.alloc(npix * sizeof(double)));
  H_CUDA(cudaMemcpy(a.p, DI, npix * channels * sizeof(double), cudaMemcpyHostToDevice));
  robust_error_kernel<<<grid_for(npix), kRedThreads>>>(a.as<double>(), npix, channels, lambda_ * lambda_, robust_type, b.as<double>());
  H_CUDA(cudaGetLastError());
  H_CUDA(cudaMemcpy(rho_out, b.p, npix * sizeof(double), cudaMemcpyDeviceToHost));
  return ICA_OK;
}

int ica_steepest_descent_host(const double* Ix, const double* Iy, const double* J, int32_t height, int32_t width,
                              int32_t channels, int32_t nparams, double* DIJ_out) {
  if (!Ix || !Iy || !J || !DIJ_out || height < 1 || width < 1 || channels < 1 || nparams < 1 || nparams > ICA_MAX_PARAMS) {
    set_error("invalid argument"); return ICA_ERR_INVALID;
  }
  if (int rc = require_gpu_()) return rc;
  const long long npix = (long long)height * width;
  DevBuf dx, dy, dj, dout;
  H_CUDA(dx.alloc(npix * channels * 8)); H_CUDA(dy.alloc(npix * channels * 8));
  H_CUDA(dj.alloc(npix * 2 * nparams * 8)); H_CUDA(dout.alloc(npix * channels * nparams * 8));
  H_CUDA(cudaMemcpy(dx.p, Ix, npix * channels * 8, cudaMemcpyHostToDevice));
  H_CUDA(cudaMemcpy(dy.p, Iy, npix * channels * 8, cudaMemcpyHostToDevice));
  H_CUDA(cudaMemcpy(dj.p, J, npix * 2 * nparams * 8, cudaMemcpyHostToDevice));
  steepest_descent_kernel<<<grid_for(npix * channels * nparams), kRedThreads>>>(dx.as<double>(), dy.as<double>(), dj.as<double>(), npix,
                                                                                channels, nparams, dout.as<double>());
  H_CUDA(cudaGetLastError());
  H_CUDA(cudaMemcpy(DIJ_out, dout.p, npix * channels * nparams * 8, cudaMemcpyDeviceToHost));
  return ICA_OK;
}

// H (DI == NULL, n x n out) or b (DI given, n out); rho may be NULL (quadratic)
int ica_dij_reduce_host(const double* DIJ, const double* DI, const double* rho, int32_t height, int32_t width,
                        int32_t channels, int32_t nparams, double* out) {
  if (!DIJ || !out || height < 1 || width < 1 || channels < 1 || nparams < 1 || nparams > ICA_MAX_PARAMS) {
    set_error("invalid argument"); return ICA_ERR_INVALID;
  }
  if (int rc = require_gpu_()) return rc;
  const long long npix = (long long)height * width;
  const int nout = DI ? nparams : nparams * nparams;
  const int blocks = (int)std::min<long long>((npix + kRedThreads - 1) / kRedThreads, 148 * 4);
  DevBuf dj, dd, dr, dp, dout;
  H_CUDA(dj.alloc(npix * channels * nparams * 8));
  H_CUDA(cudaMemcpy(dj.p, DIJ, npix * channels * nparams * 8, cudaMemcpyHostToDevice));
  if (DI) { H_CUDA(dd.alloc(npix * channels * 8)); H_CUDA(cudaMemcpy(dd.p, DI, npix * channels * 8, cudaMemcpyHostToDevice)); }
  if (rho) { H_CUDA(dr.alloc(npix * 8)); H_CUDA(cudaMemcpy(dr.p, rho, npix * 8, cudaMemcpyHostToDevice)); }
  H_CUDA(dp.alloc((size_t)blocks * nout * 8)); H_CUDA(dout.alloc(nout * 8));
  dij_reduce_kernel<<<blocks, kRedThreads>>>(dj.as<double>(), DI ? dd.as<double>() : nullptr, rho ? dr.as<double>() : nullptr, npix,
                                             channels, nparams, dp.as<double>());
  H_CUDA(cudaGetLastError());
  final_sum_kernel<<<1, 64>>>(dp.as<double>(), blocks, nout, dout.as<double>());
  H_CUDA(cudaGetLastError());
  H_CUDA(cudaMemcpy(out, dout.p, nout * 8, cudaMemcpyDeviceToHost));
  return ICA_OK;
}

int ica_transform_image_host(const double* image, int32_t height, int32_t width, int32_t channels, const double* matrix9,
                             double* out) {
  if (!image || !matrix9 || !out || height < 1 || width < 1 || channels < 1) { set_error("invalid argument"); return ICA_ERR_INVALID; }
  if (int rc = require_gpu_()) return rc;
  const long long n = (long long)height * width * channels;
  DevBuf di, dout, dm, dmm;
  H_CUDA(di.alloc(n * 8)); H_CUDA(dout.alloc(n * 8)); H_CUDA(dm.alloc(9 * 8)); H_CUDA(dmm.alloc(4 * 8));
  H_CUDA(cudaMemcpy(di.p, image, n * 8, cudaMemcpyHostToDevice));
  H_CUDA(cudaMemcpy(dm.p, matrix9, 9 * 8, cudaMemcpyHostToDevice));
  const double init[4] = {1e300, -1e300, 1e300, -1e300};
  H_CUDA(cudaMemcpy(dmm.p, init, sizeof(init), cudaMemcpyHostToDevice));
  warp_bilinear_kernel<<<grid_for((long long)height * width), kRedThreads>>>(di.as<double>(), width, height, channels, dm.as<double>(),
                                                                             dout.as<double>(), dmm.as<double>());
  H_CUDA(cudaGetLastError());
  clip_like_skimage_kernel<<<grid_for(n), kRedThreads>>>(dout.as<double>(), n, dmm.as<double>());
  H_CUDA(cudaGetLastError());
  H_CUDA(cudaMemcpy(out, dout.p, n * 8, cudaMemcpyDeviceToHost));
  return ICA_OK;
}

int ica_warp_ipol_host(const double* image, int32_t height, int32_t width, int32_t channels, const double* params,
                       int32_t nparams, int32_t nanifoutside, int32_t delta, double* out) {
  if (!image || !params || !out || height < 1 || width < 1 || channels < 1) { set_error("invalid argument"); return ICA_ERR_INVALID; }
  if (nparams != 2 && nparams != 3 && nparams != 4 && nparams != 6 && nparams != 8) { set_error("Invalid transformation type"); return ICA_ERR_INVALID; }
  if (int rc = require_gpu_()) return rc;
  const long long n = (long long)height * width * channels;
  DevBuf di, dout, dp;
  H_CUDA(di.alloc(n * 8)); H_CUDA(dout.alloc(n * 8)); H_CUDA(dp.alloc(8 * 8));
  H_CUDA(cudaMemcpy(di.p, image, n * 8, cudaMemcpyHostToDevice));
  H_CUDA(cudaMemcpy(dp.p, params, nparams * 8, cudaMemcpyHostToDevice));
  warp_ipol_kernel<<<grid_for((long long)height * width), kRedThreads>>>(di.as<double>(), width, height, channels, dp.as<double>(), nparams,
                                                                         nanifoutside, delta, dout.as<double>());
  H_CUDA(cudaGetLastError());
  H_CUDA(cudaMemcpy(out, dout.p, n * 8, cudaMemcpyDeviceToHost));
  return ICA_OK;
}

}  // extern "C"
