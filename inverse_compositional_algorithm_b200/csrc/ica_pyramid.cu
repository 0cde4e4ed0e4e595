// K0: one pyramid level with the semantics of
//   skimage.transform.rescale(img, nu, mode='constant', cval=0, order=3, anti_aliasing=True,
//                             channel_axis=2, preserve_range=True)
// as called at src/inverse_compositional_algorithm.py:333-336 (third-party arithmetic:
// scikit-image 0.24 -> scipy.ndimage.gaussian_filter + scipy.ndimage.zoom).
//
// Per axis the chain  Gaussian(sigma=(f-1)/2, zero extension) -> pad 12 zeros -> cubic B-spline
// prefilter (IIR, pole sqrt(3)-2, mirror initialisation on the padded line) -> B-spline
// evaluation at (o+1/2) f - 1/2  is LINEAR, so it is one banded matrix A (n_out x n_in).  The
// host builds A exactly in fp64 by pushing unit impulses through that chain (build_resample_1d),
// keeps the band above 1e-9 of the largest weight (about 30 taps for nu = 1/2) and the device
// applies  out = clip(A_y * in * A_x^T)  as two FIR passes with per-output-row weights.
// The channel axis is prefiltered by scipy too but evaluated at integer positions, which returns
// the samples unchanged, so it is skipped.  Bound: HBM (read level s, write level s+1).
#include <math.h>
#include <vector>
#include <algorithm>
#include "ica_pyramid.cuh"

namespace ica {

// round-half-to-even like np.round (src/zoom.py:20-21 and skimage's output shape)
int round_half_even(double v) { return (int)nearbyint(v); }

int zoomed_size(int n, double factor) { return round_half_even((double)n * factor); }

void build_resample_1d(int n_in, int n_out, Resample1D* out, double rel_threshold) {
  const double f = (double)n_in / (double)n_out;
  const double sigma = f > 1.0 ? (f - 1.0) / 2.0 : 0.0;
  // scipy.ndimage._gaussian_kernel1d, truncate = 4
  std::vector<double> gk;
  int radius = 0;
  if (sigma > 1e-15) {
    radius = (int)(4.0 * sigma + 0.5);
    gk.resize(2 * radius + 1);
    double sum = 0.0;
    for (int i = -radius; i <= radius; ++i) { gk[i + radius] = exp(-0.5 / (sigma * sigma) * i * i); sum += gk[i + radius]; }
    for (auto& w : gk) w /= sum;
  }
  const int N = n_in + 2 * kSplinePad;
  const double z = sqrt(3.0) - 2.0;
  // evaluation taps and weights per output sample
  std::vector<int> efirst(n_out);
  std::vector<double> ew(4 * (size_t)n_out);
  for (int o = 0; o < n_out; ++o) {
    const double cc = (o + 0.5) * f - 0.5 + kSplinePad;
    const double fl = floor(cc);
    const double t = cc - fl;
    efirst[o] = (int)fl - 1;
    ew[4 * o + 0] = (1 - t) * (1 - t) * (1 - t) / 6.0;
    ew[4 * o + 1] = (3 * t * t * t - 6 * t * t + 4) / 6.0;
    ew[4 * o + 2] = (-3 * t * t * t + 3 * t * t + 3 * t + 1) / 6.0;
    ew[4 * o + 3] = t * t * t / 6.0;
  }
  // banded storage around the centre of each output row
  // half-width kept around each output's centre: the Gaussian radius plus the prefilter tails (|z|^k < 1e-9 at k = 16)
  const int RB = std::max(48, radius + 24);
  const int BWD = 2 * RB + 1;
  std::vector<double> band((size_t)n_out * BWD, 0.0);
  std::vector<int> bstart(n_out);
  for (int o = 0; o < n_out; ++o) bstart[o] = (int)floor((o + 0.5) * f - 0.5) - RB;
  std::vector<double> c(N);
  for (int j = 0; j < n_in; ++j) {
    std::fill(c.begin(), c.end(), 0.0);
    // Gaussian response of the impulse at j (correlate, zero extension, output domain [0,n_in))
    if (radius > 0) {
      for (int i = std::max(0, j - radius); i <= std::min(n_in - 1, j + radius); ++i) c[kSplinePad + i] = gk[j - i + radius];
    } else {
      c[kSplinePad + j] = 1.0;
    }
    // scipy ni_splines.c apply_filter (order 3: one pole, gain (1-z)(1-1/z) = 6), mirror init
    for (int i = 0; i < N; ++i) c[i] *= 6.0;
    {
      double zi = z;
      const double zn1 = pow(z, (double)(N - 1));
      double c0 = c[0] + zn1 * c[N - 1];
      for (int i = 1; i < N - 1; ++i) { c0 += zi * (c[i] + zn1 * c[N - 1 - i]); zi *= z; }
      c[0] = c0 / (1.0 - zn1 * zn1);
    }
    for (int i = 1; i < N; ++i) c[i] += z * c[i - 1];
    c[N - 1] = (z * c[N - 2] + c[N - 1]) * z / (z * z - 1.0);
    for (int i = N - 2; i >= 0; --i) c[i] = z * (c[i + 1] - c[i]);
    // evaluate the outputs whose band can contain j
    const int olo = std::max(0, (int)floor((j - RB - 2) / f) - 1);
    const int ohi = std::min(n_out - 1, (int)ceil((j + RB + 2) / f) + 1);
    for (int o = olo; o <= ohi; ++o) {
      const int col = j - bstart[o];
      if (col < 0 || col >= BWD) continue;
      double s = 0.0;
      for (int k = 0; k < 4; ++k) {
        const int idx = efirst[o] + k;
        if (idx >= 0 && idx < N) s += ew[4 * o + k] * c[idx];
      }
      band[(size_t)o * BWD + col] = s;
    }
  }
  // keep the significant band
  double mx = 0.0;
  for (double v : band) mx = std::max(mx, fabs(v));
  const double thr = rel_threshold * mx;
  std::vector<int> first(n_out), last(n_out);
  int taps = 1;
  for (int o = 0; o < n_out; ++o) {
    int fi = BWD, la = -1;
    for (int k = 0; k < BWD; ++k) if (fabs(band[(size_t)o * BWD + k]) > thr) { fi = std::min(fi, k); la = std::max(la, k); }
    if (la < 0) { fi = RB; la = RB; }
    first[o] = bstart[o] + fi; last[o] = bstart[o] + la;
    taps = std::max(taps, la - fi + 1);
  }
  taps = std::min(taps, n_in);
  out->n_in = n_in; out->n_out = n_out; out->taps = taps;
  out->start.resize(n_out);
  out->weights.assign((size_t)n_out * taps, 0.f);
  for (int o = 0; o < n_out; ++o) {
    int st = first[o] - (taps - (last[o] - first[o] + 1)) / 2;   // centre the kept band
    st = std::max(0, std::min(st, n_in - taps));
    out->start[o] = st;
    for (int k = 0; k < taps; ++k) {
      const int col = st + k - bstart[o];
      out->weights[(size_t)o * taps + k] = (col >= 0 && col < BWD) ? (float)band[(size_t)o * BWD + col] : 0.f;
    }
  }
}

// The IPOL-style level of the reference's zoom.zoom_out (src/zoom.py:29-60; dead code there, PARITY UNPINNED) along one
// axis, as a banded operator built exactly in fp64 like build_resample_1d:
//   scipy.ndimage.gaussian_filter(sigma = sigma_zero * sqrt(1/factor^2 - 1)): defaults mode='reflect', truncate 4
//   scipy.ndimage.map_coordinates(order=3, mode='nearest') at o / factor: pad 12 samples with the edge value,
//   cubic-B-spline prefilter with scipy's 'reflect' initialisation (ni_splines.c, the one it uses for mode nearest),
//   B-spline evaluation at floor(x)-1 .. floor(x)+2, indices clamped to the padded line.
void build_zoom_out_1d(int n_in, double factor, double sigma_zero, Resample1D* out, double rel_threshold) {
  const int n_out = std::max(round_half_even((double)n_in * factor), 1);
  const double sigma = sigma_zero * sqrt(std::max(0.0, 1.0 / (factor * factor) - 1.0));
  std::vector<double> gk;
  int radius = 0;
  if (sigma > 1e-15) {
    radius = (int)(4.0 * sigma + 0.5);
    gk.resize(2 * radius + 1);
    double sum = 0.0;
    for (int i = -radius; i <= radius; ++i) { gk[i + radius] = exp(-0.5 / (sigma * sigma) * i * i); sum += gk[i + radius]; }
    for (auto& w : gk) w /= sum;
  }
  const int N = n_in + 2 * kSplinePad;
  const double z = sqrt(3.0) - 2.0;
  auto reflect = [n_in](int i) {   // scipy 'reflect': (d c b a | a b c d | d c b a)
    const int period = 2 * n_in;
    i %= period; if (i < 0) i += period;
    return i < n_in ? i : period - 1 - i;
  };
  // banded storage around each output's centre (Gaussian radius + prefilter tails, |z|^k < 1e-9 at k = 16)
  const int RB = radius + 24;
  const int BWD = 2 * RB + 1;
  std::vector<double> band((size_t)n_out * BWD, 0.0);
  std::vector<int> bstart(n_out);
  for (int o = 0; o < n_out; ++o) bstart[o] = (int)floor((double)o / factor) - RB;
  std::vector<double> g(n_in), c(N);
  for (int j = 0; j < n_in; ++j) {
    // Gaussian response of the impulse at j with reflect extension: g[i] = sum_k gk[k] * [reflect(i + k) == j]
    std::fill(g.begin(), g.end(), 0.0);
    if (radius > 0) {
      for (int i = std::max(0, j - radius); i <= std::min(n_in - 1, j + radius); ++i) {
        double v = 0.0;
        for (int k = -radius; k <= radius; ++k) if (reflect(i + k) == j) v += gk[k + radius];
        g[i] = v;
      }
    } else {
      g[j] = 1.0;
    }
    for (int i = 0; i < N; ++i) c[i] = 6.0 * g[std::min(std::max(i - kSplinePad, 0), n_in - 1)];   // edge padding, gain
    {  // _init_causal_reflect
      const double c0 = c[0];
      const double zn = pow(z, (double)N);
      double zi = z;
      double acc = c[0] + zn * c[N - 1];
      for (int i = 1; i < N; ++i) { acc += zi * (c[i] + zn * c[N - 1 - i]); zi *= z; }
      c[0] = acc * z / (1.0 - zn * zn) + c0;
    }
    for (int i = 1; i < N; ++i) c[i] += z * c[i - 1];
    c[N - 1] *= z / (z - 1.0);   // _init_anticausal_reflect
    for (int i = N - 2; i >= 0; --i) c[i] = z * (c[i + 1] - c[i]);
    const int olo = std::max(0, (int)floor((j - RB - 2) * factor) - 1);
    const int ohi = std::min(n_out - 1, (int)ceil((j + RB + 2) * factor) + 1);
    for (int o = olo; o <= ohi; ++o) {
      const int col = j - bstart[o];
      if (col < 0 || col >= BWD) continue;
      const double x = (double)o / factor + kSplinePad;
      const double fl = floor(x);
      const double t = x - fl;
      const double w[4] = {(1 - t) * (1 - t) * (1 - t) / 6.0, (3 * t * t * t - 6 * t * t + 4) / 6.0,
                           (-3 * t * t * t + 3 * t * t + 3 * t + 1) / 6.0, t * t * t / 6.0};
      double sacc = 0.0;
      for (int k = 0; k < 4; ++k) sacc += w[k] * c[std::min(std::max((int)fl - 1 + k, 0), N - 1)];
      band[(size_t)o * BWD + col] = sacc;
    }
  }
  double mx = 0.0;
  for (double v : band) mx = std::max(mx, fabs(v));
  const double thr = rel_threshold * mx;
  std::vector<int> first(n_out), last(n_out);
  int taps = 1;
  for (int o = 0; o < n_out; ++o) {
    int fi = BWD, la = -1;
    for (int k = 0; k < BWD; ++k) if (fabs(band[(size_t)o * BWD + k]) > thr) { fi = std::min(fi, k); la = std::max(la, k); }
    if (la < 0) { fi = RB; la = RB; }
    first[o] = bstart[o] + fi; last[o] = bstart[o] + la;
    taps = std::max(taps, la - fi + 1);
  }
  taps = std::min(taps, n_in);
  out->n_in = n_in; out->n_out = n_out; out->taps = taps;
  out->start.resize(n_out);
  out->weights.assign((size_t)n_out * taps, 0.f);
  for (int o = 0; o < n_out; ++o) {
    int st = first[o] - (taps - (last[o] - first[o] + 1)) / 2;
    st = std::max(0, std::min(st, n_in - taps));
    out->start[o] = st;
    for (int k = 0; k < taps; ++k) {
      const int col = st + k - bstart[o];
      out->weights[(size_t)o * taps + k] = (col >= 0 && col < BWD) ? (float)band[(size_t)o * BWD + col] : 0.f;
    }
  }
}

// Rows of a banded operator whose weights are identical up to a shift of 2 samples per output
// (exact 2:1 levels away from the borders): start[o] = 2*o + s0 and the same `taps` weights.
void detect_uniform_rows(const Resample1D& r, FastRows* f) {
  f->lo = f->hi = 0; f->s0 = 0; f->taps = 0;
  if (r.n_out < 16 || r.taps > kFastTapsMax) return;
  const int mid = r.n_out / 2;
  const int s0 = r.start[mid] - 2 * mid;
  const float* wm = &r.weights[(size_t)mid * r.taps];
  float mx = 0.f;
  for (int k = 0; k < r.taps; ++k) mx = std::max(mx, fabsf(wm[k]));
  auto same = [&](int o) {
    if (r.start[o] != 2 * o + s0) return false;
    const float* w = &r.weights[(size_t)o * r.taps];
    for (int k = 0; k < r.taps; ++k) if (fabsf(w[k] - wm[k]) > 2e-8f * mx) return false;
    return true;
  };
  int lo = mid, hi = mid + 1;
  while (lo > 0 && same(lo - 1)) --lo;
  while (hi < r.n_out && same(hi)) ++hi;
  if (hi - lo < 8) return;
  f->lo = lo; f->hi = hi; f->s0 = s0; f->taps = r.taps;
  for (int k = 0; k < kFastTapsMax; ++k) f->w[k] = k < r.taps ? wm[k] : 0.f;
}

namespace {

constexpr int kVR = 4;        // output rows per thread in the vertical pass
constexpr int kMaxTaps = 128;
constexpr int kFR = 8;        // outputs per thread along the filtered axis in the uniform (fast) horizontal kernel
#ifndef ICA_KFRV
#define ICA_KFRV 8
#endif
constexpr int kFRV = ICA_KFRV;       // output rows per thread in the uniform vertical kernel (each input row is read (TP+14)/16 times)

struct FastW { float w[kFastTapsMax]; };
// the same weights paired for packed fp32 arithmetic on two neighbouring outputs of a 2:1 level: wp[j] = (w[j], w[j-2]),
// zero outside [0, taps) -- an input sample j contributes to output i with weight w[j - 2 i]
struct FastW2 { float2 wp[kFastTapsMax + 2]; };

// logical index of the general kernels -> output index, skipping the range [lo, hi) the fast kernel covers
__device__ __forceinline__ int skip_range(int i, int lo, int hi) { return i < lo ? i : i + (hi - lo); }

// ---- vertical pass, general rows: tmp[oy][j] = sum_k Wy[oy][k] * in[sy[oy]+k][j],  j over nx*C floats
__global__ void __launch_bounds__(256) pyr_vertical_kernel(
    const float* __restrict__ in0a, const float* __restrict__ in0b, int nset, long long in_stride, int in_pitch,
    int ncols, int ny_out, int skip_lo, int skip_hi,
    const float* __restrict__ W, const int* __restrict__ start, int taps,
    float* __restrict__ tmp, long long tmp_stride) {
  __shared__ float sw[kVR][kMaxTaps];
  __shared__ int sst[kVR], soy[kVR];
  const int img = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  // row groups never straddle the range [skip_lo, skip_hi) the fast kernel covers
  const int gA = (skip_lo + kVR - 1) / kVR;
  if (threadIdx.x < kVR) {
    int oy;
    if ((int)blockIdx.y < gA) { oy = blockIdx.y * kVR + threadIdx.x; if (oy >= skip_lo) oy = -1; }
    else { oy = skip_hi + ((int)blockIdx.y - gA) * kVR + threadIdx.x; if (oy >= ny_out) oy = -1; }
    soy[threadIdx.x] = oy;
  }
  __syncthreads();
  if (threadIdx.x < kVR) {
    int oy = soy[threadIdx.x];
    if (oy < 0) oy = soy[0];     // inactive slots reuse the group's first row (always valid) for the loop bounds
    sst[threadIdx.x] = start[oy];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kVR * taps; i += blockDim.x) {
    const int r = i / taps, k = i % taps;
    sw[r][k] = soy[r] >= 0 ? W[(long long)soy[r] * taps + k] : 0.f;
  }
  __syncthreads();
  if (j >= ncols) return;
  int row_lo = sst[0], row_hi = sst[0] + taps;
#pragma unroll
  for (int r = 1; r < kVR; ++r) { row_lo = min(row_lo, sst[r]); row_hi = max(row_hi, sst[r] + taps); }
  const float* in = (img < nset ? in0a + (long long)img * in_stride : in0b + (long long)(img - nset) * in_stride) + j;
  float acc[kVR];
#pragma unroll
  for (int r = 0; r < kVR; ++r) acc[r] = 0.f;
  // batches of 8 independent loads in flight: the border rows are few, so this kernel is pure load latency
  for (int row0 = row_lo; row0 < row_hi; row0 += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = row0 + u < row_hi ? __ldg(in + (long long)(row0 + u) * in_pitch) : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int r = 0; r < kVR; ++r) {
        const int k = row0 + u - sst[r];
        if (k >= 0 && k < taps && row0 + u < row_hi) acc[r] = fmaf(sw[r][k], v[u], acc[r]);
      }
    }
  }
  float* o = tmp + (long long)img * tmp_stride + j;
#pragma unroll
  for (int r = 0; r < kVR; ++r) if (soy[r] >= 0) o[(long long)soy[r] * ncols] = acc[r];
}

// ---- vertical pass, uniform rows (exact 2:1, interior): 4 float columns x kFR output rows per
// thread, every input row loaded once per thread (LDG.128), weights as constant-bank operands.
template <int TP>
__global__ void __launch_bounds__(128) pyr_vertical_fast_kernel(
    const float* __restrict__ in0a, const float* __restrict__ in0b, int nset, long long in_stride, int in_pitch,
    int col4_off, int ncols4, int oy_lo, int ngroups, int s0,
    const FastW fw, float* __restrict__ tmp, long long tmp_stride, int tmp_pitch) {
  const int img = blockIdx.z;
  const int j4l = blockIdx.x * blockDim.x + threadIdx.x;   // float4 column inside the strip [col4_off, col4_off + ncols4)
  if (j4l >= ncols4 || (int)blockIdx.y >= ngroups) return;
  const int j4 = col4_off + j4l;
  const int oy = oy_lo + blockIdx.y * kFRV;
  const float* inb = img < nset ? in0a + (long long)img * in_stride : in0b + (long long)(img - nset) * in_stride;
  const float4* in = reinterpret_cast<const float4*>(inb + (long long)(2 * oy + s0) * in_pitch) + j4;
  const int p4 = in_pitch >> 2;
  float4 acc[kFRV];
#pragma unroll
  for (int r = 0; r < kFRV; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < TP + 2 * (kFRV - 1); ++j) {
    const float4 v = __ldg(in + (long long)j * p4);
#pragma unroll
    for (int r = 0; r < kFRV; ++r) {
      const int k = j - 2 * r;
      if (k >= 0 && k < TP) {
        const float w = fw.w[k];
        acc[r].x = fmaf(w, v.x, acc[r].x); acc[r].y = fmaf(w, v.y, acc[r].y);
        acc[r].z = fmaf(w, v.z, acc[r].z); acc[r].w = fmaf(w, v.w, acc[r].w);
      }
    }
  }
  float4* o = reinterpret_cast<float4*>(tmp + (long long)img * tmp_stride + (long long)oy * tmp_pitch) + j4;
  const int t4 = tmp_pitch >> 2;
#pragma unroll
  for (int r = 0; r < kFRV; ++r) o[(long long)r * t4] = acc[r];
}

// ---- horizontal pass, uniform columns: kFR consecutive outputs per thread, inputs loaded once
template <int C, int TP>
__global__ void __launch_bounds__(128) pyr_horizontal_fast_kernel(
    const float* __restrict__ tmp, long long tmp_stride, int ncols_in, int rs_lo, int rs_hi, int ox_lo, int ngroups, int s0, const FastW fw,
    float* __restrict__ out0a, float* __restrict__ out0b, int nset, long long out_stride, int out_pitch,
    const MinMaxKeys* __restrict__ mm_parent, MinMaxKeys* __restrict__ mm_child, int mm_stride) {
  const int img = blockIdx.z;
  const int oy = skip_range(blockIdx.y, rs_lo, rs_hi);   // all rows but [rs_lo, rs_hi) (those belong to the fused kernel)
  const int g = blockIdx.x * blockDim.x + threadIdx.x;   // (group of kFR columns) * C + channel
  const int set = img < nset ? 0 : 1, pb = img - set * nset;
  float* out0 = (set ? out0b : out0a) + (long long)pb * out_stride;
  const MinMaxKeys pk = mm_parent[(long long)pb * mm_stride + set];
  const float lo = key_float(pk.lo), hi = key_float(pk.hi);
  const bool active = g < ngroups * C;
  float vmin = 0.f, vmax = 0.f;
  if (active) {
    const int og = g / C, c = g - og * C;
    const int ox = ox_lo + og * kFR;
    const float* row = tmp + (long long)img * tmp_stride + (long long)oy * ncols_in + (long long)(2 * ox + s0) * C + c;
    float acc[kFR];
#pragma unroll
    for (int r = 0; r < kFR; ++r) acc[r] = 0.f;
#pragma unroll
    for (int j = 0; j < TP + 2 * (kFR - 1); ++j) {
      const float v = __ldg(row + j * C);
#pragma unroll
      for (int r = 0; r < kFR; ++r) {
        const int k = j - 2 * r;
        if (k >= 0 && k < TP) acc[r] = fmaf(fw.w[k], v, acc[r]);
      }
    }
    float* o = out0 + (long long)oy * out_pitch + ox * C + c;
    vmin = 3.4e38f; vmax = -3.4e38f;
#pragma unroll
    for (int r = 0; r < kFR; ++r) {
      const float val = fminf(fmaxf(acc[r], lo), hi);
      o[r * C] = val;
      vmin = fminf(vmin, val); vmax = fmaxf(vmax, val);
    }
  }
  // two-sided reduction: feed min and max through the same helper
  {
    unsigned kmin = active ? float_key(vmin) : 0xffffffffu;
    unsigned kmax = active ? float_key(vmax) : 0u;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
      kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    __shared__ unsigned smin[4], smax[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { smin[warp] = kmin; smax[warp] = kmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { kmin = min(kmin, smin[w]); kmax = max(kmax, smax[w]); }
      MinMaxKeys* ck = mm_child + (long long)pb * mm_stride + set;
      if (kmin <= kmax) { atomicMin(&ck->lo, kmin); atomicMax(&ck->hi, kmax); }
    }
  }
}

// ---- border rows, 16-byte aligned images: 4 float columns x kVR output rows per thread.  The rows of a group read a
// common range of input rows; their weights are laid out zero-padded over that range ([input row][output row] in
// shared memory), so the inner loop is one LDG.128 + one LDS.128 + 16 FMA with no predicates.
constexpr int kVSpan = kMaxTaps + 16;
__global__ void __launch_bounds__(128) pyr_vertical_border4_kernel(
    const float* __restrict__ in0a, const float* __restrict__ in0b, int nset, long long in_stride, int in_pitch,
    int ncols4, int ny_out, int skip_lo, int skip_hi,
    const float* __restrict__ W, const int* __restrict__ start, int taps,
    float* __restrict__ tmp, long long tmp_stride, int tmp_pitch) {
  __shared__ float4 sw[kVSpan];   // sw[i] = weights of the group's (up to) 4 output rows for input row row_lo + i
  __shared__ int sst[kVR], soy[kVR], srange[2];
  const int img = blockIdx.z;
  const int gA = (skip_lo + kVR - 1) / kVR;   // row groups never straddle [skip_lo, skip_hi)
  if (threadIdx.x < kVR) {
    int oy;
    if ((int)blockIdx.y < gA) { oy = blockIdx.y * kVR + threadIdx.x; if (oy >= skip_lo) oy = -1; }
    else { oy = skip_hi + ((int)blockIdx.y - gA) * kVR + threadIdx.x; if (oy >= ny_out) oy = -1; }
    soy[threadIdx.x] = oy;
    sst[threadIdx.x] = oy >= 0 ? start[oy] : -1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int lo = 0x7fffffff, hi = 0;
    for (int r = 0; r < kVR; ++r) if (soy[r] >= 0) { lo = min(lo, sst[r]); hi = max(hi, sst[r] + taps); }
    srange[0] = lo; srange[1] = min(hi, lo + kVSpan);   // (spans beyond kVSpan do not occur: starts of neighbours differ by <= 2)
  }
  __syncthreads();
  const int row_lo = srange[0], n = srange[1] - srange[0];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float w[kVR];
#pragma unroll
    for (int r = 0; r < kVR; ++r) {
      const int k = row_lo + i - sst[r];
      w[r] = (soy[r] >= 0 && k >= 0 && k < taps) ? W[(long long)soy[r] * taps + k] : 0.f;
    }
    sw[i] = make_float4(w[0], w[1], w[2], w[3]);
  }
  __syncthreads();
  const int j4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (j4 >= ncols4) return;
  const float* inb = img < nset ? in0a + (long long)img * in_stride : in0b + (long long)(img - nset) * in_stride;
  const float4* in = reinterpret_cast<const float4*>(inb + (long long)row_lo * in_pitch) + j4;
  const int p4 = in_pitch >> 2;
  float4 acc[kVR];
#pragma unroll
  for (int r = 0; r < kVR; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int i = 0; i < n; ++i) {
    const float4 v = __ldg(in + (long long)i * p4);
    const float4 w = sw[i];
    acc[0].x = fmaf(w.x, v.x, acc[0].x); acc[0].y = fmaf(w.x, v.y, acc[0].y); acc[0].z = fmaf(w.x, v.z, acc[0].z); acc[0].w = fmaf(w.x, v.w, acc[0].w);
    acc[1].x = fmaf(w.y, v.x, acc[1].x); acc[1].y = fmaf(w.y, v.y, acc[1].y); acc[1].z = fmaf(w.y, v.z, acc[1].z); acc[1].w = fmaf(w.y, v.w, acc[1].w);
    acc[2].x = fmaf(w.z, v.x, acc[2].x); acc[2].y = fmaf(w.z, v.y, acc[2].y); acc[2].z = fmaf(w.z, v.z, acc[2].z); acc[2].w = fmaf(w.z, v.w, acc[2].w);
    acc[3].x = fmaf(w.w, v.x, acc[3].x); acc[3].y = fmaf(w.w, v.y, acc[3].y); acc[3].z = fmaf(w.w, v.z, acc[3].z); acc[3].w = fmaf(w.w, v.w, acc[3].w);
  }
  float4* o = reinterpret_cast<float4*>(tmp + (long long)img * tmp_stride) + j4;
  const int t4 = tmp_pitch >> 2;
#pragma unroll
  for (int r = 0; r < kVR; ++r) if (soy[r] >= 0) o[(long long)soy[r] * t4] = acc[r];
}

// ---- border columns: kHR output rows per thread (the weight of a tap is loaded once and used for all of them)
constexpr int kHR = 8;
template <int C>
__global__ void __launch_bounds__(128) pyr_horizontal_border_kernel(
    const float* __restrict__ tmp, long long tmp_stride, int ncols_in, int nx_out, int r0, int nrows, int rs_lo, int rs_hi,
    int skip_lo, int skip_hi, const float* __restrict__ Wt /* [taps][nx_out] */, const int* __restrict__ start, int taps,
    float* __restrict__ out0a, float* __restrict__ out0b, int nset, long long out_stride, int out_pitch,
    const MinMaxKeys* __restrict__ mm_parent, MinMaxKeys* __restrict__ mm_child, int mm_stride) {
  const int img = blockIdx.z;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;   // logical flat index over the border columns
  const int ncol = nx_out - (skip_hi - skip_lo);
  const int set = img < nset ? 0 : 1, pb = img - set * nset;
  float* out0 = (set ? out0b : out0a) + (long long)pb * out_stride;
  const MinMaxKeys pk = mm_parent[(long long)pb * mm_stride + set];
  const float lo = key_float(pk.lo), hi = key_float(pk.hi);
  const bool active = e < ncol * C;
  float vmin = 3.4e38f, vmax = -3.4e38f;
  if (active) {
    const int oxl = e / C, c = e - oxl * C;
    const int ox = skip_range(oxl, skip_lo, skip_hi);
    const int st = __ldg(start + ox);
    const float* rows[kHR];
    int oys[kHR];
#pragma unroll
    for (int i = 0; i < kHR; ++i) {
      const int rl = blockIdx.y * kHR + i;                       // row of this launch
      oys[i] = rl < nrows ? skip_range(r0 + rl, rs_lo, rs_hi) : -1;
      rows[i] = tmp + (long long)img * tmp_stride + (long long)(oys[i] >= 0 ? oys[i] : oys[0]) * ncols_in + st * C + c;
    }
    float acc[kHR];
#pragma unroll
    for (int i = 0; i < kHR; ++i) acc[i] = 0.f;
    const float* wp = Wt + ox;
#pragma unroll 2
    for (int k = 0; k < taps; ++k) {
      const float w = __ldg(wp + (long long)k * nx_out);
#pragma unroll
      for (int i = 0; i < kHR; ++i) acc[i] = fmaf(w, __ldg(rows[i] + k * C), acc[i]);
    }
#pragma unroll
    for (int i = 0; i < kHR; ++i) {
      if (oys[i] < 0) continue;
      const float val = fminf(fmaxf(acc[i], lo), hi);
      out0[(long long)oys[i] * out_pitch + ox * C + c] = val;
      vmin = fminf(vmin, val); vmax = fmaxf(vmax, val);
    }
  }
  {
    unsigned kmin = vmin <= vmax ? float_key(vmin) : 0xffffffffu;
    unsigned kmax = vmin <= vmax ? float_key(vmax) : 0u;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
      kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    __shared__ unsigned smin[4], smax[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { smin[warp] = kmin; smax[warp] = kmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int)((blockDim.x + 31) >> 5); ++w) { kmin = min(kmin, smin[w]); kmax = max(kmax, smax[w]); }
      MinMaxKeys* ck = mm_child + (long long)pb * mm_stride + set;
      if (kmin <= kmax) { atomicMin(&ck->lo, kmin); atomicMax(&ck->hi, kmax); }
    }
  }
}

// ---- both passes in one kernel for the region where rows AND columns are uniform (exact 2:1, interior): a block
// produces kFRV output rows x OWB output pixels.  Stage 1 = the vertical pass on the input window of the strip
// (one float4 column and kFRV output rows per thread, as in pyr_vertical_fast_kernel) into shared memory; stage 2 =
// the horizontal pass from shared memory (kFH consecutive outputs of one channel per task), clip, min/max, store.
// The intermediate image never goes to global memory.  Shared rows are padded (one float every 2*kFH*C) so that the
// strided reads of stage 2 spread over the banks.
constexpr int kFH = 4;
template <int C> struct FusedCfg {
  static constexpr int OWB = C == 3 ? 64 : 192;            // output pixels per strip (192 floats per output row)
  static constexpr int PADN = 2 * kFH * C;                 // floats between pads
  static constexpr int TASKS_PER_ROW = (OWB / kFH) * C;    // 48
};
// GATHER (level 0 -> 1 when the caller has not computed level 0's min / max): the block also takes the min / max of the
// input rows and float4 columns it OWNS (consecutive blocks' windows overlap; the owned parts tile the union exactly once)
// into mm_gather -- the separate pass over level 0 disappears -- and, the parent's range being unknown until every block
// is done, does not clip: the caller clamps the child level afterwards in the (rare) case that it leaves the range.
template <int C, int TP, bool GATHER>
__global__ void __launch_bounds__(128) pyr_fused_fast_kernel(
    const float* __restrict__ in0a, const float* __restrict__ in0b, int nset, long long in_stride, int in_pitch, int ncols,
    int oy_lo, int s0y, const FastW fwy, int ox_lo, int ox_hi, int s0x, const FastW2 fwx,
    float* __restrict__ out0a, float* __restrict__ out0b, long long out_stride, int out_pitch,
    const MinMaxKeys* __restrict__ mm_parent, MinMaxKeys* __restrict__ mm_child, int mm_stride,
    MinMaxKeys* __restrict__ mm_gather) {
  constexpr int OWB = FusedCfg<C>::OWB, PADN = FusedCfg<C>::PADN, TPR = FusedCfg<C>::TASKS_PER_ROW;
  constexpr int WF = ((2 * (OWB - 1) + TP) * C + 3 + 3) / 4 * 4;   // window floats incl. alignment slack, float4 multiple
  constexpr int NW4 = WF / 4;
  constexpr int SP = WF + WF / PADN + 4;                           // padded shared row
  static_assert(NW4 <= 128, "one float4 column per thread");
  __shared__ float stmp[kFRV][SP];
  const int img = blockIdx.z;
  const int set = img < nset ? 0 : 1, pb = img - set * nset;
  const int ox0 = ox_lo + blockIdx.x * OWB;
  const int oy = oy_lo + blockIdx.y * kFRV;
  const int fstart = (2 * ox0 + s0x) * C;          // first float of the strip's footprint in an input row
  const int f0 = fstart & ~3;                      // 16-byte aligned window start
  const int t = threadIdx.x;
  // Shared position of the window float g (counted from the strip's first footprint float `fstart`): g + g / PADN.
  // A horizontal task reads g = PADN * og + c + C * j, whose pad count is og + j / (2 * kFH) -- a compile-time offset.
  const int shift = fstart - f0;
  // ---- stage 1: vertical pass (packed fp32: the float4 column is two FFMA2 lanes)
  {
    const float* inb = (set ? in0b : in0a) + (long long)pb * in_stride;
    const bool act = t < NW4 && f0 + 4 * t < ncols;
    constexpr int kOwnJ0 = (TP - 2) / 2, kOwnT0 = (NW4 - 2 * OWB * C / 4) / 2;     // owned rows / float4 columns of the window
    static_assert(kOwnT0 >= 0, "a block's window is at least as wide as the stride between blocks");
    const bool towned = GATHER && t >= kOwnT0 && t < kOwnT0 + 2 * OWB * C / 4;
    float gmin = 3.4e38f, gmax = -3.4e38f;       // NaNs are skipped by fminf / fmaxf (nanmin / nanmax)
    float2 acc[kFRV][2];
#pragma unroll
    for (int r = 0; r < kFRV; ++r) { acc[r][0] = make_float2(0.f, 0.f); acc[r][1] = make_float2(0.f, 0.f); }
    if (act) {
      const float4* in = reinterpret_cast<const float4*>(inb + (long long)(2 * oy + s0y) * in_pitch + f0) + t;
      const int p4 = in_pitch >> 2;
#pragma unroll
      for (int j = 0; j < TP + 2 * (kFRV - 1); ++j) {
        const float4 v = __ldg(in + (long long)j * p4);
        const float2 vlo2 = make_float2(v.x, v.y), vhi2 = make_float2(v.z, v.w);
        if (GATHER && j >= kOwnJ0 && j < kOwnJ0 + 2 * kFRV && towned) {
          gmin = fminf(fminf(gmin, fminf(v.x, v.y)), fminf(v.z, v.w));
          gmax = fmaxf(fmaxf(gmax, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
        }
#pragma unroll
        for (int r = 0; r < kFRV; ++r) {
          const int k = j - 2 * r;
          if (k >= 0 && k < TP) {
            const float2 w2 = make_float2(fwy.w[k], fwy.w[k]);
            acc[r][0] = __ffma2_rn(w2, vlo2, acc[r][0]);
            acc[r][1] = __ffma2_rn(w2, vhi2, acc[r][1]);
          }
        }
      }
    }
    if (GATHER) {      // block min / max of the owned inputs -> the parent level's keys
      unsigned kmin = gmin <= gmax ? float_key(gmin) : 0xffffffffu, kmax = gmin <= gmax ? float_key(gmax) : 0u;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
      }
      if ((t & 31) == 0 && kmin <= kmax) {
        MinMaxKeys* gk = mm_gather + (long long)pb * mm_stride + set;
        atomicMin(&gk->lo, kmin); atomicMax(&gk->hi, kmax);
      }
    }
    if (t < NW4) {
      int pos[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { const int g = 4 * t + i - shift; pos[i] = g >= 0 ? g + g / PADN : -1; }
#pragma unroll
      for (int r = 0; r < kFRV; ++r) {
        if (pos[0] >= 0) stmp[r][pos[0]] = acc[r][0].x;
        if (pos[1] >= 0) stmp[r][pos[1]] = acc[r][0].y;
        if (pos[2] >= 0) stmp[r][pos[2]] = acc[r][1].x;
        if (pos[3] >= 0) stmp[r][pos[3]] = acc[r][1].y;
      }
    }
  }
  __syncthreads();
  // ---- stage 2: horizontal pass from shared memory
  float* out0 = (set ? out0b : out0a) + (long long)pb * out_stride;
  float lo = -3.4028235e38f, hi = 3.4028235e38f;
  if (!GATHER) {
    const MinMaxKeys pk = mm_parent[(long long)pb * mm_stride + set];
    lo = key_float(pk.lo); hi = key_float(pk.hi);
  }
  float vmin = 3.4e38f, vmax = -3.4e38f;
#pragma unroll 1
  for (int q = t; q < kFRV * TPR; q += 128) {
    const int r = q / TPR, rem = q - r * TPR;
    const int og = rem / C, c = rem - og * C;
    const int ox = ox0 + og * kFH;
    if (ox >= ox_hi) continue;                       // partial last strip (ox_hi - ox_lo is a multiple of kFH)
    const float* src = &stmp[r][(PADN + 1) * og + c];
    // outputs (0, 1) and (2, 3) as two packed accumulators: per input sample two FFMA2 with the paired weights instead
    // of up to four FFMA (element by element the same fmaf sequence: the terms outside a band carry weight zero)
    static_assert(kFH == 4, "two packed accumulator pairs");
    float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < TP + 2 * (kFH - 1); ++j) {
      const float v = src[C * j + j / (2 * kFH)];
      const float2 v2 = make_float2(v, v);
      if (j < TP + 2) a01 = __ffma2_rn(fwx.wp[j], v2, a01);
      if (j >= 4) a23 = __ffma2_rn(fwx.wp[j - 4], v2, a23);
    }
    const float acc[kFH] = {a01.x, a01.y, a23.x, a23.y};
    float* o = out0 + (long long)(oy + r) * out_pitch + ox * C + c;
#pragma unroll
    for (int i = 0; i < kFH; ++i) {
      const float val = fminf(fmaxf(acc[i], lo), hi);
      o[i * C] = val;
      vmin = fminf(vmin, val); vmax = fmaxf(vmax, val);
    }
  }
  {
    unsigned kmin = vmin <= vmax ? float_key(vmin) : 0xffffffffu;
    unsigned kmax = vmin <= vmax ? float_key(vmax) : 0u;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
      kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    __shared__ unsigned smin[4], smax[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { smin[warp] = kmin; smax[warp] = kmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 4; ++w) { kmin = min(kmin, smin[w]); kmax = max(kmax, smax[w]); }
      MinMaxKeys* ck = mm_child + (long long)pb * mm_stride + set;
      if (kmin <= kmax) { atomicMin(&ck->lo, kmin); atomicMax(&ck->hi, kmax); }
    }
  }
}

// Min / max of the part of a level that the GATHER blocks of the fused kernel do not own: rows outside [r0, r1) in full,
// and the floats outside [f0, f1) of the rows inside.  grid (blocks, images)
__global__ void __launch_bounds__(256) minmax_frame_kernel(const float* __restrict__ in0a, const float* __restrict__ in0b, int nset,
                                                            long long in_stride, int in_pitch, int ncols, int ny, int r0, int r1,
                                                            int f0, int f1, MinMaxKeys* __restrict__ mm, int mm_stride) {
  const int img = blockIdx.y;
  const int set = img < nset ? 0 : 1, pb = img - set * nset;
  const float* in = (set ? in0b : in0a) + (long long)pb * in_stride;
  float fmin_ = 3.4e38f, fmax_ = -3.4e38f;
  const int side = f0 + (ncols - f1);                 // floats per inside row that belong to the frame
  for (int y = blockIdx.x; y < ny; y += gridDim.x) {
    const float* row = in + (long long)y * in_pitch;
    if (y < r0 || y >= r1) {
      for (int i = threadIdx.x; i < ncols; i += blockDim.x) { const float v = __ldg(row + i); fmin_ = fminf(fmin_, v); fmax_ = fmaxf(fmax_, v); }
    } else {
      for (int i = threadIdx.x; i < side; i += blockDim.x) {
        const float v = __ldg(row + (i < f0 ? i : f1 + (i - f0)));
        fmin_ = fminf(fmin_, v); fmax_ = fmaxf(fmax_, v);
      }
    }
  }
  unsigned kmin = 0xffffffffu, kmax = 0u;
  if (fmin_ <= fmax_) { kmin = float_key(fmin_); kmax = float_key(fmax_); }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  __shared__ unsigned smin[8], smax[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { smin[warp] = kmin; smax[warp] = kmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { kmin = min(kmin, smin[w]); kmax = max(kmax, smax[w]); }
    MinMaxKeys* k = mm + (long long)pb * mm_stride + set;
    if (kmin <= kmax) { atomicMin(&k->lo, kmin); atomicMax(&k->hi, kmax); }
  }
}

// Deferred clip of a level that was built without its parent's range (GATHER): decide per image whether any value
// left [parent min, parent max] (skimage's clip), clamp those images, then clamp their keys.
__global__ void clip_decide_kernel(const MinMaxKeys* __restrict__ mm_parent, const MinMaxKeys* __restrict__ mm_child, int mm_stride,
                                   int nimg, int nset, int* __restrict__ flags) {
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= nimg) return;
  const int set = img < nset ? 0 : 1, pb = img - set * nset;
  const MinMaxKeys p = mm_parent[(long long)pb * mm_stride + set], c = mm_child[(long long)pb * mm_stride + set];
  flags[img] = (c.lo <= c.hi && (c.lo < p.lo || c.hi > p.hi)) ? 1 : 0;
}
__global__ void __launch_bounds__(256) clip_apply_kernel(float* __restrict__ out0a, float* __restrict__ out0b, int nset, long long out_stride,
                                                          int out_pitch, int ncols, int ny, const MinMaxKeys* __restrict__ mm_parent,
                                                          MinMaxKeys* __restrict__ mm_child, int mm_stride, const int* __restrict__ flags) {
  const int img = blockIdx.y;
  if (!flags[img]) return;
  const int set = img < nset ? 0 : 1, pb = img - set * nset;
  const MinMaxKeys p = mm_parent[(long long)pb * mm_stride + set];
  const float lo = key_float(p.lo), hi = key_float(p.hi);
  float* out = (set ? out0b : out0a) + (long long)pb * out_stride;
  for (int y = blockIdx.x; y < ny; y += gridDim.x)
    for (int i = threadIdx.x; i < ncols; i += blockDim.x) {
      float* q = out + (long long)y * out_pitch + i;
      *q = fminf(fmaxf(*q, lo), hi);
    }
}
__global__ void clip_keys_kernel(const MinMaxKeys* __restrict__ mm_parent, MinMaxKeys* __restrict__ mm_child, int mm_stride, int nimg,
                                 int nset, const int* __restrict__ flags) {
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= nimg || !flags[img]) return;
  const int set = img < nset ? 0 : 1, pb = img - set * nset;
  const MinMaxKeys p = mm_parent[(long long)pb * mm_stride + set];
  MinMaxKeys* c = mm_child + (long long)pb * mm_stride + set;
  c->lo = max(c->lo, p.lo); c->hi = min(c->hi, p.hi);      // keys are order-preserving: the range of the clamped values
}

__global__ void minmax_reset_kernel(MinMaxKeys* mm, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) { mm[i].lo = 0xffffffffu; mm[i].hi = 0u; }
}

// min/max of a whole image (level 0): grid (blocks, images)
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ img0, long long stride, long long count,
                                                      MinMaxKeys* __restrict__ mm, int mm_stride) {
  const float* img = img0 + (long long)blockIdx.y * stride;
  unsigned kmin = 0xffffffffu, kmax = 0u;
  const bool vec = ((reinterpret_cast<unsigned long long>(img) & 15ull) == 0);
  const long long n4 = vec ? count / 4 : 0;
  float fmin_ = 3.4e38f, fmax_ = -3.4e38f;       // NaNs are skipped by fminf/fmaxf (nanmin / nanmax)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(img) + i);
    fmin_ = fminf(fminf(fmin_, v.x), fminf(v.y, fminf(v.z, v.w)));
    fmax_ = fmaxf(fmaxf(fmax_, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
  }
  for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const float v = __ldg(img + i);
    fmin_ = fminf(fmin_, v); fmax_ = fmaxf(fmax_, v);
  }
  if (fmin_ <= fmax_) { kmin = float_key(fmin_); kmax = float_key(fmax_); }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  __shared__ unsigned smin[8], smax[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { smin[warp] = kmin; smax[warp] = kmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { kmin = min(kmin, smin[w]); kmax = max(kmax, smax[w]); }
    MinMaxKeys* k = mm + (long long)blockIdx.y * mm_stride;
    atomicMin(&k->lo, kmin);
    atomicMax(&k->hi, kmax);
  }
}

// dtype conversion of host-format inputs (uint8 / float64 -> float32) fused with the min/max of the converted image
// (the level-0 clip range): grid (blocks, images); NaNs are skipped like np.nanmin / np.nanmax
template <typename T>
__global__ void __launch_bounds__(256) convert_minmax_kernel(const T* __restrict__ in0, float* __restrict__ out0, long long count,
                                                              MinMaxKeys* __restrict__ mm, int mm_stride) {
  const T* in = in0 + (long long)blockIdx.y * count;
  float* out = out0 + (long long)blockIdx.y * count;
  float fmin_ = 3.4e38f, fmax_ = -3.4e38f;
  const bool vec = sizeof(T) == 1 && (count & 3) == 0 &&
                   (((reinterpret_cast<unsigned long long>(in)) & 3ull) == 0) && (((reinterpret_cast<unsigned long long>(out)) & 15ull) == 0);
  if (vec) {
    const long long n4 = count >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(in) + i);
      const float4 v = make_float4((float)u.x, (float)u.y, (float)u.z, (float)u.w);
      reinterpret_cast<float4*>(out)[i] = v;
      fmin_ = fminf(fminf(fmin_, v.x), fminf(v.y, fminf(v.z, v.w)));
      fmax_ = fmaxf(fmaxf(fmax_, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
      const float v = (float)in[i];
      out[i] = v;
      fmin_ = fminf(fmin_, v); fmax_ = fmaxf(fmax_, v);
    }
  }
  unsigned kmin = 0xffffffffu, kmax = 0u;
  if (fmin_ <= fmax_) { kmin = float_key(fmin_); kmax = float_key(fmax_); }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  __shared__ unsigned smin[8], smax[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { smin[warp] = kmin; smax[warp] = kmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { kmin = min(kmin, smin[w]); kmax = max(kmax, smax[w]); }
    MinMaxKeys* k = mm + (long long)blockIdx.y * mm_stride;
    if (kmin <= kmax) { atomicMin(&k->lo, kmin); atomicMax(&k->hi, kmax); }
  }
}

// RGB -> luminance ingest: host images are [H][W][3] of type T, the plan works on one channel.
// Y = 0.2125 R + 0.7154 G + 0.0721 B (the weights of skimage.color.rgb2gray, the reference's image toolkit), evaluated
// in float64 left to right and rounded once to float32; fused with the min/max of the result like the plain conversions.
template <typename T>
__global__ void __launch_bounds__(256) convert_luma_minmax_kernel(const T* __restrict__ in0, float* __restrict__ out0, long long npix,
                                                                   MinMaxKeys* __restrict__ mm, int mm_stride) {
  const T* in = in0 + (long long)blockIdx.y * npix * 3;
  float* out = out0 + (long long)blockIdx.y * npix;
  float fmin_ = 3.4e38f, fmax_ = -3.4e38f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const double r = (double)in[3 * i], g = (double)in[3 * i + 1], b = (double)in[3 * i + 2];
    const float v = (float)__dadd_rn(__dadd_rn(__dmul_rn(0.2125, r), __dmul_rn(0.7154, g)), __dmul_rn(0.0721, b));
    out[i] = v;
    fmin_ = fminf(fmin_, v); fmax_ = fmaxf(fmax_, v);
  }
  unsigned kmin = 0xffffffffu, kmax = 0u;
  if (fmin_ <= fmax_) { kmin = float_key(fmin_); kmax = float_key(fmax_); }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  __shared__ unsigned smin[8], smax[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { smin[warp] = kmin; smax[warp] = kmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { kmin = min(kmin, smin[w]); kmax = max(kmax, smax[w]); }
    MinMaxKeys* k = mm + (long long)blockIdx.y * mm_stride;
    if (kmin <= kmax) { atomicMin(&k->lo, kmin); atomicMax(&k->hi, kmax); }
  }
}

}  // namespace

__global__ void widen_f32_f64_kernel(const float* __restrict__ in, double* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = (double)in[i];
}
// float32 -> float64 on the device (the reference returns DI / Iw as float64 arrays: the widening happens before the
// device->host copy instead of in a host pass over the result)
cudaError_t launch_widen_f64(const float* in, double* out, long long n, cudaStream_t stream) {
  widen_f32_f64_kernel<<<(int)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, stream>>>(in, out, n);
  return cudaGetLastError();
}

// dtype: 0 = float32, 1 = uint8, 2 = float64 RGB input; out: float32 luminance, npix pixels per image
cudaError_t launch_convert_luma(const void* in, int dtype, float* out, long long npix, int nimg, MinMaxKeys* mm, int mm_stride,
                                cudaStream_t stream) {
  const int blocks = (int)std::max<long long>(1, std::min<long long>((npix + 256 * 8 - 1) / (256 * 8), 256));
  const dim3 grid(blocks, nimg);
  if (dtype == 1) convert_luma_minmax_kernel<unsigned char><<<grid, 256, 0, stream>>>(static_cast<const unsigned char*>(in), out, npix, mm, mm_stride);
  else if (dtype == 2) convert_luma_minmax_kernel<double><<<grid, 256, 0, stream>>>(static_cast<const double*>(in), out, npix, mm, mm_stride);
  else convert_luma_minmax_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(in), out, npix, mm, mm_stride);
  return cudaGetLastError();
}

cudaError_t launch_minmax_reset(MinMaxKeys* mm, int count, cudaStream_t stream) {
  minmax_reset_kernel<<<(count + 255) / 256, 256, 0, stream>>>(mm, count);
  return cudaGetLastError();
}

cudaError_t launch_minmax(const float* img0, long long stride, long long count, int nimg, MinMaxKeys* mm,
                          int mm_stride, cudaStream_t stream) {
  int blocks = (int)std::min<long long>((count + 256 * 16 - 1) / (256 * 16), 256);
  if (blocks < 1) blocks = 1;
  minmax_kernel<<<dim3(blocks, nimg), 256, 0, stream>>>(img0, stride, count, mm, mm_stride);
  return cudaGetLastError();
}

// Groups of kFR outputs of the uniform range whose (zero-padded) TP-tap footprint stays inside [0, n_in):
// the padding taps have weight 0, but 0 * (whatever lies past the data) must never be formed.
static void fast_groups(const FastRows& f, int n_in, int TP, int FR, int* lo, int* ngroups) {
  int l = f.lo;
  while (2 * l + f.s0 < 0) ++l;
  const int omax = (n_in - f.s0 - TP - 2 * (FR - 1)) / 2;   // last admissible group base
  int ng = 0;
  if (f.hi - l >= FR && omax >= l) ng = std::min((f.hi - l) / FR, (omax - l) / FR + 1);
  *lo = l; *ngroups = std::max(ng, 0);
}

// vertical fast pass on the float4 columns [col4_off, col4_off + ncols4) of the rows lo .. lo + ngroups * kFRV
template <int TP>
static void launch_vfast(const float* in0a, const float* in0b, int nset, long long in_stride, int in_pitch, int ncols,
                         int col4_off, int ncols4, const FastRows& f, int lo, int ngroups, float* tmp, long long tmp_stride,
                         cudaStream_t stream) {
  FastW fw;
  for (int k = 0; k < kFastTapsMax; ++k) fw.w[k] = f.w[k];
  dim3 grid((ncols4 + 127) / 128, ngroups, 2 * nset);
  pyr_vertical_fast_kernel<TP><<<grid, 128, 0, stream>>>(in0a, in0b, nset, in_stride, in_pitch, col4_off, ncols4, lo, ngroups, f.s0,
                                                          fw, tmp, tmp_stride, ncols);
}

// horizontal fast pass on all output rows except [rs_lo, rs_hi)
template <int C, int TP>
static void launch_hfast(const float* tmp, long long tmp_stride, int ncols, const FastRows& f, int lo, int ngroups, int ny_out,
                         int rs_lo, int rs_hi, float* out0a, float* out0b, int nset, long long out_stride, int out_pitch,
                         const MinMaxKeys* mm_parent, MinMaxKeys* mm_child, int mm_stride, cudaStream_t stream) {
  FastW fw;
  for (int k = 0; k < kFastTapsMax; ++k) fw.w[k] = f.w[k];
  dim3 grid((ngroups * C + 127) / 128, ny_out - (rs_hi - rs_lo), 2 * nset);
  pyr_horizontal_fast_kernel<C, TP><<<grid, 128, 0, stream>>>(tmp, tmp_stride, ncols, rs_lo, rs_hi, lo, ngroups, f.s0, fw, out0a,
                                                               out0b, nset, out_stride, out_pitch, mm_parent, mm_child, mm_stride);
}

template <int C, int TP>
static void launch_fused(const float* in0a, const float* in0b, int nset, long long in_stride, int in_pitch, int ncols,
                         const FastRows& fy, int vlo, int ngv, const FastRows& fx, int hlo, int hhi,
                         float* out0a, float* out0b, long long out_stride, int out_pitch,
                         const MinMaxKeys* mm_parent, MinMaxKeys* mm_child, int mm_stride, cudaStream_t stream,
                         MinMaxKeys* mm_gather, int* own /* r0, r1, f0, f1 of the gathered region */) {
  FastW fwy;
  FastW2 fwx;
  for (int k = 0; k < kFastTapsMax; ++k) fwy.w[k] = fy.w[k];
  for (int j = 0; j < kFastTapsMax + 2; ++j)
    fwx.wp[j] = make_float2(j < TP ? fx.w[j] : 0.f, (j >= 2 && j - 2 < TP) ? fx.w[j - 2] : 0.f);
  constexpr int OWB = FusedCfg<C>::OWB;
  dim3 grid((hhi - hlo + OWB - 1) / OWB, ngv, 2 * nset);
  if (mm_gather) {
    pyr_fused_fast_kernel<C, TP, true><<<grid, 128, 0, stream>>>(in0a, in0b, nset, in_stride, in_pitch, ncols, vlo, fy.s0, fwy, hlo, hhi,
                                                                  fx.s0, fwx, out0a, out0b, out_stride, out_pitch, mm_parent, mm_child,
                                                                  mm_stride, mm_gather);
    // the region the blocks own (same arithmetic as the kernel): rows and floats, clamped to the image
    constexpr int WF = ((2 * (OWB - 1) + TP) * C + 3 + 3) / 4 * 4, NW4 = WF / 4;
    constexpr int kOwnJ0 = (TP - 2) / 2, kOwnT0 = (NW4 - 2 * OWB * C / 4) / 2;
    const int f00 = ((2 * hlo + fx.s0) * C) & ~3;
    own[0] = 2 * vlo + fy.s0 + kOwnJ0;
    own[1] = own[0] + 2 * kFRV * ngv;
    own[2] = std::min(ncols, f00 + 4 * kOwnT0);
    own[3] = std::min(ncols, f00 + 4 * kOwnT0 + 2 * OWB * C * (int)grid.x);
  } else {
    pyr_fused_fast_kernel<C, TP, false><<<grid, 128, 0, stream>>>(in0a, in0b, nset, in_stride, in_pitch, ncols, vlo, fy.s0, fwy, hlo, hhi,
                                                                   fx.s0, fwx, out0a, out0b, out_stride, out_pitch, mm_parent, mm_child,
                                                                   mm_stride, nullptr);
  }
}

// True when launch_pyr_down would take the fused interior kernel for this level, i.e. when it can gather the parent's
// min / max itself (same conditions as in launch_pyr_down).
bool pyr_level_gathers(const float* in0a, const float* in0b, long long in_stride, int in_pitch, int nx_in, int ny_in, int channels,
                       const DeviceResample& ry, const DeviceResample& rx, const float* tmp, long long tmp_stride) {
  const int ncols = nx_in * channels;
  const bool valign = (ncols % 4 == 0) && (in_pitch % 4 == 0) && (in_stride % 4 == 0) && (tmp_stride % 4 == 0) &&
                      ((reinterpret_cast<unsigned long long>(in0a) | reinterpret_cast<unsigned long long>(in0b) |
                        reinterpret_cast<unsigned long long>(tmp)) & 15ull) == 0;
  const int tmax = std::max(ry.fast.taps, rx.fast.taps);
  const int TP = tmax <= 32 ? 32 : kFastTapsMax;
  int vlo = 0, ngv = 0, fhl = 0, fhn = 0;
  if (valign && ry.fast.hi - ry.fast.lo >= kFRV) fast_groups(ry.fast, ny_in, TP, kFRV, &vlo, &ngv);
  if (ngv > 0 && rx.fast.hi - rx.fast.lo >= kFH) fast_groups(rx.fast, nx_in, TP, kFH, &fhl, &fhn);
  return ngv > 0 && fhn > 0;
}

// Deferred clip after a gathering level: clamp the child images whose values left the parent's range, and their keys.
cudaError_t launch_clip_fixup(float* out0a, float* out0b, long long out_stride, int out_pitch, int nx_out, int ny_out, int channels,
                              int nset, const MinMaxKeys* mm_parent, MinMaxKeys* mm_child, int mm_stride, int* flags,
                              cudaStream_t stream) {
  const int nimg = 2 * nset;
  clip_decide_kernel<<<(nimg + 127) / 128, 128, 0, stream>>>(mm_parent, mm_child, mm_stride, nimg, nset, flags);
  clip_apply_kernel<<<dim3(std::min(ny_out, 32), nimg), 256, 0, stream>>>(out0a, out0b, nset, out_stride, out_pitch, nx_out * channels,
                                                                           ny_out, mm_parent, mm_child, mm_stride, flags);
  clip_keys_kernel<<<(nimg + 127) / 128, 128, 0, stream>>>(mm_parent, mm_child, mm_stride, nimg, nset, flags);
  return cudaGetLastError();
}

// One level for `nset` pairs: both image sets (a = I1, b = I2) in the same launches.  mm_parent / mm_child
// point at the pair's (level, image 0) key of the parent / child level; image 1's key follows it.
cudaError_t launch_pyr_down(const float* in0a, const float* in0b, long long in_stride, int in_pitch, int nx_in, int ny_in,
                            int channels, const DeviceResample& ry, const DeviceResample& rx, float* tmp,
                            long long tmp_stride, float* out0a, float* out0b, long long out_stride, int out_pitch, int nset,
                            const MinMaxKeys* mm_parent, MinMaxKeys* mm_child, int mm_stride, cudaStream_t stream,
                            int* launches, cudaStream_t border_stream, MinMaxKeys* mm_gather, const MinMaxKeys* mm_open) {
  // everything but the fused interior kernel (border rows / columns: ~5 % of the outputs in small, latency-bound
  // launches) can run on a second stream next to it; the caller joins the two streams after the level
  cudaStream_t bs = border_stream ? border_stream : stream;
  const int ncols = nx_in * channels;
  const int ny_out = ry.n_out, nx_out = rx.n_out;
  int nl = 0;
  const bool valign = (ncols % 4 == 0) && (in_pitch % 4 == 0) && (in_stride % 4 == 0) && (tmp_stride % 4 == 0) &&
                      ((reinterpret_cast<unsigned long long>(in0a) | reinterpret_cast<unsigned long long>(in0b) |
                        reinterpret_cast<unsigned long long>(tmp)) & 15ull) == 0;
  // uniform (exact 2:1) row range [vlo, vhi) and column range [hlo, hhi); TP covers both operators' taps
  const int tmax = std::max(ry.fast.taps, rx.fast.taps);
  const int TP = tmax <= 32 ? 32 : kFastTapsMax;
  int vlo = 0, vhi = 0, ngv = 0, hlo = 0, hhi = 0;
  if (valign && ry.fast.hi - ry.fast.lo >= kFRV) {
    fast_groups(ry.fast, ny_in, TP, kFRV, &vlo, &ngv);
    vhi = vlo + ngv * kFRV;
  }
  // the fused kernel needs both ranges; its horizontal tasks produce kFH outputs
  int fhl = 0, fhn = 0;
  if (ngv > 0 && rx.fast.hi - rx.fast.lo >= kFH) fast_groups(rx.fast, nx_in, TP, kFH, &fhl, &fhn);
  const bool fused = ngv > 0 && fhn > 0;
  // gather mode (see pyr_level_gathers): the parent's min / max is produced by this level's own first read; nothing is
  // clipped here (the border kernels get an open range), the caller clamps afterwards if a value left the range
  if (mm_gather && (!fused || !mm_open)) return cudaErrorInvalidValue;
  if (mm_gather) mm_parent = mm_open;
  int own[4] = {0, 0, 0, 0};
#define ICA_TP_C(fn, ...) do { if (channels == 3) { if (TP == 32) fn<3, 32>(__VA_ARGS__); else fn<3, kFastTapsMax>(__VA_ARGS__); } \
                               else { if (TP == 32) fn<1, 32>(__VA_ARGS__); else fn<1, kFastTapsMax>(__VA_ARGS__); } } while (0)
#define ICA_VF(...) do { if (TP == 32) launch_vfast<32>(__VA_ARGS__); else launch_vfast<kFastTapsMax>(__VA_ARGS__); } while (0)
  auto hgeneral = [&](int r0, int nrows, int rs_lo, int rs_hi, int skip_lo, int skip_hi) {
    const int ne = (nx_out - (skip_hi - skip_lo)) * channels;
    if (ne <= 0 || nrows <= 0) return;
    const int threads = ne >= 128 ? 128 : ((ne + 31) / 32) * 32;
    dim3 grid((ne + threads - 1) / threads, (nrows + kHR - 1) / kHR, 2 * nset);
    if (channels == 3)
      pyr_horizontal_border_kernel<3><<<grid, threads, 0, bs>>>(tmp, tmp_stride, ncols, nx_out, r0, nrows, rs_lo, rs_hi, skip_lo,
                                                                     skip_hi, rx.weights_t, rx.start, rx.taps, out0a, out0b, nset,
                                                                     out_stride, out_pitch, mm_parent, mm_child, mm_stride);
    else
      pyr_horizontal_border_kernel<1><<<grid, threads, 0, bs>>>(tmp, tmp_stride, ncols, nx_out, r0, nrows, rs_lo, rs_hi, skip_lo,
                                                                     skip_hi, rx.weights_t, rx.start, rx.taps, out0a, out0b, nset,
                                                                     out_stride, out_pitch, mm_parent, mm_child, mm_stride);
    ++nl;
  };
  if (fused) {
    hlo = fhl; hhi = hlo + fhn * kFH;
    // (1) interior x interior: both passes in one kernel, no intermediate image
    ICA_TP_C(launch_fused, in0a, in0b, nset, in_stride, in_pitch, ncols, ry.fast, vlo, ngv, rx.fast, hlo, hhi, out0a, out0b,
             out_stride, out_pitch, mm_parent, mm_child, mm_stride, stream, mm_gather, own);
    ++nl;
    if (mm_gather) {   // the frame of the parent level that no block owns
      minmax_frame_kernel<<<dim3(std::min(ny_in, 64), 2 * nset), 256, 0, bs>>>(in0a, in0b, nset, in_stride, in_pitch, ncols, ny_in, own[0],
                                                                                own[1], own[2], own[3], mm_gather, mm_stride);
      ++nl;
    }
    // (2) interior rows x border columns: vertical fast pass on the two column strips the border outputs read
    int in_left = 0, in_right = nx_in;
    for (int ox = 0; ox < hlo; ++ox) in_left = std::max(in_left, rx.start_host[ox] + rx.taps);
    for (int ox = hhi; ox < nx_out; ++ox) in_right = std::min(in_right, rx.start_host[ox]);
    const int l4 = std::min(ncols / 4, (in_left * channels + 3) / 4);            // strip [0, l4) in float4 columns
    const int r4 = std::max(l4, std::min(ncols / 4, (in_right * channels) / 4)); // strip [r4, ncols/4)
    if (hlo > 0 && l4 > 0) { ICA_VF(in0a, in0b, nset, in_stride, in_pitch, ncols, 0, l4, ry.fast, vlo, ngv, tmp, tmp_stride, bs); ++nl; }
    if (hhi < nx_out && ncols / 4 - r4 > 0) {
      ICA_VF(in0a, in0b, nset, in_stride, in_pitch, ncols, r4, ncols / 4 - r4, ry.fast, vlo, ngv, tmp, tmp_stride, bs); ++nl;
    }
    hgeneral(vlo, vhi - vlo, 0x7fffffff, 0x7fffffff, hlo, hhi);
  } else if (ngv > 0) {
    ICA_VF(in0a, in0b, nset, in_stride, in_pitch, ncols, 0, ncols / 4, ry.fast, vlo, ngv, tmp, tmp_stride, bs);
    ++nl;
  }
  // (3) border rows (all rows when nothing is uniform): general vertical pass ...
  if (ny_out - (vhi - vlo) > 0) {
    const int ngroups = (vlo + kVR - 1) / kVR + (ny_out - vhi + kVR - 1) / kVR;
    // the 4-row kernel keeps the weights of a row group over the group's common input range in shared memory
    // (kVSpan rows): operators whose rows start far apart (strong down-scaling) take the scalar kernel instead
    bool span_ok = (int)ry.start_host.size() == ny_out;
    for (int g = 0; span_ok && g < ngroups; ++g) {
      const int gA = (vlo + kVR - 1) / kVR;
      int lo = 0x7fffffff, hi = 0;
      for (int r = 0; r < kVR; ++r) {
        const int oy = g < gA ? g * kVR + r : vhi + (g - gA) * kVR + r;
        if ((g < gA && oy >= vlo) || oy >= ny_out) continue;
        lo = std::min(lo, ry.start_host[oy]); hi = std::max(hi, ry.start_host[oy] + ry.taps);
      }
      if (hi - lo > kVSpan) span_ok = false;
    }
    if (valign && span_ok) {
      dim3 grid((ncols / 4 + 127) / 128, ngroups, 2 * nset);
      pyr_vertical_border4_kernel<<<grid, 128, 0, bs>>>(in0a, in0b, nset, in_stride, in_pitch, ncols / 4, ny_out, vlo, vhi,
                                                             ry.weights, ry.start, ry.taps, tmp, tmp_stride, ncols);
    } else {
      dim3 grid((ncols + 255) / 256, ngroups, 2 * nset);
      pyr_vertical_kernel<<<grid, 256, 0, bs>>>(in0a, in0b, nset, in_stride, in_pitch, ncols, ny_out, vlo, vhi, ry.weights,
                                                     ry.start, ry.taps, tmp, tmp_stride);
    }
    ++nl;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // ... and the horizontal pass of every row the fused kernel did not produce
  const int rs_lo = fused ? vlo : 0, rs_hi = fused ? vhi : 0;
  int glo = 0, ghi = 0;   // column range of the uniform horizontal kernel on those rows
  if (ny_out - (rs_hi - rs_lo) > 0) {
    if (rx.fast.hi - rx.fast.lo >= kFR) {
      int lo = 0, ngroups = 0;
      fast_groups(rx.fast, nx_in, TP, kFR, &lo, &ngroups);
      if (ngroups > 0) {
        glo = lo; ghi = glo + ngroups * kFR;
        ICA_TP_C(launch_hfast, tmp, tmp_stride, ncols, rx.fast, lo, ngroups, ny_out, rs_lo, rs_hi, out0a, out0b, nset, out_stride,
                 out_pitch, mm_parent, mm_child, mm_stride, bs);
        ++nl;
      }
    }
    hgeneral(0, ny_out - (rs_hi - rs_lo), rs_lo, rs_hi, glo, ghi);
  }
#undef ICA_TP_C
#undef ICA_VF
  if (launches) *launches = nl;
  return cudaGetLastError();
}

// `nimg` images of `count` elements each; mm[img * mm_stride] receives the min/max keys of the converted image
cudaError_t launch_convert_u8(const unsigned char* in, float* out, long long count, int nimg, MinMaxKeys* mm, int mm_stride,
                              cudaStream_t stream) {
  const int blocks = (int)std::max<long long>(1, std::min<long long>((count / 4 + 256 * 8 - 1) / (256 * 8), 256));
  convert_minmax_kernel<unsigned char><<<dim3(blocks, nimg), 256, 0, stream>>>(in, out, count, mm, mm_stride);
  return cudaGetLastError();
}
cudaError_t launch_convert_f64(const double* in, float* out, long long count, int nimg, MinMaxKeys* mm, int mm_stride,
                               cudaStream_t stream) {
  const int blocks = (int)std::max<long long>(1, std::min<long long>((count + 256 * 8 - 1) / (256 * 8), 256));
  convert_minmax_kernel<double><<<dim3(blocks, nimg), 256, 0, stream>>>(in, out, count, mm, mm_stride);
  return cudaGetLastError();
}

int max_taps() { return kMaxTaps; }

}  // namespace ica
