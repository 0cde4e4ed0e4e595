// Synthetic image pairs with a known ground-truth motion, generated on the GPU.
//
// The reference fabricates its test pairs with transformation.transform_image (src/transformation.py:266-318) inside
// the notebooks; at thousands of registrations per second a host-side generator and the host->device copies of its
// output become the bottleneck, so the benchmark-sized inputs are produced where they are consumed:
//   texture  = white Gaussian noise on (H + 2m) x (W + 2m) x C, blurred by a separable Gaussian (sigma 2, radius 8,
//              periodic), affinely normalised to [0, 255];
//   I2       = centre crop of the texture;
//   I1(x)    = texture(x'(x; p_gt) + m) by Catmull-Rom interpolation (clamped indices) + N(0, noise_sigma) noise,
//              optionally a square of uniform noise (occlusion), clamped to [0, 255]; both optionally rounded to
//              8-bit values.
// Randomness is a counter-based hash (splitmix64 of seed, pair, stream and element index + Box-Muller), so a pixel's
// value depends only on (seed, pair, position): any pair can be regenerated alone, and
// inverse_compositional_algorithm_b200/synthetic.py holds a numpy mirror of exactly this pipeline for the CPU legs
// of the benchmark and the tests.  Input synthesis only: nothing here is on the registration path.
#include <stdint.h>
#include <math.h>
#include <algorithm>
#include <vector>
#include "ica_common.cuh"
#include "ica_transform.cuh"

namespace ica {
namespace {

constexpr int kGenRadius = 8;       // Gaussian radius (sigma = 2, truncate 4)
constexpr int kGenTile = 32;        // output tile of the fused noise + blur kernel
constexpr int kGenHalo = kGenTile + 2 * kGenRadius;

__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// stream 0: texture noise, 1: additive noise of I1, 2: occlusion
__host__ __device__ inline unsigned long long gen_key(unsigned long long seed, int pair, int stream) {
  return splitmix64(seed ^ splitmix64(((unsigned long long)(unsigned)pair << 8) | (unsigned)stream));
}
__device__ __forceinline__ float gen_uniform(unsigned long long key, unsigned long long idx) {   // [0, 1)
  return (float)(splitmix64(key + idx * 0xD1342543DE82EF95ull) >> 40) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float gen_normal(unsigned long long key, unsigned long long idx) {
  const unsigned long long h = splitmix64(key + idx * 0xD1342543DE82EF95ull);
  const float u1 = ((float)(h >> 40) + 1.0f) * (1.0f / 16777216.0f);            // (0, 1]
  const float u2 = (float)((h >> 16) & 0xFFFFFFull) * (1.0f / 16777216.0f);     // [0, 1)
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

struct GenWeights { float w[2 * kGenRadius + 1]; };

// texture[pair][y][x][c] (unnormalised) = separable Gaussian blur of the hash noise, periodic; min / max per pair.
// One block = one 32 x 32 tile of one channel: noise -> shared, horizontal pass -> shared, vertical pass -> global.
__global__ void __launch_bounds__(256) gen_texture_kernel(float* __restrict__ tex, int Ht, int Wt, int C, int pair0,
                                                          unsigned long long seed, GenWeights gw, MinMaxKeys* __restrict__ mm) {
  __shared__ float sn[kGenHalo][kGenHalo + 1];
  __shared__ float sh[kGenHalo][kGenTile + 1];
  __shared__ unsigned smin[8], smax[8];
  const int pl = blockIdx.z / C, c = blockIdx.z % C;      // pair within this launch, channel
  const unsigned long long key = gen_key(seed, pair0 + pl, 0);
  const int x0 = blockIdx.x * kGenTile, y0 = blockIdx.y * kGenTile;
  for (int i = threadIdx.x; i < kGenHalo * kGenHalo; i += blockDim.x) {
    const int ly = i / kGenHalo, lx = i % kGenHalo;
    int gy = (y0 + ly - kGenRadius) % Ht, gx = (x0 + lx - kGenRadius) % Wt;
    if (gy < 0) gy += Ht;
    if (gx < 0) gx += Wt;
    sn[ly][lx] = gen_normal(key, ((unsigned long long)gy * Wt + gx) * C + c);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kGenHalo * kGenTile; i += blockDim.x) {
    const int ly = i / kGenTile, lx = i % kGenTile;
    float a = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * kGenRadius; ++k) a = fmaf(gw.w[k], sn[ly][lx + k], a);
    sh[ly][lx] = a;
  }
  __syncthreads();
  float vmin = 3.4e38f, vmax = -3.4e38f;
  float* out = tex + (long long)pl * Ht * Wt * C;
  for (int i = threadIdx.x; i < kGenTile * kGenTile; i += blockDim.x) {
    const int ly = i / kGenTile, lx = i % kGenTile;
    const int y = y0 + ly, x = x0 + lx;
    if (y >= Ht || x >= Wt) continue;
    float a = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * kGenRadius; ++k) a = fmaf(gw.w[k], sh[ly + k][lx], a);
    out[((long long)y * Wt + x) * C + c] = a;
    vmin = fminf(vmin, a); vmax = fmaxf(vmax, a);
  }
  unsigned kmin = vmin <= vmax ? float_key(vmin) : 0xffffffffu, kmax = vmin <= vmax ? float_key(vmax) : 0u;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { smin[warp] = kmin; smax[warp] = kmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { kmin = min(kmin, smin[w]); kmax = max(kmax, smax[w]); }
    if (kmin <= kmax) { atomicMin(&mm[pl].lo, kmin); atomicMax(&mm[pl].hi, kmax); }
  }
}

__global__ void gen_reset_kernel(MinMaxKeys* mm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { mm[i].lo = 0xffffffffu; mm[i].hi = 0u; }
}

struct GenPairDesc {
  double m[9];            // params2matrix(p_gt): x'(x) in I2 / texture coordinates (before the margin)
  int occ_x, occ_y, occ_side;
  int pad_;
};

__device__ __forceinline__ float keys_w(float t, int j) {
  const float t2 = t * t;
  switch (j) {
    case 0: return t * fmaf(t, fmaf(-0.5f, t, 1.0f), -0.5f);
    case 1: return fmaf(t2, fmaf(1.5f, t, -2.5f), 1.0f);
    case 2: return t * fmaf(t, fmaf(-1.5f, t, 2.0f), 0.5f);
    default: return t2 * fmaf(0.5f, t, -0.5f);
  }
}

// I1 and I2 of the pairs of this launch from their (unnormalised) textures
__global__ void __launch_bounds__(256) gen_pair_kernel(const float* __restrict__ tex, int Ht, int Wt, int C, int H, int W,
                                                       int margin, int pair0, unsigned long long seed,
                                                       const GenPairDesc* __restrict__ desc, const MinMaxKeys* __restrict__ mm,
                                                       float noise_sigma, int quantize, float* __restrict__ I1,
                                                       float* __restrict__ I2) {
  const int pl = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const float lo = key_float(mm[pl].lo), hi = key_float(mm[pl].hi);
  const float scale = 255.0f / (hi - lo);
  const float* t = tex + (long long)pl * Ht * Wt * C;
  const GenPairDesc d = desc[pl];
  // x'(x; p_gt) in fp64 (the ground truth is only as good as this evaluation)
  const double zz = d.m[6] * x + d.m[7] * y + d.m[8];
  const double xs = (d.m[0] * x + d.m[1] * y + d.m[2]) / zz + margin;
  const double ys = (d.m[3] * x + d.m[4] * y + d.m[5]) / zz + margin;
  const double fxs = floor(xs), fys = floor(ys);
  const int cx = (int)fxs, cy = (int)fys;
  const float tx = (float)(xs - fxs), ty = (float)(ys - fys);
  float wx[4], wy[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { wx[j] = keys_w(tx, j); wy[j] = keys_w(ty, j); }
  const unsigned long long kn = gen_key(seed, pair0 + pl, 1), ko = gen_key(seed, pair0 + pl, 2);
  const bool occ = d.occ_side > 0 && x >= d.occ_x && x < d.occ_x + d.occ_side && y >= d.occ_y && y < d.occ_y + d.occ_side;
  const long long o = (((long long)pl * H + y) * W + x) * C;      // I1 / I2 point at this launch's first pair
  for (int c = 0; c < C; ++c) {
    float v2 = (t[((long long)(y + margin) * Wt + (x + margin)) * C + c] - lo) * scale;
    float a = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int yy = min(max(cy - 1 + q, 0), Ht - 1);
      float h = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int xx = min(max(cx - 1 + j, 0), Wt - 1);
        h = fmaf(wx[j], t[((long long)yy * Wt + xx) * C + c], h);
      }
      a = fmaf(wy[q], h, a);
    }
    float v1 = (a - lo) * scale;
    const unsigned long long idx = ((unsigned long long)y * W + x) * C + c;
    if (noise_sigma > 0.f) v1 = fmaf(noise_sigma, gen_normal(kn, idx), v1);
    if (occ) v1 = 255.0f * gen_uniform(ko, idx);
    v1 = fminf(fmaxf(v1, 0.f), 255.f);
    if (quantize) { v1 = rintf(v1); v2 = rintf(v2); }
    I1[o + c] = v1;
    I2[o + c] = v2;
  }
}

}  // namespace
}  // namespace ica

using namespace ica;

extern "C" {

// B pairs into I1 / I2 (float32 [B][H][W][C], device).  p_gt: [B][8] ground-truth parameters (host), ttypes [B]
// (host); occ_xy: [B][2] top-left corner of the occluding square (host) or NULL, occ_side its side (0 = none).
// A value depends only on (seed, pair index, position): pair_offset shifts the pair indices, so that rank r of a
// sharded job generates pairs [r*B, (r+1)*B) of one global set.  Runs on `stream` and returns when it is done.
int ica_generate_pairs_device(float* I1_dev, float* I2_dev, int32_t batch, int32_t height, int32_t width, int32_t channels,
                              const int32_t* ttypes, const double* p_gt, const int32_t* occ_xy, int32_t occ_side,
                              uint64_t seed, int32_t pair_offset, int32_t margin, double noise_sigma, int32_t quantize,
                              void* stream_) {
  if (!I1_dev || !I2_dev || !ttypes || !p_gt || batch < 1 || height < 1 || width < 1 || (channels != 1 && channels != 3) ||
      margin < 0 || margin > 4096) {
    set_error("bad argument"); return ICA_ERR_INVALID;
  }
  for (int b = 0; b < batch; ++b) if (nparams_of(ttypes[b]) < 0) { set_error("Unknown transform type"); return ICA_ERR_INVALID; }
  cudaStream_t stream = (cudaStream_t)stream_;
  const int Ht = height + 2 * margin, Wt = width + 2 * margin;
  GenWeights gw;
  {
    const double sigma = 2.0;
    double sum = 0.0, w[2 * kGenRadius + 1];
    for (int k = -kGenRadius; k <= kGenRadius; ++k) { w[k + kGenRadius] = exp(-0.5 * k * k / (sigma * sigma)); sum += w[k + kGenRadius]; }
    for (int k = 0; k <= 2 * kGenRadius; ++k) gw.w[k] = (float)(w[k] / sum);
  }
  // textures of a group of pairs at a time (256 MiB of scratch at most)
  const long long tex_floats = (long long)Ht * Wt * channels;
  const int group = (int)std::max<long long>(1, std::min<long long>(batch, (64ll << 20) / tex_floats));
  float* tex = nullptr; MinMaxKeys* mm = nullptr; GenPairDesc* desc = nullptr;
  cudaError_t e = cudaMalloc(&tex, (size_t)group * tex_floats * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&mm, (size_t)batch * sizeof(MinMaxKeys));
  if (e == cudaSuccess) e = cudaMalloc(&desc, (size_t)batch * sizeof(GenPairDesc));
  std::vector<GenPairDesc> hd(batch);
  for (int b = 0; b < batch && e == cudaSuccess; ++b) {
    params2matrix(p_gt + (size_t)b * ICA_MAX_PARAMS, ttypes[b], hd[b].m);
    hd[b].occ_side = occ_xy ? occ_side : 0;
    hd[b].occ_x = occ_xy ? occ_xy[2 * b] : 0;
    hd[b].occ_y = occ_xy ? occ_xy[2 * b + 1] : 0;
    hd[b].pad_ = 0;
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(desc, hd.data(), (size_t)batch * sizeof(GenPairDesc), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) { gen_reset_kernel<<<(batch + 255) / 256, 256, 0, stream>>>(mm, batch); e = cudaGetLastError(); }
  for (int b0 = 0; b0 < batch && e == cudaSuccess; b0 += group) {
    const int nb = std::min(group, batch - b0);
    dim3 gt((Wt + kGenTile - 1) / kGenTile, (Ht + kGenTile - 1) / kGenTile, nb * channels);
    gen_texture_kernel<<<gt, 256, 0, stream>>>(tex, Ht, Wt, channels, pair_offset + b0, seed, gw, mm + b0);
    dim3 blk(32, 8), gp((width + 31) / 32, (height + 7) / 8, nb);
    gen_pair_kernel<<<gp, blk, 0, stream>>>(tex, Ht, Wt, channels, height, width, margin, pair_offset + b0, seed, desc + b0,
                                            mm + b0, (float)noise_sigma, quantize, I1_dev + (long long)b0 * height * width * channels,
                                            I2_dev + (long long)b0 * height * width * channels);
    e = cudaGetLastError();
  }
  // the scratch is released once the stream has consumed it
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFree(tex); cudaFree(mm); cudaFree(desc);
  if (e != cudaSuccess) { set_error("ica_generate_pairs_device: %s", cudaGetErrorString(e)); return ICA_ERR_CUDA; }
  return ICA_OK;
}

}  // extern "C"
