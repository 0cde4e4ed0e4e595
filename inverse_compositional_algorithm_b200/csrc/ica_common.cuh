// Shared definitions of the B200 inverse-compositional library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ica_b200.h"

namespace ica {

// src/constants.py:1-6
constexpr int    kMaxIter     = 30;
constexpr double kLambda0     = 80.0;
constexpr double kLambdaN     = 5.0;
constexpr double kLambdaRatio = 0.9;
constexpr int    kSplinePad   = 12;  // scipy.ndimage _prepad_for_spline_filter (grid-constant)

enum Transform { TRANSLATION = 1, EUCLIDEAN = 2, SIMILARITY = 3, AFFINITY = 4, HOMOGRAPHY = 5 };
enum Robust { QUADRATIC = 0, TRUNCATED_QUADRATIC = 1, GERMAN_MCCLURE = 2, LORENTZIAN = 3, CHARBONNIER = 4 };

// ---- moment layout -------------------------------------------------------------------------
// The per-iteration kernel does not accumulate the n x n Hessian directly.  With
//   S = sum_c grad(I1_c) grad(I1_c)^T  (2x2, 3 unique),   v = sum_c grad(I1_c) * DI_c  (2),
// H = sum_px rho' J^T S J and b = sum_px rho' J^T v, and every entry of the Jacobians of
// src/derivatives.py:31-68 is a monomial in the integer pixel coordinates (x, y).  So H and b
// are exact linear combinations of the moments
//   MH[ij][a][b] = sum_px rho' S_ij x^a y^b  (a+b <= DH),   MB[i][a][b] = sum_px rho' v_i x^a y^b (a+b <= DH/2)
// with DH = 0 (translation), 2 (euclidean/similarity/affinity) or 4 (homography).  A warp owns
// one image row at a time: lanes accumulate the x-moments in fp32, one transposing shuffle
// reduction leaves value k on lane k, and lane k folds in y^b in fp64.
constexpr int kYPow = 5;                 // y^0..y^4
constexpr int kMaxRowVals = 21;          // 3*(4+1) + 2*(2+1)
constexpr int kAccStride = kMaxRowVals * kYPow;   // doubles per block partial (105)

__host__ __device__ inline int row_vals_h(int dh) { return 3 * (dh + 1); }
__host__ __device__ inline int row_vals(int dh) { return 3 * (dh + 1) + 2 * (dh / 2 + 1); }

struct LevelDesc {
  int nx, ny;           // level shape
  int pitch;            // floats per image row
  long long offset;     // float offset of this level inside a pair's pyramid slab (levels >= 1)
  int tiles_x, tiles_y; // tiling used by the per-iteration kernel
};

// Row-sharded mode (one large pair split across ranks): rank r of n owns the tile rows [ty0, ty1) of every
// level; first = index of its first tile, return value = number of its tiles (n = 1: the whole level).
__host__ __device__ inline int band_tiles(const LevelDesc& L, int rank, int n, int* first) {
  const int ty0 = (int)((long long)rank * L.tiles_y / n), ty1 = (int)((long long)(rank + 1) * L.tiles_y / n);
  *first = ty0 * L.tiles_x;
  return (ty1 - ty0) * L.tiles_x;
}

// Chunks (= partial slots, work items) of a level with `ntiles` tiles.  unit == 0 (tile kernel, ica_iterate.cu):
// min(ntiles, max_chunks) chunks of nearly equal size.  unit > 0 (march kernel, ica_march.cu): chunks of unit * m tiles
// (one or `mpref` tiles per consumer warp; more when max_chunks would be exceeded), the last one shorter.  Neither
// depends on what else is in the batch.
__host__ __device__ inline int chunk_tiles(int ntiles, int max_chunks, int unit, int mpref) {
  int m = ntiles >= 8 * unit ? (mpref > 1 ? mpref : 1) : 1;
  const int need = (ntiles + unit * max_chunks - 1) / (unit * max_chunks);
  if (need > m) m = need;
  return unit * m;
}
__host__ __device__ inline int chunk_count(int ntiles, int max_chunks, int unit, int mpref) {
  if (unit <= 0) return ntiles < max_chunks ? ntiles : max_chunks;
  if (ntiles <= 0) return 0;
  const int tpc = chunk_tiles(ntiles, max_chunks, unit, mpref);
  return (ntiles + tpc - 1) / tpc;
}
// tiles [*b, *e) of chunk `chunk` (relative to the first tile of the rank's band)
__host__ __device__ inline void chunk_range(int chunk, int ntiles, int nch, int max_chunks, int unit, int mpref, int* b, int* e) {
  if (unit <= 0) {
    *b = (int)((long long)chunk * ntiles / nch);
    *e = (int)((long long)(chunk + 1) * ntiles / nch);
  } else {
    const int tpc = chunk_tiles(ntiles, max_chunks, unit, mpref);
    *b = chunk * tpc;
    *e = *b + tpc < ntiles ? *b + tpc : ntiles;
  }
}

// Device-side state of one image pair (one "registration").
struct PairState {
  double p[ICA_MAX_PARAMS];       // current parameters at `scale`
  double p_prev[ICA_MAX_PARAMS];  // parameters used by the last warp (for DI / Iw, ica.py:261)
  double hinv[ICA_MAX_PARAMS * ICA_MAX_PARAMS];  // quadratic loop: H^-1 of the current scale
  double lambda_it;
  double err;
  int scale;                      // current scale, -1 when finished
  int iter;                       // iterations done at this scale
  int ttype;
  int nparams;
  int total_iters;
  int traj_count;
  int iters_per_scale[ICA_MAX_SCALES];
  unsigned int ticket;            // blocks of this pair that finished the current launch
  unsigned int pad_;
};

struct MinMaxKeys { unsigned int lo, hi; };  // order-preserving keys of floats (see float_key)

__host__ __device__ inline unsigned int float_key(float f) {
#ifdef __CUDA_ARCH__
  unsigned int b = __float_as_uint(f);
#else
  union { float f; unsigned int u; } c; c.f = f; unsigned int b = c.u;
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ inline float key_float(unsigned int k) {
  unsigned int b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; unsigned int u; } c; c.u = b; return c.f;
#endif
}

void set_error(const char* fmt, ...);

#define ICA_CUDA_CHECK(expr)                                                        \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ica::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),       \
                     __FILE__, __LINE__);                                           \
      return ICA_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

}  // namespace ica
