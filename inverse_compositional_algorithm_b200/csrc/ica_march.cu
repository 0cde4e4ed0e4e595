// K2, second generation: the fused per-iteration kernel as a COLUMN MARCH with a register-resident bicubic window.
//
// Same contract as ica_iterate_kernel (ica_iterate.cu): one launch = one iteration of every still-active pair; per
// chunk of tiles it writes the [K][kYPow] fp64 moment sums from which ica_solve_kernel assembles H and b
// (src/inverse_compositional_algorithm.py:109-131, 225-259; bi.bicubic_interpolation_skimage, io.robust_error_function,
// io.independent_vector[_robust], de.hessian[_robust]).
//
// Why a second kernel: the first one fetches the 4 x 4 x C taps of every pixel from shared memory (48 + 15 loads per
// RGB pixel) and is bound by the shared-memory data pipe at ~0.25 of the HBM roofline (profiles/README.md).  Here
//   * a lane owns one image COLUMN of a 32-column tile and marches down its rows.  The taps of pixel (x, y+1) are the
//     taps of (x, y) shifted by one row, so the lane keeps a 5 x 5 x C window of I2 in REGISTERS and loads one new
//     window row (5 C floats) per pixel instead of 16 C.  The window is one column and one row larger than the
//     footprint: the pixel's own 4 x 4 taps sit at offset (dcx, dcy) in {0,1}^2 inside it, which absorbs the drift of
//     the projected position along the march (rotation, scale, perspective) -- the Keys weights are zero-padded to
//     five entries, so the arithmetic of the valid taps is unchanged (adding exact zeros).  A pixel whose footprint
//     leaves the window (drift > 1 px inside one 16-row tile: > 3.5 degrees of rotation) samples global memory.
//   * the window rows are static register names: the march is unrolled by five rows (the window's rotation period).
//   * I1: the centre value is carried down the column in registers; left / right / below come from shared memory.
//   * moments: x is fixed per lane, so lanes accumulate the y-moments (rho' S y^b, rho' v y^b) in fp32 and fold in
//     x^a once per column run (the mirror image of the first kernel's row-wise scheme).
//   * every consumer warp is its own pipeline: it computes the I2 box its tile needs from its lanes' own projections,
//     issues its own tiled TMA copies (5 window rows + 6 I1 rows per block, three blocks in flight, zero fill outside
//     the image; the NaN footprint of skimage's cval is evaluated analytically) and waits on its own mbarriers.
//     There is no producer warp and no barrier between warps inside a chunk; a twelfth warp fetches work items and
//     prepares the chunk descriptors ahead of the consumers.
// Per RGB pixel: 15 + 9 shared loads instead of 63, ~210 instead of ~320 instructions.
#include <atomic>
#include <type_traits>
#include "ica_device.cuh"
#include "ica_transform.cuh"
#include "ica_iterate.cuh"

namespace ica {

namespace {

#ifndef ICA_MARCH_WARPS
#define ICA_MARCH_WARPS 11
#endif
#ifndef ICA_MARCH_ROWS
#define ICA_MARCH_ROWS 16
#endif
constexpr int MW = ICA_MARCH_WARPS;          // consumer warps (+ 1 scheduler warp)
constexpr int kMThreads = (MW + 1) * 32;
constexpr int kMConsumerThreads = MW * 32;
constexpr int MTW = 32;                      // tile width: one column per lane
constexpr int MTH = ICA_MARCH_ROWS;          // tile height: rows of one march
constexpr int MNS = 3;                       // blocks in flight per warp
constexpr int MRB = 5;                       // I2 rows per block = rotation period of the register window
constexpr int MR1 = 6;                       // I1 rows per block (block j holds the rows y0 - 4 + 5 j ... + 5)
constexpr int MHALO = 4;                     // the I1 box starts at x0 - 4 (16-byte box corner)
constexpr int kMaxSpread = 5;                // rows by which the lanes' windows of one tile may differ
constexpr int kItemPairBitsM = 20;

// Accounting of one run (tools/march_stats.py; library built with -DICA_MARCH_STATS): per CTA, accumulated over the
// launches, in the plan's debug buffer -- tiles, tiles without a staged box, warp steps, pixels on the global-memory
// path, and the cycles the consumer warps spend waiting / issuing / folding.
#ifdef ICA_MARCH_STATS
#define MSTAT_ADD(slot, val) do { if (P.dbg_time && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(P.dbg_time) + blockIdx.x * 16 + (slot), (unsigned long long)(val)); } while (0)
#define MSTAT_CLK() clock64()
#else
#define MSTAT_ADD(slot, val) do { (void)(val); } while (0)
#define MSTAT_CLK() 0ll
#endif

template <int C> struct MGeo {
  static constexpr int S2W = C == 3 ? 160 : 64;          // floats per I2 window row (== 0 mod 32 banks)
  static constexpr int S1W = C == 3 ? 112 : 48;          // floats per I1 row (x0-4 ... x0+32 and padding)
  static constexpr int BWPX = S2W / C;                   // window width in pixels (53 / 64)
  static constexpr int kRing2 = MNS * MRB * S2W;         // floats
  static constexpr int kRing1 = MNS * MR1 * S1W;
  static constexpr int kWarpFloats = kRing2 + kRing1;
  static constexpr unsigned kBytes2 = MRB * S2W * 4u, kBytes1 = MR1 * S1W * 4u;
};

template <int DH> struct MVals {
  static constexpr int HW = DH + 1, BWN = DH / 2 + 1;
  static constexpr int K = 3 * HW + 2 * BWN;
  static constexpr int NS_ = (DH + 1) * (DH + 2) / 2;          // (a, b) with a + b <= DH
  static constexpr int NB_ = (DH / 2 + 1) * (DH / 2 + 2) / 2;
  static constexpr int NF = 3 * NS_ + 2 * NB_;                 // folded sums per flush (57 / 24 / 5)
};

// fold index -> entry of the [K][kYPow] accumulator table (moment (ij, a) x y-power b)
struct FoldTab { unsigned char e[64]; };
constexpr FoldTab make_fold(int dh) {
  FoldTab t{};
  const int hw = dh + 1, bwn = dh / 2 + 1;
  int f = 0;
  for (int ij = 0; ij < 3; ++ij)
    for (int b = 0; b <= dh; ++b)
      for (int a = 0; a + b <= dh; ++a) t.e[f++] = (unsigned char)((ij * hw + a) * kYPow + b);
  for (int i = 0; i < 2; ++i)
    for (int b = 0; b <= dh / 2; ++b)
      for (int a = 0; a + b <= dh / 2; ++a) t.e[f++] = (unsigned char)((3 * hw + i * bwn + a) * kYPow + b);
  return t;
}
__constant__ FoldTab kFold[3] = {make_fold(0), make_fold(2), make_fold(4)};

// Chunk descriptor, prepared by the scheduler warp
struct __align__(16) MDesc {
  float coef[8];          // d00, m01, m02, m10, d11, m12, m20, m21 (WarpCoef)
  float4 fl;              // lo, hi: clip range of I2 at this level; lambda^2
  double m64[9];          // warp matrix in fp64 (tie-break path of project_px)
  int nx, ny, pitch, need_h;
  int t_begin, t_end, ty0, band_rows;
  int pair, chunk, stop, nch;
  int gxlo, gxspan, fxlo, fxspan;
  const float* I2;        // global image, for the pixels whose footprint leaves the register window
  const char* tm1;        // tensor maps of the level: I1, then I2 (+128 bytes)
};

// ---------------------------------------------------------------- PTX helpers (as in ica_iterate.cu)
__device__ __forceinline__ unsigned m_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void m_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(m_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void m_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(m_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void m_mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(m_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void m_mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "MLAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra MLAB_DONE;\n\t"
      "bra MLAB_WAIT;\n\t"
      "MLAB_DONE:\n\t}" ::"r"(m_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void m_tma_load_2d(void* dst, const void* tmap, int c0, int c1, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(m_smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(m_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void m_fence_tensormap_acquire(const void* tmap) {
  asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void m_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ long long m_gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void m_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kMConsumerThreads) : "memory"); }

// Transposing warp reduction (see ica_iterate.cu): on return lane l holds in v[0] the sum over all lanes of value
// index (l >> log2(32 / NP)).
template <int NP>
__device__ __forceinline__ float m_transpose_reduce(float (&v)[NP], int lane) {
  int off = 16;
#pragma unroll
  for (int h = NP / 2; h >= 1; h >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float keep = up ? v[i + h] : v[i];
      const float send = up ? v[i] : v[i + h];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
  float r = v[0];
  for (; off >= 1; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
  return r;
}

template <int C>
__device__ __noinline__ float m_sample_global_slow(const float* __restrict__ img, int pitch, int nx, int ny, int cx,
                                                   int cy, int ch, float wx0, float wx1, float wx2, float wx3,
                                                   float wy0, float wy1, float wy2, float wy3) {
  const float wx[4] = {wx0, wx1, wx2, wx3}, wy[4] = {wy0, wy1, wy2, wy3};
  return sample_global<C>(img, pitch, nx, ny, cx, cy, ch, wx, wy);
}

// ============================================================ scheduler warp
// Claims work items (chunks) and prepares their descriptors two chunks ahead of the consumers.
__device__ void m_scheduler_loop(const IterParams& P, MDesc* desc, volatile int* ready, volatile int* done, double* pm64,
                                 int par, int lane) {
  SchedHdr* const hdr = P.hdr + par;
  const int* const item_pair = P.item_pair + (long long)__ldcg(&hdr->list) * P.B * P.max_chunks;
  bool first_item = true;
  for (int seq = 0;; ++seq) {
    // descriptor slot seq % 3 belonged to chunk seq - 3: wait until the consumers are done with it
    if (seq >= 3) { while (*done < seq - 2) __nanosleep(64); }
    int item = (int)blockIdx.x;        // first item of this CTA: static (the dynamic counter starts at the grid size)
    if (!first_item) {
      if (lane == 0) item = atomicAdd(&hdr->counter, 1);
      item = __shfl_sync(0xffffffffu, item, 0);
    }
    first_item = false;
    MDesc& D = desc[seq % 3];
    if (item >= __ldcg(&hdr->total)) {
      if (lane == 0) { D.stop = 1; __threadfence_block(); *ready = seq + 1; }
      break;
    }
    const int packed = __ldcg(item_pair + item);
    const int pair = packed & ((1 << kItemPairBitsM) - 1), chunk = packed >> kItemPairBitsM;
    const PairState* stp = P.state + pair;      // written by the solve of the previous iteration: read from L2
    const int s = __ldcg(&stp->scale);
    const int st_ttype = __ldcg(&stp->ttype), st_iter = __ldcg(&stp->iter);
    const double st_lambda = __ldcg(&stp->lambda_it);
    if (lane == 0) {
      double pp[ICA_MAX_PARAMS];
#pragma unroll
      for (int i = 0; i < ICA_MAX_PARAMS; ++i) pp[i] = __ldcg(&stp->p[i]);
      warp_matrix(pp, st_ttype, pm64);
    }
    __syncwarp();
    const WarpCoef coef = make_warp_coef(pm64);
    const MinMaxKeys mm = P.mm[(pair * P.nscales + s) * 2 + 1];
    const LevelDesc L = P.lv[s];
    const int ty0 = (int)((long long)P.shard_rank * L.tiles_y / P.shard_n);
    const int ty1 = (int)((long long)(P.shard_rank + 1) * L.tiles_y / P.shard_n);
    const int ntiles = (ty1 - ty0) * L.tiles_x;
    const int nch = chunk_count(ntiles, P.max_chunks, P.chunk_unit, P.chunk_m);
    int t_begin, t_end;
    chunk_range(chunk, ntiles, nch, P.max_chunks, P.chunk_unit, P.chunk_m, &t_begin, &t_end);
    const char* tm1 = static_cast<const char*>(P.tmaps) + ((long long)(pair * P.nscales + s) * 2) * 128;
    if (lane == 0) { m_fence_tensormap_acquire(tm1); m_fence_tensormap_acquire(tm1 + 128); }
    if (lane < 9) D.m64[lane] = pm64[lane];
    if (lane == 0) {
      D.coef[0] = coef.d00; D.coef[1] = coef.m01; D.coef[2] = coef.m02; D.coef[3] = coef.m10;
      D.coef[4] = coef.d11; D.coef[5] = coef.m12; D.coef[6] = coef.m20; D.coef[7] = coef.m21;
      D.fl = make_float4(key_float(mm.lo), key_float(mm.hi), (float)(st_lambda * st_lambda), 0.0f);
      D.nx = L.nx; D.ny = L.ny; D.pitch = L.pitch; D.need_h = (P.robust_loop || st_iter == 0) ? 1 : 0;
      D.t_begin = t_begin; D.t_end = t_end; D.ty0 = ty0; D.band_rows = ty1 - ty0;
      D.pair = pair; D.chunk = chunk; D.stop = 0; D.nch = nch;
      // columns inside the discarded frame (ica.py:85-93) and, of those, the ones with a central x-difference
      const int fxlo = P.frame ? P.delta : 0;
      const int fxspan = max(0, L.nx - 2 * fxlo);
      const int gxlo = max(fxlo, 1), gxhi = min(fxlo + fxspan, L.nx - 1);
      D.gxlo = gxlo; D.gxspan = max(0, gxhi - gxlo); D.fxlo = fxlo; D.fxspan = fxspan;
      D.I2 = s == 0 ? P.I2_0 + (long long)pair * P.in_stride : P.pyr2 + (long long)pair * P.pyr_stride + L.offset;
      D.tm1 = tm1;
    }
    __syncwarp();
    if (lane == 0) { __threadfence_block(); *ready = seq + 1; }
  }
}

// Geometry of one 32 x MTH tile as a consumer warp sees it; two records per warp in shared memory (the tile being
// marched and the prefetched one), so that nothing of it occupies registers during the march
struct MTile {
  int x0, y0, nrows;      // tile corner, rows inside the image
  int bx0, by0;           // corner of the staged I2 box
  int nst;                // blocks the tile needs
  int ng;                 // groups of five march times (the first four times of a tile only fill the window)
  unsigned q0;            // stream index of the tile's first block
  int fits;               // the I2 box covers every lane's window for the whole march
  int issued;             // blocks issued so far
  int valid, pad;
  int2 XY[32];            // per lane: left column of the register window; its top row at step 0
};

// ============================================================ the kernel
template <int C, int DH>
__global__ void __launch_bounds__(kMThreads, 1) ica_march_kernel(const __grid_constant__ IterParams P) {
  using G = MGeo<C>;
  using V = MVals<DH>;
  constexpr int HW = V::HW, BWN = V::BWN, K = V::K, NF = V::NF;
  constexpr int NENT = K * kYPow;
  constexpr int S1W = G::S1W, S2W = G::S2W;

  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) unsigned long long s_full[MW][MNS];
  __shared__ MDesc s_desc[3];
  __shared__ MTile s_tile[MW][2];
  __shared__ double s_pm64[9];
  __shared__ int s_par, s_ready, s_done;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  if (tid == 0) {
    if (blockIdx.x == 0 && P.cond_handle) {      // safety net of the device-side loop (see ica_iterate_kernel)
      const int n = atomicAdd(P.loop_count + 1, 1);
      if (n > P.max_launches + 8) cudaGraphSetConditional(P.cond_handle, 0u);
    }
    s_par = __ldcg(P.loop_count) & 1;
    s_ready = 0; s_done = 0;
    for (int w = 0; w < MW; ++w) for (int i = 0; i < MNS; ++i) m_mbar_init(&s_full[w][i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // fp64 accumulators of the chunk in progress, double-buffered over consecutive chunks: [2][MW][K][kYPow]
  double* const accs0 = reinterpret_cast<double*>(smem + MW * G::kWarpFloats);
  constexpr int kAccSet = MW * NENT;
  __syncthreads();
  const int par = s_par;
  SchedHdr* const hdr = P.hdr + par;
  if ((int)blockIdx.x >= __ldcg(&hdr->total)) return;
  if (tid == 0) atomicMin(&hdr->t0, m_gtime());

  if (warp == MW) {
    m_scheduler_loop(P, s_desc, &s_ready, &s_done, s_pm64, par, lane);
    return;
  }

  // ------------------------------------------------------------------ consumers
  const int delta = P.delta;
  const bool frame = P.frame != 0;
  const bool robust = P.robust_loop != 0;
  const float chm = P.ch_mult;
  const int rtype = P.robust_type;
  const bool ipol = P.ipol_warp != 0, ipol_nan = P.ipol_nan != 0;
  float* const ring2 = smem + warp * G::kWarpFloats;       // [MNS * MRB rows][S2W]
  float* const ring1 = ring2 + G::kRing2;                   // [MNS][MR1 rows][S1W]
  unsigned long long* const full = s_full[warp];
  double* myaccs = accs0 + warp * NENT;                      // this warp's table in the current set
  MTile* const tiles = s_tile[warp];

  unsigned q_issue = 0, q_done = 0;       // blocks issued / released by this warp so far
  // per-lane y-moment accumulators of the column `vcol` (-1: empty)
  float vy[K];
#pragma unroll
  for (int i = 0; i < K; ++i) vy[i] = 0.0f;
  int vcol = -1;

  // once per column run: fold in x^a, add the lanes up (transposing shuffle reduction, fixed order), fp64 accumulate
  auto flush_col = [&]() {
    const float xf = (float)(vcol + lane);
    float xp[5];
    xp[0] = 1.0f; xp[1] = xf; xp[2] = xf * xf; xp[3] = xp[2] * xf; xp[4] = xp[2] * xp[2];
    float prod[NF];
#pragma unroll
    for (int ij = 0; ij < 3; ++ij)
#pragma unroll
      for (int b = 0; b <= DH; ++b)
#pragma unroll
        for (int a = 0; a <= DH; ++a)
          if (a + b <= DH) prod[ij * V::NS_ + b * (DH + 1) - b * (b - 1) / 2 + a] = vy[ij * HW + b] * xp[a];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int b = 0; b <= DH / 2; ++b)
#pragma unroll
        for (int a = 0; a <= DH / 2; ++a)
          if (a + b <= DH / 2) prod[3 * V::NS_ + i * V::NB_ + b * (DH / 2 + 1) - b * (b - 1) / 2 + a] = vy[3 * HW + i * BWN + b] * xp[a];
#pragma unroll
    for (int i = 0; i < K; ++i) vy[i] = 0.0f;
    constexpr int NP = NF <= 8 ? 8 : (NF <= 16 ? 16 : 32);
    constexpr int SH = NP == 8 ? 2 : (NP == 16 ? 1 : 0);
#pragma unroll
    for (int r = 0; r * NP < NF; ++r) {
      float t[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) t[i] = (r * NP + i < NF) ? prod[(r * NP + i < NF) ? r * NP + i : 0] : 0.0f;
      const float tot = m_transpose_reduce<NP>(t, lane);
      const int f = r * NP + (lane >> SH);
      if ((lane & ((1 << SH) - 1)) == 0 && f < NF) myaccs[kFold[DH / 2].e[f]] += (double)tot;
    }
    vcol = -1;
  };

  int nitems = 0;
  const long long st_t0 = MSTAT_CLK();
  for (int seq = 0;; ++seq) {
    // ---------------- the next chunk
    const long long st_d0 = MSTAT_CLK();
    while (*reinterpret_cast<volatile int*>(&s_ready) <= seq) __nanosleep(32);
    MSTAT_ADD(8, MSTAT_CLK() - st_d0);
    __threadfence_block();
    __syncwarp();
    const MDesc& D = s_desc[seq % 3];
    if (D.stop) { MSTAT_ADD(5, MSTAT_CLK() - st_t0); break; }
    // this warp's table of the set this chunk accumulates into (its previous contents were summed two chunks ago:
    // the summing threads pass the barrier in between only after they are done with it)
    for (int i = lane; i < NENT; i += 32) myaccs[i] = 0.0;
    __syncwarp();
    const int nx = D.nx, ny = D.ny;
    const bool need_h = D.need_h != 0;
    const int ntile = D.t_end - D.t_begin;
    const int i0 = D.t_begin + (int)((long long)warp * ntile / MW);
    const int i1 = D.t_begin + (int)((long long)(warp + 1) * ntile / MW);
    const char* const tm1 = D.tm1;

    // geometry of tile `ti` (band-relative index; column-major so that a warp's tiles form vertical runs) into record t
    auto tile_geo = [&](int ti, MTile* t) {
      WarpCoef coef;
      coef.d00 = D.coef[0]; coef.m01 = D.coef[1]; coef.m02 = D.coef[2]; coef.m10 = D.coef[3];
      coef.d11 = D.coef[4]; coef.m12 = D.coef[5]; coef.m20 = D.coef[6]; coef.m21 = D.coef[7];
      const int tx = ti / D.band_rows, ty = D.ty0 + ti - tx * D.band_rows;
      const int x0 = tx * MTW, y0 = ty * MTH;
      const int nrows = min(MTH, ny - y0);
      const int xl = min(x0 + lane, nx - 1);
      int cxa, cya, cxb, cyb; float ta, tb;
      const bool oka = project_px(coef, D.m64, xl, y0, cxa, cya, ta, tb);
      const bool okb = project_px(coef, D.m64, xl, y0 + nrows - 1, cxb, cyb, ta, tb);
      const int X0 = min(cxa, cxb) - 1;
      const int Yt = y0 + min(cya - y0, cyb - (y0 + nrows - 1)) - 1;
      const int mnx = __reduce_min_sync(0xffffffffu, X0), mxx = __reduce_max_sync(0xffffffffu, X0);
      const int mny = __reduce_min_sync(0xffffffffu, Yt), mxy = __reduce_max_sync(0xffffffffu, Yt);
      const bool okall = __all_sync(0xffffffffu, oka && okb);
      const int bx0 = (mnx >> 2) << 2;                        // floor to a multiple of 4 pixels (16-byte box corner)
      const int spread = mxy - mny;
      const bool fits = okall && (mxx + 5 - bx0 <= G::BWPX) && spread <= kMaxSpread &&
                        mnx > -(1 << 24) && mxx < (1 << 24) && mny > -(1 << 24) && mxy < (1 << 24);
      const int ng = (nrows + 4 + MRB - 1) / MRB;
      t->XY[lane] = make_int2(X0, Yt);
      if (lane == 0) {
        t->x0 = x0; t->y0 = y0; t->nrows = nrows; t->bx0 = bx0; t->by0 = mny;
        t->ng = ng; t->nst = fits ? max(ng, (spread + nrows + 3) / MRB + 1) : ng;
        t->q0 = q_issue; t->fits = fits ? 1 : 0; t->issued = 0; t->valid = 1;
      }
      __syncwarp();
    };
    // block j of a tile: I1 rows y0 - 4 + 5 j ... (6 rows), I2 rows by0 + 5 j ... (5 rows)
    auto issue_block = [&](const MTile* t, int j) {
      if (lane == 0) {
        const unsigned slot = q_issue % MNS;
        unsigned long long* bar = &full[slot];
        const bool fits = t->fits != 0;
        m_fence_proxy_async();            // this warp's reads of the slot are ordered before the copy engine's writes
        m_mbar_expect_tx(bar, G::kBytes1 + (fits ? G::kBytes2 : 0u));
        m_tma_load_2d(ring1 + slot * (MR1 * S1W), tm1, (t->x0 - MHALO) * C, t->y0 - 4 + MRB * j, bar);
        if (fits) m_tma_load_2d(ring2 + slot * (MRB * S2W), tm1 + 128, t->bx0 * C, t->by0 + MRB * j, bar);
        m_mbar_arrive(bar);
      }
      ++q_issue;
    };

    int cs = 0;                           // which of the two records is the tile being marched
    int ti = i0;
    // issue whatever fits into the ring: the rest of the current tile first, then the head of the next tile of this chunk
    auto try_issue = [&]() {
      const long long st_i0 = MSTAT_CLK();
      __syncwarp();
      MTile* const cur = &tiles[cs];
      MTile* const nxt = &tiles[cs ^ 1];
      int ci = cur->issued, ni = nxt->valid ? nxt->issued : 0;
      const int cn = cur->nst;
      while (q_issue - q_done < (unsigned)MNS) {
        if (ci < cn) { issue_block(cur, ci); ++ci; }
        else {
          if (!nxt->valid) {
            if (ti + 1 >= i1) break;
            tile_geo(ti + 1, nxt);
            ni = 0;
          }
          if (ni < nxt->nst) { issue_block(nxt, ni); ++ni; }
          else break;
        }
      }
      __syncwarp();
      if (lane == 0) { cur->issued = ci; if (nxt->valid) nxt->issued = ni; }
      __syncwarp();
      MSTAT_ADD(6, MSTAT_CLK() - st_i0);
    };

    if (lane == 0) { tiles[0].valid = 0; tiles[1].valid = 0; }
    __syncwarp();
    if (i0 < i1) tile_geo(i0, &tiles[0]);
    for (; ti < i1; ++ti) {
      if (ti > i0) {      // adopt the prefetched tile
        if (lane == 0) tiles[cs].valid = 0;
        cs ^= 1;
        __syncwarp();
        if (!tiles[cs].valid) tile_geo(ti, &tiles[cs]);
      }
      try_issue();
      const MTile* const cur = &tiles[cs];
      const int x0 = cur->x0, y0 = cur->y0, nrows = cur->nrows, nst = cur->nst;
      const unsigned q0 = cur->q0;
      const bool tfits = cur->fits != 0;
      MSTAT_ADD(0, 1); if (!tfits) MSTAT_ADD(1, 1);
      if (vcol >= 0 && vcol != x0) { const long long st_f0 = MSTAT_CLK(); flush_col(); MSTAT_ADD(7, MSTAT_CLK() - st_f0); }
      vcol = x0;

      // ---------------- per-tile constants of this lane
      const int x = x0 + lane;
      const bool colok = x < nx;
      const int xc = colok ? x : nx - 1;
      const float mcgx = (colok && (unsigned)(x - D.gxlo) < (unsigned)D.gxspan) ? 0.5f : 0.0f;   // x-gradient defined in this column
      const float mcfx = (colok && (unsigned)(x - D.fxlo) < (unsigned)D.fxspan) ? 0.5f : 0.0f;   // column inside the frame
      // ring row of this lane's window row at march time 0, in bytes; advanced by one row per time step.  (The lane's
      // window corner (X0, Yt) stays in the tile record: two registers less in the march.)
      int roff;
      {
        const int dl = tfits ? (cur->XY[lane].y - cur->by0) : 0;
        int r = (int)(q0 % MNS) * MRB + dl;
        r = r >= MNS * MRB ? r - MNS * MRB : r;
        roff = r * (S2W * 4);
      }
      const float* const c1 = ring1 + (lane + MHALO) * C;
      float win[5][5][C];                  // [slot][column][channel]: the register window
#pragma unroll
      for (int sl = 0; sl < 5; ++sl)
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
          for (int ch = 0; ch < C; ++ch) win[sl][i][ch] = 0.0f;

      // ---------------- the march: time tau loads window row tau into slot tau % 5; for tau >= 4 it processes image
      // row y0 + tau - 4.  Group g = tau / 5 reads blocks g and g + 1.
      int kk = 0, g = 0;
      const float* b1 = c1;
      float i1prev[C];                     // I1 of the row above (carried across the block boundary: row -1 of a block is not staged)
#pragma unroll
      for (int ch = 0; ch < C; ++ch) i1prev[ch] = 0.0f;
      const int ntau = nrows + 4;
#pragma unroll 1
      for (int tau = 0; tau < ntau; ++tau) {
        if (kk == 0) {
          const unsigned qg = q0 + g;
          const long long st_w0 = MSTAT_CLK();
          m_mbar_wait(&full[qg % MNS], (qg / MNS) & 1u);
          if (g + 1 < nst) m_mbar_wait(&full[(qg + 1) % MNS], ((qg + 1) / MNS) & 1u);
          MSTAT_ADD(4, MSTAT_CLK() - st_w0);
          b1 = c1 + (qg % MNS) * (MR1 * S1W);
        }
        const int2 xy = cur->XY[lane];       // X0, Yt
        if (tfits) {
          const float* rp = reinterpret_cast<const float*>(reinterpret_cast<const char*>(ring2) + roff + (xy.x - cur->bx0) * (C * 4));
          switch (kk) {
#define ICA_MARCH_LOAD(KK) case KK: _Pragma("unroll") for (int i = 0; i < 5; ++i) _Pragma("unroll") for (int ch = 0; ch < C; ++ch) win[KK][i][ch] = rp[i * C + ch]; break;
            ICA_MARCH_LOAD(0) ICA_MARCH_LOAD(1) ICA_MARCH_LOAD(2) ICA_MARCH_LOAD(3)
            default: _Pragma("unroll") for (int i = 0; i < 5; ++i) _Pragma("unroll") for (int ch = 0; ch < C; ++ch) win[4][i][ch] = rp[i * C + ch]; break;
#undef ICA_MARCH_LOAD
          }
          roff += S2W * 4;
          roff = roff >= MNS * MRB * S2W * 4 ? 0 : roff;
        }
        const int y = y0 + tau - 4;
        const bool yin = !frame || (y >= delta && y < ny - delta);
        // I1: row kk of block g is the pixel's own row (ring rows y0 - 4 + 5 g ...)
        const float* r1 = b1 + kk * S1W;
        float i1own[C], i1up[C];
        if (tau >= 4) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) { i1own[ch] = r1[ch]; i1up[ch] = kk != 0 ? r1[ch - S1W] : i1prev[ch]; }
        }
        if (tau >= 4 && yin) {      // (uniform) rows of the discarded frame contribute nothing
          int cx, cy; float tx, ty;
          bool pok;
          {
            const float4 q0v = *reinterpret_cast<const float4*>(&D.coef[0]);
            const float4 q1v = *reinterpret_cast<const float4*>(&D.coef[4]);
            WarpCoef coef;
            coef.d00 = q0v.x; coef.m01 = q0v.y; coef.m02 = q0v.z; coef.m10 = q0v.w;
            coef.d11 = q1v.x; coef.m12 = q1v.y; coef.m20 = q1v.z; coef.m21 = q1v.w;
            pok = project_px(coef, D.m64, xc, y, cx, cy, tx, ty);
          }
          const int dcx = cx - 1 - xy.x, dcy = cy - 1 - (xy.y + tau - 4);
          const bool fast = tfits && (unsigned)dcx < 2u && (unsigned)dcy < 2u;
          // Keys weights, zero-padded to five entries: the 4 x 4 footprint sits at (dcx, dcy) inside the 5 x 5 window.
          // The row weights are rotated on to the slots: window row j lives in slot (kk + 1 + j) % 5.
          float wx5[5], wys[5];
          {
            float w0, w1, w2v, w3;
            keys_weights(tx, w0, w1, w2v, w3);
            const bool sx = dcx != 0;
            wx5[0] = sx ? 0.0f : w0; wx5[1] = sx ? w0 : w1; wx5[2] = sx ? w1 : w2v; wx5[3] = sx ? w2v : w3; wx5[4] = sx ? w3 : 0.0f;
            keys_weights(ty, w0, w1, w2v, w3);
            // slot sl holds window row (sl - kk - 1) mod 5, whose weight is entry (that row - dcy) of (w0, w1, w2, w3, 0)
            int rot = kk + 1 + (dcy != 0 ? 1 : 0);
            rot = rot >= 5 ? rot - 5 : rot;
            float a0 = w0, a1 = w1, a2 = w2v, a3 = w3, a4 = 0.0f;    // rotate right by `rot`
            if (rot & 1) { const float t = a4; a4 = a3; a3 = a2; a2 = a1; a1 = a0; a0 = t; }
            if (rot & 2) { const float t3 = a3, t4 = a4; a4 = a2; a3 = a1; a2 = a0; a1 = t4; a0 = t3; }
            if (rot & 4) { const float t = a0; a0 = a1; a1 = a2; a2 = a3; a3 = a4; a4 = t; }
            wys[0] = a0; wys[1] = a1; wys[2] = a2; wys[3] = a3; wys[4] = a4;
          }
          // horizontal pass per slot (slots paired for the packed arithmetic), folded straight into the vertical sum
          float iw[C];
          if constexpr (C == 3) {
            float2 arg = make_float2(0.0f, 0.0f);
            float ab = 0.0f;
#pragma unroll
            for (int sp = 0; sp < 2; ++sp) {
              const int s0 = 2 * sp, s1 = 2 * sp + 1;
              float2 h0, h1, hb;
#pragma unroll
              for (int i = 0; i < 5; ++i) {
                const float2 wv = make_float2(wx5[i], wx5[i]);
                const float2 v0 = make_float2(win[s0][i][0], win[s0][i][1]);
                const float2 v1 = make_float2(win[s1][i][0], win[s1][i][1]);
                const float2 vb = make_float2(win[s0][i][2], win[s1][i][2]);
                h0 = i == 0 ? __fmul2_rn(wv, v0) : __ffma2_rn(wv, v0, h0);
                h1 = i == 0 ? __fmul2_rn(wv, v1) : __ffma2_rn(wv, v1, h1);
                hb = i == 0 ? __fmul2_rn(wv, vb) : __ffma2_rn(wv, vb, hb);
              }
              arg = __ffma2_rn(make_float2(wys[s0], wys[s0]), h0, arg);
              arg = __ffma2_rn(make_float2(wys[s1], wys[s1]), h1, arg);
              ab = fmaf(wys[s0], hb.x, ab);
              ab = fmaf(wys[s1], hb.y, ab);
            }
            {
              float2 h4; float hb4;
#pragma unroll
              for (int i = 0; i < 5; ++i) {
                const float2 wv = make_float2(wx5[i], wx5[i]);
                const float2 v4 = make_float2(win[4][i][0], win[4][i][1]);
                h4 = i == 0 ? __fmul2_rn(wv, v4) : __ffma2_rn(wv, v4, h4);
                hb4 = i == 0 ? wx5[i] * win[4][i][2] : fmaf(wx5[i], win[4][i][2], hb4);
              }
              arg = __ffma2_rn(make_float2(wys[4], wys[4]), h4, arg);
              ab = fmaf(wys[4], hb4, ab);
            }
            iw[0] = arg.x; iw[1] = arg.y; iw[2] = ab;
          } else {
            float a = 0.0f;
#pragma unroll
            for (int sp = 0; sp < 2; ++sp) {
              const int s0 = 2 * sp, s1 = 2 * sp + 1;
              float2 h;
#pragma unroll
              for (int i = 0; i < 5; ++i) {
                const float2 wv = make_float2(wx5[i], wx5[i]);
                const float2 v = make_float2(win[s0][i][0], win[s1][i][0]);
                h = i == 0 ? __fmul2_rn(wv, v) : __ffma2_rn(wv, v, h);
              }
              a = fmaf(wys[s0], h.x, a);
              a = fmaf(wys[s1], h.y, a);
            }
            float h4;
#pragma unroll
            for (int i = 0; i < 5; ++i) h4 = i == 0 ? wx5[i] * win[4][i][0] : fmaf(wx5[i], win[4][i][0], h4);
            iw[0] = fmaf(wys[4], h4, a);
          }
          // NaN footprint of skimage's warp (cval = nan): any of the 4 x 4 taps outside the image
          bool valid = pok && colok && cx >= 1 && cy >= 1 && cx + 2 <= nx - 1 && cy + 2 <= ny - 1;
#ifdef ICA_MARCH_STATS
          { const unsigned sm_ = __ballot_sync(__activemask(), !fast && colok); MSTAT_ADD(2, 1); if (sm_) MSTAT_ADD(3, __popc(sm_)); }
#endif
          if (!fast && colok) {
            // (rare) the footprint left the register window, or the tile has no staged box: sample global memory
            float wx[4], wy[4];
            keys_weights(tx, wx[0], wx[1], wx[2], wx[3]);
            keys_weights(ty, wy[0], wy[1], wy[2], wy[3]);
#pragma unroll
            for (int ch = 0; ch < C; ++ch)
              iw[ch] = pok ? m_sample_global_slow<C>(D.I2, D.pitch, nx, ny, cx, cy, ch, wx[0], wx[1], wx[2], wx[3], wy[0], wy[1], wy[2], wy[3])
                           : __int_as_float(0x7fc00000);
          }
          bool in = true;
          if (ipol) {     // IPOL-style warp domain (bi.py:144): the projected point lies in [delta, n - 1 - delta]
            const int hx = nx - 1 - delta, hy = ny - 1 - delta;
            in = pok && colok && cx >= delta && cy >= delta && (cx < hx || (cx == hx && tx == 0.0f)) && (cy < hy || (cy == hy && ty == 0.0f));
            valid = colok && (in || !ipol_nan);
          }
          const float4 fl = D.fl;          // lo, hi, lambda^2
          const bool gyrow = y >= 1 && y <= ny - 2;
          const float mgx = mcgx, mgy = gyrow ? mcfx : 0.0f;
          float sxx = 0.f, sxy = 0.f, syy = 0.f, vx = 0.f, vyv = 0.f, t2 = 0.f;
#pragma unroll
          for (int ch = 0; ch < C; ++ch) {
            float iwv = iw[ch];
            if (!ipol) iwv = fminf(fmaxf(iwv, fl.x), fl.y);          // clip (Q1)
            else iwv = in ? iwv : 0.0f;
            const float gx = mgx * (r1[ch + C] - r1[ch - C]);
            const float gy = mgy * (r1[ch + S1W] - i1up[ch]);
            const float di = valid ? iwv - i1own[ch] : 0.0f;         // non-finite -> 0 (io.py:72, 134)
            if (need_h) { sxx = fmaf(gx, gx, sxx); sxy = fmaf(gx, gy, sxy); syy = fmaf(gy, gy, syy); }
            vx = fmaf(gx, di, vx); vyv = fmaf(gy, di, vyv);
            t2 = fmaf(di, di, t2);
          }
          // robust weight and y-moments.  A gray image stands for its x3 replication
          const float rho = robust ? rho_prime(t2 * chm, fl.z, rtype) : 1.0f;
          const float scl = rho * chm;
          const float fy = (float)y;
          const float y2 = fy * fy;
          const float2 yp01 = make_float2(1.0f, fy);
          const float2 yp23 = make_float2(y2, y2 * fy);
          const float y4 = y2 * y2;
          auto acc = [&](int base, int npow, float w) {
            const float2 wv = make_float2(w, w);
            if (npow >= 2) {
              const float2 r = __ffma2_rn(wv, yp01, make_float2(vy[base], vy[base + 1]));
              vy[base] = r.x; vy[base + 1] = r.y;
            } else {
              vy[base] = fmaf(w, 1.0f, vy[base]);
            }
            if (npow == 3) vy[base + 2] = fmaf(w, y2, vy[base + 2]);
            if (npow >= 4) {
              const float2 r = __ffma2_rn(wv, yp23, make_float2(vy[base + 2], vy[base + 3]));
              vy[base + 2] = r.x; vy[base + 3] = r.y;
            }
            if (npow == 5) vy[base + 4] = fmaf(w, y4, vy[base + 4]);
          };
          if (need_h) { acc(0 * HW, HW, scl * sxx); acc(1 * HW, HW, scl * sxy); acc(2 * HW, HW, scl * syy); }
          acc(3 * HW, BWN, scl * vx); acc(3 * HW + BWN, BWN, scl * vyv);
        }
        if (tau >= 4) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) i1prev[ch] = i1own[ch];
        }
        if (++kk == MRB || tau + 1 == ntau) {
          // block g is not read again; the last group also retires the blocks it only touched
          q_done += (tau + 1 == ntau) ? (unsigned)(nst - g) : 1u;
          kk = 0; ++g;
          try_issue();
        }
      }
    }
    if (vcol >= 0) { const long long st_f0 = MSTAT_CLK(); flush_col(); MSTAT_ADD(7, MSTAT_CLK() - st_f0); }
    ++nitems;
    const long long st_e0 = MSTAT_CLK();

    // ---------------- chunk partial: [K][kYPow] doubles, warps summed in fixed order (see ica_iterate_kernel)
    const int pair = D.pair, chunk = D.chunk;
    m_consumer_sync();
    {
      const double* accs = accs0 + (nitems & 1 ? 0 : kAccSet);   // nitems was just incremented: the set of this chunk
      double* out = P.partials + ((long long)pair * P.max_chunks + chunk) * kAccStride;
      for (int i = tid; i < NENT; i += kMConsumerThreads) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < MW; ++w) sum += accs[w * NENT + i];
        out[i] = sum;
      }
    }
    myaccs += (nitems & 1) ? kAccSet : -kAccSet;     // the other set for the next chunk (zeroed by its owner at the top)
    if (tid == 0) { atomicMax(&hdr->t1, m_gtime()); *reinterpret_cast<volatile int*>(&s_done) = seq + 1; }
    MSTAT_ADD(9, MSTAT_CLK() - st_e0); MSTAT_ADD(10, 1);
  }
}

template <int C, int DH>
cudaError_t launch_march_t(const IterParams& P, int grid, cudaStream_t stream) {
  constexpr size_t smem = (size_t)MW * MGeo<C>::kWarpFloats * sizeof(float) +
                          2 * (size_t)MW * MVals<DH>::K * kYPow * sizeof(double);
  static std::atomic<unsigned long long> configured{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    cudaError_t e = cudaFuncSetAttribute(ica_march_kernel<C, DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.fetch_or(bit, std::memory_order_release);
  }
  ica_march_kernel<C, DH><<<grid, kMThreads, smem, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace

int march_tile_w() { return MTW; }
int march_tile_h() { return MTH; }
int march_chunk_unit() { return MW; }
void march_stage_boxes(int channels, int* w1, int* h1, int* w2, int* h2) {
  *w1 = channels == 3 ? MGeo<3>::S1W : MGeo<1>::S1W; *h1 = MR1;
  *w2 = channels == 3 ? MGeo<3>::S2W : MGeo<1>::S2W; *h2 = MRB;
}

cudaError_t launch_march(const IterParams& P, int channels, int dh, int grid, cudaStream_t stream) {
  if (channels == 3) {
    if (dh == 4) return launch_march_t<3, 4>(P, grid, stream);
    if (dh == 2) return launch_march_t<3, 2>(P, grid, stream);
    return launch_march_t<3, 0>(P, grid, stream);
  }
  if (dh == 4) return launch_march_t<1, 4>(P, grid, stream);
  if (dh == 2) return launch_march_t<1, 2>(P, grid, stream);
  return launch_march_t<1, 0>(P, grid, stream);
}

}  // namespace ica
