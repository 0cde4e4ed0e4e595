// K2 + K3: the fused per-iteration kernel of the inverse compositional loop (sm_100a).
//
// One launch = one iteration of EVERY still-active image pair of the batch, each at its own
// scale.  Replaces, per iteration (src/inverse_compositional_algorithm.py:109-131, 225-259):
//   bi.bicubic_interpolation_skimage  (warp of I2 by p, Catmull-Rom, NaN footprint, clip)
//   DI = Iw - I1
//   io.robust_error_function          (rho'(sum_c DI_c^2))
//   io.independent_vector[_robust]    (b)
//   de.hessian_robust / de.hessian    (H; gradients, Jacobian and steepest-descent images are
//                                      recomputed per pixel, never stored: ica.py:81-100)
// and, in ica_solve_kernel (one block per pair, launched after it):
//   de.inverse_hessian, io.parametric_solve, tr.update_transform, the lambda schedule, the
//   stopping rule and zm.zoom_in_parameters at a scale change.
//
// Execution model
//   * The work list of a launch (built by ica_schedule_kernel for the first iteration, by the last block of
//     ica_solve_kernel afterwards): pair b at a level with T tiles of 64 x 11 pixels contributes min(T, max_chunks)
//     chunks of consecutive tiles.  The chunk decomposition of a pair never depends on what else is in the batch.
//   * ica_iterate_kernel is PERSISTENT (2 CTAs per SM, 12 warps each at 80 registers) and WARP-SPECIALISED.  CTAs claim
//     chunks through an atomic counter.  Warp 11 is the producer: per tile it projects the tile corners and stages the
//     I1 patch (+halo) and the window of I2 the tile can reach with two tiled TMA copies (cp.async.bulk.tensor.2d ->
//     UTMALDG, tensor maps per pair / level / image) that complete on a "full" mbarrier; the copy engine fills what
//     lies outside the image (NaN for I2 = skimage's cval, so the NaN footprint falls out of the arithmetic; 0 for
//     I1).  Warps 0-10 consume the stage and release it through an "empty" mbarrier.  Two stages: the copies of tile
//     t+1 (even of the next chunk) overlap the arithmetic of tile t; no block-wide barrier in the tile loop.
//   * A consumer warp owns one image row of the tile (two pixels per lane, packed fp32).  Lanes accumulate the
//     x-moments of that row in fp32 and keep them across the consecutive tiles of the row; once per row segment they
//     are transposed through shared memory so that moment k lands on lane k, which folds in y^b in fp64.  Chunk
//     partials go to fixed slots; ica_solve_kernel sums them in a fixed order (deterministic, batch-invariant).
// Bound named by the north star: HBM (read I1 once + I2 once per pixel-iteration = 2*C*4 bytes); measured: DRAM
// traffic = 1.01x that, the kernel is limited by instruction issue and shared-memory wavefronts.  No tensor cores.
#include <atomic>
#include "ica_device.cuh"
#include "ica_transform.cuh"
#include "ica_iterate.cuh"

namespace ica {

namespace {

// Per-CTA timeline stamps and wait accounting (tools/timeline.py) are compiled in only with -DICA_TIMELINE=1
// (ICA_TIMELINE=1 python -m inverse_compositional_algorithm_b200.build): they cost registers in the per-pixel loop.
#ifdef ICA_TIMELINE
constexpr bool kTimeline = true;
#else
constexpr bool kTimeline = false;
#endif
#ifndef ICA_BLOCKS_PER_SM
#define ICA_BLOCKS_PER_SM 2
#endif
constexpr int kBlocksPerSM = ICA_BLOCKS_PER_SM;      // 2 CTAs x 12 warps per SM at 80 registers per thread (16 warps at 64 registers measured 9 % slower)
#ifndef ICA_CONSUMER_WARPS
#define ICA_CONSUMER_WARPS 11
#endif
constexpr int kConsumerWarps = ICA_CONSUMER_WARPS;   // + 1 producer warp = 12 warps: warps are allocated in groups of 4
constexpr int kRowsPerWarp = 1;      // rows of a tile per consumer warp
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kThreads = kConsumerThreads + 32;   // + one producer warp
constexpr int kChunkRing = 8;       // chunk records in flight per CTA (> the deepest pipeline + 2)
constexpr int kItemPairBits = 20;   // a work item = pair | chunk << 20 (one read tells the producer both)
constexpr int TW = 64;            // tile width  (2 pixels per lane: x0+lane, x0+32+lane)
constexpr int TH = kRowsPerWarp * kConsumerWarps;   // tile height (row y0 + warp + rr * kConsumerWarps for warp, rr)
constexpr int HALO = 4;           // the I1 patch starts at x0-4: the inner coordinate of a TMA box must be a multiple of 16 bytes
constexpr int S1ROWS = TH + 2;
#ifndef ICA_BH
#define ICA_BH 24
#endif
#ifndef ICA_PRODUCER_SLEEP
#define ICA_PRODUCER_SLEEP 200     // ns between the producer's polls of an `empty` barrier (tuning hook)
#endif
#ifndef ICA_S2W_RGB
#define ICA_S2W_RGB 256           // (tuning hook: 224 = 74 pixels measured no faster)
#endif
#ifndef ICA_STAGES_RGB
#define ICA_STAGES_RGB 2
#endif
#ifndef ICA_STAGES_GRAY
#define ICA_STAGES_GRAY 4
#endif
constexpr int BH_MAX = ICA_BH;    // rows of the staged I2 window; taller windows (strong rotation / zoom) take the global-memory path

template <int DH> struct RowVals { static constexpr int K = 3 * (DH + 1) + 2 * (DH / 2 + 1); };

// One stage = the TMA boxes of a tile: the I2 window (S2W floats x BH_MAX rows) and the I1 patch (S1W floats x
// S1ROWS rows).  A box is at most 256 elements wide, its rows are multiples of 16 bytes, and S2W == 0 (mod 32
// banks), so lanes of a warp that sit on different window rows never collide.
template <int C> struct Stage {
  static constexpr int S2W = C == 3 ? ICA_S2W_RGB : 96;    // floats per window row
  static constexpr int BWPX = S2W / C;                     // window width in pixels (85 RGB, 96 gray)
  static constexpr int S1W = (TW + 2 * HALO) * C;          // 216 (RGB) / 72 (gray) floats per patch row
  static constexpr int kFloats = (BH_MAX * S2W + S1ROWS * S1W + 31) / 32 * 32;   // I2 window, then I1 patch; 128-byte multiple
  // staged tiles in flight per CTA: RGB tiles are compute-heavy (a third stage measured no gain); gray tiles are consumed
  // in ~2 us, where two stages leave the consumers waiting for the producer 11-17 % of the time
  static constexpr int kStages = C == 3 ? ICA_STAGES_RGB : ICA_STAGES_GRAY;
};

// Everything a consumer needs to know about a staged tile (written by the producer), grouped so that a consumer
// fetches it with a few 128-bit shared loads where it needs it (the register file is the scarce resource of this
// kernel: nothing tile-constant is kept in registers across the tap loads).  The warp coefficients are stored
// duplicated, (c, c): they are operands of packed fp32 arithmetic on the lane's two pixels.
struct __align__(16) TileCtl {
  float c8[8];            // d00, m01, m02, m10, d11, m12, m20, m21 of WarpCoef (broadcast operands of the packed arithmetic)
  int4 geo;               // x0, y0, nx, ny
  int4 win;               // bx0, by0, bw - 3, bh - 3: the staged I2 window as corner + unsigned limits (0, 0: no window)
  int4 msk;               // gxlo, gxspan: columns with an in-frame x-gradient; fxlo, fxspan: columns inside the frame
  int4 flg;               // need_h, last, stop, pitch
  float4 fl;              // lo, hi: clip range of I2 at this level; lambda^2; unused
  double m64[9];          // warp matrix in fp64 (tie-break path of project_px)
  int pair, chunk, nch, scale;
  const float* I2;        // global image, for the pixels whose taps leave the staged window
};

// ---------------------------------------------------------------- mbarrier / bulk-copy PTX
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LAB_DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "LAB_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 2-D tiled TMA load: the box described by the tensor map, with its corner at (c0 floats, c1 rows); elements outside
// the image are filled by the copy engine (NaN for I2 = skimage's cval, 0 for I1), negative corners included.
// the producer is usually a tile ahead: poll politely so that its spinning does not take issue slots from the consumers
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long* bar, unsigned parity) {
  for (;;) {
    unsigned done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(ICA_PRODUCER_SLEEP);
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, int c0, int c1, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// 128-bit shared loads that stay where they are written (volatile asm): tile constants are fetched at the point of use
__device__ __forceinline__ int4 lds_i4(const void* p) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ float4 lds_f4(const void* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ void fence_tensormap_acquire(const void* tmap) {
  asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define ICA_STAMP(slot) do { if (kTimeline && P.dbg_time && it == 0 && tid == 0) P.dbg_time[blockIdx.x * 16 + (slot)] = gtime(); } while (0)
// barrier among the consumer threads only (the producer warp never joins it)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory"); }

// Transposing warp reduction: every lane holds NP partial values v[0..NP-1]; on return lane l holds in v[0] the sum
// over all 32 lanes of value index (l >> log2(32 / NP)), NP in {8, 16, 32}.  Fixed order -> deterministic.
template <int NP>
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[NP], int lane) {
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

// Jacobian entry k as a monomial, in registers (same table as ica_transform.cuh: jacobian_monomials)
__device__ __forceinline__ void mono_of(int ttype, int k, Mono& jx, Mono& jy) {
  Mono tx[ICA_MAX_PARAMS], ty[ICA_MAX_PARAMS];
  jacobian_monomials(ttype, tx, ty);
  jx = tx[0]; jy = ty[0];
#pragma unroll
  for (int i = 1; i < ICA_MAX_PARAMS; ++i) if (i == k) { jx = tx[i]; jy = ty[i]; }
}

// Generic (rare) path kept out of line so that its address arithmetic is not hoisted into the fast path.
template <int C>
__device__ __noinline__ float sample_global_slow(const float* __restrict__ img, int pitch, int nx, int ny, int cx,
                                                 int cy, int ch, float wx0, float wx1, float wx2, float wx3,
                                                 float wy0, float wy1, float wy2, float wy3) {
  const float wx[4] = {wx0, wx1, wx2, wx3}, wy[4] = {wy0, wy1, wy2, wy3};
  return sample_global<C>(img, pitch, nx, ny, cx, cy, ch, wx, wy);
}

// ============================================================ producer warp
template <int C>
__device__ __forceinline__ void producer_loop(const IterParams& P, float* stages, int stage_floats,
                                              unsigned long long* full, unsigned long long* empty, TileCtl* tctl,
                                              double* pm64, int4* cinfo, int par, int lane) {
  constexpr int S1W = Stage<C>::S1W, S2W = Stage<C>::S2W, kStages = Stage<C>::kStages;
  unsigned k = 0;   // tiles staged so far by this CTA
  unsigned cseq = 0;   // chunks started so far by this CTA
  SchedHdr* const hdr = P.hdr + par;
  const int* const item_pair = P.item_pair + (long long)__ldcg(&hdr->list) * P.B * P.max_chunks;
  bool first_item = !P.fused;
  const bool pdbg = kTimeline && P.dbg_time != nullptr && lane == 0;   // profiling hook: cycles spent fetching work / waiting for a free stage
  long long pd_fetch = 0, pd_empty = 0, pd_proj = 0, pd_ctl = 0, pd_issue = 0;
  for (;;) {
    // dynamic work distribution: chunks are handed out by an atomic counter (reset by the scheduler)
    const long long pf0 = pdbg ? clock64() : 0;
    int item = (int)blockIdx.x;        // first item of this CTA: static (the dynamic counter starts at the grid size)
    if (!first_item) {
      if (lane == 0) item = atomicAdd(&hdr->counter, 1);
      item = __shfl_sync(0xffffffffu, item, 0);
    }
    first_item = false;
    // (the total is re-read: with the fused solve a CTA that starts late may find the list of the NEXT iteration here)
    if (item >= __ldcg(&hdr->total)) break;
    const int packed = __ldcg(item_pair + item);
    const int pair = packed & ((1 << kItemPairBits) - 1), chunk = packed >> kItemPairBits;
    // the pair's state was written by the solve of the previous iteration, possibly during this very launch: read
    // it from L2, never through a (per-SM, possibly stale) L1 line
    const PairState* stp = P.state + pair;
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
    const float lo = key_float(mm.lo), hi = key_float(mm.hi);
    const float lambda2 = (float)(st_lambda * st_lambda);
    const int need_h = (P.robust_loop || st_iter == 0) ? 1 : 0;
    const LevelDesc L = P.lv[s];
    const int nx = L.nx, ny = L.ny, pitch = L.pitch;
    const float* __restrict__ I2 = s == 0 ? P.I2_0 + (long long)pair * P.in_stride
                                          : P.pyr2 + (long long)pair * P.pyr_stride + L.offset;
    // tensor maps of this pair's level: [pair][scale][0 = I1 (zero fill), 1 = I2 (NaN fill)], 128 bytes each
    const char* tm1 = static_cast<const char*>(P.tmaps) + ((long long)(pair * P.nscales + s) * 2) * 128;
    const char* tm2 = tm1 + 128;
    // the maps live in global memory and are rewritten (cudaMemcpyAsync) when the level-0 images move or a new plan
    // reuses the allocation: acquire them through the tensormap proxy before the first copy that uses them
    if (lane == 0) { fence_tensormap_acquire(tm1); fence_tensormap_acquire(tm2); }
    int t_first;
    const int ntiles = band_tiles(L, P.shard_rank, P.shard_n, &t_first);
    const int nch = ntiles < P.max_chunks ? ntiles : P.max_chunks;
    const int t_begin = t_first + (int)((long long)chunk * ntiles / nch);
    const int t_end = t_first + (int)((long long)(chunk + 1) * ntiles / nch);
    // what the consumers need at the chunk's epilogue, in a ring that outlives the stages (the producer is at most
    // kStages tiles, hence chunks, ahead): published by the `full` barrier of the chunk's first tile
    if (lane == 0) cinfo[cseq & (kChunkRing - 1)] = make_int4(pair, chunk, nch, 0);
    ++cseq;

    if (kTimeline && P.dbg_time && k == 0 && lane == 0) P.dbg_time[blockIdx.x * 16 + 9] = gtime();
    if (pdbg) pd_fetch += clock64() - pf0;
    for (int tile = t_begin; tile < t_end; ++tile, ++k) {
      const int sidx = k % kStages;
      const unsigned use = k / kStages;
      const long long pe0 = pdbg ? clock64() : 0;
      if (use >= 1) mbar_wait_backoff(&empty[sidx], (use - 1) & 1);   // consumers released the previous use
      const long long pt1 = pdbg ? clock64() : 0;
      if (pdbg) pd_empty += pt1 - pe0;
      float* s2 = stages + sidx * stage_floats;
      float* s1 = s2 + BH_MAX * S2W;
      unsigned long long* bar = &full[sidx];
      TileCtl& tc = tctl[sidx];
      const int x0 = (tile % L.tiles_x) * TW;
      const int y0 = (tile / L.tiles_x) * TH;
      // window of I2 reachable from this tile: project the four corners of its in-image part
      const int xe_ = min(x0 + TW, nx) - 1, ye_ = min(y0 + TH, ny) - 1;
      const int px = (lane & 1) ? xe_ : x0, py = (lane & 2) ? ye_ : y0;
      int cx, cy; float tx, ty;
      const bool ok = project_px(coef, pm64, px, py, cx, cy, tx, ty);
      int mnx = cx, mxx = cx, mny = cy, mxy = cy, okall = ok ? 1 : 0;
#pragma unroll
      for (int o = 1; o < 4; o <<= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        okall &= __shfl_xor_sync(0xffffffffu, okall, o);
      }
      const int bx0 = ((mnx - 2) >> 2) << 2;                   // floor to a multiple of 4 pixels (16-byte box corner)
      const int by0 = mny - 2;                                 // one pixel of margin around the corners' taps
      const int bw = mxx + 3 - bx0 + 1;
      const int bh = mxy + 3 - by0 + 1;
      const bool fits = okall && bw <= Stage<C>::BWPX && bh <= BH_MAX && bw > 0 && bh > 0;
      const long long pt2 = pdbg ? clock64() : 0;
      if (tile < t_begin + kStages) {   // first use of this stage's control block by the chunk: per-chunk constants
        if (lane < 9) tc.m64[lane] = pm64[lane];
        if (lane == 0) {
          tc.c8[0] = coef.d00; tc.c8[1] = coef.m01; tc.c8[2] = coef.m02; tc.c8[3] = coef.m10;
          tc.c8[4] = coef.d11; tc.c8[5] = coef.m12; tc.c8[6] = coef.m20; tc.c8[7] = coef.m21;
          tc.fl = make_float4(lo, hi, lambda2, 0.0f);
          // columns inside the discarded frame (ica.py:85-93) and, of those, the ones with a central x-difference
          const int fxlo = P.frame ? P.delta : 0;
          const int fxspan = max(0, nx - 2 * fxlo);
          const int gxlo = max(fxlo, 1), gxhi = min(fxlo + fxspan, nx - 1);
          tc.msk = make_int4(gxlo, max(0, gxhi - gxlo), fxlo, fxspan);
          tc.pair = pair; tc.chunk = chunk; tc.nch = nch; tc.scale = s; tc.I2 = I2;
        }
      }
      if (lane == 0) {
        tc.geo = make_int4(x0, y0, nx, ny);
        tc.win = fits ? make_int4(bx0, by0, bw - 3, bh - 3) : make_int4(0, 0, 0, 0);
        tc.flg = make_int4(need_h, tile + 1 == t_end ? 1 : 0, 0, pitch);
      }
      const long long pt3 = pdbg ? clock64() : 0;
      __syncwarp();                       // the control block is complete before the arrival below
      if (lane == 0) {
        // two tiled TMA copies per tile; pixels outside the image are filled by the copy engine
        constexpr unsigned kBytes1 = S1ROWS * S1W * 4u, kBytes2 = BH_MAX * S2W * 4u;
        mbar_expect_tx(bar, kBytes1 + (fits ? kBytes2 : 0u));
        tma_load_2d(s1, tm1, (x0 - HALO) * C, y0 - 1, bar);
        if (fits) tma_load_2d(s2, tm2, bx0 * C, by0, bar);
        mbar_arrive(bar);                 // release; the phase completes when the copied bytes have landed too
      }
      if (pdbg) { const long long pt4 = clock64(); pd_proj += pt2 - pt1; pd_ctl += pt3 - pt2; pd_issue += pt4 - pt3; }
    }
  }
  if (pdbg) {
    long long* d = P.dbg_time + blockIdx.x * 16;
    d[7] = pd_fetch; d[8] = pd_empty; d[10] = pd_proj; d[11] = pd_ctl; d[12] = pd_issue;
  }
  // no more work: hand the consumers a stop marker through the next stage
  {
    const int sidx = k % kStages;
    const unsigned use = k / kStages;
    if (use >= 1) mbar_wait(&empty[sidx], (use - 1) & 1);
    if (lane == 0) { tctl[sidx].flg = make_int4(0, 1, 1, 0); mbar_arrive(&full[sidx]); }
  }
}

// ============================================================ scheduling and the per-pair solve
// Both run either in the stand-alone kernels (ica_schedule_kernel, ica_solve_kernel: host-driven loop, row-sharded
// mode, parity hooks) or, fused, inside ica_iterate_kernel by the consumer warps of the CTA that finishes a pair's last
// chunk.  BlockSync abstracts the barrier: the whole block, or the consumer warps only (named barrier 1).
struct SyncBlock { __device__ __forceinline__ void operator()() const { __syncthreads(); } };
struct SyncConsumers { __device__ __forceinline__ void operator()() const { consumer_sync(); } };

// Work list of the NEXT iteration into the buffer `np`: chunk_start[b] = exclusive prefix sum of chunks per pair,
// chunk_start[B] = total, item_pair[i] = pair of work item i; publishes the number of unfinished pairs, advances the
// iteration counter (which selects the buffer the next launch reads) and sets the condition of the CUDA-graph while
// node.  Executed by `nthr` threads (a multiple of 32, <= 1024) that share `sync`.
template <typename Sync>
__device__ void schedule_block(const IterParams& P, int* s_warp, int* s_scal, bool first, int tid, int nthr, Sync sync) {
  const int lane = tid & 31, warp = tid >> 5;
  const int nwarp = nthr >> 5;
  const int B = P.B;
  const int cnt_old = first ? -1 : __ldcg(P.loop_count);
  const int np = (cnt_old + 1) & 1;
  // No pair changed scale or finished in this iteration: the next work list is the current one.  Re-announce it under the
  // next header instead of rebuilding it (not with the fused solve, whose late CTAs need the two lists to alternate).
  if (!first && !P.fused && __ldcg(P.sched_dirty) == 0) {
    if (tid == 0) {
      const SchedHdr* ho = P.hdr + (cnt_old & 1);
      SchedHdr* hn = P.hdr + np;
      hn->total = ho->total; hn->npairs = ho->npairs; hn->list = ho->list; hn->counter = P.grid_ctas;
      hn->t0 = 0x7fffffffffffffffll; hn->t1 = 0;
      const long long t0 = __ldcg(&ho->t0), t1 = __ldcg(&ho->t1);
      if (t1 > t0 && t1 > 0) { P.kernel_ns[0] += t1 - t0; P.kernel_ns[1] += 1; }
      const int cnt = cnt_old + 1;
      *P.solve_ticket = 0;
      __threadfence();
      *P.loop_count = cnt;
      // (the number of unfinished pairs did not change either; a rank of the row-sharded mode may own no work item yet go on)
      if (P.cond_handle) cudaGraphSetConditional(P.cond_handle, (__ldcg(P.n_active) > 0 && cnt < P.max_launches) ? 1u : 0u);
    }
    return;
  }
  int* const chunk_start = P.chunk_start + (long long)np * (B + 1);
  int* const item_pair = P.item_pair + (long long)np * B * P.max_chunks;
  if (tid == 0) { s_scal[0] = 0; s_scal[1] = 0; s_scal[2] = 0; }   // carry, active pairs, pairs with work
  sync();
  for (int base = 0; base < B; base += nthr) {
    const int b = base + tid;
    int c = 0;
    bool act = false;   // a rank whose band is empty at a coarse level still counts the pair as unfinished
    if (b < B) {
      const int s = __ldcg(&P.state[b].scale);
      if (s >= 0) { int tf; const int nt = band_tiles(P.lv[s], P.shard_rank, P.shard_n, &tf); c = chunk_count(nt, P.max_chunks, P.chunk_unit, P.chunk_m); }
      act = s >= 0;
    }
    const unsigned actmask = __ballot_sync(0xffffffffu, act);
    const unsigned workmask = __ballot_sync(0xffffffffu, c > 0);
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    sync();
    if (warp == 0) {
      const int w = lane < nwarp ? s_warp[lane] : 0;
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
      s_warp[lane] = wi - w;   // exclusive
    }
    sync();
    const int excl = s_scal[0] + s_warp[warp] + incl - c;
    if (b < B) chunk_start[b] = excl;
    // the items of a pair are written by the whole warp (up to 256-1024 per pair: one thread would take microseconds)
#pragma unroll 1
    for (int src = 0; src < 32; ++src) {
      const int cb = __shfl_sync(0xffffffffu, c, src), eb = __shfl_sync(0xffffffffu, excl, src);
      for (int i = lane; i < cb; i += 32) item_pair[eb + i] = (base + (warp << 5) + src) | (i << kItemPairBits);   // (pair, chunk)
    }
    if (lane == 0 && actmask) atomicAdd(&s_scal[1], __popc(actmask));
    if (lane == 0 && workmask) atomicAdd(&s_scal[2], __popc(workmask));
    sync();
    if (tid == nthr - 1) s_scal[0] = excl + c;
    sync();
  }
  __threadfence();     // the list is complete before the header and the iteration counter announce it
  sync();
  if (tid == 0) {
    chunk_start[B] = s_scal[0];
    *P.n_active = s_scal[1];
    SchedHdr* hn = P.hdr + np;
    hn->total = s_scal[0]; hn->npairs = s_scal[2]; hn->list = np;
    *P.sched_dirty = 0;
    // CTA i takes item i first (no atomic on its critical path), then claims dynamically.  Not with the fused solve: a CTA
    // of the previous launch that starts late may already work on this list, so there every item is claimed atomically
    hn->counter = P.fused ? 0 : P.grid_ctas;
    hn->t0 = 0x7fffffffffffffffll; hn->t1 = 0;
    // device-side bookkeeping of the loop: iterations done, time spent in the streaming phase of the iterate kernel
    int cnt = 0;
    if (first) { P.kernel_ns[0] = 0; P.kernel_ns[1] = 0; P.loop_count[1] = 0; }
    else {
      cnt = cnt_old + 1;
      const SchedHdr* ho = P.hdr + (cnt_old & 1);
      const long long t0 = __ldcg(&ho->t0), t1 = __ldcg(&ho->t1);
      if (t1 > t0 && t1 > 0) { P.kernel_ns[0] += t1 - t0; P.kernel_ns[1] += 1; }
    }
    *P.solve_ticket = 0;
    __threadfence();
    *P.loop_count = cnt;
    // the condition of the CUDA-graph while node: no host round trip until every pair has converged
    if (!first && P.cond_handle) cudaGraphSetConditional(P.cond_handle, (s_scal[1] > 0 && cnt < P.max_launches) ? 1u : 0u);
  }
}

__global__ void __launch_bounds__(1024) ica_schedule_kernel(const __grid_constant__ IterParams P) {
  __shared__ int s_warp[32];
  __shared__ int s_scal[4];
  schedule_block(P, s_warp, s_scal, true, (int)threadIdx.x, (int)blockDim.x, SyncBlock());
}

// K3: sums a pair's chunk partials in a fixed order, assembles H and b, de.inverse_hessian + io.parametric_solve +
// tr.update_transform, the lambda schedule, the stopping rule and zm.zoom_in_parameters at a scale change
// (ica.py:223-259, 102-131).  NW warps of the calling block take part (tid in [0, 32 NW)); s_part is scratch for
// NW x NENT doubles.  Returns false when the caller must not go on to the scheduling step (row-sharded mode 1).
struct SolveShared {
  double mom[kAccStride];
  double aug[ICA_MAX_PARAMS][2 * ICA_MAX_PARAMS + 1];
  double vec[2 * ICA_MAX_PARAMS];
  PairState st;                    // the pair's state is staged here and written back once
  int warp_scan[32];
  int scal[4];
  unsigned int ticket;
};

// Row-sharded mode, exchange through peer memory (one process per GPU, buffers mapped with CUDA IPC over NVLink): every
// rank stores its band's moment sums into its slot of EVERY rank's buffer, raises the slot's sequence flag, waits until
// all slots of its own buffer carry this iteration's sequence number and adds them in rank order -- identical sums, and
// therefore identical parameters, on all ranks with no host round trip and no collective launch.  Two slot sets
// alternate by the parity of the sequence number: a rank can only start iteration i+2 after every rank has published
// i+1, i.e. after every rank has finished reading iteration i.
template <typename Sync>
__device__ void exchange_moments(const IterParams& P, int pair, int tid, int nthr, double* s_mom, int nent, Sync sync) {
  const int W = P.x_world, me = P.x_rank;
  const unsigned long long seq = __ldcg(P.x_seq_base) + (unsigned long long)__ldcg(P.loop_count) + 1ull;
  const long long par = (long long)(seq & 1ull);
  const long long t0 = gtime();
  for (int e = tid; e < W * nent; e += nthr) {
    const int r = e / nent, k = e - r * nent;
    double* dst = P.x_peers[r] + ((par * W + me) * P.B + pair) * kXSlot;
    dst[k] = s_mom[k];
  }
  __threadfence_system();     // the sums are visible system-wide before the flags
  sync();
  if (tid < W) {
    volatile unsigned long long* f =
        reinterpret_cast<volatile unsigned long long*>(P.x_peers[tid] + ((par * W + me) * P.B + pair) * kXSlot + (kXSlot - 1));
    *f = seq;
  }
  const double* mine = P.x_peers[me];
  if (tid < W) {
    const volatile unsigned long long* f =
        reinterpret_cast<const volatile unsigned long long*>(mine + ((par * W + tid) * P.B + pair) * kXSlot + (kXSlot - 1));
    while (*f != seq) {
      if (gtime() - t0 > 4000000000ll) { *P.x_error = 1; break; }     // 4 s: a peer is gone; fail instead of hanging the GPU
    }
  }
  __threadfence_system();
  sync();
  if (tid < nent) {
    double a = 0.0;
    for (int r = 0; r < W; ++r) a += *reinterpret_cast<const volatile double*>(mine + ((par * W + r) * P.B + pair) * kXSlot + tid);
    s_mom[tid] = a;
  }
  sync();
  if (tid == 0 && P.x_ns) { atomicAdd(reinterpret_cast<unsigned long long*>(&P.x_ns[0]), (unsigned long long)(gtime() - t0)); atomicAdd(reinterpret_cast<unsigned long long*>(&P.x_ns[1]), 1ull); }
}

template <int DH, int NW, typename Sync>
__device__ bool solve_pair(const IterParams& P, int pair, int tid, double* s_part, SolveShared& sh, Sync sync, bool stamp) {
  constexpr int K = RowVals<DH>::K;
  constexpr int HW = DH + 1;
  constexpr int NENT = K * kYPow;
  constexpr int kStateWords = (int)(sizeof(PairState) / 8);
  static_assert(sizeof(PairState) % 8 == 0, "PairState is copied as 8-byte words");
  static_assert(NW * 32 >= kAccStride && NW * 32 >= kStateWords, "one thread per moment / state word");
  double* const s_mom = sh.mom;
  double (*s_aug)[2 * ICA_MAX_PARAMS + 1] = sh.aug;
  double* const s_vec = sh.vec;
  const int lane = tid & 31, warp = tid >> 5;
#define SOLVE_STAMP(slot) do { if (kTimeline && stamp && P.dbg_time && tid == 0) P.dbg_time[(long long)P.dbg_row * 16 + (slot)] = gtime(); } while (0)
  SOLVE_STAMP(0);
  if (tid < kStateWords)
    reinterpret_cast<unsigned long long*>(&sh.st)[tid] = __ldcg(reinterpret_cast<const unsigned long long*>(&P.state[pair]) + tid);
  sync();
  SOLVE_STAMP(1);
  PairState& st = sh.st;
  const int s = st.scale;
  const bool robust = P.robust_loop != 0;
  if (s >= 0) {   // the pair took part in the iteration that just ran
    const LevelDesc L = P.lv[s];
    const int nx = L.nx, ny = L.ny;
    int t_first;
    const int ntiles = band_tiles(L, P.shard_rank, P.shard_n, &t_first);
    const int nch = chunk_count(ntiles, P.max_chunks, P.chunk_unit, P.chunk_m);
    const bool need_h = P.robust_loop || st.iter == 0;
    const int ttype = st.ttype;
    const int n = nparams_of(ttype);
    // warp 0 fetches its rows of the assembly table now; the latency hides behind the partial sums
    AsmEntry asm_e[3];
    if (warp == 0) {
#pragma unroll
      for (int m = 0; m < 3; ++m) { const int e = lane + 32 * m; if (e < 72) asm_e[m] = P.asm_tab[ttype * 72 + e]; }
    }
    if (P.solve_mode == 2) {   // row-sharded: the moments were summed over ranks by the caller's allreduce
      if (tid < NENT) s_mom[tid] = P.ext_moments[(long long)pair * kAccStride + tid];
      sync();
    } else {
      // fixed summation order: warp w sums its contiguous range of chunks, then warps in order
      const int c0 = (int)((long long)warp * nch / NW), c1 = (int)((long long)(warp + 1) * nch / NW);
      const double* src = P.partials + (long long)pair * P.max_chunks * kAccStride;
      constexpr int NJ = (NENT + 31) / 32;
      constexpr int UN = 8;                                    // chunks in flight per lane
      double sum[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) sum[j] = 0.0;
      int c = c0;
      for (; c + UN <= c1; c += UN) {
        double a[UN][NJ];
#pragma unroll
        for (int u = 0; u < UN; ++u)
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const int e = lane + 32 * j;
            a[u][j] = e < NENT ? __ldcg(src + (long long)(c + u) * kAccStride + e) : 0.0;
          }
#pragma unroll
        for (int u = 0; u < UN; ++u)
#pragma unroll
          for (int j = 0; j < NJ; ++j) sum[j] += a[u][j];
      }
      for (; c < c1; ++c) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) { const int e = lane + 32 * j; if (e < NENT) sum[j] += __ldcg(src + (long long)c * kAccStride + e); }
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) { const int e = lane + 32 * j; if (e < NENT) s_part[warp * NENT + e] = sum[j]; }
      sync();
      if (tid < NENT) {
        double tsum = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) tsum += s_part[w * NENT + tid];
        // quadratic loop after the first iteration of a scale: the H moments were not gathered
        s_mom[tid] = (!need_h && tid < 3 * HW * kYPow) ? 0.0 : tsum;
      }
      sync();
    }
    SOLVE_STAMP(2);
    if (P.solve_mode == 3) exchange_moments(P, pair, tid, NW * 32, s_mom, NENT, sync);   // row-sharded: add over ranks
    if (P.solve_mode == 1) {   // row-sharded: publish this rank's moment sums and stop
      if (tid < kAccStride) P.ext_moments[(long long)pair * kAccStride + tid] = tid < NENT ? s_mom[tid] : 0.0;
      return false;
    }
    if (warp == 0) {   // the n x n part is one warp's job
      // assemble H (n x n) and b (n): every entry is a fixed +-1 combination of at most 4 moments
      // (table built on the host from the Jacobian monomials, ica_transform.cuh: assemble_system)
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const int e = lane + 32 * m;
        if (e < 72) {
          const int kk = e < 64 ? e >> 3 : e - 64, l = e < 64 ? e & 7 : 0;
          double sum = 0.0;
          if (kk < n && l < n) {
#pragma unroll
            for (int q = 0; q < 4; ++q) if (asm_e[m].coef[q] != 0.0f) sum += (double)asm_e[m].coef[q] * s_mom[asm_e[m].idx[q]];
          }
          if (e < 64) s_aug[kk][l] = (kk < n && l < n) ? sum : (kk == l ? 1.0 : 0.0);   // identity padding up to 8 x 8
          else s_vec[kk] = kk < n ? sum : 0.0;
        }
      }
      __syncwarp();
      SOLVE_STAMP(3);
      if (P.dbg_Hb) {  // parity hook (ica_hessian_b_host): export, leave the state untouched
        for (int e = lane; e < n * n; e += 32) P.dbg_Hb[e] = s_aug[e / n][e % n];
        if (lane < n) P.dbg_Hb[64 + lane] = s_vec[lane];
      } else {
        // de.inverse_hessian: Gauss-Jordan with partial pivoting on the 8 x 16 matrix [H (+) I | I], lane j
        // keeps column j in registers: pivot search and row swaps are register-local, each step needs one
        // broadcast of the pivot column (element by element the arithmetic of ica_transform.cuh:
        // inverse_hessian; the identity padding leaves the n x n block untouched).  Zero matrix when a pivot
        // is exactly zero (np.linalg.LinAlgError branch, derivatives.py:127-129).
        if (need_h) {
          const int j = lane & 15;
          double a[ICA_MAX_PARAMS];
#pragma unroll
          for (int r = 0; r < ICA_MAX_PARAMS; ++r) a[r] = j < ICA_MAX_PARAMS ? s_aug[r][j] : (r == j - ICA_MAX_PARAMS ? 1.0 : 0.0);
          bool singular = false;
#pragma unroll
          for (int kk = 0; kk < ICA_MAX_PARAMS; ++kk) {
            // the lane that holds column kk finds the pivot row (largest |a[r][kk]|, r >= kk, first on ties)
            int piv = kk;
            double best = fabs(a[kk]);
#pragma unroll
            for (int r = kk + 1; r < ICA_MAX_PARAMS; ++r) { const double vv = fabs(a[r]); if (vv > best) { best = vv; piv = r; } }
            piv = __shfl_sync(0xffffffffu, piv, kk);
            best = __shfl_sync(0xffffffffu, best, kk);
            if (!(best > 0.0)) { singular = true; break; }
#pragma unroll
            for (int r = kk + 1; r < ICA_MAX_PARAMS; ++r) if (r == piv) { const double t = a[kk]; a[kk] = a[r]; a[r] = t; }
            const double pivval = __shfl_sync(0xffffffffu, a[kk], kk);
            const double inv = 1.0 / pivval;
            a[kk] *= inv;
#pragma unroll
            for (int r = 0; r < ICA_MAX_PARAMS; ++r) {
              if (r == kk) continue;
              const double f = __shfl_sync(0xffffffffu, a[r], kk);   // a[r][kk] (column kk is not rescaled except row kk)
              if (f != 0.0) a[r] -= f * a[kk];
            }
          }
          if (lane >= ICA_MAX_PARAMS && lane < 2 * ICA_MAX_PARAMS) {
            const int jc = lane - ICA_MAX_PARAMS;   // column of the inverse
#pragma unroll
            for (int r = 0; r < ICA_MAX_PARAMS; ++r) if (r < n && jc < n) st.hinv[r * n + jc] = singular ? 0.0 : a[r];
          }
          __syncwarp();
        }
        SOLVE_STAMP(4);
        if (lane < n) {                                    // io.parametric_solve (io.py:146-155)
          double a = 0.0;
          for (int j = 0; j < n; ++j) a += st.hinv[lane * n + j] * s_vec[j];
          s_vec[ICA_MAX_PARAMS + lane] = a;
        }
        __syncwarp();
        if (lane == 0) {
          double dp[ICA_MAX_PARAMS];
          double e2 = 0.0;
          for (int i = 0; i < n; ++i) { dp[i] = s_vec[ICA_MAX_PARAMS + i]; e2 += dp[i] * dp[i]; }
          const double err = sqrt(e2);
          // lambda decays after rho' was evaluated with the old value (ica.py:235-238)
          double lam = st.lambda_it;
          if (robust && P.lambda_cfg <= 0.0 && lam > kLambdaN) {
            lam *= kLambdaRatio;
            if (lam < kLambdaN) lam = kLambdaN;
          }
          for (int i = 0; i < n; ++i) st.p_prev[i] = st.p[i];
          update_transform(st.p, dp, ttype);
          const int itn = st.iter + 1;
          st.err = err;
          st.lambda_it = lam;
          st.total_iters += 1;
          if (P.traj && st.traj_count < P.traj_cap) {
            double* t = P.traj + ((long long)pair * P.traj_cap + st.traj_count) * ICA_TRAJ_STRIDE;
            t[0] = s; t[1] = itn - 1; t[2] = err; t[3] = lam;
            for (int i = 0; i < ICA_MAX_PARAMS; ++i) t[4 + i] = i < n ? st.p[i] : 0.0;
            st.traj_count += 1;
          }
          if (err > P.tol && itn < P.max_iter) {
            st.iter = itn;
          } else {  // this scale is done (ica.py:109, 225)
            *P.sched_dirty = 1;     // the pair changes level or leaves: the next work list differs
            st.iters_per_scale[s] = itn;
            if (s > 0) {
              double q[ICA_MAX_PARAMS];
              const LevelDesc Lf = P.lv[s - 1];
              zoom_in_parameters(st.p, ttype, (double)nx, (double)ny, (double)Lf.nx, (double)Lf.ny, q);
              for (int i = 0; i < n; ++i) st.p[i] = q[i];
              st.scale = s - 1;
              st.iter = 0;
              st.lambda_it = P.lambda_cfg > 0.0 ? P.lambda_cfg : kLambda0;  // new call per scale (ica.py:223)
            } else {
              st.scale = -1;
            }
          }
        }
      }
    }
    sync();
    SOLVE_STAMP(5);
    if (!P.dbg_Hb && tid < kStateWords)
      reinterpret_cast<unsigned long long*>(&P.state[pair])[tid] = reinterpret_cast<const unsigned long long*>(&sh.st)[tid];
  }
  else if (P.solve_mode == 1) {
    if (tid < kAccStride) P.ext_moments[(long long)pair * kAccStride + tid] = 0.0;
    return false;
  }
  return true;
#undef SOLVE_STAMP
}

// After a pair's solve: the caller whose pair completes the iteration (`expected` solves) builds the next work list.
template <typename Sync>
__device__ void finish_iteration(const IterParams& P, SolveShared& sh, unsigned expected, int tid, int nthr, Sync sync) {
  __threadfence();      // the pair's new state is visible before the ticket
  sync();
  if (tid == 0) sh.ticket = atomicAdd(P.solve_ticket, 1u);
  sync();
  if (sh.ticket != expected - 1) return;
  __threadfence();
  schedule_block(P, sh.warp_scan, sh.scal, false, tid, nthr, sync);
}

// The fused tail of the iterate kernel, out of line: its register needs (fp64 Gauss-Jordan, compositions) must not
// leak into the allocation of the per-pixel loop.
template <int DH>
__device__ __noinline__ void fused_tail(const IterParams& P, int pair, int tid, double* scratch, SolveShared* sh, unsigned expected) {
  solve_pair<DH, kConsumerWarps>(P, pair, tid, scratch, *sh, SyncConsumers(), false);
  finish_iteration(P, *sh, expected, tid, kConsumerThreads, SyncConsumers());
}

constexpr int kSolveThreads = 512;
constexpr int kSolveWarps = kSolveThreads / 32;

// Stand-alone K3: one block per image pair, launched after the iterate kernel (host-driven loop, row-sharded mode,
// parity hooks).  The last block to finish builds the work list of the next iteration.
template <int DH>
__global__ void __launch_bounds__(kSolveThreads) ica_solve_kernel(const __grid_constant__ IterParams P) {
  constexpr int NENT = RowVals<DH>::K * kYPow;
  __shared__ double s_part[kSolveWarps * NENT];
  __shared__ __align__(8) SolveShared sh;
  const int tid = threadIdx.x;
  if (!solve_pair<DH, kSolveWarps>(P, (int)blockIdx.x, tid, s_part, sh, SyncBlock(), blockIdx.x == 0)) return;
  finish_iteration(P, sh, gridDim.x, tid, kSolveThreads, SyncBlock());
}

// ============================================================ the kernel
// MODE 1 / 2 / 3 = the common configurations as compile-time constants (1: robust loop with the Lorentzian rho', 2:
// quadratic loop, 3: robust loop with Geman-McClure; all with the discarded frame and skimage's warp domain): their
// per-pixel code has no mode branches, which is worth 4-5 % of the kernel time.  MODE 0 reads the modes from the
// parameters.
template <int C, int DH, int MODE>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM) ica_iterate_kernel(const __grid_constant__ IterParams P) {
  constexpr int K = RowVals<DH>::K;
  constexpr int HW = DH + 1;        // x-powers kept for the Hessian moments
  constexpr int BWN = DH / 2 + 1;   // x-powers kept for the b moments
  constexpr int S1W = Stage<C>::S1W;
  constexpr int S2W = Stage<C>::S2W;
  constexpr int NENT = K * kYPow;
  constexpr int kStages = Stage<C>::kStages;
  static_assert(kChunkRing >= kStages + 2 && (kChunkRing & (kChunkRing - 1)) == 0, "chunk records outlive the stages");

  extern __shared__ __align__(128) float smem[];
  float* const stages = smem;                      // kStages x Stage<C>::kFloats
  __shared__ __align__(8) unsigned long long s_full[kStages], s_empty[kStages];
  __shared__ TileCtl tctl[kStages];
  __shared__ double s_pm64[9];
  __shared__ int4 s_cinfo[kChunkRing];
  __shared__ __align__(8) SolveShared s_solve;     // fused solve / scheduling (the CTA that finishes a pair's last chunk)
  __shared__ int s_par, s_last;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  // which of the two work lists this launch consumes: one thread decides for the CTA (with the fused solve the counter
  // can advance while a late CTA of the same launch starts; such a CTA simply works on the next iteration's list)
  if (tid == 0) {
    // safety net of the device-side loop: whatever happens to the scheduling, the graph's while node ends after
    // max_launches (+ slack) launches of this kernel
    if (blockIdx.x == 0 && P.cond_handle) {
      const int n = atomicAdd(P.loop_count + 1, 1);
      if (n > P.max_launches + 8) cudaGraphSetConditional(P.cond_handle, 0u);
    }
    s_par = __ldcg(P.loop_count) & 1;
    for (int i = 0; i < kStages; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], kConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int par = s_par;
  SchedHdr* const hdr = P.hdr + par;
  if ((int)blockIdx.x >= __ldcg(&hdr->total)) return;
  if (tid == 0) atomicMin(&hdr->t0, gtime());

  if (warp == kConsumerWarps) {
    producer_loop<C>(P, stages, Stage<C>::kFloats, s_full, s_empty, tctl, s_pm64, s_cinfo, par, lane);
    return;
  }

  // ------------------------------------------------------------------ consumers
  const int delta = P.delta;
  const bool frame = MODE != 0 ? true : P.frame != 0;
  const bool robust = (MODE == 1 || MODE == 3) ? true : (MODE == 2 ? false : P.robust_loop != 0);
  const float chm = C == 3 ? 1.0f : P.ch_mult;      // (only a gray image can stand for its RGB replication)
  const int rtype = MODE == 1 ? (int)LORENTZIAN : (MODE == 2 ? (int)QUADRATIC : (MODE == 3 ? (int)GERMAN_MCCLURE : P.robust_type));
  const bool ipol = MODE != 0 ? false : P.ipol_warp != 0, ipol_nan = MODE != 0 ? false : P.ipol_nan != 0;
  // fp64 accumulators of the chunk in progress, double-buffered over consecutive chunks: [2][kConsumerWarps][K][kYPow]
  double* const accs0 = reinterpret_cast<double*>(smem + kStages * Stage<C>::kFloats);
  constexpr int kAccSet = kConsumerWarps * K * kYPow;
  // the transposing reduction leaves moment k on the lanes k << kTrShift .. ; the first of them owns the fp64 accumulators
  constexpr int kTrN = K <= 8 ? 8 : (K <= 16 ? 16 : 32);
  constexpr int kTrShift = kTrN == 8 ? 2 : (kTrN == 16 ? 1 : 0);
  const int midx = lane >> kTrShift;
  const bool mown = (lane & ((1 << kTrShift) - 1)) == 0 && midx < K;
  double* myacc = accs0 + (warp * K + (mown ? midx : 0)) * kYPow;    // this lane's slot in the current set
  unsigned k = 0;
  int nitems = 0;
  const bool dbg = kTimeline && P.dbg_time != nullptr && tid == 0;       // profiling hook: cycles warp 0 spends waiting / in chunk epilogues
  long long dbg_wait = 0, dbg_epi = 0;
  const long long dbg_t0 = dbg ? clock64() : 0;

  if (kTimeline && P.dbg_time && tid == 0) P.dbg_time[blockIdx.x * 16 + 0] = gtime();
  // Per-lane x-moment accumulators of the image row `vrow` (-1: empty).  Consecutive tiles of a chunk usually lie on
  // the same tile row, so a warp keeps adding pixels of the same image row and pays the cross-lane reduction and
  // the fp64 fold once per row segment of the chunk instead of once per tile.
  float v[K];
#pragma unroll
  for (int i = 0; i < K; ++i) v[i] = 0.0f;
  int vrow = -1;
  // once per row segment: a transposing shuffle reduction leaves the row's moment k on the lane that owns it (fixed
  // summation order), which folds in y^b in fp64; the fp64 accumulators live in shared memory
  auto flush_row = [&]() {
    float t32[kTrN];
#pragma unroll
    for (int i = 0; i < kTrN; ++i) t32[i] = i < K ? v[i] : 0.0f;
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = 0.0f;
    const float tot = warp_transpose_reduce<kTrN>(t32, lane);
    if (mown) {
      const double yd = (double)vrow, t = (double)tot;
      double yp = 1.0;
#pragma unroll
      for (int b = 0; b < kYPow; ++b) { myacc[b] = fma(t, yp, myacc[b]); yp *= yd; }
    }
    vrow = -1;
  };
  for (int it = 0;; ++it) {
    if (mown) {
#pragma unroll
      for (int b = 0; b < kYPow; ++b) myacc[b] = 0.0;
    }
    bool last;
    bool stop = false;
    do {
      const int sidx = k % kStages;
      const long long w0 = dbg ? clock64() : 0;
      mbar_wait(&s_full[sidx], (k / kStages) & 1);
      if (dbg) dbg_wait += clock64() - w0;
      const TileCtl* tc = &tctl[sidx];
      {
        const int4 flg = lds_i4(&tc->flg);             // need_h, last, stop, pitch
        if (flg.z) { stop = true; break; }             // uniform: the producer ran out of work
        last = flg.y != 0;
      }
      if (k == 0) ICA_STAMP(1);
      const float* s2 = stages + sidx * Stage<C>::kFloats;
      const float* s1 = s2 + BH_MAX * S2W;
#pragma unroll 1
      for (int rr = 0; rr < TH / kConsumerWarps; ++rr) {
        const int ly = warp + rr * kConsumerWarps;
        const int4 geo = lds_i4(&tc->geo);             // x0, y0, nx, ny
        const int y = geo.y + ly;
        if (vrow >= 0 && vrow != y) flush_row();
        // rows of the discarded frame (ica.py:85-93) have no gradient: they add exact zeros to every moment -- skip them
        if (y < geo.w && (!frame || (y >= delta && y < geo.w - delta))) {
          vrow = y;
          const int nx = geo.z, ny = geo.w;
          const bool need_hr = (MODE == 1 || MODE == 3) ? true : lds_i4(&tc->flg).x != 0;
          // moments of one pixel: v[] += (rho' S, rho' v) * x^a.  Consecutive powers share a packed FFMA2 (the weight
          // is its broadcast operand): element by element the same fmaf as the scalar form.
          auto add_moments = [&](float scl, float sxx, float sxy, float syy, float vx, float vy, float xf) {
            const float x2 = xf * xf;
            const float2 xp01 = make_float2(1.0f, xf);
            const float2 xp23 = make_float2(x2, x2 * xf);
            const float x4 = x2 * x2;
            auto acc = [&](int base, int npow, float w) {
              const float2 w2 = make_float2(w, w);
              if (npow >= 2) {
                const float2 r = __ffma2_rn(w2, xp01, make_float2(v[base], v[base + 1]));
                v[base] = r.x; v[base + 1] = r.y;
              } else {
                v[base] = fmaf(w, 1.0f, v[base]);
              }
              if (npow == 3) v[base + 2] = fmaf(w, x2, v[base + 2]);
              if (npow >= 4) {
                const float2 r = __ffma2_rn(w2, xp23, make_float2(v[base + 2], v[base + 3]));
                v[base + 2] = r.x; v[base + 3] = r.y;
              }
              if (npow == 5) v[base + 4] = fmaf(w, x4, v[base + 4]);
            };
            if (need_hr) {
              acc(0 * HW, HW, scl * sxx); acc(1 * HW, HW, scl * sxy); acc(2 * HW, HW, scl * syy);
            }
            acc(3 * HW, BWN, scl * vx); acc(3 * HW + BWN, BWN, scl * vy);
          };
          // ---- the lane's two pixels A = (x0+lane, y), B = (x0+32+lane, y): projection in packed fp32 (same
          // arithmetic per pixel as project_px; no validity test: a staged window implies z > 0 and bounded
          // displacements over the whole tile, and lanes without a window take the generic path below)
          const int xA = geo.x + lane, xB = xA + 32;
          int cxA, cyA, cxB, cyB;
          float2 tx2, ty2;
          {
            const float4 qa = lds_f4(&tc->c8[0]), qb = lds_f4(&tc->c8[4]);   // d00, m01, m02, m10 / d11, m12, m20, m21
            const float2 c_d00 = make_float2(qa.x, qa.x), c_m01 = make_float2(qa.y, qa.y), c_m02 = make_float2(qa.z, qa.z);
            const float2 c_m10 = make_float2(qa.w, qa.w), c_d11 = make_float2(qb.x, qb.x), c_m12 = make_float2(qb.y, qb.y);
            const float fy = (float)y, fxa = (float)xA;
            const float2 fy2 = make_float2(fy, fy), nfy2 = make_float2(-fy, -fy);
            const float2 fx2 = make_float2(fxa, fxa + 32.0f), nfx2 = make_float2(-fxa, -(fxa + 32.0f));
            float2 nx_ = __ffma2_rn(c_d00, fx2, __ffma2_rn(c_m01, fy2, c_m02));
            float2 ny_ = __ffma2_rn(c_m10, fx2, __ffma2_rn(c_d11, fy2, c_m12));
            float2 dx2 = nx_, dy2 = ny_;
            if (DH == 4) {
              // (only a batch with a homography has moment degree 4: for the affine family the bottom row of the matrix is
              // (0, 0, 1), z is exactly 1 and the perspective terms below are exact no-ops, so they are not issued)
              const float2 zm1 = __ffma2_rn(make_float2(qb.z, qb.z), fx2, __fmul2_rn(make_float2(qb.w, qb.w), fy2));
              nx_ = __ffma2_rn(nfx2, zm1, nx_);
              ny_ = __ffma2_rn(nfy2, zm1, ny_);
              const float zA = 1.0f + zm1.x, zB = 1.0f + zm1.y;
              const float2 rz = make_float2(fast_rcp(zA), fast_rcp(zB));
              dx2 = __fmul2_rn(nx_, rz); dy2 = __fmul2_rn(ny_, rz);
            }
            const float flxA = floorf(dx2.x), flxB = floorf(dx2.y), flyA = floorf(dy2.x), flyB = floorf(dy2.y);
            tx2 = make_float2(dx2.x - flxA, dx2.y - flxB);
            ty2 = make_float2(dy2.x - flyA, dy2.y - flyB);
            cxA = xA + (int)flxA; cxB = xB + (int)flxB; cyA = y + (int)flyA; cyB = y + (int)flyB;
            // tie band of project_px, one test for both pixels (the wider of the two bands)
            const float amax = fmaxf(fabsf(dx2.x) + fabsf(dy2.x), fabsf(dx2.y) + fabsf(dy2.y));
            const float kTie = fmaf(1.0e-6f, amax, 2.5e-4f);
            const float tmin = fminf(fminf(tx2.x, tx2.y), fminf(ty2.x, ty2.y));
            const float tmax = fmaxf(fmaxf(tx2.x, tx2.y), fmaxf(ty2.x, ty2.y));
            if (tmin < kTie || tmax > 1.0f - kTie) {
              // (rare) a coordinate next to an integer: the exact fp64 evaluation decides the tap set
              // (kept inline: as an out-of-line call its by-reference results live in local memory, measured 3 % slower)
              WarpCoef coef;
              coef.d00 = qa.x; coef.m01 = qa.y; coef.m02 = qa.z; coef.m10 = qa.w; coef.d11 = qb.x; coef.m12 = qb.y; coef.m20 = qb.z; coef.m21 = qb.w;
              float a, b;
              project_px(coef, tc->m64, xA, y, cxA, cyA, a, b); tx2.x = a; ty2.x = b;
              project_px(coef, tc->m64, xB, y, cxB, cyB, a, b); tx2.y = a; ty2.y = b;
            }
          }
          bool insmA, insmB;
          int offA, offB;    // float offsets of the first tap inside the staged I2 window
          {
            const int4 win = lds_i4(&tc->win);   // bx0, by0, bw - 3, bh - 3 (0, 0 without a window)
            // 0 <= c - 1 - b0 <= extent - 4, as one unsigned comparison per axis
            const int oxA = cxA - 1 - win.x, oyA = cyA - 1 - win.y, oxB = cxB - 1 - win.x, oyB = cyB - 1 - win.y;
            insmA = (unsigned)oxA < (unsigned)win.z && (unsigned)oyA < (unsigned)win.w;
            insmB = (unsigned)oxB < (unsigned)win.z && (unsigned)oyB < (unsigned)win.w;
            offA = oyA * S2W + oxA * C;
            offB = oyB * S2W + oxB * C;
          }
          const bool fastlane = insmA && insmB && xB < nx;
          if (__all_sync(0xffffffffu, fastlane)) {
            // ===== straight-line path: both pixels in one instruction stream, packed fp32 (FFMA2)
            // phase 1: the 2 x 16 taps of every channel -> warped values (few live registers besides the loads)
            float2 iw[C];
            {
              float2 wx[4], wy[4];
              keys_weights2(tx2, wx[0], wx[1], wx[2], wx[3]);
              keys_weights2(ty2, wy[0], wy[1], wy[2], wy[3]);
              const float* tA = s2 + offA;
              const float* tB = s2 + offB;
#pragma unroll
              for (int ch = 0; ch < C; ++ch) {
                float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float* rA = tA + q * S2W + ch;
                  const float* rB = tB + q * S2W + ch;
                  float2 h = __fmul2_rn(wx[0], make_float2(rA[0], rB[0]));
                  h = __ffma2_rn(wx[1], make_float2(rA[C], rB[C]), h);
                  h = __ffma2_rn(wx[2], make_float2(rA[2 * C], rB[2 * C]), h);
                  h = __ffma2_rn(wx[3], make_float2(rA[3 * C], rB[3 * C]), h);
                  acc2 = __ffma2_rn(wy[q], h, acc2);
                }
                iw[ch] = acc2;
              }
            }
            // IPOL-style warp domain (bi.py:144): the projected point must lie in [delta, n - 1 - delta]; with the exact
            // integer parts this is a test on (c, t).  Valid pixels have their 4 x 4 taps inside the image (delta >= 2).
            bool okA = true, okB = true, inA = true, inB = true;
            if (ipol) {
              const int hx = nx - 1 - delta, hy = ny - 1 - delta;
              inA = cxA >= delta && cyA >= delta && (cxA < hx || (cxA == hx && tx2.x == 0.0f)) && (cyA < hy || (cyA == hy && ty2.x == 0.0f));
              inB = cxB >= delta && cyB >= delta && (cxB < hx || (cxB == hx && tx2.y == 0.0f)) && (cyB < hy || (cyB == hy && ty2.y == 0.0f));
              okA = inA || !ipol_nan; okB = inB || !ipol_nan;
            }
            // phase 2: residual, gradient of I1 (masks fold the 1/2, the frame and the image border), S, v
            float2 mgx, mgy;
            {
              const int4 msk = lds_i4(&tc->msk);   // gxlo, gxspan, fxlo, fxspan
              // (rows of the frame never get here: see the row test above)
              mgx = make_float2((unsigned)(xA - msk.x) < (unsigned)msk.y ? 0.5f : 0.0f,
                                (unsigned)(xB - msk.x) < (unsigned)msk.y ? 0.5f : 0.0f);
              if (MODE != 0) {
                // with a frame (delta >= 1) every remaining row has a central y-difference, and the columns with an
                // x-gradient are exactly the columns inside the frame: one mask serves both gradients
                mgy = mgx;
              } else {
                const bool gyrow = y >= 1 && y <= ny - 2;
                mgy = make_float2((gyrow && (unsigned)(xA - msk.z) < (unsigned)msk.w) ? 0.5f : 0.0f,
                                  (gyrow && (unsigned)(xB - msk.z) < (unsigned)msk.w) ? 0.5f : 0.0f);
              }
            }
            const float2 nmgx = make_float2(-mgx.x, -mgx.y), nmgy = make_float2(-mgy.x, -mgy.y);
            const float* cA = s1 + (ly + 1) * S1W + (lane + HALO) * C;
            const float* cB = cA + 32 * C;
            const float4 fl = lds_f4(&tc->fl);     // lo, hi, lambda^2
            float2 sxx = make_float2(0.f, 0.f), sxy = sxx, syy = sxx, vx = sxx, vy = sxx, t2 = sxx;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
              bool vA, vB;
              float iwA, iwB;
              if (!ipol) {
                vA = iw[ch].x == iw[ch].x; vB = iw[ch].y == iw[ch].y;            // NaN footprint
                iwA = fminf(fmaxf(iw[ch].x, fl.x), fl.y); iwB = fminf(fmaxf(iw[ch].y, fl.x), fl.y);   // clip (Q1)
              } else {
                vA = okA; vB = okB;
                iwA = inA ? iw[ch].x : 0.0f; iwB = inB ? iw[ch].y : 0.0f;
              }
              const float2 gx = __ffma2_rn(make_float2(cA[ch + C], cB[ch + C]), mgx, __fmul2_rn(make_float2(cA[ch - C], cB[ch - C]), nmgx));
              const float2 gy = __ffma2_rn(make_float2(cA[ch + S1W], cB[ch + S1W]), mgy, __fmul2_rn(make_float2(cA[ch - S1W], cB[ch - S1W]), nmgy));
              const float2 di = make_float2(vA ? iwA - cA[ch] : 0.0f, vB ? iwB - cB[ch] : 0.0f);   // non-finite -> 0 (io.py:72, 134)
              if (need_hr) { sxx = __ffma2_rn(gx, gx, sxx); sxy = __ffma2_rn(gx, gy, sxy); syy = __ffma2_rn(gy, gy, syy); }
              vx = __ffma2_rn(gx, di, vx); vy = __ffma2_rn(gy, di, vy);
              t2 = __ffma2_rn(di, di, t2);
            }
            // phase 3: robust weight and moments.  A gray image stands for its x3 replication
            // (SURVEY Q12): every channel sum triples
            const float rhoA = robust ? rho_prime(t2.x * chm, fl.z, rtype) : 1.0f;
            const float rhoB = robust ? rho_prime(t2.y * chm, fl.z, rtype) : 1.0f;
            add_moments(rhoA * chm, sxx.x, sxy.x, syy.x, vx.x, vy.x, (float)xA);
            add_moments(rhoB * chm, sxx.y, sxy.y, syy.y, vx.y, vy.y, (float)xB);
          } else {
            // ===== generic path: image edges, windows that do not fit, degenerate projections
            const bool yin = !frame || (y >= delta && y < ny - delta);
            const int4 win = lds_i4(&tc->win);
            const float4 fl = lds_f4(&tc->fl);
            WarpCoef coef;
            coef.d00 = tc->c8[0]; coef.m01 = tc->c8[1]; coef.m02 = tc->c8[2]; coef.m10 = tc->c8[3];
            coef.d11 = tc->c8[4]; coef.m12 = tc->c8[5]; coef.m20 = tc->c8[6]; coef.m21 = tc->c8[7];
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
              const int x = half ? xB : xA;
              if (x >= nx) continue;
              const int lx = lane + half * 32;
              int cx, cy; float tx, ty;
              const bool pok = project_px(coef, tc->m64, x, y, cx, cy, tx, ty);
              const int ox = cx - 1 - win.x, oy = cy - 1 - win.y;
              const bool insm = (unsigned)ox < (unsigned)win.z && (unsigned)oy < (unsigned)win.w;
              float wxs[4], wys[4];
              keys_weights(tx, wxs[0], wxs[1], wxs[2], wxs[3]);
              keys_weights(ty, wys[0], wys[1], wys[2], wys[3]);
              const bool inframe = yin && (!frame || (x >= delta && x < nx - delta));
              const float* t2base = s2 + oy * S2W + ox * C;
              const float* c1 = s1 + (ly + 1) * S1W + (lx + HALO) * C;
              const bool gxok = inframe && x >= 1 && x <= nx - 2;
              const bool gyok = inframe && y >= 1 && y <= ny - 2;
              const float lo = fl.x, hi = fl.y;
              float sxx = 0.f, sxy = 0.f, syy = 0.f, vx = 0.f, vy = 0.f, t2 = 0.f;
#pragma unroll
              for (int ch = 0; ch < C; ++ch) {
                float iwv;
                if (!pok) {
                  iwv = __int_as_float(0x7fc00000);
                } else if (insm) {
                  float a = 0.0f;
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const float* r = t2base + q * S2W + ch;
                    float hsum = wxs[0] * r[0] + wxs[1] * r[C] + wxs[2] * r[2 * C] + wxs[3] * r[3 * C];
                    a = fmaf(wys[q], hsum, a);
                  }
                  iwv = a;
                } else {
                  iwv = sample_global_slow<C>(tc->I2, lds_i4(&tc->flg).w, nx, ny, cx, cy, ch, wxs[0], wxs[1], wxs[2], wxs[3], wys[0], wys[1], wys[2], wys[3]);
                }
                bool valid;
                if (!ipol) {
                  valid = iwv == iwv;                      // NaN footprint
                  iwv = fminf(fmaxf(iwv, lo), hi);         // clip (only used when valid)
                } else {
                  const int hx = nx - 1 - delta, hy = ny - 1 - delta;
                  const bool in = pok && cx >= delta && cy >= delta && (cx < hx || (cx == hx && tx == 0.0f)) && (cy < hy || (cy == hy && ty == 0.0f));
                  valid = in || !ipol_nan;
                  iwv = in ? iwv : 0.0f;
                }
                const float i1c = c1[ch];
                const float gx = gxok ? 0.5f * (c1[ch + C] - c1[ch - C]) : 0.0f;
                const float gy = gyok ? 0.5f * (c1[ch + S1W] - c1[ch - S1W]) : 0.0f;
                const float di = valid ? iwv - i1c : 0.0f; // non-finite -> 0 (io.py:72, 134)
                if (need_hr) { sxx = fmaf(gx, gx, sxx); sxy = fmaf(gx, gy, sxy); syy = fmaf(gy, gy, syy); }
                vx = fmaf(gx, di, vx); vy = fmaf(gy, di, vy);
                t2 = fmaf(di, di, t2);
              }
              const float rho = robust ? rho_prime(t2 * chm, fl.z, rtype) : 1.0f;
              add_moments(rho * chm, sxx, sxy, syy, vx, vy, (float)x);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sidx]);   // this warp is done with the stage (and its TileCtl)
      ++k;
    } while (!last);
    if (stop) break;
    if (vrow >= 0) flush_row();   // the chunk's last row segment
    const int4 ci = s_cinfo[nitems & (kChunkRing - 1)];     // pair, chunk, chunks of the pair (written by the producer)
    const int pair = ci.x, chunk = ci.y, nch = ci.z;
    ++nitems;
    ICA_STAMP(2);
    const long long e0 = dbg ? clock64() : 0;

    // ---------------- chunk partial: [K][kYPow] doubles, warps summed in fixed order.  One barrier per chunk: the
    // next chunk accumulates into the other set, and this set is only zeroed again after the next chunk's barrier,
    // which the summing threads reach after they are done with it.
    consumer_sync();
    {
      const double* accs = accs0 + (nitems & 1 ? 0 : kAccSet);   // nitems was just incremented: the set of this chunk
      double* out = P.partials + ((long long)pair * P.max_chunks + chunk) * kAccStride;
      for (int i = tid; i < NENT; i += kConsumerThreads) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < kConsumerWarps; ++w) sum += accs[w * NENT + i];
        out[i] = sum;
      }
    }
    myacc += (nitems & 1) ? kAccSet : -kAccSet;     // the other set for the next chunk
    if (tid == 0) atomicMax(&hdr->t1, gtime());     // end of the streaming phase so far (bench: kernel time)
    if (P.fused) {
      // ---------------- fused K3: the CTA that completes the pair's last chunk sums the pair's partials, solves,
      // composes and -- if it was the iteration's last pair -- builds the next work list; everybody else moves on.
      __threadfence();                              // this CTA's partial is visible before its ticket
      consumer_sync();
      if (tid == 0) {
        const unsigned t = atomicAdd(&P.pair_ticket[pair], 1u);
        s_last = (t == (unsigned)nch - 1u) ? 1 : 0;
        if (s_last) P.pair_ticket[pair] = 0;        // for the next iteration
      }
      consumer_sync();
      if (s_last) {
        __threadfence();
        // scratch: the accumulator set the NEXT chunk will use is idle now (it is zeroed at the top of the loop)
        double* scratch = accs0 + ((nitems & 1) ? kAccSet : 0);
        const unsigned expected = (unsigned)__ldcg(&hdr->npairs);
        fused_tail<DH>(P, pair, tid, scratch, &s_solve, expected);
        consumer_sync();
      }
    }
    if (dbg) dbg_epi += clock64() - e0;
  }
  if (dbg) {
    long long* d = P.dbg_time + blockIdx.x * 16;
    d[3] = dbg_wait; d[4] = dbg_epi; d[5] = 0; d[6] = clock64() - dbg_t0; d[15] = k;
  }
  if (kTimeline && P.dbg_time && tid == 32) { P.dbg_time[blockIdx.x * 16 + 13] = gtime(); P.dbg_time[blockIdx.x * 16 + 14] = nitems; }
}

// Resets the per-pair state at the start of a run (ica.py:319-337: ps[0] = p, ps[s>0] = 0;
// the coarse-to-fine loop starts at the coarsest scale).
__global__ void ica_init_state_kernel(PairState* state, const double* p_in, const int* ttypes,
                                      int B, int nscales, double lambda_cfg, int* n_active, unsigned int* pair_ticket) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) *n_active = B;
  if (b >= B) return;
  if (pair_ticket) pair_ticket[b] = 0;
  PairState& st = state[b];
  const int tt = ttypes[b];
  const int n = nparams_of(tt);
  for (int i = 0; i < ICA_MAX_PARAMS; ++i) {
    st.p[i] = (nscales == 1 && i < n) ? p_in[b * ICA_MAX_PARAMS + i] : 0.0;
    st.p_prev[i] = st.p[i];
  }
  for (int i = 0; i < ICA_MAX_PARAMS * ICA_MAX_PARAMS; ++i) st.hinv[i] = 0.0;
  st.lambda_it = lambda_cfg > 0.0 ? lambda_cfg : kLambda0;
  st.err = 1e10;
  st.scale = nscales - 1;
  st.iter = 0;
  st.ttype = tt;
  st.nparams = n;
  st.total_iters = 0;
  st.traj_count = 0;
  for (int i = 0; i < ICA_MAX_SCALES; ++i) st.iters_per_scale[i] = 0;
  st.ticket = 0;
}

__global__ void ica_export_results_kernel(const PairState* state, int B, double* p_out, double* err_out,
                                          int* iters_out, int nscales) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const PairState& st = state[b];
  for (int i = 0; i < ICA_MAX_PARAMS; ++i) p_out[b * ICA_MAX_PARAMS + i] = i < st.nparams ? st.p[i] : 0.0;
  if (err_out) err_out[b] = st.err;
  if (iters_out) for (int s = 0; s < nscales; ++s) iters_out[b * nscales + s] = st.iters_per_scale[s];
}

// K5: Iw and DI of the reference's return value: the warp with the parameters of the LAST
// iteration (p_prev, i.e. before the final update; ica.py:227-251, 261) at the finest scale.
template <int C>
__global__ void ica_warp_out_kernel(const float* __restrict__ I1_0, const float* __restrict__ I2_0,
                                    long long in_stride, int nx, int ny, int pitch,
                                    const PairState* state, const MinMaxKeys* mm, int nscales,
                                    float* __restrict__ Iw, float* __restrict__ DI, int ipol, int ipol_nan, int delta) {
  const int pair = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  __shared__ WarpCoef sc;
  __shared__ double m[9];
  __shared__ float slo, shi;
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    warp_matrix(state[pair].p_prev, state[pair].ttype, m);
    sc = make_warp_coef(m);
    const MinMaxKeys k = mm[(pair * nscales + 0) * 2 + 1];
    slo = key_float(k.lo); shi = key_float(k.hi);
  }
  __syncthreads();
  if (x >= nx || y >= ny) return;
  const float* I1 = I1_0 + (long long)pair * in_stride;
  const float* I2 = I2_0 + (long long)pair * in_stride;
  int cx, cy; float tx, ty;
  const bool pok = project_px(sc, m, x, y, cx, cy, tx, ty);
  float wx[4], wy[4];
  keys_weights(tx, wx[0], wx[1], wx[2], wx[3]);
  keys_weights(ty, wy[0], wy[1], wy[2], wy[3]);
  const long long o = (long long)pair * in_stride + (long long)y * pitch + x * C;
#pragma unroll
  for (int ch = 0; ch < C; ++ch) {
    float iw = pok ? sample_global<C>(I2, pitch, nx, ny, cx, cy, ch, wx, wy) : __int_as_float(0x7fc00000);
    if (!ipol) {
      if (iw == iw) iw = fminf(fmaxf(iw, slo), shi);
    } else {      // IPOL-style domain on the projected point, no clip (bi.py:144-150)
      const int hx = nx - 1 - delta, hy = ny - 1 - delta;
      const bool in = pok && cx >= delta && cy >= delta && (cx < hx || (cx == hx && tx == 0.0f)) && (cy < hy || (cy == hy && ty == 0.0f));
      if (!in) iw = ipol_nan ? __int_as_float(0x7fc00000) : 0.0f;
    }
    Iw[o + ch] = iw;
    DI[o + ch] = iw - I1[(long long)y * pitch + x * C + ch];
  }
}

// Stand-alone warp by an explicit matrix (ica_warp_host; bi.bicubic_interpolation_skimage).
struct Mat9 { double m[9]; };

template <int C>
__global__ void ica_warp_matrix_kernel(const float* __restrict__ img, int nx, int ny, WarpCoef coef, Mat9 mat,
                                       const MinMaxKeys* mm, float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= nx || y >= ny) return;
  const float lo = key_float(mm->lo), hi = key_float(mm->hi);
  int cx, cy; float tx, ty;
  const bool pok = project_px(coef, mat.m, x, y, cx, cy, tx, ty);
  float wx[4], wy[4];
  keys_weights(tx, wx[0], wx[1], wx[2], wx[3]);
  keys_weights(ty, wy[0], wy[1], wy[2], wy[3]);
#pragma unroll
  for (int ch = 0; ch < C; ++ch) {
    float iw = pok ? sample_global<C>(img, nx * C, nx, ny, cx, cy, ch, wx, wy) : __int_as_float(0x7fc00000);
    if (iw == iw) iw = fminf(fmaxf(iw, lo), hi);
    out[((long long)y * nx + x) * C + ch] = iw;
  }
}

// Gradient + frame as separate images (helper API / parity hook; the loop never stores them).
template <int C>
__global__ void ica_gradient_kernel(const float* __restrict__ img, int nx, int ny, int delta, int frame,
                                    float* __restrict__ Ix, float* __restrict__ Iy) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= nx || y >= ny) return;
  const bool infr = frame && (x < delta || x >= nx - delta || y < delta || y >= ny - delta);
#pragma unroll
  for (int ch = 0; ch < C; ++ch) {
    const long long o = ((long long)y * nx + x) * C + ch;
    float gx = 0.f, gy = 0.f;
    if (x >= 1 && x <= nx - 2) gx = 0.5f * (img[o + C] - img[o - C]);
    if (y >= 1 && y <= ny - 2) gy = 0.5f * (img[o + (long long)nx * C] - img[o - (long long)nx * C]);
    if (infr) { gx = __int_as_float(0x7fc00000); gy = gx; }
    Ix[o] = gx; Iy[o] = gy;
  }
}

template <int C, int DH, int MODE>
cudaError_t launch_iterate_m(const IterParams& P, int grid, cudaStream_t stream) {
  constexpr size_t smem = Stage<C>::kStages * (size_t)Stage<C>::kFloats * sizeof(float) +
                          2 * (size_t)kConsumerWarps * RowVals<DH>::K * kYPow * sizeof(double);
  // the attribute belongs to the (device, function) pair: configure once per device
  static std::atomic<unsigned long long> configured{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    cudaError_t e = cudaFuncSetAttribute(ica_iterate_kernel<C, DH, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.fetch_or(bit, std::memory_order_release);
  }
  ica_iterate_kernel<C, DH, MODE><<<grid, kThreads, smem, stream>>>(P);
  return cudaGetLastError();
}

template <int C, int DH>
cudaError_t launch_iterate_t(const IterParams& P, int grid, cudaStream_t stream) {
  if (P.frame && !P.ipol_warp) {
    if (P.robust_loop && P.robust_type == LORENTZIAN) return launch_iterate_m<C, DH, 1>(P, grid, stream);
    if (P.robust_loop && P.robust_type == GERMAN_MCCLURE) return launch_iterate_m<C, DH, 3>(P, grid, stream);
    if (!P.robust_loop) return launch_iterate_m<C, DH, 2>(P, grid, stream);
  }
  return launch_iterate_m<C, DH, 0>(P, grid, stream);
}

}  // namespace

void build_assembly_table(int dh, AsmEntry* tab) {
  const int hw = dh + 1, bwn = dh / 2 + 1, boff = 3 * hw;
  for (int i = 0; i < 6 * 72; ++i) for (int q = 0; q < 4; ++q) { tab[i].idx[q] = 0; tab[i].coef[q] = 0.f; }
  for (int ttype = TRANSLATION; ttype <= HOMOGRAPHY; ++ttype) {
    if (moment_degree_of(ttype) > dh) continue;
    Mono jx[ICA_MAX_PARAMS], jy[ICA_MAX_PARAMS];
    jacobian_monomials(ttype, jx, jy);
    const int n = nparams_of(ttype);
    for (int k = 0; k < n; ++k) {
      for (int l = 0; l < n; ++l) {
        AsmEntry& e = tab[ttype * 72 + k * 8 + l];
        const Mono* a[4] = {&jx[k], &jx[k], &jy[k], &jy[k]};
        const Mono* b[4] = {&jx[l], &jy[l], &jx[l], &jy[l]};
        const int ij[4] = {0, 1, 1, 2};
        for (int q = 0; q < 4; ++q) {
          if (a[q]->coef && b[q]->coef) {
            e.coef[q] = (float)(a[q]->coef * b[q]->coef);
            e.idx[q] = (ij[q] * hw + a[q]->a + b[q]->a) * kYPow + a[q]->b + b[q]->b;
          }
        }
      }
      AsmEntry& e = tab[ttype * 72 + 64 + k];
      if (jx[k].coef) { e.coef[0] = (float)jx[k].coef; e.idx[0] = (boff + 0 * bwn + jx[k].a) * kYPow + jx[k].b; }
      if (jy[k].coef) { e.coef[1] = (float)jy[k].coef; e.idx[1] = (boff + 1 * bwn + jy[k].a) * kYPow + jy[k].b; }
    }
  }
}

int iterate_tile_w() { return TW; }
int iterate_tile_h() { return TH; }
void iterate_stage_boxes(int channels, int* w1, int* h1, int* w2, int* h2) {
  *w1 = channels == 3 ? Stage<3>::S1W : Stage<1>::S1W; *h1 = S1ROWS;
  *w2 = channels == 3 ? Stage<3>::S2W : Stage<1>::S2W; *h2 = BH_MAX;
}
int iterate_blocks_per_sm() { return kBlocksPerSM; }

cudaError_t launch_schedule(const IterParams& P, cudaStream_t stream) {
  ica_schedule_kernel<<<1, 1024, 0, stream>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_solve(const IterParams& P, int dh, cudaStream_t stream) {
  if (dh == 4) ica_solve_kernel<4><<<P.B, kSolveThreads, 0, stream>>>(P);
  else if (dh == 2) ica_solve_kernel<2><<<P.B, kSolveThreads, 0, stream>>>(P);
  else ica_solve_kernel<0><<<P.B, kSolveThreads, 0, stream>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_iterate(const IterParams& P, int channels, int dh, int grid, cudaStream_t stream) {
  if (channels == 3) {
    if (dh == 4) return launch_iterate_t<3, 4>(P, grid, stream);
    if (dh == 2) return launch_iterate_t<3, 2>(P, grid, stream);
    return launch_iterate_t<3, 0>(P, grid, stream);
  }
  if (dh == 4) return launch_iterate_t<1, 4>(P, grid, stream);
  if (dh == 2) return launch_iterate_t<1, 2>(P, grid, stream);
  return launch_iterate_t<1, 0>(P, grid, stream);
}

cudaError_t launch_init_state(PairState* state, const double* p_in, const int* ttypes, int B, int nscales,
                              double lambda_cfg, int* n_active, unsigned int* pair_ticket, cudaStream_t stream) {
  ica_init_state_kernel<<<(B + 127) / 128, 128, 0, stream>>>(state, p_in, ttypes, B, nscales, lambda_cfg, n_active, pair_ticket);
  return cudaGetLastError();
}

cudaError_t launch_export_results(const PairState* state, int B, double* p_out, double* err_out, int* iters_out,
                                  int nscales, cudaStream_t stream) {
  ica_export_results_kernel<<<(B + 127) / 128, 128, 0, stream>>>(state, B, p_out, err_out, iters_out, nscales);
  return cudaGetLastError();
}

cudaError_t launch_warp_out(const float* I1_0, const float* I2_0, long long in_stride, int nx, int ny, int channels,
                            const PairState* state, const MinMaxKeys* mm, int nscales, int B, float* Iw, float* DI,
                            int ipol_warp, int ipol_nan, int delta, cudaStream_t stream) {
  dim3 block(32, 8), grid((nx + 31) / 32, (ny + 7) / 8, B);
  if (channels == 3)
    ica_warp_out_kernel<3><<<grid, block, 0, stream>>>(I1_0, I2_0, in_stride, nx, ny, nx * 3, state, mm, nscales, Iw, DI, ipol_warp, ipol_nan, delta);
  else
    ica_warp_out_kernel<1><<<grid, block, 0, stream>>>(I1_0, I2_0, in_stride, nx, ny, nx, state, mm, nscales, Iw, DI, ipol_warp, ipol_nan, delta);
  return cudaGetLastError();
}

cudaError_t launch_warp_matrix(const float* img, int nx, int ny, int channels, const double* m9, const MinMaxKeys* mm,
                               float* out, cudaStream_t stream) {
  dim3 block(32, 8), grid((nx + 31) / 32, (ny + 7) / 8);
  const WarpCoef coef = make_warp_coef(m9);
  Mat9 mat;
  for (int i = 0; i < 9; ++i) mat.m[i] = m9[i];
  if (channels == 3) ica_warp_matrix_kernel<3><<<grid, block, 0, stream>>>(img, nx, ny, coef, mat, mm, out);
  else ica_warp_matrix_kernel<1><<<grid, block, 0, stream>>>(img, nx, ny, coef, mat, mm, out);
  return cudaGetLastError();
}

cudaError_t launch_gradient(const float* img, int nx, int ny, int channels, int delta, int frame, float* Ix,
                            float* Iy, cudaStream_t stream) {
  dim3 block(32, 8), grid((nx + 31) / 32, (ny + 7) / 8);
  if (channels == 3) ica_gradient_kernel<3><<<grid, block, 0, stream>>>(img, nx, ny, delta, frame, Ix, Iy);
  else ica_gradient_kernel<1><<<grid, block, 0, stream>>>(img, nx, ny, delta, frame, Ix, Iy);
  return cudaGetLastError();
}

}  // namespace ica
