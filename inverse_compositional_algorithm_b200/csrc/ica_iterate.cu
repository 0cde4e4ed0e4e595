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
// and, in the block that delivers a pair's last chunk (ticket pattern, nobody spins):
//   de.inverse_hessian, io.parametric_solve, tr.update_transform, the lambda schedule, the
//   stopping rule and zm.zoom_in_parameters at a scale change.
//
// Execution model
//   * ica_schedule_kernel (1 block) turns the per-pair scales into a work list: pair b at a
//     level with T tiles contributes min(T, max_chunks) chunks of consecutive 64x16 tiles.
//   * ica_iterate_kernel is PERSISTENT: 2 CTAs per SM, each CTA walks the chunk list with a
//     fixed stride, so a ragged batch (pairs at different scales, different iteration counts)
//     still fills all 148 SMs and a single straggler at the finest scale is spread over the
//     whole chip.
//   * per tile the I1 patch (+halo) and the window of I2 the warp can touch are staged into
//     shared memory with 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) completing on an
//     mbarrier; two stages, so the copy of tile t+1 overlaps the arithmetic of tile t.  Rows or
//     columns of the window that fall outside the image are filled by ordinary stores (NaN for
//     I2 = skimage's cval, so the NaN footprint falls out of the arithmetic).
//   * a warp owns one image row at a time; x-moments are accumulated per lane in fp32, a
//     transposing shuffle reduction leaves moment k on lane k, which folds in y^b in fp64.
//     Chunk partials go to fixed slots, the last chunk of a pair sums them in a fixed order
//     (deterministic) and runs the n x n solve / compose epilogue.
// Bound: HBM (read I1 once + I2 once per pixel-iteration = 2*C*4 bytes); no tensor cores.
#include "ica_device.cuh"
#include "ica_transform.cuh"
#include "ica_iterate.cuh"

namespace ica {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int TW = 64;            // tile width  (2 pixels per lane: x0+lane, x0+32+lane)
constexpr int TH = 16;            // tile height (2 rows per warp: y0+warp, y0+8+warp)
constexpr int HALO = 4;           // I1 patch starts at x0-4 so that row starts are 16-byte aligned
constexpr int S1PX = TW + 2 * HALO;
constexpr int S1ROWS = TH + 2;
constexpr int BW_MAX = 88;        // staged I2 window (pixels, multiple of 4); larger -> global path
constexpr int BH_MAX = 28;
constexpr int kSchedCap = 4096;   // pairs whose chunk table is cached in shared memory

template <int DH> struct RowVals { static constexpr int K = 3 * (DH + 1) + 2 * (DH / 2 + 1);
                                   static constexpr int NP = K > 16 ? 32 : (K > 8 ? 16 : 8); };

template <int C> struct Stage {
  static constexpr int S1W = S1PX * C;
  static constexpr int S2W = BW_MAX * C;
  static constexpr int kFloats = BH_MAX * S2W + S1ROWS * S1W;   // I2 window, then I1 patch
};

struct PairCtx {          // per-pair constants of the chunk being processed (shared memory)
  double m64[9];          // warp matrix in fp64 (tie-break path of project_px)
  WarpCoef coef;
  float lo, hi;           // clip range of I2 at this level
  float lambda2;
  int scale, need_h, pair;
};

struct StageCtl { int bx0, by0, bw, bh, fits; };

// ---------------------------------------------------------------- mbarrier / bulk-copy PTX
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_arrive(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One image-row segment [xa, xa+w) of row yy: the part that lies inside the image and whose
// byte range is a multiple of 16 goes through a bulk copy ([xs, xe4), `bytes`); the rest
// (out-of-image pixels, a ragged right end, or everything when bulk copies are not possible)
// is filled by fill_row_rest with ordinary loads/stores.
struct RowPlan { int xs, xe4; unsigned bytes; };

template <int C>
__device__ __forceinline__ RowPlan plan_row(int yy, int xa, int w, int nx, int ny, bool bulk_ok) {
  RowPlan r; r.xs = xa; r.xe4 = xa; r.bytes = 0;
  if (yy < 0 || yy >= ny || !bulk_ok) return r;
  int xs = max(xa, 0), xe = min(xa + w, nx);
  if (xe <= xs) return r;
  int xe4 = xs + ((xe - xs) & ~3);
  r.xs = xs; r.xe4 = xe4; r.bytes = (unsigned)(xe4 - xs) * C * 4u;
  return r;
}

template <int C>
__device__ __forceinline__ void fill_row_rest(float* dst, const float* __restrict__ img, int pitch, int yy, int xa,
                                              int w, int nx, int ny, int xs, int xe4, float fill, int lane) {
  const bool rowin = yy >= 0 && yy < ny;
  const float* src = img + (long long)yy * pitch + (long long)xa * C;
  const int left = (xs - xa) * C;            // floats before the bulk part
  const int right0 = (xe4 - xa) * C;         // first float after it
  const int total = w * C;
  for (int i = lane; i < left; i += 32) {
    const int xx = xa + i / C;
    dst[i] = (rowin && xx >= 0 && xx < nx) ? __ldg(src + i) : fill;
  }
  for (int i = right0 + lane; i < total; i += 32) {
    const int xx = xa + i / C;
    dst[i] = (rowin && xx >= 0 && xx < nx) ? __ldg(src + i) : fill;
  }
}

template <int C, int DH>
__global__ void __launch_bounds__(kThreads, 2) ica_iterate_kernel(const IterParams P) {
  constexpr int K = RowVals<DH>::K;
  constexpr int NP = RowVals<DH>::NP;
  constexpr int HW = DH + 1;        // x-powers kept for the Hessian moments
  constexpr int BWN = DH / 2 + 1;   // x-powers kept for the b moments
  constexpr int S1W = Stage<C>::S1W;
  constexpr int S2W = Stage<C>::S2W;
  constexpr int NENT = K * kYPow;

  extern __shared__ __align__(128) float smem[];
  float* const stage0 = smem;
  float* const stage1 = smem + Stage<C>::kFloats;
  int* const s_chunk_start = reinterpret_cast<int*>(smem + 2 * Stage<C>::kFloats);
  __shared__ __align__(8) unsigned long long s_bar[2];
  __shared__ PairCtx ctx;
  __shared__ StageCtl sctl[2];
  __shared__ unsigned int s_ticket;
  __shared__ int s_pair, s_chunk, s_flag;
  __shared__ double s_mom[kAccStride];
  __shared__ double s_aug[ICA_MAX_PARAMS][2 * ICA_MAX_PARAMS + 1];
  __shared__ double s_vec[2 * ICA_MAX_PARAMS];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int B = P.B;
  const int total_chunks = P.chunk_start[B];
  if ((int)blockIdx.x >= total_chunks) return;

  const bool sched_in_smem = B <= kSchedCap;
  if (sched_in_smem)
    for (int i = tid; i <= B; i += kThreads) s_chunk_start[i] = P.chunk_start[i];
  if (tid == 0) {
    mbar_init(&s_bar[0], 2);
    mbar_init(&s_bar[1], 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned phase0 = 0, phase1 = 0;

  const int delta = P.delta;
  const bool frame = P.frame != 0;
  const bool robust = P.robust_loop != 0;
  const float chm = P.ch_mult;
  const int rtype = P.robust_type;

  for (int item = blockIdx.x; item < total_chunks; item += gridDim.x) {
    // ---------------- decode the work item: pair and chunk, per-pair constants
    if (tid == 0) {
      const int* cs = sched_in_smem ? s_chunk_start : P.chunk_start;
      int lo_ = 0, hi_ = B;   // largest b with cs[b] <= item
      while (hi_ - lo_ > 1) { const int mid = (lo_ + hi_) >> 1; if (cs[mid] <= item) lo_ = mid; else hi_ = mid; }
      s_pair = lo_;
      s_chunk = item - cs[lo_];
      const PairState& st = P.state[lo_];
      ctx.pair = lo_;
      ctx.scale = st.scale;
      warp_matrix(st.p, st.ttype, ctx.m64);
      ctx.coef = make_warp_coef(ctx.m64);
      const MinMaxKeys mm = P.mm[(lo_ * P.nscales + st.scale) * 2 + 1];
      ctx.lo = key_float(mm.lo);
      ctx.hi = key_float(mm.hi);
      ctx.lambda2 = (float)(st.lambda_it * st.lambda_it);
      ctx.need_h = (P.robust_loop || st.iter == 0) ? 1 : 0;
    }
    __syncthreads();
    const int pair = s_pair, chunk = s_chunk;
    const int s = ctx.scale;
    const LevelDesc L = P.lv[s];
    const int nx = L.nx, ny = L.ny, pitch = L.pitch;
    const float* __restrict__ I1 = s == 0 ? P.I1_0 + (long long)pair * P.in_stride
                                          : P.pyr1 + (long long)pair * P.pyr_stride + L.offset;
    const float* __restrict__ I2 = s == 0 ? P.I2_0 + (long long)pair * P.in_stride
                                          : P.pyr2 + (long long)pair * P.pyr_stride + L.offset;
    const bool bulk_ok = (pitch & 3) == 0 && ((((unsigned long long)I1) | ((unsigned long long)I2)) & 15ull) == 0;
    const int ntiles = L.tiles_x * L.tiles_y;
    const int nch = ntiles < P.max_chunks ? ntiles : P.max_chunks;
    const int t_begin = (int)((long long)chunk * ntiles / nch);
    const int t_end = (int)((long long)(chunk + 1) * ntiles / nch);
    const bool need_h = ctx.need_h != 0;
    const WarpCoef coef = ctx.coef;
    const float lo = ctx.lo, hi = ctx.hi, lambda2 = ctx.lambda2;

    // ---------------- staging of one tile into one stage (warps 0 and 1 issue, the others go on)
    auto issue_tile = [&](int tile, int sidx) {
      const int x0 = (tile % L.tiles_x) * TW;
      const int y0 = (tile / L.tiles_x) * TH;
      float* s2 = sidx ? stage1 : stage0;
      float* s1 = s2 + BH_MAX * S2W;
      unsigned long long* bar = &s_bar[sidx];
      if (warp == 0) {
        // window of I2 reachable from this tile: project the four corners of its in-image part
        const int xe_ = min(x0 + TW, nx) - 1, ye_ = min(y0 + TH, ny) - 1;
        const int px = (lane & 1) ? xe_ : x0, py = (lane & 2) ? ye_ : y0;
        int cx, cy; float tx, ty;
        const bool ok = project_px(coef, ctx.m64, px, py, cx, cy, tx, ty);
        int mnx = cx, mxx = cx, mny = cy, mxy = cy, okall = ok ? 1 : 0;
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
          mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
          mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
          okall &= __shfl_xor_sync(0xffffffffu, okall, o);
        }
        const int bx0 = ((mnx - 2) >> 2) << 2;                   // floor to a multiple of 4 pixels
        const int bw = ((mxx + 3 - bx0 + 1) + 3) & ~3;
        const int by0 = mny - 2;
        const int bh = mxy + 3 - by0 + 1;
        const bool fits = okall && bw <= BW_MAX && bh <= BH_MAX && bw > 0 && bh > 0;
        if (lane == 0) { sctl[sidx].bx0 = bx0; sctl[sidx].by0 = by0; sctl[sidx].bw = bw; sctl[sidx].bh = bh; sctl[sidx].fits = fits ? 1 : 0; }
        RowPlan rp; rp.xs = bx0; rp.xe4 = bx0; rp.bytes = 0;
        if (fits && lane < bh) rp = plan_row<C>(by0 + lane, bx0, bw, nx, ny, bulk_ok);
        unsigned tot = rp.bytes;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0) { fence_proxy_async(); mbar_expect_tx_arrive(bar, tot); }
        __syncwarp();
        if (rp.bytes) bulk_g2s(s2 + lane * S2W + (rp.xs - bx0) * C, I2 + (long long)(by0 + lane) * pitch + (long long)rp.xs * C, rp.bytes, bar);
        if (fits) {
          const float qnan = __int_as_float(0x7fc00000);         // skimage cval outside the image
          unsigned need = __ballot_sync(0xffffffffu, lane < bh && (int)(rp.bytes / (4u * C)) != bw);
          while (need) {
            const int r = __ffs(need) - 1; need &= need - 1;
            const int rxs = __shfl_sync(0xffffffffu, rp.xs, r), rxe = __shfl_sync(0xffffffffu, rp.xe4, r);
            fill_row_rest<C>(s2 + r * S2W, I2, pitch, by0 + r, bx0, bw, nx, ny, rxs, rxe, qnan, lane);
          }
        }
      } else if (warp == 1) {
        const int xa = x0 - HALO;
        RowPlan rp; rp.xs = xa; rp.xe4 = xa; rp.bytes = 0;
        if (lane < S1ROWS) rp = plan_row<C>(y0 - 1 + lane, xa, S1PX, nx, ny, bulk_ok);
        unsigned tot = rp.bytes;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0) { fence_proxy_async(); mbar_expect_tx_arrive(bar, tot); }
        __syncwarp();
        if (rp.bytes) bulk_g2s(s1 + lane * S1W + (rp.xs - xa) * C, I1 + (long long)(y0 - 1 + lane) * pitch + (long long)rp.xs * C, rp.bytes, bar);
        unsigned need = __ballot_sync(0xffffffffu, lane < S1ROWS && (int)(rp.bytes / (4u * C)) != S1PX);
        while (need) {
          const int r = __ffs(need) - 1; need &= need - 1;
          const int rxs = __shfl_sync(0xffffffffu, rp.xs, r), rxe = __shfl_sync(0xffffffffu, rp.xe4, r);
          fill_row_rest<C>(s1 + r * S1W, I1, pitch, y0 - 1 + r, xa, S1PX, nx, ny, rxs, rxe, 0.0f, lane);
        }
      }
    };

    double acc[kYPow];
#pragma unroll
    for (int i = 0; i < kYPow; ++i) acc[i] = 0.0;

    issue_tile(t_begin, 0);
    __syncthreads();   // ordinary stores of the first tile's fill + its StageCtl are visible
    for (int tile = t_begin; tile < t_end; ++tile) {
      const int sidx = (tile - t_begin) & 1;
      if (tile + 1 < t_end) issue_tile(tile + 1, sidx ^ 1);      // overlaps with this tile's arithmetic
      if (sidx == 0) { mbar_wait(&s_bar[0], phase0); phase0 ^= 1; } else { mbar_wait(&s_bar[1], phase1); phase1 ^= 1; }
      const float* s2 = sidx ? stage1 : stage0;
      const float* s1 = s2 + BH_MAX * S2W;
      const int x0 = (tile % L.tiles_x) * TW;
      const int y0 = (tile / L.tiles_x) * TH;
      const int bx0 = sctl[sidx].bx0, by0 = sctl[sidx].by0, bw = sctl[sidx].bw, bh = sctl[sidx].bh;
      const bool fits = sctl[sidx].fits != 0;

#pragma unroll 1
      for (int rr = 0; rr < TH / kWarps; ++rr) {
        const int ly = warp + rr * kWarps;
        const int y = y0 + ly;
        float v[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) v[i] = 0.0f;
        if (y < ny) {
          const bool yin = !frame || (y >= delta && y < ny - delta);
#pragma unroll
          for (int half = 0; half < TW / 32; ++half) {
            const int lx = lane + half * 32;
            const int x = x0 + lx;
            if (x >= nx) continue;
            const bool inframe = yin && (!frame || (x >= delta && x < nx - delta));
            int cx, cy; float tx, ty;
            const bool pok = project_px(coef, ctx.m64, x, y, cx, cy, tx, ty);
            float wx[4], wy[4];
            keys_weights(tx, wx[0], wx[1], wx[2], wx[3]);
            keys_weights(ty, wy[0], wy[1], wy[2], wy[3]);
            const bool insm = fits && (cx - 1 >= bx0) && (cx + 2 < bx0 + bw) && (cy - 1 >= by0) && (cy + 2 < by0 + bh);
            const float* t2base = s2 + (cy - 1 - by0) * S2W + (cx - 1 - bx0) * C;
            const float* c1 = s1 + (ly + 1) * S1W + (lx + HALO) * C;
            const bool gxok = inframe && x >= 1 && x <= nx - 2;
            const bool gyok = inframe && y >= 1 && y <= ny - 2;
            float sxx = 0.f, sxy = 0.f, syy = 0.f, vx = 0.f, vy = 0.f, t2 = 0.f;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
              float iw;
              if (!pok) {
                iw = __int_as_float(0x7fc00000);
              } else if (insm) {
                float a = 0.0f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float* r = t2base + q * S2W + ch;
                  float hsum = wx[0] * r[0] + wx[1] * r[C] + wx[2] * r[2 * C] + wx[3] * r[3 * C];
                  a = fmaf(wy[q], hsum, a);
                }
                iw = a;
              } else {
                iw = sample_global<C>(I2, pitch, nx, ny, cx, cy, ch, wx, wy);
              }
              const bool valid = iw == iw;               // NaN footprint
              iw = fminf(fmaxf(iw, lo), hi);             // clip (only used when valid)
              const float i1c = c1[ch];
              const float gx = gxok ? 0.5f * (c1[ch + C] - c1[ch - C]) : 0.0f;
              const float gy = gyok ? 0.5f * (c1[ch + S1W] - c1[ch - S1W]) : 0.0f;
              const float di = valid ? iw - i1c : 0.0f;  // non-finite -> 0 (io.py:72, 134)
              if (need_h) { sxx = fmaf(gx, gx, sxx); sxy = fmaf(gx, gy, sxy); syy = fmaf(gy, gy, syy); }
              vx = fmaf(gx, di, vx); vy = fmaf(gy, di, vy);
              t2 = fmaf(di, di, t2);
            }
            // gray image standing for its x3 replication: every channel sum triples
            const float rho = robust ? rho_prime(t2 * chm, lambda2, rtype) : 1.0f;
            const float sc = rho * chm;
            const float xf = (float)x;
            if (need_h) {
              float wq[3] = {sc * sxx, sc * sxy, sc * syy};
#pragma unroll
              for (int q = 0; q < 3; ++q) {
                float xp = 1.0f;
#pragma unroll
                for (int a = 0; a < HW; ++a) { v[q * HW + a] = fmaf(wq[q], xp, v[q * HW + a]); xp *= xf; }
              }
            }
            {
              float uq[2] = {sc * vx, sc * vy};
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                float xp = 1.0f;
#pragma unroll
                for (int a = 0; a < BWN; ++a) { v[3 * HW + q * BWN + a] = fmaf(uq[q], xp, v[3 * HW + q * BWN + a]); xp *= xf; }
              }
            }
          }
        }
        // moment k of the row lands on lane (k << log2(32/NP)); fold in y^b in fp64
        const float tot = warp_transpose_reduce<NP>(v, lane);
        if (y < ny) {
          const double yd = (double)y, t = (double)tot;
          double yp = 1.0;
#pragma unroll
          for (int b = 0; b < kYPow; ++b) { acc[b] = fma(t, yp, acc[b]); yp *= yd; }
        }
      }
      __syncthreads();   // stage sidx is free again; fills of the next stage are visible
    }

    // ---------------- chunk partial: [K][kYPow] doubles, warps summed in fixed order
    double* red = reinterpret_cast<double*>(smem);   // 8 warps * NENT doubles; the stages are idle here
    constexpr int SH = (NP == 32) ? 0 : (NP == 16 ? 1 : 2);
    {
      const int k = lane >> SH;
      if ((lane & ((1 << SH) - 1)) == 0 && k < K) {
#pragma unroll
        for (int b = 0; b < kYPow; ++b) red[(warp * K + k) * kYPow + b] = acc[b];
      }
    }
    __syncthreads();
    {
      double* out = P.partials + ((long long)pair * P.max_chunks + chunk) * kAccStride;
      for (int i = tid; i < NENT; i += kThreads) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) sum += red[w * NENT + i];
        out[i] = sum;
      }
    }
    // ---------------- arrive; the block that delivers the pair's last chunk runs the epilogue
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&P.state[pair].ticket, 1u);
    __syncthreads();
    if (s_ticket != (unsigned)(nch - 1)) continue;   // block-uniform
    __threadfence();

    // ================= K3: reduce the chunks, solve, compose, schedule =================
    PairState& st = P.state[pair];
    {
      // fixed summation order: NG groups of consecutive chunks per entry, groups combined in order
      constexpr int NG = kThreads / NENT;            // 2 (DH=4), 3 (DH=2), 10 (DH=0)
      double* part = reinterpret_cast<double*>(smem);
      const int e = tid % NENT, gI = tid / NENT;
      if (gI < NG) {
        const int c0 = (int)((long long)gI * nch / NG), c1 = (int)((long long)(gI + 1) * nch / NG);
        const double* src = P.partials + (long long)pair * P.max_chunks * kAccStride + e;
        double sum = 0.0;
        int c = c0;
        for (; c + 4 <= c1; c += 4) {
          const double a0 = __ldcg(src + (long long)(c + 0) * kAccStride), a1 = __ldcg(src + (long long)(c + 1) * kAccStride);
          const double a2 = __ldcg(src + (long long)(c + 2) * kAccStride), a3 = __ldcg(src + (long long)(c + 3) * kAccStride);
          sum += a0; sum += a1; sum += a2; sum += a3;
        }
        for (; c < c1; ++c) sum += __ldcg(src + (long long)c * kAccStride);
        part[gI * NENT + e] = sum;
      }
      __syncthreads();
      if (tid < NENT) {
        double sum = 0.0;
#pragma unroll
        for (int g2 = 0; g2 < NG; ++g2) sum += part[g2 * NENT + tid];
        // quadratic loop after the first iteration of a scale: the H moments were not gathered
        s_mom[tid] = (!need_h && tid < 3 * HW * kYPow) ? 0.0 : sum;
      }
      __syncthreads();
    }
    const int ttype = st.ttype;
    const int n = nparams_of(ttype);
    // assemble H (n x n) and b (n) from the moments, one entry per thread (same sums as
    // ica_transform.cuh: assemble_system)
    if (tid < n * n + n) {
      Mono jx[ICA_MAX_PARAMS], jy[ICA_MAX_PARAMS];
      jacobian_monomials(ttype, jx, jy);
      constexpr int hw = DH + 1, bwn = DH / 2 + 1, boff = 3 * hw;
      if (tid < n * n) {
        const int k = tid / n, l = tid % n;
        double sum = 0.0;
        if (jx[k].coef && jx[l].coef) sum += (double)(jx[k].coef * jx[l].coef) * s_mom[(0 * hw + jx[k].a + jx[l].a) * kYPow + jx[k].b + jx[l].b];
        if (jx[k].coef && jy[l].coef) sum += (double)(jx[k].coef * jy[l].coef) * s_mom[(1 * hw + jx[k].a + jy[l].a) * kYPow + jx[k].b + jy[l].b];
        if (jy[k].coef && jx[l].coef) sum += (double)(jy[k].coef * jx[l].coef) * s_mom[(1 * hw + jy[k].a + jx[l].a) * kYPow + jy[k].b + jx[l].b];
        if (jy[k].coef && jy[l].coef) sum += (double)(jy[k].coef * jy[l].coef) * s_mom[(2 * hw + jy[k].a + jy[l].a) * kYPow + jy[k].b + jy[l].b];
        s_aug[k][l] = sum;
        s_aug[k][n + l] = (k == l) ? 1.0 : 0.0;
      } else {
        const int k = tid - n * n;
        double sum = 0.0;
        if (jx[k].coef) sum += (double)jx[k].coef * s_mom[(boff + 0 * bwn + jx[k].a) * kYPow + jx[k].b];
        if (jy[k].coef) sum += (double)jy[k].coef * s_mom[(boff + 1 * bwn + jy[k].a) * kYPow + jy[k].b];
        s_vec[k] = sum;
      }
    }
    if (tid == 0) s_flag = 0;
    __syncthreads();
    if (P.dbg_Hb) {  // parity hook (ica_hessian_b_host): export, leave the state untouched
      if (tid < n * n) P.dbg_Hb[tid] = s_aug[tid / n][tid % n];
      if (tid < n) P.dbg_Hb[64 + tid] = s_vec[tid];
      if (tid == 0) st.ticket = 0;
      __syncthreads();
      continue;
    }
    // de.inverse_hessian: Gauss-Jordan with partial pivoting on [H | I], one thread per entry
    // (element by element the arithmetic of ica_transform.cuh: inverse_hessian); zero matrix when
    // a pivot is exactly zero (np.linalg.LinAlgError branch, derivatives.py:127-129)
    if (need_h) {
      const int i = tid / (2 * ICA_MAX_PARAMS), j = tid % (2 * ICA_MAX_PARAMS);
      const bool act = tid < ICA_MAX_PARAMS * 2 * ICA_MAX_PARAMS && i < n && j < 2 * n;
      for (int k = 0; k < n; ++k) {
        if (tid == 0) {
          int piv = k; double best = fabs(s_aug[k][k]);
          for (int r = k + 1; r < n; ++r) { const double vv = fabs(s_aug[r][k]); if (vv > best) { best = vv; piv = r; } }
          if (!(best > 0.0)) s_flag = 1;
          s_chunk = piv;
        }
        __syncthreads();
        const int piv = s_chunk;
        if (s_flag) break;                       // block-uniform
        const bool swp = act && piv != k && (i == k || i == piv);
        double other = 0.0;
        if (swp) other = s_aug[i == k ? piv : k][j];
        __syncthreads();
        if (swp) s_aug[i][j] = other;
        __syncthreads();
        const double inv = 1.0 / s_aug[k][k];
        __syncthreads();
        if (act && i == k) s_aug[k][j] *= inv;
        __syncthreads();
        const double f = act ? s_aug[i][k] : 0.0;
        const double pk = act ? s_aug[k][j] : 0.0;
        __syncthreads();
        if (act && i != k && f != 0.0) s_aug[i][j] -= f * pk;
        __syncthreads();
      }
      __syncthreads();
      if (tid < n * n) st.hinv[tid] = s_flag ? 0.0 : s_aug[tid / n][n + tid % n];
      __syncthreads();
    }
    if (tid < n) {                                 // io.parametric_solve (io.py:146-155)
      double a = 0.0;
      for (int j = 0; j < n; ++j) a += st.hinv[tid * n + j] * s_vec[j];
      s_vec[ICA_MAX_PARAMS + tid] = a;
    }
    __syncthreads();
    if (tid == 0) {
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
      const int it = st.iter + 1;
      st.err = err;
      st.lambda_it = lam;
      st.total_iters += 1;
      if (P.traj && st.traj_count < P.traj_cap) {
        double* t = P.traj + ((long long)pair * P.traj_cap + st.traj_count) * ICA_TRAJ_STRIDE;
        t[0] = s; t[1] = it - 1; t[2] = err; t[3] = lam;
        for (int i = 0; i < ICA_MAX_PARAMS; ++i) t[4 + i] = i < n ? st.p[i] : 0.0;
        st.traj_count += 1;
      }
      if (err > P.tol && it < P.max_iter) {
        st.iter = it;
      } else {  // this scale is done (ica.py:109, 225)
        st.iters_per_scale[s] = it;
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
      st.ticket = 0;
    }
    __syncthreads();
  }
}

// Work list of the next launch: chunk_start[b] = exclusive prefix sum of chunks per pair,
// chunk_start[B] = total; also publishes the number of unfinished pairs.  One block.
__global__ void __launch_bounds__(1024) ica_schedule_kernel(const IterParams P) {
  __shared__ int s_warp[32];
  __shared__ int s_carry, s_act;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = P.B;
  if (tid == 0) { s_carry = 0; s_act = 0; }
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    const int b = base + tid;
    int c = 0;
    if (b < B) {
      const int s = P.state[b].scale;
      if (s >= 0) { const int nt = P.lv[s].tiles_x * P.lv[s].tiles_y; c = nt < P.max_chunks ? nt : P.max_chunks; }
    }
    const unsigned actmask = __ballot_sync(0xffffffffu, c > 0);
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int w = s_warp[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
      s_warp[lane] = wi - w;   // exclusive
    }
    __syncthreads();
    const int excl = s_carry + s_warp[warp] + incl - c;
    if (b < B) P.chunk_start[b] = excl;
    if (lane == 0 && actmask) atomicAdd(&s_act, __popc(actmask));
    __syncthreads();
    if (tid == 1023) s_carry = excl + c;
    __syncthreads();
  }
  if (tid == 0) { P.chunk_start[B] = s_carry; *P.n_active = s_act; }
}

// Resets the per-pair state at the start of a run (ica.py:319-337: ps[0] = p, ps[s>0] = 0;
// the coarse-to-fine loop starts at the coarsest scale).
__global__ void ica_init_state_kernel(PairState* state, const double* p_in, const int* ttypes,
                                      int B, int nscales, double lambda_cfg, int* n_active) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) *n_active = B;
  if (b >= B) return;
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
                                    float* __restrict__ Iw, float* __restrict__ DI) {
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
    if (iw == iw) iw = fminf(fmaxf(iw, slo), shi);
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

template <int C, int DH>
cudaError_t launch_iterate_t(const IterParams& P, int grid, size_t smem, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ica_iterate_kernel<C, DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(2 * Stage<C>::kFloats * sizeof(float) + (kSchedCap + 1) * sizeof(int)));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  ica_iterate_kernel<C, DH><<<grid, kThreads, smem, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace

int iterate_tile_w() { return TW; }
int iterate_tile_h() { return TH; }

cudaError_t launch_schedule(const IterParams& P, cudaStream_t stream) {
  ica_schedule_kernel<<<1, 1024, 0, stream>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_iterate(const IterParams& P, int channels, int dh, int grid, cudaStream_t stream) {
  const int nsched = (P.B <= kSchedCap ? P.B : 0) + 1;
  const size_t per_stage = (size_t)(channels == 3 ? Stage<3>::kFloats : Stage<1>::kFloats) * sizeof(float);
  const size_t smem = 2 * per_stage + (size_t)nsched * sizeof(int);
  if (channels == 3) {
    if (dh == 4) return launch_iterate_t<3, 4>(P, grid, smem, stream);
    if (dh == 2) return launch_iterate_t<3, 2>(P, grid, smem, stream);
    return launch_iterate_t<3, 0>(P, grid, smem, stream);
  }
  if (dh == 4) return launch_iterate_t<1, 4>(P, grid, smem, stream);
  if (dh == 2) return launch_iterate_t<1, 2>(P, grid, smem, stream);
  return launch_iterate_t<1, 0>(P, grid, smem, stream);
}

cudaError_t launch_init_state(PairState* state, const double* p_in, const int* ttypes, int B, int nscales,
                              double lambda_cfg, int* n_active, cudaStream_t stream) {
  ica_init_state_kernel<<<(B + 127) / 128, 128, 0, stream>>>(state, p_in, ttypes, B, nscales, lambda_cfg, n_active);
  return cudaGetLastError();
}

cudaError_t launch_export_results(const PairState* state, int B, double* p_out, double* err_out, int* iters_out,
                                  int nscales, cudaStream_t stream) {
  ica_export_results_kernel<<<(B + 127) / 128, 128, 0, stream>>>(state, B, p_out, err_out, iters_out, nscales);
  return cudaGetLastError();
}

cudaError_t launch_warp_out(const float* I1_0, const float* I2_0, long long in_stride, int nx, int ny, int channels,
                            const PairState* state, const MinMaxKeys* mm, int nscales, int B, float* Iw, float* DI,
                            cudaStream_t stream) {
  dim3 block(32, 8), grid((nx + 31) / 32, (ny + 7) / 8, B);
  if (channels == 3)
    ica_warp_out_kernel<3><<<grid, block, 0, stream>>>(I1_0, I2_0, in_stride, nx, ny, nx * 3, state, mm, nscales, Iw, DI);
  else
    ica_warp_out_kernel<1><<<grid, block, 0, stream>>>(I1_0, I2_0, in_stride, nx, ny, nx, state, mm, nscales, Iw, DI);
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
