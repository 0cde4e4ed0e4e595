// K2 + K3: the fused per-iteration kernel of the inverse compositional loop.
//
// One launch = one iteration of EVERY still-active image pair of the batch, each at its own
// scale.  Replaces, per iteration (src/inverse_compositional_algorithm.py:109-131, 225-259):
//   bi.bicubic_interpolation_skimage  (warp of I2 by p, Catmull-Rom, NaN footprint, clip)
//   DI = Iw - I1
//   io.robust_error_function          (rho'(sum_c DI_c^2))
//   io.independent_vector[_robust]    (b)
//   de.hessian_robust / de.hessian    (H; gradients, Jacobian and steepest-descent images are
//                                      recomputed per pixel, never stored: ica.py:81-100)
// and, in the last block of each pair to finish (ticket pattern, no spinning):
//   de.inverse_hessian, io.parametric_solve, tr.update_transform, the lambda schedule, the
//   stopping rule and zm.zoom_in_parameters at a scale change.
//
// Bound: HBM (read I1 once + I2 once per pixel-iteration = 2*C*4 bytes); no tensor cores.
#include "ica_device.cuh"
#include "ica_transform.cuh"
#include "ica_iterate.cuh"

namespace ica {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int TW = 64;           // tile width  (2 pixels per lane: x0+lane, x0+32+lane)
constexpr int TH = 16;           // tile height (2 rows per warp: y0+warp, y0+8+warp)
constexpr int BW_MAX = 80;       // staged I2 window (pixels); larger windows fall back to global
constexpr int BH_MAX = 32;

template <int DH> struct RowVals { static constexpr int K = 3 * (DH + 1) + 2 * (DH / 2 + 1);
                                   static constexpr int NP = K > 16 ? 32 : (K > 8 ? 16 : 8); };

struct BlockCtl {
  double m64[9];         // warp matrix in fp64 (tie-break path of project_px)
  WarpCoef coef;
  float lo, hi;          // clip range of I2 at this level
  float lambda2;
  int scale, iter, ttype;
  int need_h;
  int bx0, by0, bw, bh, fits;
  unsigned int ticket;
};

template <int C, int DH>
__global__ void __launch_bounds__(kThreads, 2) ica_iterate_kernel(const IterParams P) {
  constexpr int K = RowVals<DH>::K;
  constexpr int NP = RowVals<DH>::NP;
  constexpr int HW = DH + 1;        // x-powers kept for the Hessian moments
  constexpr int BWN = DH / 2 + 1;   // x-powers kept for the b moments
  constexpr int S1W = (TW + 2) * C;
  constexpr int S2W = BW_MAX * C;

  __shared__ float s1[(TH + 2) * S1W];
  __shared__ __align__(16) float s2[BH_MAX * S2W];
  __shared__ BlockCtl ctl;

  const int pair = blockIdx.y;
  const int g = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  PairState& st = P.state[pair];

  if (tid == 0) {
    ctl.scale = st.scale;
    if (ctl.scale >= 0) {
      warp_matrix(st.p, st.ttype, ctl.m64);
      ctl.coef = make_warp_coef(ctl.m64);
      const MinMaxKeys mm = P.mm[(pair * P.nscales + ctl.scale) * 2 + 1];
      ctl.lo = key_float(mm.lo);
      ctl.hi = key_float(mm.hi);
      ctl.lambda2 = (float)(st.lambda_it * st.lambda_it);
      ctl.iter = st.iter;
      ctl.ttype = st.ttype;
      ctl.need_h = (P.robust_loop || st.iter == 0) ? 1 : 0;
    }
  }
  __syncthreads();
  const int s = ctl.scale;
  if (s < 0) return;  // finished pair: nothing reads or writes its state any more

  const LevelDesc L = P.lv[s];
  const int nx = L.nx, ny = L.ny, pitch = L.pitch;
  const float* __restrict__ I1 = s == 0 ? P.I1_0 + (long long)pair * P.in_stride
                                        : P.pyr1 + (long long)pair * P.pyr_stride + L.offset;
  const float* __restrict__ I2 = s == 0 ? P.I2_0 + (long long)pair * P.in_stride
                                        : P.pyr2 + (long long)pair * P.pyr_stride + L.offset;
  const int ntiles = L.tiles_x * L.tiles_y;
  const int nblk = ntiles < P.G ? ntiles : P.G;
  const bool need_h = ctl.need_h != 0;
  const bool robust = P.robust_loop != 0;
  const WarpCoef coef = ctl.coef;
  const float lo = ctl.lo, hi = ctl.hi, lambda2 = ctl.lambda2;
  const int delta = P.delta;
  const bool frame = P.frame != 0;
  const float chm = P.ch_mult;

  double acc[kYPow];
#pragma unroll
  for (int i = 0; i < kYPow; ++i) acc[i] = 0.0;

  for (int tile = g; tile < ntiles; tile += nblk) {
    const int x0 = (tile % L.tiles_x) * TW;
    const int y0 = (tile / L.tiles_x) * TH;
    __syncthreads();  // previous tile's readers are done with s1/s2/ctl
    // ---- window of I2 touched by this tile: project the corners of the in-image part
    if (tid < 4) {
      int xe = min(x0 + TW, nx) - 1, ye = min(y0 + TH, ny) - 1;
      int px = (tid & 1) ? xe : x0, py = (tid & 2) ? ye : y0;
      int cx, cy; float tx, ty;
      bool ok = project_px(coef, ctl.m64, px, py, cx, cy, tx, ty);
      int mnx = cx, mxx = cx, mny = cy, mxy = cy, okall = ok ? 1 : 0;
#pragma unroll
      for (int o = 1; o < 4; o <<= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xfu, mnx, o)); mxx = max(mxx, __shfl_xor_sync(0xfu, mxx, o));
        mny = min(mny, __shfl_xor_sync(0xfu, mny, o)); mxy = max(mxy, __shfl_xor_sync(0xfu, mxy, o));
        okall &= __shfl_xor_sync(0xfu, okall, o);
      }
      if (tid == 0) {
        ctl.bx0 = mnx - 2; ctl.by0 = mny - 2;
        ctl.bw = mxx + 3 - ctl.bx0 + 1; ctl.bh = mxy + 3 - ctl.by0 + 1;
        ctl.fits = (okall && ctl.bw <= BW_MAX && ctl.bh <= BH_MAX) ? 1 : 0;
      }
    }
    // ---- stage the I1 tile with a 1-pixel halo (gradients), zero outside the image
    for (int r = warp; r < TH + 2; r += kWarps) {
      const int yy = y0 - 1 + r;
      const bool rowin = yy >= 0 && yy < ny;
      const float* src = I1 + (long long)yy * pitch + (x0 - 1) * C;
      for (int i = lane; i < S1W; i += 32) {
        const int xx = x0 - 1 + i / C;
        float v = 0.0f;
        if (rowin && xx >= 0 && xx < nx) v = __ldg(src + i);
        s1[r * S1W + i] = v;
      }
    }
    __syncthreads();
    const int bx0 = ctl.bx0, by0 = ctl.by0, bw = ctl.bw, bh = ctl.bh;
    const bool fits = ctl.fits != 0;
    if (fits) {
      const int roww = bw * C;
      for (int r = warp; r < bh; r += kWarps) {
        const int yy = by0 + r;
        const bool rowin = yy >= 0 && yy < ny;
        const float* src = I2 + (long long)yy * pitch + (long long)bx0 * C;
        for (int i = lane; i < roww; i += 32) {
          const int xx = bx0 + i / C;
          float v = __int_as_float(0x7fc00000);  // NaN = skimage cval outside the image
          if (rowin && xx >= 0 && xx < nx) v = __ldg(src + i);
          s2[r * S2W + i] = v;
        }
      }
    }
    __syncthreads();

    // ---- two image rows per warp
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
          // projected position and taps
          int cx, cy; float tx, ty;
          const bool pok = project_px(coef, ctl.m64, x, y, cx, cy, tx, ty);
          float wx[4], wy[4];
          keys_weights(tx, wx[0], wx[1], wx[2], wx[3]);
          keys_weights(ty, wy[0], wy[1], wy[2], wy[3]);
          const bool insm = fits && (cx - 1 >= bx0) && (cx + 2 < bx0 + bw) && (cy - 1 >= by0) &&
                            (cy + 2 < by0 + bh);
          const float* t2base = s2 + (cy - 1 - by0) * S2W + (cx - 1 - bx0) * C;
          const float* c1 = s1 + (ly + 1) * S1W + (lx + 1) * C;
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
          const float rho = robust ? rho_prime(t2 * chm, lambda2, P.robust_type) : 1.0f;
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
      // value k of the row lands on lane (k << log2(32/NP)); fold in y^b in fp64
      const float tot = warp_transpose_reduce<NP>(v, lane);
      if (y < ny) {
        const double yd = (double)y, t = (double)tot;
        double yp = 1.0;
#pragma unroll
        for (int b = 0; b < kYPow; ++b) { acc[b] = fma(t, yp, acc[b]); yp *= yd; }
      }
    }
  }

  // ---- block partial: [K][kYPow] doubles, warps summed in fixed order
  __syncthreads();
  double* red = reinterpret_cast<double*>(s2);  // 8 warps * K * 5 doubles <= 6720 B
  constexpr int SH = (NP == 32) ? 0 : (NP == 16 ? 1 : 2);
  if (g < nblk) {
    const int k = lane >> SH;
    if ((lane & ((1 << SH) - 1)) == 0 && k < K) {
#pragma unroll
      for (int b = 0; b < kYPow; ++b) red[(warp * K + k) * kYPow + b] = acc[b];
    }
    __syncthreads();
    double* out = P.partials + ((long long)pair * P.G + g) * kAccStride;
    for (int i = tid; i < K * kYPow; i += kThreads) {
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) sum += red[w * K * kYPow + i];
      out[i] = sum;
    }
  }
  // ---- arrive; the last of the pair's G blocks runs the solve/update epilogue
  __threadfence();
  __syncthreads();
  if (tid == 0) ctl.ticket = atomicAdd(&st.ticket, 1u);
  __syncthreads();
  if (ctl.ticket != (unsigned)(P.G - 1)) return;
  __threadfence();

  double* mom = reinterpret_cast<double*>(s2) + 1024;  // past `red`
  for (int i = tid; i < K * kYPow; i += kThreads) {
    const double* src = P.partials + (long long)pair * P.G * kAccStride + i;
    double sum = 0.0;
    for (int b = 0; b < nblk; ++b) sum += __ldcg(src + (long long)b * kAccStride);
    mom[i] = sum;
  }
  __syncthreads();
  if (tid != 0) return;

  // ================= K3: solve, compose, schedule (single thread, fp64) =================
  const int ttype = st.ttype;
  const int n = nparams_of(ttype);
  double H[ICA_MAX_PARAMS * ICA_MAX_PARAMS], bvec[ICA_MAX_PARAMS], dp[ICA_MAX_PARAMS];
  if (!need_h) {  // quadratic loop after the first iteration: moments of H were not gathered
    for (int i = 0; i < 3 * HW * kYPow; ++i) mom[i] = 0.0;
  }
  assemble_system(mom, DH, ttype, H, bvec);
  if (P.dbg_Hb) {  // parity hook (ica_hessian_b_host): export, do not touch the state
    for (int i = 0; i < n * n; ++i) P.dbg_Hb[i] = H[i];
    for (int i = 0; i < n; ++i) P.dbg_Hb[64 + i] = bvec[i];
    st.ticket = 0;
    return;
  }
  double* hinv = st.hinv;
  if (need_h) inverse_hessian(H, n, hinv);  // robust: every iteration; quadratic: once per scale
  double e2 = 0.0;
  for (int i = 0; i < n; ++i) {             // io.parametric_solve (io.py:146-155)
    double a = 0.0;
    for (int j = 0; j < n; ++j) a += hinv[i * n + j] * bvec[j];
    dp[i] = a; e2 += a * a;
  }
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
  if (P.traj) {
    double* t = P.traj + ((long long)pair * P.traj_cap + st.traj_count) * ICA_TRAJ_STRIDE;
    if (st.traj_count < P.traj_cap) {
      t[0] = s; t[1] = it - 1; t[2] = err; t[3] = lam;
      for (int i = 0; i < ICA_MAX_PARAMS; ++i) t[4 + i] = i < n ? st.p[i] : 0.0;
      st.traj_count += 1;
    }
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
      atomicSub(P.n_active, 1);
    }
  }
  st.ticket = 0;
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

}  // namespace

int iterate_tile_w() { return TW; }
int iterate_tile_h() { return TH; }

cudaError_t launch_iterate(const IterParams& P, int B, int channels, int dh, cudaStream_t stream) {
  dim3 grid(P.G, B), block(kThreads);
#define ICA_LAUNCH(CC, DD) ica_iterate_kernel<CC, DD><<<grid, block, 0, stream>>>(P)
  if (channels == 3) {
    if (dh == 4) ICA_LAUNCH(3, 4); else if (dh == 2) ICA_LAUNCH(3, 2); else ICA_LAUNCH(3, 0);
  } else {
    if (dh == 4) ICA_LAUNCH(1, 4); else if (dh == 2) ICA_LAUNCH(1, 2); else ICA_LAUNCH(1, 0);
  }
#undef ICA_LAUNCH
  return cudaGetLastError();
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
