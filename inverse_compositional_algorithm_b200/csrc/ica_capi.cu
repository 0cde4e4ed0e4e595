// C-ABI of libica_b200.so (see include/ica_b200.h): plan management and entry points.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <map>
#include <mutex>
#include <thread>
#include <chrono>
#include <vector>
#include <cuda.h>
#include "ica_common.cuh"
#include "ica_transform.cuh"
#include "ica_device.cuh"
#include "ica_iterate.cuh"
#include "ica_pyramid.cuh"

namespace ica {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace ica

using namespace ica;

#define ICA_LAUNCH_CHECK(expr)                                                      \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ica::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),       \
                     __FILE__, __LINE__);                                           \
      return ICA_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

struct ica_plan {
  ica_config cfg;
  int B, H, W, C, nscales, dh;
  int device = 0;                    // the CUDA device the plan's buffers, stream and graph live on
  int max_chunks = 0, grid = 0;
  int march = 0;                     // 1: column-march K2 (ica_march.cu, opt-in: ICA_K2=march); 0: the tile kernel
  int chunk_m = 2;                   // march kernel: preferred tiles per consumer warp and chunk (ICA_CHUNK_M)
  int shard_rank = 0, shard_n = 1;   // row-sharded mode (ica_plan_set_row_shard)
  int* chunk_start = nullptr;
  int* item_pair = nullptr;
  unsigned int* solve_ticket = nullptr;
  int* loop_count = nullptr;
  long long* kernel_ns = nullptr;   // [2] accumulated streaming-phase time of the iterate kernel (ns), launches
  SchedHdr* hdr = nullptr;          // [2] headers of the double-buffered work lists
  unsigned int* pair_ticket = nullptr;   // [B]
  // row-sharded mode with the exchange in the device-side loop (peer memory, CUDA IPC)
  double* xbuf = nullptr;            // this rank's exchange buffer [2][world][B][kXSlot]
  double** x_peers_dev = nullptr;    // device array [world] of every rank's buffer as mapped here
  void* x_peer_ptr[64] = {nullptr};  // host copy (peers opened with cudaIpcOpenMemHandle; own buffer at [rank])
  int x_world = 0, x_rank = 0;
  unsigned long long x_seq = 0;      // sequence numbers consumed so far
  unsigned long long* x_seq_dev = nullptr;   // device: sequence base of the current run
  unsigned long long* x_seq_host = nullptr;  // pinned staging of it
  int* x_error = nullptr;
  long long* x_ns = nullptr;         // [2] device: accumulated exchange time, exchanges
  int graph_mode = -1;               // solve mode the loop graph was built for
  int fused = 0;                    // solve inside the iterate kernel (one launch per iteration; opt-in, ICA_FUSE=1)
  int* h_loop = nullptr;            // pinned: {iterations, pad} and kernel ns copied after a run
  long long* h_kns = nullptr;       // pinned [2]
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  const float *graph_I1 = nullptr, *graph_I2 = nullptr;
  int graph_dh = -1;
  int use_graph = 1;
  AsmEntry* asm_tab = nullptr;
  int asm_dh = -1;
  long long* dbg_time = nullptr;   // optional per-CTA timeline (ica_plan_debug_timeline)
  LevelDesc lv[ICA_MAX_SCALES];
  long long in_stride = 0, pyr_stride = 0;
  float *pyr1 = nullptr, *pyr2 = nullptr;
  float* tmp = nullptr;
  long long tmp_stride = 0;
  int tmp_images = 0;
  long long tmp_floats = 0;
  DeviceResample ry[ICA_MAX_SCALES], rx[ICA_MAX_SCALES];
  PairState* state = nullptr;
  MinMaxKeys* mm = nullptr;
  MinMaxKeys* mm_open = nullptr;    // IPOL pyramid (no clip): open parent ranges
  MinMaxKeys* gather_keys = nullptr;   // open ranges for the border kernels of a gathering level (same layout as mm)
  int* clip_flags = nullptr;           // [2B] deferred-clip decisions
  double* partials = nullptr;
  double* traj = nullptr;
  int traj_cap = 0;
  int* n_active = nullptr;
  int* h_n_active = nullptr;  // pinned
  int* ttypes_dev = nullptr;
  std::vector<int> ttypes;
  double* p_dev = nullptr;
  double* err_dev = nullptr;
  int* iters_dev = nullptr;
  // host-entry staging
  float *in1_dev = nullptr, *in2_dev = nullptr;
  void* raw_dev = nullptr;
  size_t raw_bytes = 0;
  float *DI_dev = nullptr, *Iw_dev = nullptr;
  double* out64_dev = nullptr;       // float64 staging of DI and Iw (host entry with ICA_DTYPE_OUT_F64), 2 x nimg doubles
  const float *last_I1 = nullptr, *last_I2 = nullptr;
  // K2 stages its tiles with tiled TMA copies: one tensor map per (pair, level, image)
  void* tmaps_dev = nullptr;                 // CUtensorMap [B][nscales][2], 128 bytes each
  std::vector<unsigned char> tmaps_host;
  const float *tm_I1 = nullptr, *tm_I2 = nullptr;   // level-0 images the maps currently describe
  float *pad1 = nullptr, *pad2 = nullptr;    // level-0 copies with 16-byte rows (only when the caller's rows are not)
  bool mm_ready = false;                     // the keys were reset and level 0's min/max filled by the ingest conversion
  const float *k2_I1 = nullptr, *k2_I2 = nullptr;   // K2's view of level 0 in the current run
  long long k2_stride = 0;
  int k2_pitch = 0;
  cudaStream_t stream = nullptr;  // own stream of the host entry
  size_t device_bytes = 0;
  long long launches = 0;
  int last_launches_per_iter = 2;    // 1 when the last run used the fused solve
  // optional timing
  int timing = 0;
  cudaEvent_t ev_host0 = nullptr, ev_host1 = nullptr, ev_h2d_done = nullptr;  // bracket ica_plan_run_host on its stream
  std::vector<cudaEvent_t> ev_iter, ev_pyr;
  int n_ev_iter = 0, n_ev_pyr = 0;
};

namespace {

template <typename T>
int dev_alloc(ica_plan* pl, T** ptr, size_t count) {
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  cudaError_t e = cudaMalloc((void**)ptr, bytes);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    return ICA_ERR_ALLOC;
  }
  if (pl) pl->device_bytes += bytes;
  return ICA_OK;
}

int upload_resample(ica_plan* pl, const Resample1D& r, DeviceResample* d) {
  d->n_in = r.n_in; d->n_out = r.n_out; d->taps = r.taps;
  d->start_host = r.start;
  detect_uniform_rows(r, &d->fast);
  int rc;
  if ((rc = dev_alloc(pl, &d->start, r.start.size()))) return rc;
  if ((rc = dev_alloc(pl, &d->weights, r.weights.size()))) return rc;
  if ((rc = dev_alloc(pl, &d->weights_t, r.weights.size()))) return rc;
  std::vector<float> wt(r.weights.size());
  for (int o = 0; o < r.n_out; ++o)
    for (int k = 0; k < r.taps; ++k) wt[(size_t)k * r.n_out + o] = r.weights[(size_t)o * r.taps + k];
  ICA_CUDA_CHECK(cudaMemcpy(d->start, r.start.data(), r.start.size() * sizeof(int), cudaMemcpyHostToDevice));
  ICA_CUDA_CHECK(cudaMemcpy(d->weights, r.weights.data(), r.weights.size() * sizeof(float), cudaMemcpyHostToDevice));
  ICA_CUDA_CHECK(cudaMemcpy(d->weights_t, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
  return ICA_OK;
}

int upload_assembly(ica_plan* pl) {
  if (pl->asm_dh == pl->dh) return ICA_OK;
  std::vector<AsmEntry> tab((size_t)6 * 72);
  build_assembly_table(pl->dh, tab.data());
  ICA_CUDA_CHECK(cudaMemcpy(pl->asm_tab, tab.data(), tab.size() * sizeof(AsmEntry), cudaMemcpyHostToDevice));
  pl->asm_dh = pl->dh;
  return ICA_OK;
}

void free_resample(DeviceResample* d) {
  cudaFree(d->start); cudaFree(d->weights); cudaFree(d->weights_t);
  d->start = nullptr; d->weights = nullptr; d->weights_t = nullptr;
}

int validate_config(const ica_config* c) {
  if (!c) { set_error("config is NULL"); return ICA_ERR_INVALID; }
  if (c->batch < 1 || c->batch >= (1 << 20)) { set_error("batch must be in [1, 2^20)"); return ICA_ERR_INVALID; }
  if (c->height < 1 || c->width < 1) { set_error("image shape must be positive"); return ICA_ERR_INVALID; }
  if (c->channels != 1 && c->channels != 3) { set_error("channels must be 1 or 3"); return ICA_ERR_INVALID; }
  if (c->nscales < 1 || c->nscales > ICA_MAX_SCALES) { set_error("nscales must be in [1, %d]", ICA_MAX_SCALES); return ICA_ERR_INVALID; }
  if (c->nscales > 1 && !(c->nu > 0.0 && c->nu < 1.0)) { set_error("nu must be in (0, 1)"); return ICA_ERR_INVALID; }
  if (nparams_of(c->transform_type) < 0) { set_error("Unknown transform type"); return ICA_ERR_INVALID; }
  if (c->robust_type < 0 || c->robust_type > 4) { set_error("Unknown type for robust error function"); return ICA_ERR_INVALID; }
  if (!(c->tol < 0.01)) { set_error("TOL must be positive and very small (less than 0.01)"); return ICA_ERR_INVALID; }
  if (c->max_iter < 1) { set_error("max_iter must be >= 1"); return ICA_ERR_INVALID; }
  if (c->delta < 0) { set_error("delta must be >= 0"); return ICA_ERR_INVALID; }
  if ((c->flags & ICA_FLAG_IPOL_WARP) && c->delta < 2) {
    set_error("the IPOL warp domain needs delta >= 2 on the registration path (clamped neighbours are only in the bicubic_interpolation_image helper)");
    return ICA_ERR_INVALID;
  }
  return ICA_OK;
}

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n < 1) {
    set_error("no CUDA device is visible (%s); libica_b200 has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return ICA_ERR_NO_DEVICE;
  }
  return ICA_OK;
}

// a plan belongs to the device that was current when it was created: make it current for the calling thread
int enter_plan_device(const ica_plan* pl) {
  int cur = -1;
  if (cudaGetDevice(&cur) == cudaSuccess && cur == pl->device) return ICA_OK;
  ICA_CUDA_CHECK(cudaSetDevice(pl->device));
  return ICA_OK;
}

// ---- tensor maps (driver entry point fetched through the runtime: no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// float32 image [ny][pitch] of which nx*C floats per row are valid; box = boxw floats x boxh rows
int encode_image_map(void* out128, const float* base, int nx, int ny, int C, int pitch, int boxw, int boxh, bool nan_fill) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return ICA_ERR_CUDA; }
  alignas(64) CUtensorMap m;
  const cuuint64_t gdim[2] = {(cuuint64_t)nx * C, (cuuint64_t)ny};
  const cuuint64_t gstr[1] = {(cuuint64_t)pitch * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)boxw, (cuuint32_t)boxh};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for a %d x %d x %d image, pitch %d, box %d x %d", (int)r, nx, ny, C, pitch, boxw, boxh);
    return ICA_ERR_CUDA;
  }
  static_assert(sizeof(CUtensorMap) == 128, "tensor maps are passed as 128-byte records");
  memcpy(out128, &m, 128);
  return ICA_OK;
}

// maps of levels [s_lo, s_hi) of every pair into the host copy
int encode_level_maps(ica_plan* pl, int s_lo, int s_hi) {
  int w1, h1, w2, h2;
  if (pl->march) march_stage_boxes(pl->C, &w1, &h1, &w2, &h2); else iterate_stage_boxes(pl->C, &w1, &h1, &w2, &h2);
  for (int b = 0; b < pl->B; ++b)
    for (int s = s_lo; s < s_hi; ++s) {
      const LevelDesc& L = pl->lv[s];
      const float* i1 = s == 0 ? pl->k2_I1 + (long long)b * pl->k2_stride : pl->pyr1 + (long long)b * pl->pyr_stride + L.offset;
      const float* i2 = s == 0 ? pl->k2_I2 + (long long)b * pl->k2_stride : pl->pyr2 + (long long)b * pl->pyr_stride + L.offset;
      const int pitch = s == 0 ? pl->k2_pitch : L.pitch;
      unsigned char* rec = pl->tmaps_host.data() + ((size_t)(b * pl->nscales + s) * 2) * 128;
      if (int rc = encode_image_map(rec, i1, L.nx, L.ny, pl->C, pitch, w1, h1, false)) return rc;
      // (the march kernel evaluates the NaN footprint analytically: its zero-padded weights must never meet a NaN)
      if (int rc = encode_image_map(rec + 128, i2, L.nx, L.ny, pl->C, pitch, w2, h2, !pl->march)) return rc;
    }
  return ICA_OK;
}

// K2's view of level 0 (the caller's images, or padded copies when their rows are not multiples of 16 bytes) and the
// tensor maps that go with it; re-encoded only when the images moved.
int prepare_level0(ica_plan* pl, const float* I1, const float* I2, cudaStream_t stream) {
  const int row = pl->W * pl->C;
  const bool aligned = (row % 4 == 0) && (((reinterpret_cast<unsigned long long>(I1) | reinterpret_cast<unsigned long long>(I2)) & 15ull) == 0);
  if (aligned) {
    pl->k2_I1 = I1; pl->k2_I2 = I2; pl->k2_pitch = row; pl->k2_stride = pl->in_stride;
  } else {
    const int pitch = (row + 3) / 4 * 4;
    const long long stride = (long long)pitch * pl->H;
    if (!pl->pad1) {
      if (int rc = dev_alloc(pl, &pl->pad1, (size_t)pl->B * stride)) return rc;
      if (int rc = dev_alloc(pl, &pl->pad2, (size_t)pl->B * stride)) return rc;
    }
    ICA_CUDA_CHECK(cudaMemcpy2DAsync(pl->pad1, (size_t)pitch * 4, I1, (size_t)row * 4, (size_t)row * 4, (size_t)pl->B * pl->H,
                                     cudaMemcpyDeviceToDevice, stream));
    ICA_CUDA_CHECK(cudaMemcpy2DAsync(pl->pad2, (size_t)pitch * 4, I2, (size_t)row * 4, (size_t)row * 4, (size_t)pl->B * pl->H,
                                     cudaMemcpyDeviceToDevice, stream));
    pl->k2_I1 = pl->pad1; pl->k2_I2 = pl->pad2; pl->k2_pitch = pitch; pl->k2_stride = stride;
  }
  if (pl->tm_I1 != pl->k2_I1 || pl->tm_I2 != pl->k2_I2) {
    if (int rc = encode_level_maps(pl, 0, 1)) return rc;
    ICA_CUDA_CHECK(cudaMemcpyAsync(pl->tmaps_dev, pl->tmaps_host.data(), pl->tmaps_host.size(), cudaMemcpyHostToDevice, stream));
    pl->tm_I1 = pl->k2_I1; pl->tm_I2 = pl->k2_I2;
  }
  return ICA_OK;
}

// K2 of this plan: the tile kernel (default) or the column-march kernel (ICA_K2=march)
cudaError_t launch_k2(const ica_plan* pl, const IterParams& P, cudaStream_t stream) {
  return pl->march ? launch_march(P, pl->C, pl->dh, pl->grid, stream) : launch_iterate(P, pl->C, pl->dh, pl->grid, stream);
}

// kernel parameters of the current run (prepare_level0 first)
void fill_iter_params(const ica_plan* pl, const float* /*I1*/, const float* /*I2*/, IterParams* P) {
  memset(P, 0, sizeof(*P));
  P->I1_0 = pl->k2_I1; P->I2_0 = pl->k2_I2; P->in_stride = pl->k2_stride;
  P->pyr1 = pl->pyr1; P->pyr2 = pl->pyr2; P->pyr_stride = pl->pyr_stride;
  for (int s = 0; s < pl->nscales; ++s) P->lv[s] = pl->lv[s];
  P->lv[0].pitch = pl->k2_pitch;
  P->tmaps = pl->tmaps_dev;
  P->nscales = pl->nscales;
  P->state = pl->state; P->mm = pl->mm; P->partials = pl->partials;
  P->traj = (pl->cfg.flags & ICA_FLAG_RECORD_TRAJECTORY) ? pl->traj : nullptr;
  P->dbg_Hb = nullptr;
  P->dbg_time = pl->dbg_time;
  P->dbg_row = pl->grid;
  P->n_active = pl->n_active;
  P->traj_cap = pl->traj_cap;
  P->chunk_start = pl->chunk_start;
  P->item_pair = pl->item_pair;
  P->hdr = pl->hdr;
  P->pair_ticket = pl->pair_ticket;
  P->grid_ctas = pl->grid;
  P->fused = 0;                      // the callers that run the fused loop set it
  P->solve_ticket = pl->solve_ticket;
  P->sched_dirty = pl->loop_count + 2;
  P->asm_tab = pl->asm_tab;
  P->cond_handle = 0;
  P->loop_count = pl->loop_count;
  P->max_launches = pl->nscales * pl->cfg.max_iter;
  P->kernel_ns = pl->kernel_ns;
  P->shard_rank = pl->shard_rank; P->shard_n = pl->shard_n;
  P->solve_mode = 0; P->ext_moments = nullptr;
  P->x_peers = pl->x_peers_dev; P->x_world = pl->x_world; P->x_rank = pl->x_rank; P->x_seq_base = pl->x_seq_dev;
  P->x_error = pl->x_error; P->x_ns = pl->x_ns;
  P->B = pl->B;
  P->max_chunks = pl->max_chunks;
  P->chunk_unit = pl->march ? march_chunk_unit() : 0;
  P->chunk_m = pl->chunk_m;
  P->robust_type = pl->cfg.robust_type;
  P->robust_loop = pl->cfg.robust_loop;
  P->lambda_cfg = pl->cfg.lambda_;
  P->tol = pl->cfg.tol;
  P->max_iter = pl->cfg.max_iter;
  P->delta = pl->cfg.delta;
  P->frame = (pl->cfg.nanifoutside != 0 && pl->cfg.delta > 0) ? 1 : 0;
  P->ch_mult = (pl->C == 1 && pl->cfg.gray_as_rgb) ? 3.0f : 1.0f;
  P->ipol_warp = (pl->cfg.flags & ICA_FLAG_IPOL_WARP) ? 1 : 0;
  P->ipol_nan = pl->cfg.nanifoutside != 0 ? 1 : 0;
}

// CUDA graph of the iteration loop: one conditional WHILE node whose body is {iterate, solve}; the solve
// kernel's scheduling block sets the condition on the device, so the whole coarse-to-fine loop of every
// pair runs without a host round trip.  Rebuilt when the image pointers or the moment degree change.
int ensure_loop_graph(ica_plan* pl, const float* I1, const float* I2, int solve_mode = 0) {
  I1 = pl->k2_I1; I2 = pl->k2_I2;   // the graph bakes K2's view of level 0
  if (pl->graph_exec && pl->graph_I1 == I1 && pl->graph_I2 == I2 && pl->graph_dh == pl->dh && pl->graph_mode == solve_mode) return ICA_OK;
  if (pl->graph_exec) { cudaGraphExecDestroy(pl->graph_exec); pl->graph_exec = nullptr; }
  if (pl->graph) { cudaGraphDestroy(pl->graph); pl->graph = nullptr; }
  ICA_CUDA_CHECK(cudaGraphCreate(&pl->graph, 0));
  cudaGraphConditionalHandle handle;
  ICA_CUDA_CHECK(cudaGraphConditionalHandleCreate(&handle, pl->graph, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams cp = {};
  cp.type = cudaGraphNodeTypeConditional;
  cp.conditional.handle = handle;
  cp.conditional.type = cudaGraphCondTypeWhile;
  cp.conditional.size = 1;
  cudaGraphNode_t node;
  ICA_CUDA_CHECK(cudaGraphAddNode(&node, pl->graph, nullptr, 0, &cp));
  cudaGraph_t body = cp.conditional.phGraph_out[0];
  IterParams P;
  fill_iter_params(pl, I1, I2, &P);
  P.cond_handle = (unsigned long long)handle;
  P.fused = solve_mode == 0 ? pl->fused : 0;
  P.solve_mode = solve_mode;
  ICA_CUDA_CHECK(cudaStreamBeginCaptureToGraph(pl->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  cudaError_t e1 = launch_k2(pl, P, pl->stream);
  cudaError_t e2 = P.fused ? cudaSuccess : launch_solve(P, pl->dh, pl->stream);   // fused: the iterate kernel solves
  cudaGraph_t captured = nullptr;
  cudaError_t e3 = cudaStreamEndCapture(pl->stream, &captured);
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
    set_error("building the loop graph failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
    return ICA_ERR_CUDA;
  }
  ICA_CUDA_CHECK(cudaGraphInstantiate(&pl->graph_exec, pl->graph, 0));
  pl->graph_I1 = I1; pl->graph_I2 = I2; pl->graph_dh = pl->dh; pl->graph_mode = solve_mode;
  return ICA_OK;
}

// minmax of level 0 and all pyramid levels of both images
int build_pyramids(ica_plan* pl, const float* I1, const float* I2, cudaStream_t stream) {
  const int B = pl->B, ns = pl->nscales;
  // Level 0's min / max (the clip range of level 1 and of the warps at level 0): done by the host entry's
  // conversion kernels, else gathered by the first pyramid level's own read of level 0 when that level takes the fused
  // kernel (no separate pass over both images: -25 MB per pair), else by a separate pass.
  bool gather0 = false;
  if (!pl->mm_ready) {
    ICA_LAUNCH_CHECK(launch_minmax_reset(pl->mm, B * ns * 2, stream));
    pl->launches += 1;
    gather0 = ns > 1 && pl->gather_keys && !getenv("ICA_NO_GATHER") &&
              pyr_level_gathers(I1, I2, pl->in_stride, pl->lv[0].pitch, pl->lv[0].nx, pl->lv[0].ny, pl->C, pl->ry[0], pl->rx[0], pl->tmp,
                                (long long)pl->lv[1].ny * pl->lv[0].nx * pl->C);
    if (!gather0) {
      const long long n0 = (long long)pl->H * pl->W * pl->C;
      const float* src[2] = {I1, I2};
      for (int which = 0; which < 2; ++which) {
        ICA_LAUNCH_CHECK(launch_minmax(src[which], pl->in_stride, n0, B, pl->mm + which, ns * 2, stream));
        pl->launches += 1;
      }
    }
  }
  pl->mm_ready = false;
  for (int s = 0; s + 1 < ns; ++s) {
    const LevelDesc& Li = pl->lv[s];
    const LevelDesc& Lo = pl->lv[s + 1];
    // pairs per launch group: both images of a pair share the launch; the intermediate (vertical-pass) buffer of a
    // group stays within the budget the plan allocated (L2-resident between the two passes for the big levels)
    const long long tmp_per_img = (long long)Lo.ny * Li.nx * pl->C;
    const int group = (int)std::max<long long>(1, std::min<long long>(B, pl->tmp_floats / std::max<long long>(1, 2 * tmp_per_img)));
    for (int b0 = 0; b0 < B; b0 += group) {
      const int nset = std::min(group, B - b0);
      const float* ina = s == 0 ? I1 + (long long)b0 * pl->in_stride : pl->pyr1 + (long long)b0 * pl->pyr_stride + Li.offset;
      const float* inb = s == 0 ? I2 + (long long)b0 * pl->in_stride : pl->pyr2 + (long long)b0 * pl->pyr_stride + Li.offset;
      const long long istr = s == 0 ? pl->in_stride : pl->pyr_stride;
      float* outa = pl->pyr1 + (long long)b0 * pl->pyr_stride + Lo.offset;
      float* outb = pl->pyr2 + (long long)b0 * pl->pyr_stride + Lo.offset;
      int nl = 0;
      if (pl->timing && pl->n_ev_pyr + 2 <= (int)pl->ev_pyr.size()) cudaEventRecord(pl->ev_pyr[pl->n_ev_pyr++], stream);
      const bool gather = gather0 && s == 0;
      ICA_LAUNCH_CHECK(launch_pyr_down(ina, inb, istr, Li.pitch, Li.nx, Li.ny, pl->C, pl->ry[s], pl->rx[s], pl->tmp,
                                       tmp_per_img, outa, outb, pl->pyr_stride, Lo.pitch, nset,
                                       (pl->mm_open ? pl->mm_open : pl->mm) + ((long long)b0 * ns + s) * 2, pl->mm + ((long long)b0 * ns + s + 1) * 2,
                                       ns * 2, stream, &nl, nullptr, gather ? pl->mm + ((long long)b0 * ns + s) * 2 : nullptr,
                                       gather ? pl->gather_keys + ((long long)b0 * ns + s) * 2 : nullptr));
      if (gather && !pl->mm_open) {   // deferred clip of level 1 to level 0's range (zoom_out levels are never clipped)
        ICA_LAUNCH_CHECK(launch_clip_fixup(outa, outb, pl->pyr_stride, Lo.pitch, Lo.nx, Lo.ny, pl->C, nset,
                                           pl->mm + ((long long)b0 * ns + s) * 2, pl->mm + ((long long)b0 * ns + s + 1) * 2, ns * 2,
                                           pl->clip_flags + 2 * b0, stream));
        nl += 3;
      }
      if (pl->timing && pl->n_ev_pyr + 1 <= (int)pl->ev_pyr.size()) cudaEventRecord(pl->ev_pyr[pl->n_ev_pyr++], stream);
      pl->launches += nl;
    }
  }
  return ICA_OK;
}

}  // namespace

extern "C" {

const char* ica_last_error(void) { return g_err; }
int ica_version(void) { return 100; }

int ica_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int ica_set_device(int device) {
  if (int rc = require_device()) return rc;
  ICA_CUDA_CHECK(cudaSetDevice(device));
  return ICA_OK;
}

int ica_get_device(int* device_out) {
  if (!device_out) return ICA_ERR_INVALID;
  if (int rc = require_device()) return rc;
  ICA_CUDA_CHECK(cudaGetDevice(device_out));
  return ICA_OK;
}

int ica_get_constants(double* out5) {
  if (!out5) return ICA_ERR_INVALID;
  out5[0] = kMaxIter; out5[1] = kLambda0; out5[2] = kLambdaN; out5[3] = kLambdaRatio; out5[4] = kSplinePad;
  return ICA_OK;
}

int ica_plan_destroy(ica_plan* pl) {
  if (!pl) return ICA_OK;
  cudaFree(pl->pyr1); cudaFree(pl->pyr2); cudaFree(pl->tmp); cudaFree(pl->tmaps_dev); cudaFree(pl->pad1); cudaFree(pl->pad2);
  for (int s = 0; s < ICA_MAX_SCALES; ++s) { free_resample(&pl->ry[s]); free_resample(&pl->rx[s]); }
  for (int r = 0; r < pl->x_world; ++r) if (r != pl->x_rank && pl->x_peer_ptr[r]) cudaIpcCloseMemHandle(pl->x_peer_ptr[r]);
  cudaFree(pl->xbuf); cudaFree(pl->x_peers_dev); cudaFree(pl->x_error); cudaFree(pl->x_ns); cudaFree(pl->x_seq_dev);
  if (pl->x_seq_host) cudaFreeHost(pl->x_seq_host);
  cudaFree(pl->mm_open); cudaFree(pl->gather_keys); cudaFree(pl->clip_flags);
  cudaFree(pl->state); cudaFree(pl->mm); cudaFree(pl->partials); cudaFree(pl->chunk_start); cudaFree(pl->item_pair); cudaFree(pl->solve_ticket); cudaFree(pl->asm_tab); cudaFree(pl->loop_count); cudaFree(pl->kernel_ns); cudaFree(pl->hdr); cudaFree(pl->pair_ticket);
  if (pl->h_loop) cudaFreeHost(pl->h_loop);
  if (pl->h_kns) cudaFreeHost(pl->h_kns);
  if (pl->graph_exec) cudaGraphExecDestroy(pl->graph_exec);
  if (pl->graph) cudaGraphDestroy(pl->graph); cudaFree(pl->dbg_time); cudaFree(pl->traj); cudaFree(pl->n_active);
  if (pl->h_n_active) cudaFreeHost(pl->h_n_active);
  cudaFree(pl->ttypes_dev); cudaFree(pl->p_dev); cudaFree(pl->err_dev); cudaFree(pl->iters_dev);
  cudaFree(pl->in1_dev); cudaFree(pl->in2_dev); cudaFree(pl->raw_dev); cudaFree(pl->DI_dev); cudaFree(pl->Iw_dev); cudaFree(pl->out64_dev);
  for (auto e : pl->ev_iter) cudaEventDestroy(e);
  for (auto e : pl->ev_pyr) cudaEventDestroy(e);
  if (pl->ev_host0) cudaEventDestroy(pl->ev_host0);
  if (pl->ev_host1) cudaEventDestroy(pl->ev_host1);
  if (pl->ev_h2d_done) cudaEventDestroy(pl->ev_h2d_done);
  if (pl->stream) cudaStreamDestroy(pl->stream);
  delete pl;
  return ICA_OK;
}

int ica_plan_create(const ica_config* cfg, ica_plan** plan_out) {
  if (!plan_out) { set_error("plan_out is NULL"); return ICA_ERR_INVALID; }
  *plan_out = nullptr;
  if (int rc = validate_config(cfg)) return rc;
  if (int rc = require_device()) return rc;
  ica_plan* pl = new ica_plan();
  pl->cfg = *cfg;
  if (pl->cfg.robust_type != QUADRATIC) pl->cfg.robust_loop = 1;
  pl->B = cfg->batch; pl->H = cfg->height; pl->W = cfg->width; pl->C = cfg->channels; pl->nscales = cfg->nscales;
  pl->ttypes.assign(pl->B, cfg->transform_type);
  pl->dh = moment_degree_of(cfg->transform_type);
  {   // which K2: ICA_K2=march selects the column-march kernel (ica_march.cu).  Measured on B200 (round 2,
      // profiles/README.md) it needs a third of the shared-memory loads but 1.5x the instructions of the tile kernel and
      // runs at 0.12 instead of 0.24 of the HBM roofline: it stays opt-in.
    const char* e = getenv("ICA_K2");
    pl->march = (e && (e[0] == 'm' || e[0] == 'M')) ? 1 : 0;
    const char* m = getenv("ICA_CHUNK_M");
    if (m && atoi(m) > 0) pl->chunk_m = atoi(m);
  }
  const int TWv = pl->march ? march_tile_w() : iterate_tile_w(), THv = pl->march ? march_tile_h() : iterate_tile_h();
  // level shapes (zoom.zoom_size, src/zoom.py:8-22; skimage uses the same rounding)
  long long off = 0;
  int nx = pl->W, ny = pl->H;
  for (int s = 0; s < pl->nscales; ++s) {
    if (s > 0) { nx = std::max(zoomed_size(nx, cfg->nu), 1); ny = std::max(zoomed_size(ny, cfg->nu), 1); }
    LevelDesc& L = pl->lv[s];
    L.nx = nx; L.ny = ny;
    // levels >= 1 live in plan-owned slabs: pad rows to 16 bytes so the TMA bulk copies apply
    L.pitch = s == 0 ? nx * pl->C : ((nx * pl->C + 3) / 4) * 4;
    L.offset = s == 0 ? 0 : off;
    L.tiles_x = (nx + TWv - 1) / TWv; L.tiles_y = (ny + THv - 1) / THv;
    if (s > 0) off += ((long long)L.pitch * ny + 31) / 32 * 32;
  }
  pl->in_stride = (long long)pl->H * pl->W * pl->C;
  pl->pyr_stride = off;
  const int nt0 = pl->lv[0].tiles_x * pl->lv[0].tiles_y;
  // chunks (= partial slots) per pair: enough for one straggler pair to spread over the chip, few enough that the chunk
  // epilogues and the per-pair sum of the partials stay cheap (measured at 1024^2: 128 chunks 9830 pairs/s, 256 chunks
  // 9370, single-pair latency unchanged); independent of the batch size, so a pair's result does not depend on what
  // it is batched with
  int mc = cfg->blocks_per_pair > 0 ? cfg->blocks_per_pair : (nt0 > 16384 ? 1024 : 128);
  if (cfg->blocks_per_pair <= 0) { const char* e = getenv("ICA_CHUNKS"); if (e && atoi(e) > 0) mc = atoi(e); }   // tuning hook
  pl->max_chunks = std::max(1, std::min(std::min(mc, nt0), 2047));   // (work items pack the chunk index into 11 bits)
  {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    pl->device = dev;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    pl->grid = (pl->march ? 1 : iterate_blocks_per_sm()) * sms;   // persistent kernel: every CTA resident
  }
  int rc = ICA_OK;
#define TRY(expr) do { if ((rc = (expr)) != ICA_OK) { ica_plan_destroy(pl); return rc; } } while (0)
#define TRY_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { set_error("%s failed: %s", #expr, cudaGetErrorString(e__)); ica_plan_destroy(pl); return ICA_ERR_CUDA; } } while (0)
  if (pl->nscales > 1) {
    TRY(dev_alloc(pl, &pl->pyr1, (size_t)pl->B * pl->pyr_stride));
    TRY(dev_alloc(pl, &pl->pyr2, (size_t)pl->B * pl->pyr_stride));
    pl->tmp_stride = (long long)pl->lv[1].ny * pl->lv[0].nx * pl->C;
    // intermediate (vertical-pass) image: only the border rows and columns of a level are ever written when the fused
    // kernel applies, so it is sized for the whole batch (one launch group per level) up to 4 GiB
    const long long budget = 4096ll << 20;
    pl->tmp_images = (int)std::max<long long>(1, std::min<long long>(2ll * pl->B, budget / std::max<long long>(1, pl->tmp_stride * 4)));
    pl->tmp_floats = std::max<long long>((long long)pl->tmp_images * pl->tmp_stride, 2 * pl->tmp_stride);
    pl->tmp_floats = (pl->tmp_floats + 3) / 4 * 4;
    TRY(dev_alloc(pl, &pl->tmp, (size_t)pl->tmp_floats));
    for (int s = 0; s + 1 < pl->nscales; ++s) {
      Resample1D r;
      const bool ipol_pyr = (cfg->flags & ICA_FLAG_IPOL_PYRAMID) != 0;   // zoom.zoom_out levels instead of skimage rescale
      if (ipol_pyr) build_zoom_out_1d(pl->lv[s].ny, cfg->nu, 0.6, &r); else build_resample_1d(pl->lv[s].ny, pl->lv[s + 1].ny, &r);
      if (r.taps > max_taps()) { set_error("resampling operator too wide (%d taps)", r.taps); ica_plan_destroy(pl); return ICA_ERR_INVALID; }
      TRY(upload_resample(pl, r, &pl->ry[s]));
      if (ipol_pyr) build_zoom_out_1d(pl->lv[s].nx, cfg->nu, 0.6, &r); else build_resample_1d(pl->lv[s].nx, pl->lv[s + 1].nx, &r);
      if (r.taps > max_taps()) { set_error("resampling operator too wide (%d taps)", r.taps); ica_plan_destroy(pl); return ICA_ERR_INVALID; }
      TRY(upload_resample(pl, r, &pl->rx[s]));
    }
  }
  TRY(dev_alloc(pl, &pl->state, (size_t)pl->B));
  TRY(dev_alloc(pl, &pl->mm, (size_t)pl->B * pl->nscales * 2));
  {
    TRY(dev_alloc(pl, &pl->gather_keys, (size_t)pl->B * pl->nscales * 2));
    TRY(dev_alloc(pl, &pl->clip_flags, (size_t)pl->B * 2));
    std::vector<MinMaxKeys> open_keys((size_t)pl->B * pl->nscales * 2);
    for (auto& k : open_keys) { k.lo = float_key(-3.4028235e38f); k.hi = float_key(3.4028235e38f); }
    TRY_CUDA(cudaMemcpy(pl->gather_keys, open_keys.data(), open_keys.size() * sizeof(MinMaxKeys), cudaMemcpyHostToDevice));
  }
  if (cfg->flags & ICA_FLAG_IPOL_PYRAMID) {   // zoom_out does not clip: the pyramid kernels get an open parent range
    TRY(dev_alloc(pl, &pl->mm_open, (size_t)pl->B * pl->nscales * 2));
    std::vector<MinMaxKeys> open_keys((size_t)pl->B * pl->nscales * 2);
    for (auto& k : open_keys) { k.lo = float_key(-3.4028235e38f); k.hi = float_key(3.4028235e38f); }
    TRY_CUDA(cudaMemcpy(pl->mm_open, open_keys.data(), open_keys.size() * sizeof(MinMaxKeys), cudaMemcpyHostToDevice));
  }
  TRY(dev_alloc(pl, &pl->partials, (size_t)pl->B * pl->max_chunks * kAccStride));
  TRY(dev_alloc(pl, &pl->chunk_start, 2 * ((size_t)pl->B + 1)));                 // double-buffered work lists
  TRY(dev_alloc(pl, &pl->item_pair, 2 * (size_t)pl->B * pl->max_chunks));
  TRY(dev_alloc(pl, &pl->hdr, 2));
  TRY_CUDA(cudaMemset(pl->hdr, 0, 2 * sizeof(SchedHdr)));
  TRY(dev_alloc(pl, &pl->pair_ticket, (size_t)pl->B));
  TRY_CUDA(cudaMemset(pl->pair_ticket, 0, (size_t)pl->B * sizeof(unsigned int)));
  TRY(dev_alloc(pl, &pl->solve_ticket, 1));
  TRY(dev_alloc(pl, &pl->loop_count, 4));      // iterations, iterate launches (safety net), work-list dirty flag
  TRY_CUDA(cudaMemset(pl->loop_count, 0, 4 * sizeof(int)));
  TRY(dev_alloc(pl, &pl->kernel_ns, 2));
  TRY_CUDA(cudaMemset(pl->kernel_ns, 0, 2 * sizeof(long long)));
  TRY_CUDA(cudaMallocHost((void**)&pl->h_loop, 2 * sizeof(int)));
  TRY_CUDA(cudaMallocHost((void**)&pl->h_kns, 2 * sizeof(long long)));
  pl->h_loop[0] = 0; pl->h_kns[0] = 0; pl->h_kns[1] = 0;
  pl->use_graph = (cfg->flags & ICA_FLAG_HOST_LOOP) ? 0 : (getenv("ICA_NO_GRAPH") ? 0 : 1);
  // Fused solve (the CTA that finishes a pair's last chunk solves it inside the iterate kernel: one launch per
  // iteration) is available but off by default: measured on B200 (round 2, profiles/README.md) it is SLOWER than the
  // separate 512-thread solve launch -- 28.5 vs 26.9 ms per 256-pair step, 2.39 vs 2.12 ms for a single pair -- the
  // solve is a latency chain that runs worse at the iterate kernel's 80 registers, and it stalls a streaming CTA.
  pl->fused = (getenv("ICA_FUSE") && !pl->march) ? 1 : 0;
  TRY(dev_alloc(pl, &pl->asm_tab, (size_t)6 * 72));
  TRY(upload_assembly(pl));
  TRY_CUDA(cudaMemset(pl->solve_ticket, 0, sizeof(unsigned int)));
  pl->traj_cap = pl->nscales * cfg->max_iter;
  if (cfg->flags & ICA_FLAG_RECORD_TRAJECTORY) TRY(dev_alloc(pl, &pl->traj, (size_t)pl->B * pl->traj_cap * ICA_TRAJ_STRIDE));
  TRY(dev_alloc(pl, &pl->n_active, 1));
  TRY_CUDA(cudaMallocHost((void**)&pl->h_n_active, sizeof(int)));
  TRY(dev_alloc(pl, &pl->ttypes_dev, (size_t)pl->B));
  TRY_CUDA(cudaMemcpy(pl->ttypes_dev, pl->ttypes.data(), pl->B * sizeof(int), cudaMemcpyHostToDevice));
  TRY(dev_alloc(pl, &pl->p_dev, (size_t)pl->B * ICA_MAX_PARAMS));
  TRY(dev_alloc(pl, &pl->err_dev, (size_t)pl->B));
  TRY(dev_alloc(pl, &pl->iters_dev, (size_t)pl->B * pl->nscales));
  if (cfg->flags & ICA_FLAG_WRITE_DI_IW) {
    TRY(dev_alloc(pl, &pl->DI_dev, (size_t)pl->B * pl->in_stride));
    TRY(dev_alloc(pl, &pl->Iw_dev, (size_t)pl->B * pl->in_stride));
  }
  pl->tmaps_host.assign((size_t)pl->B * pl->nscales * 2 * 128, 0);
  TRY_CUDA(cudaMalloc(&pl->tmaps_dev, pl->tmaps_host.size()));
  pl->device_bytes += pl->tmaps_host.size();
  if (pl->nscales > 1) TRY(encode_level_maps(pl, 1, pl->nscales));   // levels >= 1 live in the plan's slabs
  TRY_CUDA(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
  TRY_CUDA(cudaEventCreate(&pl->ev_host0));
  TRY_CUDA(cudaEventCreate(&pl->ev_host1));
  TRY_CUDA(cudaEventCreateWithFlags(&pl->ev_h2d_done, cudaEventDisableTiming));
  TRY_CUDA(cudaMemset(pl->state, 0, pl->B * sizeof(PairState)));
#undef TRY
#undef TRY_CUDA
  *plan_out = pl;
  return ICA_OK;
}

int ica_plan_set_transform_types(ica_plan* pl, const int32_t* types, int32_t count) {
  if (!pl || !types || count != pl->B) { set_error("transform type array must have one entry per pair"); return ICA_ERR_INVALID; }
  int dh = 0;
  for (int i = 0; i < count; ++i) {
    if (nparams_of(types[i]) < 0) { set_error("Unknown transform type"); return ICA_ERR_INVALID; }
    dh = std::max(dh, moment_degree_of(types[i]));
  }
  pl->ttypes.assign(types, types + count);
  pl->dh = dh;
  if (int rc = upload_assembly(pl)) return rc;
  ICA_CUDA_CHECK(cudaMemcpy(pl->ttypes_dev, pl->ttypes.data(), count * sizeof(int), cudaMemcpyHostToDevice));
  return ICA_OK;
}

int ica_plan_level_shapes(const ica_plan* pl, int32_t* nx_out, int32_t* ny_out) {
  if (!pl) return ICA_ERR_INVALID;
  for (int s = 0; s < pl->nscales; ++s) { if (nx_out) nx_out[s] = pl->lv[s].nx; if (ny_out) ny_out[s] = pl->lv[s].ny; }
  return ICA_OK;
}

size_t ica_plan_device_bytes(const ica_plan* pl) { return pl ? pl->device_bytes : 0; }
int64_t ica_plan_last_launch_count(const ica_plan* pl) {
  // valid once the run's stream work has completed (the iteration count is copied back asynchronously)
  return pl ? pl->launches + (long long)pl->last_launches_per_iter * pl->h_loop[0] : 0;
}

int ica_plan_enable_timing(ica_plan* pl, int32_t enable) {
  if (!pl) return ICA_ERR_INVALID;
  pl->timing = enable < 0 ? 0 : enable;   // 1: events around the pyramid launches; 2: also host loop + events per iterate launch
  if (enable && pl->ev_iter.empty()) {
    const int n_it = 2 * (pl->nscales * pl->cfg.max_iter + 8);
    const int n_py = 4 * pl->nscales * ((pl->B + std::max(1, pl->tmp_images) - 1) / std::max(1, pl->tmp_images)) + 8;
    pl->ev_iter.resize(n_it); pl->ev_pyr.resize(n_py);
    for (auto& e : pl->ev_iter) ICA_CUDA_CHECK(cudaEventCreate(&e));
    for (auto& e : pl->ev_pyr) ICA_CUDA_CHECK(cudaEventCreate(&e));
  }
  return ICA_OK;
}

int ica_plan_get_timing(ica_plan* pl, float* iterate_ms, int32_t* iterate_launches, float* pyramid_ms,
                        int32_t* pyramid_launches) {
  if (!pl) return ICA_ERR_INVALID;
  float it = 0.f, py = 0.f;
  for (int i = 0; i + 1 < pl->n_ev_iter; i += 2) {
    float ms = 0.f;
    ICA_CUDA_CHECK(cudaEventSynchronize(pl->ev_iter[i + 1]));
    ICA_CUDA_CHECK(cudaEventElapsedTime(&ms, pl->ev_iter[i], pl->ev_iter[i + 1]));
    it += ms;
  }
  for (int i = 0; i + 1 < pl->n_ev_pyr; i += 2) {
    float ms = 0.f;
    ICA_CUDA_CHECK(cudaEventSynchronize(pl->ev_pyr[i + 1]));
    ICA_CUDA_CHECK(cudaEventElapsedTime(&ms, pl->ev_pyr[i], pl->ev_pyr[i + 1]));
    py += ms;
  }
  if (pl->n_ev_iter == 0) {   // graph loop: device-side %globaltimer span of every iterate launch
    it = (float)(pl->h_kns[0] * 1e-6);
    if (iterate_launches) *iterate_launches = (int)pl->h_kns[1];
  } else if (iterate_launches) *iterate_launches = pl->n_ev_iter / 2;
  if (iterate_ms) *iterate_ms = it;
  if (pyramid_ms) *pyramid_ms = py;
  if (pyramid_launches) *pyramid_launches = pl->n_ev_pyr;  // two kernels per recorded pair
  return ICA_OK;
}

// ---- row-sharded mode: one (large) pair per plan entry, split by bands of tile rows across ranks.  Every rank
// holds both full images and runs the whole pipeline except that K2 only visits its band; the caller sums the
// per-pair moment vectors over ranks (NCCL allreduce) between ica_plan_shard_partial and ica_plan_shard_solve,
// after which every rank performs the identical solve and keeps an identical state.
int ica_moment_stride(void) { return kAccStride; }

int ica_row_band(int32_t tiles_y, int32_t rank, int32_t nranks, int32_t* ty0, int32_t* ty1) {
  if (tiles_y < 0 || nranks < 1 || rank < 0 || rank >= nranks) { set_error("invalid band request"); return ICA_ERR_INVALID; }
  if (ty0) *ty0 = (int)((long long)rank * tiles_y / nranks);
  if (ty1) *ty1 = (int)((long long)(rank + 1) * tiles_y / nranks);
  return ICA_OK;
}

int ica_plan_set_row_shard(ica_plan* pl, int32_t rank, int32_t nranks) {
  if (!pl || nranks < 1 || rank < 0 || rank >= nranks) { set_error("invalid row shard (rank %d of %d)", rank, nranks); return ICA_ERR_INVALID; }
  pl->shard_rank = rank; pl->shard_n = nranks;
  if (pl->graph_exec) { cudaGraphExecDestroy(pl->graph_exec); pl->graph_exec = nullptr; }   // kernel arguments are baked in
  return ICA_OK;
}

int ica_plan_shard_begin(ica_plan* pl, const float* I1, const float* I2, const double* p_in, void* stream_) {
  if (!pl || !I1 || !I2 || !p_in) { set_error("NULL argument"); return ICA_ERR_INVALID; }
  if (int rc = enter_plan_device(pl)) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  pl->launches = 0; pl->n_ev_iter = 0; pl->n_ev_pyr = 0;
  pl->last_I1 = I1; pl->last_I2 = I2;
  if (int rc = prepare_level0(pl, I1, I2, stream)) return rc;
  if (int rc = build_pyramids(pl, I1, I2, stream)) return rc;
  ICA_LAUNCH_CHECK(launch_init_state(pl->state, p_in, pl->ttypes_dev, pl->B, pl->nscales, pl->cfg.lambda_, pl->n_active, pl->pair_ticket, stream));
  IterParams P;
  fill_iter_params(pl, I1, I2, &P);
  ICA_LAUNCH_CHECK(launch_schedule(P, stream));
  pl->launches += 2;
  return ICA_OK;
}

int ica_plan_shard_partial(ica_plan* pl, double* moments, void* stream_) {
  if (!pl || !moments || !pl->last_I1) { set_error("ica_plan_shard_begin must be called first"); return ICA_ERR_INVALID; }
  cudaStream_t stream = (cudaStream_t)stream_;
  IterParams P;
  fill_iter_params(pl, pl->last_I1, pl->last_I2, &P);
  P.solve_mode = 1; P.ext_moments = moments;
  ICA_LAUNCH_CHECK(launch_k2(pl, P, stream));
  ICA_LAUNCH_CHECK(launch_solve(P, pl->dh, stream));
  pl->launches += 2;
  return ICA_OK;
}

int ica_plan_shard_solve(ica_plan* pl, const double* moments, int32_t* n_active_out, void* stream_) {
  if (!pl || !moments || !pl->last_I1) { set_error("ica_plan_shard_begin must be called first"); return ICA_ERR_INVALID; }
  cudaStream_t stream = (cudaStream_t)stream_;
  IterParams P;
  fill_iter_params(pl, pl->last_I1, pl->last_I2, &P);
  P.solve_mode = 2; P.ext_moments = const_cast<double*>(moments);
  ICA_LAUNCH_CHECK(launch_solve(P, pl->dh, stream));
  pl->launches += 1;
  if (n_active_out) {
    ICA_CUDA_CHECK(cudaMemcpyAsync(pl->h_n_active, pl->n_active, sizeof(int), cudaMemcpyDeviceToHost, stream));
    ICA_CUDA_CHECK(cudaStreamSynchronize(stream));
    *n_active_out = *pl->h_n_active;
  }
  return ICA_OK;
}

int ica_plan_shard_finish(ica_plan* pl, double* p_out, void* stream_) {
  if (!pl || !p_out || !pl->last_I1) { set_error("ica_plan_shard_begin must be called first"); return ICA_ERR_INVALID; }
  cudaStream_t stream = (cudaStream_t)stream_;
  ICA_CUDA_CHECK(cudaMemcpyAsync(pl->h_loop, pl->loop_count, sizeof(int), cudaMemcpyDeviceToHost, stream));
  ICA_LAUNCH_CHECK(launch_export_results(pl->state, pl->B, p_out, pl->err_dev, pl->iters_dev, pl->nscales, stream));
  pl->launches += 1;
  if (pl->cfg.flags & ICA_FLAG_WRITE_DI_IW) {
    ICA_LAUNCH_CHECK(launch_warp_out(pl->last_I1, pl->last_I2, pl->in_stride, pl->W, pl->H, pl->C, pl->state, pl->mm, pl->nscales,
                                     pl->B, pl->Iw_dev, pl->DI_dev, (pl->cfg.flags & ICA_FLAG_IPOL_WARP) ? 1 : 0, pl->cfg.nanifoutside != 0, pl->cfg.delta, stream));
    pl->launches += 1;
  }
  return ICA_OK;
}

// ---- row-sharded mode with the per-iteration exchange INSIDE the device-side loop (BASELINE configs[4]): every rank
// owns an exchange buffer; ica_plan_xchg_create returns its CUDA IPC handle, the caller gathers the handles of all ranks
// (torch.distributed / MPI / a file -- 64 bytes each) and passes them to ica_plan_xchg_connect, which maps the peers'
// buffers (NVLink peer access).  ica_plan_run_row_sharded then runs the whole registration as ONE graph launch per rank:
// K2 on the rank's band, K3 adds the ranks' moment sums through the mapped buffers (no NCCL call, no host round trip).
int ica_plan_xchg_create(ica_plan* pl, int32_t world, int32_t rank, void* ipc_handle_out64) {
  if (!pl || world < 1 || world > 64 || rank < 0 || rank >= world) { set_error("invalid exchange group (rank %d of %d)", rank, world); return ICA_ERR_INVALID; }
  if (int rc = enter_plan_device(pl)) return rc;
  if (pl->xbuf) { set_error("the exchange buffer of this plan exists already"); return ICA_ERR_INVALID; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles travel as 64-byte records");
  const size_t n = (size_t)2 * world * pl->B * kXSlot;
  if (int rc = dev_alloc(pl, &pl->xbuf, n)) return rc;
  ICA_CUDA_CHECK(cudaMemset(pl->xbuf, 0, n * sizeof(double)));
  if (int rc = dev_alloc(pl, &pl->x_peers_dev, (size_t)world)) return rc;
  if (int rc = dev_alloc(pl, &pl->x_error, 1)) return rc;
  if (int rc = dev_alloc(pl, &pl->x_ns, 2)) return rc;
  if (int rc = dev_alloc(pl, &pl->x_seq_dev, 1)) return rc;
  ICA_CUDA_CHECK(cudaMallocHost((void**)&pl->x_seq_host, sizeof(unsigned long long)));
  ICA_CUDA_CHECK(cudaMemset(pl->x_error, 0, sizeof(int)));
  ICA_CUDA_CHECK(cudaMemset(pl->x_ns, 0, 2 * sizeof(long long)));
  pl->x_world = world; pl->x_rank = rank;
  pl->x_peer_ptr[rank] = pl->xbuf;
  if (ipc_handle_out64) {
    cudaIpcMemHandle_t h;
    ICA_CUDA_CHECK(cudaIpcGetMemHandle(&h, pl->xbuf));
    memcpy(ipc_handle_out64, &h, 64);
  }
  if (int rc = ica_plan_set_row_shard(pl, rank, world)) return rc;
  if (world == 1) {
    ICA_CUDA_CHECK(cudaMemcpy(pl->x_peers_dev, pl->x_peer_ptr, sizeof(void*), cudaMemcpyHostToDevice));
  }
  return ICA_OK;
}

int ica_plan_xchg_connect(ica_plan* pl, const void* ipc_handles /* [world][64] */) {
  if (!pl || !pl->xbuf || !ipc_handles) { set_error("ica_plan_xchg_create must be called first"); return ICA_ERR_INVALID; }
  if (int rc = enter_plan_device(pl)) return rc;
  for (int r = 0; r < pl->x_world; ++r) {
    if (r == pl->x_rank || pl->x_peer_ptr[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(ipc_handles) + (size_t)r * 64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e)); return ICA_ERR_CUDA; }
    pl->x_peer_ptr[r] = p;
  }
  ICA_CUDA_CHECK(cudaMemcpy(pl->x_peers_dev, pl->x_peer_ptr, (size_t)pl->x_world * sizeof(void*), cudaMemcpyHostToDevice));
  return ICA_OK;
}

// mean exchange time (publish -> all ranks' sums seen, i.e. wire latency + waiting for the slowest rank) in microseconds,
// number of exchanges since the last call, and whether a peer timed out
int ica_plan_xchg_stats(ica_plan* pl, double* mean_us_out, int64_t* count_out, int32_t* error_out) {
  if (!pl || !pl->xbuf) { set_error("no exchange buffer"); return ICA_ERR_INVALID; }
  if (int rc = enter_plan_device(pl)) return rc;
  long long ns[2] = {0, 0}; int err = 0;
  ICA_CUDA_CHECK(cudaDeviceSynchronize());
  ICA_CUDA_CHECK(cudaMemcpy(ns, pl->x_ns, sizeof(ns), cudaMemcpyDeviceToHost));
  ICA_CUDA_CHECK(cudaMemcpy(&err, pl->x_error, sizeof(int), cudaMemcpyDeviceToHost));
  ICA_CUDA_CHECK(cudaMemset(pl->x_ns, 0, sizeof(ns)));
  if (mean_us_out) *mean_us_out = ns[1] > 0 ? 1e-3 * (double)ns[0] / (double)ns[1] : 0.0;
  if (count_out) *count_out = ns[1];
  if (error_out) *error_out = err;
  return ICA_OK;
}

int ica_plan_run_row_sharded(ica_plan* pl, const float* I1, const float* I2, double* p_inout, void* stream_) {
  if (!pl || !I1 || !I2 || !p_inout) { set_error("NULL argument"); return ICA_ERR_INVALID; }
  if (!pl->xbuf || !pl->x_peers_dev) { set_error("ica_plan_xchg_create / ica_plan_xchg_connect must be called first"); return ICA_ERR_INVALID; }
  if (int rc = enter_plan_device(pl)) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  pl->launches = 0; pl->n_ev_iter = 0; pl->n_ev_pyr = 0;
  pl->last_I1 = I1; pl->last_I2 = I2;
  const int max_launches = pl->nscales * pl->cfg.max_iter;
  // sequence numbers of this run's exchanges: every rank consumes the same range, so they agree without talking
  ICA_CUDA_CHECK(cudaStreamSynchronize(stream));          // (a previous run on this stream may still read the staging word)
  *pl->x_seq_host = pl->x_seq;
  ICA_CUDA_CHECK(cudaMemcpyAsync(pl->x_seq_dev, pl->x_seq_host, sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
  pl->x_seq += (unsigned long long)max_launches + 2;
  if (int rc = prepare_level0(pl, I1, I2, stream)) return rc;
  if (int rc = build_pyramids(pl, I1, I2, stream)) return rc;
  ICA_LAUNCH_CHECK(launch_init_state(pl->state, p_inout, pl->ttypes_dev, pl->B, pl->nscales, pl->cfg.lambda_, pl->n_active,
                                     pl->pair_ticket, stream));
  IterParams P;
  fill_iter_params(pl, I1, I2, &P);
  ICA_LAUNCH_CHECK(launch_schedule(P, stream));
  pl->launches += 2;

  pl->last_launches_per_iter = 2;
  if (pl->use_graph) {
    if (int rc = ensure_loop_graph(pl, I1, I2, 3)) return rc;
    ICA_CUDA_CHECK(cudaGraphLaunch(pl->graph_exec, stream));
  } else {
    // host-driven variant (profilers): every rank launches the full count; finished iterations are empty
    P.solve_mode = 3;
    for (int it = 0; it < max_launches; ++it) {
      ICA_LAUNCH_CHECK(launch_k2(pl, P, stream));
      ICA_LAUNCH_CHECK(launch_solve(P, pl->dh, stream));
    }
  }
  ICA_CUDA_CHECK(cudaMemcpyAsync(pl->h_loop, pl->loop_count, sizeof(int), cudaMemcpyDeviceToHost, stream));
  ICA_CUDA_CHECK(cudaMemcpyAsync(pl->h_kns, pl->kernel_ns, 2 * sizeof(long long), cudaMemcpyDeviceToHost, stream));
  ICA_LAUNCH_CHECK(launch_export_results(pl->state, pl->B, p_inout, pl->err_dev, pl->iters_dev, pl->nscales, stream));
  pl->launches += 1;
  return ICA_OK;
}

int ica_plan_run_device(ica_plan* pl, const float* I1, const float* I2, double* p_inout, void* stream_) {
  if (!pl || !I1 || !I2 || !p_inout) { set_error("NULL argument"); return ICA_ERR_INVALID; }
  if (pl->shard_n != 1) { set_error("the plan is row-sharded: use ica_plan_shard_begin/partial/solve/finish"); return ICA_ERR_INVALID; }
  if (int rc = enter_plan_device(pl)) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  pl->launches = 0;
  pl->n_ev_iter = 0; pl->n_ev_pyr = 0;
  pl->last_I1 = I1; pl->last_I2 = I2;
  if (int rc = prepare_level0(pl, I1, I2, stream)) return rc;
  if (int rc = build_pyramids(pl, I1, I2, stream)) return rc;
  ICA_LAUNCH_CHECK(launch_init_state(pl->state, p_inout, pl->ttypes_dev, pl->B, pl->nscales, pl->cfg.lambda_,
                                     pl->n_active, pl->pair_ticket, stream));
  pl->launches += 1;
  IterParams P;
  fill_iter_params(pl, I1, I2, &P);
  const int max_launches = pl->nscales * pl->cfg.max_iter;
  // timing mode 2 brackets every iterate launch with CUDA events: keep the solve in its own launch there
  const bool fused = pl->fused && pl->timing < 2;
  P.fused = fused ? 1 : 0;            // (before the first schedule: the work counter's start value depends on it)
  pl->last_launches_per_iter = fused ? 1 : 2;
  ICA_LAUNCH_CHECK(launch_schedule(P, stream));   // work list of the first iteration
  pl->launches += 1;
  if (pl->use_graph && pl->timing < 2) {
    // device-side loop: CUDA-graph while node, condition set by the solve kernel
    if (int rc = ensure_loop_graph(pl, I1, I2)) return rc;
    ICA_CUDA_CHECK(cudaGraphLaunch(pl->graph_exec, stream));
  } else {
    const int poll_every = 8;
    int done = 0;
    // a pair needs at least one launch per scale, so the first poll can wait that long
    int next_poll = std::max(pl->nscales, poll_every);
    for (int it = 0; it < max_launches && !done; ++it) {
      const bool timed = pl->timing >= 2 && pl->n_ev_iter + 2 <= (int)pl->ev_iter.size();
      if (timed) cudaEventRecord(pl->ev_iter[pl->n_ev_iter++], stream);
      ICA_LAUNCH_CHECK(launch_k2(pl, P, stream));
      if (timed) cudaEventRecord(pl->ev_iter[pl->n_ev_iter++], stream);
      // per-pair solve / compose; its last block publishes the next work list and the number of unfinished pairs
      if (!fused) ICA_LAUNCH_CHECK(launch_solve(P, pl->dh, stream));
      if (it + 1 >= next_poll && it + 1 < max_launches) {
        ICA_CUDA_CHECK(cudaMemcpyAsync(pl->h_n_active, pl->n_active, sizeof(int), cudaMemcpyDeviceToHost, stream));
        ICA_CUDA_CHECK(cudaStreamSynchronize(stream));
        if (*pl->h_n_active <= 0) done = 1;
        next_poll = it + 1 + poll_every;
      }
    }
  }
  // iterations executed and time spent in the iterate kernel (device-side counters), read lazily by the getters
  ICA_CUDA_CHECK(cudaMemcpyAsync(pl->h_loop, pl->loop_count, sizeof(int), cudaMemcpyDeviceToHost, stream));
  ICA_CUDA_CHECK(cudaMemcpyAsync(pl->h_kns, pl->kernel_ns, 2 * sizeof(long long), cudaMemcpyDeviceToHost, stream));
  ICA_LAUNCH_CHECK(launch_export_results(pl->state, pl->B, p_inout, pl->err_dev, pl->iters_dev, pl->nscales, stream));
  pl->launches += 1;
  if (pl->cfg.flags & ICA_FLAG_WRITE_DI_IW) {
    ICA_LAUNCH_CHECK(launch_warp_out(I1, I2, pl->in_stride, pl->W, pl->H, pl->C, pl->state, pl->mm, pl->nscales, pl->B,
                                     pl->Iw_dev, pl->DI_dev, (pl->cfg.flags & ICA_FLAG_IPOL_WARP) ? 1 : 0, pl->cfg.nanifoutside != 0,
                                     pl->cfg.delta, stream));
    pl->launches += 1;
  }
  return ICA_OK;
}

// Host->device copies of concurrent ica_plan_run_host calls (different plans, different threads) take turns on the
// PCIe link: while one plan's kernels run, the next plan's inputs are copied, instead of all copies sharing the link
// and every plan starting late.  (Ordering the copies on the device with a chained event instead of this host-side
// hand-over was measured slower.)
static std::mutex g_h2d_mutex[64];   // one per device: copies to different GPUs do not wait for one another



// The host entry waits on events by polling with a yield in between: a lone caller sees the event as fast as a spinning
// cudaEventSynchronize would, while many concurrent callers (several plans per GPU, several ranks per host) give their
// core away instead of spinning on it (blocking-sync events were measured to add ~0.4 ms to a single call).
static int wait_event_polite(cudaEvent_t ev) {
  // the first ~30 us are polled with yields (a lone short call sees its event at once); after that the thread sleeps
  // 50 us between polls: with many plans in flight (8 per rank x 8 ranks on one host in the 8-GPU bench) the waiting
  // threads leave the cores to the ones that are enqueueing work instead of storming the scheduler with yields
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    const cudaError_t e = cudaEventQuery(ev);
    if (e == cudaSuccess) return ICA_OK;
    if (e != cudaErrorNotReady) { set_error("cudaEventQuery failed: %s", cudaGetErrorString(e)); return ICA_ERR_CUDA; }
    if (std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(30)) std::this_thread::yield();
    else std::this_thread::sleep_for(std::chrono::microseconds(50));
  }
}

// both images of the batch host -> device; u8 / f64 inputs go through two staging areas so that the two copies run
// back to back and the link is released before the conversions to float32 are even scheduled
static int upload_pair(ica_plan* pl, const void* h1, const void* h2, int dtype_in, cudaStream_t stream) {
  const bool luma = (dtype_in & ICA_DTYPE_RGB_TO_LUMA) != 0;   // RGB host images, one-channel plan
  const int dtype = dtype_in & 0xf;
  const long long n = (long long)pl->B * pl->in_stride * (luma ? 3 : 1);
  if (dtype == 0 && !luma) {
    ICA_CUDA_CHECK(cudaMemcpyAsync(pl->in1_dev, h1, n * sizeof(float), cudaMemcpyHostToDevice, stream));
    ICA_CUDA_CHECK(cudaMemcpyAsync(pl->in2_dev, h2, n * sizeof(float), cudaMemcpyHostToDevice, stream));
    ICA_CUDA_CHECK(cudaEventRecord(pl->ev_h2d_done, stream));
    return ICA_OK;
  }
  const size_t esz = dtype == 1 ? 1 : (dtype == 2 ? 8 : 4);
  const size_t area = ((size_t)n * esz + 255) / 256 * 256;
  if (pl->raw_bytes < 2 * area) {
    cudaFree(pl->raw_dev); pl->raw_dev = nullptr; pl->device_bytes -= pl->raw_bytes; pl->raw_bytes = 0;
    cudaError_t e = cudaMalloc(&pl->raw_dev, 2 * area);
    if (e != cudaSuccess) { set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); return ICA_ERR_ALLOC; }
    pl->raw_bytes = 2 * area; pl->device_bytes += pl->raw_bytes;
  }
  unsigned char* r1 = static_cast<unsigned char*>(pl->raw_dev);
  unsigned char* r2 = r1 + area;
  ICA_CUDA_CHECK(cudaMemcpyAsync(r1, h1, (size_t)n * esz, cudaMemcpyHostToDevice, stream));
  ICA_CUDA_CHECK(cudaMemcpyAsync(r2, h2, (size_t)n * esz, cudaMemcpyHostToDevice, stream));
  ICA_CUDA_CHECK(cudaEventRecord(pl->ev_h2d_done, stream));
  // conversion to float32 fused with the level-0 min/max of every image (saves the separate pass over level 0)
  const int ns = pl->nscales;
  ICA_LAUNCH_CHECK(launch_minmax_reset(pl->mm, pl->B * ns * 2, stream));
  if (luma) {
    ICA_LAUNCH_CHECK(launch_convert_luma(r1, dtype, pl->in1_dev, pl->in_stride, pl->B, pl->mm + 0, ns * 2, stream));
    ICA_LAUNCH_CHECK(launch_convert_luma(r2, dtype, pl->in2_dev, pl->in_stride, pl->B, pl->mm + 1, ns * 2, stream));
  } else if (dtype == 1) {
    ICA_LAUNCH_CHECK(launch_convert_u8(r1, pl->in1_dev, pl->in_stride, pl->B, pl->mm + 0, ns * 2, stream));
    ICA_LAUNCH_CHECK(launch_convert_u8(r2, pl->in2_dev, pl->in_stride, pl->B, pl->mm + 1, ns * 2, stream));
  } else {
    ICA_LAUNCH_CHECK(launch_convert_f64(reinterpret_cast<const double*>(r1), pl->in1_dev, pl->in_stride, pl->B, pl->mm + 0, ns * 2, stream));
    ICA_LAUNCH_CHECK(launch_convert_f64(reinterpret_cast<const double*>(r2), pl->in2_dev, pl->in_stride, pl->B, pl->mm + 1, ns * 2, stream));
  }
  pl->mm_ready = true;
  return ICA_OK;
}

int ica_plan_run_host(ica_plan* pl, const void* I1_host, const void* I2_host, int32_t dtype, double* p_inout_host,
                      double* err_out, int32_t* iters_out, float* DI_out, float* Iw_out) {
  if (!pl || !I1_host || !I2_host || !p_inout_host) { set_error("NULL argument"); return ICA_ERR_INVALID; }
  const bool out64 = (dtype & ICA_DTYPE_OUT_F64) != 0;     // DI_out / Iw_out are float64 arrays (widened on the device)
  dtype &= ~ICA_DTYPE_OUT_F64;
  if ((dtype & ~(0xf | ICA_DTYPE_RGB_TO_LUMA)) || (dtype & 0xf) > 2) { set_error("dtype must be 0 (f32), 1 (u8) or 2 (f64), optionally | ICA_DTYPE_RGB_TO_LUMA"); return ICA_ERR_INVALID; }
  if ((dtype & ICA_DTYPE_RGB_TO_LUMA) && pl->C != 1) { set_error("ICA_DTYPE_RGB_TO_LUMA needs a one-channel plan (the host images are RGB)"); return ICA_ERR_INVALID; }
  if ((DI_out || Iw_out) && !(pl->cfg.flags & ICA_FLAG_WRITE_DI_IW)) {
    set_error("DI/Iw requested but the plan was created without ICA_FLAG_WRITE_DI_IW"); return ICA_ERR_INVALID;
  }
  if (int rc = enter_plan_device(pl)) return rc;
  cudaStream_t stream = pl->stream;
  const size_t nimg = (size_t)pl->B * pl->in_stride;
  if (!pl->in1_dev) {
    if (int rc = dev_alloc(pl, &pl->in1_dev, nimg)) return rc;
    if (int rc = dev_alloc(pl, &pl->in2_dev, nimg)) return rc;
  }
  {
    std::lock_guard<std::mutex> lock(g_h2d_mutex[pl->device & 63]);
    ICA_CUDA_CHECK(cudaEventRecord(pl->ev_host0, stream));
    ICA_CUDA_CHECK(cudaMemcpyAsync(pl->p_dev, p_inout_host, (size_t)pl->B * ICA_MAX_PARAMS * sizeof(double),
                                   cudaMemcpyHostToDevice, stream));
    if (int rc = upload_pair(pl, I1_host, I2_host, dtype, stream)) return rc;
    if (int rc = wait_event_polite(pl->ev_h2d_done)) return rc;   // the link is free for the next caller
  }
  const int rc_run = ica_plan_run_device(pl, pl->in1_dev, pl->in2_dev, pl->p_dev, stream);
  pl->mm_ready = false;
  if (rc_run) return rc_run;
  pl->launches += (dtype == 0 ? 0 : 3);   // key reset + the two conversion kernels (0 | luma counts too: dtype != 0)
  ICA_CUDA_CHECK(cudaMemcpyAsync(p_inout_host, pl->p_dev, (size_t)pl->B * ICA_MAX_PARAMS * sizeof(double),
                                 cudaMemcpyDeviceToHost, stream));
  if (err_out) ICA_CUDA_CHECK(cudaMemcpyAsync(err_out, pl->err_dev, pl->B * sizeof(double), cudaMemcpyDeviceToHost, stream));
  if (iters_out) ICA_CUDA_CHECK(cudaMemcpyAsync(iters_out, pl->iters_dev, (size_t)pl->B * pl->nscales * sizeof(int),
                                                cudaMemcpyDeviceToHost, stream));
  if (out64 && (DI_out || Iw_out)) {
    if (!pl->out64_dev) { if (int rc = dev_alloc(pl, &pl->out64_dev, 2 * nimg)) return rc; }
    if (DI_out) {
      ICA_LAUNCH_CHECK(launch_widen_f64(pl->DI_dev, pl->out64_dev, (long long)nimg, stream));
      ICA_CUDA_CHECK(cudaMemcpyAsync(DI_out, pl->out64_dev, nimg * sizeof(double), cudaMemcpyDeviceToHost, stream));
    }
    if (Iw_out) {
      ICA_LAUNCH_CHECK(launch_widen_f64(pl->Iw_dev, pl->out64_dev + nimg, (long long)nimg, stream));
      ICA_CUDA_CHECK(cudaMemcpyAsync(Iw_out, pl->out64_dev + nimg, nimg * sizeof(double), cudaMemcpyDeviceToHost, stream));
    }
  } else {
    if (DI_out) ICA_CUDA_CHECK(cudaMemcpyAsync(DI_out, pl->DI_dev, nimg * sizeof(float), cudaMemcpyDeviceToHost, stream));
    if (Iw_out) ICA_CUDA_CHECK(cudaMemcpyAsync(Iw_out, pl->Iw_dev, nimg * sizeof(float), cudaMemcpyDeviceToHost, stream));
  }
  ICA_CUDA_CHECK(cudaEventRecord(pl->ev_host1, stream));
  if (int rc = wait_event_polite(pl->ev_host1)) return rc;
  return ICA_OK;
}

int ica_plan_last_host_run_ms(ica_plan* pl, float* ms_out) {
  if (!pl || !ms_out) return ICA_ERR_INVALID;
  ICA_CUDA_CHECK(cudaEventElapsedTime(ms_out, pl->ev_host0, pl->ev_host1));
  return ICA_OK;
}

int ica_plan_debug_timeline(ica_plan* pl, long long* host_out, int32_t enable) {
  if (!pl) return ICA_ERR_INVALID;
  const size_t n = (size_t)(pl->grid + 1) * 16;
  if (enable && !pl->dbg_time) {
    if (int rc = dev_alloc(pl, &pl->dbg_time, n)) return rc;
    ICA_CUDA_CHECK(cudaMemset(pl->dbg_time, 0, n * sizeof(long long)));
  }
  if (host_out && pl->dbg_time) {
    ICA_CUDA_CHECK(cudaDeviceSynchronize());
    ICA_CUDA_CHECK(cudaMemcpy(host_out, pl->dbg_time, n * sizeof(long long), cudaMemcpyDeviceToHost));
  }
  if (!enable && pl->dbg_time) { cudaFree(pl->dbg_time); pl->dbg_time = nullptr; }
  // the loop graph bakes the kernel arguments: rebuild it next run
  if (pl->graph_exec) { cudaGraphExecDestroy(pl->graph_exec); pl->graph_exec = nullptr; }
  return pl->grid + 1;
}

int ica_plan_get_results(ica_plan* pl, double* p_out, double* err_out, int32_t* iters_out) {
  if (!pl) return ICA_ERR_INVALID;
  ICA_CUDA_CHECK(cudaDeviceSynchronize());
  ICA_LAUNCH_CHECK(launch_export_results(pl->state, pl->B, pl->p_dev, pl->err_dev, pl->iters_dev, pl->nscales, 0));
  if (p_out) ICA_CUDA_CHECK(cudaMemcpy(p_out, pl->p_dev, (size_t)pl->B * ICA_MAX_PARAMS * sizeof(double), cudaMemcpyDeviceToHost));
  if (err_out) ICA_CUDA_CHECK(cudaMemcpy(err_out, pl->err_dev, pl->B * sizeof(double), cudaMemcpyDeviceToHost));
  if (iters_out) ICA_CUDA_CHECK(cudaMemcpy(iters_out, pl->iters_dev, (size_t)pl->B * pl->nscales * sizeof(int), cudaMemcpyDeviceToHost));
  return ICA_OK;
}

int ica_plan_get_trajectory(ica_plan* pl, double* traj_out, int32_t* count_out) {
  if (!pl || !pl->traj) { set_error("plan was created without ICA_FLAG_RECORD_TRAJECTORY"); return ICA_ERR_INVALID; }
  ICA_CUDA_CHECK(cudaDeviceSynchronize());
  if (traj_out) ICA_CUDA_CHECK(cudaMemcpy(traj_out, pl->traj, (size_t)pl->B * pl->traj_cap * ICA_TRAJ_STRIDE * sizeof(double), cudaMemcpyDeviceToHost));
  if (count_out) {
    std::vector<PairState> h(pl->B);
    ICA_CUDA_CHECK(cudaMemcpy(h.data(), pl->state, pl->B * sizeof(PairState), cudaMemcpyDeviceToHost));
    for (int b = 0; b < pl->B; ++b) count_out[b] = h[b].traj_count;
  }
  return ICA_OK;
}

int ica_plan_get_di_iw_device(ica_plan* pl, const float** DI_dev, const float** Iw_dev) {
  if (!pl || !pl->DI_dev) { set_error("plan was created without ICA_FLAG_WRITE_DI_IW"); return ICA_ERR_INVALID; }
  if (DI_dev) *DI_dev = pl->DI_dev;
  if (Iw_dev) *Iw_dev = pl->Iw_dev;
  return ICA_OK;
}

int ica_plan_get_level_device(ica_plan* pl, int32_t which, int32_t pair, int32_t scale, const float** ptr_out,
                              int32_t* pitch_out) {
  if (!pl || which < 0 || which > 1 || pair < 0 || pair >= pl->B || scale < 0 || scale >= pl->nscales || !pl->last_I1) {
    set_error("bad level query"); return ICA_ERR_INVALID;
  }
  const float* base0 = which == 0 ? pl->last_I1 : pl->last_I2;
  const float* pyr = which == 0 ? pl->pyr1 : pl->pyr2;
  if (ptr_out) *ptr_out = scale == 0 ? base0 + (long long)pair * pl->in_stride
                                     : pyr + (long long)pair * pl->pyr_stride + pl->lv[scale].offset;
  if (pitch_out) *pitch_out = pl->lv[scale].pitch;
  return ICA_OK;
}

// ------------------------------------------------------------------ stateless entry points
int ica_zoom_size(int32_t nx, int32_t ny, double factor, int32_t* nxx, int32_t* nyy) {
  if (nxx) *nxx = zoomed_size(nx, factor);
  if (nyy) *nyy = zoomed_size(ny, factor);
  return ICA_OK;
}

int ica_resample_operator(int32_t n_in, int32_t n_out, int32_t* taps_out, int32_t* start_out, float* weights_out,
                          int32_t weights_capacity, int32_t* fast_range_out) {
  if (n_in < 1 || n_out < 1 || !taps_out) { set_error("bad argument"); return ICA_ERR_INVALID; }
  Resample1D r;
  build_resample_1d(n_in, n_out, &r);
  *taps_out = r.taps;
  if (start_out) for (int o = 0; o < n_out; ++o) start_out[o] = r.start[o];
  if (weights_out) {
    if ((long long)weights_capacity < (long long)n_out * r.taps) { set_error("weights buffer too small"); return ICA_ERR_INVALID; }
    for (size_t i = 0; i < r.weights.size(); ++i) weights_out[i] = r.weights[i];
  }
  if (fast_range_out) {
    FastRows f;
    detect_uniform_rows(r, &f);
    fast_range_out[0] = f.lo; fast_range_out[1] = f.hi; fast_range_out[2] = f.s0;
  }
  return ICA_OK;
}

// The banded operator of zoom.zoom_out (src/zoom.py:29-60) along one axis of length n_in, built on the host in fp64
// (no device needed); same output convention as ica_resample_operator.
int ica_zoom_out_operator(int32_t n_in, double factor, int32_t* n_out, int32_t* taps_out, int32_t* start_out,
                          float* weights_out, int32_t weights_capacity) {
  if (n_in < 1 || !(factor > 0.0 && factor < 1.0) || !taps_out || !n_out) { set_error("bad argument"); return ICA_ERR_INVALID; }
  Resample1D r;
  build_zoom_out_1d(n_in, factor, 0.6 /* constants.ZOOM_SIGMA_ZERO */, &r);
  *taps_out = r.taps; *n_out = r.n_out;
  if (start_out) for (int o = 0; o < r.n_out; ++o) start_out[o] = r.start[o];
  if (weights_out) {
    if ((long long)weights_capacity < (long long)r.n_out * r.taps) { set_error("weights buffer too small"); return ICA_ERR_INVALID; }
    for (size_t i = 0; i < r.weights.size(); ++i) weights_out[i] = r.weights[i];
  }
  return ICA_OK;
}

int ica_apply_operators_host(const float* image, int32_t height, int32_t width, int32_t channels,
                             const int32_t* ystart, const float* yweights, int32_t ytaps, int32_t ny_out,
                             const int32_t* xstart, const float* xweights, int32_t xtaps, int32_t nx_out,
                             int32_t clip_to_input_range, float* out);

// zoom.zoom_out (src/zoom.py:29-60) of one float32 image [H][W][C]: both operators built here (C++), applied by the
// pyramid kernels; out is [round(H f)][round(W f)][C], no clipping (the reference function does not clip).
int ica_zoom_out_host(const float* image, int32_t height, int32_t width, int32_t channels, double factor, float* out,
                      int32_t* out_h, int32_t* out_w) {
  if (!image || !out || height < 1 || width < 1 || (channels != 1 && channels != 3) || !(factor > 0.0 && factor < 1.0)) {
    set_error("bad argument"); return ICA_ERR_INVALID;
  }
  Resample1D ry, rx;
  build_zoom_out_1d(height, factor, 0.6, &ry);
  build_zoom_out_1d(width, factor, 0.6, &rx);
  if (out_h) *out_h = ry.n_out;
  if (out_w) *out_w = rx.n_out;
  return ica_apply_operators_host(image, height, width, channels, ry.start.data(), ry.weights.data(), ry.taps, ry.n_out,
                                  rx.start.data(), rx.weights.data(), rx.taps, rx.n_out, 0, out);
}

// out = A_y * image * A_x^T for caller-supplied banded operators (rows of `taps` weights starting at start[o]); the
// two-pass / fused kernels of the pyramid do the work.  Used by the Python mirror of zoom.zoom_out.
int ica_apply_operators_host(const float* image, int32_t height, int32_t width, int32_t channels,
                             const int32_t* ystart, const float* yweights, int32_t ytaps, int32_t ny_out,
                             const int32_t* xstart, const float* xweights, int32_t xtaps, int32_t nx_out,
                             int32_t clip_to_input_range, float* out) {
  if (!image || !out || !ystart || !yweights || !xstart || !xweights || height < 1 || width < 1 || (channels != 1 && channels != 3) ||
      ny_out < 1 || nx_out < 1 || ytaps < 1 || xtaps < 1) { set_error("bad argument"); return ICA_ERR_INVALID; }
  if (ytaps > max_taps() || xtaps > max_taps() || ytaps > height || xtaps > width) {
    set_error("operator band too wide (%d / %d taps; at most %d and the image size)", ytaps, xtaps, max_taps());
    return ICA_ERR_INVALID;
  }
  for (int o = 0; o < ny_out; ++o) if (ystart[o] < 0 || ystart[o] + ytaps > height) { set_error("row operator reads outside the image"); return ICA_ERR_INVALID; }
  for (int o = 0; o < nx_out; ++o) if (xstart[o] < 0 || xstart[o] + xtaps > width) { set_error("column operator reads outside the image"); return ICA_ERR_INVALID; }
  if (int rc = require_device()) return rc;
  Resample1D ry, rx;
  ry.n_in = height; ry.n_out = ny_out; ry.taps = ytaps; ry.start.assign(ystart, ystart + ny_out);
  ry.weights.assign(yweights, yweights + (size_t)ny_out * ytaps);
  rx.n_in = width; rx.n_out = nx_out; rx.taps = xtaps; rx.start.assign(xstart, xstart + nx_out);
  rx.weights.assign(xweights, xweights + (size_t)nx_out * xtaps);
  DeviceResample dy, dx;
  float *d_in = nullptr, *d_out = nullptr, *d_tmp = nullptr;
  MinMaxKeys* d_mm = nullptr;
  const size_t n_in = (size_t)height * width * channels;
  const int out_pitch = (nx_out * channels + 3) / 4 * 4;
  const size_t n_out = (size_t)ny_out * out_pitch;
  const long long tmp_stride = ((long long)ny_out * width * channels + 3) / 4 * 4;
  int rc = upload_resample(nullptr, ry, &dy);
  if (!rc) rc = upload_resample(nullptr, rx, &dx);
  if (!rc) rc = dev_alloc<float>(nullptr, &d_in, n_in);
  if (!rc) rc = dev_alloc<float>(nullptr, &d_out, 2 * n_out);
  if (!rc) rc = dev_alloc<float>(nullptr, &d_tmp, (size_t)(2 * tmp_stride));
  if (!rc) rc = dev_alloc<MinMaxKeys>(nullptr, &d_mm, 4);
  cudaError_t e = cudaSuccess;
  if (!rc) {
    e = cudaMemcpy(d_in, image, n_in * sizeof(float), cudaMemcpyHostToDevice);
    // keys: [0],[1] parent (image a, b), [2],[3] child
    MinMaxKeys open_range[4];
    for (auto& k : open_range) { k.lo = float_key(-3.4028235e38f); k.hi = float_key(3.4028235e38f); }
    if (e == cudaSuccess) e = cudaMemcpy(d_mm, open_range, sizeof(open_range), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && clip_to_input_range) {
      e = launch_minmax_reset(d_mm, 2, 0);
      if (e == cudaSuccess) e = launch_minmax(d_in, 0, (long long)n_in, 2, d_mm, 1, 0);   // the same image stands for both sets
    }
    if (e == cudaSuccess)
      e = launch_pyr_down(d_in, d_in, (long long)n_in, width * channels, width, height, channels, dy, dx, d_tmp, tmp_stride, d_out,
                          d_out + n_out, (long long)n_out, out_pitch, 1, d_mm, d_mm + 2, 2, 0, nullptr);
    if (e == cudaSuccess)
      e = cudaMemcpy2D(out, (size_t)nx_out * channels * sizeof(float), d_out, (size_t)out_pitch * sizeof(float),
                       (size_t)nx_out * channels * sizeof(float), ny_out, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("ica_apply_operators_host: %s", cudaGetErrorString(e)); rc = ICA_ERR_CUDA; }
  }
  free_resample(&dy); free_resample(&dx);
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_tmp); cudaFree(d_mm);
  return rc;
}

int ica_nparams(int32_t t) { int n = nparams_of(t); if (n < 0) set_error("Unknown transform type"); return n < 0 ? ICA_ERR_INVALID : n; }

int ica_params2matrix(const double* p, int32_t t, double* m9) {
  if (!p || !m9 || nparams_of(t) < 0) { set_error("Unknown transform type"); return ICA_ERR_INVALID; }
  params2matrix(p, t, m9);
  return ICA_OK;
}

int ica_update_transform(double* p, const double* dp, int32_t t) {
  if (!p || !dp || nparams_of(t) < 0) { set_error("Unknown transform type"); return ICA_ERR_INVALID; }
  update_transform(p, dp, t);
  return ICA_OK;
}

int ica_zoom_in_parameters(const double* p, int32_t t, double nx, double ny, double nxx, double nyy, double* out) {
  if (!p || !out || nparams_of(t) < 0) { set_error("Unsupported transformation type"); return ICA_ERR_INVALID; }
  zoom_in_parameters(p, t, nx, ny, nxx, nyy, out);
  return ICA_OK;
}

int ica_inverse_hessian(const double* H, int32_t n, double* Hinv) {
  if (!H || !Hinv || n < 1 || n > ICA_MAX_PARAMS) { set_error("n must be in [1, 8]"); return ICA_ERR_INVALID; }
  inverse_hessian(H, n, Hinv);
  return ICA_OK;
}

int ica_warp_host(const float* image, int32_t height, int32_t width, int32_t channels, const double* m9, float* out) {
  if (!image || !m9 || !out || height < 1 || width < 1 || (channels != 1 && channels != 3)) { set_error("bad argument"); return ICA_ERR_INVALID; }
  if (int rc = require_device()) return rc;
  const size_t n = (size_t)height * width * channels;
  float *d_in = nullptr, *d_out = nullptr; MinMaxKeys* d_mm = nullptr;
  int rc = ICA_OK;
  if ((rc = dev_alloc<float>(nullptr, &d_in, n)) || (rc = dev_alloc<float>(nullptr, &d_out, n)) || (rc = dev_alloc<MinMaxKeys>(nullptr, &d_mm, 1))) {
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_mm); return rc;
  }
  cudaError_t e = cudaMemcpy(d_in, image, n * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_minmax_reset(d_mm, 1, 0);
  if (e == cudaSuccess) e = launch_minmax(d_in, 0, (long long)n, 1, d_mm, 1, 0);
  if (e == cudaSuccess) e = launch_warp_matrix(d_in, width, height, channels, m9, d_mm, d_out, 0);
  if (e == cudaSuccess) e = cudaMemcpy(out, d_out, n * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_mm);
  if (e != cudaSuccess) { set_error("ica_warp_host: %s", cudaGetErrorString(e)); return ICA_ERR_CUDA; }
  return ICA_OK;
}

int ica_gradient_host(const float* image, int32_t height, int32_t width, int32_t channels, int32_t delta,
                      int32_t nanifoutside, float* Ix, float* Iy) {
  if (!image || !Ix || !Iy || height < 1 || width < 1 || (channels != 1 && channels != 3)) { set_error("bad argument"); return ICA_ERR_INVALID; }
  if (int rc = require_device()) return rc;
  const size_t n = (size_t)height * width * channels;
  float *d_in = nullptr, *d_x = nullptr, *d_y = nullptr;
  int rc = ICA_OK;
  if ((rc = dev_alloc<float>(nullptr, &d_in, n)) || (rc = dev_alloc<float>(nullptr, &d_x, n)) || (rc = dev_alloc<float>(nullptr, &d_y, n))) {
    cudaFree(d_in); cudaFree(d_x); cudaFree(d_y); return rc;
  }
  cudaError_t e = cudaMemcpy(d_in, image, n * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_gradient(d_in, width, height, channels, delta, (nanifoutside && delta > 0) ? 1 : 0, d_x, d_y, 0);
  if (e == cudaSuccess) e = cudaMemcpy(Ix, d_x, n * sizeof(float), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(Iy, d_y, n * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(d_in); cudaFree(d_x); cudaFree(d_y);
  if (e != cudaSuccess) { set_error("ica_gradient_host: %s", cudaGetErrorString(e)); return ICA_ERR_CUDA; }
  return ICA_OK;
}

int ica_rescale_host(const float* image, int32_t height, int32_t width, int32_t channels, double nu, float* out,
                     int32_t* out_h, int32_t* out_w) {
  if (!image || !out || height < 1 || width < 1 || (channels != 1 && channels != 3) || !(nu > 0.0 && nu < 1.0)) {
    set_error("bad argument"); return ICA_ERR_INVALID;
  }
  ica_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.batch = 1; cfg.height = height; cfg.width = width; cfg.channels = channels; cfg.nscales = 2; cfg.nu = nu;
  cfg.transform_type = TRANSLATION; cfg.robust_type = QUADRATIC; cfg.tol = 1e-3; cfg.max_iter = 1;
  ica_plan* pl = nullptr;
  if (int rc = ica_plan_create(&cfg, &pl)) return rc;
  const size_t n = (size_t)height * width * channels;
  int rc = dev_alloc(pl, &pl->in1_dev, n);
  if (!rc) rc = dev_alloc(pl, &pl->in2_dev, n);
  if (!rc) {
    cudaError_t e = cudaMemcpy(pl->in1_dev, image, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(pl->in2_dev, image, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("ica_rescale_host: %s", cudaGetErrorString(e)); rc = ICA_ERR_CUDA; }
  }
  if (!rc) rc = build_pyramids(pl, pl->in1_dev, pl->in2_dev, 0);
  if (!rc) {
    const LevelDesc& L = pl->lv[1];
    cudaError_t e = cudaMemcpy2D(out, (size_t)L.nx * channels * sizeof(float), pl->pyr1 + L.offset,
                                 (size_t)L.pitch * sizeof(float), (size_t)L.nx * channels * sizeof(float), L.ny,
                                 cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("ica_rescale_host: %s", cudaGetErrorString(e)); rc = ICA_ERR_CUDA; }
    if (out_h) *out_h = L.ny;
    if (out_w) *out_w = L.nx;
  }
  ica_plan_destroy(pl);
  return rc;
}

int ica_hessian_b_host(const float* I1, const float* I2, int32_t height, int32_t width, int32_t channels,
                       int32_t gray_as_rgb, int32_t transform_type, const double* p, int32_t robust_type,
                       double lambda_, int32_t delta, int32_t nanifoutside, double* H_out, double* b_out) {
  if (!I1 || !I2 || !p || !H_out || !b_out) { set_error("NULL argument"); return ICA_ERR_INVALID; }
  ica_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.batch = 1; cfg.height = height; cfg.width = width; cfg.channels = channels; cfg.gray_as_rgb = gray_as_rgb;
  cfg.nscales = 1; cfg.nu = 0.5; cfg.transform_type = transform_type; cfg.robust_type = robust_type;
  cfg.robust_loop = 1; cfg.lambda_ = lambda_ > 0 ? lambda_ : kLambda0; cfg.tol = 1e-3; cfg.max_iter = 1;
  cfg.delta = delta; cfg.nanifoutside = nanifoutside;
  ica_plan* pl = nullptr;
  if (int rc = ica_plan_create(&cfg, &pl)) return rc;
  const size_t n = (size_t)height * width * channels;
  const int np = nparams_of(transform_type);
  double ph[ICA_MAX_PARAMS] = {0};
  for (int i = 0; i < np; ++i) ph[i] = p[i];
  double* d_dbg = nullptr;
  int rc = dev_alloc(pl, &pl->in1_dev, n);
  if (!rc) rc = dev_alloc(pl, &pl->in2_dev, n);
  if (!rc) rc = dev_alloc(pl, &d_dbg, 72);
  cudaError_t e = cudaSuccess;
  if (!rc) {
    e = cudaMemcpy(pl->in1_dev, I1, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(pl->in2_dev, I2, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(pl->p_dev, ph, sizeof(ph), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("ica_hessian_b_host: %s", cudaGetErrorString(e)); rc = ICA_ERR_CUDA; }
  }
  if (!rc) rc = prepare_level0(pl, pl->in1_dev, pl->in2_dev, 0);
  if (!rc) rc = build_pyramids(pl, pl->in1_dev, pl->in2_dev, 0);
  if (!rc) {
    e = launch_init_state(pl->state, pl->p_dev, pl->ttypes_dev, 1, 1, cfg.lambda_, pl->n_active, pl->pair_ticket, 0);
    IterParams P;
    fill_iter_params(pl, pl->in1_dev, pl->in2_dev, &P);
    P.dbg_Hb = d_dbg;
    if (e == cudaSuccess) e = launch_schedule(P, 0);
    if (e == cudaSuccess) e = launch_k2(pl, P, 0);
    if (e == cudaSuccess) e = launch_solve(P, pl->dh, 0);
    double hb[72];
    if (e == cudaSuccess) e = cudaMemcpy(hb, d_dbg, sizeof(hb), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("ica_hessian_b_host: %s", cudaGetErrorString(e)); rc = ICA_ERR_CUDA; }
    else { memcpy(H_out, hb, sizeof(double) * np * np); memcpy(b_out, hb + 64, sizeof(double) * np); }
  }
  cudaFree(d_dbg);
  ica_plan_destroy(pl);
  return rc;
}

}  // extern "C"
