// Host-visible interface of ica_iterate.cu
#pragma once
#include "ica_common.cuh"

namespace ica {

// H[k][l] (entry k*8+l) and b[k] (entry 64+k) as +-1 combinations of at most 4 moment sums
struct AsmEntry { int idx[4]; float coef[4]; };

constexpr int kXSlot = 128;     // doubles per (rank, pair) exchange slot: 105 moment sums, the last word is the flag

// Header of one (double-buffered) work list
struct SchedHdr {
  int total;              // work items of the list
  int npairs;             // pairs that own at least one item
  int counter;            // next unclaimed item (dynamic work distribution)
  int list;               // which of the two list buffers holds the items (a list that did not change is reused)
  long long t0, t1;       // %globaltimer: first CTA start / last chunk done of the launch that consumed the list
};

struct IterParams {
  const float* I1_0;      // level-0 inputs [B][H][W][C]
  const float* I2_0;
  long long in_stride;    // floats per pair at level 0
  const float* pyr1;      // levels >= 1: [B][pyr_stride]
  const float* pyr2;
  long long pyr_stride;
  LevelDesc lv[ICA_MAX_SCALES];
  int nscales;
  PairState* state;       // [B]
  const MinMaxKeys* mm;   // [B][nscales][2]  (0 = I1, 1 = I2)
  double* partials;       // [B][max_chunks][kAccStride]
  double* traj;           // [B][traj_cap][ICA_TRAJ_STRIDE] or nullptr
  int dbg_row;            // row of dbg_time the solve kernel (block 0) stamps
  long long* dbg_time;    // nullptr, or [grid][16] globaltimer stamps of each CTA's first item (profiling hook)
  double* dbg_Hb;         // nullptr, or 72 doubles: H (<=64) then b (8); state left untouched
  int* n_active;
  // Work lists are double-buffered by the parity of *loop_count (the number of iterations scheduled so far): the list
  // of iteration i+1 is written while stragglers of iteration i may still poll the counter of list i.
  int* chunk_start;       // [2][B+1] exclusive prefix sums of the chunks per pair
  int* item_pair;         // [2][B*max_chunks] pair of each work item
  SchedHdr* hdr;          // [2] totals, work counter and time stamps of each list
  unsigned int* pair_ticket;   // [B] chunks of a pair that are done in the current iteration (fused solve)
  int grid_ctas;          // CTAs of the iterate launch (the work counter starts there: CTA i takes item i first)
  int fused;              // 1: the CTA that finishes a pair's last chunk solves it inside the iterate kernel (no solve launch)
  const AsmEntry* asm_tab;      // [6 transform codes][72]
  const void* tmaps;            // CUtensorMap [B][nscales][2] (I1 zero-filled, I2 NaN-filled boxes), 128 bytes each
  unsigned long long cond_handle;   // cudaGraphConditionalHandle of the while node (0 outside a graph)
  int* loop_count;              // iterations executed in this run (device)
  int max_launches;
  long long* kernel_ns;         // [2] accumulated time of the streaming phase of the iterate kernel (ns: first CTA start ->
                                //     last chunk partial written) and number of launches of this run
  unsigned int* solve_ticket;   // pairs solved in the current iteration (the last one schedules the next)
  int* sched_dirty;             // set when a pair changed scale or finished in this iteration: the work list must be rebuilt
  int shard_rank, shard_n;      // row-sharded mode: this rank's band of tile rows (0, 1 = whole image)
  int solve_mode;               // 0: sum partials + solve; 1: sum partials -> ext_moments only; 2: solve from ext_moments;
                                // 3: sum partials, exchange them with the other ranks through peer memory, solve
  double* ext_moments;          // [B][kAccStride] moment sums exchanged between ranks (row-sharded mode)
  // peer exchange (mode 3): x_peers[r] = rank r's exchange buffer [2 parities][x_world ranks][B pairs][kXSlot doubles],
  // mapped into this process (CUDA IPC over NVLink); slot word kXSlot-1 is the sequence flag of the slot
  double* const* x_peers;
  int x_world, x_rank;
  const unsigned long long* x_seq_base;   // device: first sequence number of this run (monotonic over the plan's life)
  int* x_error;                      // set when a peer did not answer within the time-out
  long long* x_ns;                   // [2] accumulated exchange time (publish -> all peers seen), exchanges
  int B;
  int max_chunks;         // partial slots per pair
  int chunk_unit;         // 0: tile kernel (ica_iterate.cu); > 0: march kernel, consumer warps per CTA (ica_common.cuh: chunk_count)
  int chunk_m;            // march kernel: preferred tiles per consumer warp and chunk
  int traj_cap;
  int robust_type;
  int robust_loop;        // 1: rho' and H every iteration; 0: quadratic loop (H at iter 0 of a scale)
  double lambda_cfg;
  double tol;
  int max_iter;
  int delta;
  int frame;              // nanifoutside && delta > 0
  float ch_mult;          // 3 for a gray image standing for its RGB replication, else 1
  int ipol_warp;          // 1: IPOL-style warp domain (bicubic_interpolation_image, bi.py:121-152): a pixel is valid iff its
                          //    projected point lies in [delta, n-1-delta]; no clip; 0 instead of NaN when ipol_nan == 0
  int ipol_nan;           // IPOL warp: value outside the domain is NaN (discarded) / 0 (kept as a residual of -I1)
};

void build_assembly_table(int dh, AsmEntry* tab /* [6*72] */);
int iterate_tile_w();
int iterate_tile_h();
int iterate_blocks_per_sm();
// TMA boxes of one staged tile, in floats x rows: I1 patch (w1 x h1) and I2 window (w2 x h2)
void iterate_stage_boxes(int channels, int* w1, int* h1, int* w2, int* h2);
cudaError_t launch_schedule(const IterParams& P, cudaStream_t stream);
cudaError_t launch_solve(const IterParams& P, int dh, cudaStream_t stream);
cudaError_t launch_iterate(const IterParams& P, int channels, int dh, int grid, cudaStream_t stream);
// second-generation K2 (ica_march.cu): same contract as launch_iterate, own tile shape and TMA boxes
int march_tile_w();
int march_tile_h();
int march_chunk_unit();
void march_stage_boxes(int channels, int* w1, int* h1, int* w2, int* h2);
cudaError_t launch_march(const IterParams& P, int channels, int dh, int grid, cudaStream_t stream);
cudaError_t launch_init_state(PairState* state, const double* p_in, const int* ttypes, int B, int nscales,
                              double lambda_cfg, int* n_active, unsigned int* pair_ticket, cudaStream_t stream);
cudaError_t launch_export_results(const PairState* state, int B, double* p_out, double* err_out, int* iters_out,
                                  int nscales, cudaStream_t stream);
cudaError_t launch_warp_out(const float* I1_0, const float* I2_0, long long in_stride, int nx, int ny, int channels,
                            const PairState* state, const MinMaxKeys* mm, int nscales, int B, float* Iw, float* DI,
                            int ipol_warp, int ipol_nan, int delta, cudaStream_t stream);
cudaError_t launch_warp_matrix(const float* img, int nx, int ny, int channels, const double* m9, const MinMaxKeys* mm,
                               float* out, cudaStream_t stream);
cudaError_t launch_gradient(const float* img, int nx, int ny, int channels, int delta, int frame, float* Ix,
                            float* Iy, cudaStream_t stream);

}  // namespace ica
