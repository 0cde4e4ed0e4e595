// Host-visible interface of ica_pyramid.cu
#pragma once
#include <vector>
#include "ica_common.cuh"

namespace ica {

// Banded 1-D resampling operator (host): out[o] = sum_k weights[o*taps+k] * in[start[o]+k]
struct Resample1D {
  int n_in = 0, n_out = 0, taps = 0;
  std::vector<int> start;
  std::vector<float> weights;
};

constexpr int kFastTapsMax = 40;
// Output rows [lo, hi) of an operator that share one weight vector: start[o] = 2*o + s0 (exact 2:1 interior)
struct FastRows { int lo = 0, hi = 0, s0 = 0, taps = 0; float w[kFastTapsMax] = {0}; };

// The same operator on the device; weights_t is the [taps][n_out] transpose for the horizontal pass
struct DeviceResample {
  int n_in = 0, n_out = 0, taps = 0;
  int* start = nullptr;
  float* weights = nullptr;
  float* weights_t = nullptr;
  FastRows fast;
  std::vector<int> start_host;   // host copy of start[] (footprints of the border outputs)
};

int round_half_even(double v);
int zoomed_size(int n, double factor);   // src/zoom.py:8-22
void build_resample_1d(int n_in, int n_out, Resample1D* out, double rel_threshold = 1e-9);
// zoom.zoom_out (src/zoom.py:29-60) along one axis: Gaussian (reflect) + cubic-spline resampling at o / factor (nearest)
void build_zoom_out_1d(int n_in, double factor, double sigma_zero, Resample1D* out, double rel_threshold = 1e-9);
void detect_uniform_rows(const Resample1D& r, FastRows* f);
int max_taps();

cudaError_t launch_minmax_reset(MinMaxKeys* mm, int count, cudaStream_t stream);
cudaError_t launch_minmax(const float* img0, long long stride, long long count, int nimg, MinMaxKeys* mm,
                          int mm_stride, cudaStream_t stream);
// one level for `nset` pairs, both image sets per launch; tmp holds ny_out x (nx_in*C) floats per image (2*nset images)
cudaError_t launch_pyr_down(const float* in0a, const float* in0b, long long in_stride, int in_pitch, int nx_in, int ny_in,
                            int channels, const DeviceResample& ry, const DeviceResample& rx, float* tmp,
                            long long tmp_stride, float* out0a, float* out0b, long long out_stride, int out_pitch, int nset,
                            const MinMaxKeys* mm_parent, MinMaxKeys* mm_child, int mm_stride, cudaStream_t stream,
                            int* launches = nullptr, cudaStream_t border_stream = nullptr, MinMaxKeys* mm_gather = nullptr,
                            const MinMaxKeys* mm_open = nullptr);
// gather mode of launch_pyr_down (mm_gather != nullptr): the level's first read also produces the PARENT's min / max
// (no separate pass over the parent), nothing is clipped, and launch_clip_fixup clamps afterwards where needed
bool pyr_level_gathers(const float* in0a, const float* in0b, long long in_stride, int in_pitch, int nx_in, int ny_in, int channels,
                       const DeviceResample& ry, const DeviceResample& rx, const float* tmp, long long tmp_stride);
cudaError_t launch_clip_fixup(float* out0a, float* out0b, long long out_stride, int out_pitch, int nx_out, int ny_out, int channels,
                              int nset, const MinMaxKeys* mm_parent, MinMaxKeys* mm_child, int mm_stride, int* flags,
                              cudaStream_t stream);
// uint8 / float64 -> float32 for nimg images of `count` elements, fused with each image's min/max keys
cudaError_t launch_widen_f64(const float* in, double* out, long long n, cudaStream_t stream);
// RGB host images of dtype (0 f32, 1 u8, 2 f64) -> one-channel float32 luminance + its min/max
cudaError_t launch_convert_luma(const void* in, int dtype, float* out, long long npix, int nimg, MinMaxKeys* mm, int mm_stride,
                                cudaStream_t stream);
cudaError_t launch_convert_u8(const unsigned char* in, float* out, long long count, int nimg, MinMaxKeys* mm, int mm_stride,
                              cudaStream_t stream);
cudaError_t launch_convert_f64(const double* in, float* out, long long count, int nimg, MinMaxKeys* mm, int mm_stride,
                               cudaStream_t stream);

}  // namespace ica
