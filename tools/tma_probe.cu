// Probe: 2-D tiled TMA load of a float image with out-of-bounds fill, tensor map passed (A) as a __grid_constant__
// kernel parameter and (B) through global memory.  nvcc -arch=sm_100a tools/tma_probe.cu -o /tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cmath>
#include <cstdlib>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__constant__ int BW_d, BH_d;
static int BW = 256, BH = 22;

__device__ void do_load(const void* tmap, float* out, int c0, int c1) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BW_d * BH_d * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(sm)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
  for (int i = threadIdx.x; i < BW_d * BH_d; i += blockDim.x) out[i] = sm[i];
}

__global__ void kernel_param(const __grid_constant__ CUtensorMap tm, float* out, int c0, int c1) { do_load(&tm, out, c0, c1); }
__global__ void kernel_global(const void* tm, float* out, int c0, int c1) { do_load(tm, out, c0, c1); }

int main(int argc, char** argv) {
  BW = argc > 1 ? atoi(argv[1]) : 256; BH = argc > 2 ? atoi(argv[2]) : 22;
  const int fillnan = argc > 3 ? atoi(argv[3]) : 1;
  const int c0 = argc > 4 ? atoi(argv[4]) : -6, c1 = argc > 5 ? atoi(argv[5]) : -2;
  const int swz = argc > 6 ? atoi(argv[6]) : 0;
  cudaMemcpyToSymbol(BW_d, &BW, 4); cudaMemcpyToSymbol(BH_d, &BH, 4);
  printf("box %d x %d fillnan %d c0 %d c1 %d swizzle %d\n", BW, BH, fillnan, c0, c1, swz);
  const int nx = 100, ny = 40, C = 3, pitch = 300;
  std::vector<float> img((size_t)ny * pitch);
  for (int y = 0; y < ny; ++y) for (int i = 0; i < pitch; ++i) img[(size_t)y * pitch + i] = y * 1000 + i;
  float* d_img; cudaMalloc(&d_img, img.size() * 4); cudaMemcpy(d_img, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
  float* d_out; cudaMalloc(&d_out, BW * BH * 4);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(e), (int)q, fnp);
  EncodeTiledFn fn = (EncodeTiledFn)fnp;
  alignas(64) CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)nx * C, (cuuint64_t)ny}, gstr[1] = {(cuuint64_t)pitch * 4};
  cuuint32_t box[2] = {(cuuint32_t)BW, (cuuint32_t)BH}, estr[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_img, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  fillnan ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  void* d_tm; cudaMalloc(&d_tm, 128); cudaMemcpy(d_tm, &m, 128, cudaMemcpyHostToDevice);
  std::vector<float> out(BW * BH);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(d_out, 0, BW * BH * 4);
    if (variant == 0) kernel_param<<<1, 128, BW * BH * 4>>>(m, d_out, c0, c1);
    else kernel_global<<<1, 128, BW * BH * 4>>>(d_tm, d_out, c0, c1);
    e = cudaDeviceSynchronize();
    printf("variant %s: %s\n", variant ? "global" : "param", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(out.data(), d_out, BW * BH * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r_ = 0; r_ < BH; ++r_) for (int i = 0; i < BW; ++i) {
      const int y = c1 + r_, x = c0 + i;
      const float v = out[r_ * BW + i];
      const bool inside = y >= 0 && y < ny && x >= 0 && x < nx * C;
      if (inside ? (v != y * 1000 + x) : (fillnan ? !std::isnan(v) : v != 0.f)) { if (bad < 5) printf("  mismatch r=%d i=%d v=%f\n", r_, i, v); ++bad; }
    }
    printf("  mismatches: %d (row 2: %f %f %f ... %f)\n", bad, out[2 * BW + 5], out[2 * BW + 6], out[2 * BW + 7], out[2 * BW + 255]);
  }
  return 0;
}
