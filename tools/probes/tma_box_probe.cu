// How fast does one SM's TMA unit deliver 4-D tensor boxes with short inner rows?  Every CTA (one per SM) streams
// boxes of a (B, H, W, 3C) bf16 tensor into a 4-deep shared-memory ring: inner extent 32 channels (64 B, SWIZZLE_64B)
// or 64 channels (128 B, SWIZZLE_128B), 64 rows per box.  Prints GB/s for the whole chip.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
constexpr int kC = 96, kH = 64, kW = 64, kB = 128;
// boxes_per_stage boxes of `box_bytes` each per stage; issuers = number of lanes of warp 0 that issue (round-robin)
__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap map, int inner_ch, int boxes_per_stage,
                                                int box_bytes, int iters, int issuers) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[4];
  const uint32_t sb = smem_u32(smem);
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  const int nwin = kB * (kH / 8) * (kW / 8);
  const int nslots = 3 * kC / inner_ch;
  for (int it = 0; it < iters + 4; ++it) {
    if (it >= 4) mbar_wait(smem_u32(&bars[(it - 4) & 3]), ((it - 4) >> 2) & 1);   // consume the stage loaded 4 iterations ago
    if (it < iters) {
      const int s = it & 3;
      if (lane == 0) mbar_expect_tx(smem_u32(&bars[s]), boxes_per_stage * box_bytes);
      __syncwarp();
      for (int b = lane; b < boxes_per_stage; b += issuers) {
        if (lane >= issuers) break;
        const long long id = ((long long)(blockIdx.x + it * gridDim.x) * boxes_per_stage + b);
        const int win = (int)((id / nslots) % nwin), slot = (int)(id % nslots);
        const int bimg = win / 64, wh = (win / 8) % 8, ww = win % 8;
        tma_load_4d(sb + (s * boxes_per_stage + b) * box_bytes, &map, smem_u32(&bars[s]), slot * inner_ch, ww * 8, wh * 8, bimg);
      }
    }
  }
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const size_t n = (size_t)kB * kH * kW * 3 * kC;
  __nv_bfloat16* d;
  CK(cudaMalloc(&d, n * 2));
  CK(cudaMemset(d, 0, n * 2));
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 6 * 8192 + 1024));
  struct Cfg { int inner, rows_w, rows_h, boxes, issuers; CUtensorMapSwizzle sw; const char* name; };
  Cfg cfgs[] = {{32, 8, 8, 6, 1, CU_TENSOR_MAP_SWIZZLE_64B, "64 B rows x 64, 6 boxes/stage, 1 issuer"},
                {32, 8, 8, 6, 6, CU_TENSOR_MAP_SWIZZLE_64B, "64 B rows x 64, 6 boxes/stage, 6 issuers"},
                {64, 8, 8, 3, 1, CU_TENSOR_MAP_SWIZZLE_128B, "128 B rows x 64, 3 boxes/stage, 1 issuer"},
                {64, 8, 8, 6, 6, CU_TENSOR_MAP_SWIZZLE_128B, "128 B rows x 64, 6 boxes/stage, 6 issuers"},
                {32, 8, 4, 12, 12, CU_TENSOR_MAP_SWIZZLE_64B, "64 B rows x 32, 12 boxes/stage, 12 issuers"}};
  for (auto& c : cfgs) {
    CUtensorMap map;
    cuuint64_t dims[4] = {3 * kC, kW, kH, kB};
    cuuint64_t strides[3] = {3 * kC * 2, (cuuint64_t)kW * 3 * kC * 2, (cuuint64_t)kH * kW * 3 * kC * 2};
    cuuint32_t box[4] = {(cuuint32_t)c.inner, (cuuint32_t)c.rows_w, (cuuint32_t)c.rows_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int box_bytes = c.inner * 2 * c.rows_w * c.rows_h;
    const int iters = 400;
    probe<<<148, 128, 4 * c.boxes * box_bytes + 1024>>>(map, c.inner, c.boxes, box_bytes, 20, c.issuers);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    probe<<<148, 128, 4 * c.boxes * box_bytes + 1024>>>(map, c.inner, c.boxes, box_bytes, iters, c.issuers);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double bytes = 148.0 * iters * c.boxes * box_bytes;
    printf("%-48s %8.1f GB/s  (%.1f B/clk/SM at 1.965 GHz)\n", c.name, bytes / ms / 1e6, bytes / ms / 1e6 / 148 / 1.965);
  }
  return 0;
}
