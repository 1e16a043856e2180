// Where does tcgen05.mma with M = 64 (cta_group::1) put its accumulator rows in tensor memory?
// A = [64 x 16] bf16 with A[r][0] = r + 1, A[r][1] = 1; B = [64 x 16] with (test 0) B[n][0] = 1 or (test 1) B[n][1] = n + 1,
// so D[r][n] = r + 1 (row map) or n + 1 (column map).  TMEM is pre-filled with -1; all 128 lanes x 64 columns are dumped.
// nvcc -gencode arch=compute_100a,code=sm_100a tools/probes/umma_m64_probe.cu -o tools/probes/umma_m64_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef __nv_bfloat16 bf16;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sw64_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
#define TMEM_LD32(taddr, r)                                                                                         \
  asm volatile(                                                                                                     \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,"   \
      "%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),      \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),     \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                   \
      : "r"(taddr))
#define TMEM_ST32(taddr, r)                                                                                         \
  asm volatile(                                                                                                     \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,"    \
      "%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"                                               \
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),          \
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),   \
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), \
        "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                      \
      : "memory")

__global__ void __launch_bounds__(128, 1) probe(int test, int lane_off, float* out) {
  extern __shared__ __align__(1024) unsigned char smem[];  // A tile at 0, B tile at 4096 (64 rows x 64 B, SWIZZLE_64B)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2048; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int r = threadIdx.x;
    bf16* arow = reinterpret_cast<bf16*>(smem + r * 64 + ((0 ^ ((r >> 1) & 3)) << 4));         // chunk 0 of row r
    bf16* brow = reinterpret_cast<bf16*>(smem + 4096 + r * 64 + ((0 ^ ((r >> 1) & 3)) << 4));
    arow[0] = __float2bfloat16((float)(r + 1));
    arow[1] = __float2bfloat16(1.0f);
    if (test == 0) brow[0] = __float2bfloat16(1.0f); else brow[1] = __float2bfloat16((float)(r + 1));
  }
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t v[32];
  for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(-1.0f);
  TMEM_ST32(tl, v);
  TMEM_ST32(tl + 32, v);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t d = tmem + ((uint32_t)lane_off << 16);
    const uint64_t a = sw64_desc(smem_u32(smem)), b = sw64_desc(smem_u32(smem + 4096));
    const uint32_t id = idesc_bf16(64, 64);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int h = 0; h < 2; ++h) {
    TMEM_LD32(tl + 32 * h, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int e = 0; e < 32; ++e) out[(warp * 32 + lane) * 64 + 32 * h + e] = __uint_as_float(v[e]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64));
}

int main(int argc, char** argv) {
  const int lane_off = argc > 1 ? atoi(argv[1]) : 0;  // lane field of the accumulator address (0, 16, 32, 64 ...)
  float* d;
  CK(cudaMalloc(&d, 128 * 64 * sizeof(float)));
  static float h[2][128 * 64];
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 + 1024));
  for (int t = 0; t < 2; ++t) {
    probe<<<1, 128, 8192 + 1024>>>(t, lane_off, d);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h[t], d, sizeof(h[t]), cudaMemcpyDeviceToHost));
  }
  printf("# M = 64, N = 64, accumulator lane offset %d: per TMEM lane the accumulator row it holds (from D = r + 1), whether all\n", lane_off);
  printf("# 64 columns of the lane carry that row, and the column held at TMEM columns 0, 1, 31, 32, 63 (from D = n + 1)\n");
  for (int L = 0; L < 128; ++L) {
    const float r0 = h[0][L * 64];
    bool same = true;
    for (int c = 1; c < 64; ++c) same = same && h[0][L * 64 + c] == r0;
    printf("lane %3d: row %4.0f %s   cols %3.0f %3.0f %3.0f %3.0f %3.0f\n", L, r0 - 1, same ? "uniform" : "MIXED", h[1][L * 64] - 1,
           h[1][L * 64 + 1] - 1, h[1][L * 64 + 31] - 1, h[1][L * 64 + 32] - 1, h[1][L * 64 + 63] - 1);
  }
  return 0;
}
