// Probe for the tcgen05 / TMA building blocks of the window-attention kernel (sm_100a):
//   1. a 4-D tiled tensor map over qkv (B, H, W, 3C) bf16 with SWIZZLE_64B; box = (32 ch, 8, 8, 1) = one window's
//      q / k / v tile of one head, 64 rows x 64 B;
//   2. S = [Q_a; Q_b] [K_a; K_b]^T by tcgen05.mma (M = 128, N = 128, K = 32, both operands K-major SW64 in smem);
//   3. tcgen05.ld of S, P = bf16(S * 1/64) written back with tcgen05.st, O = P V with A from TMEM and V as an
//      MN-major SW64 smem operand (M = 128, N = 32, K = 64), once per half.
// Prints max |error| of S and O against a host reference.  nvcc -arch=sm_100a tools/probes/umma_probe.cu -o umma_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

typedef __nv_bfloat16 bf16;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
    if (!done && ++spins > (1u << 16)) __trap();
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

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

// K-major / MN-major SWIZZLE_64B shared-memory operand descriptor: 64-byte rows, 8-row groups 512 B apart
__device__ __forceinline__ uint64_t sw64_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);   // start address
  d |= (uint64_t)1 << 16;                   // leading byte offset (unused for one swizzle atom), in 16-byte units
  d |= (uint64_t)(512 >> 4) << 32;          // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
  d |= (uint64_t)4 << 61;                   // SWIZZLE_64B
  return d;
}
// instruction descriptor, kind::f16: bf16 x bf16 -> f32
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int kTile = 4096;  // 64 rows x 64 B
constexpr int kC = 96, kH = 16, kW = 16, kB = 2;

__global__ void __launch_bounds__(192, 1) probe_kernel(const __grid_constant__ CUtensorMap map, float* out_s, float* out_o) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar_tma = smem_u32(&bars[0]), bar_s = smem_u32(&bars[1]), bar_p = smem_u32(&bars[2]), bar_o = smem_u32(&bars[3]);
  if (threadIdx.x == 0) {
    mbar_init(bar_tma, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // unit a: image 0, window (wh 0, ww 1), head 0 ; unit b: image 1, window (wh 1, ww 0), head 2
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(bar_tma, 6 * kTile);
    for (int part = 0; part < 3; ++part) {
      tma_load_4d(sb + (2 * part + 0) * kTile, &map, bar_tma, part * kC + 0 * 32, 8, 0, 0);
      tma_load_4d(sb + (2 * part + 1) * kTile, &map, bar_tma, part * kC + 2 * 32, 0, 8, 1);
    }
  }
  if (warp == 1) {
    mbar_wait(bar_tma, 0);
    tc_fence_after();
    if (lane == 0) {
      const uint32_t id_s = idesc_bf16(128, 128, 0, 0);
      for (int k = 0; k < 2; ++k)
        umma_ss(tmem + 0, sw64_desc(sb + 0 * kTile + 32 * k), sw64_desc(sb + 2 * kTile + 32 * k), id_s, k > 0);
      umma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_p, 0);
    tc_fence_after();
    if (lane == 0) {
      const uint32_t id_o = idesc_bf16(128, 32, 0, 1);
      for (int half = 0; half < 2; ++half)
        for (int ks = 0; ks < 4; ++ks)
          umma_ts(tmem + 160 + 32 * half, tmem + 128 + 8 * ks, sw64_desc(sb + (4 + half) * kTile + 1024 * ks), id_o, ks > 0);
      umma_commit(bar_o);
    }
    __syncwarp();
  }
  if (warp >= 2) {
    const int quad = warp & 3;            // TMEM lane quadrant this warp may touch
    const int row = quad * 32 + lane;     // accumulator row
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    mbar_wait(bar_s, 0);
    tc_fence_after();
    uint32_t r[32];
    uint32_t pk[32];
    const int cbase = row < 64 ? 0 : 64;  // own diagonal block
    for (int c = 0; c < 4; ++c) {
      TMEM_LD32(lane_addr + 32 * c, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) out_s[row * 128 + 32 * c + j] = __uint_as_float(r[j]);
      if (32 * c >= cbase && 32 * c < cbase + 64) {
        const int o = (32 * c - cbase) / 2;
        for (int j = 0; j < 16; ++j) {
          __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(r[2 * j]) * (1.f / 64), __uint_as_float(r[2 * j + 1]) * (1.f / 64));
          pk[o + j] = *reinterpret_cast<uint32_t*>(&h);
        }
      }
    }
    TMEM_ST32(lane_addr + 128, pk);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    mbar_arrive(bar_p);
    mbar_wait(bar_o, 0);
    tc_fence_after();
    TMEM_LD32(lane_addr + 160 + (row < 64 ? 0 : 32), r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out_o[row * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const size_t n = (size_t)kB * kH * kW * 3 * kC;
  std::vector<bf16> h(n);
  std::vector<float> hf(n);
  srand(1);
  for (size_t i = 0; i < n; ++i) {
    h[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
    hf[i] = __bfloat162float(h[i]);
  }
  bf16* d;
  CK(cudaMalloc(&d, n * 2));
  CK(cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice));
  float *ds, *d_o;
  CK(cudaMalloc(&ds, 128 * 128 * 4));
  CK(cudaMalloc(&d_o, 128 * 32 * 4));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap map;
  cuuint64_t dims[4] = {3 * kC, kW, kH, kB};
  cuuint64_t strides[3] = {3 * kC * 2, (cuuint64_t)kW * 3 * kC * 2, (cuuint64_t)kH * kW * 3 * kC * 2};
  cuuint32_t box[4] = {32, 8, 8, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * kTile + 1024));
  probe_kernel<<<1, 192, 6 * kTile + 1024>>>(map, ds, d_o);
  CK(cudaDeviceSynchronize());
  std::vector<float> S(128 * 128), O(128 * 32);
  CK(cudaMemcpy(S.data(), ds, S.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(O.data(), d_o, O.size() * 4, cudaMemcpyDeviceToHost));
  // reference
  auto tok = [&](int unit, int i) {  // unit a: b0 (wh0, ww1); unit b: b1 (wh1, ww0)
    const int ih = i / 8, iw = i % 8;
    const int b = unit, r0 = unit == 0 ? 0 : 8, c0 = unit == 0 ? 8 : 0;
    return ((size_t)(b * kH + r0 + ih) * kW + c0 + iw) * 3 * kC;
  };
  auto el = [&](int unit, int part, int i, int ch) { return hf[tok(unit, i) + part * kC + (unit == 0 ? 0 : 2) * 32 + ch]; };
  double es = 0, eo = 0;
  for (int row = 0; row < 128; ++row) {
    const int u = row / 64, i = row % 64;
    std::vector<float> p(64);
    for (int col = 0; col < 128; ++col) {
      const int uc = col / 64, j = col % 64;
      double acc = 0;
      for (int ch = 0; ch < 32; ++ch) acc += (double)el(u, 0, i, ch) * el(uc, 1, j, ch);
      es = fmax(es, fabs(acc - S[row * 128 + col]));
      if (uc == u) p[j] = __bfloat162float(__float2bfloat16((float)(S[row * 128 + col] * (1.f / 64))));
    }
    for (int ch = 0; ch < 32; ++ch) {
      double acc = 0;
      for (int j = 0; j < 64; ++j) acc += (double)p[j] * el(u, 2, j, ch);
      eo = fmax(eo, fabs(acc - O[row * 32 + ch]));
    }
  }
  printf("umma_probe: max|S err| = %.3e   max|O err| = %.3e   S[0][0]=%f S[127][127]=%f O[0][0]=%f\n", es, eo, S[0], S[128 * 128 - 1], O[0]);
  printf(es < 1e-3 && eo < 1e-3 ? "PROBE OK\n" : "PROBE MISMATCH\n");
  return 0;
}
