// Probe (B200): how fast can one SM pull 36 KB tiles with (a) cp.async.bulk requests of various sizes issued
// by 1 lane / 32 lanes / several warps, (b) LDGSTS (cp.async 16 B).  Prints cycles per tile and GB/s.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// mode 0: bulk copies; `nwarps` warps issue, `lanes` lanes per warp active; each request `bytes`; tile = 36864 B
// Two tiles in flight (double buffer).  Each CTA streams `ntiles` tiles from its own region of `src`.
__global__ void __launch_bounds__(512, 1) probe_bulk(const uint8_t* src, size_t cta_stride, int ntiles, int bytes, int nwarps, int lanes,
                                                     long long* cycles_out, int* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bars[2];
  const int tile_bytes = 36864;
  const int nreq = tile_bytes / bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(&bars[0]);
  if (threadIdx.x == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const uint8_t* base = src + (size_t)blockIdx.x * cta_stride;
  const uint32_t sbase = smem_u32(smem);
  long long t0 = clock64();
  int acc = 0;
  for (int t = 0; t < ntiles + 1; ++t) {
    if (t < ntiles) {
      const int s = t & 1;
      if (threadIdx.x == 0) mbar_expect_tx(bar0 + 8 * s, tile_bytes);
      __syncthreads();
      if (warp < nwarps && lane < lanes) {
        const int issuer = warp * lanes + lane, nissuers = nwarps * lanes;
        for (int r = issuer; r < nreq; r += nissuers)
          bulk_g2s(sbase + s * tile_bytes + r * bytes, base + (size_t)t * tile_bytes + (size_t)r * bytes, bytes, bar0 + 8 * s);
      }
    }
    if (t > 0) {
      const int s = (t - 1) & 1;
      mbar_wait(bar0 + 8 * s, ((t - 1) >> 1) & 1);
      acc += smem[s * tile_bytes + threadIdx.x * 16];
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cycles_out[blockIdx.x] = t1 - t0; sink[blockIdx.x] = acc; }
}

// mode 1: LDGSTS 16 B per lane, `nwarps` warps issue, commit/wait groups, double buffered
__global__ void __launch_bounds__(512, 1) probe_ldgsts(const uint8_t* src, size_t cta_stride, int ntiles, int nwarps, long long* cycles_out, int* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tile_bytes = 36864;
  const int warp = threadIdx.x >> 5;
  const uint8_t* base = src + (size_t)blockIdx.x * cta_stride;
  const uint32_t sbase = smem_u32(smem);
  long long t0 = clock64();
  int acc = 0;
  for (int t = 0; t < ntiles + 1; ++t) {
    if (t < ntiles && warp < nwarps) {
      const int s = t & 1;
      for (int c = threadIdx.x; c < tile_bytes / 16; c += nwarps * 32) {
        // swizzled destination like the real kernel would use
        const int row = c / 36, ch = c % 36;
        const uint32_t dst = sbase + s * tile_bytes + row * 576 + ((ch ^ (row & 3)) % 36) * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(base + (size_t)t * tile_bytes + (size_t)c * 16) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (t > 0) {
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncthreads();
      acc += smem[((t - 1) & 1) * tile_bytes + threadIdx.x * 16];
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cycles_out[blockIdx.x] = t1 - t0; sink[blockIdx.x] = acc; }
}

int main() {
  const int ctas = 148, ntiles = 400;
  const size_t cta_stride = (size_t)ntiles * 36864;
  uint8_t* src; long long* cyc; int* sink;
  cudaMalloc(&src, cta_stride * ctas); cudaMemset(src, 1, cta_stride * ctas);
  cudaMalloc(&cyc, ctas * sizeof(long long)); cudaMalloc(&sink, ctas * sizeof(int));
  cudaFuncSetAttribute(probe_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 36864);
  cudaFuncSetAttribute(probe_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 36864);
  long long h[148];
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int sizes[] = {64, 192, 576, 1152, 4608, 36864};
  const int cfgs[][2] = {{1, 1}, {1, 32}, {4, 8}, {4, 32}, {12, 32}};
  for (int grid : {1, 148}) {
    for (int bytes : sizes)
      for (auto& c : cfgs) {
        if (36864 / bytes < c[0] * c[1] && !(c[0] == 1 && c[1] == 1)) continue;
        probe_bulk<<<grid, 512, 2 * 36864>>>(src, cta_stride, ntiles, bytes, c[0], c[1], cyc, sink);
        cudaEventRecord(e0);
        probe_bulk<<<grid, 512, 2 * 36864>>>(src, cta_stride, ntiles, bytes, c[0], c[1], cyc, sink);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("bulk grid=%d bytes=%d cfg=%dx%d ERROR %s\n", grid, bytes, c[0], c[1], cudaGetErrorString(err)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
        printf("bulk   grid=%3d req=%5dB x%4d issuers=%2dw x%2dl : %8.0f cyc/tile  %7.1f GB/s total  (%.3f ms)\n", grid, bytes, 36864 / bytes, c[0], c[1],
               avg / ntiles, (double)grid * ntiles * 36864 / ms / 1e6, ms);
      }
    for (int nw : {1, 4, 12, 16}) {
      probe_ldgsts<<<grid, 512, 2 * 36864>>>(src, cta_stride, ntiles, nw, cyc, sink);
      cudaEventRecord(e0);
      probe_ldgsts<<<grid, 512, 2 * 36864>>>(src, cta_stride, ntiles, nw, cyc, sink);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("ldgsts ERROR %s\n", cudaGetErrorString(err)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
      printf("ldgsts grid=%3d warps=%2d : %8.0f cyc/tile  %7.1f GB/s total  (%.3f ms)\n", grid, nw, avg / ntiles, (double)grid * ntiles * 36864 / ms / 1e6, ms);
    }
  }
  return 0;
}
