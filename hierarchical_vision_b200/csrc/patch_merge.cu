// PatchMerging's 2x2 strided gather (reference swinv2.py:484-491): x (B, H*W, C) -> (B, H/2*W/2, 4C)
// with channel block m of output token (b, i, j) taken from token (b, 2i + (m&1), 2j + (m>>1)).
// Forward and backward are the same permutation run in opposite directions, as pure 16-byte
// vector copies: every byte is read once and written once, writes fully coalesced.
#include "hv_common.cuh"

namespace hv {
namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

// FORWARD: dst = merged (B, H/2*W/2, 4C), src = x.   !FORWARD: dst = dx, src = d merged.
template <bool FORWARD>
__global__ void __launch_bounds__(kThreads) patch_merge_copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                                    int H, int W, int cv /* 16B vectors per C */,
                                                                    int64_t total_vecs) {
  const int H2 = H >> 1, W2 = W >> 1;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  int64_t base = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  for (; base < total_vecs; base += stride * kUnroll) {
    uint4 v[kUnroll];
    int64_t dsti[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t id = base + u * stride;
      dsti[u] = -1;
      if (id < total_vecs) {
        // id enumerates the merged tensor: ((b*H2 + i)*W2 + j)*4 + m, then the vector inside C
        const int c = (int)(id % cv);
        int64_t t = id / cv;
        const int m = (int)(t & 3);
        t >>= 2;
        const int j = (int)(t % W2);
        t /= W2;
        const int i = (int)(t % H2);
        const int b = (int)(t / H2);
        const int64_t xi = merge_src_token(H, W, b, i, j, m) * cv + c;
        if (FORWARD) { v[u] = src[xi]; dsti[u] = id; }
        else         { v[u] = src[id]; dsti[u] = xi; }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (dsti[u] >= 0) dst[dsti[u]] = v[u];
  }
}

int launch(bool forward, const void* src, void* dst, int B, int H, int W, int C, int dtype, cudaStream_t st) {
  if (B <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) HV_FAIL(HV_ERR_SHAPE, "patch_merge: x size (%d*%d) are not even.", H, W);
  if (dtype != HV_F32 && dtype != HV_BF16) HV_FAIL(HV_ERR_DTYPE, "patch_merge: dtype %d", dtype);
  const int esz = dtype == HV_F32 ? 4 : 2;
  if ((C * esz) % 16 != 0) HV_FAIL(HV_ERR_SHAPE, "patch_merge: C*elem_size=%d must be a multiple of 16", C * esz);
  if (!aligned16(src) || !aligned16(dst)) HV_FAIL(HV_ERR_ALIGN, "patch_merge: pointers must be 16-byte aligned");
  const int cv = C * esz / 16;
  const int64_t total = (int64_t)B * H * W * cv;
  int64_t blocks = (total + (int64_t)kThreads * kUnroll - 1) / ((int64_t)kThreads * kUnroll);
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (forward)
    patch_merge_copy_kernel<true><<<(int)blocks, kThreads, 0, st>>>((const uint4*)src, (uint4*)dst, H, W, cv, total);
  else
    patch_merge_copy_kernel<false><<<(int)blocks, kThreads, 0, st>>>((const uint4*)src, (uint4*)dst, H, W, cv, total);
  HV_LAUNCH_OK("patch_merge_copy_kernel");
  return HV_OK;
}

}  // namespace

int patch_merge_gather_fwd(const void* x, void* out, int B, int H, int W, int C, int dtype, cudaStream_t st) {
  return launch(true, x, out, B, H, W, C, dtype, st);
}
int patch_merge_gather_bwd(const void* dout, void* dx, int B, int H, int W, int C, int dtype, cudaStream_t st) {
  return launch(false, dout, dx, B, H, W, C, dtype, st);
}

}  // namespace hv
