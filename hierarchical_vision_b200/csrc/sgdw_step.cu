// One-pass DecoupledSGDW step over every parameter of the model (reference optim.py:16-44, Composer's DecoupledSGDW rule as
// restated in train.FlatSGD): for each element, with the gradient read from the flat fp32 gradient buffer and scaled by the
// gradient-clipping coefficient on the fly,
//     g' = g * coef;  buf = momentum * buf + g';  p = p * (1 - lr * wd / lr0) - lr * buf
// -- the same fp32 operations in the same order as the multi-tensor (foreach) path, so the results are bit-identical, but
// p, buf and g make ONE trip through HBM (5 x 4 bytes per parameter) instead of the ~14 of clip-scale + five foreach passes.
// Parameters live in separate allocations: a table of (p, buf, offset into the flat gradients, numel, wd / lr0) and a list of
// 4096-element chunks (tensor, offset) drive a grid-stride loop.  lr and coef are device scalars (CUDA-graph friendly).
#include "hv_common.cuh"

namespace hv {
namespace {

struct SgdwTensor {
  float* p;
  float* buf;
  long long goff;
  int n;
  float wd_scale;  // weight_decay / initial_lr (0: no decay)
};
static_assert(sizeof(SgdwTensor) == 32, "table layout is shared with train.py");

constexpr int kChunk = 4096;
constexpr int kThreads = 256;

__device__ __forceinline__ float upd_one(float g, float& b, float p, float c, float mom, float nlr, float decay) {
  const float gs = __fmul_rn(g, c);
  b = __fadd_rn(__fmul_rn(b, mom), gs);
  return __fadd_rn(__fmul_rn(p, decay), __fmul_rn(b, nlr));
}

__global__ void __launch_bounds__(kThreads)
sgdw_step_kernel(const SgdwTensor* __restrict__ tens, const int2* __restrict__ chunks, int nchunks, const float* __restrict__ flat,
                 const float* __restrict__ lr_p, const float* __restrict__ coef_p, float momentum) {
  const float lr = __ldg(lr_p), c = coef_p ? __ldg(coef_p) : 1.0f, nlr = -lr;
  for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const int2 cd = __ldg(&chunks[ch]);
    const SgdwTensor t = tens[cd.x];
    const int off = cd.y, n = min(kChunk, t.n - off);
    const float decay = __fsub_rn(1.0f, __fmul_rn(lr, t.wd_scale));
    float* p = t.p + off;
    float* b = t.buf + off;
    const float* g = flat + t.goff + off;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(g)) & 15u) == 0;
    if (vec) {
      const int n4 = n >> 2;
      for (int i = threadIdx.x; i < n4; i += kThreads) {
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
        float4 bv = reinterpret_cast<float4*>(b)[i], pv = reinterpret_cast<float4*>(p)[i];
        pv.x = upd_one(gv.x, bv.x, pv.x, c, momentum, nlr, decay);
        pv.y = upd_one(gv.y, bv.y, pv.y, c, momentum, nlr, decay);
        pv.z = upd_one(gv.z, bv.z, pv.z, c, momentum, nlr, decay);
        pv.w = upd_one(gv.w, bv.w, pv.w, c, momentum, nlr, decay);
        reinterpret_cast<float4*>(b)[i] = bv;
        reinterpret_cast<float4*>(p)[i] = pv;
      }
      for (int i = (n4 << 2) + threadIdx.x; i < n; i += kThreads) {
        float bb = b[i];
        p[i] = upd_one(g[i], bb, p[i], c, momentum, nlr, decay);
        b[i] = bb;
      }
    } else {
      for (int i = threadIdx.x; i < n; i += kThreads) {
        float bb = b[i];
        p[i] = upd_one(g[i], bb, p[i], c, momentum, nlr, decay);
        b[i] = bb;
      }
    }
  }
}

}  // namespace

int sgdw_step(const void* table, const void* chunks, int nchunks, const float* flat, const float* lr, const float* coef,
              float momentum, cudaStream_t st) {
  if (nchunks <= 0) return HV_OK;
  int grid = 8 * num_sms();
  if (grid > nchunks) grid = nchunks;
  sgdw_step_kernel<<<grid, kThreads, 0, st>>>(static_cast<const SgdwTensor*>(table), static_cast<const int2*>(chunks), nchunks, flat,
                                              lr, coef, momentum);
  HV_LAUNCH_OK("sgdw_step_kernel");
  return HV_OK;
}

}  // namespace hv
