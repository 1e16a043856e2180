// Fused cross entropy with label smoothing, forward and gradient in one launch: the tail of the training step
// (reference models.py:121-152 `loss`; hierarchy.py:65-94 MultitaskCrossEntropy = dot(coeffs, CE per tier);
// algorithmic.py:88-119, 160-164 label smoothing: target = onehot * (1 - a) + a / classes).
//
//   loss_row[r]   = lse_r - (1 - a) x[r, t_r] - (a / n) sum_c x[r, c]
//   dlogits[r, c] = scale * (softmax(x_r)[c] - (1 - a) [c == t_r] - a / n)          (scale = tier coefficient / rows)
//
// One CTA per row: one pass over the row for (max, sum exp, sum x) with an online merge, one more (L2-resident) pass to
// write the gradient.  Replaces, per tier, the cast + log_softmax + nll_loss + their backward kernels (~6 launches and
// four passes over the (rows, classes) logits); the result is what the host side multiplies by the upstream scalar.
#include "hv_common.cuh"

namespace hv {
namespace {

constexpr int kThreads = 256;

struct Acc { float m, s, x; };  // running max, sum exp(x - m), sum x
__device__ __forceinline__ Acc merge(const Acc& a, const Acc& b) {
  Acc r;
  r.m = fmaxf(a.m, b.m);
  r.s = a.s * __expf(a.m - r.m) + b.s * __expf(b.m - r.m);
  r.x = a.x + b.x;
  return r;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) ce_fwd_grad_kernel(const T* __restrict__ logits, const int64_t* __restrict__ target,
                                                              float* __restrict__ loss_rows, T* __restrict__ dlogits,
                                                              int classes, float smoothing, float scale) {
  __shared__ Acc s_acc[kThreads / 32];
  __shared__ Acc s_row;
  const int64_t r = blockIdx.x;
  const T* x = logits + r * classes;
  Acc a = {-3.0e38f, 0.f, 0.f};
  for (int c = threadIdx.x; c < classes; c += kThreads) {
    const float v = to_f32(x[c]);
    const float m = fmaxf(a.m, v);
    a.s = a.s * __expf(a.m - m) + __expf(v - m);
    a.m = m;
    a.x += v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Acc b;
    b.m = __shfl_xor_sync(0xffffffffu, a.m, o);
    b.s = __shfl_xor_sync(0xffffffffu, a.s, o);
    b.x = __shfl_xor_sync(0xffffffffu, a.x, o);
    a = merge(a, b);
  }
  if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    Acc t = s_acc[0];
    for (int w = 1; w < kThreads / 32; ++w) t = merge(t, s_acc[w]);
    s_row = t;
  }
  __syncthreads();
  const Acc t = s_row;
  const int64_t tgt = target[r];
  const float lse = t.m + __logf(t.s);
  const float un = smoothing / (float)classes;
  if (threadIdx.x == 0) {
    const float xt = (tgt >= 0 && tgt < classes) ? to_f32(x[tgt]) : 0.f;
    loss_rows[r] = lse - (1.0f - smoothing) * xt - un * t.x;
  }
  T* g = dlogits + r * classes;
  for (int c = threadIdx.x; c < classes; c += kThreads) {
    const float p = __expf(to_f32(x[c]) - lse);
    g[c] = from_f32<T>(scale * (p - (c == tgt ? 1.0f - smoothing : 0.f) - un));
  }
}

}  // namespace

int ce_fwd_grad(const void* logits, const int64_t* target, float* loss_rows, void* dlogits, int64_t rows, int classes,
                float smoothing, float scale, int dtype, cudaStream_t st) {
  if (rows <= 0 || classes <= 0) HV_FAIL(HV_ERR_SHAPE, "cross_entropy: rows=%lld classes=%d", (long long)rows, classes);
  if (rows > 2147483647LL) HV_FAIL(HV_ERR_SHAPE, "cross_entropy: too many rows");
  if (!(smoothing >= 0.f && smoothing < 1.f)) HV_FAIL(HV_ERR_SHAPE, "cross_entropy: label smoothing %f not in [0, 1)", smoothing);
  if (dtype == HV_F32)
    ce_fwd_grad_kernel<float><<<(unsigned)rows, kThreads, 0, st>>>((const float*)logits, target, loss_rows, (float*)dlogits, classes,
                                                                   smoothing, scale);
  else if (dtype == HV_BF16)
    ce_fwd_grad_kernel<bf16><<<(unsigned)rows, kThreads, 0, st>>>((const bf16*)logits, target, loss_rows, (bf16*)dlogits, classes,
                                                                  smoothing, scale);
  else
    HV_FAIL(HV_ERR_DTYPE, "cross_entropy: dtype %d", dtype);
  HV_LAUNCH_OK("ce_fwd_grad_kernel");
  return HV_OK;
}

}  // namespace hv
