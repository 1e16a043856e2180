// Continuous position bias table of SwinV2's WindowAttention (reference swinv2.py:141-145, 233-246):
//   table[r, h] = 16 * sigmoid( W2[h, :] . relu(W1 coords[r] + b1) ),   coords (M, 2), W1 (HID, 2), b1 (HID), W2 (heads, HID)
// with M = (2 ws - 1)^2 (225 for the 8x8 window) and HID = 512.  The reference runs this as two Linear layers, a ReLU,
// a gather and a sigmoid per block and per step (~5 tiny kernels forward, ~8 backward); here it is one kernel each way.
// The gather through relative_position_index happens inside the attention kernels (closed form), so the table is
// all that is needed.  Deterministic: no atomics, fixed summation order.
#include "hv_common.cuh"

namespace hv {
namespace {

constexpr int kHidMax = 512;
constexpr int kHeadsMax = 32;

// One CTA per table row r: HID threads compute the hidden activations, then warp w reduces head w, w + nwarps, ...
// MODE 0: out[r, h] = 16 sigmoid(z)        MODE 1: out[r, h] = dtable[r, h] * 16 sigmoid(z) (1 - sigmoid(z))  (= dz)
template <int MODE>
__global__ void __launch_bounds__(kHidMax) cpb_rows_kernel(const float* __restrict__ coords, const float* __restrict__ w1,
                                                          const float* __restrict__ b1, const float* __restrict__ w2,
                                                          const float* __restrict__ dtable, float* __restrict__ out,
                                                          int hid, int heads) {
  __shared__ float act[kHidMax];
  const int r = blockIdx.x, j = threadIdx.x;
  const float c0 = coords[2 * r], c1 = coords[2 * r + 1];
  if (j < hid) act[j] = fmaxf(fmaf(w1[2 * j], c0, fmaf(w1[2 * j + 1], c1, b1[j])), 0.f);
  __syncthreads();
  const int warp = j >> 5, lane = j & 31, nwarps = blockDim.x >> 5;
  for (int h = warp; h < heads; h += nwarps) {
    float s = 0.f;
    for (int k = lane; k < hid; k += 32) s = fmaf(w2[h * hid + k], act[k], s);
    s = warp_sum(s);
    if (lane == 0) {
      const float sg = 1.0f / (1.0f + __expf(-s));
      out[r * heads + h] = MODE == 0 ? 16.0f * sg : dtable[r * heads + h] * 16.0f * sg * (1.0f - sg);
    }
  }
}

// 32 hidden units per CTA, 8 row chunks per unit (thread = unit x chunk): every thread accumulates dW2[:, j],
// dW1[j, :], db1[j] over its rows in registers, the 8 partials are summed through shared memory in a fixed order.
// gridDim.y > 1 splits the table rows between CTAs (the 961-row table of a 16 x 16 window is a 120-step serial chain per
// thread otherwise, 80 us on 16 SMs): the CTAs then accumulate into the zeroed outputs with fp32 atomics.
constexpr int kChunks = 8;
__global__ void __launch_bounds__(32 * kChunks) cpb_bwd_hidden_kernel(const float* __restrict__ coords, const float* __restrict__ w1,
                                                                       const float* __restrict__ b1, const float* __restrict__ w2,
                                                                       const float* __restrict__ dz, float* __restrict__ dw1,
                                                                       float* __restrict__ db1, float* __restrict__ dw2, int M,
                                                                       int hid, int heads) {
  // one buffer, two lives: first the staged dz rows (every thread reads all of them: from global that was a serial chain
  // of ~700 dependent L2 loads per thread, 21 us for this tiny kernel), then the partial sums of the row chunks
  __shared__ float buf[kChunks * (kHeadsMax + 3) * 32];
  float (*part)[kHeadsMax + 3][32] = reinterpret_cast<float (*)[kHeadsMax + 3][32]>(buf);
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r_begin = blockIdx.y * rows_per, r_end = min(M, r_begin + rows_per);
  const bool staged = rows_per * heads <= kChunks * (kHeadsMax + 3) * 32;
  if (staged)
    for (int i = threadIdx.x; i < (r_end - r_begin) * heads; i += 32 * kChunks) buf[i] = __ldg(&dz[r_begin * heads + i]);
  __syncthreads();
  const float* dzs = staged ? buf - r_begin * heads : dz;
  const int jl = threadIdx.x & 31, rc = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + jl;
  const bool live = j < hid;
  const float wa = live ? w1[2 * j] : 0.f, wb = live ? w1[2 * j + 1] : 0.f, bb = live ? b1[j] : 0.f;
  float w2j[kHeadsMax], g2[kHeadsMax];
#pragma unroll
  for (int h = 0; h < kHeadsMax; ++h) {
    w2j[h] = (live && h < heads) ? w2[h * hid + j] : 0.f;
    g2[h] = 0.f;
  }
  float ga = 0.f, gb = 0.f, gbias = 0.f;
  for (int r = r_begin + rc; r < r_end; r += kChunks) {
    const float c0 = __ldg(&coords[2 * r]), c1 = __ldg(&coords[2 * r + 1]);
    const float pre = fmaf(wa, c0, fmaf(wb, c1, bb));
    const float a = fmaxf(pre, 0.f);
    float da = 0.f;
#pragma unroll
    for (int h = 0; h < kHeadsMax; ++h) {
      if (h < heads) {
        const float d = dzs[r * heads + h];  // same address for every lane of the warp: broadcast
        g2[h] = fmaf(d, a, g2[h]);
        da = fmaf(d, w2j[h], da);
      }
    }
    if (pre > 0.f) {
      ga = fmaf(da, c0, ga);
      gb = fmaf(da, c1, gb);
      gbias += da;
    }
  }
  __syncthreads();  // everybody is done with the staged dz
#pragma unroll
  for (int h = 0; h < kHeadsMax; ++h) part[rc][h][jl] = g2[h];
  part[rc][kHeadsMax][jl] = ga;
  part[rc][kHeadsMax + 1][jl] = gb;
  part[rc][kHeadsMax + 2][jl] = gbias;
  __syncthreads();
  // thread (rc, jl) finishes slots rc, rc + 8, ... of unit jl
  for (int slot = rc; slot < kHeadsMax + 3; slot += kChunks) {
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) sum += part[c][slot][jl];
    if (!live) continue;
    float* dst = nullptr;
    if (slot < kHeadsMax) { if (slot < heads) dst = &dw2[slot * hid + j]; }
    else if (slot == kHeadsMax) dst = &dw1[2 * j];
    else if (slot == kHeadsMax + 1) dst = &dw1[2 * j + 1];
    else dst = &db1[j];
    if (dst == nullptr) continue;
    if (gridDim.y == 1) *dst = sum;
    else atomicAdd(dst, sum);
  }
}

int check(int M, int hid, int heads) {
  if (M <= 0 || hid <= 0 || hid > kHidMax || hid % 32 != 0 || heads <= 0 || heads > kHeadsMax)
    HV_FAIL(HV_ERR_SHAPE, "cpb_bias: M=%d hidden=%d (<= %d, multiple of 32) heads=%d (<= %d)", M, hid, kHidMax, heads, kHeadsMax);
  return HV_OK;
}

}  // namespace

int cpb_bias_fwd(const float* coords, const float* w1, const float* b1, const float* w2, float* table, int M, int hid,
                 int heads, cudaStream_t st) {
  const int rc = check(M, hid, heads);
  if (rc) return rc;
  cpb_rows_kernel<0><<<M, hid, 0, st>>>(coords, w1, b1, w2, nullptr, table, hid, heads);
  HV_LAUNCH_OK("cpb_rows_kernel<0>");
  return HV_OK;
}

int cpb_bias_bwd(const float* coords, const float* w1, const float* b1, const float* w2, const float* dtable, float* dw1,
                 float* db1, float* dw2, float* workspace, int M, int hid, int heads, cudaStream_t st) {
  const int rc = check(M, hid, heads);
  if (rc) return rc;
  cpb_rows_kernel<1><<<M, hid, 0, st>>>(coords, w1, b1, w2, dtable, workspace, hid, heads);
  HV_LAUNCH_OK("cpb_rows_kernel<1>");
  const int splits = M >= 512 ? 8 : 1;  // the (2 * 16 - 1)^2 table of 16 x 16 windows
  if (splits > 1) {
    HV_CUDA_OK(cudaMemsetAsync(dw1, 0, sizeof(float) * 2 * hid, st));
    HV_CUDA_OK(cudaMemsetAsync(db1, 0, sizeof(float) * hid, st));
    HV_CUDA_OK(cudaMemsetAsync(dw2, 0, sizeof(float) * heads * hid, st));
  }
  cpb_bwd_hidden_kernel<<<dim3((hid + 31) / 32, splits), 32 * kChunks, 0, st>>>(coords, w1, b1, w2, workspace, dw1, db1, dw2, M, hid, heads);
  HV_LAUNCH_OK("cpb_bwd_hidden_kernel");
  return HV_OK;
}

}  // namespace hv
