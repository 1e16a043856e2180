// Mlp activation of the SwinV2 block (reference swinv2.py:60-63): a = GELU_erf(h + b1), where h = x W1^T is the
// fc1 GEMM output computed WITHOUT its bias.  Folding the bias add into the activation kernel makes the
// gradient of fc1.bias a by-product of the activation backward (column sums of d h accumulated in registers)
// instead of a separate full-tensor reduction over the largest activation of the block.
//
// Pure streaming kernels (HBM-bound): a half-warp owns a strip of 16 consecutive 16-byte vectors of one row,
// so a warp instruction covers two rows x 256 contiguous bytes; every lane always sees the same columns, which
// is what lets the bias (forward) and the d-bias partial sums (backward) live in registers.
#include "hv_common.cuh"

namespace hv {
namespace {

constexpr int kThreads = 256;
constexpr int kRowsPerIter = 4;  // per half-warp: vectors in flight per lane

// erf via Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7): one reciprocal and one exp2; the same exp(-x^2/2)
// is the Gaussian density the derivative needs.
struct GeluParts { float cdf, pdf_x; };  // Phi(x), x * phi(x)
__device__ __forceinline__ GeluParts gelu_parts(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = exp2f(-z * z * 1.4426950408889634f);  // exp(-x^2/2)
  const float erf_abs = fmaf(-poly, e, 1.0f);
  GeluParts p;
  p.cdf = 0.5f * (1.0f + copysignf(erf_abs, x));
  p.pdf_x = x * e * 0.3989422804014327f;
  return p;
}

// bf16 activations: Phi(x) ~ 0.5 (1 + tanh(x (c1 + c3 x^2 + c5 x^4))), coefficients fitted to the exact erf form
// (max |GELU error| 2.8e-5, derivative 1.2e-4, plus tanh.approx's 2^-11): an order of magnitude below the bf16
// rounding of the result (2^-9), at less than half the issue slots of the erf evaluation -- these kernels are
// issue-bound, not HBM-bound, with the erf form.  fp32 activations keep the erf form above.
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ GeluParts gelu_parts_bf16(float x) {
  constexpr float c1 = 0.7974857091903687f, c3 = 0.03703207150101662f, c5 = -0.000356393022229895f;
  // c5 < 0: unclamped, the polynomial turns around at |x| ~ 8.3 and changes sign at ~11 (GELU(12) would come out as 0).
  // With x^2 capped at 64 the tanh argument is monotone (1.7 x beyond |x| = 8) and tanh saturates to +-1, sech^2 to 0.
  const float x2 = fminf(x * x, 64.0f);
  const float t = tanh_fast(x * fmaf(x2, fmaf(x2, c5, c3), c1));
  GeluParts p;
  p.cdf = fmaf(0.5f, t, 0.5f);
  const float up = fmaf(x2, fmaf(x2, 5.0f * c5, 3.0f * c3), c1);
  p.pdf_x = (0.5f * x) * up * fmaf(-t, t, 1.0f);  // x * d/dx of the fitted Phi
  return p;
}
// Two elements per instruction with sm_100's packed fp32 arithmetic (FMUL2 / FFMA2 / FADD2): the bf16 kernels are bound
// by issue slots, and everything around the tanh is fp32 multiply-add work on independent neighbours.
struct GeluParts2 { float2 cdf, pdf_x; };
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
template <bool DERIV>
__device__ __forceinline__ GeluParts2 gelu_parts_bf16x2(float2 x) {
  constexpr float c1 = 0.7974857091903687f, c3 = 0.03703207150101662f, c5 = -0.000356393022229895f;
  float2 x2 = __fmul2_rn(x, x);
  x2 = make_float2(fminf(x2.x, 64.0f), fminf(x2.y, 64.0f));  // see gelu_parts_bf16
  const float2 u = __fmul2_rn(x, __ffma2_rn(x2, __ffma2_rn(x2, splat2(c5), splat2(c3)), splat2(c1)));
  const float2 t = make_float2(tanh_fast(u.x), tanh_fast(u.y));
  GeluParts2 p;
  p.cdf = __ffma2_rn(splat2(0.5f), t, splat2(0.5f));
  if (DERIV) {
    const float2 up = __ffma2_rn(x2, __ffma2_rn(x2, splat2(5.0f * c5), splat2(3.0f * c3)), splat2(c1));
    const float2 sech2 = __ffma2_rn(make_float2(-t.x, -t.y), t, splat2(1.0f));
    p.pdf_x = __fmul2_rn(__fmul2_rn(__fmul2_rn(splat2(0.5f), x), up), sech2);  // x * d/dx of the fitted Phi
  } else {
    p.pdf_x = splat2(0.f);
  }
  return p;
}
template <typename T> __device__ __forceinline__ GeluParts gelu_eval(float x);
template <> __device__ __forceinline__ GeluParts gelu_eval<float>(float x) { return gelu_parts(x); }
template <> __device__ __forceinline__ GeluParts gelu_eval<bf16>(float x) { return gelu_parts_bf16(x); }

template <typename T> struct V16;  // one packed 16-byte vector
template <> struct V16<float> {
  static constexpr int n = 4;
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
  __device__ __forceinline__ void get(float* f) const { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
  __device__ __forceinline__ void set(const float* f) { v = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct V16<bf16> {
  static constexpr int n = 8;
  uint4 v;
  __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = v; }
  __device__ __forceinline__ void get(float* f) const {
    f[0] = bf16lo_to_f32(v.x); f[1] = bf16hi_to_f32(v.x); f[2] = bf16lo_to_f32(v.y); f[3] = bf16hi_to_f32(v.y);
    f[4] = bf16lo_to_f32(v.z); f[5] = bf16hi_to_f32(v.z); f[6] = bf16lo_to_f32(v.w); f[7] = bf16hi_to_f32(v.w);
  }
  __device__ __forceinline__ void set(const float* f) {
    v = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
};

// strips = vectors_per_row / 16; warp w handles strip (w % strips) of rows rgroup*2R + ..., rgroup = w / strips
template <typename T, bool BACKWARD>
__global__ void __launch_bounds__(kThreads, BACKWARD ? 3 : 4)  // backward: 85 registers (no spills) at 24 warps per SM
bias_gelu_kernel(const T* __restrict__ h, const T* __restrict__ dout, const float* __restrict__ bias, T* __restrict__ out,
                 float* __restrict__ partials, int64_t rows, int cols, int strips, int rgroups) {
  constexpr int VE = V16<T>::n;
  constexpr int R = kRowsPerIter;
  const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
  const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int strip = (int)(warp % strips);
  const int64_t rgroup = warp / strips;
  if (rgroup >= rgroups) return;
  const int col = (strip * 16 + hl) * VE;
  float b[VE], acc[VE];
#pragma unroll
  for (int e = 0; e < VE; ++e) {
    b[e] = bias[col + e];
    acc[e] = 0.f;
  }
  for (int64_t r0 = rgroup * (2 * R); r0 < rows; r0 += (int64_t)rgroups * (2 * R)) {
    V16<T> hv_[R], gv_[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const int64_t r = r0 + 2 * rr + half;
      if (r < rows) {
        hv_[rr].load(h + r * cols + col);
        if (BACKWARD) gv_[rr].load(dout + r * cols + col);
      }
    }
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const int64_t r = r0 + 2 * rr + half;
      if (r < rows) {
        float x[VE], g[VE], o[VE];
        hv_[rr].get(x);
        if (BACKWARD) gv_[rr].get(g);
        if (sizeof(T) == 2) {
#pragma unroll
          for (int e = 0; e < VE; e += 2) {
            const float2 xe = __fadd2_rn(make_float2(x[e], x[e + 1]), make_float2(b[e], b[e + 1]));
            const GeluParts2 p = gelu_parts_bf16x2<BACKWARD>(xe);
            float2 oe;
            if (BACKWARD) {
              oe = __fmul2_rn(make_float2(g[e], g[e + 1]), __fadd2_rn(p.cdf, p.pdf_x));
              const float2 a2 = __fadd2_rn(make_float2(acc[e], acc[e + 1]), oe);
              acc[e] = a2.x; acc[e + 1] = a2.y;
            } else {
              oe = __fmul2_rn(xe, p.cdf);
            }
            o[e] = oe.x; o[e + 1] = oe.y;
          }
        } else {
#pragma unroll
          for (int e = 0; e < VE; ++e) {
            const float xe = x[e] + b[e];
            const GeluParts p = gelu_eval<T>(xe);
            if (BACKWARD) {
              o[e] = g[e] * (p.cdf + p.pdf_x);
              acc[e] += o[e];
            } else {
              o[e] = xe * p.cdf;
            }
          }
        }
        V16<T> ov;
        ov.set(o);
        ov.store(out + r * cols + col);
      }
    }
  }
  if (BACKWARD) {
#pragma unroll
    for (int e = 0; e < VE; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
    if (half == 0) {
      float* dst = partials + rgroup * cols + col;
#pragma unroll
      for (int e = 0; e < VE; ++e) dst[e] = acc[e];
    }
  }
}

// out[c] = sum over `nrows` partial rows: block = 32 columns x 8 row-groups
__global__ void __launch_bounds__(256) colsum_rows_kernel(const float* __restrict__ partials, int nrows, int ncols,
                                                          float* __restrict__ out) {
  __shared__ float sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < ncols)
    for (int r = ty; r < nrows; r += 8) s += partials[(int64_t)r * ncols + c];
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < ncols) {
#pragma unroll
    for (int k = 1; k < 8; ++k) s += sm[k][tx];
    out[c] = s;
  }
}

struct Plan { int strips, rgroups, blocks; };
bool make_plan(int64_t rows, int cols, int ve, bool backward, Plan& p) {
  if (cols % (16 * ve) != 0) return false;
  p.strips = cols / (16 * ve);
  const int64_t max_warps = (int64_t)num_sms() * (backward ? 3 : 4) * (kThreads / 32);  // resident CTAs per SM
  int64_t rg = max_warps / p.strips;
  const int64_t need = (rows + 2 * kRowsPerIter - 1) / (2 * kRowsPerIter);
  if (rg > need) rg = need;
  if (rg < 1) rg = 1;
  p.rgroups = (int)rg;
  const int64_t warps = rg * p.strips;
  p.blocks = (int)((warps + (kThreads / 32) - 1) / (kThreads / 32));
  return true;
}

}  // namespace

size_t bias_gelu_bwd_workspace_bytes(int64_t rows, int cols) {
  (void)rows;
  // rgroups <= resident warps / strips, so rgroups * cols <= resident warps * 16 * VE floats (VE <= 8)
  return (size_t)num_sms() * 4 * (kThreads / 32) * 16 * 8 * sizeof(float);
}

int bias_gelu_fwd(const void* h, const float* bias, void* out, int64_t rows, int cols, int dtype, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) HV_FAIL(HV_ERR_SHAPE, "bias_gelu: rows=%lld cols=%d", (long long)rows, cols);
  if (dtype != HV_F32 && dtype != HV_BF16) HV_FAIL(HV_ERR_DTYPE, "bias_gelu: dtype %d", dtype);
  if (!aligned16(h) || !aligned16(out)) HV_FAIL(HV_ERR_ALIGN, "bias_gelu: pointers must be 16-byte aligned");
  Plan p;
  if (!make_plan(rows, cols, dtype == HV_F32 ? 4 : 8, false, p))
    HV_FAIL(HV_ERR_SHAPE, "bias_gelu: cols=%d must be a multiple of %d", cols, dtype == HV_F32 ? 64 : 128);
  if (dtype == HV_F32)
    bias_gelu_kernel<float, false><<<p.blocks, kThreads, 0, st>>>((const float*)h, nullptr, bias, (float*)out, nullptr, rows, cols, p.strips, p.rgroups);
  else
    bias_gelu_kernel<bf16, false><<<p.blocks, kThreads, 0, st>>>((const bf16*)h, nullptr, bias, (bf16*)out, nullptr, rows, cols, p.strips, p.rgroups);
  HV_LAUNCH_OK("bias_gelu_kernel<fwd>");
  return HV_OK;
}

int bias_gelu_bwd(const void* dout, const void* h, const float* bias, void* dh, float* dbias, void* workspace,
                  size_t workspace_bytes, int64_t rows, int cols, int dtype, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) HV_FAIL(HV_ERR_SHAPE, "bias_gelu_bwd: rows=%lld cols=%d", (long long)rows, cols);
  if (dtype != HV_F32 && dtype != HV_BF16) HV_FAIL(HV_ERR_DTYPE, "bias_gelu_bwd: dtype %d", dtype);
  if (!aligned16(h) || !aligned16(dout) || !aligned16(dh)) HV_FAIL(HV_ERR_ALIGN, "bias_gelu_bwd: pointers must be 16-byte aligned");
  Plan p;
  if (!make_plan(rows, cols, dtype == HV_F32 ? 4 : 8, true, p))
    HV_FAIL(HV_ERR_SHAPE, "bias_gelu_bwd: cols=%d must be a multiple of %d", cols, dtype == HV_F32 ? 64 : 128);
  if (workspace == nullptr || workspace_bytes < (size_t)p.rgroups * cols * sizeof(float))
    HV_FAIL(HV_ERR_WORKSPACE, "bias_gelu_bwd: workspace of %zu bytes required", bias_gelu_bwd_workspace_bytes(rows, cols));
  float* partials = static_cast<float*>(workspace);
  if (dtype == HV_F32)
    bias_gelu_kernel<float, true><<<p.blocks, kThreads, 0, st>>>((const float*)h, (const float*)dout, bias, (float*)dh, partials, rows, cols, p.strips, p.rgroups);
  else
    bias_gelu_kernel<bf16, true><<<p.blocks, kThreads, 0, st>>>((const bf16*)h, (const bf16*)dout, bias, (bf16*)dh, partials, rows, cols, p.strips, p.rgroups);
  HV_LAUNCH_OK("bias_gelu_kernel<bwd>");
  colsum_rows_kernel<<<(cols + 31) / 32, 256, 0, st>>>(partials, p.rgroups, cols, dbias);
  HV_LAUNCH_OK("colsum_rows_kernel");
  return HV_OK;
}

}  // namespace hv
