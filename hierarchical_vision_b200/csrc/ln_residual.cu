// Res-post-norm: out = shortcut + keep_scale[sample] * LayerNorm(y + bias)   (reference swinv2.py:431, 434;
// `bias` is the bias of the Linear that produced y -- attn.proj.bias / mlp.fc2.bias, swinv2.py:262, 64 -- folded
// in here so that its gradient falls out of this kernel's backward instead of a separate column reduction)
// and, with shortcut == nullptr, the plain LayerNorm of PatchMerging / PatchEmbed (swinv2.py:494, 656).
//
// HBM-bound streaming kernels.  A row (token) is owned by a group of GS lanes; every lane moves K 16-byte
// vectors of y per row and R rows per loop iteration (all loads issued before the first use, so each lane
// keeps R*K*(2..3) independent 16-byte requests in flight); statistics are reduced with warp shuffles;
// gamma/beta/bias live in registers for the lifetime of the thread.  Backward keeps its d-gamma / d-beta /
// d-bias partial sums in registers across all the rows a thread visits, folds them once per CTA and a second
// tiny kernel sums the per-CTA rows (deterministic two-stage reduction; no global atomics).
#include "hv_common.cuh"

namespace hv {
namespace {

constexpr int kThreads = 256;
constexpr int kFinalizeRows = 8;  // row-groups per finalize block

// Packed fp32 arithmetic of sm_100 (FADD2 / FMUL2 / FFMA2: two elements per instruction).  With bf16 activations these
// kernels move 6 bytes per element and are bound by issue slots, not by HBM; neighbouring elements of a vector are
// independent, so all the normalisation arithmetic pairs up.
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// One 16-byte global vector kept PACKED in registers (4 regs) until it is needed as fp32: keeps the
// register footprint of the R*K vectors a lane has in flight small enough for 3-4 CTAs per SM.
template <typename T> struct Pk;
template <> struct Pk<float> {
  static constexpr int n = 4;
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
  __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void get(float* f) const { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
  __device__ __forceinline__ void set(const float* f) { v = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Pk<bf16> {
  static constexpr int n = 8;
  uint4 v;
  __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = v; }
  __device__ __forceinline__ void zero() { v = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void get(float* f) const {
    f[0] = bf16lo_to_f32(v.x); f[1] = bf16hi_to_f32(v.x); f[2] = bf16lo_to_f32(v.y); f[3] = bf16hi_to_f32(v.y);
    f[4] = bf16lo_to_f32(v.z); f[5] = bf16hi_to_f32(v.z); f[6] = bf16lo_to_f32(v.w); f[7] = bf16hi_to_f32(v.w);
  }
  __device__ __forceinline__ void set(const float* f) {
    v = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
};
// VE consecutive elements of T (VE = elements of the *y* vector; T may be wider than y's type)
template <typename T, int VE> struct PkRow {
  static constexpr int kVecs = VE / Pk<T>::n;
  Pk<T> p[kVecs];
  __device__ __forceinline__ void load(const T* q) {
#pragma unroll
    for (int i = 0; i < kVecs; ++i) p[i].load(q + i * Pk<T>::n);
  }
  __device__ __forceinline__ void store(T* q) const {
#pragma unroll
    for (int i = 0; i < kVecs; ++i) p[i].store(q + i * Pk<T>::n);
  }
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < kVecs; ++i) p[i].zero();
  }
  __device__ __forceinline__ void get(float (&f)[VE]) const {
#pragma unroll
    for (int i = 0; i < kVecs; ++i) p[i].get(f + i * Pk<T>::n);
  }
  __device__ __forceinline__ void set(const float (&f)[VE]) {
#pragma unroll
    for (int i = 0; i < kVecs; ++i) p[i].set(f + i * Pk<T>::n);
  }
};

template <typename TY, typename TR, int GS, int K, int R, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
ln_residual_fwd_kernel(const TY* __restrict__ y, const TR* __restrict__ shortcut, const float* __restrict__ gamma,
                       const float* __restrict__ beta, const float* __restrict__ bias, const float* __restrict__ keep_scale,
                       TR* __restrict__ out, float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows, int C,
                       int64_t rows_per_sample, float eps) {
  constexpr int VE = 16 / sizeof(TY);
  constexpr int RPW = 32 / GS;  // rows per warp per sub-iteration
  const int lane = threadIdx.x & 31, gl = lane % GS, gi = lane / GS;
  const int vpr = C / VE;  // vectors per row
  float gam[K][VE], bet[K][VE], bia[K][VE];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int v = gl + k * GS;
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      gam[k][e] = (v < vpr) ? gamma[v * VE + e] : 0.f;
      bet[k][e] = (v < vpr) ? beta[v * VE + e] : 0.f;
      bia[k][e] = (v < vpr && bias != nullptr) ? bias[v * VE + e] : 0.f;
    }
  }
  const int64_t warp_global = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)gridDim.x * (kThreads / 32);
  const float inv_c = 1.0f / (float)C;
  for (int64_t r0 = warp_global * (RPW * R); r0 < rows; r0 += warp_stride * (RPW * R)) {
    PkRow<TY, VE> xv[R][K];
    PkRow<TR, VE> rv[R][K];
    // phase 1: every y vector of the R rows in flight at once
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const int64_t r = r0 + rr * RPW + gi;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int v = gl + k * GS;
        if (r < rows && v < vpr) xv[rr][k].load(y + r * C + v * VE);
        else xv[rr][k].zero();
      }
    }
    // phase 2: the shortcut vectors follow immediately (independent of phase 1's data)
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const int64_t r = r0 + rr * RPW + gi;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int v = gl + k * GS;
        if (shortcut != nullptr && r < rows && v < vpr) rv[rr][k].load(shortcut + r * C + v * VE);
        else rv[rr][k].zero();
      }
    }
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const int64_t r = r0 + rr * RPW + gi;
      const bool row_ok = r < rows;
      float2 sum2 = f2s(0.f);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float x[VE];
        xv[rr][k].get(x);
#pragma unroll
        for (int e = 0; e < VE; e += 2) sum2 = f2add(sum2, f2add(f2(x[e], x[e + 1]), f2(bia[k][e], bia[k][e + 1])));
      }
      const float mean = group_sum<GS>(sum2.x + sum2.y) * inv_c;
      const float2 nmean = f2s(-mean);
      float2 sq2 = f2s(0.f);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (gl + k * GS < vpr) {
          float x[VE];
          xv[rr][k].get(x);
#pragma unroll
          for (int e = 0; e < VE; e += 2) {
            const float2 dlt = f2add(f2add(f2(x[e], x[e + 1]), f2(bia[k][e], bia[k][e + 1])), nmean);
            sq2 = f2fma(dlt, dlt, sq2);
          }
        }
      }
      const float rstd = rsqrtf(group_sum<GS>(sq2.x + sq2.y) * inv_c + eps);
      const float ks = (keep_scale != nullptr && row_ok) ? keep_scale[r / rows_per_sample] : 1.0f;
      const float2 rs2 = f2s(rstd), nmr2 = f2s(-mean * rstd), ks2 = f2s(ks);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int v = gl + k * GS;
        if (row_ok && v < vpr) {
          float x[VE], res[VE];
          xv[rr][k].get(x);
          rv[rr][k].get(res);
#pragma unroll
          for (int e = 0; e < VE; e += 2) {
            // res + ks * (((x + bias) * rstd - mean * rstd) * gamma + beta)
            const float2 xh = f2fma(f2add(f2(x[e], x[e + 1]), f2(bia[k][e], bia[k][e + 1])), rs2, nmr2);
            const float2 o = f2fma(ks2, f2fma(xh, f2(gam[k][e], gam[k][e + 1]), f2(bet[k][e], bet[k][e + 1])), f2(res[e], res[e + 1]));
            res[e] = o.x;
            res[e + 1] = o.y;
          }
          PkRow<TR, VE> o;
          o.set(res);
          o.store(out + r * C + v * VE);
        }
      }
      if (row_ok && gl == 0) {
        mean_out[r] = mean;
        rstd_out[r] = rstd;
      }
    }
  }
}

// workspace layout: [gridDim.x][3][C] float partials (dgamma, dbeta, dbias)
template <typename TY, typename TR, int GS, int K, int R, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
ln_residual_bwd_kernel(const TR* __restrict__ dout, const TY* __restrict__ y, const float* __restrict__ gamma,
                       const float* __restrict__ bias, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                       const float* __restrict__ keep_scale, TY* __restrict__ dy, float* __restrict__ partials, int64_t rows,
                       int C, int64_t rows_per_sample) {
  constexpr int VE = 16 / sizeof(TY);
  constexpr int RPW = 32 / GS;
  extern __shared__ float red[];  // [3][C] per CTA
  const int lane = threadIdx.x & 31, gl = lane % GS, gi = lane / GS;
  const int vpr = C / VE;
  float gam[K][VE], bia[K][VE], dg[K][VE], db[K][VE], dbi[K][VE];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int v = gl + k * GS;
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      gam[k][e] = (v < vpr) ? gamma[v * VE + e] : 0.f;
      bia[k][e] = (v < vpr && bias != nullptr) ? bias[v * VE + e] : 0.f;
      dg[k][e] = 0.f;
      db[k][e] = 0.f;
      dbi[k][e] = 0.f;
    }
  }
  for (int c = threadIdx.x; c < 3 * C; c += kThreads) red[c] = 0.f;
  __syncthreads();
  const int64_t warp_global = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)gridDim.x * (kThreads / 32);
  const float inv_c = 1.0f / (float)C;
  for (int64_t r0 = warp_global * (RPW * R); r0 < rows; r0 += warp_stride * (RPW * R)) {
    PkRow<TY, VE> yv[R][K];
    PkRow<TR, VE> gv[R][K];
    float mean[R], rstd[R], ks[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const int64_t r = r0 + rr * RPW + gi;
      const bool row_ok = r < rows;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int v = gl + k * GS;
        if (row_ok && v < vpr) {
          yv[rr][k].load(y + r * C + v * VE);
          gv[rr][k].load(dout + r * C + v * VE);
        } else {
          yv[rr][k].zero();
          gv[rr][k].zero();
        }
      }
      mean[rr] = row_ok ? mean_in[r] : 0.f;
      rstd[rr] = row_ok ? rstd_in[r] : 0.f;
      ks[rr] = (keep_scale != nullptr && row_ok) ? keep_scale[r / rows_per_sample] : 1.0f;
    }
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const int64_t r = r0 + rr * RPW + gi;
      const bool row_ok = r < rows;
      float xh[K][VE], gx[K][VE];
      float2 s1 = f2s(0.f), s2 = f2s(0.f);
      const float2 ks2 = f2s(ks[rr]), rs2 = f2s(rstd[rr]), nmr2 = f2s(-mean[rr] * rstd[rr]);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        // rows / vectors outside the tensor were loaded as zeros (and mean = rstd = 0, gamma = bias = 0 there), so every
        // sum below receives exact zeros from them without a select
        float yy[VE], gg[VE];
        yv[rr][k].get(yy);
        gv[rr][k].get(gg);
#pragma unroll
        for (int e = 0; e < VE; e += 2) {
          const float2 g = f2mul(f2(gg[e], gg[e + 1]), ks2);
          const float2 xhat = f2fma(f2add(f2(yy[e], yy[e + 1]), f2(bia[k][e], bia[k][e + 1])), rs2, nmr2);
          xh[k][e] = xhat.x; xh[k][e + 1] = xhat.y;
          const float2 dg2 = f2fma(g, xhat, f2(dg[k][e], dg[k][e + 1]));
          dg[k][e] = dg2.x; dg[k][e + 1] = dg2.y;
          const float2 db2 = f2add(f2(db[k][e], db[k][e + 1]), g);
          db[k][e] = db2.x; db[k][e + 1] = db2.y;
          const float2 gxv = f2mul(g, f2(gam[k][e], gam[k][e + 1]));
          gx[k][e] = gxv.x; gx[k][e + 1] = gxv.y;
          s1 = f2fma(gxv, xhat, s1);
          s2 = f2add(s2, gxv);
        }
      }
      const float c1 = group_sum<GS>(s1.x + s1.y) * inv_c;
      const float c2 = group_sum<GS>(s2.x + s2.y) * inv_c;
      const float2 nc1 = f2s(-c1), nc2 = f2s(-c2);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int v = gl + k * GS;
        if (row_ok && v < vpr) {
          float o[VE];
#pragma unroll
          for (int e = 0; e < VE; e += 2) {
            // rstd * (gx - c2 - xhat * c1)
            const float2 t = f2mul(rs2, f2fma(f2(xh[k][e], xh[k][e + 1]), nc1, f2add(f2(gx[k][e], gx[k][e + 1]), nc2)));
            o[e] = t.x; o[e + 1] = t.y;
            const float2 d2 = f2add(f2(dbi[k][e], dbi[k][e + 1]), t);
            dbi[k][e] = d2.x; dbi[k][e + 1] = d2.y;
          }
          PkRow<TY, VE> ov;
          ov.set(o);
          ov.store(dy + r * C + v * VE);
        }
      }
    }
  }
  // fold the 32/GS row groups of a warp (shuffles), then the warps of the CTA (shared-memory atomics on
  // distinct columns), then one plain store per column per CTA.
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      float a = dg[k][e], b = db[k][e], c = dbi[k][e];
#pragma unroll
      for (int o = GS; o < 32; o <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
      }
      const int v = gl + k * GS;
      if (gi == 0 && v < vpr) {
        atomicAdd(&red[v * VE + e], a);
        atomicAdd(&red[C + v * VE + e], b);
        atomicAdd(&red[2 * C + v * VE + e], c);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 3 * C; c += kThreads) partials[(int64_t)blockIdx.x * 3 * C + c] = red[c];
}

// Sum the per-CTA partial rows: block = 32 columns x kFinalizeRows row-groups, coalesced 128-byte reads.
__global__ void __launch_bounds__(32 * kFinalizeRows) ln_param_grad_finalize_kernel(const float* __restrict__ partials, int nblocks,
                                                                                   int C, float* __restrict__ dgamma,
                                                                                   float* __restrict__ dbeta,
                                                                                   float* __restrict__ dbias) {
  __shared__ float sm[kFinalizeRows][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < 3 * C)
    for (int b = ty; b < nblocks; b += kFinalizeRows) s += partials[(int64_t)b * 3 * C + c];
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < 3 * C) {
#pragma unroll
    for (int k = 1; k < kFinalizeRows; ++k) s += sm[k][tx];
    if (c < C) dgamma[c] = s;
    else if (c < 2 * C) dbeta[c - C] = s;
    else if (dbias != nullptr) dbias[c - 2 * C] = s;
  }
}

int grid_for_rows(int64_t rows, int rows_per_warp_iter, int ctas_per_sm) {
  const int64_t warps = (rows + rows_per_warp_iter - 1) / rows_per_warp_iter;
  const int64_t blocks = (warps + (kThreads / 32) - 1) / (kThreads / 32);
  const int64_t cap = (int64_t)num_sms() * ctas_per_sm;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}
constexpr int kBwdCtasPerSm = 2;  // also bounds the number of partial rows the finalize kernel sums

struct Shape { int gs, k; };
bool pick_shape(int vpr, bool exact_fit, Shape& s) {
  // rows of 12 / 24 / 48 vectors (C = 96 / 192 / 384 in bf16: stages 0-2 of SwinV2-T) fill a lane group exactly with three
  // vectors per lane; the power-of-two groups below would leave a quarter of the lanes idle
  if (exact_fit && (vpr == 12 || vpr == 24 || vpr == 48)) { s.gs = vpr / 3; s.k = 3; return true; }
  static const Shape table[] = {{8, 1}, {16, 1}, {32, 1}, {32, 2}, {32, 3}, {32, 4}, {32, 6}, {32, 8}};
  for (const Shape& t : table)
    if (vpr <= t.gs * t.k) { s = t; return true; }
  return false;
}

template <typename TY, typename TR, int GS, int K>
int run_fwd(const void* y, const void* sc, const float* gamma, const float* beta, const float* bias, const float* ks,
            void* out, float* mean, float* rstd, int64_t rows, int C, int64_t rps, float eps, cudaStream_t st) {
  constexpr int R = K == 1 ? 4 : ((K == 2 || GS < 32) ? 2 : 1);
  constexpr int MINB = (K <= 2 || GS < 32) ? 2 : 1;
  const int grid = grid_for_rows(rows, (32 / GS) * R, 2 * MINB);
  ln_residual_fwd_kernel<TY, TR, GS, K, R, MINB><<<grid, kThreads, 0, st>>>((const TY*)y, (const TR*)sc, gamma, beta, bias, ks,
                                                                      (TR*)out, mean, rstd, rows, C, rps, eps);
  HV_LAUNCH_OK("ln_residual_fwd_kernel");
  return HV_OK;
}

template <typename TY, typename TR, int GS, int K>
int run_bwd(const void* dout, const void* y, const float* gamma, const float* bias, const float* mean, const float* rstd,
            const float* ks, void* dy, float* dgamma, float* dbeta, float* dbias, float* partials, int64_t rows, int C,
            int64_t rps, cudaStream_t st) {
  // rows in flight per lane group: enough 16-byte requests outstanding to cover the HBM latency at 2 CTAs per SM
  constexpr int R = K == 1 ? 4 : (K == 3 ? 2 : 1);
  constexpr int MINB = K <= 2 ? 2 : 1;
  const int grid = grid_for_rows(rows, (32 / GS) * R, kBwdCtasPerSm);
  ln_residual_bwd_kernel<TY, TR, GS, K, R, MINB><<<grid, kThreads, 3 * C * sizeof(float), st>>>(
      (const TR*)dout, (const TY*)y, gamma, bias, mean, rstd, ks, (TY*)dy, partials, rows, C, rps);
  HV_LAUNCH_OK("ln_residual_bwd_kernel");
  ln_param_grad_finalize_kernel<<<(3 * C + 31) / 32, 32 * kFinalizeRows, 0, st>>>(partials, grid, C, dgamma, dbeta, dbias);
  HV_LAUNCH_OK("ln_param_grad_finalize_kernel");
  return HV_OK;
}

#define HV_LN_DISPATCH_SHAPE(FN, TY, TR, ...)                                  \
  switch (shape.gs * 100 + shape.k) {                                          \
    case 403:  return FN<TY, TR, 4, 3>(__VA_ARGS__);                           \
    case 803:  return FN<TY, TR, 8, 3>(__VA_ARGS__);                           \
    case 1603: return FN<TY, TR, 16, 3>(__VA_ARGS__);                          \
    case 801:  return FN<TY, TR, 8, 1>(__VA_ARGS__);                           \
    case 1601: return FN<TY, TR, 16, 1>(__VA_ARGS__);                          \
    case 3201: return FN<TY, TR, 32, 1>(__VA_ARGS__);                          \
    case 3202: return FN<TY, TR, 32, 2>(__VA_ARGS__);                          \
    case 3203: return FN<TY, TR, 32, 3>(__VA_ARGS__);                          \
    case 3204: return FN<TY, TR, 32, 4>(__VA_ARGS__);                          \
    case 3206: return FN<TY, TR, 32, 6>(__VA_ARGS__);                          \
    default:   return FN<TY, TR, 32, 8>(__VA_ARGS__);                          \
  }

int check_common(int64_t rows, int C, int y_dtype, int res_dtype, bool backward, Shape& shape) {
  if (rows <= 0 || C <= 0) HV_FAIL(HV_ERR_SHAPE, "ln_residual: rows=%lld C=%d", (long long)rows, C);
  if (!((y_dtype == HV_F32 && res_dtype == HV_F32) || (y_dtype == HV_BF16 && (res_dtype == HV_BF16 || res_dtype == HV_F32))))
    HV_FAIL(HV_ERR_DTYPE, "ln_residual: unsupported dtype pair y=%d residual=%d", y_dtype, res_dtype);
  const int ve = y_dtype == HV_F32 ? 4 : 8;
  if (C % ve != 0) HV_FAIL(HV_ERR_SHAPE, "ln_residual: C=%d must be a multiple of %d", C, ve);
  // exact-fit groups: always in the backward; in the forward only where measured faster (same-type residual, 24 vectors
  // per row or fp32 rows -- the other cases run out of registers with three vectors and two tensors per lane)
  const int vpr = C / ve;
  const bool exact_fit = backward || (y_dtype == res_dtype && (vpr == 24 || y_dtype == HV_F32));
  if (!pick_shape(vpr, exact_fit, shape)) HV_FAIL(HV_ERR_SHAPE, "ln_residual: C=%d too wide (max %d)", C, 256 * ve);
  return HV_OK;
}

}  // namespace

size_t ln_residual_bwd_workspace_bytes(int64_t rows, int C) {
  (void)rows;
  return (size_t)num_sms() * kBwdCtasPerSm * 3 * (size_t)C * sizeof(float);
}

int ln_residual_fwd(const void* y, const void* shortcut, const float* gamma, const float* beta, const float* bias,
                    const float* keep_scale, void* out, float* mean, float* rstd, int64_t rows, int C,
                    int64_t rows_per_sample, float eps, int y_dtype, int res_dtype, cudaStream_t st) {
  Shape shape;
  int rc = check_common(rows, C, y_dtype, res_dtype, false, shape);
  if (rc) return rc;
  if (!aligned16(y) || !aligned16(out) || (shortcut && !aligned16(shortcut))) HV_FAIL(HV_ERR_ALIGN, "ln_residual_fwd: pointers must be 16-byte aligned");
  if (rows_per_sample <= 0) rows_per_sample = rows;
  if (y_dtype == HV_F32) {
    HV_LN_DISPATCH_SHAPE(run_fwd, float, float, y, shortcut, gamma, beta, bias, keep_scale, out, mean, rstd, rows, C, rows_per_sample, eps, st)
  } else if (res_dtype == HV_BF16) {
    HV_LN_DISPATCH_SHAPE(run_fwd, bf16, bf16, y, shortcut, gamma, beta, bias, keep_scale, out, mean, rstd, rows, C, rows_per_sample, eps, st)
  } else {
    HV_LN_DISPATCH_SHAPE(run_fwd, bf16, float, y, shortcut, gamma, beta, bias, keep_scale, out, mean, rstd, rows, C, rows_per_sample, eps, st)
  }
}

int ln_residual_bwd(const void* dout, const void* y, const float* gamma, const float* bias, const float* mean,
                    const float* rstd, const float* keep_scale, void* dy, float* dgamma, float* dbeta, float* dbias,
                    void* workspace, size_t workspace_bytes, int64_t rows, int C, int64_t rows_per_sample, int y_dtype,
                    int res_dtype, cudaStream_t st) {
  Shape shape;
  int rc = check_common(rows, C, y_dtype, res_dtype, true, shape);
  if (rc) return rc;
  if (!aligned16(y) || !aligned16(dout) || !aligned16(dy)) HV_FAIL(HV_ERR_ALIGN, "ln_residual_bwd: pointers must be 16-byte aligned");
  if (workspace == nullptr || workspace_bytes < ln_residual_bwd_workspace_bytes(rows, C))
    HV_FAIL(HV_ERR_WORKSPACE, "ln_residual_bwd: workspace of %zu bytes required", ln_residual_bwd_workspace_bytes(rows, C));
  if (rows_per_sample <= 0) rows_per_sample = rows;
  float* partials = static_cast<float*>(workspace);
  if (y_dtype == HV_F32) {
    HV_LN_DISPATCH_SHAPE(run_bwd, float, float, dout, y, gamma, bias, mean, rstd, keep_scale, dy, dgamma, dbeta, dbias, partials, rows, C, rows_per_sample, st)
  } else if (res_dtype == HV_BF16) {
    HV_LN_DISPATCH_SHAPE(run_bwd, bf16, bf16, dout, y, gamma, bias, mean, rstd, keep_scale, dy, dgamma, dbeta, dbias, partials, rows, C, rows_per_sample, st)
  } else {
    HV_LN_DISPATCH_SHAPE(run_bwd, bf16, float, dout, y, gamma, bias, mean, rstd, keep_scale, dy, dgamma, dbeta, dbias, partials, rows, C, rows_per_sample, st)
  }
}

}  // namespace hv
