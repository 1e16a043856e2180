// Res-post-norm: out = shortcut + keep_scale[sample] * LayerNorm(y)   (reference swinv2.py:431, 434)
// and, with shortcut == nullptr, the plain LayerNorm of PatchMerging (swinv2.py:494).
//
// HBM-bound streaming kernels: a row (token) is owned by a group of GS lanes, every lane moves
// K 16-byte vectors of y per row, statistics are reduced with warp shuffles, gamma/beta live in
// registers for the lifetime of the thread.  Backward keeps its d-gamma / d-beta partial sums in
// registers across all the rows a thread visits and reduces them once per CTA (deterministic
// two-stage reduction through a workspace; no atomics).
#include "hv_common.cuh"

namespace hv {
namespace {

constexpr int kThreads = 256;

template <typename T, int VE> struct RowVec {  // VE consecutive elements of T <-> fp32 registers
  static constexpr int kVecs = VE * sizeof(T) / 16;
  __device__ __forceinline__ static void load(const T* p, float (&f)[VE]) {
#pragma unroll
    for (int v = 0; v < kVecs; ++v) {
      float t[Vec16<T>::n];
      Vec16<T>::load(p + v * Vec16<T>::n, t);
#pragma unroll
      for (int e = 0; e < Vec16<T>::n; ++e) f[v * Vec16<T>::n + e] = t[e];
    }
  }
  __device__ __forceinline__ static void store(T* p, const float (&f)[VE]) {
#pragma unroll
    for (int v = 0; v < kVecs; ++v) {
      float t[Vec16<T>::n];
#pragma unroll
      for (int e = 0; e < Vec16<T>::n; ++e) t[e] = f[v * Vec16<T>::n + e];
      Vec16<T>::store(p + v * Vec16<T>::n, t);
    }
  }
};

template <typename TY, typename TR, int GS, int K>
__global__ void __launch_bounds__(kThreads) ln_residual_fwd_kernel(const TY* __restrict__ y, const TR* __restrict__ shortcut,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta,
                                                                   const float* __restrict__ keep_scale, TR* __restrict__ out,
                                                                   float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                                   int64_t rows, int C, int64_t rows_per_sample, float eps) {
  constexpr int VE = 16 / sizeof(TY);
  constexpr int RPW = 32 / GS;  // rows per warp per iteration
  const int lane = threadIdx.x & 31, gl = lane % GS, gi = lane / GS;
  const int vpr = C / VE;  // vectors per row
  float gam[K][VE], bet[K][VE];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int v = gl + k * GS;
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      gam[k][e] = (v < vpr) ? gamma[v * VE + e] : 0.f;
      bet[k][e] = (v < vpr) ? beta[v * VE + e] : 0.f;
    }
  }
  const int64_t warp_global = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)gridDim.x * (kThreads / 32);
  const float inv_c = 1.0f / (float)C;
  for (int64_t r0 = warp_global * RPW; r0 < rows; r0 += warp_stride * RPW) {
    const int64_t r = r0 + gi;
    const bool row_ok = r < rows;
    float x[K][VE];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int v = gl + k * GS;
      if (row_ok && v < vpr) {
        RowVec<TY, VE>::load(y + r * C + v * VE, x[k]);
      } else {
#pragma unroll
        for (int e = 0; e < VE; ++e) x[k][e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < VE; ++e) sum += x[k][e];
    }
    const float mean = group_sum<GS>(sum) * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int v = gl + k * GS;
      if (v < vpr) {
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          const float dlt = x[k][e] - mean;
          sq = fmaf(dlt, dlt, sq);
        }
      }
    }
    const float rstd = rsqrtf(group_sum<GS>(sq) * inv_c + eps);
    const float ks = (keep_scale != nullptr && row_ok) ? keep_scale[r / rows_per_sample] : 1.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int v = gl + k * GS;
      if (row_ok && v < vpr) {
        float res[VE];
        if (shortcut != nullptr) {
          RowVec<TR, VE>::load(shortcut + r * C + v * VE, res);
        } else {
#pragma unroll
          for (int e = 0; e < VE; ++e) res[e] = 0.f;
        }
#pragma unroll
        for (int e = 0; e < VE; ++e) res[e] += ks * fmaf((x[k][e] - mean) * rstd, gam[k][e], bet[k][e]);
        RowVec<TR, VE>::store(out + r * C + v * VE, res);
      }
    }
    if (row_ok && gl == 0) {
      mean_out[r] = mean;
      rstd_out[r] = rstd;
    }
  }
}

// workspace layout: [gridDim.x][2][C] float partials (dgamma, dbeta)
template <typename TY, typename TR, int GS, int K>
__global__ void __launch_bounds__(kThreads) ln_residual_bwd_kernel(const TR* __restrict__ dout, const TY* __restrict__ y,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ mean_in,
                                                                   const float* __restrict__ rstd_in,
                                                                   const float* __restrict__ keep_scale, TY* __restrict__ dy,
                                                                   float* __restrict__ partials, int64_t rows, int C,
                                                                   int64_t rows_per_sample) {
  constexpr int VE = 16 / sizeof(TY);
  constexpr int RPW = 32 / GS;
  extern __shared__ float red[];  // [2][C] per CTA
  const int lane = threadIdx.x & 31, gl = lane % GS, gi = lane / GS;
  const int vpr = C / VE;
  float gam[K][VE], dg[K][VE], db[K][VE];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int v = gl + k * GS;
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      gam[k][e] = (v < vpr) ? gamma[v * VE + e] : 0.f;
      dg[k][e] = 0.f;
      db[k][e] = 0.f;
    }
  }
  for (int c = threadIdx.x; c < 2 * C; c += kThreads) red[c] = 0.f;
  __syncthreads();
  const int64_t warp_global = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)gridDim.x * (kThreads / 32);
  const float inv_c = 1.0f / (float)C;
  for (int64_t r0 = warp_global * RPW; r0 < rows; r0 += warp_stride * RPW) {
    const int64_t r = r0 + gi;
    const bool row_ok = r < rows;
    const float mean = row_ok ? mean_in[r] : 0.f;
    const float rstd = row_ok ? rstd_in[r] : 0.f;
    const float ks = (keep_scale != nullptr && row_ok) ? keep_scale[r / rows_per_sample] : 1.0f;
    float xh[K][VE], gx[K][VE];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int v = gl + k * GS;
      if (row_ok && v < vpr) {
        float gv[VE];
        RowVec<TY, VE>::load(y + r * C + v * VE, xh[k]);
        RowVec<TR, VE>::load(dout + r * C + v * VE, gv);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          const float g = gv[e] * ks;
          xh[k][e] = (xh[k][e] - mean) * rstd;
          dg[k][e] = fmaf(g, xh[k][e], dg[k][e]);
          db[k][e] += g;
          gx[k][e] = g * gam[k][e];
          s1 = fmaf(gx[k][e], xh[k][e], s1);
          s2 += gx[k][e];
        }
      } else {
#pragma unroll
        for (int e = 0; e < VE; ++e) { xh[k][e] = 0.f; gx[k][e] = 0.f; }
      }
    }
    const float c1 = group_sum<GS>(s1) * inv_c;
    const float c2 = group_sum<GS>(s2) * inv_c;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int v = gl + k * GS;
      if (row_ok && v < vpr) {
        float o[VE];
#pragma unroll
        for (int e = 0; e < VE; ++e) o[e] = rstd * (gx[k][e] - c2 - xh[k][e] * c1);
        RowVec<TY, VE>::store(dy + r * C + v * VE, o);
      }
    }
  }
  // fold the 32/GS row groups of a warp, then the warps of the CTA (shared-memory atomics: 8 warps,
  // distinct columns within a warp instruction), then one plain store per column per CTA.
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      float a = dg[k][e], b = db[k][e];
#pragma unroll
      for (int o = GS; o < 32; o <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      const int v = gl + k * GS;
      if (gi == 0 && v < vpr) {
        atomicAdd(&red[v * VE + e], a);
        atomicAdd(&red[C + v * VE + e], b);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += kThreads) partials[(int64_t)blockIdx.x * 2 * C + c] = red[c];
}

__global__ void ln_param_grad_finalize_kernel(const float* __restrict__ partials, int nblocks, int C,
                                              float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 2 * C) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partials[(int64_t)b * 2 * C + c];
  if (c < C) dgamma[c] = s; else dbeta[c - C] = s;
}

int grid_for_rows(int64_t rows, int rows_per_warp) {
  const int64_t warps = (rows + rows_per_warp - 1) / rows_per_warp;
  const int64_t blocks = (warps + (kThreads / 32) - 1) / (kThreads / 32);
  const int64_t cap = (int64_t)num_sms() * 8;  // 8 CTAs of 256 threads = 2048 threads per SM
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

struct Shape { int gs, k; };
bool pick_shape(int vpr, Shape& s) {
  static const Shape table[] = {{8, 1}, {16, 1}, {32, 1}, {32, 2}, {32, 3}, {32, 4}, {32, 6}, {32, 8}};
  for (const Shape& t : table)
    if (vpr <= t.gs * t.k) { s = t; return true; }
  return false;
}

template <typename TY, typename TR, int GS, int K>
int run_fwd(const void* y, const void* sc, const float* gamma, const float* beta, const float* ks, void* out, float* mean,
            float* rstd, int64_t rows, int C, int64_t rps, float eps, cudaStream_t st) {
  const int grid = grid_for_rows(rows, 32 / GS);
  ln_residual_fwd_kernel<TY, TR, GS, K><<<grid, kThreads, 0, st>>>((const TY*)y, (const TR*)sc, gamma, beta, ks, (TR*)out,
                                                                   mean, rstd, rows, C, rps, eps);
  HV_LAUNCH_OK("ln_residual_fwd_kernel");
  return HV_OK;
}

template <typename TY, typename TR, int GS, int K>
int run_bwd(const void* dout, const void* y, const float* gamma, const float* mean, const float* rstd, const float* ks,
            void* dy, float* dgamma, float* dbeta, float* partials, int64_t rows, int C, int64_t rps, cudaStream_t st) {
  const int grid = grid_for_rows(rows, 32 / GS);
  ln_residual_bwd_kernel<TY, TR, GS, K><<<grid, kThreads, 2 * C * sizeof(float), st>>>(
      (const TR*)dout, (const TY*)y, gamma, mean, rstd, ks, (TY*)dy, partials, rows, C, rps);
  HV_LAUNCH_OK("ln_residual_bwd_kernel");
  ln_param_grad_finalize_kernel<<<(2 * C + 127) / 128, 128, 0, st>>>(partials, grid, C, dgamma, dbeta);
  HV_LAUNCH_OK("ln_param_grad_finalize_kernel");
  return HV_OK;
}

#define HV_LN_DISPATCH_SHAPE(FN, TY, TR, ...)                                  \
  switch (shape.gs * 100 + shape.k) {                                          \
    case 801:  return FN<TY, TR, 8, 1>(__VA_ARGS__);                           \
    case 1601: return FN<TY, TR, 16, 1>(__VA_ARGS__);                          \
    case 3201: return FN<TY, TR, 32, 1>(__VA_ARGS__);                          \
    case 3202: return FN<TY, TR, 32, 2>(__VA_ARGS__);                          \
    case 3203: return FN<TY, TR, 32, 3>(__VA_ARGS__);                          \
    case 3204: return FN<TY, TR, 32, 4>(__VA_ARGS__);                          \
    case 3206: return FN<TY, TR, 32, 6>(__VA_ARGS__);                          \
    default:   return FN<TY, TR, 32, 8>(__VA_ARGS__);                          \
  }

int check_common(int64_t rows, int C, int y_dtype, int res_dtype, Shape& shape) {
  if (rows <= 0 || C <= 0) HV_FAIL(HV_ERR_SHAPE, "ln_residual: rows=%lld C=%d", (long long)rows, C);
  if (!((y_dtype == HV_F32 && res_dtype == HV_F32) || (y_dtype == HV_BF16 && (res_dtype == HV_BF16 || res_dtype == HV_F32))))
    HV_FAIL(HV_ERR_DTYPE, "ln_residual: unsupported dtype pair y=%d residual=%d", y_dtype, res_dtype);
  const int ve = y_dtype == HV_F32 ? 4 : 8;
  if (C % ve != 0) HV_FAIL(HV_ERR_SHAPE, "ln_residual: C=%d must be a multiple of %d", C, ve);
  if (!pick_shape(C / ve, shape)) HV_FAIL(HV_ERR_SHAPE, "ln_residual: C=%d too wide (max %d)", C, 256 * ve);
  return HV_OK;
}

}  // namespace

size_t ln_residual_bwd_workspace_bytes(int64_t rows, int C) {
  (void)rows;
  return (size_t)num_sms() * 8 * 2 * (size_t)C * sizeof(float);
}

int ln_residual_fwd(const void* y, const void* shortcut, const float* gamma, const float* beta, const float* keep_scale,
                    void* out, float* mean, float* rstd, int64_t rows, int C, int64_t rows_per_sample, float eps,
                    int y_dtype, int res_dtype, cudaStream_t st) {
  Shape shape;
  int rc = check_common(rows, C, y_dtype, res_dtype, shape);
  if (rc) return rc;
  if (!aligned16(y) || !aligned16(out) || (shortcut && !aligned16(shortcut))) HV_FAIL(HV_ERR_ALIGN, "ln_residual_fwd: pointers must be 16-byte aligned");
  if (rows_per_sample <= 0) rows_per_sample = rows;
  if (y_dtype == HV_F32) {
    HV_LN_DISPATCH_SHAPE(run_fwd, float, float, y, shortcut, gamma, beta, keep_scale, out, mean, rstd, rows, C, rows_per_sample, eps, st)
  } else if (res_dtype == HV_BF16) {
    HV_LN_DISPATCH_SHAPE(run_fwd, bf16, bf16, y, shortcut, gamma, beta, keep_scale, out, mean, rstd, rows, C, rows_per_sample, eps, st)
  } else {
    HV_LN_DISPATCH_SHAPE(run_fwd, bf16, float, y, shortcut, gamma, beta, keep_scale, out, mean, rstd, rows, C, rows_per_sample, eps, st)
  }
}

int ln_residual_bwd(const void* dout, const void* y, const float* gamma, const float* mean, const float* rstd,
                    const float* keep_scale, void* dy, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                    int64_t rows, int C, int64_t rows_per_sample, int y_dtype, int res_dtype, cudaStream_t st) {
  Shape shape;
  int rc = check_common(rows, C, y_dtype, res_dtype, shape);
  if (rc) return rc;
  if (!aligned16(y) || !aligned16(dout) || !aligned16(dy)) HV_FAIL(HV_ERR_ALIGN, "ln_residual_bwd: pointers must be 16-byte aligned");
  if (workspace == nullptr || workspace_bytes < ln_residual_bwd_workspace_bytes(rows, C))
    HV_FAIL(HV_ERR_WORKSPACE, "ln_residual_bwd: workspace of %zu bytes required", ln_residual_bwd_workspace_bytes(rows, C));
  if (rows_per_sample <= 0) rows_per_sample = rows;
  float* partials = static_cast<float*>(workspace);
  if (y_dtype == HV_F32) {
    HV_LN_DISPATCH_SHAPE(run_bwd, float, float, dout, y, gamma, mean, rstd, keep_scale, dy, dgamma, dbeta, partials, rows, C, rows_per_sample, st)
  } else if (res_dtype == HV_BF16) {
    HV_LN_DISPATCH_SHAPE(run_bwd, bf16, bf16, dout, y, gamma, mean, rstd, keep_scale, dy, dgamma, dbeta, partials, rows, C, rows_per_sample, st)
  } else {
    HV_LN_DISPATCH_SHAPE(run_bwd, bf16, float, dout, y, gamma, mean, rstd, keep_scale, dy, dgamma, dbeta, partials, rows, C, rows_per_sample, st)
  }
}

}  // namespace hv
