// Generic fused shifted-window scaled-cosine attention (any window size up to 16x16, any head
// dim up to 64, fp32 or bf16 I/O, fp32 CUDA-core math).  One CTA per (window, head); thread i owns
// query slot i in the forward / dQ pass and key slot j in the dK/dV pass.
//
// This is the accuracy path (fp32 activations, odd window sizes such as the reference default 7,
// explicit masks of WindowAttention.forward(x, mask)); the bf16 N=64/d=32 shapes of SwinV2-T/B
// dispatch to the tensor-core kernel in wattn_mma64.cu instead.
//
// The cyclic shift, window partition and their inverses (reference swinv2.py:399-412, 420-429)
// exist only as address arithmetic (hv::window_slot_to_token); the shifted-window mask
// (swinv2.py:357-388) is evaluated from region ids, never read from memory.
#include "hv_common.cuh"

namespace hv {
namespace {

struct GenericSmem {
  // layout computed identically on host and device
  int pitch;
  size_t off_tok, off_region, off_bias, off_vec, off_rows;
  size_t bytes;
};

__host__ __device__ inline GenericSmem generic_layout(int N, int d, int ws, int n_row_arrays, int n_vecs, int n_tabs) {
  GenericSmem L;
  L.pitch = d + 1;
  size_t o = 0;
  L.off_tok = o;    o += sizeof(int64_t) * N;
  L.off_region = o; o += sizeof(int) * N;
  L.off_bias = o;   o += sizeof(float) * (2 * ws - 1) * (2 * ws - 1) * n_tabs;
  L.off_vec = o;    o += sizeof(float) * N * n_vecs;
  o = (o + 15) & ~size_t(15);
  L.off_rows = o;   o += sizeof(float) * N * L.pitch * n_row_arrays;
  L.bytes = o;
  return L;
}

template <typename T, int DMAX>
__global__ void __launch_bounds__(256) wattn_generic_fwd_kernel(const T* __restrict__ qkv, const float* __restrict__ bias_table,
                                                                const float* __restrict__ tau, const float* __restrict__ mask,
                                                                int mask_windows, T* __restrict__ out, float* __restrict__ lse,
                                                                Geom g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = g.N, d = g.d, ws = g.ws;
  const GenericSmem L = generic_layout(N, d, ws, 3, 1, 1);
  int64_t* tok = reinterpret_cast<int64_t*>(smem_raw + L.off_tok);
  int* region = reinterpret_cast<int*>(smem_raw + L.off_region);
  float* bias_s = reinterpret_cast<float*>(smem_raw + L.off_bias);
  float* rk_s = reinterpret_cast<float*>(smem_raw + L.off_vec);
  float* Qs = reinterpret_cast<float*>(smem_raw + L.off_rows);
  float* Ks = Qs + N * L.pitch;
  float* Vs = Ks + N * L.pitch;
  const int pitch = L.pitch;

  const int head = blockIdx.x % g.heads;
  const int brow = blockIdx.x / g.heads;  // b * nW + win
  const int b = brow / g.nW, win = brow - b * g.nW;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int M = (2 * ws - 1) * (2 * ws - 1);
  const int64_t qkv_stride = 3 * (int64_t)g.C;

  for (int s = tid; s < N; s += nthreads) {
    tok[s] = window_slot_to_token(g, b, win, s);
    region[s] = (g.shift > 0) ? window_slot_region(g, win, s) : 0;
  }
  for (int r = tid; r < M; r += nthreads) bias_s[r] = bias_table[r * g.heads + head];
  __syncthreads();
  {
    // reuse tok[] as element offsets
    for (int e = tid; e < N * d; e += nthreads) {
      const int slot = e / d, c = e - slot * d;
      const T* p = qkv + tok[slot] * qkv_stride + head * d + c;
      Qs[slot * pitch + c] = to_f32(p[0]);
      Ks[slot * pitch + c] = to_f32(p[g.C]);
      Vs[slot * pitch + c] = to_f32(p[2 * g.C]);
    }
  }
  __syncthreads();
  const bool active = tid < N;
  const int i = active ? tid : 0;
  // normalise K row i in place (F.normalize: x / max(||x||, eps), swinv2.py:229)
  if (active) {
    float ss = 0.f;
    for (int c = 0; c < d; ++c) ss += Ks[i * pitch + c] * Ks[i * pitch + c];
    const float r = 1.0f / fmaxf(sqrtf(ss), kNormEps);
    for (int c = 0; c < d; ++c) Ks[i * pitch + c] *= r;
    rk_s[i] = r;
  }
  __syncthreads();
  if (!active) return;

  float q[DMAX], o[DMAX];
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < DMAX; ++c) {
    q[c] = (c < d) ? Qs[i * pitch + c] : 0.f;
    ss += q[c] * q[c];
    o[c] = 0.f;
  }
  const float qscale = tau[head] / fmaxf(sqrtf(ss), kNormEps);
#pragma unroll
  for (int c = 0; c < DMAX; ++c) q[c] *= qscale;

  const int ih = i / ws, iw = i - ih * ws;
  const int reg_i = region[i];
  const bool closed_mask = (mask == nullptr) && (g.shift > 0);
  const float* mrow = mask ? mask + ((int64_t)(brow % mask_windows) * N + i) * N : nullptr;
  float m = -INFINITY, l = 0.f;
  int jh = 0, jw = 0;
  for (int j = 0; j < N; ++j) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < DMAX; ++c)
      if (c < d) s = fmaf(q[c], Ks[j * pitch + c], s);
    s += bias_s[(ih - jh + ws - 1) * (2 * ws - 1) + (iw - jw + ws - 1)];
    if (closed_mask) s += (region[j] != reg_i) ? kMaskValue : 0.f;
    if (mrow) s += mrow[j];
    const float m_new = fmaxf(m, s);
    const float corr = expf(m - m_new);
    const float p = expf(s - m_new);
    l = l * corr + p;
#pragma unroll
    for (int c = 0; c < DMAX; ++c)
      if (c < d) o[c] = fmaf(o[c], corr, p * Vs[j * pitch + c]);
    m = m_new;
    if (++jw == ws) { jw = 0; ++jh; }
  }
  const float inv_l = 1.0f / l;
  T* op = out + tok[i] * (int64_t)g.C + head * d;
#pragma unroll
  for (int c = 0; c < DMAX; ++c)
    if (c < d) op[c] = from_f32<T>(o[c] * inv_l);
  lse[((int64_t)brow * g.heads + head) * N + i] = m + logf(l);
}

template <typename T, int DMAX>
__global__ void __launch_bounds__(256) wattn_generic_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ out,
                                                                const T* __restrict__ dout, const float* __restrict__ lse,
                                                                const float* __restrict__ bias_table,
                                                                const float* __restrict__ tau, const float* __restrict__ mask,
                                                                int mask_windows, T* __restrict__ dqkv,
                                                                float* __restrict__ dbias_table, float* __restrict__ dtau,
                                                                Geom g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = g.N, d = g.d, ws = g.ws;
  const GenericSmem L = generic_layout(N, d, ws, 4, 5, 2);
  int64_t* tok = reinterpret_cast<int64_t*>(smem_raw + L.off_tok);
  int* region = reinterpret_cast<int*>(smem_raw + L.off_region);
  float* bias_s = reinterpret_cast<float*>(smem_raw + L.off_bias);
  float* lse_s = reinterpret_cast<float*>(smem_raw + L.off_vec);
  float* D_s = lse_s + N;
  float* rq_s = D_s + N;
  float* rk_s = rq_s + N;
  float* red_s = rk_s + N;  // N floats scratch (block reduction)
  float* Qs = reinterpret_cast<float*>(smem_raw + L.off_rows);
  float* Ks = Qs + N * L.pitch;
  float* Vs = Ks + N * L.pitch;
  float* Gs = Vs + N * L.pitch;  // dO rows
  const int pitch = L.pitch;

  const int head = blockIdx.x % g.heads;
  const int brow = blockIdx.x / g.heads;
  const int b = brow / g.nW, win = brow - b * g.nW;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int Mtab = (2 * ws - 1) * (2 * ws - 1);
  const int64_t qkv_stride = 3 * (int64_t)g.C;
  const float tau_h = tau[head];
  float* dbias_s = bias_s + Mtab;

  for (int s = tid; s < N; s += nthreads) {
    tok[s] = window_slot_to_token(g, b, win, s);
    region[s] = (g.shift > 0) ? window_slot_region(g, win, s) : 0;
    lse_s[s] = lse[((int64_t)brow * g.heads + head) * N + s];
  }
  for (int r = tid; r < Mtab; r += nthreads) {
    bias_s[r] = bias_table[r * g.heads + head];
    dbias_s[r] = 0.f;
  }
  __syncthreads();
  for (int e = tid; e < N * d; e += nthreads) {
    const int slot = e / d, c = e - slot * d;
    const T* p = qkv + tok[slot] * qkv_stride + head * d + c;
    Qs[slot * pitch + c] = to_f32(p[0]);
    Ks[slot * pitch + c] = to_f32(p[g.C]);
    Vs[slot * pitch + c] = to_f32(p[2 * g.C]);
    Gs[slot * pitch + c] = to_f32(dout[tok[slot] * (int64_t)g.C + head * d + c]);
  }
  __syncthreads();
  const bool active = tid < N;
  const int t = active ? tid : 0;
  if (active) {
    float sq = 0.f, sk = 0.f, dd = 0.f;
    const T* orow = out + tok[t] * (int64_t)g.C + head * d;
    for (int c = 0; c < d; ++c) {
      sq += Qs[t * pitch + c] * Qs[t * pitch + c];
      sk += Ks[t * pitch + c] * Ks[t * pitch + c];
      dd += Gs[t * pitch + c] * to_f32(orow[c]);
    }
    const float rq = 1.0f / fmaxf(sqrtf(sq), kNormEps);
    const float rk = 1.0f / fmaxf(sqrtf(sk), kNormEps);
    for (int c = 0; c < d; ++c) {
      Qs[t * pitch + c] *= rq;  // q-hat
      Ks[t * pitch + c] *= rk;  // k-hat
    }
    rq_s[t] = rq; rk_s[t] = rk; D_s[t] = dd;
  }
  __syncthreads();

  const bool closed_mask = (mask == nullptr) && (g.shift > 0);
  const float* mbase = mask ? mask + (int64_t)(brow % mask_windows) * N * N : nullptr;
  const int th = t / ws, tw = t - th * ws;
  const int reg_t = region[t];
  float dtau_acc = 0.f;

  // ---- pass 1: thread = query slot i -> dq_i, dbias, dtau
  // dbias: for a fixed j the threads (i) of the CTA hit distinct table entries, so the shared-memory
  // atomics see almost no same-address contention; the table is flushed to global once per CTA.
  if (active) {
    float qh[DMAX], g_i[DMAX], acc[DMAX];
#pragma unroll
    for (int c = 0; c < DMAX; ++c) {
      qh[c] = (c < d) ? Qs[t * pitch + c] : 0.f;
      g_i[c] = (c < d) ? Gs[t * pitch + c] : 0.f;
      acc[c] = 0.f;
    }
    const float lse_i = lse_s[t], D_i = D_s[t];
    int jh = 0, jw = 0;
    for (int j = 0; j < N; ++j) {
      float cosv = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < DMAX; ++c)
        if (c < d) {
          cosv = fmaf(qh[c], Ks[j * pitch + c], cosv);
          dp = fmaf(g_i[c], Vs[j * pitch + c], dp);
        }
      const int r = (th - jh + ws - 1) * (2 * ws - 1) + (tw - jw + ws - 1);
      float s = fmaf(tau_h, cosv, bias_s[r]);
      if (closed_mask) s += (region[j] != reg_t) ? kMaskValue : 0.f;
      if (mbase) s += mbase[(int64_t)t * N + j];
      const float p = expf(s - lse_i);
      const float ds = p * (dp - D_i);
      dtau_acc = fmaf(ds, cosv, dtau_acc);
      atomicAdd(&dbias_s[r], ds);
#pragma unroll
      for (int c = 0; c < DMAX; ++c)
        if (c < d) acc[c] = fmaf(ds, Ks[j * pitch + c], acc[c]);
      if (++jw == ws) { jw = 0; ++jh; }
    }
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < DMAX; ++c) {
      acc[c] *= tau_h;  // d q-hat
      dot = fmaf(qh[c], acc[c], dot);
    }
    const float rq = rq_s[t];
    T* dq = dqkv + tok[t] * qkv_stride + head * d;
#pragma unroll
    for (int c = 0; c < DMAX; ++c)
      if (c < d) dq[c] = from_f32<T>((acc[c] - qh[c] * dot) * rq);
  }

  // ---- pass 2: thread = key slot j -> dk_j, dv_j
  if (active) {
    float kh[DMAX], v_j[DMAX], dk[DMAX], dv[DMAX];
#pragma unroll
    for (int c = 0; c < DMAX; ++c) {
      kh[c] = (c < d) ? Ks[t * pitch + c] : 0.f;
      v_j[c] = (c < d) ? Vs[t * pitch + c] : 0.f;
      dk[c] = 0.f; dv[c] = 0.f;
    }
    int ih = 0, iw = 0;
    for (int i = 0; i < N; ++i) {
      float cosv = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < DMAX; ++c)
        if (c < d) {
          cosv = fmaf(Qs[i * pitch + c], kh[c], cosv);
          dp = fmaf(Gs[i * pitch + c], v_j[c], dp);
        }
      const int r = (ih - th + ws - 1) * (2 * ws - 1) + (iw - tw + ws - 1);
      float s = fmaf(tau_h, cosv, bias_s[r]);
      if (closed_mask) s += (region[i] != reg_t) ? kMaskValue : 0.f;
      if (mbase) s += mbase[(int64_t)i * N + t];
      const float p = expf(s - lse_s[i]);
      const float ds = p * (dp - D_s[i]);
#pragma unroll
      for (int c = 0; c < DMAX; ++c)
        if (c < d) {
          dv[c] = fmaf(p, Gs[i * pitch + c], dv[c]);
          dk[c] = fmaf(ds, Qs[i * pitch + c], dk[c]);
        }
      if (++iw == ws) { iw = 0; ++ih; }
    }
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < DMAX; ++c) {
      dk[c] *= tau_h;
      dot = fmaf(kh[c], dk[c], dot);
    }
    const float rk = rk_s[t];
    T* dkp = dqkv + tok[t] * qkv_stride + g.C + head * d;
    T* dvp = dkp + g.C;
#pragma unroll
    for (int c = 0; c < DMAX; ++c)
      if (c < d) {
        dkp[c] = from_f32<T>((dk[c] - kh[c] * dot) * rk);
        dvp[c] = from_f32<T>(dv[c]);
      }
  }

  // ---- d tau: block reduction, one atomic per CTA
  red_s[tid < N ? tid : 0] = 0.f;
  __syncthreads();
  if (active) red_s[t] = dtau_acc;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int k = 0; k < N; ++k) s += red_s[k];
    atomicAdd(&dtau[head], s);
  }
  for (int r = tid; r < Mtab; r += nthreads) atomicAdd(&dbias_table[r * g.heads + head], dbias_s[r]);
}

template <typename T>
int launch_fwd(const Geom& g, const void* qkv, const float* bias_table, const float* tau, const float* mask,
               int mask_windows, void* out, float* lse, cudaStream_t st) {
  const int threads = ((g.N + 31) / 32) * 32;
  const GenericSmem L = generic_layout(g.N, g.d, g.ws, 3, 1, 1);
  if (L.bytes > 227 * 1024) HV_FAIL(HV_ERR_SHAPE, "generic window attention: N=%d d=%d needs %zu B smem", g.N, g.d, L.bytes);
  const int blocks = g.B * g.nW * g.heads;
  auto run = [&](auto kern) -> int {
    HV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.bytes));
    kern<<<blocks, threads, L.bytes, st>>>((const T*)qkv, bias_table, tau, mask, mask_windows, (T*)out, lse, g);
    HV_LAUNCH_OK("wattn_generic_fwd_kernel");
    return HV_OK;
  };
  return g.d <= 32 ? run(wattn_generic_fwd_kernel<T, 32>) : run(wattn_generic_fwd_kernel<T, 64>);
}

template <typename T>
int launch_bwd(const Geom& g, const void* qkv, const void* out, const void* dout, const float* lse,
               const float* bias_table, const float* tau, const float* mask, int mask_windows, void* dqkv,
               float* dbias_table, float* dtau, cudaStream_t st) {
  const int threads = ((g.N + 31) / 32) * 32;
  const GenericSmem L = generic_layout(g.N, g.d, g.ws, 4, 5, 2);
  if (L.bytes > 227 * 1024) HV_FAIL(HV_ERR_SHAPE, "generic window attention bwd: N=%d d=%d needs %zu B smem", g.N, g.d, L.bytes);
  const int blocks = g.B * g.nW * g.heads;
  const int Mtab = (2 * g.ws - 1) * (2 * g.ws - 1);
  HV_CUDA_OK(cudaMemsetAsync(dbias_table, 0, sizeof(float) * Mtab * g.heads, st));
  HV_CUDA_OK(cudaMemsetAsync(dtau, 0, sizeof(float) * g.heads, st));
  auto run = [&](auto kern) -> int {
    HV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.bytes));
    kern<<<blocks, threads, L.bytes, st>>>((const T*)qkv, (const T*)out, (const T*)dout, lse, bias_table, tau, mask,
                                           mask_windows, (T*)dqkv, dbias_table, dtau, g);
    HV_LAUNCH_OK("wattn_generic_bwd_kernel");
    return HV_OK;
  };
  return g.d <= 32 ? run(wattn_generic_bwd_kernel<T, 32>) : run(wattn_generic_bwd_kernel<T, 64>);
}

}  // namespace

int wattn_generic_fwd(const Geom& g, int dtype, const void* qkv, const float* bias_table, const float* tau,
                      const float* mask, int mask_windows, void* out, float* lse, cudaStream_t st) {
  if (g.N > 256 || g.d > 64) HV_FAIL(HV_ERR_SHAPE, "window attention supports ws<=16 and head dim<=64 (got ws=%d d=%d)", g.ws, g.d);
  return dtype == HV_F32 ? launch_fwd<float>(g, qkv, bias_table, tau, mask, mask_windows, out, lse, st)
                         : launch_fwd<bf16>(g, qkv, bias_table, tau, mask, mask_windows, out, lse, st);
}

int wattn_generic_bwd(const Geom& g, int dtype, const void* qkv, const void* out, const void* dout, const float* lse,
                      const float* bias_table, const float* tau, const float* mask, int mask_windows, void* dqkv,
                      float* dbias_table, float* dtau, cudaStream_t st) {
  if (g.N > 256 || g.d > 64) HV_FAIL(HV_ERR_SHAPE, "window attention supports ws<=16 and head dim<=64 (got ws=%d d=%d)", g.ws, g.d);
  return dtype == HV_F32
             ? launch_bwd<float>(g, qkv, out, dout, lse, bias_table, tau, mask, mask_windows, dqkv, dbias_table, dtau, st)
             : launch_bwd<bf16>(g, qkv, out, dout, lse, bias_table, tau, mask, mask_windows, dqkv, dbias_table, dtau, st);
}

}  // namespace hv
