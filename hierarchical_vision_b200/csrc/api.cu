// C ABI of libhv_swin.so (declared in include/hv_swin.h): argument validation, dispatch, error text.
#include <stdarg.h>
#include <string.h>

#include "hv_common.cuh"

namespace hv {

// ---- kernels implemented in the other translation units ---------------------------------
int wattn_generic_fwd(const Geom& g, int dtype, const void* qkv, const float* bias_table, const float* tau,
                      const float* mask, int mask_windows, void* out, float* lse, cudaStream_t st);
int wattn_generic_bwd(const Geom& g, int dtype, const void* qkv, const void* out, const void* dout, const float* lse,
                      const float* bias_table, const float* tau, const float* mask, int mask_windows, void* dqkv,
                      float* dbias_table, float* dtau, cudaStream_t st);
bool wattn_mma64_supported(const Geom& g, int dtype);
bool wattn_tc64_supported(const Geom& g, int dtype);
int wattn_fwd_variant_set(int v);
int wattn_fwd_variant_get();
bool wattn_tc64_fwd2_supported(const Geom& g, int dtype);
int sgdw_step(const void* table, const void* chunks, int nchunks, const float* flat, const float* lr, const float* coef,
              float momentum, cudaStream_t st);
bool mlp_dgelu_gemm_supported(int64_t M, int N, int K);
size_t mlp_dgelu_gemm_workspace_bytes(int N);
int mlp_dgelu_gemm(const void* dy, const void* w2, const void* h, const float* b1, void* dh, float* db1, void* workspace,
                   size_t workspace_bytes, int64_t M, int N, int K, cudaStream_t st);
int mlp_fc1_gelu_gemm(const void* x, const void* w1, const float* b1, void* h, void* act, int64_t M, int N, int K, cudaStream_t st);
bool wattn_tc256_supported(const Geom& g, int dtype);
void window16_tile_token_index(const Geom& g, int64_t* out);
int wattn_tc256_variant_set(int v);
int wattn_tc256_fwd(const Geom& g, const void* qkv, const float* bias_table, const float* tau, void* out, float* stats,
                    cudaStream_t st);
size_t wattn_tc256_bwd_workspace_bytes(const Geom& g);
int wattn_tc256_bwd(const Geom& g, const void* qkv, const void* out, const void* dout, const float* stats,
                    const float* bias_table, const float* tau, void* dqkv, float* dbias_table, float* dtau, void* workspace,
                    size_t workspace_bytes, cudaStream_t st);
int wattn_tc64_fwd2(const Geom& g, const void* qkv, const float* bias_table, const float* tau, void* out, float* stats,
                    cudaStream_t st);
int wattn_tc64_fwd(const Geom& g, const void* qkv, const float* bias_table, const float* tau, void* out, float* lse,
                   cudaStream_t st);
size_t wattn_mma64_bwd_workspace_bytes(const Geom& g);
size_t dq_colsum_workspace_bytes(int C);
int dq_colsum(const void* dqkv, int64_t tokens, int C, float* out, void* workspace, size_t workspace_bytes, cudaStream_t st);
int wattn_mma64_heads_per_cta(int heads);
bool wattn_tc64_bwd_supported(const Geom& g, int dtype);
int wattn_bwd_variant_set(int v);
size_t wattn_tc64_bwd_workspace_bytes(const Geom& g);
int wattn_tc64_bwd(const Geom& g, const void* qkv, const void* dout, const float* stats, const float* bias_table,
                   const float* tau, void* dqkv, float* dbias_table, float* dtau, float* dq_colsum, void* workspace,
                   size_t workspace_bytes, cudaStream_t st);
int wattn_mma64_fwd(const Geom& g, const void* qkv, const float* bias_table, const float* tau, const float* mask,
                    int mask_windows, void* out, float* lse, cudaStream_t st);
int wattn_mma64_bwd(const Geom& g, const void* qkv, const void* out, const void* dout, const float* lse,
                    const float* bias_table, const float* tau, const float* mask, int mask_windows, void* dqkv,
                    float* dbias_table, float* dtau, float* dq_colsum, void* workspace, size_t workspace_bytes,
                    cudaStream_t st);
size_t ln_residual_bwd_workspace_bytes(int64_t rows, int C);
int ln_residual_fwd(const void* y, const void* shortcut, const float* gamma, const float* beta, const float* bias,
                    const float* keep_scale, void* out, float* mean, float* rstd, int64_t rows, int C,
                    int64_t rows_per_sample, float eps, int y_dtype, int res_dtype, cudaStream_t st);
int ln_residual_bwd(const void* dout, const void* y, const float* gamma, const float* bias, const float* mean,
                    const float* rstd, const float* keep_scale, void* dy, float* dgamma, float* dbeta, float* dbias,
                    void* workspace, size_t workspace_bytes, int64_t rows, int C, int64_t rows_per_sample, int y_dtype,
                    int res_dtype, cudaStream_t st);
size_t bias_gelu_bwd_workspace_bytes(int64_t rows, int cols);
int bias_gelu_fwd(const void* h, const float* bias, void* out, int64_t rows, int cols, int dtype, cudaStream_t st);
int bias_gelu_bwd(const void* dout, const void* h, const float* bias, void* dh, float* dbias, void* workspace,
                  size_t workspace_bytes, int64_t rows, int cols, int dtype, cudaStream_t st);
int patch_merge_gather_fwd(const void* x, void* out, int B, int H, int W, int C, int dtype, cudaStream_t st);
int patch_merge_gather_bwd(const void* dout, void* dx, int B, int H, int W, int C, int dtype, cudaStream_t st);
int cpb_bias_fwd(const float* coords, const float* w1, const float* b1, const float* w2, float* table, int M, int hid,
                 int heads, cudaStream_t st);
int cpb_bias_bwd(const float* coords, const float* w1, const float* b1, const float* w2, const float* dtable, float* dw1,
                 float* db1, float* dw2, float* workspace, int M, int hid, int heads, cudaStream_t st);
int ce_fwd_grad(const void* logits, const int64_t* target, float* loss_rows, void* dlogits, int64_t rows, int classes,
                float smoothing, float scale, int dtype, cudaStream_t st);
int patch_rows(const void* img, int img_dtype, const float* scale, const float* shift, void* out, int out_dtype, int B,
               int Cin, int H, int W, int P, cudaStream_t st);

// ---- error text / device info --------------------------------------------------------------
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_arch_state[64];  // 0 unknown, 1 ok, 2 bad
static int g_sm_count[64];

static int query_device(int& dev) {
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    set_error("no current CUDA device (libhv_swin has no CPU fallback)");
    return HV_ERR_CUDA;
  }
  if (g_arch_state[dev] == 0) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
      set_error("cudaGetDeviceProperties failed");
      return HV_ERR_CUDA;
    }
    g_sm_count[dev] = p.multiProcessorCount;
    g_arch_state[dev] = (p.major == 10 && p.minor == 0) ? 1 : 2;
  }
  return HV_OK;
}

int check_device_arch() {
  int dev;
  int rc = query_device(dev);
  if (rc) return rc;
  if (g_arch_state[dev] != 1) {
    set_error("libhv_swin is built for sm_100a (B200) only; current device is not compute capability 10.0");
    return HV_ERR_ARCH;
  }
  return HV_OK;
}

int num_sms() {
  int dev;
  if (query_device(dev)) return 148;
  return g_sm_count[dev] > 0 ? g_sm_count[dev] : 148;
}

static int make_checked_geom(int B, int H, int W, int C, int heads, int ws, int shift, Geom& g) {
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || heads <= 0 || ws <= 0)
    HV_FAIL(HV_ERR_SHAPE, "window_attn: non-positive size (B=%d H=%d W=%d C=%d heads=%d ws=%d)", B, H, W, C, heads, ws);
  if (C % heads != 0) HV_FAIL(HV_ERR_SHAPE, "window_attn: C=%d not divisible by heads=%d", C, heads);
  if (H % ws != 0 || W % ws != 0) HV_FAIL(HV_ERR_SHAPE, "window_attn: resolution %dx%d not divisible by window %d", H, W, ws);
  if (shift < 0 || shift >= ws) HV_FAIL(HV_ERR_SHAPE, "shift_size must in 0-window_size (shift=%d ws=%d)", shift, ws);
  g = make_geom(B, H, W, C, heads, ws, shift);
  return HV_OK;
}

}  // namespace hv

using namespace hv;

extern "C" {

int hv_abi_version(void) { return HV_ABI_VERSION; }
const char* hv_last_error(void) { return g_err; }
int hv_compiled_arch(void) { return 100; }

int hv_window_attn_fwd_variant(int variant) {
  if (variant < -1 || variant > 2) HV_FAIL(HV_ERR_SHAPE, "hv_window_attn_fwd_variant: variant %d", variant);
  wattn_fwd_variant_set(variant);
  return HV_OK;
}

int hv_window_attn_bwd_variant(int variant) {
  if (variant < -1 || variant > 1) HV_FAIL(HV_ERR_SHAPE, "hv_window_attn_bwd_variant: variant %d", variant);
  wattn_bwd_variant_set(variant);
  return HV_OK;
}

int hv_window_attn_tc256_variant(int variant) {
  if (variant < -1 || variant > 1) HV_FAIL(HV_ERR_SHAPE, "hv_window_attn_tc256_variant: variant %d", variant);
  wattn_tc256_variant_set(variant);
  return HV_OK;
}

int hv_window_attn_kernel_kind(int C, int heads, int ws, int dtype) {
  if (heads <= 0 || C % heads) return 0;
  Geom g = make_geom(1, ws, ws, C, heads, ws, 0);
  return (wattn_mma64_supported(g, dtype) || wattn_tc256_supported(g, dtype)) ? 1 : 0;
}

size_t hv_window_attn_stats_floats(int B, int H, int W, int C, int heads, int ws, int dtype) {
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || heads <= 0 || ws <= 0 || C % heads || H % ws || W % ws) return 0;
  const Geom g = make_geom(B, H, W, C, heads, ws, 0);
  const size_t plane = (size_t)g.B * g.nW * g.heads * g.N;
  return (wattn_mma64_supported(g, dtype) || wattn_tc256_supported(g, dtype)) ? 3 * plane : plane;
}

int hv_window_attn_kernel_name(int B, int H, int W, int C, int heads, int ws, int shift, int dtype, int backward, char* out,
                               int out_len) {
  if (!out || out_len <= 0) HV_FAIL(HV_ERR_NULL, "hv_window_attn_kernel_name: out is NULL");
  Geom g;
  int rc = make_checked_geom(B, H, W, C, heads, ws, shift, g);
  if (rc) return rc;
  // the same decisions as hv_window_attn_fwd / _bwd (mask == NULL)
  if (wattn_tc256_supported(g, dtype))
    snprintf(out, out_len, "wattn_tc256_%s_kernel", backward ? "bwd" : "fwd");
  else if (!wattn_mma64_supported(g, dtype))
    snprintf(out, out_len, "wattn_generic_%s_kernel<%s>", backward ? "bwd" : "fwd", dtype == HV_BF16 ? "bf16" : "float");
  else if (!backward && wattn_tc64_supported(g, dtype) && wattn_fwd_variant_get() != 2 && wattn_tc64_fwd2_supported(g, dtype))
    snprintf(out, out_len, "wattn_tc64_fwd2_kernel<%s>", shift > 0 ? "true" : "false");
  else if (!backward && wattn_tc64_supported(g, dtype))
    snprintf(out, out_len, "wattn_tc64_fwd_kernel");
  else if (backward && wattn_tc64_bwd_supported(g, dtype))
    snprintf(out, out_len, "wattn_tc64_bwd_kernel<%s>", shift > 0 ? "true" : "false");
  else
    snprintf(out, out_len, "wattn_mma64_%s_kernel<%d>", backward ? "bwd" : "fwd", wattn_mma64_heads_per_cta(heads));
  return HV_OK;
}

int hv_relative_position_index(int ws, int64_t* out) {
  if (!out) HV_FAIL(HV_ERR_NULL, "hv_relative_position_index: out is NULL");
  if (ws <= 0) HV_FAIL(HV_ERR_SHAPE, "ws=%d", ws);
  const int N = ws * ws;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) out[(size_t)i * N + j] = rel_pos_index(ws, i, j);
  return HV_OK;
}

int hv_shift_window_mask(int H, int W, int ws, int shift, float* out) {
  if (!out) HV_FAIL(HV_ERR_NULL, "hv_shift_window_mask: out is NULL");
  Geom g;
  int rc = make_checked_geom(1, H, W, ws, 1, ws, shift, g);
  if (rc) return rc;
  if (shift == 0) HV_FAIL(HV_ERR_SHAPE, "hv_shift_window_mask: shift must be > 0 (the reference has attn_mask=None otherwise)");
  const int N = g.N;
  for (int win = 0; win < g.nW; ++win)
    for (int i = 0; i < N; ++i) {
      const int ri = window_slot_region(g, win, i);
      for (int j = 0; j < N; ++j)
        out[((size_t)win * N + i) * N + j] = (window_slot_region(g, win, j) != ri) ? -100.0f : 0.0f;
    }
  return HV_OK;
}

int hv_window_token_index(int B, int H, int W, int ws, int shift, int64_t* out) {
  if (!out) HV_FAIL(HV_ERR_NULL, "hv_window_token_index: out is NULL");
  Geom g;
  int rc = make_checked_geom(B, H, W, ws, 1, ws, shift, g);
  if (rc) return rc;
  for (int b = 0; b < B; ++b)
    for (int win = 0; win < g.nW; ++win)
      for (int s = 0; s < g.N; ++s) out[((size_t)b * g.nW + win) * g.N + s] = window_slot_to_token(g, b, win, s);
  return HV_OK;
}

int hv_window16_tile_token_index(int B, int H, int W, int shift, int64_t* out) {
  if (!out) HV_FAIL(HV_ERR_NULL, "hv_window16_tile_token_index: out is NULL");
  Geom g;
  int rc = make_checked_geom(B, H, W, 32, 1, 16, shift, g);
  if (rc) return rc;
  if (shift != 0 && shift != 8) HV_FAIL(HV_ERR_SHAPE, "hv_window16_tile_token_index: shift %d (the N = 256 kernels take 0 or 8)", shift);
  window16_tile_token_index(g, out);
  return HV_OK;
}

int hv_merge_token_index(int B, int H, int W, int64_t* out) {
  if (!out) HV_FAIL(HV_ERR_NULL, "hv_merge_token_index: out is NULL");
  if (B <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) HV_FAIL(HV_ERR_SHAPE, "x size (%d*%d) are not even.", H, W);
  size_t o = 0;
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < H / 2; ++i)
      for (int j = 0; j < W / 2; ++j)
        for (int m = 0; m < 4; ++m) out[o++] = merge_src_token(H, W, b, i, j, m);
  return HV_OK;
}

static int attn_common_checks(const char* what, int dtype, const float* mask, int mask_windows, const Geom& g) {
  if (dtype != HV_F32 && dtype != HV_BF16) HV_FAIL(HV_ERR_DTYPE, "%s: dtype %d", what, dtype);
  if (mask != nullptr) {
    if (mask_windows <= 0 || (g.B * g.nW) % mask_windows != 0)
      HV_FAIL(HV_ERR_SHAPE, "%s: %d windows not divisible by mask_windows=%d", what, g.B * g.nW, mask_windows);
  }
  return check_device_arch();
}

int hv_window_attn_fwd(const void* qkv, const float* bias_table, const float* tau, const float* mask, int mask_windows,
                       void* out, float* lse, int B, int H, int W, int C, int heads, int ws, int shift, int dtype,
                       void* stream) {
  if (!qkv || !bias_table || !tau || !out || !lse) HV_FAIL(HV_ERR_NULL, "hv_window_attn_fwd: NULL argument");
  Geom g;
  int rc = make_checked_geom(B, H, W, C, heads, ws, shift, g);
  if (rc) return rc;
  rc = attn_common_checks("hv_window_attn_fwd", dtype, mask, mask_windows, g);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // tensor-core kernels (bf16, 8x8 window, head dim 32, in-kernel shift mask): tcgen05/TMEM/TMA forward when the
  // shift is even, otherwise the mma.sync one; both write lse in log2 units for wattn_mma64_bwd
  if (mask == nullptr && wattn_mma64_supported(g, dtype)) {
    if (wattn_tc64_supported(g, dtype)) {
      if (wattn_fwd_variant_get() != 2 && wattn_tc64_fwd2_supported(g, dtype))
        return wattn_tc64_fwd2(g, qkv, bias_table, tau, out, lse, st);
      return wattn_tc64_fwd(g, qkv, bias_table, tau, out, lse, st);  // first-generation kernel: any even shift
    }
    return wattn_mma64_fwd(g, qkv, bias_table, tau, mask, mask_windows, out, lse, st);
  }
  // 16 x 16 windows, head dim 32, bf16 (SwinV2-B): tcgen05 / TMEM / TMA kernels, statistics in tile order for wattn_tc256_bwd
  if (mask == nullptr && wattn_tc256_supported(g, dtype)) return wattn_tc256_fwd(g, qkv, bias_table, tau, out, lse, st);
  return wattn_generic_fwd(g, dtype, qkv, bias_table, tau, mask, mask_windows, out, lse, st);
}

size_t hv_window_attn_bwd_workspace_bytes(int B, int H, int W, int C, int heads, int ws, int dtype) {
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || heads <= 0 || ws <= 0 || C % heads || H % ws || W % ws) return 0;
  Geom g = make_geom(B, H, W, C, heads, ws, 0);
  if (wattn_mma64_supported(g, dtype)) {  // either tensor-core backward may be selected at call time
    const size_t a = wattn_mma64_bwd_workspace_bytes(g), b = wattn_tc64_bwd_workspace_bytes(g);
    return a > b ? a : b;
  }
  if (wattn_tc256_supported(g, dtype)) return wattn_tc256_bwd_workspace_bytes(g);
  return 16;  // the generic kernel reduces with atomics
}

int hv_window_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const float* bias_table,
                       const float* tau, const float* mask, int mask_windows, void* dqkv, float* dbias_table, float* dtau,
                       float* dq_colsum, void* workspace, size_t workspace_bytes, int B, int H, int W, int C, int heads,
                       int ws, int shift, int dtype, void* stream) {
  if (!qkv || !out || !dout || !lse || !bias_table || !tau || !dqkv || !dbias_table || !dtau)
    HV_FAIL(HV_ERR_NULL, "hv_window_attn_bwd: NULL argument");
  Geom g;
  int rc = make_checked_geom(B, H, W, C, heads, ws, shift, g);
  if (rc) return rc;
  rc = attn_common_checks("hv_window_attn_bwd", dtype, mask, mask_windows, g);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mask == nullptr && wattn_mma64_supported(g, dtype)) {
    if (wattn_tc64_bwd_supported(g, dtype))
      return wattn_tc64_bwd(g, qkv, dout, lse, bias_table, tau, dqkv, dbias_table, dtau, dq_colsum, workspace,
                            workspace_bytes, st);
    return wattn_mma64_bwd(g, qkv, out, dout, lse, bias_table, tau, mask, mask_windows, dqkv, dbias_table, dtau, dq_colsum,
                           workspace, workspace_bytes, st);
  }
  if (mask == nullptr && dq_colsum == nullptr && wattn_tc256_supported(g, dtype))
    return wattn_tc256_bwd(g, qkv, out, dout, lse, bias_table, tau, dqkv, dbias_table, dtau, workspace, workspace_bytes, st);
  if (dq_colsum != nullptr)
    HV_FAIL(HV_ERR_SHAPE, "hv_window_attn_bwd: dq_colsum is only produced by the tensor-core kernel "
                          "(hv_window_attn_kernel_kind() == 1 and mask == NULL)");
  return wattn_generic_bwd(g, dtype, qkv, out, dout, lse, bias_table, tau, mask, mask_windows, dqkv, dbias_table, dtau, st);
}

size_t hv_dq_colsum_workspace_bytes(int C) { return C > 0 ? dq_colsum_workspace_bytes(C) : 0; }

int hv_dq_colsum(const void* dqkv, float* dq_colsum_out, void* workspace, size_t workspace_bytes, int64_t tokens, int C,
                 int dtype, void* stream) {
  if (!dqkv || !dq_colsum_out) HV_FAIL(HV_ERR_NULL, "hv_dq_colsum: NULL argument");
  if (dtype != HV_BF16) HV_FAIL(HV_ERR_DTYPE, "hv_dq_colsum: bf16 only (the tensor-core attention path)");
  if (tokens <= 0 || C <= 0) HV_FAIL(HV_ERR_SHAPE, "hv_dq_colsum: tokens=%lld C=%d", (long long)tokens, C);
  int rc = check_device_arch();
  if (rc) return rc;
  return dq_colsum(dqkv, tokens, C, dq_colsum_out, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int hv_ln_residual_fwd(const void* y, const void* shortcut, const float* gamma, const float* beta, const float* bias,
                       const float* keep_scale, void* out, float* mean, float* rstd, int64_t rows, int C,
                       int64_t rows_per_sample, float eps, int y_dtype, int res_dtype, void* stream) {
  if (!y || !gamma || !beta || !out || !mean || !rstd) HV_FAIL(HV_ERR_NULL, "hv_ln_residual_fwd: NULL argument");
  int rc = check_device_arch();
  if (rc) return rc;
  return ln_residual_fwd(y, shortcut, gamma, beta, bias, keep_scale, out, mean, rstd, rows, C, rows_per_sample, eps,
                         y_dtype, res_dtype, static_cast<cudaStream_t>(stream));
}

size_t hv_ln_residual_bwd_workspace_bytes(int64_t rows, int C) { return ln_residual_bwd_workspace_bytes(rows, C); }

int hv_ln_residual_bwd(const void* dout, const void* y, const float* gamma, const float* bias, const float* mean,
                       const float* rstd, const float* keep_scale, void* dy, float* dgamma, float* dbeta, float* dbias,
                       void* workspace, size_t workspace_bytes, int64_t rows, int C, int64_t rows_per_sample, int y_dtype,
                       int res_dtype, void* stream) {
  if (!dout || !y || !gamma || !mean || !rstd || !dy || !dgamma || !dbeta) HV_FAIL(HV_ERR_NULL, "hv_ln_residual_bwd: NULL argument");
  int rc = check_device_arch();
  if (rc) return rc;
  return ln_residual_bwd(dout, y, gamma, bias, mean, rstd, keep_scale, dy, dgamma, dbeta, dbias, workspace, workspace_bytes,
                         rows, C, rows_per_sample, y_dtype, res_dtype, static_cast<cudaStream_t>(stream));
}

int hv_sgdw_step(const void* table, const void* chunks, int nchunks, const float* flat_grad, const float* lr, const float* clip_coef,
                 float momentum, void* stream) {
  if (!table || !chunks || !flat_grad || !lr) HV_FAIL(HV_ERR_NULL, "hv_sgdw_step: NULL argument");
  if (nchunks < 0) HV_FAIL(HV_ERR_SHAPE, "hv_sgdw_step: nchunks=%d", nchunks);
  int rc = check_device_arch();
  if (rc) return rc;
  return sgdw_step(table, chunks, nchunks, flat_grad, lr, clip_coef, momentum, static_cast<cudaStream_t>(stream));
}

size_t hv_mlp_dgelu_gemm_workspace_bytes(int64_t rows, int hidden, int C) {
  return mlp_dgelu_gemm_supported(rows, hidden, C) ? mlp_dgelu_gemm_workspace_bytes(hidden) : 0;
}

int hv_mlp_dgelu_gemm(const void* dy, const void* w2, const void* h, const float* b1, void* dh, float* db1, void* workspace,
                      size_t workspace_bytes, int64_t rows, int hidden, int C, int dtype, void* stream) {
  if (!dy || !w2 || !h || !b1 || !dh || !db1) HV_FAIL(HV_ERR_NULL, "hv_mlp_dgelu_gemm: NULL argument");
  if (dtype != HV_BF16) HV_FAIL(HV_ERR_DTYPE, "hv_mlp_dgelu_gemm: bf16 activations only");
  int rc = check_device_arch();
  if (rc) return rc;
  return mlp_dgelu_gemm(dy, w2, h, b1, dh, db1, workspace, workspace_bytes, rows, hidden, C, static_cast<cudaStream_t>(stream));
}

int hv_mlp_fc1_gelu_gemm(const void* x, const void* w1, const float* b1, void* h, void* act, int64_t rows, int hidden, int C,
                         int dtype, void* stream) {
  if (!x || !w1 || !b1 || !h || !act) HV_FAIL(HV_ERR_NULL, "hv_mlp_fc1_gelu_gemm: NULL argument");
  if (dtype != HV_BF16) HV_FAIL(HV_ERR_DTYPE, "hv_mlp_fc1_gelu_gemm: bf16 activations only");
  int rc = check_device_arch();
  if (rc) return rc;
  return mlp_fc1_gelu_gemm(x, w1, b1, h, act, rows, hidden, C, static_cast<cudaStream_t>(stream));
}

int hv_bias_gelu_fwd(const void* h, const float* bias, void* out, int64_t rows, int cols, int dtype, void* stream) {
  if (!h || !bias || !out) HV_FAIL(HV_ERR_NULL, "hv_bias_gelu_fwd: NULL argument");
  int rc = check_device_arch();
  if (rc) return rc;
  return bias_gelu_fwd(h, bias, out, rows, cols, dtype, static_cast<cudaStream_t>(stream));
}

size_t hv_bias_gelu_bwd_workspace_bytes(int64_t rows, int cols) { return bias_gelu_bwd_workspace_bytes(rows, cols); }

int hv_bias_gelu_bwd(const void* dout, const void* h, const float* bias, void* dh, float* dbias, void* workspace,
                     size_t workspace_bytes, int64_t rows, int cols, int dtype, void* stream) {
  if (!dout || !h || !bias || !dh || !dbias) HV_FAIL(HV_ERR_NULL, "hv_bias_gelu_bwd: NULL argument");
  int rc = check_device_arch();
  if (rc) return rc;
  return bias_gelu_bwd(dout, h, bias, dh, dbias, workspace, workspace_bytes, rows, cols, dtype,
                       static_cast<cudaStream_t>(stream));
}

int hv_patch_merge_gather_fwd(const void* x, void* out, int B, int H, int W, int C, int dtype, void* stream) {
  if (!x || !out) HV_FAIL(HV_ERR_NULL, "hv_patch_merge_gather_fwd: NULL argument");
  int rc = check_device_arch();
  if (rc) return rc;
  return patch_merge_gather_fwd(x, out, B, H, W, C, dtype, static_cast<cudaStream_t>(stream));
}

int hv_patch_merge_gather_bwd(const void* dout, void* dx, int B, int H, int W, int C, int dtype, void* stream) {
  if (!dout || !dx) HV_FAIL(HV_ERR_NULL, "hv_patch_merge_gather_bwd: NULL argument");
  int rc = check_device_arch();
  if (rc) return rc;
  return patch_merge_gather_bwd(dout, dx, B, H, W, C, dtype, static_cast<cudaStream_t>(stream));
}

int hv_patch_rows(const void* img, int img_dtype, const float* scale, const float* shift, void* out, int out_dtype,
                  int B, int Cin, int H, int W, int P, void* stream) {
  if (!img || !out) HV_FAIL(HV_ERR_NULL, "hv_patch_rows: NULL argument");
  if ((scale == nullptr) != (shift == nullptr)) HV_FAIL(HV_ERR_NULL, "hv_patch_rows: scale and shift go together");
  int rc = check_device_arch();
  if (rc) return rc;
  return patch_rows(img, img_dtype, scale, shift, out, out_dtype, B, Cin, H, W, P, static_cast<cudaStream_t>(stream));
}

int hv_cross_entropy_fwd_grad(const void* logits, const int64_t* target, float* loss_rows, void* dlogits, int64_t rows,
                              int classes, float smoothing, float scale, int dtype, void* stream) {
  if (!logits || !target || !loss_rows || !dlogits) HV_FAIL(HV_ERR_NULL, "hv_cross_entropy_fwd_grad: NULL argument");
  int rc = check_device_arch();
  if (rc) return rc;
  return ce_fwd_grad(logits, target, loss_rows, dlogits, rows, classes, smoothing, scale, dtype, static_cast<cudaStream_t>(stream));
}

int hv_cpb_bias_fwd(const float* coords, const float* w1, const float* b1, const float* w2, float* table, int M, int hidden,
                    int heads, void* stream) {
  if (!coords || !w1 || !b1 || !w2 || !table) HV_FAIL(HV_ERR_NULL, "hv_cpb_bias_fwd: NULL argument");
  int rc = check_device_arch();
  if (rc) return rc;
  return cpb_bias_fwd(coords, w1, b1, w2, table, M, hidden, heads, static_cast<cudaStream_t>(stream));
}

int hv_cpb_bias_bwd(const float* coords, const float* w1, const float* b1, const float* w2, const float* dtable, float* dw1,
                    float* db1, float* dw2, float* workspace, int M, int hidden, int heads, void* stream) {
  if (!coords || !w1 || !b1 || !w2 || !dtable || !dw1 || !db1 || !dw2 || !workspace)
    HV_FAIL(HV_ERR_NULL, "hv_cpb_bias_bwd: NULL argument");
  int rc = check_device_arch();
  if (rc) return rc;
  return cpb_bias_bwd(coords, w1, b1, w2, dtable, dw1, db1, dw2, workspace, M, hidden, heads, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
