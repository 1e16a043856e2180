/* hv_swin.h -- C ABI of libhv_swin.so: the sm_100a (B200) implementation of the SwinV2
 * windowed-attention hot path of samuelstevens/hierarchical-vision.
 *
 * The reference implements this path as PyTorch nn.Modules in `swinv2.py`; the drop-in
 * boundary is those classes (hierarchical_vision_b200/swinv2.py keeps their names, ctor
 * arguments, forward signatures and state_dict keys).  This header is what sits underneath:
 * plain pointers and sizes, no torch types.  Each entry point names the reference lines
 * (file:line in /root/reference) whose work it replaces.
 *
 * Conventions
 *   - every pointer marked "device" is CUDA device memory owned by the caller; the library
 *     never allocates, frees or keeps a pointer after the call returns;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); calls only enqueue;
 *   - return value 0 = success, otherwise an HV_ERR_* code; hv_last_error() gives the message
 *     of the last failing call on the calling thread.  No exceptions cross the boundary;
 *   - there is no CPU fallback and no other GPU backend: on anything but sm_100 the device
 *     entry points return HV_ERR_ARCH;
 *   - dtype codes describe the activation tensors; small parameter tensors (bias table, tau,
 *     LayerNorm gamma/beta, statistics, gradients of those) are always float32.
 */
#ifndef HV_SWIN_H_
#define HV_SWIN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HV_ABI_VERSION 4

#if defined(__GNUC__)
#define HV_API __attribute__((visibility("default")))
#else
#define HV_API
#endif

enum { HV_F32 = 0, HV_BF16 = 1, HV_U8 = 2 /* images only */ };

enum {
  HV_OK = 0,
  HV_ERR_SHAPE = 1,     /* unsupported / inconsistent sizes                         */
  HV_ERR_ALIGN = 2,     /* pointer or row pitch not 16-byte aligned                 */
  HV_ERR_DTYPE = 3,     /* unknown dtype code or unsupported combination            */
  HV_ERR_WORKSPACE = 4, /* workspace missing or smaller than hv_*_workspace_bytes() */
  HV_ERR_ARCH = 5,      /* device is not sm_100 (B200)                              */
  HV_ERR_CUDA = 6,      /* a CUDA runtime call or launch failed                     */
  HV_ERR_NULL = 7       /* required pointer is NULL                                 */
};

/* ---- library info -------------------------------------------------------------------- */
HV_API int hv_abi_version(void);
HV_API const char* hv_last_error(void);
/* 100 when built for sm_100a */
HV_API int hv_compiled_arch(void);
/* Which attention kernel family a geometry dispatches to: 0 = generic CUDA-core kernel, 1 = tensor-core kernels (bf16,
 * head dim 32, window 8 (N = 64) or window 16 (N = 256)).  For kind 1 the module wraps the qkv Linear and the attention in
 * one autograd node and takes d(q_bias) from hv_dq_colsum.  Host-only query. */
HV_API int hv_window_attn_kernel_kind(int C, int heads, int ws, int dtype);

/* Forward kernel of the tensor-core path (kind 1): 0 = mma.sync + cp.async kernel, 1 = tcgen05 / TMEM / TMA kernels (the
 * second-generation kernel with TMA stores for shift 0 or ws / 2, the first-generation one for other even shifts; odd
 * shifts fall back to 0), 2 = the first-generation tcgen05 kernel wherever valid, -1 = decided by the HV_ATTN_TCGEN05
 * environment variable (0 / 1 as above; unset = 1).  All write the same outputs and statistics.  A test /
 * benchmarking switch: process-wide, read at launch time, not meant to be flipped while other threads launch. */
HV_API int hv_window_attn_fwd_variant(int variant);
/* Backward kernel of the tensor-core path: 0 = mma.sync + cp.async kernel, 1 = tcgen05 / TMEM / TMA kernel (shift 0 or
 * ws / 2; falls back to 0 otherwise), -1 = decided by the HV_ATTN_TCGEN05_BWD environment variable (unset = automatic:
 * the tcgen05 kernel wherever it is valid).  Same outputs and workspace; process-wide test / benchmarking switch. */
HV_API int hv_window_attn_bwd_variant(int variant);
/* 16 x 16 windows with head dim 32 in bf16 (SwinV2-B at window 16, reference swinv2.py:105-283 with window_size = 16,
 * shift 0 or 8): 1 = tcgen05 / TMEM / TMA kernels wattn_tc256_{fwd,bwd}_kernel, 0 = the generic CUDA-core kernels,
 * -1 = decided by the HV_ATTN_TC256 environment variable (unset = 1).  Forward and backward of one call pair must run
 * under the same setting (the statistics layouts differ).  Process-wide test / benchmarking switch. */
HV_API int hv_window_attn_tc256_variant(int variant);
/* Number of float32 elements of the `lse` statistics buffer of hv_window_attn_fwd / _bwd for a geometry: B*nW*heads*N
 * for the generic kernel; three such planes for the tensor-core kernels (row log-sum-exp | r_i = 1 / |q_i| |
 * c_j = tau log2(e) / |k_j|, each (B*nW, heads, N); window-slot order for N = 64, the kernels' tile order for N = 256).  0 on invalid sizes.  Host-only query. */
HV_API size_t hv_window_attn_stats_floats(int B, int H, int W, int C, int heads, int ws, int dtype);

/* Name of the kernel hv_window_attn_fwd (backward = 0) / hv_window_attn_bwd (backward = 1) launches for a geometry with
 * mask == NULL under the current variant settings, e.g. "wattn_tc64_bwd_kernel<true>" (tcgen05 / TMEM / TMA, shifted
 * layer), "wattn_mma64_fwd_kernel<3>", "wattn_generic_bwd_kernel<float>".  For measurement records (bench.py names the
 * kernel its roofline line is about from the dispatch, not from a literal).  Host-only query. */
HV_API int hv_window_attn_kernel_name(int B, int H, int W, int C, int heads, int ws, int shift, int dtype, int backward,
                                      char* out, int out_len);

/* ---- host-side integer maps (CPU; same arithmetic the kernels use on the device) ------ */
/* relative_position_index (N,N) int64 -- reference swinv2.py:175-190 */
HV_API int hv_relative_position_index(int ws, int64_t* out);
/* attn_mask (nW,N,N) float32 in {0,-100} -- reference swinv2.py:357-388 (requires shift > 0) */
HV_API int hv_shift_window_mask(int H, int W, int ws, int shift, float* out);
/* token index feeding window row r, slot i: roll(-shift) + window_partition, (B*nW, N) int64
 * -- reference swinv2.py:69-83, 399-412; also the scatter map of window_reverse + roll(+shift)
 * (swinv2.py:86-102, 420-429) */
HV_API int hv_window_token_index(int B, int H, int W, int ws, int shift, int64_t* out);
/* Host-only: image token index (b * H * W + row * W + col) feeding tile row t of window `win` of image b in the TILE order
 * of the 16 x 16-window kernels (two column parts of 8 x 16 tokens: t = 128 part + 8 ih + iw % 8), out[(b * nW + win) * 256 + t]:
 * the same arithmetic the kernels' TMA box coordinates come from, for bit-exact checks against torch.roll +
 * window_partition (reference swinv2.py:69-83, 399-412).  shift 0 or 8. */
HV_API int hv_window16_tile_token_index(int B, int H, int W, int shift, int64_t* out);

/* PatchMerging concat sources (B*H/2*W/2, 4) int64 -- reference swinv2.py:486-490 */
HV_API int hv_merge_token_index(int B, int H, int W, int64_t* out);

/* ---- fused shifted-window scaled-cosine attention --------------------------------------
 * Replaces, in SwinTransformerBlock.forward / WindowAttention.forward:
 *   torch.roll + window_partition            swinv2.py:399-412
 *   head split, F.normalize(q) @ F.normalize(k)^T, * exp(clamped logit_scale)   :221-231
 *   + 16*sigmoid(cpb table)[relative_position_index]                            :236-247
 *   + shifted-window mask, softmax                                              :249-257
 *   attn @ v, head merge                                                        :261
 *   window_reverse + torch.roll back                                            :420-429
 * and the autograd of all of it.  The per-token linears (qkv, proj) stay outside.
 *
 *   qkv   device, (B, H*W, 3C) `dtype`, image token order, channel = t*C + head*d + j (t=q,k,v)
 *   bias_table device float32 ((2ws-1)^2, heads) = 16*sigmoid(cpb_mlp(relative_coords_table))
 *   tau   device float32 (heads) = exp(min(logit_scale, log 100))
 *   mask  NULL -> the shift mask is generated in-kernel from (H, W, ws, shift);
 *         else device float32 (mask_windows, N, N) added to window (row % mask_windows) and the
 *         in-kernel shift mask is NOT applied (WindowAttention.forward(x, mask) semantics)
 *   out   device, (B, H*W, C) `dtype`, image token order
 *   lse   device float32, hv_window_attn_stats_floats() elements: row statistics saved for hv_window_attn_bwd.  Opaque
 *         to the caller: (B*nW, heads, N) row log-sum-exp in natural-log units (generic kernel); for the tensor-core
 *         kernels the log2-unit log-sum-exp followed by two more planes with the row scales of the cosine logits
 *         (1 / |q_i| and tau log2(e) / |k_j|), which the backward kernel reads back instead of recomputing them.  The
 *         pair fwd/bwd of one geometry always dispatches to the same kernel kind.
 */
HV_API int hv_window_attn_fwd(const void* qkv, const float* bias_table, const float* tau, const float* mask,
                       int mask_windows, void* out, float* lse, int B, int H, int W, int C, int heads,
                       int ws, int shift, int dtype, void* stream);

/* Workspace (bytes) hv_window_attn_bwd needs for its partial reductions. */
HV_API size_t hv_window_attn_bwd_workspace_bytes(int B, int H, int W, int C, int heads, int ws, int dtype);

/*   out   device (B, H*W, C) `dtype`: the forward's output (read by the generic and mma.sync kernels; the tcgen05
 *         kernel forms D = rowsum(P o dP) itself and never touches it)
 *   dout  device (B, H*W, C) `dtype`: gradient of `out`
 *   dqkv  device (B, H*W, 3C) `dtype`: fully overwritten
 *   dbias_table device float32 ((2ws-1)^2, heads): overwritten with d loss / d bias_table
 *   dtau  device float32 (heads): overwritten with d loss / d tau
 *   dq_colsum  NULL, or device float32 (C): overwritten with the sum over all tokens of the q third of dqkv,
 *         i.e. the gradient of WindowAttention.q_bias (swinv2.py:193-195, 211-220), so that no separate column
 *         reduction over dqkv is needed (v_bias needs none: softmax rows sum to one, so v_bias passes through the
 *         attention as a plain additive term of `out`; k has no bias).  Only the tensor-core kernel
 *         (hv_window_attn_kernel_kind() == 1, mask == NULL) produces it; otherwise it must be NULL.
 */
HV_API int hv_window_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                       const float* bias_table, const float* tau, const float* mask, int mask_windows,
                       void* dqkv, float* dbias_table, float* dtau, float* dq_colsum, void* workspace,
                       size_t workspace_bytes, int B, int H, int W, int C, int heads, int ws, int shift, int dtype,
                       void* stream);

/* Gradient of WindowAttention.q_bias (swinv2.py:193-195, 211-220) as its own streaming pass: dq_colsum[c] = sum over all
 * tokens of dqkv[token, c], c < C, for dqkv (tokens, 3C) bf16.  The same numbers hv_window_attn_bwd writes when its
 * dq_colsum argument is not NULL (it calls this); separate so that a caller can time / schedule it apart from the
 * attention kernel.  HBM-bound: tokens * C * 2 bytes. */
HV_API size_t hv_dq_colsum_workspace_bytes(int C);
HV_API int hv_dq_colsum(const void* dqkv, float* dq_colsum, void* workspace, size_t workspace_bytes, int64_t tokens, int C,
                        int dtype, void* stream);

/* ---- res-post-norm: out = shortcut + keep_scale[sample] * LayerNorm(y + bias) ------------
 * Replaces `shortcut + drop_path(norm1(x))` / `x + drop_path(norm2(mlp(x)))`, swinv2.py:431, 434,
 * and (shortcut == NULL) the plain LayerNorm of PatchMerging / PatchEmbed, swinv2.py:494, 656.
 * `bias` (optional) is the bias of the Linear that produced y (attn.proj.bias swinv2.py:262,
 * mlp.fc2.bias swinv2.py:64): the caller runs that Linear without bias and hands the bias here, so
 * that its gradient is produced by hv_ln_residual_bwd instead of a separate column reduction.
 *   y        device (rows, C) y_dtype          shortcut  device (rows, C) res_dtype or NULL
 *   gamma, beta  device float32 (C)            bias      device float32 (C) or NULL
 *   keep_scale   device float32 (rows / rows_per_sample) or NULL (DropPath factor per sample)
 *   out      device (rows, C) res_dtype        mean, rstd device float32 (rows), saved for backward
 */
HV_API int hv_ln_residual_fwd(const void* y, const void* shortcut, const float* gamma, const float* beta,
                       const float* bias, const float* keep_scale, void* out, float* mean, float* rstd,
                       int64_t rows, int C, int64_t rows_per_sample, float eps, int y_dtype, int res_dtype,
                       void* stream);

HV_API size_t hv_ln_residual_bwd_workspace_bytes(int64_t rows, int C);

/*   dout device (rows, C) res_dtype (this is also d shortcut)   dy device (rows, C) y_dtype
 *   dgamma, dbeta device float32 (C), overwritten; dbias device float32 (C) or NULL (= column sums of dy) */
HV_API int hv_ln_residual_bwd(const void* dout, const void* y, const float* gamma, const float* bias,
                       const float* mean, const float* rstd, const float* keep_scale, void* dy, float* dgamma,
                       float* dbeta, float* dbias, void* workspace, size_t workspace_bytes, int64_t rows, int C,
                       int64_t rows_per_sample, int y_dtype, int res_dtype, void* stream);

/* One-pass DecoupledSGDW step over all parameters (reference optim.py:16-44, configs.py:45; Composer's rule):
 * g' = g * clip_coef; buf = momentum * buf + g'; p = p * (1 - lr * wd / lr0) - lr * buf, bit-identical to the multi-tensor
 * path of train.FlatSGD.  `table`: device array of 32-byte records {float* p; float* buf; int64 grad_offset; int32 numel;
 * float wd_over_lr0}; `chunks`: device array of int32 pairs (record index, element offset), one per 4096 elements of a
 * tensor; flat_grad: the flat fp32 gradient buffer; lr, clip_coef (may be NULL = 1): device scalars. */
HV_API int hv_sgdw_step(const void* table, const void* chunks, int nchunks, const float* flat_grad, const float* lr,
                        const float* clip_coef, float momentum, void* stream);

/* Backward of the Mlp hidden activation fused into the GEMM that produces its upstream gradient (reference swinv2.py:43-66
 * differentiated): dh = (dy W2) * GELU'(h + b1) and db1 = column sums of dh, with dy (rows, C), w2 = fc2.weight (C, hidden)
 * as stored, h = x W1^T (rows, hidden) without the fc1 bias, all bf16; b1, db1 fp32 (hidden).  Replaces the dgrad GEMM of
 * fc2 plus hv_bias_gelu_bwd: the (rows, hidden) gradient of the activation output is never written.  tcgen05 / TMEM / TMA.
 * Requires rows % 128 == 0, hidden % 128 == 0, hidden <= 4096, C % 8 == 0 (hv_mlp_dgelu_gemm_workspace_bytes returns 0
 * otherwise; callers then use the two-kernel path). */
HV_API size_t hv_mlp_dgelu_gemm_workspace_bytes(int64_t rows, int hidden, int C);
/* Forward counterpart (reference swinv2.py:60-62): h = x W1^T (rows, hidden) bf16 WITHOUT the fc1 bias (what the backward
 * kernels read) and act = GELU(h + b1) from one tcgen05 GEMM; x (rows, C), w1 = fc1.weight (hidden, C) as stored.  Replaces
 * the fc1 GEMM + hv_bias_gelu_fwd (h is written once and never re-read in the forward).  Same shape conditions as
 * hv_mlp_dgelu_gemm (query them with hv_mlp_dgelu_gemm_workspace_bytes != 0). */
HV_API int hv_mlp_fc1_gelu_gemm(const void* x, const void* w1, const float* b1, void* h, void* act, int64_t rows, int hidden,
                                int C, int dtype, void* stream);
HV_API int hv_mlp_dgelu_gemm(const void* dy, const void* w2, const void* h, const float* b1, void* dh, float* db1,
                             void* workspace, size_t workspace_bytes, int64_t rows, int hidden, int C, int dtype, void* stream);

/* ---- Mlp activation with the fc1 bias folded in: out = GELU_erf(h + bias) ------------------
 * Replaces the bias add of `fc1` plus `self.act(x)` (nn.GELU, exact erf form), swinv2.py:61-62, and their
 * autograd; the caller runs fc1 as a bias-free GEMM.  Backward also yields d fc1.bias (column sums of dh).
 *   h, out, dout, dh  device (rows, cols) `dtype`; cols a multiple of 64 (fp32) / 128 (bf16)
 *   bias, dbias       device float32 (cols) */
HV_API int hv_bias_gelu_fwd(const void* h, const float* bias, void* out, int64_t rows, int cols, int dtype, void* stream);
HV_API size_t hv_bias_gelu_bwd_workspace_bytes(int64_t rows, int cols);
HV_API int hv_bias_gelu_bwd(const void* dout, const void* h, const float* bias, void* dh, float* dbias, void* workspace,
                     size_t workspace_bytes, int64_t rows, int cols, int dtype, void* stream);

/* ---- PatchMerging 2x2 gather ------------------------------------------------------------
 * Replaces the four strided slices + torch.cat of swinv2.py:484-491 (forward) and their
 * autograd (backward = the inverse permutation; every input token is read exactly once).
 *   x  device (B, H*W, C)      out device (B, H/2*W/2, 4C)       same dtype */
HV_API int hv_patch_merge_gather_fwd(const void* x, void* out, int B, int H, int W, int C, int dtype, void* stream);
HV_API int hv_patch_merge_gather_bwd(const void* dout, void* dx, int B, int H, int W, int C, int dtype, void* stream);

/* ---- continuous position bias table -----------------------------------------------------
 * table[r, h] = 16 * sigmoid(cpb_mlp(relative_coords_table)[r, h]) with cpb_mlp = Linear(2, hidden) + ReLU +
 * Linear(hidden, heads, no bias) (swinv2.py:141-145, 233-246); the gather through relative_position_index is done by
 * the attention kernels.  Replaces ~5 (forward) / ~8 (backward) tiny library kernels per block and step.
 *   coords device float32 (M, 2)   w1 (hidden, 2)   b1 (hidden)   w2 (heads, hidden)   table, dtable (M, heads)
 *   hidden <= 512 and a multiple of 32, heads <= 32;  workspace: M * heads floats (backward) */
HV_API int hv_cpb_bias_fwd(const float* coords, const float* w1, const float* b1, const float* w2, float* table, int M,
                    int hidden, int heads, void* stream);
HV_API int hv_cpb_bias_bwd(const float* coords, const float* w1, const float* b1, const float* w2, const float* dtable,
                    float* dw1, float* db1, float* dw2, float* workspace, int M, int hidden, int heads, void* stream);

/* ---- cross entropy with label smoothing, loss rows and gradient in one launch ---------------
 * The tail of the step: reference models.py:121-152 (`loss`), hierarchy.py:65-94 (MultitaskCrossEntropy =
 * dot(coeffs, CE per tier); call once per tier with scale = coeff / rows), algorithmic.py:88-119, 160-164 (label
 * smoothing: target = onehot (1 - a) + a / classes).
 *   logits    device (rows, classes) `dtype` (float32 / bfloat16)     target  device int64 (rows), class indices
 *   loss_rows device float32 (rows): lse - (1 - a) x[t] - a mean(x)   (the caller averages / weights them)
 *   dlogits   device (rows, classes) `dtype`: scale * (softmax - (1 - a) onehot - a / classes) */
HV_API int hv_cross_entropy_fwd_grad(const void* logits, const int64_t* target, float* loss_rows, void* dlogits, int64_t rows,
                                     int classes, float smoothing, float scale, int dtype, void* stream);

/* ---- PatchEmbed input gather -------------------------------------------------------------
 * Left operand of PatchEmbed's Conv2d(in_chans, embed_dim, kernel = stride = patch) seen as a per-patch GEMM
 * (swinv2.py:648-657): out[(b, ph, pw), (c, dy, dx)] = img[b, c, ph*P + dy, pw*P + dx] * scale[c] + shift[c].
 * `scale`/`shift` (device float32 (Cin), or both NULL) fold the reference's on-device normalisation
 * (data.py:130-136: (x - mean) / std) into the gather, so a uint8 batch is read exactly once.
 *   img  device (B, Cin, H, W), img_dtype in {HV_U8, HV_F32, HV_BF16}; only Cin = 3, P = 4
 *   out  device (B * H/P * W/P, Cin*P*P) out_dtype in {HV_F32, HV_BF16}; multiply by proj.weight.view(E, -1)^T */
HV_API int hv_patch_rows(const void* img, int img_dtype, const float* scale, const float* shift, void* out,
                  int out_dtype, int B, int Cin, int H, int W, int P, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HV_SWIN_H_ */
