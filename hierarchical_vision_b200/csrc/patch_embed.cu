// PatchEmbed input gather (reference swinv2.py:648-657): Conv2d(3 -> E, kernel = stride = P) over an NCHW image is
// a per-patch GEMM, tokens x (Cin*P*P) . (Cin*P*P) x E.  This kernel produces its left operand: for every patch
// token (b, ph, pw) the Cin*P*P pixels in the conv weight's own (c, dy, dx) order, optionally normalised on the
// fly ((x - mean_c) / std_c, the reference's device transform data.py:130-136, when the image is still uint8),
// in the GEMM's compute dtype.  It replaces cuDNN's NCHW->NHWC transposes around the convolution (three full
// passes over the image and the token tensor); the GEMM stays a library GEMM and LayerNorm is K4a.
//
// A CTA owns kTok consecutive tokens: reads are coalesced along image rows, the (kTok x K) tile is transposed
// through shared memory and written as one contiguous, 16-byte-vectorised block.
#include "hv_common.cuh"

namespace hv {
namespace {

constexpr int kTok = 64;
constexpr int kThreads = 256;

template <typename TIn> __device__ __forceinline__ float load_px(const TIn* p);
template <> __device__ __forceinline__ float load_px<unsigned char>(const unsigned char* p) { return (float)__ldg(p); }
template <> __device__ __forceinline__ float load_px<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_px<bf16>(const bf16* p) { return __bfloat162float(*p); }

template <typename TIn, typename TOut, int P, int CIN>
__global__ void __launch_bounds__(kThreads)
patch_rows_kernel(const TIn* __restrict__ img, const float* __restrict__ scale, const float* __restrict__ shift,
                  TOut* __restrict__ out, int H, int W, int64_t ntok) {
  constexpr int K = CIN * P * P;
  static_assert((kTok * P) % kThreads == 0 || kThreads % (kTok * P) == 0, "tile shape");
  __shared__ __align__(16) TOut tile[kTok * K];
  const int Hp = H / P, Wp = W / P;
  const int64_t tok0 = (int64_t)blockIdx.x * kTok;
  // pixel slot of this thread inside one (c, dy) image-row pass: token tl, column dx
  const int tl = threadIdx.x / P, dx = threadIdx.x % P;
  const int64_t tok = tok0 + tl;
  const bool live = tl < kTok && tok < ntok;
  int b = 0, ph = 0, pw = 0;
  if (live) {
    pw = (int)(tok % Wp);
    const int64_t t2 = tok / Wp;
    ph = (int)(t2 % Hp);
    b = (int)(t2 / Hp);
  }
#pragma unroll
  for (int c = 0; c < CIN; ++c) {
    const float sc = scale ? scale[c] : 1.0f, sh = shift ? shift[c] : 0.0f;
#pragma unroll
    for (int dy = 0; dy < P; ++dy) {
      if (live) {
        const TIn* src = img + (((int64_t)b * CIN + c) * H + (ph * P + dy)) * W + pw * P + dx;
        tile[tl * K + (c * P + dy) * P + dx] = from_f32<TOut>(fmaf(load_px<TIn>(src), sc, sh));
      }
    }
  }
  __syncthreads();
  // contiguous block of min(kTok, ntok - tok0) * K elements
  const int64_t rem = ntok - tok0;
  const int nel = (int)(rem < kTok ? rem : kTok) * K;
  constexpr int VE = 16 / sizeof(TOut);
  const int nvec = nel / VE;  // K * sizeof(TOut) is a multiple of 16 for P = 4, CIN = 3
  uint4* dst = reinterpret_cast<uint4*>(out + tok0 * K);
  const uint4* srcv = reinterpret_cast<const uint4*>(tile);
  for (int i = threadIdx.x; i < nvec; i += kThreads) dst[i] = srcv[i];
}

template <typename TIn, typename TOut>
int launch(const void* img, const float* scale, const float* shift, void* out, int B, int H, int W, cudaStream_t st) {
  const int64_t ntok = (int64_t)B * (H / 4) * (W / 4);
  const int64_t blocks = (ntok + kTok - 1) / kTok;
  if (blocks > 0x7fffffff) HV_FAIL(HV_ERR_SHAPE, "patch_rows: too many tokens");
  patch_rows_kernel<TIn, TOut, 4, 3><<<(int)blocks, kThreads, 0, st>>>((const TIn*)img, scale, shift, (TOut*)out, H, W, ntok);
  HV_LAUNCH_OK("patch_rows_kernel");
  return HV_OK;
}

}  // namespace

int patch_rows(const void* img, int img_dtype, const float* scale, const float* shift, void* out, int out_dtype, int B,
               int Cin, int H, int W, int P, cudaStream_t st) {
  if (Cin != 3 || P != 4) HV_FAIL(HV_ERR_SHAPE, "patch_rows: only in_chans=3, patch_size=4 (got %d, %d)", Cin, P);
  if (B <= 0 || H <= 0 || W <= 0 || H % P || W % P) HV_FAIL(HV_ERR_SHAPE, "patch_rows: image %dx%d not divisible by %d", H, W, P);
  if (!aligned16(out)) HV_FAIL(HV_ERR_ALIGN, "patch_rows: out must be 16-byte aligned");
  if (out_dtype != HV_F32 && out_dtype != HV_BF16) HV_FAIL(HV_ERR_DTYPE, "patch_rows: out dtype %d", out_dtype);
#define HV_PR(TIN)                                                                                   \
  return out_dtype == HV_F32 ? launch<TIN, float>(img, scale, shift, out, B, H, W, st)              \
                             : launch<TIN, bf16>(img, scale, shift, out, B, H, W, st)
  switch (img_dtype) {
    case HV_U8: HV_PR(unsigned char);
    case HV_F32: HV_PR(float);
    case HV_BF16: HV_PR(bf16);
    default: HV_FAIL(HV_ERR_DTYPE, "patch_rows: image dtype %d", img_dtype);
  }
#undef HV_PR
}

}  // namespace hv
