// Tensor-core fused shifted-window scaled-cosine attention for the SwinV2 hot shape:
// window 8x8 (N = 64 tokens), head dim 32, bf16 activations (every stage of SwinV2-T, stage 3 of -B).
//
// Replaces reference swinv2.py:399-412 (roll + window_partition), 221-261 (head split, cosine logits,
// logit scale, position bias, shift mask, softmax, attn @ v, head merge), 420-429 (window_reverse +
// roll back) and their autograd.  Data layout: qkv (B, H*W, 3C) and out (B, H*W, C) stay in IMAGE token
// order in HBM; the window gather / scatter is address arithmetic in the producer and the epilogue.
//
// Structure (persistent CTAs, one per SM, HG heads per CTA):
//   * 4 producer warps (one per SM sub-partition, 40 registers each after `setmaxnreg.dec`): per window,
//     16-byte `cp.async` (LDGSTS) copies straight from the rolled image position of every token into a
//     padded, bank-conflict-free shared-memory tile; completion is signalled on an mbarrier
//     (`cp.async.mbarrier.arrive.noinc`), multi-stage full/empty ring.  (Measured on B200,
//     tools/probes/tma_probe.cu: `cp.async.bulk` requests of one token row segment (64-576 B) cost
//     ~60-90 issue cycles each per warp and top out at 2.7 TB/s from one warp, while LDGSTS from 4 warps
//     reaches the 6.7 TB/s copy ceiling; bulk/TMA only wins for >= 4 KB contiguous requests, which a
//     partitioned head group of a rolled window never has.)
//   * 4 compute warps per head (152 registers each after `setmaxnreg.inc`): each owns 16 query rows
//     (forward) / 16 key rows (backward) of one (window, head) and keeps S/P entirely in registers
//     (mma.sync m16n8k16 bf16, fp32 accumulate).  The kernel is issue-bound on the CUDA cores, so every
//     reduction that can ride on the (otherwise ~80 % idle) tensor pipe does: squared row norms and
//     D = dO.o are diagonals of 16x16 self products, the softmax row sums are a ones-column MMA, the
//     subtraction of D from dP is an extra k-step against a (hi, lo) bf16 split of -D, and d(bias) is
//     accumulated over all windows of the CTA by MMAs against an identity selector.
//   * the position bias is block-Toeplitz in (ih - jh, iw - jw): 18 registers per thread hold every value
//     it ever needs; the shift mask is two 16-bit patterns per thread; softmax uses ex2 and skips the
//     running maximum for heads whose logit range provably fits fp32 (|logit| <= tau + bias range).
//   * backward recomputes S from q,k and the saved row log-sum-exp; dS goes through shared memory once
//     (bf16, stmatrix) to be re-read transposed for dQ; d(bias), d(tau) and the column sums of dq
//     (= gradient of q_bias) are reduced deterministically (per-CTA partials + a reduce kernel).
// The kernel is HBM-bound by bytes (37 FLOP/B): the design goal is bytes in flight and as few
// issue slots per logit as possible, not MMA rate.
#include <stdlib.h>

#include "hv_common.cuh"

namespace hv {
namespace {

constexpr int kN = 64;       // tokens per window
constexpr int kWs = 8;       // window side
constexpr int kTab = 225;    // (2*8-1)^2 bias-table rows
constexpr int kOstPitch = 80;                 // bytes per row of the per-warp 16x32 bf16 staging tile
constexpr int kOstBytes = 16 * kOstPitch;
constexpr float kNoMaxRange = 64.0f;  // log2 units: skip the running max when 2*tau2 + bias range <= this

// ---------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(0x989680u)  // suspend-time hint: sleep in hardware instead of spinning
        : "memory");
    if (done) break;
    if (++spins > (1u << 20)) __trap();  // a lost arrival must abort the kernel, never hang the GPU
  }
}
// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only) and its mbarrier completion hook
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// arrive on `bar` (without incrementing the pending count) once all prior cp.async of this thread landed
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void stsm_x4(uint32_t addr, const uint32_t (&r)[4]) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_fast(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rsqrt_fast(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t hmul2_bf16(uint32_t a, uint32_t b) {  // packed bf16x2 product
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
template <int N> __device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N> __device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
__device__ __forceinline__ float dot2(uint32_t w, float x, float y) {  // packed pair . (x, y)
  return fmaf(bf16lo_to_f32(w), x, bf16hi_to_f32(w) * y);
}

// Row-wise dot products x_i . y_i of two 16 x 32 bf16 tiles held as m16n8k16 A fragments (two k-steps each),
// exact in fp32, on the tensor pipe: rows 0-7 (8-15) of Y reinterpreted as a B operand are registers (0, 2)
// ((1, 3)) of the same fragment, so X Y^T costs four MMAs and the wanted values are its diagonal.  Row g
// (= lane / 4) lands in the quad's lane t = g / 2, element g % 2; a shuffle broadcasts it to the quad.
// Returns s0 = x_g . y_g and s1 = x_{g+8} . y_{g+8} in every lane.
__device__ __forceinline__ void rowdot_mma(const uint32_t (&x)[2][4], const uint32_t (&y)[2][4], int lane, float& s0,
                                           float& s1) {
  float n0[4] = {0.f, 0.f, 0.f, 0.f}, n1[4] = {0.f, 0.f, 0.f, 0.f};
  mma_bf16(n0, x[0], y[0][0], y[0][2]);
  mma_bf16(n1, x[0], y[0][1], y[0][3]);
  mma_bf16(n0, x[1], y[1][0], y[1][2]);
  mma_bf16(n1, x[1], y[1][1], y[1][3]);
  const bool odd = (lane >> 2) & 1;
  const float v0 = odd ? n0[1] : n0[0];
  const float v1 = odd ? n1[3] : n1[2];
  const int src = (lane & ~3) | (lane >> 3);
  s0 = __shfl_sync(0xffffffffu, v0, src);
  s1 = __shfl_sync(0xffffffffu, v1, src);
}

// Lane address pieces for ldmatrix.x4 (byte offsets relative to a [row][pitch] bf16 tile)
//  A operand (16 rows x 16 k):   row = (l&7) + 8*((l>>3)&1), k byte = (l>>4)*16
//  B operand from [n][k] rows (8 n x 32 k):  row = l&7, k byte = (l>>3)*16
//  B operand from [k][n] rows via .trans (16 k x 16 n): row = (l&7) + 8*((l>>3)&1), n byte = (l>>4)*16
__device__ __forceinline__ int lane_row16(int l) { return (l & 7) + 8 * ((l >> 3) & 1); }

// Store a 16 x 32 fp32 accumulator tile (rows g, g+8 of the warp's block; 4 n-tiles) as bf16 to two
// global rows: fragments -> per-warp shared staging -> one 16-byte store per row per lane.
template <int PITCH = kOstPitch>
__device__ __forceinline__ void store_tile_bf16(const float (&acc)[4][4], uint32_t ost, int g_, int t_, bf16* row0,
                                                bf16* row1) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    sts32(ost + g_ * PITCH + nt * 16 + t_ * 4, pack_bf16x2(acc[nt][0], acc[nt][1]));
    sts32(ost + (g_ + 8) * PITCH + nt * 16 + t_ * 4, pack_bf16x2(acc[nt][2], acc[nt][3]));
  }
  __syncwarp();
  const uint4 v0 = lds128(ost + g_ * PITCH + t_ * 16);
  const uint4 v1 = lds128(ost + (g_ + 8) * PITCH + t_ * 16);
  *reinterpret_cast<uint4*>(row0 + t_ * 8) = v0;
  *reinterpret_cast<uint4*>(row1 + t_ * 8) = v1;
  __syncwarp();
}

// 16-bit patterns (bit 2*nt+e <-> column slot 8*nt + 2*t + e) telling which columns lie in the
// wrapped band of the window along h / along w; see hv_index.h::window_slot_region.
__device__ __forceinline__ void column_band_bits(int t_, int hi_thr, uint32_t& colH, uint32_t& colW) {
  colH = 0; colW = 0;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (nt >= hi_thr) colH |= 1u << (2 * nt + e);          // column slot row = nt
      if (2 * t_ + e >= hi_thr) colW |= 1u << (2 * nt + e);  // column slot col = 2t+e
    }
}

// Window cursor of a persistent CTA: (image b, window win) advance by a fixed stride per iteration, so the
// per-iteration update is two adds and a compare instead of integer divisions.
struct Cursor {
  int b, win, step_b, step_w;
  __device__ __forceinline__ void init(const Geom& g, int first_row, int stride) {
    b = first_row / g.nW; win = first_row - b * g.nW;
    step_b = stride / g.nW; step_w = stride - step_b * g.nW;
  }
  __device__ __forceinline__ void next(const Geom& g) {
    b += step_b; win += step_w;
    if (win >= g.nW) { win -= g.nW; ++b; }
  }
};
// Stage cursor of the full/empty ring
template <int STAGES> struct Ring {
  int s = 0;
  uint32_t ph = 0;
  __device__ __forceinline__ void next() {
    if (++s == STAGES) { s = 0; ph ^= 1u; }
  }
};
// Window (wh, ww) from win: exact for win < 2^16 (float reciprocal + 0.5 guard).
__device__ __forceinline__ void window_rc(const Geom& g, float inv_nWw, int win, int& wh, int& ww) {
  wh = __float2int_rz((win + 0.5f) * inv_nWw);
  ww = win - wh * g.nWw;
}
// Token index of slot (ih, iw) of the window whose first shifted row/col are row0/col0.
__device__ __forceinline__ int64_t tile_token(const Geom& g, int b, int row0, int col0, int ih, int iw) {
  int r = row0 + ih; if (r >= g.H) r -= g.H;
  int c = col0 + iw; if (c >= g.W) c -= g.W;
  return ((int64_t)b * g.H + r) * g.W + c;
}
// 1 / max(sqrt(ss), 1e-12): F.normalize's denominator (reference swinv2.py:229)
__device__ __forceinline__ float inv_norm(float ss) { return rsqrt_fast(fmaxf(ss, 1e-24f)); }

// Register split between the roles (only when the CTA fills the register file: 16 warps x 128)
template <int HG> struct Regs {
  static constexpr bool kSplit = (HG == 3);
  static constexpr int kCompute = 152, kProducer = 40;        // forward
  static constexpr int kComputeBwd = 152, kProducerBwd = 56;  // backward: the producers also run the pre-pass
};

// Shift-mask patterns of one thread (rows g, g+8 of a 16-row block wq; columns 8nt+2t+e)
struct MaskBits {
  uint32_t colH, colW;
  bool r0H, r1H, rW;
  __device__ __forceinline__ void init(int wq, int g_, int t_, int hi_thr) {
    column_band_bits(t_, hi_thr, colH, colW);
    r0H = (2 * wq) >= hi_thr; r1H = (2 * wq + 1) >= hi_thr; rW = g_ >= hi_thr;
  }
  __device__ __forceinline__ void apply(float (&acc)[8][4], bool bottom, bool right, float neg) const {
    // Only edge windows come here (warp-uniform).  The convergence point keeps ptxas from if-converting the 64
    // predicated adds into the straight-line path that every interior window executes.
    __syncwarp();
    const uint32_t m0 = (bottom ? (r0H ? ~colH : colH) : 0u) | (right ? (rW ? ~colW : colW) : 0u);
    const uint32_t m1 = (bottom ? (r1H ? ~colH : colH) : 0u) | (right ? (rW ? ~colW : colW) : 0u);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (m0 & (1u << (2 * nt + e))) acc[nt][e] += neg;
        if (m1 & (1u << (2 * nt + e))) acc[nt][2 + e] += neg;
      }
  }
};

// =============================================================================== forward
template <int HG> struct FwdCfg {
  static constexpr int kWarps = 4 * HG;             // compute warps
  static constexpr int kProducers = 4;              // producer warps
  static constexpr int kThreads = (kWarps + kProducers) * 32;
  static constexpr int kPitch = HG * 192 + 16;  // [q | k | v] x HG heads (64 B each) + 16 B pad: odd multiple of 16
  static constexpr int kStageBytes = kN * kPitch;
  static constexpr int kStages = 4;
  static constexpr int kOffOst = kStages * kStageBytes;
  static constexpr int kOffCvec = kOffOst + kWarps * kOstBytes;  // [2][HG][64] float
  static constexpr int kOffBar = kOffCvec + 2 * HG * kN * 4;
  static constexpr int kSmem = kOffBar + 2 * kStages * 8;
};

template <int HG>
__global__ void __launch_bounds__(FwdCfg<HG>::kThreads, 1)
wattn_mma64_fwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ bias_table, const float* __restrict__ tau,
                       bf16* __restrict__ out, float* __restrict__ lse, Geom g, int ctas_per_group) {
  using Cfg = FwdCfg<HG>;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nHG = g.heads / HG;
  const int hgrp = blockIdx.x % nHG;
  const int cta = blockIdx.x / nHG;
  const int nrows = g.B * g.nW;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + Cfg::kOffBar, bar_empty = bar_full + 8 * Cfg::kStages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(bar_full + 8 * s, Cfg::kProducers * 32);
      mbar_init(bar_empty + 8 * s, Cfg::kWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= Cfg::kWarps) {
    // ------------------------------------------------------------------ producer warps
    // Each producer warp owns two of the window's eight token rows; the (token, chunk) a lane handles in a given
    // instruction is the same for every window, so its shared/global offsets are precomputed once.
    if constexpr (Regs<HG>::kSplit) reg_dealloc<Regs<HG>::kProducer>();
    const int pw = warp - Cfg::kWarps;
    // One window row of one part (q, k or v) = 8 tokens x 4*HG chunks = HG full-warp instructions, so instruction
    // (part, p) copies the same (token, chunk) pattern p for every part: HG lane constants instead of kRowInstr.
    int p_iw[HG], p_soff[HG], p_goff[HG];
#pragma unroll
    for (int p = 0; p < HG; ++p) {
      const int q = lane + 32 * p;
      p_iw[p] = q / (4 * HG);
      const int within = q - p_iw[p] * (4 * HG);
      p_soff[p] = p_iw[p] * Cfg::kPitch + within * 16;
      p_goff[p] = hgrp * (HG * 32) + within * 8;
    }
    const int tok_stride = 3 * g.C;
    const float inv_nWw = 1.0f / (float)g.nWw;
    Cursor cur;
    cur.init(g, cta, ctas_per_group);
    Ring<Cfg::kStages> ring;
    for (int row = cta; row < nrows; row += ctas_per_group, cur.next(g), ring.next()) {
      int wh, ww;
      window_rc(g, inv_nWw, cur.win, wh, ww);
      const int col0 = ww * kWs + g.shift;
      if (pw == 0) mbar_wait(bar_empty + 8 * ring.s, ring.ph ^ 1);  // one waiter; the other producers sleep in the barrier
      named_bar_sync(10, Cfg::kProducers * 32);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int ih = 2 * pw + r;
        int irow = wh * kWs + g.shift + ih;
        if (irow >= g.H) irow -= g.H;
        const bf16* rowp = qkv + ((int64_t)cur.b * g.H + irow) * g.W * tok_stride;
        const uint32_t dst = sbase + ring.s * Cfg::kStageBytes + ih * (kWs * Cfg::kPitch);
#pragma unroll
        for (int p = 0; p < HG; ++p) {
          int col = col0 + p_iw[p];
          if (col >= g.W) col -= g.W;
          const bf16* src = rowp + col * tok_stride + p_goff[p];
#pragma unroll
          for (int part = 0; part < 3; ++part) cp_async16(dst + p_soff[p] + part * (HG * 64), src + part * g.C);
        }
      }
      cp_async_arrive(bar_full + 8 * ring.s);
    }
    return;
  }

  // -------------------------------------------------------------------- compute warps
  if constexpr (Regs<HG>::kSplit) reg_alloc<Regs<HG>::kCompute>();
  const int hh = warp >> 2, wq = warp & 3;
  const int head = hgrp * HG + hh;
  const int g_ = lane >> 2, t_ = lane & 3;
  const int i0 = 16 * wq + g_, i1 = i0 + 8;  // own query slots: (ih 2wq, iw g), (ih 2wq+1, iw g)
  const float tau2 = __ldg(&tau[head]) * kLog2e;
  // Logits are tau2*cos + bias2 with |cos| <= 1: if 2*tau2 + (bias range) stays far inside the fp32 exponent range,
  // exp2(logit - (tau2 + max bias)) can neither overflow nor lose the row, and the running maximum is skipped.
  float bmx = -3.0e38f, bmn = 3.0e38f;
  {
    float bv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) bv[k] = __ldg(&bias_table[min(lane + 32 * k, kTab - 1) * g.heads + head]);  // 8 loads in flight
#pragma unroll
    for (int k = 0; k < 8; ++k) { bmx = fmaxf(bmx, kLog2e * bv[k]); bmn = fminf(bmn, kLog2e * bv[k]); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    bmx = fmaxf(bmx, __shfl_xor_sync(0xffffffffu, bmx, o));
    bmn = fminf(bmn, __shfl_xor_sync(0xffffffffu, bmn, o));
  }
  const bool use_max = !(2.0f * tau2 + (bmx - bmn) <= kNoMaxRange);
  const float off = use_max ? 0.f : tau2 + bmx;
  // The bias is block-Toeplitz: rows (ih = 2wq + rh, iw = g), columns (jh = nt, jw = 2t + e) only involve
  // d = nt - rh + 1 in [0, 8] and e, so 18 registers hold every value this thread needs (pre-shifted by -off).
  float bias2[9][2];
#pragma unroll
  for (int d = 0; d < 9; ++d)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int r = (2 * wq + 1 - d + 7) * 15 + (g_ - 2 * t_ - e + 7);
      bias2[d][e] = kLog2e * __ldg(&bias_table[r * g.heads + head]) - off;
    }
  const int hi_thr = kWs - g.shift;
  MaskBits mb;
  mb.init(wq, g_, t_, hi_thr);
  const int nWh = g.H / kWs;
  const float kNeg = kMaskValue * kLog2e;

  const uint32_t ost = sbase + Cfg::kOffOst + warp * kOstBytes;
  float* cvec_base = reinterpret_cast<float*>(smem + Cfg::kOffCvec);
  const int arow = lane_row16(lane), acolb = (lane >> 4) * 16;
  const int brow = lane & 7, bcolb = (lane >> 3) * 16;
  const uint32_t own = (16 * wq + arow) * Cfg::kPitch + acolb;

  const float inv_nWw = 1.0f / (float)g.nWw;
  const uint32_t ones_b = (g_ == 0) ? 0x3F803F80u : 0u;  // B fragment of a ones column: row sums of P by MMA
  Cursor cur;
  cur.init(g, cta, ctas_per_group);
  Ring<Cfg::kStages> ring;
  int it = 0;
  for (int row = cta; row < nrows; row += ctas_per_group, ++it, cur.next(g), ring.next()) {
    int wh, ww;
    window_rc(g, inv_nWw, cur.win, wh, ww);
    const int row0 = wh * kWs + g.shift, col0 = ww * kWs + g.shift;
    if (wq == 0) mbar_wait(bar_full + 8 * ring.s, ring.ph);  // one waiter per head; the rest sleep in the barrier
    named_bar_sync(1 + hh, 128);
    const uint32_t st = sbase + ring.s * Cfg::kStageBytes;
    const uint32_t qb = st + hh * 64, kb_ = st + HG * 64 + hh * 64, vb_ = st + 2 * HG * 64 + hh * 64;

    // --- own 16 key rows and 16 query rows: inverse norms (key norms, scaled by tau*log2e, go to a shared vector)
    uint32_t ka[2][4], qa[2][4], kf[2][4];
    ldsm_x4(kb_ + own, ka[0]);
    ldsm_x4(kb_ + own + 32, ka[1]);
    ldsm_x4(qb + own, qa[0]);
    ldsm_x4(qb + own + 32, qa[1]);
    ldsm_x4(kb_ + brow * Cfg::kPitch + bcolb, kf[0]);
    float* cvec = cvec_base + ((it & 1) * HG + hh) * kN;
    float r0, r1;
    {
      float s0, s1;
      rowdot_mma(ka, ka, lane, s0, s1);
      const float c0 = tau2 * inv_norm(s0), c1 = tau2 * inv_norm(s1);
      if (t_ == 0) {
        cvec[i0] = c0;
        cvec[i1] = c1;
      }
      rowdot_mma(qa, qa, lane, s0, s1);
      r0 = inv_norm(s0);
      r1 = inv_norm(s1);
      if (t_ == 0) {  // row scales beside the log-sum-exp (planes 1: r, 2: c) for the tcgen05 backward kernel
        const int64_t plane = (int64_t)nrows * g.heads * kN;
        float* sp = lse + plane + ((int64_t)row * g.heads + head) * kN;
        sp[i0] = r0;
        sp[i1] = r1;
        sp[plane + i0] = c0;
        sp[plane + i1] = c1;
      }
    }
    named_bar_sync(1 + hh, 128);

    // --- S = Q K^T (raw dot products), K fragments double-buffered ahead of the MMAs
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt + 1 < 8) ldsm_x4(kb_ + (8 * (nt + 1) + brow) * Cfg::kPitch + bcolb, kf[(nt + 1) & 1]);
      acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      mma_bf16(acc[nt], qa[0], kf[nt & 1][0], kf[nt & 1][1]);
      mma_bf16(acc[nt], qa[1], kf[nt & 1][2], kf[nt & 1][3]);
    }
    uint32_t vf[2][4];
    ldsm_x4_t(vb_ + arow * Cfg::kPitch + acolb, vf[0]);
    // --- logits in the log2 domain: tau*cos + bias (+ mask)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 c = *reinterpret_cast<const float2*>(&cvec[8 * nt + 2 * t_]);
      acc[nt][0] = fmaf(acc[nt][0] * r0, c.x, bias2[nt + 1][0]);
      acc[nt][1] = fmaf(acc[nt][1] * r0, c.y, bias2[nt + 1][1]);
      acc[nt][2] = fmaf(acc[nt][2] * r1, c.x, bias2[nt][0]);
      acc[nt][3] = fmaf(acc[nt][3] * r1, c.y, bias2[nt][1]);
    }
    if (g.shift > 0) {
      const bool bottom = wh == nWh - 1, right = ww == g.nWw - 1;
      if (bottom || right) mb.apply(acc, bottom, right, kNeg);
    }
    // --- softmax numerators (rows g, g+8)
    float mx0 = 0.f, mx1 = 0.f;
    if (use_max) {
      mx0 = acc[0][0]; mx1 = acc[0][2];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        mx0 = fmaxf(mx0, fmaxf(acc[nt][0], acc[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(acc[nt][2], acc[nt][3]));
      }
      mx0 = quad_max(mx0);
      mx1 = quad_max(mx1);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        acc[nt][0] -= mx0; acc[nt][1] -= mx0;
        acc[nt][2] -= mx1; acc[nt][3] -= mx1;
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      acc[nt][0] = ex2(acc[nt][0]);
      acc[nt][1] = ex2(acc[nt][1]);
      acc[nt][2] = ex2(acc[nt][2]);
      acc[nt][3] = ex2(acc[nt][3]);
    }
    // --- O = P V ; the row sums l = P 1 ride along as a fifth n-tile whose B operand is a ones column
    float o[4][4], lacc[4];
    lacc[0] = lacc[1] = lacc[2] = lacc[3] = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(acc[2 * ks][0], acc[2 * ks][1]);
      pa[1] = pack_bf16x2(acc[2 * ks][2], acc[2 * ks][3]);
      pa[2] = pack_bf16x2(acc[2 * ks + 1][0], acc[2 * ks + 1][1]);
      pa[3] = pack_bf16x2(acc[2 * ks + 1][2], acc[2 * ks + 1][3]);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int idx = 2 * ks + half;
        if (idx + 1 < 8)
          ldsm_x4_t(vb_ + (16 * ((idx + 1) >> 1) + arow) * Cfg::kPitch + ((idx + 1) & 1) * 32 + acolb, vf[(idx + 1) & 1]);
        mma_bf16(o[2 * half], pa, vf[idx & 1][0], vf[idx & 1][1]);
        mma_bf16(o[2 * half + 1], pa, vf[idx & 1][2], vf[idx & 1][3]);
      }
      mma_bf16(lacc, pa, ones_b, ones_b);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * ring.s);  // the stage is free for the producer

    // column 0 of the ones tile sits in the t == 0 lane of every quad
    const float l0 = __shfl_sync(0xffffffffu, lacc[0], lane & ~3);
    const float l1 = __shfl_sync(0xffffffffu, lacc[2], lane & ~3);
    const float inv0 = rcp_fast(l0), inv1 = rcp_fast(l1);
    if (t_ == 0) {  // row log-sum-exp in log2 units (what the backward kernel consumes)
      float* lp = lse + ((int64_t)row * g.heads + head) * kN;
      lp[i0] = off + mx0 + lg2_fast(l0);
      lp[i1] = off + mx1 + lg2_fast(l1);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      o[nt][0] *= inv0; o[nt][1] *= inv0;
      o[nt][2] *= inv1; o[nt][3] *= inv1;
    }
    bf16* orow0 = out + tile_token(g, cur.b, row0, col0, 2 * wq, g_) * g.C + head * 32;
    bf16* orow1 = out + tile_token(g, cur.b, row0, col0, 2 * wq + 1, g_) * g.C + head * 32;
    store_tile_bf16(o, ost, g_, t_, orow0, orow1);
  }
}

// =============================================================================== backward
template <int HG> struct BwdCfg {
  static constexpr int kWarps = 4 * HG;
  static constexpr int kProducers = 4;
  static constexpr int kThreads = (kWarps + kProducers) * 32;
  static constexpr int kPitch = HG * 320 + 16;  // [q | k | v | o | dO] x HG heads + pad: odd multiple of 16
  static constexpr int kLseOff = kN * kPitch;   // HG x 64 fp32 row log-sum-exp behind the token rows
  // per-head vectors over the window's 64 rows, written by the producer warps' pre-pass:
  // (r, lse) float2 | -D as (hi, lo) bf16 pair | r as bf16 | c float
  static constexpr int kVecBytes = kN * 8 + kN * 4 + kN * 2 + kN * 4 + kN * 2 /* pad to 16 B multiple */;
  static constexpr int kVecOff = kLseOff + HG * kN * 4;
  static constexpr int kStageBytes = kVecOff + HG * kVecBytes;
  static constexpr int kStages = 3;
  static constexpr int kDsPitch = 144;                      // bytes per dS row (64 bf16 + 16 B pad)
  static constexpr int kOffDs = kStages * kStageBytes;      // [HG][64 j][64 i] bf16
  static constexpr int kOffRed = kOffDs + HG * kN * kDsPitch;   // per-warp d(tau) partials, [HG][4][32] column sums
  static constexpr int kRedBytes = kWarps * 4 + kWarps * 32 * 4;
  static constexpr int kOffBar = kOffRed + ((kRedBytes + 15) / 16) * 16;
  static constexpr int kSmem = kOffBar + 3 * kStages * 8;       // full | empty | pre
  static constexpr int kBinBytes = kWarps * 32 * 32 * 4;        // end-of-kernel d(bias) bins, reuse the stage ring
  static_assert(kBinBytes <= kStages * kStageBytes, "d(bias) bins must fit in the stage ring");
};

template <int HG>
__global__ void __launch_bounds__(BwdCfg<HG>::kThreads, 1)
wattn_mma64_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ out, const bf16* __restrict__ dout,
                       const float* __restrict__ lse, const float* __restrict__ bias_table, const float* __restrict__ tau,
                       bf16* __restrict__ dqkv, float* __restrict__ ws_dbias, float* __restrict__ ws_dtau,
                       float* __restrict__ ws_colsum, Geom g, int ctas_per_group) {
  using Cfg = BwdCfg<HG>;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nHG = g.heads / HG;
  const int hgrp = blockIdx.x % nHG;
  const int cta = blockIdx.x / nHG;
  const int nrows = g.B * g.nW;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + Cfg::kOffBar, bar_empty = bar_full + 8 * Cfg::kStages;
  const uint32_t bar_pre = bar_empty + 8 * Cfg::kStages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(bar_full + 8 * s, Cfg::kProducers * 32);
      mbar_init(bar_empty + 8 * s, Cfg::kWarps);
      mbar_init(bar_pre + 8 * s, Cfg::kProducers);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= Cfg::kWarps) {
    // ------------------------------------------------------------------ producer warps
    // q,k,v (from qkv), o (from out), dO (from dout) token segments and the row log-sum-exp of the window;
    // same lane-constant chunk schedule as the forward producer (two window rows per producer warp).
    if constexpr (Regs<HG>::kSplit) reg_dealloc<Regs<HG>::kProducerBwd>();
    const int pw = warp - Cfg::kWarps;
    // One window row of one part (q, k, v, o or dO) = 8 tokens x 4*HG chunks = HG full-warp instructions, so
    // instruction (part, p) copies the same (token, chunk) pattern p for every part: HG lane constants.
    int p_iw[HG], p_soff[HG], p_goff[HG];
#pragma unroll
    for (int p = 0; p < HG; ++p) {
      const int q = lane + 32 * p;
      p_iw[p] = q / (4 * HG);
      const int within = q - p_iw[p] * (4 * HG);
      p_soff[p] = p_iw[p] * Cfg::kPitch + within * 16;
      p_goff[p] = hgrp * (HG * 32) + within * 8;
    }
    const float inv_nWw = 1.0f / (float)g.nWw;
    const int g_ = lane >> 2, t_ = lane & 3;
    const int arow = lane_row16(lane), acolb = (lane >> 4) * 16;
    const uint32_t own = (16 * pw + arow) * Cfg::kPitch + acolb;
    // Pre-pass over the 16 token rows this warp loaded itself (all HG heads): 1/|k|, 1/|q|, D = dO . O on the tensor
    // pipe, published as per-stage vectors.  It runs one stage behind the copies, so the compute warps find the
    // vectors ready and never wait for each other before the first MMA of a window.
    auto prepass = [&](int s) {
      const uint32_t st = sbase + s * Cfg::kStageBytes;
      unsigned char* stg = smem + s * Cfg::kStageBytes;
#pragma unroll 1
      for (int hh = 0; hh < HG; ++hh) {
        const uint32_t qb = st + hh * 64, kb_ = st + HG * 64 + hh * 64;
        const uint32_t ob = st + 3 * HG * 64 + hh * 64, gb = st + 4 * HG * 64 + hh * 64;
        uint32_t ka[2][4], qa[2][4], oa[2][4], ga[2][4];
        ldsm_x4(kb_ + own, ka[0]);
        ldsm_x4(kb_ + own + 32, ka[1]);
        ldsm_x4(qb + own, qa[0]);
        ldsm_x4(qb + own + 32, qa[1]);
        ldsm_x4(ob + own, oa[0]);
        ldsm_x4(ob + own + 32, oa[1]);
        ldsm_x4(gb + own, ga[0]);
        ldsm_x4(gb + own + 32, ga[1]);
        float c0, c1, r0, r1, d0, d1;
        rowdot_mma(ka, ka, lane, c0, c1);
        rowdot_mma(qa, qa, lane, r0, r1);
        rowdot_mma(ga, oa, lane, d0, d1);
        if (t_ == 0) {
          const float* lse_s = reinterpret_cast<const float*>(stg + Cfg::kLseOff) + hh * kN;
          unsigned char* vecs = stg + Cfg::kVecOff + hh * Cfg::kVecBytes;
          float2* rl = reinterpret_cast<float2*>(vecs);
          uint32_t* dhl = reinterpret_cast<uint32_t*>(vecs + kN * 8);
          bf16* rb16 = reinterpret_cast<bf16*>(vecs + kN * 12);
          float* cv = reinterpret_cast<float*>(vecs + kN * 14);
          const int j0 = 16 * pw + g_, j1 = j0 + 8;
          r0 = inv_norm(r0); r1 = inv_norm(r1);
          rl[j0] = make_float2(r0, lse_s[j0]);
          rl[j1] = make_float2(r1, lse_s[j1]);
          rb16[j0] = __float2bfloat16_rn(r0);
          rb16[j1] = __float2bfloat16_rn(r1);
          cv[j0] = inv_norm(c0);
          cv[j1] = inv_norm(c1);
          const bf16 h0 = __float2bfloat16_rn(-d0), h1 = __float2bfloat16_rn(-d1);
          dhl[j0] = pack_bf16x2(__bfloat162float(h0), -d0 - __bfloat162float(h0));
          dhl[j1] = pack_bf16x2(__bfloat162float(h1), -d1 - __bfloat162float(h1));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pre + 8 * s);
    };
    Cursor cur;
    cur.init(g, cta, ctas_per_group);
    Ring<Cfg::kStages> ring;
    int prev_s = -1;
    for (int row = cta; row < nrows; row += ctas_per_group, cur.next(g), ring.next()) {
      int wh, ww;
      window_rc(g, inv_nWw, cur.win, wh, ww);
      const int col0 = ww * kWs + g.shift;
      if (pw == 0) mbar_wait(bar_empty + 8 * ring.s, ring.ph ^ 1);  // one waiter; the other producers sleep in the barrier
      named_bar_sync(10, Cfg::kProducers * 32);
      const uint32_t st = sbase + ring.s * Cfg::kStageBytes;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int ih = 2 * pw + r;
        int irow = wh * kWs + g.shift + ih;
        if (irow >= g.H) irow -= g.H;
        const int rowtok = (cur.b * g.H + irow) * g.W;  // token index < 2^31 (checked on the host)
        const uint32_t dst = st + ih * (kWs * Cfg::kPitch);
#pragma unroll
        for (int p = 0; p < HG; ++p) {
          int col = col0 + p_iw[p];
          if (col >= g.W) col -= g.W;
          const int64_t tok = rowtok + col;
          const bf16* s3 = qkv + tok * (3 * g.C) + p_goff[p];
          const int64_t o1 = tok * g.C + p_goff[p];
          const uint32_t d = dst + p_soff[p];
          cp_async16(d, s3);
          cp_async16(d + HG * 64, s3 + g.C);
          cp_async16(d + 2 * HG * 64, s3 + 2 * g.C);
          cp_async16(d + 3 * HG * 64, out + o1);
          cp_async16(d + 4 * HG * 64, dout + o1);
        }
      }
      if (lane < 4 * HG) {  // the row log-sum-exp of this warp's own 16 rows: 64 B per head
        const int hh = lane >> 2, ch = lane & 3;
        const float* lrow = lse + ((int64_t)row * g.heads + hgrp * HG + hh) * kN + 16 * pw + 4 * ch;
        cp_async16(st + Cfg::kLseOff + (hh * kN + 16 * pw + 4 * ch) * 4, lrow);
      }
      cp_async_arrive(bar_full + 8 * ring.s);
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (prev_s >= 0) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");  // the previous stage's copies of this warp have landed
        __syncwarp();
        prepass(prev_s);
      }
      prev_s = ring.s;
    }
    if (prev_s >= 0) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      prepass(prev_s);
    }
    return;
  }

  // -------------------------------------------------------------------- compute warps
  if constexpr (Regs<HG>::kSplit) reg_alloc<Regs<HG>::kComputeBwd>();
  const int hh = warp >> 2, wk = warp & 3;
  const int head = hgrp * HG + hh;
  const int g_ = lane >> 2, t_ = lane & 3;
  const int j0 = 16 * wk + g_, j1 = j0 + 8;  // own key slots in the S^T pass == own query slots in the dQ pass
  const float tau_h = __ldg(&tau[head]);
  const float tau2 = tau_h * kLog2e;
  const int hi_thr = kWs - g.shift;
  MaskBits mb;
  mb.init(wk, g_, t_, hi_thr);
  const int nWh = g.H / kWs;
  const float kNeg = kMaskValue * kLog2e;

  const uint32_t dsT = sbase + Cfg::kOffDs + hh * kN * Cfg::kDsPitch;
  // Bias, S^T orientation: rows = keys (jh = 2wk + rh, jw = g), columns = queries (ih = nt, iw = 2t + e);
  // d = nt - rh + 1 in [0, 8]
  float bias2[9][2];
#pragma unroll
  for (int d = 0; d < 9; ++d)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int r = (d - 1 - 2 * wk + 7) * 15 + (2 * t_ + e - g_ + 7);
      bias2[d][e] = kLog2e * __ldg(&bias_table[r * g.heads + head]);
    }
  // d(bias) of this warp's 16 x 64 block of (key, query) pairs, accumulated over all windows by the tensor pipe
  float dbacc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) dbacc[nt][0] = dbacc[nt][1] = dbacc[nt][2] = dbacc[nt][3] = 0.f;
  float qsum[8];  // column sums (over this thread's rows, all windows) of dq
#pragma unroll
  for (int i = 0; i < 8; ++i) qsum[i] = 0.f;
  float dtau_acc = 0.f;

  const int arow = lane_row16(lane), acolb = (lane >> 4) * 16;
  const int brow = lane & 7, bcolb = (lane >> 3) * 16;
  const uint32_t own = (16 * wk + arow) * Cfg::kPitch + acolb;
  // constant fragments: A = ones in k = 0, 1 (picks the (hi, lo) row of the -D operand); B = identity selector
  uint32_t ones_a[4];
  ones_a[0] = ones_a[1] = (t_ == 0) ? 0x3F803F80u : 0u;
  ones_a[2] = ones_a[3] = 0u;
  const uint32_t sel = ((2 * t_ == g_) ? 0x00003F80u : 0u) | ((2 * t_ + 1 == g_) ? 0x3F800000u : 0u);

  const float inv_nWw = 1.0f / (float)g.nWw;
  Cursor cur;
  cur.init(g, cta, ctas_per_group);
  Ring<Cfg::kStages> ring;
  for (int row = cta; row < nrows; row += ctas_per_group, cur.next(g), ring.next()) {
    int wh, ww;
    window_rc(g, inv_nWw, cur.win, wh, ww);
    const int row0 = wh * kWs + g.shift, col0 = ww * kWs + g.shift;
    // own token rows: slot j0 = (ih 2*wk, iw g_), j1 = (ih 2*wk+1, iw g_)
    const int tok0 = (int)tile_token(g, cur.b, row0, col0, 2 * wk, g_);  // < 2^31 (checked on the host)
    const int tok1 = (int)tile_token(g, cur.b, row0, col0, 2 * wk + 1, g_);
    // the producers' pre-pass barrier implies the stage is full (each producer arrives after its own copies landed);
    // one waiter per head, the rest sleep in the named barrier, which also fences the previous tile's dS buffer
    if (wk == 0) mbar_wait(bar_pre + 8 * ring.s, ring.ph);
    named_bar_sync(1 + hh, 128);
    const uint32_t st = sbase + ring.s * Cfg::kStageBytes;
    const uint32_t qb = st + hh * 64, kb_ = st + HG * 64 + hh * 64, vb_ = st + 2 * HG * 64 + hh * 64;
    const uint32_t ob = st + 3 * HG * 64 + hh * 64, gb = st + 4 * HG * 64 + hh * 64;
    // per-stage vectors published by the producers: (1/|q_i|, lse2_i) | -D_i as bf16 (hi, lo) | 1/|q_i| bf16 | 1/|k_j|
    const uint32_t vec_u = st + Cfg::kVecOff + hh * Cfg::kVecBytes;
    const uint32_t rl_u = vec_u, dhl_u = vec_u + kN * 8, rb_u = vec_u + kN * 12;
    const float* vecf = reinterpret_cast<const float*>(smem + ring.s * Cfg::kStageBytes + Cfg::kVecOff + hh * Cfg::kVecBytes);
    // output staging: the O segment of this warp's own 16 token rows is dead once the producers' pre-pass is done
    const uint32_t ost = ob + (16 * wk) * Cfg::kPitch;

    uint32_t ka[2][4], qf[2][4];
    ldsm_x4(kb_ + own, ka[0]);
    ldsm_x4(kb_ + own + 32, ka[1]);
    ldsm_x4(qb + brow * Cfg::kPitch + bcolb, qf[0]);
    const float c0 = vecf[kN * 14 / 4 + j0], c1 = vecf[kN * 14 / 4 + j1];
    const float r0 = vecf[2 * j0], r1 = vecf[2 * j1];

    // --- S^T = K Q^T for own 16 keys (rows) x 64 queries (columns), then P^T
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt + 1 < 8) ldsm_x4(qb + (8 * (nt + 1) + brow) * Cfg::kPitch + bcolb, qf[(nt + 1) & 1]);
      acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      mma_bf16(acc[nt], ka[0], qf[nt & 1][0], qf[nt & 1][1]);
      mma_bf16(acc[nt], ka[1], qf[nt & 1][2], qf[nt & 1][3]);
    }
    uint32_t gf[2][4];
    ldsm_x4_t(gb + arow * Cfg::kPitch + acolb, gf[0]);
    const float cs0 = c0 * tau2, cs1 = c1 * tau2;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const uint4 w = lds128(rl_u + (8 * nt + 2 * t_) * 8);  // (r_i, lse_i, r_i+1, lse_i+1)
      const float rx = __uint_as_float(w.x), lx = __uint_as_float(w.y), ry = __uint_as_float(w.z), ly = __uint_as_float(w.w);
      acc[nt][0] = fmaf(acc[nt][0] * cs0, rx, bias2[nt + 1][0]) - lx;
      acc[nt][1] = fmaf(acc[nt][1] * cs0, ry, bias2[nt + 1][1]) - ly;
      acc[nt][2] = fmaf(acc[nt][2] * cs1, rx, bias2[nt][0]) - lx;
      acc[nt][3] = fmaf(acc[nt][3] * cs1, ry, bias2[nt][1]) - ly;
    }
    if (g.shift > 0) {
      const bool bottom = wh == nWh - 1, right = ww == g.nWw - 1;
      if (bottom || right) mb.apply(acc, bottom, right, kNeg);
    }
    uint32_t pa[4][4];  // P^T as A fragments (rows = keys, k = queries)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      pa[nt >> 1][2 * (nt & 1)] = pack_bf16x2(ex2(acc[nt][0]), ex2(acc[nt][1]));
      pa[nt >> 1][2 * (nt & 1) + 1] = pack_bf16x2(ex2(acc[nt][2]), ex2(acc[nt][3]));
    }
    // --- dV = P^T dO  (own 16 keys x 32)
    {
      float dv[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int idx = 2 * ks + half;
          if (idx + 1 < 8)
            ldsm_x4_t(gb + (16 * ((idx + 1) >> 1) + arow) * Cfg::kPitch + ((idx + 1) & 1) * 32 + acolb, gf[(idx + 1) & 1]);
          mma_bf16(dv[2 * half], pa[ks], gf[idx & 1][0], gf[idx & 1][1]);
          mma_bf16(dv[2 * half + 1], pa[ks], gf[idx & 1][2], gf[idx & 1][3]);
        }
      store_tile_bf16<Cfg::kPitch>(dv, ost, g_, t_, dqkv + (int64_t)tok0 * (3 * g.C) + 2 * g.C + head * 32, dqkv + (int64_t)tok1 * (3 * g.C) + 2 * g.C + head * 32);
    }
    // --- dP^T - D = V dO^T + 1 (-D)^T : the last term is one more k-step against the (hi, lo) split of -D
    {
      uint32_t va[2][4];
      ldsm_x4(vb_ + own, va[0]);
      ldsm_x4(vb_ + own + 32, va[1]);
      ldsm_x4(gb + brow * Cfg::kPitch + bcolb, gf[0]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt + 1 < 8) ldsm_x4(gb + (8 * (nt + 1) + brow) * Cfg::kPitch + bcolb, gf[(nt + 1) & 1]);
        uint32_t dneg;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(dneg) : "r"(dhl_u + (8 * nt + g_) * 4) : "memory");
        acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        mma_bf16(acc[nt], va[0], gf[nt & 1][0], gf[nt & 1][1]);
        mma_bf16(acc[nt], va[1], gf[nt & 1][2], gf[nt & 1][3]);
        mma_bf16(acc[nt], ones_a, dneg, dneg);
      }
    }
    // --- dS^T = P^T o (dP^T - D) in packed bf16; d(bias) += dS^T via identity-selector MMAs;
    //     then the two scaled copies: dS * c_j (to shared memory, A operand of dQ) and dS * r_i (registers, A of dK)
    {
      const uint32_t cp0 = pack_bf16x2(c0, c0), cp1 = pack_bf16x2(c1, c1);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int nt = 2 * ks + hf;
          pa[ks][2 * hf] = hmul2_bf16(pa[ks][2 * hf], pack_bf16x2(acc[nt][0], acc[nt][1]));
          pa[ks][2 * hf + 1] = hmul2_bf16(pa[ks][2 * hf + 1], pack_bf16x2(acc[nt][2], acc[nt][3]));
        }
        mma_bf16(dbacc[2 * ks], pa[ks], sel, 0u);
        mma_bf16(dbacc[2 * ks + 1], pa[ks], 0u, sel);
        uint32_t dsc[4];
        dsc[0] = hmul2_bf16(pa[ks][0], cp0);
        dsc[1] = hmul2_bf16(pa[ks][1], cp1);
        dsc[2] = hmul2_bf16(pa[ks][2], cp0);
        dsc[3] = hmul2_bf16(pa[ks][3], cp1);
        stsm_x4(dsT + (16 * wk + arow) * Cfg::kDsPitch + ks * 32 + acolb, dsc);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t rp;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(rp) : "r"(rb_u + (8 * (2 * ks + hf) + 2 * t_) * 2) : "memory");
          pa[ks][2 * hf] = hmul2_bf16(pa[ks][2 * hf], rp);
          pa[ks][2 * hf + 1] = hmul2_bf16(pa[ks][2 * hf + 1], rp);
        }
      }
    }
    // --- dK: M = (dS r)^T Q, then dk = tau c (M - c^2 (k.M) k)   (scale and L2-normalisation backward)
    {
      float dk[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f;
      ldsm_x4_t(qb + arow * Cfg::kPitch + acolb, qf[0]);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int idx = 2 * ks + half;
          if (idx + 1 < 8)
            ldsm_x4_t(qb + (16 * ((idx + 1) >> 1) + arow) * Cfg::kPitch + ((idx + 1) & 1) * 32 + acolb, qf[(idx + 1) & 1]);
          mma_bf16(dk[2 * half], pa[ks], qf[idx & 1][0], qf[idx & 1][1]);
          mma_bf16(dk[2 * half + 1], pa[ks], qf[idx & 1][2], qf[idx & 1][3]);
        }
      ldsm_x4(kb_ + own, ka[0]);
      ldsm_x4(kb_ + own + 32, ka[1]);
      float e0 = 0.f, e1 = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        e0 += dot2(ka[ks][0], dk[2 * ks][0], dk[2 * ks][1]) + dot2(ka[ks][2], dk[2 * ks + 1][0], dk[2 * ks + 1][1]);
        e1 += dot2(ka[ks][1], dk[2 * ks][2], dk[2 * ks][3]) + dot2(ka[ks][3], dk[2 * ks + 1][2], dk[2 * ks + 1][3]);
      }
      const float tc0 = tau_h * c0, tc1 = tau_h * c1;
      e0 = -quad_sum(e0) * c0 * c0 * tc0;
      e1 = -quad_sum(e1) * c1 * c1 * tc1;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        dk[2 * ks][0] = fmaf(dk[2 * ks][0], tc0, e0 * bf16lo_to_f32(ka[ks][0]));
        dk[2 * ks][1] = fmaf(dk[2 * ks][1], tc0, e0 * bf16hi_to_f32(ka[ks][0]));
        dk[2 * ks + 1][0] = fmaf(dk[2 * ks + 1][0], tc0, e0 * bf16lo_to_f32(ka[ks][2]));
        dk[2 * ks + 1][1] = fmaf(dk[2 * ks + 1][1], tc0, e0 * bf16hi_to_f32(ka[ks][2]));
        dk[2 * ks][2] = fmaf(dk[2 * ks][2], tc1, e1 * bf16lo_to_f32(ka[ks][1]));
        dk[2 * ks][3] = fmaf(dk[2 * ks][3], tc1, e1 * bf16hi_to_f32(ka[ks][1]));
        dk[2 * ks + 1][2] = fmaf(dk[2 * ks + 1][2], tc1, e1 * bf16lo_to_f32(ka[ks][3]));
        dk[2 * ks + 1][3] = fmaf(dk[2 * ks + 1][3], tc1, e1 * bf16hi_to_f32(ka[ks][3]));
      }
      store_tile_bf16<Cfg::kPitch>(dk, ost, g_, t_, dqkv + (int64_t)tok0 * (3 * g.C) + g.C + head * 32, dqkv + (int64_t)tok1 * (3 * g.C) + g.C + head * 32);
    }
    named_bar_sync(1 + hh, 128);  // dS c of all 64 keys is in shared memory

    // --- dQ: M = (dS c) K for own 16 queries, dq = tau r (M - r^2 (q.M) q) ; d(tau) += r (q.M)
    {
      float dq[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
      uint32_t da[2][4];
      // A = dS[i][j] read transposed from dsT[j][i]: matrix m -> j half (m>>1), i half (m&1)
      const uint32_t da_addr = dsT + ((lane & 7) + 8 * (lane >> 4)) * Cfg::kDsPitch + (16 * wk + 8 * ((lane >> 3) & 1)) * 2;
      ldsm_x4_t(da_addr, da[0]);
      ldsm_x4_t(kb_ + arow * Cfg::kPitch + acolb, qf[0]);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (ks + 1 < 4) ldsm_x4_t(da_addr + 16 * (ks + 1) * Cfg::kDsPitch, da[(ks + 1) & 1]);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int idx = 2 * ks + half;
          if (idx + 1 < 8)
            ldsm_x4_t(kb_ + (16 * ((idx + 1) >> 1) + arow) * Cfg::kPitch + ((idx + 1) & 1) * 32 + acolb, qf[(idx + 1) & 1]);
          mma_bf16(dq[2 * half], da[ks & 1], qf[idx & 1][0], qf[idx & 1][1]);
          mma_bf16(dq[2 * half + 1], da[ks & 1], qf[idx & 1][2], qf[idx & 1][3]);
        }
      }
      uint32_t qa[2][4];
      ldsm_x4(qb + own, qa[0]);
      ldsm_x4(qb + own + 32, qa[1]);
      float e0 = 0.f, e1 = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        e0 += dot2(qa[ks][0], dq[2 * ks][0], dq[2 * ks][1]) + dot2(qa[ks][2], dq[2 * ks + 1][0], dq[2 * ks + 1][1]);
        e1 += dot2(qa[ks][1], dq[2 * ks][2], dq[2 * ks][3]) + dot2(qa[ks][3], dq[2 * ks + 1][2], dq[2 * ks + 1][3]);
      }
      e0 = quad_sum(e0);
      e1 = quad_sum(e1);
      if (t_ == 0) dtau_acc += fmaf(e0, r0, e1 * r1);
      const float tr0 = tau_h * r0, tr1 = tau_h * r1;
      e0 *= -r0 * r0 * tr0;
      e1 *= -r1 * r1 * tr1;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        dq[2 * ks][0] = fmaf(dq[2 * ks][0], tr0, e0 * bf16lo_to_f32(qa[ks][0]));
        dq[2 * ks][1] = fmaf(dq[2 * ks][1], tr0, e0 * bf16hi_to_f32(qa[ks][0]));
        dq[2 * ks + 1][0] = fmaf(dq[2 * ks + 1][0], tr0, e0 * bf16lo_to_f32(qa[ks][2]));
        dq[2 * ks + 1][1] = fmaf(dq[2 * ks + 1][1], tr0, e0 * bf16hi_to_f32(qa[ks][2]));
        dq[2 * ks][2] = fmaf(dq[2 * ks][2], tr1, e1 * bf16lo_to_f32(qa[ks][1]));
        dq[2 * ks][3] = fmaf(dq[2 * ks][3], tr1, e1 * bf16hi_to_f32(qa[ks][1]));
        dq[2 * ks + 1][2] = fmaf(dq[2 * ks + 1][2], tr1, e1 * bf16lo_to_f32(qa[ks][3]));
        dq[2 * ks + 1][3] = fmaf(dq[2 * ks + 1][3], tr1, e1 * bf16hi_to_f32(qa[ks][3]));
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        qsum[2 * nt] += dq[nt][0] + dq[nt][2];
        qsum[2 * nt + 1] += dq[nt][1] + dq[nt][3];
      }
      store_tile_bf16<Cfg::kPitch>(dq, ost, g_, t_, dqkv + (int64_t)tok0 * (3 * g.C) + head * 32, dqkv + (int64_t)tok1 * (3 * g.C) + head * 32);
      if (lane == 0) mbar_arrive(bar_empty + 8 * ring.s);  // staging lives in the stage: release it only now
    }
  }

  // ---- end of kernel: fold this CTA's d(bias) accumulators onto the 225-row table (fixed summation order),
  //      d(tau) and the dq / dv column sums.  The stage ring is idle by now (every load was consumed).
  named_bar_sync(9, Cfg::kWarps * 32);
  float* bins = reinterpret_cast<float*>(smem);  // [HG][4 wk][32 slots (nt*4 + c)][32 lanes]
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int c = 0; c < 4; ++c) bins[((hh * 4 + wk) * 32 + nt * 4 + c) * 32 + lane] = dbacc[nt][c];
  // column sums: reduce over g (lanes with equal t), lanes 0-3 then hold columns 8nt + 2t + e
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) qsum[i] += __shfl_xor_sync(0xffffffffu, qsum[i], o);
  }
  dtau_acc = warp_sum(dtau_acc);
  float* dtau_s = reinterpret_cast<float*>(smem + Cfg::kOffRed);
  float* csum_s = dtau_s + Cfg::kWarps;  // [HG][4 wk][32]
  if (lane == 0) dtau_s[warp] = dtau_acc;
  if (lane < 4) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) csum_s[(hh * 4 + wk) * 32 + 8 * nt + 2 * lane + e] = qsum[2 * nt + e];
  }
  named_bar_sync(9, Cfg::kWarps * 32);
  const int tid_h = wk * 32 + lane;
  for (int r = tid_h; r < kTab; r += 128) {
    const int dh = r / 15 - 7, dw = r % 15 - 7;  // (ih - jh, iw - jw)
    float sum = 0.f;
    for (int jh = 0; jh < kWs; ++jh) {
      const int ih = jh + dh;
      if (ih < 0 || ih >= kWs) continue;
      for (int jw = 0; jw < kWs; ++jw) {
        const int iw = jw + dw;
        if (iw < 0 || iw >= kWs) continue;
        sum += bins[((hh * 4 + (jh >> 1)) * 32 + ih * 4 + (jh & 1) * 2 + (iw & 1)) * 32 + jw * 4 + (iw >> 1)];
      }
    }
    ws_dbias[((int64_t)cta * g.heads + head) * kTab + r] = sum;
  }
  if (tid_h == 0) ws_dtau[cta * g.heads + head] = dtau_s[4 * hh] + dtau_s[4 * hh + 1] + dtau_s[4 * hh + 2] + dtau_s[4 * hh + 3];
  if (tid_h < 32) {
    float s = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < 4; ++w2) s += csum_s[(hh * 4 + w2) * 32 + tid_h];
    ws_colsum[(int64_t)cta * g.C + head * 32 + tid_h] = s;
  }
}

// Sum the per-CTA partials: one thread per (head, table row), per (head) for d(tau), per channel for the dq column sums.
__global__ void wattn_mma64_reduce_kernel(const float* __restrict__ ws_dbias, const float* __restrict__ ws_dtau,
                                          const float* __restrict__ ws_colsum, int nparts, int heads, int C,
                                          float* __restrict__ dbias_table, float* __restrict__ dtau,
                                          float* __restrict__ dq_colsum) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int per_head = kTab + 1;
  const int n_tab = heads * per_head;
  if (idx < n_tab) {
    const int head = idx / per_head, r = idx - head * per_head;
    float s = 0.f;
    if (r < kTab) {
      for (int c = 0; c < nparts; ++c) s += ws_dbias[((int64_t)c * heads + head) * kTab + r];
      dbias_table[r * heads + head] = s;
    } else {
      for (int c = 0; c < nparts; ++c) s += ws_dtau[c * heads + head];
      dtau[head] = s;
    }
  } else if (idx < n_tab + C && dq_colsum != nullptr) {
    const int k = idx - n_tab;
    float s = 0.f;
    for (int c = 0; c < nparts; ++c) s += ws_colsum[(int64_t)c * C + k];
    dq_colsum[k] = s;
  }
}

int pick_hg(int heads) {
  static const int forced = []() { const char* e = getenv("HV_ATTN_HEADS_PER_CTA"); return e ? atoi(e) : 0; }();
  if (forced >= 1 && forced <= 3 && heads % forced == 0) return forced;  // tuning knob for experiments
  return heads % 3 == 0 ? 3 : (heads % 2 == 0 ? 2 : 1);
}

int ctas_per_group(const Geom& g, int hg) {
  const int nHG = g.heads / hg;
  int per = num_sms() / nHG;
  if (per < 1) per = 1;
  const int nrows = g.B * g.nW;
  return per < nrows ? per : nrows;
}

template <int HG>
int launch_fwd(const Geom& g, const void* qkv, const float* bias_table, const float* tau, void* out, float* lse,
               cudaStream_t st) {
  using Cfg = FwdCfg<HG>;
  auto kern = wattn_mma64_fwd_kernel<HG>;
  static thread_local int attr_dev = -1;  // once per device and thread: keeps the call out of CUDA-graph captures
  int dev = 0;
  HV_CUDA_OK(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    HV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    attr_dev = dev;
  }
  const int per = ctas_per_group(g, HG);
  const int grid = per * (g.heads / HG);
  kern<<<grid, Cfg::kThreads, Cfg::kSmem, st>>>((const bf16*)qkv, bias_table, tau, (bf16*)out, lse, g, per);
  HV_LAUNCH_OK("wattn_mma64_fwd_kernel");
  return HV_OK;
}

template <int HG>
int launch_bwd(const Geom& g, const void* qkv, const void* out, const void* dout, const float* lse, const float* bias_table,
               const float* tau, void* dqkv, float* dbias_table, float* dtau, float* dq_colsum, float* ws, cudaStream_t st) {
  using Cfg = BwdCfg<HG>;
  auto kern = wattn_mma64_bwd_kernel<HG>;
  static thread_local int attr_dev = -1;
  int dev = 0;
  HV_CUDA_OK(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    HV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    attr_dev = dev;
  }
  const int per = ctas_per_group(g, HG);
  const int grid = per * (g.heads / HG);
  float* ws_dbias = ws;
  float* ws_dtau = ws_dbias + (size_t)per * g.heads * kTab;
  float* ws_colsum = ws_dtau + (size_t)per * g.heads;
  kern<<<grid, Cfg::kThreads, Cfg::kSmem, st>>>((const bf16*)qkv, (const bf16*)out, (const bf16*)dout, lse, bias_table, tau,
                                                (bf16*)dqkv, ws_dbias, ws_dtau, ws_colsum, g, per);
  HV_LAUNCH_OK("wattn_mma64_bwd_kernel");
  const int n = g.heads * (kTab + 1) + g.C;
  wattn_mma64_reduce_kernel<<<(n + 127) / 128, 128, 0, st>>>(ws_dbias, ws_dtau, ws_colsum, per, g.heads, g.C, dbias_table,
                                                            dtau, dq_colsum);
  HV_LAUNCH_OK("wattn_mma64_reduce_kernel");
  return HV_OK;
}

}  // namespace

int wattn_mma64_heads_per_cta(int heads) { return pick_hg(heads); }

bool wattn_mma64_supported(const Geom& g, int dtype) {
  // token index must fit 31 bits and the per-image window count 16 bits (Cursor / window_rc arithmetic)
  return dtype == HV_BF16 && g.ws == kWs && g.d == 32 && g.C % 32 == 0 && g.nW < 65536 &&
         (int64_t)g.B * g.H * g.W < (int64_t(1) << 31);
}

size_t wattn_mma64_bwd_workspace_bytes(const Geom& g) {
  // per-CTA partials (bias table, tau, dq/dv column sums); sized for the largest grid (one CTA per SM)
  return (size_t)num_sms() * (g.heads * (kTab + 1) + g.C) * sizeof(float) + 256;
}

int wattn_mma64_fwd(const Geom& g, const void* qkv, const float* bias_table, const float* tau, const float* mask,
                    int mask_windows, void* out, float* lse, cudaStream_t st) {
  (void)mask; (void)mask_windows;
  if (!aligned16(qkv) || !aligned16(out)) HV_FAIL(HV_ERR_ALIGN, "window_attn: qkv/out must be 16-byte aligned");
  switch (pick_hg(g.heads)) {
    case 3: return launch_fwd<3>(g, qkv, bias_table, tau, out, lse, st);
    case 2: return launch_fwd<2>(g, qkv, bias_table, tau, out, lse, st);
    default: return launch_fwd<1>(g, qkv, bias_table, tau, out, lse, st);
  }
}

int wattn_mma64_bwd(const Geom& g, const void* qkv, const void* out, const void* dout, const float* lse,
                    const float* bias_table, const float* tau, const float* mask, int mask_windows, void* dqkv,
                    float* dbias_table, float* dtau, float* dq_colsum, void* workspace, size_t workspace_bytes,
                    cudaStream_t st) {
  (void)mask; (void)mask_windows;
  if (!aligned16(qkv) || !aligned16(out) || !aligned16(dout) || !aligned16(dqkv) || !aligned16(lse))
    HV_FAIL(HV_ERR_ALIGN, "window_attn_bwd: tensors must be 16-byte aligned");
  if (workspace == nullptr || workspace_bytes < wattn_mma64_bwd_workspace_bytes(g))
    HV_FAIL(HV_ERR_WORKSPACE, "window_attn_bwd: workspace of %zu bytes required", wattn_mma64_bwd_workspace_bytes(g));
  float* ws = static_cast<float*>(workspace);
  switch (pick_hg(g.heads)) {
    case 3: return launch_bwd<3>(g, qkv, out, dout, lse, bias_table, tau, dqkv, dbias_table, dtau, dq_colsum, ws, st);
    case 2: return launch_bwd<2>(g, qkv, out, dout, lse, bias_table, tau, dqkv, dbias_table, dtau, dq_colsum, ws, st);
    default: return launch_bwd<1>(g, qkv, out, dout, lse, bias_table, tau, dqkv, dbias_table, dtau, dq_colsum, ws, st);
  }
}

}  // namespace hv
