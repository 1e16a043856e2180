// Tensor-core fused shifted-window scaled-cosine attention for the SwinV2 hot shape:
// window 8x8 (N = 64 tokens), head dim 32, bf16 activations (every stage of SwinV2-T, stage 3 of -B).
//
// Replaces reference swinv2.py:399-412 (roll + window_partition), 221-261 (head split, cosine logits,
// logit scale, position bias, shift mask, softmax, attn @ v, head merge), 420-429 (window_reverse +
// roll back) and their autograd.  Data layout: qkv (B, H*W, 3C) and out (B, H*W, C) stay in IMAGE token
// order in HBM; the window gather / scatter is address arithmetic in the TMA producer and the epilogue.
//
// Structure (persistent CTAs, one per SM, HG heads per CTA):
//   * 4 producer warps (one per SM sub-partition): per window, 16-byte `cp.async` (LDGSTS) copies straight
//     from the rolled image position of every token into a padded, bank-conflict-free shared-memory
//     tile; completion is signalled on an mbarrier (`cp.async.mbarrier.arrive.noinc`), multi-stage
//     full/empty ring.  (Measured on B200, tools/probes/tma_probe.cu: `cp.async.bulk` requests of one
//     token row segment (64-576 B) cost ~60-90 issue cycles each per warp and top out at 2.7 TB/s from
//     one warp, while LDGSTS from 4 warps reaches the 6.7 TB/s copy ceiling; bulk/TMA only wins for
//     >= 4 KB contiguous requests, which a partitioned head group of a rolled window never has.)
//   * 4 compute warps per head: each owns 16 query rows (forward) / 16 key rows (backward) of one
//     (window, head) and keeps S/P entirely in registers (mma.sync m16n8k16 bf16, fp32 accumulate);
//     the position bias lives in registers (forward) or shared memory (backward) for the whole
//     kernel, the shift mask is two 16-bit patterns per thread, softmax uses ex2.
//   * backward recomputes S from q,k and the saved row log-sum-exp; dS~ goes through shared memory
//     once (bf16) to be re-read transposed for dQ; d(bias) is accumulated in registers over all
//     windows of the CTA and folded to the ((2ws-1)^2, heads) table deterministically.
// The kernel is HBM-bound (37 FLOP/B): the design goal is bytes in flight, not MMA rate.
#include <stdlib.h>

#include "hv_common.cuh"

namespace hv {
namespace {

constexpr int kN = 64;       // tokens per window
constexpr int kWs = 8;       // window side
constexpr int kTab = 225;    // (2*8-1)^2 bias-table rows
constexpr int kOstPitch = 80;                 // bytes per row of the per-warp 16x32 bf16 staging tile
constexpr int kOstBytes = 16 * kOstPitch;

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
__device__ __forceinline__ float sq2(uint32_t w) {  // sum of squares of a packed bf16 pair
  const float a = bf16lo_to_f32(w), b = bf16hi_to_f32(w);
  return fmaf(a, a, b * b);
}
__device__ __forceinline__ float dot2(uint32_t w, float x, float y) {  // packed pair . (x, y)
  return fmaf(bf16lo_to_f32(w), x, bf16hi_to_f32(w) * y);
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
__device__ __forceinline__ float inv_norm(float ss) { return rsqrtf(fmaxf(ss, 1e-24f)); }

template <int HG> struct FwdCfg {
  static constexpr int kWarps = 4 * HG;             // compute warps
  static constexpr int kProducers = 4;              // producer warps
  static constexpr int kThreads = (kWarps + kProducers) * 32;
  static constexpr int kPitch = HG * 192 + 16;  // [q | k | v] x HG heads (64 B each) + 16 B pad: odd multiple of 16
  static constexpr int kCpr = 12 * HG;          // 16-byte chunks per token row
  static constexpr int kRowInstr = kWs * kCpr / 32;  // full-warp LDGSTS per window row (8 tokens)
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
    // Each producer warp owns two of the window's eight token rows.  One window row = 8 tokens x kCpr
    // 16-byte chunks = kRowInstr full-warp LDGSTS instructions; the (token, chunk) a lane handles in the
    // m-th of them is the same for every window, so its shared/global offsets are precomputed once.
    const int pw = warp - Cfg::kWarps;
    int soff[Cfg::kRowInstr], gpack[Cfg::kRowInstr];
#pragma unroll
    for (int m = 0; m < Cfg::kRowInstr; ++m) {
      const int q = lane + 32 * m;
      const int iw = q / Cfg::kCpr, c = q - iw * Cfg::kCpr;
      const int part = c / (4 * HG), within = c - part * (4 * HG);
      soff[m] = iw * Cfg::kPitch + c * 16;
      gpack[m] = (part * g.C + hgrp * (HG * 32) + within * 8) | (iw << 24);
    }
    const int tok_stride = 3 * g.C;
    const float inv_nWw = 1.0f / (float)g.nWw;
    Cursor cur;
    cur.init(g, cta, ctas_per_group);
    int it = 0;
    for (int row = cta; row < nrows; row += ctas_per_group, ++it, cur.next(g)) {
      const int s = it % Cfg::kStages;
      const uint32_t ph = (it / Cfg::kStages) & 1;
      int wh, ww;
      window_rc(g, inv_nWw, cur.win, wh, ww);
      const int col0 = ww * kWs + g.shift;
      if (pw == 0) mbar_wait(bar_empty + 8 * s, ph ^ 1);  // one waiter; the other producers sleep in the barrier
      named_bar_sync(10, Cfg::kProducers * 32);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int ih = 2 * pw + r;
        int irow = wh * kWs + g.shift + ih;
        if (irow >= g.H) irow -= g.H;
        const bf16* rowp = qkv + ((int64_t)cur.b * g.H + irow) * g.W * tok_stride;
        const uint32_t dst = sbase + s * Cfg::kStageBytes + ih * (kWs * Cfg::kPitch);
#pragma unroll
        for (int m = 0; m < Cfg::kRowInstr; ++m) {
          int col = col0 + (gpack[m] >> 24);
          if (col >= g.W) col -= g.W;
          cp_async16(dst + soff[m], rowp + col * tok_stride + (gpack[m] & 0xffffff));
        }
      }
      cp_async_arrive(bar_full + 8 * s);
    }
    return;
  }

  // -------------------------------------------------------------------- compute warps
  const int hh = warp >> 2, wq = warp & 3;
  const int head = hgrp * HG + hh;
  const int g_ = lane >> 2, t_ = lane & 3;
  const int i0 = 16 * wq + g_, i1 = i0 + 8;  // own query slots
  float bias2[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = 8 * nt + 2 * t_ + e;
      bias2[nt][e] = kLog2e * __ldg(&bias_table[rel_pos_index(kWs, i0, j) * g.heads + head]);
      bias2[nt][2 + e] = kLog2e * __ldg(&bias_table[rel_pos_index(kWs, i1, j) * g.heads + head]);
    }
  const float tau2 = __ldg(&tau[head]) * kLog2e;
  const int hi_thr = kWs - g.shift;
  uint32_t colH, colW;
  column_band_bits(t_, hi_thr, colH, colW);
  const bool r0H = (2 * wq) >= hi_thr, r1H = (2 * wq + 1) >= hi_thr, rW = g_ >= hi_thr;
  const int nWh = g.H / kWs;
  const float kNeg = kMaskValue * kLog2e;

  const uint32_t ost = sbase + Cfg::kOffOst + warp * kOstBytes;
  float* cvec_base = reinterpret_cast<float*>(smem + Cfg::kOffCvec);
  const int arow = lane_row16(lane), acolb = (lane >> 4) * 16;
  const int brow = lane & 7, bcolb = (lane >> 3) * 16;

  const float inv_nWw = 1.0f / (float)g.nWw;
  const uint32_t ones_b = (g_ == 0) ? 0x3F803F80u : 0u;  // B fragment of a ones column: row sums of P by MMA
  Cursor cur;
  cur.init(g, cta, ctas_per_group);
  int it = 0;
  for (int row = cta; row < nrows; row += ctas_per_group, ++it, cur.next(g)) {
    const int s = it % Cfg::kStages;
    const uint32_t ph = (it / Cfg::kStages) & 1;
    int wh, ww;
    window_rc(g, inv_nWw, cur.win, wh, ww);
    const int row0 = wh * kWs + g.shift, col0 = ww * kWs + g.shift;
    if (wq == 0) mbar_wait(bar_full + 8 * s, ph);  // one waiter per head; the rest sleep in the barrier
    named_bar_sync(1 + hh, 128);
    const uint32_t st = sbase + s * Cfg::kStageBytes;
    const uint32_t qb = st + hh * 64, kb_ = st + HG * 64 + hh * 64, vb_ = st + 2 * HG * 64 + hh * 64;

    // --- inverse norms of this warp's 16 key rows -> shared vector (scaled by tau*log2e)
    float* cvec = cvec_base + ((it & 1) * HG + hh) * kN;
    {
      uint32_t ka[2][4];
      ldsm_x4(kb_ + (16 * wq + arow) * Cfg::kPitch + acolb, ka[0]);
      ldsm_x4(kb_ + (16 * wq + arow) * Cfg::kPitch + acolb + 32, ka[1]);
      float s0 = sq2(ka[0][0]) + sq2(ka[0][2]) + sq2(ka[1][0]) + sq2(ka[1][2]);
      float s1 = sq2(ka[0][1]) + sq2(ka[0][3]) + sq2(ka[1][1]) + sq2(ka[1][3]);
      s0 = quad_sum(s0);
      s1 = quad_sum(s1);
      if (t_ == 0) {
        cvec[i0] = tau2 * inv_norm(s0);
        cvec[i1] = tau2 * inv_norm(s1);
      }
    }
    named_bar_sync(1 + hh, 128);

    // --- Q fragments + own-row inverse norms
    uint32_t qa[2][4];
    ldsm_x4(qb + (16 * wq + arow) * Cfg::kPitch + acolb, qa[0]);
    ldsm_x4(qb + (16 * wq + arow) * Cfg::kPitch + acolb + 32, qa[1]);
    float r0 = quad_sum(sq2(qa[0][0]) + sq2(qa[0][2]) + sq2(qa[1][0]) + sq2(qa[1][2]));
    float r1 = quad_sum(sq2(qa[0][1]) + sq2(qa[0][3]) + sq2(qa[1][1]) + sq2(qa[1][3]));
    r0 = inv_norm(r0);
    r1 = inv_norm(r1);

    // --- S = Q K^T (raw dot products)
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      uint32_t kf[4];
      ldsm_x4(kb_ + (8 * nt + brow) * Cfg::kPitch + bcolb, kf);
      mma_bf16(acc[nt], qa[0], kf[0], kf[1]);
      mma_bf16(acc[nt], qa[1], kf[2], kf[3]);
    }
    // --- logits in the log2 domain: tau*cos + bias (+ mask)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 c = *reinterpret_cast<const float2*>(&cvec[8 * nt + 2 * t_]);
      acc[nt][0] = fmaf(acc[nt][0] * r0, c.x, bias2[nt][0]);
      acc[nt][1] = fmaf(acc[nt][1] * r0, c.y, bias2[nt][1]);
      acc[nt][2] = fmaf(acc[nt][2] * r1, c.x, bias2[nt][2]);
      acc[nt][3] = fmaf(acc[nt][3] * r1, c.y, bias2[nt][3]);
    }
    if (g.shift > 0) {
      const bool bottom = wh == nWh - 1, right = ww == g.nWw - 1;
      if (bottom || right) {
        const uint32_t m0 = (bottom ? (r0H ? ~colH : colH) : 0u) | (right ? (rW ? ~colW : colW) : 0u);
        const uint32_t m1 = (bottom ? (r1H ? ~colH : colH) : 0u) | (right ? (rW ? ~colW : colW) : 0u);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (m0 & (1u << (2 * nt + e))) acc[nt][e] += kNeg;
            if (m1 & (1u << (2 * nt + e))) acc[nt][2 + e] += kNeg;
          }
      }
    }
    // --- softmax (rows g, g+8)
    float mx0 = acc[0][0], mx1 = acc[0][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx0 = fmaxf(mx0, fmaxf(acc[nt][0], acc[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(acc[nt][2], acc[nt][3]));
    }
    mx0 = quad_max(mx0);
    mx1 = quad_max(mx1);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      acc[nt][0] = ex2(acc[nt][0] - mx0);
      acc[nt][1] = ex2(acc[nt][1] - mx0);
      acc[nt][2] = ex2(acc[nt][2] - mx1);
      acc[nt][3] = ex2(acc[nt][3] - mx1);
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
        uint32_t vf[4];
        ldsm_x4_t(vb_ + (16 * ks + arow) * Cfg::kPitch + half * 32 + acolb, vf);
        mma_bf16(o[2 * half], pa, vf[0], vf[1]);
        mma_bf16(o[2 * half + 1], pa, vf[2], vf[3]);
      }
      mma_bf16(lacc, pa, ones_b, ones_b);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * s);  // the stage is free for the producer

    // column 0 of the ones tile sits in the t == 0 lane of every quad
    const float l0 = __shfl_sync(0xffffffffu, lacc[0], lane & ~3);
    const float l1 = __shfl_sync(0xffffffffu, lacc[2], lane & ~3);
    const float inv0 = __fdividef(1.0f, l0), inv1 = __fdividef(1.0f, l1);
    if (t_ == 0) {
      float* lp = lse + ((int64_t)row * g.heads + head) * kN;
      lp[i0] = (mx0 + __log2f(l0)) * kLn2;
      lp[i1] = (mx1 + __log2f(l1)) * kLn2;
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      o[nt][0] *= inv0; o[nt][1] *= inv0;
      o[nt][2] *= inv1; o[nt][3] *= inv1;
    }
    // own rows: slot i0 = (ih 2*wq, iw g_), i1 = (ih 2*wq+1, iw g_)
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
  static constexpr int kCpr = 20 * HG;          // 16-byte chunks per token row
  static constexpr int kRowInstr = kWs * kCpr / 32;  // full-warp LDGSTS per window row (8 tokens)
  static constexpr int kPitch = HG * 320 + 16;  // [q | k | v | o | dO] x HG heads + pad: odd multiple of 16
  static constexpr int kLseOff = kN * kPitch;   // HG x 64 fp32 row log-sum-exp behind the token rows
  static constexpr int kStageBytes = kN * kPitch + HG * kN * 4;
  static constexpr int kStages = 3;
  static constexpr int kDsPitch = 144;                      // bytes per dS~ row (64 bf16 + 16 B pad)
  static constexpr int kOffDs = kStages * kStageBytes;      // [HG][64 j][64 i] bf16; at the end: d(bias) partials
  static constexpr int kOffVec = kOffDs + HG * kN * kDsPitch;   // r[HG][64], D[HG][64] float
  static constexpr int kOffTau = kOffVec + 2 * HG * kN * 4;     // per-warp d(tau) partials
  static constexpr int kOffBar = kOffTau + ((kWarps * 4 + 15) / 16) * 16;
  static constexpr int kSmem = kOffBar + 2 * kStages * 8;
};

template <int HG>
__global__ void __launch_bounds__(BwdCfg<HG>::kThreads, 1)
wattn_mma64_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ out, const bf16* __restrict__ dout,
                       const float* __restrict__ lse, const float* __restrict__ bias_table, const float* __restrict__ tau,
                       bf16* __restrict__ dqkv, float* __restrict__ ws_dbias, float* __restrict__ ws_dtau, Geom g,
                       int ctas_per_group) {
  using Cfg = BwdCfg<HG>;
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
    // q,k,v (from qkv), o (from out), dO (from dout) token segments and the row log-sum-exp of the window;
    // same lane-constant chunk schedule as the forward producer (two window rows per producer warp).
    const int pw = warp - Cfg::kWarps;
    // per lane-chunk: shared offset | token column, the tensor it reads (base pointer incl. channel offset)
    // and that tensor's token stride, so that one chunk costs a wrap test, one wide multiply-add and the copy
    int spack[Cfg::kRowInstr], stride_b[Cfg::kRowInstr];
    const char* gbase[Cfg::kRowInstr];
#pragma unroll
    for (int m = 0; m < Cfg::kRowInstr; ++m) {
      const int q = lane + 32 * m;
      const int iw = q / Cfg::kCpr, c = q - iw * Cfg::kCpr;
      const int part = c / (4 * HG), within = c - part * (4 * HG);
      const int ch = hgrp * (HG * 32) + within * 8;
      spack[m] = (iw * Cfg::kPitch + c * 16) | (iw << 24);
      if (part < 3) {
        gbase[m] = reinterpret_cast<const char*>(qkv + part * g.C + ch);
        stride_b[m] = 6 * g.C;
      } else {
        gbase[m] = reinterpret_cast<const char*>((part == 3 ? out : dout) + ch);
        stride_b[m] = 2 * g.C;
      }
    }
    const float inv_nWw = 1.0f / (float)g.nWw;
    Cursor cur;
    cur.init(g, cta, ctas_per_group);
    int it = 0;
    for (int row = cta; row < nrows; row += ctas_per_group, ++it, cur.next(g)) {
      const int s = it % Cfg::kStages;
      const uint32_t ph = (it / Cfg::kStages) & 1;
      int wh, ww;
      window_rc(g, inv_nWw, cur.win, wh, ww);
      const int col0 = ww * kWs + g.shift;
      if (pw == 0) mbar_wait(bar_empty + 8 * s, ph ^ 1);  // one waiter; the other producers sleep in the barrier
      named_bar_sync(10, Cfg::kProducers * 32);
      const uint32_t st = sbase + s * Cfg::kStageBytes;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int ih = 2 * pw + r;
        int irow = wh * kWs + g.shift + ih;
        if (irow >= g.H) irow -= g.H;
        const int rowtok = (cur.b * g.H + irow) * g.W;  // token index < 2^31 (checked on the host)
        const uint32_t dst = st + ih * (kWs * Cfg::kPitch);
#pragma unroll
        for (int m = 0; m < Cfg::kRowInstr; ++m) {
          int col = col0 + (spack[m] >> 24);
          if (col >= g.W) col -= g.W;
          cp_async16(dst + (spack[m] & 0xffffff), gbase[m] + (int64_t)(rowtok + col) * stride_b[m]);
        }
      }
      if (pw == 0) {
        const float* lrow = lse + ((int64_t)row * g.heads + hgrp * HG) * kN;
        for (int idx = lane; idx < HG * kN / 4; idx += 32) cp_async16(st + Cfg::kLseOff + idx * 16, lrow + idx * 4);
      }
      cp_async_arrive(bar_full + 8 * s);
    }
    return;
  }

  // -------------------------------------------------------------------- compute warps
  const int hh = warp >> 2, wk = warp & 3;
  const int head = hgrp * HG + hh;
  const int g_ = lane >> 2, t_ = lane & 3;
  const int j0 = 16 * wk + g_, j1 = j0 + 8;  // own key slots in the S^T pass == own query slots in the dQ pass
  const float tau_h = __ldg(&tau[head]);
  const float tau2 = tau_h * kLog2e;
  const int hi_thr = kWs - g.shift;
  uint32_t colH, colW;
  column_band_bits(t_, hi_thr, colH, colW);
  const bool r0H = (2 * wk) >= hi_thr, r1H = (2 * wk + 1) >= hi_thr, rW = g_ >= hi_thr;
  const int nWh = g.H / kWs;
  const float kNeg = kMaskValue * kLog2e;

  const uint32_t dsT = sbase + Cfg::kOffDs + hh * kN * Cfg::kDsPitch;
  // The bias is block-Toeplitz in (ih - jh, iw - jw).  For this thread's rows (key jh = 2*wk + rh, jw = g) and
  // columns (query ih = nt, iw = 2t + e) only d = nt - rh + 1 in [0, 8] and e in {0, 1} vary, so 18 registers hold
  // every bias value it will ever need, and 18 more accumulate d(bias) directly in table-bin space.
  float bias2[9][2], dbias[9][2];
#pragma unroll
  for (int d = 0; d < 9; ++d)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int r = (d - 1 - 2 * wk + 7) * 15 + (2 * t_ + e - g_ + 7);
      bias2[d][e] = kLog2e * __ldg(&bias_table[r * g.heads + head]);
      dbias[d][e] = 0.f;
    }
  float* rvec = reinterpret_cast<float*>(smem + Cfg::kOffVec) + hh * kN;
  float* dvec = rvec + HG * kN;
  const int arow = lane_row16(lane), acolb = (lane >> 4) * 16;
  const int brow = lane & 7, bcolb = (lane >> 3) * 16;

  float dtau_acc = 0.f;

  const float inv_nWw = 1.0f / (float)g.nWw;
  Cursor cur;
  cur.init(g, cta, ctas_per_group);
  int it = 0;
  for (int row = cta; row < nrows; row += ctas_per_group, ++it, cur.next(g)) {
    const int s = it % Cfg::kStages;
    const uint32_t ph = (it / Cfg::kStages) & 1;
    int wh, ww;
    window_rc(g, inv_nWw, cur.win, wh, ww);
    const int row0 = wh * kWs + g.shift, col0 = ww * kWs + g.shift;
    // own token rows: slot j0 = (ih 2*wk, iw g_), j1 = (ih 2*wk+1, iw g_)
    const int64_t tok0 = tile_token(g, cur.b, row0, col0, 2 * wk, g_);
    const int64_t tok1 = tile_token(g, cur.b, row0, col0, 2 * wk + 1, g_);
    if (wk == 0) mbar_wait(bar_full + 8 * s, ph);  // one waiter per head; the rest sleep in the barrier
    named_bar_sync(1 + hh, 128);
    const uint32_t st = sbase + s * Cfg::kStageBytes;
    const uint32_t qb = st + hh * 64, kb_ = st + HG * 64 + hh * 64, vb_ = st + 2 * HG * 64 + hh * 64;
    const uint32_t ob = st + 3 * HG * 64 + hh * 64, gb = st + 4 * HG * 64 + hh * 64;
    const float* lse_s = reinterpret_cast<const float*>(smem + s * Cfg::kStageBytes + Cfg::kLseOff) + hh * kN;
    const uint32_t own = (16 * wk + arow) * Cfg::kPitch + acolb;
    // output staging: the O segment of this warp's own 16 token rows is dead after the pre-pass below
    const uint32_t ost = ob + (16 * wk) * Cfg::kPitch;

    // --- pre-pass over this warp's 16 token rows: 1/|k|, 1/|q|, D = dO . O
    uint32_t ka[2][4];
    float c0, c1, r0, r1;
    {
      ldsm_x4(kb_ + own, ka[0]);
      ldsm_x4(kb_ + own + 32, ka[1]);
      c0 = quad_sum(sq2(ka[0][0]) + sq2(ka[0][2]) + sq2(ka[1][0]) + sq2(ka[1][2]));
      c1 = quad_sum(sq2(ka[0][1]) + sq2(ka[0][3]) + sq2(ka[1][1]) + sq2(ka[1][3]));
      c0 = inv_norm(c0);
      c1 = inv_norm(c1);
      uint32_t qa[2][4];
      ldsm_x4(qb + own, qa[0]);
      ldsm_x4(qb + own + 32, qa[1]);
      r0 = quad_sum(sq2(qa[0][0]) + sq2(qa[0][2]) + sq2(qa[1][0]) + sq2(qa[1][2]));
      r1 = quad_sum(sq2(qa[0][1]) + sq2(qa[0][3]) + sq2(qa[1][1]) + sq2(qa[1][3]));
      r0 = inv_norm(r0);
      r1 = inv_norm(r1);
      uint32_t oa[4], ga[4];
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        ldsm_x4(ob + own + 32 * ks, oa);
        ldsm_x4(gb + own + 32 * ks, ga);
        d0 += dot2(oa[0], bf16lo_to_f32(ga[0]), bf16hi_to_f32(ga[0])) + dot2(oa[2], bf16lo_to_f32(ga[2]), bf16hi_to_f32(ga[2]));
        d1 += dot2(oa[1], bf16lo_to_f32(ga[1]), bf16hi_to_f32(ga[1])) + dot2(oa[3], bf16lo_to_f32(ga[3]), bf16hi_to_f32(ga[3]));
      }
      d0 = quad_sum(d0);
      d1 = quad_sum(d1);
      if (t_ == 0) {
        rvec[j0] = r0; rvec[j1] = r1;
        dvec[j0] = d0; dvec[j1] = d1;
      }
    }
    named_bar_sync(1 + hh, 128);  // r, D of all 64 rows visible; previous tile's dS~ fully consumed

    // --- S^T = K Q^T for own 16 keys (rows) x 64 queries (columns), then P^T
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      uint32_t qf[4];
      ldsm_x4(qb + (8 * nt + brow) * Cfg::kPitch + bcolb, qf);
      mma_bf16(acc[nt], ka[0], qf[0], qf[1]);
      mma_bf16(acc[nt], ka[1], qf[2], qf[3]);
    }
    const float cs0 = c0 * tau2, cs1 = c1 * tau2;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int i = 8 * nt + 2 * t_;
      const float2 ri = *reinterpret_cast<const float2*>(&rvec[i]);
      acc[nt][0] = fmaf(acc[nt][0] * cs0, ri.x, bias2[nt + 1][0]);
      acc[nt][1] = fmaf(acc[nt][1] * cs0, ri.y, bias2[nt + 1][1]);
      acc[nt][2] = fmaf(acc[nt][2] * cs1, ri.x, bias2[nt][0]);
      acc[nt][3] = fmaf(acc[nt][3] * cs1, ri.y, bias2[nt][1]);
    }
    if (g.shift > 0) {
      const bool bottom = wh == nWh - 1, right = ww == g.nWw - 1;
      if (bottom || right) {
        const uint32_t m0 = (bottom ? (r0H ? ~colH : colH) : 0u) | (right ? (rW ? ~colW : colW) : 0u);
        const uint32_t m1 = (bottom ? (r1H ? ~colH : colH) : 0u) | (right ? (rW ? ~colW : colW) : 0u);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (m0 & (1u << (2 * nt + e))) acc[nt][e] += kNeg;
            if (m1 & (1u << (2 * nt + e))) acc[nt][2 + e] += kNeg;
          }
      }
    }
    uint32_t pa[4][4];  // P^T as A fragments (rows = keys, k = queries)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 ls = *reinterpret_cast<const float2*>(&lse_s[8 * nt + 2 * t_]);
      const float lx = ls.x * kLog2e, ly = ls.y * kLog2e;
      const float p0 = ex2(acc[nt][0] - lx), p1 = ex2(acc[nt][1] - ly);
      const float p2 = ex2(acc[nt][2] - lx), p3 = ex2(acc[nt][3] - ly);
      pa[nt >> 1][2 * (nt & 1)] = pack_bf16x2(p0, p1);
      pa[nt >> 1][2 * (nt & 1) + 1] = pack_bf16x2(p2, p3);
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
          uint32_t gf[4];
          ldsm_x4_t(gb + (16 * ks + arow) * Cfg::kPitch + half * 32 + acolb, gf);
          mma_bf16(dv[2 * half], pa[ks], gf[0], gf[1]);
          mma_bf16(dv[2 * half + 1], pa[ks], gf[2], gf[3]);
        }
      store_tile_bf16<Cfg::kPitch>(dv, ost, g_, t_, dqkv + tok0 * 3 * g.C + 2 * g.C + head * 32, dqkv + tok1 * 3 * g.C + 2 * g.C + head * 32);
    }
    // --- dP^T = V dO^T, dS^T = P^T o (dP^T - D)   (V / K fragments are (re)loaded where they are used:
    // 16 registers less live across the softmax keeps the kernel out of local memory at 128 registers)
    uint32_t va[2][4];
    ldsm_x4(vb_ + own, va[0]);
    ldsm_x4(vb_ + own + 32, va[1]);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      uint32_t gf[4];
      ldsm_x4(gb + (8 * nt + brow) * Cfg::kPitch + bcolb, gf);
      mma_bf16(acc[nt], va[0], gf[0], gf[1]);
      mma_bf16(acc[nt], va[1], gf[2], gf[3]);
    }
    const float ct0 = c0 * tau_h, ct1 = c1 * tau_h;
    uint32_t dsa[4][4];  // dS~^T = dS^T * tau * r_i * c_j as A fragments
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int i = 8 * nt + 2 * t_;
      const float2 di = *reinterpret_cast<const float2*>(&dvec[i]);
      const float2 ri = *reinterpret_cast<const float2*>(&rvec[i]);
      const uint32_t pw0 = pa[nt >> 1][2 * (nt & 1)], pw1 = pa[nt >> 1][2 * (nt & 1) + 1];
      const float ds0 = bf16lo_to_f32(pw0) * (acc[nt][0] - di.x);
      const float ds1 = bf16hi_to_f32(pw0) * (acc[nt][1] - di.y);
      const float ds2 = bf16lo_to_f32(pw1) * (acc[nt][2] - di.x);
      const float ds3 = bf16hi_to_f32(pw1) * (acc[nt][3] - di.y);
      dbias[nt + 1][0] += ds0; dbias[nt + 1][1] += ds1; dbias[nt][0] += ds2; dbias[nt][1] += ds3;
      const uint32_t w0 = pack_bf16x2(ds0 * ct0 * ri.x, ds1 * ct0 * ri.y);
      const uint32_t w1 = pack_bf16x2(ds2 * ct1 * ri.x, ds3 * ct1 * ri.y);
      dsa[nt >> 1][2 * (nt & 1)] = w0;
      dsa[nt >> 1][2 * (nt & 1) + 1] = w1;
      sts32(dsT + j0 * Cfg::kDsPitch + i * 2, w0);
      sts32(dsT + j1 * Cfg::kDsPitch + i * 2, w1);
    }
    // --- dK = dS~^T Q, then the L2-normalisation backward: dk = M - c^2 (k.M) k
    {
      float dk[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t qf[4];
          ldsm_x4_t(qb + (16 * ks + arow) * Cfg::kPitch + half * 32 + acolb, qf);
          mma_bf16(dk[2 * half], dsa[ks], qf[0], qf[1]);
          mma_bf16(dk[2 * half + 1], dsa[ks], qf[2], qf[3]);
        }
      ldsm_x4(kb_ + own, ka[0]);
      ldsm_x4(kb_ + own + 32, ka[1]);
      float e0 = 0.f, e1 = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        e0 += dot2(ka[ks][0], dk[2 * ks][0], dk[2 * ks][1]) + dot2(ka[ks][2], dk[2 * ks + 1][0], dk[2 * ks + 1][1]);
        e1 += dot2(ka[ks][1], dk[2 * ks][2], dk[2 * ks][3]) + dot2(ka[ks][3], dk[2 * ks + 1][2], dk[2 * ks + 1][3]);
      }
      e0 = quad_sum(e0) * c0 * c0;
      e1 = quad_sum(e1) * c1 * c1;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        dk[2 * ks][0] -= e0 * bf16lo_to_f32(ka[ks][0]);     dk[2 * ks][1] -= e0 * bf16hi_to_f32(ka[ks][0]);
        dk[2 * ks + 1][0] -= e0 * bf16lo_to_f32(ka[ks][2]); dk[2 * ks + 1][1] -= e0 * bf16hi_to_f32(ka[ks][2]);
        dk[2 * ks][2] -= e1 * bf16lo_to_f32(ka[ks][1]);     dk[2 * ks][3] -= e1 * bf16hi_to_f32(ka[ks][1]);
        dk[2 * ks + 1][2] -= e1 * bf16lo_to_f32(ka[ks][3]); dk[2 * ks + 1][3] -= e1 * bf16hi_to_f32(ka[ks][3]);
      }
      store_tile_bf16<Cfg::kPitch>(dk, ost, g_, t_, dqkv + tok0 * 3 * g.C + g.C + head * 32, dqkv + tok1 * 3 * g.C + g.C + head * 32);
    }
    named_bar_sync(1 + hh, 128);  // dS~ of all 64 keys is in shared memory

    // --- dQ = dS~ K for own 16 queries, dq = M - r^2 (q.M) q ; d(tau) += q.M
    {
      float dq[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t da[4];
        // A = dS~[i][j] read transposed from dsT[j][i]: matrix m -> j half (m>>1), i half (m&1)
        ldsm_x4_t(dsT + (16 * ks + (lane & 7) + 8 * (lane >> 4)) * Cfg::kDsPitch + (16 * wk + 8 * ((lane >> 3) & 1)) * 2, da);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t kf[4];
          ldsm_x4_t(kb_ + (16 * ks + arow) * Cfg::kPitch + half * 32 + acolb, kf);
          mma_bf16(dq[2 * half], da, kf[0], kf[1]);
          mma_bf16(dq[2 * half + 1], da, kf[2], kf[3]);
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
      if (t_ == 0) dtau_acc += e0 + e1;
      e0 *= r0 * r0;
      e1 *= r1 * r1;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        dq[2 * ks][0] -= e0 * bf16lo_to_f32(qa[ks][0]);     dq[2 * ks][1] -= e0 * bf16hi_to_f32(qa[ks][0]);
        dq[2 * ks + 1][0] -= e0 * bf16lo_to_f32(qa[ks][2]); dq[2 * ks + 1][1] -= e0 * bf16hi_to_f32(qa[ks][2]);
        dq[2 * ks][2] -= e1 * bf16lo_to_f32(qa[ks][1]);     dq[2 * ks][3] -= e1 * bf16hi_to_f32(qa[ks][1]);
        dq[2 * ks + 1][2] -= e1 * bf16lo_to_f32(qa[ks][3]); dq[2 * ks + 1][3] -= e1 * bf16hi_to_f32(qa[ks][3]);
      }
      store_tile_bf16<Cfg::kPitch>(dq, ost, g_, t_, dqkv + tok0 * 3 * g.C + head * 32, dqkv + tok1 * 3 * g.C + head * 32);
      if (lane == 0) mbar_arrive(bar_empty + 8 * s);  // staging lives in the stage: release it only now
    }
  }

  // ---- fold this CTA's d(bias) bins onto the 225-row table (fixed summation order); d(tau)
  named_bar_sync(9, Cfg::kWarps * 32);  // every compute warp is done with the dS~ buffer
  float* bins = reinterpret_cast<float*>(smem + Cfg::kOffDs);  // [HG][4 wk][32 lanes][9][2]
  {
    float* mine = bins + ((hh * 4 + wk) * 32 + lane) * 18;
#pragma unroll
    for (int d = 0; d < 9; ++d) {
      mine[2 * d] = dbias[d][0];
      mine[2 * d + 1] = dbias[d][1];
    }
  }
  dtau_acc = warp_sum(dtau_acc);
  float* dtau_s = reinterpret_cast<float*>(smem + Cfg::kOffTau);
  if (lane == 0) dtau_s[warp] = dtau_acc;
  named_bar_sync(9, Cfg::kWarps * 32);
  const int tid_h = wk * 32 + lane;
  for (int r = tid_h; r < kTab; r += 128) {
    const int dh = r / 15 - 7, dw = r % 15 - 7;  // (ih - jh, iw - jw)
    float sum = 0.f;
    for (int w2 = 0; w2 < 4; ++w2) {
      const int d = dh + 2 * w2 + 1;
      if (d < 0 || d > 8) continue;
      for (int g2 = 0; g2 < 8; ++g2) {
        const int c = dw + g2;  // = 2t + e
        if (c < 0 || c >= kWs) continue;
        sum += bins[((hh * 4 + w2) * 32 + g2 * 4 + (c >> 1)) * 18 + 2 * d + (c & 1)];
      }
    }
    ws_dbias[((int64_t)cta * g.heads + head) * kTab + r] = sum;
  }
  if (tid_h == 0)
    ws_dtau[cta * g.heads + head] =
        (dtau_s[4 * hh] + dtau_s[4 * hh + 1] + dtau_s[4 * hh + 2] + dtau_s[4 * hh + 3]) / tau_h;
}

// Sum the per-CTA partial tables: one thread per (head, table row) and one per (head) for d(tau).
__global__ void wattn_mma64_reduce_kernel(const float* __restrict__ ws_dbias, const float* __restrict__ ws_dtau, int nparts,
                                          int heads, float* __restrict__ dbias_table, float* __restrict__ dtau) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int per_head = kTab + 1;
  if (idx >= heads * per_head) return;
  const int head = idx / per_head, r = idx - head * per_head;
  float s = 0.f;
  if (r < kTab) {
    for (int c = 0; c < nparts; ++c) s += ws_dbias[((int64_t)c * heads + head) * kTab + r];
    dbias_table[r * heads + head] = s;
  } else {
    for (int c = 0; c < nparts; ++c) s += ws_dtau[c * heads + head];
    dtau[head] = s;
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
  HV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
  const int per = ctas_per_group(g, HG);
  const int grid = per * (g.heads / HG);
  kern<<<grid, Cfg::kThreads, Cfg::kSmem, st>>>((const bf16*)qkv, bias_table, tau, (bf16*)out, lse, g, per);
  HV_LAUNCH_OK("wattn_mma64_fwd_kernel");
  return HV_OK;
}

template <int HG>
int launch_bwd(const Geom& g, const void* qkv, const void* out, const void* dout, const float* lse, const float* bias_table,
               const float* tau, void* dqkv, float* dbias_table, float* dtau, float* ws, cudaStream_t st) {
  using Cfg = BwdCfg<HG>;
  auto kern = wattn_mma64_bwd_kernel<HG>;
  HV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
  const int per = ctas_per_group(g, HG);
  const int grid = per * (g.heads / HG);
  float* ws_dbias = ws;
  float* ws_dtau = ws + (size_t)per * g.heads * kTab;
  kern<<<grid, Cfg::kThreads, Cfg::kSmem, st>>>((const bf16*)qkv, (const bf16*)out, (const bf16*)dout, lse, bias_table, tau,
                                                (bf16*)dqkv, ws_dbias, ws_dtau, g, per);
  HV_LAUNCH_OK("wattn_mma64_bwd_kernel");
  const int n = g.heads * (kTab + 1);
  wattn_mma64_reduce_kernel<<<(n + 127) / 128, 128, 0, st>>>(ws_dbias, ws_dtau, per, g.heads, dbias_table, dtau);
  HV_LAUNCH_OK("wattn_mma64_reduce_kernel");
  return HV_OK;
}

}  // namespace

bool wattn_mma64_supported(const Geom& g, int dtype) {
  // token index must fit 31 bits and the per-image window count 16 bits (Cursor / window_rc arithmetic)
  return dtype == HV_BF16 && g.ws == kWs && g.d == 32 && g.C % 32 == 0 && g.nW < 65536 &&
         (int64_t)g.B * g.H * g.W < (int64_t(1) << 31);
}

size_t wattn_mma64_bwd_workspace_bytes(const Geom& g) {
  // per-CTA partial tables; sized for the largest grid (one CTA per SM)
  return (size_t)num_sms() * g.heads * (kTab + 1) * sizeof(float) + 256;
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
                    float* dbias_table, float* dtau, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  (void)mask; (void)mask_windows;
  if (!aligned16(qkv) || !aligned16(out) || !aligned16(dout) || !aligned16(dqkv) || !aligned16(lse))
    HV_FAIL(HV_ERR_ALIGN, "window_attn_bwd: tensors must be 16-byte aligned");
  if (workspace == nullptr || workspace_bytes < wattn_mma64_bwd_workspace_bytes(g))
    HV_FAIL(HV_ERR_WORKSPACE, "window_attn_bwd: workspace of %zu bytes required", wattn_mma64_bwd_workspace_bytes(g));
  float* ws = static_cast<float*>(workspace);
  switch (pick_hg(g.heads)) {
    case 3: return launch_bwd<3>(g, qkv, out, dout, lse, bias_table, tau, dqkv, dbias_table, dtau, ws, st);
    case 2: return launch_bwd<2>(g, qkv, out, dout, lse, bias_table, tau, dqkv, dbias_table, dtau, ws, st);
    default: return launch_bwd<1>(g, qkv, out, dout, lse, bias_table, tau, dqkv, dbias_table, dtau, ws, st);
  }
}

}  // namespace hv
