// tcgen05 / TMEM / TMA forward kernel of the fused shifted-window scaled-cosine attention for 16 x 16 windows (N = 256),
// head dim 32, bf16 -- SwinV2-B at window 16 (BASELINE configs[3]); reference swinv2.py:221-261.  qkv (B, H*W, 3C) and
// out (B, H*W, C) in IMAGE token order; statistics for the backward in three planes (lse in log2 units | r_i = 1 / |q_i|
// | c_j = tau log2e / |k_j|), TILE order (hv_tc_win16.cuh).
//
//   * a unit = one (window, head): q, k, v tiles of 256 rows x 64 B (SWIZZLE_64B) loaded by 4-D TMA boxes whose
//     coordinates are the cyclic shift + window partition; the epilogue writes the normalised output over the q tile and
//     one warp hands it to `cp.async.bulk.tensor` stores with the same boxes (window_reverse + un-roll);
//   * an item = one 128-query block (a column part of the window) of a unit: S = Q_blk K^T is ONE M = 128, N = 256
//     tcgen05.mma chain (2 k-steps) into 256 TMEM columns; two S buffers fill the 512 columns.  P never touches shared
//     memory: the softmax threads write it back over their own logits as packed bf16 (tcgen05.st) and O = P V (M = 128,
//     N = 32, 16 k-steps) takes its A operand from tensor memory; O lands in 32 dead logit columns of the same buffer;
//   * sixteen softmax warps: a thread owns a quarter of a logit row (64 keys = one 8 x 8 sub-block of the window):
//     tcgen05.ld 32 columns at a time, scale by 1/|q_i| * tau/|k_j|, Toeplitz position bias from the 31 x 31 table of the
//     CTA's head, shift mask (all-or-nothing per quarter row, see hv_tc_win16.cuh), exp2, bf16 pack, tcgen05.st.  Heads whose logit range
//     provably fits fp32 skip the row maximum; the others read S twice from TMEM (maximum, then exponentials) instead of
//     holding 64 logits in registers;
//   * a CTA serves one head (bias table and tau fixed) and walks windows first, first + cph, ...; three stages of 48 KB;
//     28 warps: 0 TMA loads | 1 S issuer | 2 PV issuer | 3 TMA stores | 4-7 row norms | 8-23 softmax | 24-27 epilogue.
#ifndef HV_WAIT_HINT_NS
#define HV_WAIT_HINT_NS 1000
#endif
#include "hv_tc_win16.cuh"
#include <atomic>

// sleep (ns) between polls of the off-path warps: measured slower than tight try_wait loops (0.201 vs 0.195 ms at the
// SwinV2-B stage-0 shape), so the default is none
#ifndef HV_FWD_SLEEP
#define HV_FWD_SLEEP(ns) 0
#endif

namespace hv {
namespace {
using namespace tc;

constexpr int kStage = 3 * kTile16;    // q k v; the q tile ends its life as the o tile
constexpr int kStages = 3;
constexpr int kThreads = 896;          // 28 warps
constexpr float kNoMaxRange = 64.0f;

// ---- shared memory map (dynamic, 1024-byte aligned base; no static shared memory in this kernel)
constexpr int kOffStage = 0;
constexpr int kOffBias = kOffStage + kStages * kStage;      // four alignment copies of [31][40] float: log2e * table (reversed columns) - off
constexpr int kOffVec = kOffBias + kBiasFloats16 * 4;       // [kStages][2: r, c][256] float, tile order
constexpr int kOffLsum = kOffVec + kStages * 2 * 256 * 4;   // [2 items in flight][4 quarters][128] float: partial row sums
constexpr int kOffMx = kOffLsum + 2 * 4 * 128 * 4;          // [2][128] float: off + row maximum (for the lse)
constexpr int kOffHmx = kOffMx + 2 * 128 * 4;               // [4 quarters][128] float: quarter-row maxima (use_max heads)
constexpr int kOffGeo = kOffHmx + 4 * 128 * 4;              // [4] UnitGeo16
constexpr int kOffMisc = kOffGeo + 4 * 16;                  // off, use_max
constexpr int kOffBar = kOffMisc + 16;
constexpr int kNumBars = 4 * kStages + 8;
constexpr int kOffTmem = kOffBar + kNumBars * 8;
constexpr int kSmem = kOffTmem + 16;
static_assert(kOffBias % 16 == 0 && kOffVec % 16 == 0 && kOffBar % 8 == 0, "shared-memory alignment");
static_assert(kSmem <= 227 * 1024, "shared memory budget");

// Two S buffers of 256 columns.  The softmax thread of key quarter qt writes its 64 probabilities back as 32 packed bf16
// pairs over the first half of its OWN 64 logit columns (64 qt ..): the A operand of O = P V is read from tensor memory,
// 8 columns per k-step.  O lands in columns 32-63 of the buffer (the dead second half of quarter 0's logits).
constexpr int kTmemCols = 512;
constexpr int kColO = 32;

struct FwdParams {
  Geom g;
  int cph, total;  // CTAs per head, windows (B * nW)
  int64_t plane;   // floats per statistics plane: B * nW * heads * 256
};
struct FwdMaps { CUtensorMap m[2][2]; };  // qkv, out

__global__ void __launch_bounds__(kThreads, 1)
wattn_tc256_fwd_kernel(const __grid_constant__ FwdMaps maps, const float* __restrict__ bias_table, const float* __restrict__ tau,
                       float* __restrict__ stats, const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const Geom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar0 = sb + kOffBar;
  auto bar_full = [&](int s) { return bar0 + 8 * s; };                     // q, k, v tiles landed
  auto bar_empty = [&](int s) { return bar0 + 8 * (kStages + s); };        // the stores of o have read the stage
  auto bar_norm = [&](int s) { return bar0 + 8 * (2 * kStages + s); };     // row scales r, c of the unit exist
  auto bar_written = [&](int s) { return bar0 + 8 * (3 * kStages + s); };  // epilogue wrote o over the q tile (both items)
  const uint32_t barx = bar0 + 8 * 4 * kStages;
  auto bar_s = [&](int b) { return barx + 8 * b; };            // S accumulator buffer b complete
  auto bar_o = [&](int b) { return barx + 8 * (2 + b); };      // O (columns 0-31 of buffer b) complete
  auto bar_ofree = [&](int b) { return barx + 8 * (4 + b); };  // ... and pulled out of TMEM by the epilogue
  auto bar_p = [&](int b) { return barx + 8 * (6 + b); };      // P of the item written back to tensor memory (buffer b)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmem);
  float* misc = reinterpret_cast<float*>(smem + kOffMisc);

  Work16 work;
  work.init(p.cph, p.total);
  const int head = work.head, nunits = work.nunits, nitems = 2 * work.nunits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);    // the store warp
      mbar_init(bar_norm(s), 4);     // the four norm warps
      mbar_init(bar_written(s), 8);  // the four epilogue warps, two items
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_s(b), 1);
      mbar_init(bar_o(b), 1);
      mbar_init(bar_ofree(b), 4);
    }
    mbar_init(bar_p(0), 16);
    mbar_init(bar_p(1), 16);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == 0) {
    // Logits are tau2 * cos + bias2 with |cos| <= 1: if 2 * tau2 + (bias range) stays far inside the fp32 exponent range the
    // row maximum is skipped and exp2(logit - (tau2 + max bias)) is used directly
    float bmx = -3.0e38f, bmn = 3.0e38f;
    for (int q = lane; q < kTab16 * kTab16; q += 32) {
      const float v = kLog2e * __ldg(&bias_table[q * g.heads + head]);
      bmx = fmaxf(bmx, v);
      bmn = fminf(bmn, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      bmx = fmaxf(bmx, __shfl_xor_sync(0xffffffffu, bmx, o));
      bmn = fminf(bmn, __shfl_xor_sync(0xffffffffu, bmn, o));
    }
    const float tau2 = __ldg(&tau[head]) * kLog2e;
    const bool um = !(2.0f * tau2 + (bmx - bmn) <= kNoMaxRange);
    if (lane == 0) {
      misc[0] = um ? 0.f : tau2 + bmx;
      misc[1] = um ? 1.f : 0.f;
    }
  }
  __syncthreads();
  fill_bias16(reinterpret_cast<float*>(smem + kOffBias), bias_table, g.heads, head, kLog2e, misc[0], threadIdx.x, kThreads);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const bool use_max = misc[1] != 0.f;
  const float off = misc[0];
  UnitGeo16* geo = reinterpret_cast<UnitGeo16*>(smem + kOffGeo);
  float* vecs = reinterpret_cast<float*>(smem + kOffVec);
  float* lsum = reinterpret_cast<float*>(smem + kOffLsum);
  float* mxv = reinterpret_cast<float*>(smem + kOffMx);

  if (warp < 4) {
    reg_dealloc<40>();
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer (one elected lane, warp-uniform operands)
      for (int u = 0; u < nunits; ++u) {
        const int s = u % kStages;
        mbar_wait_sleep(bar_empty(s), ((u / kStages) & 1) ^ 1, HV_FWD_SLEEP(256));
        const UnitGeo16 ug = unit_geo16(g, work.first + u * work.stride);
        if (lane == 0) geo[u & 3] = ug;
        __syncwarp();
        if (elect_one()) {
          mbar_expect_tx(bar_full(s), kStage);
#pragma unroll 1
          for (int part = 0; part < 3; ++part) {  // q, k, v
            const int c0 = part * g.C + head * 32;
            const uint32_t dst = sb + kOffStage + s * kStage + part * kTile16;
            for_each_box16(g, ug, [&](int boff, int mi, int col, int row) {
              tma_load_4d(dst + boff, &maps.m[0][mi], bar_full(s), c0, col, row, ug.b);
            });
          }
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- issuer of S = Q_blk K^T (M = 128, N = 256, two k-steps)
      const uint32_t id = idesc_bf16(128, 256, 0, 0);
      const uint64_t d_q = smem_desc(sb + kOffStage, 16, 512, 4), d_k = smem_desc(sb + kOffStage + kTile16, 16, 512, 4);
      for (int n = 0; n < nitems; ++n) {
        const int u = n >> 1, a = n & 1, s = u % kStages, buf = n & 1;
        if (a == 0) mbar_wait_fast(bar_full(s), (u / kStages) & 1);
        if (n > 1) mbar_wait_fast(bar_ofree(buf), ((n >> 1) - 1) & 1);  // the epilogue of item n-2 has drained this buffer
        tc_fence_after();
        if (elect_one()) {
          const uint64_t so = (uint64_t)((s * kStage) >> 4), ao = (uint64_t)((a * 128 * 64) >> 4);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
            umma_ss(tmem + 256 * buf, d_q + so + ao + 2 * kk, d_k + so + 2 * kk, id, kk > 0);
          umma_commit(bar_s(buf));
        }
        __syncwarp();
      }
    } else if (warp == 2) {
      // ---------------------------------------------------------------- issuer of O = P V (M = 128, N = 32, sixteen k-steps)
      const uint32_t id_o = idesc_bf16(128, 32, 0, 1);  // A = P (tensor memory, 128 lanes x 8 columns per k-step), B = v (MN-major)
      // B, MN-major view of the 256 x 64-byte v tile: 32 channels = one 64-byte atom, 8 tokens = 512 B (SBO)
      const uint64_t b_v = smem_desc(sb + kOffStage + 2 * kTile16, 16, 512, 4);
      for (int n = 0; n < nitems; ++n) {
        const int u = n >> 1, s = u % kStages, buf = n & 1;
        if ((n & 1) == 0) mbar_wait_fast(bar_full(s), (u / kStages) & 1);
        mbar_wait_fast(bar_p(buf), (n >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t so = (uint64_t)((s * kStage) >> 4);
          const uint32_t tb = tmem + 256 * buf;
#pragma unroll
          for (int ks = 0; ks < 16; ++ks)  // 16 keys per step: A = 8 columns of key quarter ks / 4, B += 1 KB
            umma_ts(tb + kColO, tb + 64 * (ks >> 2) + 8 * (ks & 3), b_v + so + (uint64_t)(64 * ks), id_o, ks > 0);
          umma_commit(bar_o(buf));
        }
        __syncwarp();
      }
    } else {
      // ---------------------------------------------------------------- warp 3: TMA stores of o (written over the q tile)
      for (int u = 0; u < nunits; ++u) {
        const int s = u % kStages;
        mbar_wait_sleep(bar_written(s), (u / kStages) & 1, HV_FWD_SLEEP(128));
        const uint32_t src = sb + kOffStage + s * kStage;
        if (elect_one()) {
          const UnitGeo16 ug = geo[u & 3];
          for_each_box16(g, ug, [&](int boff, int mi, int col, int row) {
            tma_store_4d(&maps.m[1][mi], src + boff, head * 32, col, row, ug.b);
          });
        }
        __syncwarp();
        bulk_commit();
        bulk_wait_read0();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty(s));
      }
      bulk_wait0();
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ norm warps: thread x owns rows x and x + 128 of q and of k
    const int x = 32 * (warp - 4) + lane;
    const float tau2 = __ldg(&tau[head]) * kLog2e;
    for (int u = 0; u < nunits; ++u) {
      const int s = u % kStages;
      mbar_wait_sleep(bar_full(s), (u / kStages) & 1, HV_FWD_SLEEP(128));
      const int widx = geo[u & 3].flags >> 2;
      float* vec = vecs + s * 512;
      float* sp = stats + p.plane + ((int64_t)widx * g.heads + head) * kN16;
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        const uint32_t tile = sb + kOffStage + s * kStage + part * kTile16;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int row = x + 128 * hh;
          float ss = 0.f;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const uint4 v = lds128(tile + row * 64 + ((ch ^ ((row >> 1) & 3)) << 4));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float lo = bf16lo_to_f32(w[e]), hi = bf16hi_to_f32(w[e]);
              ss = fmaf(lo, lo, ss);
              ss = fmaf(hi, hi, ss);
            }
          }
          const float sc = (part == 0 ? 1.0f : tau2) * inv_norm(ss);
          vec[part * 256 + row] = sc;
          sp[part * p.plane + row] = sc;  // planes 1: r, 2: c, for the backward kernel
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_norm(s));
    }
  } else if (warp < 24) {
    // ------------------------------------------------------------------ softmax threads: a thread owns a quarter of a logit row.
    // Query = tile row t of the item's block (window row ih, column 8 a + iw8); keys = block qt of the tile (column part
    // qt / 2, window rows 8 (qt & 1) ..), TMEM columns 64 qt ..
    reg_alloc<80>();
    const int quad = warp & 3, qt = (warp - 8) >> 2;
    const int t = 32 * quad + lane, ih = t >> 3, iw8 = t & 7;
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16) + 64 * qt;
    const float kNeg = kMaskValue * kLog2e;
    const float* bt = reinterpret_cast<const float*>(smem + kOffBias);
    float* hmx = reinterpret_cast<float*>(smem + kOffHmx);
    // bias run of key (jl, jw8) of this thread's key block for query block 0
    const float* bp0 = bias_run16(bt, (ih - 8 * (qt & 1) + 15) * kBiasStride16 + (15 - iw8 + 8 * (qt >> 1)));

    for (int n = 0; n < nitems; ++n) {
      const int u = n >> 1, a = n & 1, s = u % kStages, buf = n & 1;
      mbar_wait_fast(bar_norm(s), (u / kStages) & 1);  // row scales (and with them the tiles) of the unit exist
      const int flags = geo[u & 3].flags;
      const float* vec = vecs + s * 512;
      const float ri = vec[128 * a + t];
      const float* cv = vec + 256 + 64 * qt;
      // bias of key (jl, jw8) of the block: run of 8 floats at bp - 40 jl (16-byte aligned through the alignment copies)
      const float* bp = bp0 - 8 * a;  // a multiple of 8: same alignment copy
      // shift mask: the whole quarter row sits on the other side of a wrap than the query, or none of it (warp-uniform)
      const bool masked = ((flags & 1) && ((ih >= 8) != ((qt & 1) != 0))) || ((flags & 2) && (a != (qt >> 1)));
      const float madd = masked ? kNeg : 0.f;
      const float2 ri2 = make_float2(ri, ri);
      mbar_wait_fast(bar_s(buf), (n >> 1) & 1);
      tc_fence_after();
      const uint32_t tS = tl + 256 * buf;
      float mx = 0.f;
      if (use_max) {
        float2 mx2 = make_float2(-3.0e38f, -3.0e38f);
        uint32_t accm[2][16];
        HV_TMEM_LD16(tS, accm[0]);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // same software pipeline as the exponentiation pass below
          uint32_t (&acc)[16] = accm[c & 1];
          HV_REG_FENCE16(acc);
          if (c < 3) HV_TMEM_LD16(tS + 16 * (c + 1), accm[(c + 1) & 1]);
#pragma unroll
          for (int r2 = 0; r2 < 2; ++r2) {
            const int jl = 2 * c + r2;
            const float4 c0 = *reinterpret_cast<const float4*>(cv + 8 * jl), c1 = *reinterpret_cast<const float4*>(cv + 8 * jl + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(bp - kBiasStride16 * jl), b1 = *reinterpret_cast<const float4*>(bp - kBiasStride16 * jl + 4);
            const float2 cc[4] = {make_float2(c0.x, c0.y), make_float2(c0.z, c0.w), make_float2(c1.x, c1.y), make_float2(c1.z, c1.w)};
            const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 a2 = make_float2(__uint_as_float(acc[8 * r2 + 2 * e]), __uint_as_float(acc[8 * r2 + 2 * e + 1]));
              const float2 x2 = f2fma(f2mul(a2, ri2), cc[e], bb[e]);
              mx2.x = fmaxf(mx2.x, x2.x);
              mx2.y = fmaxf(mx2.y, x2.y);
            }
          }
          if (c < 3) tmem_wait_ld();
        }
        mx = fmaxf(mx2.x, mx2.y) + madd;
        hmx[qt * 128 + t] = mx;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory");
        mx = fmaxf(fmaxf(hmx[t], hmx[128 + t]), fmaxf(hmx[256 + t], hmx[384 + t]));
        asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory");  // all four owners have read before the next item writes
      }
      const float sub = madd - mx;
      const float2 sub2 = make_float2(sub, sub);
      float2 ls2 = make_float2(0.f, 0.f);
      // four chunks of 16 keys (two window rows), software-pipelined: the tcgen05.ld of chunk c + 1 is in flight while chunk c
      // is exponentiated; its 8 packed pairs go back into columns 8 c .. of the quarter (logits already in registers)
      uint32_t accs[2][16];
      HV_TMEM_LD16(tS, accs[0]);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t (&acc)[16] = accs[c & 1];
        HV_REG_FENCE16(acc);
        if (c < 3) HV_TMEM_LD16(tS + 16 * (c + 1), accs[(c + 1) & 1]);
        uint32_t pk[8];
#pragma unroll
        for (int r2 = 0; r2 < 2; ++r2) {
          const int jl = 2 * c + r2;
          const float4 c0 = *reinterpret_cast<const float4*>(cv + 8 * jl), c1 = *reinterpret_cast<const float4*>(cv + 8 * jl + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(bp - kBiasStride16 * jl), b1 = *reinterpret_cast<const float4*>(bp - kBiasStride16 * jl + 4);
          const float2 cc[4] = {make_float2(c0.x, c0.y), make_float2(c0.z, c0.w), make_float2(c1.x, c1.y), make_float2(c1.z, c1.w)};
          const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 a2 = make_float2(__uint_as_float(acc[8 * r2 + 2 * e]), __uint_as_float(acc[8 * r2 + 2 * e + 1]));
            float2 x2 = f2fma(f2mul(a2, ri2), cc[e], bb[e]);
            if (use_max || masked) x2 = f2add(x2, sub2);
            const float2 p2 = make_float2(ex2(x2.x), ex2(x2.y));
            ls2 = f2add(ls2, p2);
            pk[4 * r2 + e] = pack_bf16x2(p2.x, p2.y);
          }
        }
        HV_TMEM_ST8(tS + 8 * c, pk);
        if (c < 3) tmem_wait_ld();
      }
      const float ls = ls2.x + ls2.y;
      lsum[((n & 1) * 4 + qt) * 128 + t] = ls;
      if (qt == 0) mxv[(n & 1) * 128 + t] = off + mx;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p(buf));
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps: O from TMEM, normalise, write the row as
    // bf16 over the q tile (the store warp sends the tile out by TMA), lse to the statistics
    const int quad = warp & 3, t = 32 * quad + lane;
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16);
    const uint32_t swz = (uint32_t)((t >> 1) & 3);
    for (int n = 0; n < nitems; ++n) {
      const int u = n >> 1, a = n & 1, s = u % kStages, buf = n & 1;
      mbar_wait_sleep(bar_o(buf), (n >> 1) & 1, HV_FWD_SLEEP(64));
      tc_fence_after();
      uint32_t o[32];
      HV_TMEM_LD32(tl + 256 * buf + kColO, o);
      const float* lp = lsum + (n & 1) * 512 + t;
      const float l = (lp[0] + lp[128]) + (lp[256] + lp[384]);
      const float lse_off = mxv[(n & 1) * 128 + t];
      const int widx = geo[u & 3].flags >> 2;
      tmem_wait_ld();
      HV_REG_FENCE32(o);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ofree(buf));
      const float inv = rcp_fast(l);
      const uint32_t orow = sb + kOffStage + s * kStage + (128 * a + t) * 64;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(o[8 * q + 0]) * inv, __uint_as_float(o[8 * q + 1]) * inv);
        v.y = pack_bf16x2(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv);
        v.z = pack_bf16x2(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv);
        v.w = pack_bf16x2(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv);
        sts128(orow + ((q ^ swz) << 4), v);
      }
      stats[((int64_t)widx * g.heads + head) * kN16 + 128 * a + t] = lse_off + lg2_fast(l);
      fence_async_smem();  // the tile is read by the TMA engine (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_written(s));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
}

std::atomic<int> g_tc256_mode{-1};  // -1: HV_ATTN_TC256 environment (unset: on), 0: off (generic kernels), 1: on

}  // namespace

int wattn_tc256_variant_set(int v) {
  const int old = g_tc256_mode.exchange(v, std::memory_order_relaxed);
  return old;
}

bool wattn_tc256_supported(const Geom& g, int dtype) {
  static const int env = []() { const char* e = getenv("HV_ATTN_TC256"); return e == nullptr ? 1 : (atoi(e) != 0 ? 1 : 0); }();
  const int cur = g_tc256_mode.load(std::memory_order_relaxed);
  const int mode = cur < 0 ? env : cur;
  if (mode == 0) return false;
  // shift 0 or ws / 2 (the only shift SwinV2 uses, swinv2.py:560): the column parts of the tile order are the halves of the wrap
  return dtype == HV_BF16 && g.ws == kWin16 && g.d == 32 && g.C % 32 == 0 && (g.shift == 0 || g.shift == 8) &&
         g.heads <= num_sms() && (int64_t)g.B * g.H * g.W < (int64_t(1) << 31) && (int64_t)g.B * g.nW < (int64_t(1) << 28);
}

int wattn_tc256_fwd(const Geom& g, const void* qkv, const float* bias_table, const float* tau, void* out, float* stats,
                    cudaStream_t st) {
  if (!aligned16(qkv) || !aligned16(out) || !aligned16(stats))
    HV_FAIL(HV_ERR_ALIGN, "window_attn: qkv / out / statistics must be 16-byte aligned");
  struct MapKey { const void *qkv, *out; int B, H, W, C; };
  struct MapEntry { MapKey key; FwdMaps maps; };
  static thread_local MapEntry cache[32];
  static thread_local int cache_n = 0, cache_next = 0;
  const MapKey key = {qkv, out, g.B, g.H, g.W, g.C};
  const FwdMaps* mp = nullptr;
  for (int i = 0; i < cache_n; ++i) {
    const MapKey& c = cache[i].key;
    if (c.qkv == key.qkv && c.out == key.out && c.B == key.B && c.H == key.H && c.W == key.W && c.C == key.C) { mp = &cache[i].maps; break; }
  }
  if (!mp) {
    MapEntry& e = cache[cache_next];
    int rc = make_window_maps16(e.maps.m[0], qkv, g, 3 * g.C);
    if (rc) return rc;
    rc = make_window_maps16(e.maps.m[1], out, g, g.C);
    if (rc) return rc;
    e.key = key;
    mp = &e.maps;
    cache_next = (cache_next + 1) % 32;
    if (cache_n < 32) ++cache_n;
  }
  FwdParams p;
  p.g = g;
  p.total = g.B * g.nW;
  p.cph = num_sms() / g.heads;
  if (p.cph > p.total) p.cph = p.total;
  if (p.cph < 1) p.cph = 1;
  p.plane = (int64_t)g.B * g.nW * g.heads * kN16;
  static thread_local int attr_dev = -1;
  int dev = 0;
  HV_CUDA_OK(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    HV_CUDA_OK(cudaFuncSetAttribute(wattn_tc256_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_dev = dev;
  }
  wattn_tc256_fwd_kernel<<<g.heads * p.cph, kThreads, kSmem, st>>>(*mp, bias_table, tau, stats, p);
  HV_LAUNCH_OK("wattn_tc256_fwd_kernel");
  return HV_OK;
}

// Host: image token feeding tile row t of window `win` of image b in the TILE order of the N = 256 kernels -- the same
// unit_geo16 / tile_row_rc16 arithmetic the kernels' TMA coordinates come from (for bit-exact checks against the reference's
// torch.roll + window_partition, swinv2.py:399-412)
void window16_tile_token_index(const Geom& g, int64_t* out) {
  for (int widx = 0; widx < g.B * g.nW; ++widx) {
    const tc::UnitGeo16 ug = tc::unit_geo16(g, widx);
    for (int t = 0; t < tc::kN16; ++t) {
      int row, col;
      tc::tile_row_rc16(g, ug, t, row, col);
      out[(size_t)widx * tc::kN16 + t] = ((int64_t)ug.b * g.H + row) * g.W + col;
    }
  }
}

}  // namespace hv
