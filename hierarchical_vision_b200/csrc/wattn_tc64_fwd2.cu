// tcgen05 / TMEM / TMA forward kernel of the fused shifted-window scaled-cosine attention, window 8x8 (N = 64), head dim
// 32, bf16, second generation: the machine mapping of the backward kernel (wattn_tc64_bwd.cu) applied to the forward of
// reference swinv2.py:221-261.  qkv (B, H*W, 3C) and out (B, H*W, C) in IMAGE token order; statistics for the backward
// in three planes (lse in log2 units | r_i = 1 / |q_i| | c_j = tau log2e / |k_j|, window-slot order).
//
//   * cyclic shift + window partition are the coordinates of 4-D TMA tile LOADS (q, k, v tiles of a (window, head) unit:
//     64 rows x 64 B, SWIZZLE_64B); the epilogue writes the normalised output over the q tile of the stage and one warp
//     hands it to `cp.async.bulk.tensor` STORES with the same boxes: window_reverse + un-roll are coordinates too;
//   * S = Q K^T as M = 64 MMAs, one per unit: the two units of a pair interleave in the 128 TMEM lanes and share 64
//     columns (no stacked M = 128 tile whose off-diagonal half is thrown away); O = P V with P staged once to shared
//     memory as a [query][key] bf16 tile (SWIZZLE_128B, the A operand, K-major) and the v tile read MN-major;
//   * two softmax groups (8 warps each) on alternate pairs; a thread owns half a logit row: tcgen05.ld, scale by
//     1/|q_i| * tau/|k_j|, Toeplitz position bias (4 alignment copies of the 15 x 15 table per head, 10 KB instead of four
//     expanded 64 x 68 matrices), shift mask, exp2, pack, 16-byte staging stores; its half-row sum goes to the epilogue
//     through shared memory.  Heads whose logit range provably fits fp32 skip the row maximum (see wattn_mma64.cu);
//   * one token order per CTA through the window classes of hv_tc_win.cuh (interior | bottom row | right edge), seven
//     stages of 24 KB, 28 warps: 0 TMA loads | 1 S issuer | 2 PV issuer | 3 TMA stores | 4-7 row norms (tensor-pipe self
//     products on the swizzled tiles) | 8-23 softmax (2 groups) | 24-27 epilogue.  All hand-overs are mbarriers.
#define HV_WAIT_HINT_NS 1000
#include "hv_tc_win.cuh"

namespace hv {
namespace {
using namespace tc;

constexpr int kN = 64;
constexpr int kWs = 8;
constexpr int kTab = 225;
constexpr int kTile = kN * 64;         // one (window, head) q / k / v tile: 64 rows x 64 B (SWIZZLE_64B)
constexpr int kStage = 6 * kTile;      // q_a q_b k_a k_b v_a v_b; the q tiles end their life as the o tiles
#ifndef HV_FWD2_STAGES
#define HV_FWD2_STAGES 6
#endif
constexpr int kStages = HV_FWD2_STAGES;
constexpr int kThreads = 896;          // 28 warps
constexpr int kPTile = kN * 128;       // P of one unit: 64 rows x 128 B (SWIZZLE_128B)
constexpr int kBiasRow = 20;           // floats per table row (15 + alignment slack)
constexpr int kBiasCopy = 328;         // floats per alignment copy: >= 15 * 20 and = 8 (mod 32) so 8 lanes hit 8 bank groups
constexpr float kNoMaxRange = 64.0f;

// ---- shared memory map (dynamic, 1024-byte aligned base)
constexpr int kOffStage = 0;
constexpr int kOffP = kOffStage + kStages * kStage;         // [2 groups][2 units][64][128 B]
constexpr int kOffBias = kOffP + 4 * kPTile;                // [2 units][4 copies][kBiasCopy] float (log2 units, minus `off`)
constexpr int kOffVec = kOffBias + 2 * 4 * kBiasCopy * 4;   // [kStages][2: r, c][2 units][64] float, TILE order
// softmax -> epilogue hand-over, a ring of 4 pairs: the softmax of pair k + 4 starts after the PV MMAs of pair k + 2 have
// read its staging buffer, and those waited for the epilogue of pair k to drain the O buffer (after it read these)
constexpr int kOffLsum = kOffVec + kStages * 2 * 128 * 4;   // [4][2 halves][128] float: half-row sums of P
constexpr int kOffMx = kOffLsum + 4 * 2 * 128 * 4;          // [4][128] float: off + row maximum (for the lse)
constexpr int kOffHmx = kOffMx + 4 * 128 * 4;               // [2 groups][2 halves][128] float: half-row maxima (use_max heads)
constexpr int kOffGeo = kOffHmx + 2 * 2 * 128 * 4;          // [8][2] UnitGeo (kStages + 1 <= 8 pairs between producer and store warp)
constexpr int kOffSlotMap = kOffGeo + 8 * 2 * 16;           // [64] bytes
constexpr int kOffBar = kOffSlotMap + 64;
constexpr int kNumBars = 4 * kStages + 12;
constexpr int kOffTmem = kOffBar + kNumBars * 8;
constexpr int kSmem = kOffTmem + 16;
static_assert(kOffP % 1024 == 0 && kOffBias % 16 == 0 && kOffVec % 16 == 0 && kOffBar % 8 == 0, "shared-memory alignment");
static_assert(kSmem + 1024 <= 227 * 1024 && kStages <= 7, "shared memory budget");

// TMEM columns: S of a pair 64 columns (two M = 64 MMAs interleaved in the lanes), O 32 columns; two buffers each
constexpr int kColS = 0, kColO = 128, kTmemCols = 256;

struct FwdParams {
  Geom g;
  WinSchedule sched;
  int64_t plane;  // floats per statistics plane: B * nW * heads * 64
};
struct FwdMaps { CUtensorMap m[2][kNumWinMaps]; };  // qkv, out

template <int kMode>
__device__ __forceinline__ void wattn_tc64_fwd2_body(const FwdMaps& maps, const float* __restrict__ bias_table,
                                                     const float* __restrict__ tau, float* __restrict__ stats,
                                                     const FwdParams& p) {
  constexpr bool kSplit = kMode == 2;
  constexpr bool kMasked = kMode > 0;
  extern __shared__ __align__(1024) unsigned char smem[];
  const Geom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar0 = sb + kOffBar;
  auto bar_full = [&](int s) { return bar0 + 8 * s; };                     // q, k, v tiles landed
  auto bar_empty = [&](int s) { return bar0 + 8 * (kStages + s); };        // the stores of o have read the stage
  auto bar_norm = [&](int s) { return bar0 + 8 * (2 * kStages + s); };     // row scales r, c of the stage exist
  auto bar_written = [&](int s) { return bar0 + 8 * (3 * kStages + s); };  // epilogue wrote o over the q tiles
  const uint32_t barx = bar0 + 8 * 4 * kStages;
  auto bar_s = [&](int b) { return barx + 8 * b; };             // S accumulator buffer b complete
  auto bar_sfree = [&](int b) { return barx + 8 * (2 + b); };   // ... and read by its softmax group
  auto bar_staged = [&](int b) { return barx + 8 * (4 + b); };  // P staging buffer b written
  auto bar_stfree = [&](int b) { return barx + 8 * (6 + b); };  // ... and read by the PV MMAs
  auto bar_o = [&](int b) { return barx + 8 * (8 + b); };       // O accumulator buffer b complete
  auto bar_ofree = [&](int b) { return barx + 8 * (10 + b); };  // ... and pulled out of TMEM by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmem);

  __shared__ CtaWork s_work;
  __shared__ float s_off[2];
  __shared__ int s_usemax[2];
  if (threadIdx.x == 0) {
    CtaWork w0;
    w0.init(p.g, p.sched, blockIdx.x);
    s_work = w0;
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);    // the store warp
      mbar_init(bar_norm(s), 4);     // the four norm warps
      mbar_init(bar_written(s), 4);  // the four epilogue warps
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_s(b), 1);
      mbar_init(bar_sfree(b), 8);
      mbar_init(bar_staged(b), 8);
      mbar_init(bar_stfree(b), 1);
      mbar_init(bar_o(b), 1);
      mbar_init(bar_ofree(b), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // ---- one-time tables: slot of every tile row; per unit the offset of the no-maximum softmax; Toeplitz bias copies
  unsigned char* slotmap = smem + kOffSlotMap;
  if (threadIdx.x < 64) slotmap[threadIdx.x] = (unsigned char)tile_row_slot(threadIdx.x, kSplit ? g.shift : 0);
  if (warp < 2) {
    // Logits are tau2 * cos + bias2 with |cos| <= 1: if 2 * tau2 + (bias range) stays far inside the fp32 exponent range the
    // row maximum is skipped and exp2(logit - (tau2 + max bias)) is used directly
    CtaWork w0;
    w0.init(p.g, p.sched, blockIdx.x);
    const int head = warp == 0 ? w0.head_a : w0.head_b;
    float bmx = -3.0e38f, bmn = 3.0e38f;
    for (int q = lane; q < kTab; q += 32) {
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
    const bool use_max = !(2.0f * tau2 + (bmx - bmn) <= kNoMaxRange);
    if (lane == 0) {
      s_usemax[warp] = use_max ? 1 : 0;
      s_off[warp] = use_max ? 0.f : tau2 + bmx;
    }
  }
  __syncthreads();
  {
    const CtaWork w0 = s_work;
    float* bt = reinterpret_cast<float*>(smem + kOffBias);
    for (int idx = threadIdx.x; idx < 2 * 4 * kBiasCopy; idx += kThreads) {
      const int u = idx / (4 * kBiasCopy), rem = idx - u * 4 * kBiasCopy;
      const int c = rem / kBiasCopy, q = rem - c * kBiasCopy;
      const int dh = q / kBiasRow, pos = q - dh * kBiasRow;
      const int x = pos - 4 + c;  // x = 7 - iw + jw: reversed column difference
      float v = 0.f;
      if (dh < 15 && x >= 0 && x <= 14)
        v = kLog2e * __ldg(&bias_table[(dh * 15 + 14 - x) * g.heads + (u == 0 ? w0.head_a : w0.head_b)]) - s_off[u];
      bt[idx] = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const CtaWork work = s_work;
  const int npairs = work.npairs;
  UnitGeo* geo = reinterpret_cast<UnitGeo*>(smem + kOffGeo);
  float* vecs = reinterpret_cast<float*>(smem + kOffVec);
  float* lsum = reinterpret_cast<float*>(smem + kOffLsum);
  float* mxv = reinterpret_cast<float*>(smem + kOffMx);

  if (warp < 4) {
    reg_dealloc<40>();
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer (one elected lane, warp-uniform operands)
      const int nWh = g.H / kWs;
      WinCursor cur;
      cur.init(work.cross ? 2 * work.first : work.first, work.cross ? 2 * work.stride : work.stride, work.wcls, work.hcls);
      for (int k = 0; k < npairs; ++k, cur.advance()) {
        const int s = k % kStages;
        mbar_wait_fast(bar_empty(s), ((k / kStages) & 1) ^ 1);
        int ub[2], urow0[2], ucol0[2];
        bool ubottom[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          int b = cur.b, wh = cur.wh, ww = cur.ww;
          bool valid = true;
          if (u == 1 && work.cross) {
            if (cur.idx + 1 < work.ncls) cur.next(b, wh, ww);
            else valid = false;  // padding unit of an odd tail: a copy of unit 0 whose results are dropped
          }
          if (kMode == 1) wh = nWh - 1;   // class 1: the bottom row of windows
          if (kSplit) ww = g.nWw - 1;     // class 2: the last column
          const int r = (b * nWh + wh) * g.nWw + ww;
          ub[u] = b;
          urow0[u] = wh * kWs + g.shift; ucol0[u] = ww * kWs + g.shift;
          ubottom[u] = kMasked && wh == nWh - 1;
          if (lane == u) {
            UnitGeo ug;
            ug.b = b; ug.row0 = urow0[u]; ug.col0 = ucol0[u];
            ug.rflags = (r << 3) | (kSplit ? 4 : 0) | (ubottom[u] ? 2 : 0) | (valid ? 1 : 0);
            geo[(k & 7) * 2 + u] = ug;
          }
        }
        __syncwarp();
#ifdef HV_FWD2_LANE_ISSUE
        if (lane == 0) mbar_expect_tx(bar_full(s), kStage);
        __syncwarp();
        if (lane < 6) {
#else
        if (elect_one()) {
          mbar_expect_tx(bar_full(s), kStage);
#endif
          auto issue_tile = [&](int t) {
            const int u = t & 1, part = t >> 1;  // part: 0 q, 1 k, 2 v
            const int head = u == 0 ? work.head_a : work.head_b;
            const int c0 = part * g.C + head * 32;
            const uint32_t dst = sb + kOffStage + s * kStage + t * kTile;
            const int bb = u ? ub[1] : ub[0];
            for_each_box<kMode>(g, u ? ucol0[1] : ucol0[0], u ? urow0[1] : urow0[0], u ? ubottom[1] : ubottom[0],
                                [&](int off, int mi, int col, int row) { tma_load_4d(dst + off, &maps.m[0][mi], bar_full(s), c0, col, row, bb); });
          };
#ifdef HV_FWD2_LANE_ISSUE
          issue_tile(lane);
#else
          if (kMasked) {
#pragma unroll 1
            for (int t = 0; t < 6; ++t) issue_tile(t);
          } else {
#pragma unroll
            for (int t = 0; t < 6; ++t) issue_tile(t);
          }
#endif
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- issuer of S = Q K^T (one M = 64 chain per unit)
      const uint32_t id = idesc_bf16(64, 64, 0, 0);
      const uint64_t d_q = smem_desc(sb + kOffStage, 16, 512, 4), d_k = smem_desc(sb + kOffStage + 2 * kTile, 16, 512, 4);
      for (int k = 0; k < npairs; ++k) {
        const int s = k % kStages, buf = k & 1;
        mbar_wait_fast(bar_full(s), (k / kStages) & 1);
        if (k > 1) mbar_wait_fast(bar_sfree(buf), ((k >> 1) - 1) & 1);  // the group of pair k-2 has read this S buffer
        tc_fence_after();
        if (elect_one()) {
          const uint64_t so = (uint64_t)((s * kStage) >> 4);
#pragma unroll
          for (int u = 0; u < 2; ++u) {  // unit u: tiles q_u, k_u (one tile = 4 KB further), accumulator lanes + 16 u
            const uint32_t dl = (uint32_t)(16 * u) << 16;
            const uint64_t uo = (uint64_t)(u * (kTile >> 4));
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
              umma_ss(tmem + dl + kColS + 64 * buf, d_q + so + uo + 2 * kk, d_k + so + uo + 2 * kk, id, kk > 0);
          }
          umma_commit(bar_s(buf));
        }
        __syncwarp();
      }
    } else if (warp == 2) {
      // ---------------------------------------------------------------- issuer of O = P V (one M = 64 chain per unit)
      const uint32_t id_o = idesc_bf16(64, 32, 0, 1);  // A = P (K-major), B = v (MN-major)
      // A, K-major view of a [query][key] tile: query rows of 128 B, 8-row groups 1 KB apart
      const uint64_t a_p = smem_desc(sb + kOffP, 16, 1024, 2);
      // B, MN-major view of a 64 x 64-byte tile: 32 channels = one 64-byte atom, 8 tokens = 512 B (SBO)
      const uint64_t b_v = smem_desc(sb + kOffStage + 4 * kTile, 16, 512, 4);
      for (int k = 0; k < npairs; ++k) {
        const int s = k % kStages, buf = k & 1;
        mbar_wait_fast(bar_staged(buf), (k >> 1) & 1);
        if (k > 1) mbar_wait_fast(bar_ofree(buf), ((k >> 1) - 1) & 1);  // epilogue of pair k-2 has drained this O buffer
        tc_fence_after();
        if (elect_one()) {
          const uint64_t so = (uint64_t)((s * kStage) >> 4), bo = (uint64_t)(buf * ((2 * kPTile) >> 4));
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const uint32_t dl = (uint32_t)(16 * u) << 16;
            const uint64_t ao = bo + (uint64_t)(u * (kPTile >> 4)), to = so + (uint64_t)(u * (kTile >> 4));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)  // 16 keys per step: A += 32 B inside the swizzle atom, B += 1 KB
              umma_ss(tmem + dl + kColO + 32 * buf, a_p + ao + (uint64_t)(2 * ks), b_v + to + (uint64_t)(64 * ks), id_o, ks > 0);
          }
          umma_commit(bar_o(buf));
          umma_commit(bar_stfree(buf));
        }
        __syncwarp();
      }
    } else {
      // ---------------------------------------------------------------- warp 3: TMA stores of o (written over the q tiles)
      for (int k = 0; k < npairs; ++k) {
        const int s = k % kStages;
        mbar_wait_fast(bar_written(s), (k / kStages) & 1);
        const uint32_t st = sb + kOffStage + s * kStage;
        if (elect_one()) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const UnitGeo ug = geo[(k & 7) * 2 + u];
            const int head = u == 0 ? work.head_a : work.head_b;
            if (ug.rflags & 1)
              for_each_box<kMode>(g, ug.col0, ug.row0, (ug.rflags & 2) != 0, [&](int off, int mi, int col, int row) {
                tma_store_4d(&maps.m[1][mi], st + u * kTile + off, head * 32, col, row, ug.b);
              });
          }
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
    // ------------------------------------------------------------------ norm warps: one (unit, q | k) tile of every pair each
    const int u = warp & 1, part = (warp >> 1) & 1;
    const int head = u == 0 ? work.head_a : work.head_b;
    const float mult = part == 0 ? 1.0f : __ldg(&tau[head]) * kLog2e;
    const int g_ = lane >> 2, t_ = lane & 3;
    const int arow = (lane & 7) + 8 * ((lane >> 3) & 1), achunk = lane >> 4;
    for (int k = 0; k < npairs; ++k) {
      const int s = k % kStages;
      mbar_wait_fast(bar_full(s), (k / kStages) & 1);
      const uint32_t tile = sb + kOffStage + s * kStage + (2 * part + u) * kTile;
      float* vec = vecs + (s * 2 + part) * 128 + 64 * u;  // r (q tile) or c (k tile), tile order
#pragma unroll
      for (int bp = 0; bp < 2; ++bp) {
        uint32_t x[2][2][4];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int row = 16 * (2 * bp + b) + arow;
          ldsm_x4(tile + row * 64 + (((achunk) ^ ((row >> 1) & 3)) << 4), x[b][0]);
          ldsm_x4(tile + row * 64 + (((2 + achunk) ^ ((row >> 1) & 3)) << 4), x[b][1]);
        }
        float n0[2][4], n1[2][4];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
#pragma unroll
          for (int e = 0; e < 4; ++e) n0[b][e] = n1[b][e] = 0.f;
          mma_bf16(n0[b], x[b][0], x[b][0][0], x[b][0][2]);
          mma_bf16(n1[b], x[b][0], x[b][0][1], x[b][0][3]);
        }
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          mma_bf16(n0[b], x[b][1], x[b][1][0], x[b][1][2]);
          mma_bf16(n1[b], x[b][1], x[b][1][1], x[b][1][3]);
        }
        const bool odd = (lane >> 2) & 1;
        const int src = (lane & ~3) | (lane >> 3);
        float s0[2], s1[2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          s0[b] = __shfl_sync(0xffffffffu, odd ? n0[b][1] : n0[b][0], src);
          s1[b] = __shfl_sync(0xffffffffu, odd ? n1[b][3] : n1[b][2], src);
        }
        if (t_ == 0) {
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            vec[16 * (2 * bp + b) + g_] = mult * inv_norm(s0[b]);
            vec[16 * (2 * bp + b) + g_ + 8] = mult * inv_norm(s1[b]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_norm(s));
      // the row scales go out beside the log-sum-exp (planes 1: r, 2: c, slot order) for the backward kernel
      const int rf = geo[(k & 7) * 2 + u].rflags;
      if (rf & 1) {
        float* sp = stats + (1 + part) * p.plane + ((int64_t)(rf >> 3) * g.heads + head) * kN;
        sp[slotmap[lane]] = vec[lane];
        sp[slotmap[lane + 32]] = vec[lane + 32];
      }
    }
  } else if (warp < 24) {
    // ------------------------------------------------------------------ softmax threads: two groups (warps 8-15 even pairs,
    // 16-23 odd pairs); a thread owns half a logit row.  Lanes 0-15 of a warp are rows of unit a, lanes 16-31 of unit b.
    reg_alloc<80>();
    const int grp = (warp - 8) >> 3;
    const int half = ((warp - 8) >> 2) & 1;
    const int quad = warp & 3;
    const int u = lane >> 4, i = 16 * quad + (lane & 15);  // unit of the pair, tile row (query) inside the unit
    const int row = 64 * u + i;
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16);
    // warp-uniform (lanes 0-15 and 16-31 are different heads): if either head of the CTA needs the row maximum both take
    // it -- subtracting a maximum is always valid, `off` only has to match what the bias table was built with
    const bool use_max = (s_usemax[0] | s_usemax[1]) != 0;
    const float off = s_off[u];
    const float kNeg = kMaskValue * kLog2e;
    const int si = slotmap[i], ih = si >> 3, iw = si & 7;
    // Toeplitz bias: float index of (dh = ih + 7, x = 7 - iw) in the alignment copy that makes x a multiple of 4
    const int cpy = (7 - iw) & 3;
    const float* bias_base = reinterpret_cast<const float*>(smem + kOffBias) + u * 4 * kBiasCopy + cpy * kBiasCopy +
                             (ih + 7) * kBiasRow + (7 - iw - cpy) + 4 + (kSplit ? 4 * half : -(4 * half) * kBiasRow);
    // masks of a shifted layer: bit j set = key j of this thread's half sits on the other side of the wrap than the query
    uint32_t mH = 0u, mW = 0u;
    if (kMasked) {
      const int thr = kWs - g.shift;
      for (int j = 0; j < 32; ++j) {
        const int sj = slotmap[32 * half + j];
        if (((sj >> 3) >= thr) != (ih >= thr)) mH |= 1u << j;
        if (((sj & 7) >= thr) != (iw >= thr)) mW |= 1u << j;
      }
    }
    const uint32_t p_row = sb + kOffP + grp * 2 * kPTile + u * kPTile + i * 128;
    const uint32_t tS = tl + kColS + 64 * grp + 32 * half;
    float* hmx = reinterpret_cast<float*>(smem + kOffHmx) + grp * 256;

    for (int k = grp; k < npairs; k += 2) {
      const int s = k % kStages;
      mbar_wait_fast(bar_norm(s), (k / kStages) & 1);  // row scales (and with them the tiles) of the pair exist
      const int rflags = geo[(k & 7) * 2 + u].rflags;
      const float* vec = vecs + s * 2 * 128;
      const float ri = vec[row];
      const float* cv = vec + 128 + 64 * u + 32 * half;  // c_j of this half's keys, tile order
      uint32_t m = 0u;
      if (kMasked) m = ((rflags & 2) ? mH : 0u) | ((rflags & 4) ? mW : 0u);
      if (k > 1) mbar_wait_fast(bar_stfree(grp), ((k >> 1) - 1) & 1);  // the PV MMAs of pair k-2 have read this group's staging tiles
      mbar_wait_fast(bar_s(grp), (k >> 1) & 1);
      tc_fence_after();
      uint32_t acc[32];
      HV_TMEM_LD32(tS, acc);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sfree(grp));
      float sv[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        // keys 4q .. 4q + 3 of this half: slot order = window row 4 half + q / 2, columns 4 (q & 1) ..; split order = window
        // row q, columns 4 half ..
        const float4 b = *reinterpret_cast<const float4*>(bias_base + (kSplit ? -q * kBiasRow : -(q >> 1) * kBiasRow + 4 * (q & 1)));
        const float4 c = *reinterpret_cast<const float4*>(cv + 4 * q);
        sv[4 * q + 0] = fmaf(__uint_as_float(acc[4 * q + 0]) * ri, c.x, b.x);
        sv[4 * q + 1] = fmaf(__uint_as_float(acc[4 * q + 1]) * ri, c.y, b.y);
        sv[4 * q + 2] = fmaf(__uint_as_float(acc[4 * q + 2]) * ri, c.z, b.z);
        sv[4 * q + 3] = fmaf(__uint_as_float(acc[4 * q + 3]) * ri, c.w, b.w);
      }
      if (kMasked && __any_sync(0xffffffffu, m != 0u)) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if ((m >> j) & 1u) sv[j] += kNeg;
      }
      float mx = 0.f;
      if (use_max) {
        mx = sv[0];
#pragma unroll
        for (int j = 1; j < 32; ++j) mx = fmaxf(mx, sv[j]);
        hmx[half * 128 + row] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + grp * 4 + quad) : "memory");
        mx = fmaxf(mx, hmx[(half ^ 1) * 128 + row]);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + grp * 4 + quad) : "memory");  // both halves have read before the next pair writes
#pragma unroll
        for (int j = 0; j < 32; ++j) sv[j] -= mx;
      }
      float ls = 0.f;
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        uint32_t pp[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float p0 = ex2(sv[8 * c8 + 2 * w]), p1 = ex2(sv[8 * c8 + 2 * w + 1]);
          ls += p0 + p1;
          pp[w] = pack_bf16x2(p0, p1);
        }
        sts128(p_row + (uint32_t)(((4 * half + c8) ^ (i & 7)) << 4), make_uint4(pp[0], pp[1], pp[2], pp[3]));
      }
      lsum[((k & 3) * 2 + half) * 128 + row] = ls;
      if (half == 0) mxv[(k & 3) * 128 + row] = off + mx;
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_staged(grp));
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps: O from TMEM, normalise, write the row as
    // bf16 over the q tile (the store warp sends the tiles out by TMA), lse to the statistics
    const int quad = warp & 3;
    const int u = lane >> 4, t = 16 * quad + (lane & 15);  // M = 64 accumulator layout: lanes 0-15 unit a, 16-31 unit b
    const int row = 64 * u + t;
    const int head = u == 0 ? work.head_a : work.head_b;
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16);
    const int sl = slotmap[t];
    const uint32_t swz = (uint32_t)((t >> 1) & 3);
    for (int k = 0; k < npairs; ++k) {
      const int s = k % kStages, ab = k & 1;
      mbar_wait_fast(bar_o(ab), (k >> 1) & 1);
      tc_fence_after();
      uint32_t o[32];
      HV_TMEM_LD32(tl + kColO + 32 * ab, o);
      const float l = lsum[((k & 3) * 2) * 128 + row] + lsum[((k & 3) * 2 + 1) * 128 + row];
      const float lse_off = mxv[(k & 3) * 128 + row];
      const int rf = geo[(k & 7) * 2 + u].rflags;
      tmem_wait_ld();
      HV_REG_FENCE32(o);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ofree(ab));
      const float inv = rcp_fast(l);
      const uint32_t orow = sb + kOffStage + s * kStage + u * kTile + t * 64;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(o[8 * q + 0]) * inv, __uint_as_float(o[8 * q + 1]) * inv);
        v.y = pack_bf16x2(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv);
        v.z = pack_bf16x2(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv);
        v.w = pack_bf16x2(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv);
        sts128(orow + ((q ^ swz) << 4), v);
      }
      if (rf & 1) stats[((int64_t)(rf >> 3) * g.heads + head) * kN + sl] = lse_off + lg2_fast(l);
      fence_async_smem();  // the tiles are read by the TMA engine (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_written(s));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
}

template <bool kShift>
__global__ void __launch_bounds__(kThreads, 1)
wattn_tc64_fwd2_kernel(const __grid_constant__ FwdMaps maps, const float* __restrict__ bias_table, const float* __restrict__ tau,
                       float* __restrict__ stats, const __grid_constant__ FwdParams p) {
  if (!kShift) {
    wattn_tc64_fwd2_body<0>(maps, bias_table, tau, stats, p);
  } else {
    const int cls = cta_window_class(p.sched, blockIdx.x);  // a CTA serves one window class (hv_tc_win.cuh)
    if (cls == 0) wattn_tc64_fwd2_body<0>(maps, bias_table, tau, stats, p);
    else if (cls == 1) wattn_tc64_fwd2_body<1>(maps, bias_table, tau, stats, p);
    else wattn_tc64_fwd2_body<2>(maps, bias_table, tau, stats, p);
  }
}

}  // namespace

bool wattn_tc64_fwd2_supported(const Geom& g, int dtype) {
  // shifted layers: the bias lookup reads runs of four keys, i.e. the column split must sit at 4 (shift = ws / 2, the only
  // shift SwinV2 uses, swinv2.py:560)
  return dtype == HV_BF16 && g.ws == kWs && g.d == 32 && g.C % 32 == 0 && (g.shift == 0 || g.shift == 4) &&
         (int64_t)g.B * g.H * g.W < (int64_t(1) << 31) && g.W * g.C * 2 % 16 == 0;
}

int wattn_tc64_fwd2(const Geom& g, const void* qkv, const float* bias_table, const float* tau, void* out, float* stats,
                    cudaStream_t st) {
  if (!aligned16(qkv) || !aligned16(out) || !aligned16(stats))
    HV_FAIL(HV_ERR_ALIGN, "window_attn: qkv / out / statistics must be 16-byte aligned");
  struct MapKey { const void *qkv, *out; int B, H, W, C, shift; };
  struct MapEntry { MapKey key; FwdMaps maps; };
  static thread_local MapEntry cache[32];
  static thread_local int cache_n = 0, cache_next = 0;
  const MapKey key = {qkv, out, g.B, g.H, g.W, g.C, g.shift};
  const FwdMaps* mp = nullptr;
  for (int i = 0; i < cache_n; ++i) {
    const MapKey& c = cache[i].key;
    if (c.qkv == key.qkv && c.out == key.out && c.B == key.B && c.H == key.H && c.W == key.W && c.C == key.C &&
        c.shift == key.shift) { mp = &cache[i].maps; break; }
  }
  if (!mp) {
    MapEntry& e = cache[cache_next];
    int rc = make_window_maps(e.maps.m[0], qkv, g, 3 * g.C);
    if (rc) return rc;
    rc = make_window_maps(e.maps.m[1], out, g, g.C);
    if (rc) return rc;
    e.key = key;
    mp = &e.maps;
    cache_next = (cache_next + 1) % 32;
    if (cache_n < 32) ++cache_n;
  }
  FwdParams p;
  p.g = g;
  p.plane = (int64_t)g.B * g.nW * g.heads * kN;
  static double cost[3] = {1.0, 1.2, 1.3};
  static const bool cost_env = []() {
    const char* e = getenv("HV_FWD_CLASS_COST");
    if (e) sscanf(e, "%lf,%lf", &cost[1], &cost[2]);
    return e != nullptr;
  }();
  (void)cost_env;
  const int grid = plan_window_schedule(g, num_sms(), cost, p.sched);
  static thread_local int attr_dev = -1;
  int dev = 0;
  HV_CUDA_OK(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    HV_CUDA_OK(cudaFuncSetAttribute(wattn_tc64_fwd2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    HV_CUDA_OK(cudaFuncSetAttribute(wattn_tc64_fwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_dev = dev;
  }
  if (g.shift > 0)
    wattn_tc64_fwd2_kernel<true><<<grid, kThreads, kSmem, st>>>(*mp, bias_table, tau, stats, p);
  else
    wattn_tc64_fwd2_kernel<false><<<grid, kThreads, kSmem, st>>>(*mp, bias_table, tau, stats, p);
  HV_LAUNCH_OK("wattn_tc64_fwd2_kernel");
  return HV_OK;
}

}  // namespace hv
