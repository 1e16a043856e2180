// tcgen05 / TMEM / TMA forward kernel of the fused shifted-window scaled-cosine attention, window 8x8 (N = 64), head
// dim 32, bf16 (every stage of SwinV2-T).  Same contract as wattn_mma64_fwd (qkv (B, H*W, 3C) and out (B, H*W, C) in
// IMAGE token order, lse in log2 units), different machine mapping:
//
//   * the cyclic shift + window partition (swinv2.py:399-412) is the coordinate of a 4-D TMA tile load: one
//     `cp.async.bulk.tensor.4d` with box (32 channels, 8, 8, 1) and SWIZZLE_64B brings the q, k or v tile of one
//     (window, head) -- 64 rows x 64 B -- straight into the K-major layout tcgen05.mma reads.  Windows that wrap
//     around the image edge are split into 2 boxes (rows) or 16 row boxes (columns), never copied twice;
//   * two (window, head) units are stacked into one M = 128 tile: S = [Q_a; Q_b][K_a; K_b]^T is one
//     tcgen05.mma (N = 128, K = 32) whose diagonal 64 x 64 blocks are the two logit matrices, accumulated in TMEM;
//   * 128 softmax threads own one query row each (TMEM lane = row): tcgen05.ld, scale by 1/|q_i| * tau/|k_j|, add the
//     continuous position bias (expanded once per CTA into shared memory), shift mask, exp2, and write P back to TMEM
//     as bf16 (tcgen05.st) -- no shuffles, no fragments; O = P V and the row sums l = P 1 are tcgen05.mma with A
//     from TMEM and V (MN-major, the same TMA tile) / a ones tile from shared memory;
//   * warp roles: 0 TMA producer, 1 MMA issuer (one thread), 2-3 L2 norms of the q / k rows (tensor-pipe self
//     products via mma.sync on the swizzled tiles), 4-11 two softmax groups working on alternate unit pairs so that
//     the MMAs, the loads and the exponentials of neighbouring pairs overlap; everything hands over through mbarriers.
// waits with a suspend-time hint: measured 2-3 % faster for this kernel (fewer polling instructions), neutral to slightly
// negative for the backward, which keeps the plain try_wait loop
#define HV_WAIT_HINT_NS 1000
#include "hv_tc.cuh"
#include <atomic>

namespace hv {
namespace {
using namespace tc;

constexpr int kN = 64;
constexpr int kWs = 8;
constexpr int kTab = 225;
constexpr int kTile = kN * 64;        // one (window, head) q / k / v tile: 64 rows x 64 B
constexpr int kStage = 6 * kTile;     // q_a q_b k_a k_b v_a v_b
#ifndef HV_TC_STAGES
#define HV_TC_STAGES 5
#endif
constexpr int kStages = HV_TC_STAGES;
constexpr int kBiasPitch = 68;        // floats per expanded-bias row (64 + pad: conflict-free 16-byte row reads)
constexpr int kThreads = 896;  // 28 warps: 0 TMA, 1 MMA, (2-3 idle), 4-7 norms, 8-23 softmax, 24-27 epilogue
constexpr float kNoMaxRange = 64.0f;

// ---- shared memory map (dynamic, 1024-byte aligned base)
constexpr int kOffStage = 0;
constexpr int kOffOnes = kOffStage + kStages * kStage;                 // 64 x 64 B of bf16 ones
constexpr int kOffBias = kOffOnes + kTile;                             // [2 orders][2 units][64][kBiasPitch] float
#ifdef HV_TC_ONE_ORDER  // experiment: shift 0 only, no permuted bias copy
constexpr int kOffVec = kOffBias + 2 * kN * kBiasPitch * 4;
#else
constexpr int kOffVec = kOffBias + 4 * kN * kBiasPitch * 4;
#endif            // [kStages][2 units][2 (r, c)][64] float
constexpr int kOffMx = kOffVec + kStages * 2 * 2 * kN * 4;             // [2 slots][2 phases][128] float
constexpr int kOffHmx = kOffMx + 4 * 128 * 4;                          // [2 slots][2 phases][2 halves][128] float
constexpr int kOffTab = kOffHmx + 8 * 128 * 4;                         // [2 units][256] float
constexpr int kOffGeo = kOffTab + 2 * 256 * 4;                         // [8][2] UnitGeo
constexpr int kOffSlotMap = kOffGeo + 8 * 2 * 16;                      // [2 orders][64] bytes
constexpr int kOffBar = kOffSlotMap + 128;
constexpr int kNumBars = 3 * kStages + 5 * 2;
constexpr int kOffTmem = kOffBar + kNumBars * 8;
constexpr int kSmem = kOffTmem + 16;

#ifdef HV_TC_TRACE
__device__ long long* g_trace = nullptr;  // [pairs][16 events] clock64 stamps of CTA 0
#define TRACE(k, ev) do { if (blockIdx.x == 0 && lane == 0 && g_trace && (k) < 64) g_trace[(k) * 16 + (ev)] = clock64(); } while (0)
#define KO(bit) (p.ko & (bit))
#else
#define TRACE(k, ev) do { } while (0)
#define KO(bit) false
#endif

struct TcParams {
  Geom g;
  int n_same;        // head groups whose two units are heads (2g, 2g+1) of the same window
  int has_cross;     // 1: a last group pairs the odd head of two consecutive windows
  int ctas_same;     // CTAs per same-window group
  int ctas_cross;    // CTAs of the cross-window group
  int ko;            // HV_TC_TRACE builds only: knock-out bits for bottleneck experiments (results are wrong)
  int64_t plane;     // floats per statistics plane (lse | r | c): B * nW * heads * 64
};
// box shapes (w, h): 0 full (8,8) | 1 (8,8-s) 2 (8,s) row wrap | 3 (8-s,8) 4 (s,8) column wrap | 5..8 corner
struct TcMaps { CUtensorMap m[9]; };

// The work of one CTA: group, position inside the group, and the unit pairs it walks through
struct CtaWork {
  int head_a, head_b, cross, first, stride, npairs;
  __device__ __forceinline__ void init(const TcParams& p, int cta) {
    const int nrows = p.g.B * p.g.nW;
    const int same_total = p.n_same * p.ctas_same;
    if (cta < same_total) {
      const int grp = cta / p.ctas_same;
      cross = 0; head_a = 2 * grp; head_b = 2 * grp + 1;
      first = cta - grp * p.ctas_same; stride = p.ctas_same;
      npairs = first < nrows ? (nrows - first + stride - 1) / stride : 0;
    } else {
      cross = 1; head_a = head_b = p.g.heads - 1;
      first = cta - same_total; stride = p.ctas_cross;
      const int nrp = (nrows + 1) / 2;
      npairs = first < nrp ? (nrp - first + stride - 1) / stride : 0;
    }
  }
  // window row of unit `which` of pair k; valid = false for the padding unit of an odd tail
  __device__ __forceinline__ int row(int k, int which, int nrows, bool& valid) const {
    const int idx = first + k * stride;
    valid = true;
    if (!cross) return idx;
    const int r = 2 * idx + which;
    if (r >= nrows) { valid = false; return nrows - 1; }
    return r;
  }
};

struct WinPos {
  int b, row0, col0;
  bool bottom, right;  // window touches the wrapped band along h / w
};
__device__ __forceinline__ WinPos win_pos(const Geom& g, int r) {
  WinPos w;
  w.b = r / g.nW;
  const int win = r - w.b * g.nW;
  const int wh = win / g.nWw, ww = win - wh * g.nWw;
  w.row0 = wh * kWs + g.shift;
  w.col0 = ww * kWs + g.shift;
  w.bottom = g.shift > 0 && wh == g.H / kWs - 1;
  w.right = g.shift > 0 && ww == g.nWw - 1;
  return w;
}
// Window slot (ih, iw) held by tile row t.  Interior / row-wrapped windows are loaded in slot order; a column-wrapped
// window arrives as two dense boxes (columns [0, 8-s) then [8-s, 8)), i.e. in a permuted row order.  Attention does
// not care about the order of the tokens as long as bias, mask and the output address follow it.
__device__ __forceinline__ void tile_row_slot(int t, int shift, bool perm, int& ih, int& iw) {
  if (!perm) { ih = t >> 3; iw = t & 7; return; }
  const int wa = kWs - shift;
  if (t < kWs * wa) { ih = t / wa; iw = t - ih * wa; }
  else { const int t2 = t - kWs * wa; ih = t2 / shift; iw = wa + t2 - ih * shift; }
}

// TMEM columns of one pair slot (two slots = all 512 columns)
constexpr int kColS = 0, kColP = 128, kColO = 160, kColL = 224, kSlotCols = 256;

// geometry of one unit of a pair, written by the producer, read by the softmax and epilogue threads
struct UnitGeo { int b, row0, col0, rflags; };  // rflags = window row << 3 | right << 2 | bottom << 1 | valid

__global__ void __launch_bounds__(kThreads, 1)
wattn_tc64_fwd_kernel(const __grid_constant__ TcMaps maps, const float* __restrict__ bias_table,
                      const float* __restrict__ tau, bf16* __restrict__ out, float* __restrict__ lse, TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const Geom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar0 = sb + kOffBar;
  auto bar_full = [&](int s) { return bar0 + 8 * s; };
  auto bar_empty = [&](int s) { return bar0 + 8 * (kStages + s); };
  auto bar_norm = [&](int s) { return bar0 + 8 * (2 * kStages + s); };
  auto bar_s = [&](int t) { return bar0 + 8 * (3 * kStages + t); };
  auto bar_p = [&](int t) { return bar0 + 8 * (3 * kStages + 2 + t); };
  auto bar_o = [&](int t) { return bar0 + 8 * (3 * kStages + 4 + t); };
  auto bar_ofree = [&](int t) { return bar0 + 8 * (3 * kStages + 6 + t); };
  auto bar_sfree = [&](int t) { return bar0 + 8 * (3 * kStages + 8 + t); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmem);
  const int nrows = g.B * g.nW;

  // the CTA's work description lives in shared memory and every role copies it into its own registers after the role
  // split (a value computed before the split would be live across all roles and spill in the small-register ones)
  __shared__ CtaWork s_work;
  if (threadIdx.x == 0) {
    CtaWork w0;
    w0.init(p, blockIdx.x);
    s_work = w0;
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 5);  // four norm warps + the MMA warp's commit
      mbar_init(bar_norm(s), 4);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_s(t), 1);
      mbar_init(bar_p(t), 256);
      mbar_init(bar_o(t), 1);
      mbar_init(bar_ofree(t), 128);
      mbar_init(bar_sfree(t), 8);  // one arrival per softmax warp of the group: its S columns are in registers
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // ---- one-time tables: ones tile (B operand of the row-sum MMA), window slot of every tile row in both load orders,
  //      the bias table of the CTA's two heads in log2 units, and from those the expanded 64 x 64 bias matrices.
  // Logits are tau2*cos + bias2 with |cos| <= 1: if 2*tau2 + (bias range) stays far inside the fp32 exponent range the
  // running maximum is skipped and exp2(logit - (tau2 + max bias)) is used directly (see wattn_mma64.cu).
  for (int i = threadIdx.x; i < kTile / 4; i += kThreads) reinterpret_cast<uint32_t*>(smem + kOffOnes)[i] = 0x3F803F80u;
  unsigned char* slotmap = smem + kOffSlotMap;  // [2 orders][64]: ih << 3 | iw
  float* tab_s = reinterpret_cast<float*>(smem + kOffTab);  // [2 units][256]
  __shared__ float s_off[2];
  __shared__ int s_usemax[2];
  if (threadIdx.x < 128) {
    int ih, iw;
    tile_row_slot(threadIdx.x & 63, g.shift > 0 ? g.shift : 4, threadIdx.x >= 64 && g.shift > 0, ih, iw);
    slotmap[threadIdx.x] = (unsigned char)(ih << 3 | iw);
  }
  if (warp < 2) {
    CtaWork w0;
    w0.init(p, blockIdx.x);
    const int head = warp == 0 ? w0.head_a : w0.head_b;
    float bv[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) bv[q] = kLog2e * __ldg(&bias_table[min(lane + 32 * q, kTab - 1) * g.heads + head]);
    float bmx = bv[0], bmn = bv[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) { bmx = fmaxf(bmx, bv[q]); bmn = fminf(bmn, bv[q]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      bmx = fmaxf(bmx, __shfl_xor_sync(0xffffffffu, bmx, o));
      bmn = fminf(bmn, __shfl_xor_sync(0xffffffffu, bmn, o));
    }
    const float tau2 = __ldg(&tau[head]) * kLog2e;
    const bool use_max = !(2.0f * tau2 + (bmx - bmn) <= kNoMaxRange);
    const float off = use_max ? 0.f : tau2 + bmx;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (lane + 32 * q < kTab) tab_s[warp * 256 + lane + 32 * q] = bv[q] - off;
    if (lane == 0) {
      s_usemax[warp] = use_max ? 1 : 0;
      s_off[warp] = off;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // ones tile is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  {
    // expanded bias: [order (slot / permuted)][unit][row][kBiasPitch]; the permuted copy only exists when shift > 0
    float* bias_s = reinterpret_cast<float*>(smem + kOffBias);
    const int norders = g.shift > 0 ? 2 : 1;
    for (int idx = threadIdx.x; idx < norders * 2 * kN * kN; idx += kThreads) {
      const int ord = idx >> 13, u = (idx >> 12) & 1, ti = (idx >> 6) & 63, tj = idx & 63;
      const int si = slotmap[ord * 64 + ti], sj = slotmap[ord * 64 + tj];
      const int rel = ((si >> 3) - (sj >> 3) + kWs - 1) * (2 * kWs - 1) + ((si & 7) - (sj & 7) + kWs - 1);
      bias_s[((ord * 2 + u) * kN + ti) * kBiasPitch + tj] = tab_s[u * 256 + rel];
    }
  }
  __syncthreads();
  const uint32_t tmem = *tmem_slot;
  const CtaWork work = s_work;
  const int npairs = work.npairs;
  float* mxv = reinterpret_cast<float*>(smem + kOffMx);    // [2 slots][2 phases][128]: off + row max, softmax -> epilogue
  float* hmx = reinterpret_cast<float*>(smem + kOffHmx);   // [2 slots][2 phases][2 halves][128]: half-row maxima
  UnitGeo* geo = reinterpret_cast<UnitGeo*>(smem + kOffGeo);  // [8][2]

  if (warp < 4) {
    reg_dealloc<40>();
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer: lane t loads tile t of the stage
      for (int k = 0; k < npairs; ++k) {
        const int s = k % kStages;
        const uint32_t ph = (k / kStages) & 1;
        mbar_wait(bar_empty(s), ph ^ 1);
        if (warp == 0) TRACE(k, 0);  // producer: stage free
        const int which = lane & 1, part = lane >> 1;
        bool valid;
        const int r = work.row(k, which, nrows, valid);
        const WinPos w = win_pos(g, r);
        if (lane < 2) {
          UnitGeo ug;
          ug.b = w.b; ug.row0 = w.row0; ug.col0 = w.col0;
          ug.rflags = (r << 3) | (w.right ? 4 : 0) | (w.bottom ? 2 : 0) | (valid ? 1 : 0);
          geo[(k & 7) * 2 + which] = ug;
        }
        __syncwarp();
        if (lane == 0) mbar_expect_tx(bar_full(s), KO(8) ? 0 : kStage);
        __syncwarp();
        if (lane < 6 && !KO(8)) {
          const int head = which == 0 ? work.head_a : work.head_b;
          const int c0 = part * g.C + head * 32;
          const uint32_t dst = sb + kOffStage + s * kStage + lane * kTile;
          const uint32_t bar = bar_full(s);
          const bool hwrap = w.row0 + kWs > g.H, wwrap = w.col0 + kWs > g.W;
          const int sh = g.shift, wa = kWs - g.shift;
          if (!wwrap && !hwrap) {
            tma_load_4d(dst, &maps.m[0], bar, c0, w.col0, w.row0, w.b);
          } else if (!wwrap) {
            tma_load_4d(dst, &maps.m[1], bar, c0, w.col0, w.row0, w.b);
            tma_load_4d(dst + wa * kWs * 64, &maps.m[2], bar, c0, w.col0, 0, w.b);
          } else if (!hwrap) {
            tma_load_4d(dst, &maps.m[3], bar, c0, w.col0, w.row0, w.b);
            tma_load_4d(dst + kWs * wa * 64, &maps.m[4], bar, c0, 0, w.row0, w.b);
          } else {
            tma_load_4d(dst, &maps.m[5], bar, c0, w.col0, w.row0, w.b);                                  // (wa, wa)
            tma_load_4d(dst + wa * wa * 64, &maps.m[6], bar, c0, w.col0, 0, w.b);                        // (wa, s)
            tma_load_4d(dst + kWs * wa * 64, &maps.m[7], bar, c0, 0, w.row0, w.b);                       // (s, wa)
            tma_load_4d(dst + kWs * wa * 64 + sh * wa * 64, &maps.m[8], bar, c0, 0, 0, w.b);             // (s, s)
          }
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- S issuer: S(k) = [Q_a; Q_b] [K_a; K_b]^T
      // S of pair k goes in as soon as the softmax group has pulled S of pair k-2 out of TMEM (P has its own columns),
      // so the group finds its next logits waiting when it is done with the exponentials of the previous pair.
      // A warp of its own: every wait / commit costs the issuing warp ~100 cycles, PV and S issue would add up.
      const uint32_t id_s = idesc_bf16(128, 128, 0, 0);
      const uint64_t d_q = sw64_desc(sb + kOffStage), d_k = sw64_desc(sb + kOffStage + 2 * kTile);
      for (int k = 0; k < npairs; ++k) {
        const int s = k % kStages, t = k & 1;
        if (k >= 2) mbar_wait_fast(bar_sfree(t), ((k - 2) >> 1) & 1);
        mbar_wait_fast(bar_full(s), (k / kStages) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t so = (uint64_t)((s * kStage) >> 4);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
            umma_ss(tmem + kSlotCols * t + kColS, d_q + so + 2 * kk, d_k + so + 2 * kk, id_s, kk > 0);
          umma_commit(bar_s(t));
          TRACE(k, 3);  // S issued
        }
        __syncwarp();
      }
    } else if (warp == 2) {
      // ---------------------------------------------------------------- PV issuer: O = P V, l = P 1 (A from TMEM)
      const uint32_t id_o = idesc_bf16(128, 32, 0, 1);
      const uint64_t ones_desc = sw64_desc(sb + kOffOnes);
      const uint64_t d_v = sw64_desc(sb + kOffStage + 4 * kTile);
      for (int k = 0; k < npairs; ++k) {
        const int s = k % kStages, t = k & 1;
        mbar_wait_fast(bar_p(t), (k >> 1) & 1);
        TRACE(k, 7);  // MMA warp saw P
        mbar_wait_fast(bar_ofree(t), ((k >> 1) & 1) ^ 1);  // epilogue of pair k-2 has drained O / l of this slot
        tc_fence_after();
        if (elect_one()) {
          const uint64_t so = (uint64_t)((s * kStage) >> 4);
          const uint32_t tb = tmem + kSlotCols * t;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_ts(tb + kColO + 32 * half, tb + kColP + 8 * ks, d_v + so + (uint64_t)(256 * half + 64 * ks), id_o, ks > 0);
          }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ts(tb + kColL, tb + kColP + 8 * ks, ones_desc + (uint64_t)(64 * ks), id_o, ks > 0);
          umma_commit(bar_o(t));
          TRACE(k, 8);  // PV issued
          umma_commit(bar_empty(s));
        }
        __syncwarp();
      }
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
      if (warp == 4) TRACE(k, 1);  // norm warp saw full
      const uint32_t tile = sb + kOffStage + s * kStage + (2 * part + u) * kTile;
      float* vec = reinterpret_cast<float*>(smem + kOffVec) + ((s * 2 + u) * 2 + part) * kN;  // r (q tile) or c (k tile)
      // two 16-row blocks at a time with every step batched (loads, products, shuffles, rsqrt) so that the four
      // dependent chains overlap instead of running back to back
#pragma unroll
      for (int bp = 0; bp < 2; ++bp) {
        if (KO(1)) { vec[lane + 32 * bp] = mult; continue; }
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
      if (lane == 0) {
        mbar_arrive(bar_norm(s));
        if (warp == 4) TRACE(k, 2);  // norm done
        mbar_arrive(bar_empty(s));
      }
      // the row scales go out beside the log-sum-exp (planes 1: r_i = 1 / |q_i|, 2: c_j = tau log2e / |k_j|, slot order):
      // the backward kernel reads them back instead of recomputing the norms from the tiles
      const int rf = geo[(k & 7) * 2 + u].rflags;
      if (rf & 1) {
        float* sp = lse + (1 + part) * p.plane + ((int64_t)(rf >> 3) * g.heads + head) * kN;
        const unsigned char* sm = slotmap + ((rf & 4) ? 64 : 0);
        sp[sm[lane]] = vec[lane];
        sp[sm[lane + 32]] = vec[lane + 32];
      }
    }
  } else if (warp < 24) {
    // ------------------------------------------------------------------ softmax: 2 groups (alternate pairs) x 2 column
    // halves x 4 lane quadrants; a thread owns half a logit row (32 keys) of one query
    reg_alloc<80>();
    const int grp = (warp - 8) >> 3;
    const int half = ((warp - 8) >> 2) & 1;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;      // accumulator row = TMEM lane
    const int u = row >> 6, i = row & 63;  // unit of the pair, tile row inside the unit
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16) + kSlotCols * grp;
    const bool use_max = s_usemax[u] != 0;
    const float off = s_off[u];
    const float kNeg = kMaskValue * kLog2e;
    const int thr = kWs - g.shift;
    // 32-bit patterns of the wrapped band over this thread's 32 keys, xor-ed with the own row's membership:
    // slot order / permuted order
    uint32_t mHn = 0u, mWn = 0u, mHp = 0u, mWp = 0u;
    if (g.shift > 0) {
      for (int j = 0; j < 32; ++j) {
        const int sn = slotmap[32 * half + j], sp = slotmap[64 + 32 * half + j];
        if ((sn >> 3) >= thr) mHn |= 1u << j;
        if ((sn & 7) >= thr) mWn |= 1u << j;
        if ((sp >> 3) >= thr) mHp |= 1u << j;
        if ((sp & 7) >= thr) mWp |= 1u << j;
      }
      const int sn = slotmap[i], sp = slotmap[64 + i];
      if ((sn >> 3) >= thr) mHn = ~mHn;
      if ((sn & 7) >= thr) mWn = ~mWn;
      if ((sp >> 3) >= thr) mHp = ~mHp;
      if ((sp & 7) >= thr) mWp = ~mWp;
    }
    const float* bias_n = reinterpret_cast<const float*>(smem + kOffBias) + (u * kN + i) * kBiasPitch + 32 * half;
    const float* bias_p = bias_n + 2 * kN * kBiasPitch;

    for (int k = grp; k < npairs; k += 2) {
      const int s = k % kStages;
      const uint32_t tph = (k >> 1) & 1;
      mbar_wait_fast(bar_norm(s), (k / kStages) & 1);
      if (quad == 0 && half == 0) TRACE(k, 4);  // softmax saw norm
      const int rflags = geo[(k & 7) * 2 + u].rflags;
      const float* bias_row = (rflags & 4) ? bias_p : bias_n;
      const float* vec = reinterpret_cast<const float*>(smem + kOffVec) + (s * 2 + u) * 2 * kN;
      const float ri = vec[i];
      mbar_wait_fast(bar_s(grp), tph);
      if (quad == 0 && half == 0) TRACE(k, 5);  // softmax saw S
      tc_fence_after();
      uint32_t acc[32];
      HV_TMEM_LD32(tl + kColS + 64 * u + 32 * half, acc);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sfree(grp));
      float sv[32];
      if (KO(2)) {
#pragma unroll
        for (int j = 0; j < 32; ++j) sv[j] = __uint_as_float(acc[j]);
      } else
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 c = *reinterpret_cast<const float4*>(&vec[kN + 32 * half + 4 * q]);
        const float4 b = *reinterpret_cast<const float4*>(&bias_row[4 * q]);
        sv[4 * q + 0] = fmaf(__uint_as_float(acc[4 * q + 0]) * ri, c.x, b.x);
        sv[4 * q + 1] = fmaf(__uint_as_float(acc[4 * q + 1]) * ri, c.y, b.y);
        sv[4 * q + 2] = fmaf(__uint_as_float(acc[4 * q + 2]) * ri, c.z, b.z);
        sv[4 * q + 3] = fmaf(__uint_as_float(acc[4 * q + 3]) * ri, c.w, b.w);
      }
      if (rflags & 6) {  // warp-uniform: a warp's 32 rows belong to one unit
        __syncwarp();    // convergence point: keeps the predicated adds out of the interior-window path
        const bool right = rflags & 4, bottom = rflags & 2;
        const uint32_t m = (bottom ? (right ? mHp : mHn) : 0u) | (right ? mWp : 0u);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if ((m >> j) & 1u) sv[j] += kNeg;
      }
      float mx = 0.f;
      if (use_max) {
        mx = sv[0];
#pragma unroll
        for (int j = 1; j < 32; ++j) mx = fmaxf(mx, sv[j]);
        float* hm = hmx + ((grp * 2 + tph) * 2) * 128;
        hm[half * 128 + row] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + grp * 4 + quad) : "memory");
        mx = fmaxf(mx, hm[(half ^ 1) * 128 + row]);
#pragma unroll
        for (int j = 0; j < 32; ++j) sv[j] -= mx;
      }
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = KO(2) ? pack_bf16x2(sv[2 * j], sv[2 * j + 1]) : pack_bf16x2(ex2(sv[2 * j]), ex2(sv[2 * j + 1]));
      HV_TMEM_ST16(tl + kColP + 16 * half, pk);
      if (half == 0) mxv[(grp * 2 + tph) * 128 + row] = off + mx;
      tmem_wait_st();
      tc_fence_before();
      if (quad == 0 && half == 0) TRACE(k, 6);  // softmax P arrive
      mbar_arrive(bar_p(grp));
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps: O and l from TMEM, normalise, store
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int u = row >> 6, i = row & 63;
    const int head = u == 0 ? work.head_a : work.head_b;
    const int sn = slotmap[i], sp = slotmap[64 + i];
    for (int k = 0; k < npairs; ++k) {
      const int t = k & 1;
      const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16) + kSlotCols * t;
      mbar_wait_fast(bar_o(t), (k >> 1) & 1);
      if (quad == 0) TRACE(k, 9);  // epilogue saw O
      tc_fence_after();
      uint32_t o[32];
      HV_TMEM_LD32(tl + kColO + 32 * u, o);
      const float l = __uint_as_float(tmem_ld1(tl + kColL));
      tmem_wait_ld();
      const float lse_off = mxv[(t * 2 + ((k >> 1) & 1)) * 128 + row];
      const UnitGeo ug = geo[(k & 7) * 2 + u];
      tc_fence_before();
      mbar_arrive(bar_ofree(t));
      if (quad == 0) TRACE(k, 10);  // epilogue freed slot
      if ((ug.rflags & 1) && !KO(4)) {
        const float inv = rcp_fast(l);
        const int sl = (ug.rflags & 4) ? sp : sn;
        const int ih = sl >> 3, iw = sl & 7;
        int prow = ug.row0 + ih; if (prow >= g.H) prow -= g.H;
        int pcol = ug.col0 + iw; if (pcol >= g.W) pcol -= g.W;
        const int64_t tok = ((int64_t)ug.b * g.H + prow) * g.W + pcol;
        uint4* dst = reinterpret_cast<uint4*>(out + tok * g.C + head * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[8 * q + 0]) * inv, __uint_as_float(o[8 * q + 1]) * inv);
          v.y = pack_bf16x2(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv);
          v.z = pack_bf16x2(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv);
          v.w = pack_bf16x2(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv);
          dst[q] = v;
        }
        lse[((int64_t)(ug.rflags >> 3) * g.heads + head) * kN + sl] = lse_off + lg2_fast(l);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

}  // namespace

static std::atomic<int> g_fwd_variant{-1};  // -1: HV_ATTN_TCGEN05 environment variable (default automatic), 0: mma.sync, 1: tcgen05

int wattn_fwd_variant_set(int v) {
  const int old = g_fwd_variant.exchange(v, std::memory_order_relaxed);
  return old;
}
int wattn_fwd_variant_get() { return g_fwd_variant.load(std::memory_order_relaxed); }

bool wattn_tc64_supported(const Geom& g, int dtype) {
  // HV_ATTN_TCGEN05: unset or 1 = wherever valid (automatic), 0 = never (mma.sync forward)
  static const int env = []() { const char* e = getenv("HV_ATTN_TCGEN05"); return e == nullptr ? -1 : (atoi(e) != 0 ? 1 : 0); }();
  const int cur = g_fwd_variant.load(std::memory_order_relaxed);
  const int mode = cur < 0 ? env : cur;
  if (mode == 0) return false;  // 1: second-generation kernel where valid, else this one; 2: always this one
  // an odd shift would put the second half of a column-wrapped row at a 64-byte (not 128-byte) shared-memory offset
  const bool valid = dtype == HV_BF16 && g.ws == kWs && g.d == 32 && g.C % 32 == 0 && (g.shift & 1) == 0 &&
                     (int64_t)g.B * g.H * g.W < (int64_t(1) << 31) && g.W * 3 * g.C * 2 % 16 == 0;
  // automatic = wherever valid.  B200 measurements (tools/bench_kernels.py, batch 256, profiles/r02_summary.md): 0.69 vs
  // 0.61 of the HBM roofline at the stage-0 shape; at the later stages the mma.sync + cp.async forward is 2-4 % faster
  // (192-byte rows per token against this kernel's 64-byte TMA rows), which is not worth a second forward path
  (void)mode;
  return valid;
}

int wattn_tc64_fwd(const Geom& g, const void* qkv, const float* bias_table, const float* tau, void* out, float* lse,
                   cudaStream_t st) {
  if (!aligned16(qkv) || !aligned16(out)) HV_FAIL(HV_ERR_ALIGN, "window_attn: qkv/out must be 16-byte aligned");
  // the nine descriptors depend on the qkv pointer and the geometry only: a training step replays the same handful of
  // (pointer, stage) combinations, so encode once and keep the last few (cuTensorMapEncodeTiled costs ~1 us each)
  struct MapKey { const void* ptr; int B, H, W, C, shift; };
  struct MapEntry { MapKey key; TcMaps maps; };
  static thread_local MapEntry cache[32];
  static thread_local int cache_n = 0, cache_next = 0;
  const MapKey key = {qkv, g.B, g.H, g.W, g.C, g.shift};
  const TcMaps* mp = nullptr;
  for (int i = 0; i < cache_n; ++i) {
    const MapKey& c = cache[i].key;
    if (c.ptr == key.ptr && c.B == key.B && c.H == key.H && c.W == key.W && c.C == key.C && c.shift == key.shift) { mp = &cache[i].maps; break; }
  }
  if (!mp) {
    MapEntry& e = cache[cache_next];
    const int s = g.shift, wa = kWs - g.shift;
    // (w, h) of every box; with shift 0 only the first is used (the others just need to be valid descriptors)
    const int bw[9] = {kWs, kWs, kWs, s ? wa : kWs, s ? s : kWs, s ? wa : kWs, s ? wa : kWs, s ? s : kWs, s ? s : kWs};
    const int bh[9] = {kWs, s ? wa : kWs, s ? s : kWs, kWs, kWs, s ? wa : kWs, s ? s : kWs, s ? wa : kWs, s ? s : kWs};
    for (int i = 0; i < 9; ++i) {
      const int rc = make_map(&e.maps.m[i], qkv, g, 3 * g.C, bw[i], bh[i]);
      if (rc) return rc;
    }
    e.key = key;
    mp = &e.maps;
    cache_next = (cache_next + 1) % 32;
    if (cache_n < 32) ++cache_n;
  }
  const TcMaps& maps = *mp;
  TcParams p;
  p.g = g;
  p.n_same = g.heads / 2;
  p.has_cross = g.heads & 1;
  p.ko = 0;
  p.plane = (int64_t)g.B * g.nW * g.heads * kN;
#ifdef HV_TC_TRACE
  if (getenv("HV_TC_KO")) p.ko = atoi(getenv("HV_TC_KO"));
#endif
  const int nsm = num_sms();
  const int nrows = g.B * g.nW;
  if (p.n_same == 0) {
    p.ctas_same = 0;
    p.ctas_cross = nsm;
  } else if (!p.has_cross) {
    p.ctas_same = nsm / p.n_same;
    p.ctas_cross = 0;
  } else {
    p.ctas_cross = nsm / (2 * p.n_same + 1);
    if (p.ctas_cross < 1) p.ctas_cross = 1;
    p.ctas_same = (nsm - p.ctas_cross) / p.n_same;
  }
  if (p.ctas_same < 1 && p.n_same) p.ctas_same = 1;
  if (p.ctas_same > nrows) p.ctas_same = nrows;
  if (p.ctas_cross > (nrows + 1) / 2) p.ctas_cross = (nrows + 1) / 2;
  const int grid = p.n_same * p.ctas_same + p.ctas_cross;
  static thread_local int attr_dev = -1;
  int dev = 0;
  HV_CUDA_OK(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    HV_CUDA_OK(cudaFuncSetAttribute(wattn_tc64_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_dev = dev;
  }
#ifdef HV_TC_TRACE
  static long long* dtrace = nullptr;
  if (!dtrace) {
    cudaMalloc(&dtrace, 64 * 16 * sizeof(long long));
    cudaMemcpyToSymbol(g_trace, &dtrace, sizeof(dtrace));
  }
  cudaMemsetAsync(dtrace, 0, 64 * 16 * sizeof(long long), st);
#endif
  wattn_tc64_fwd_kernel<<<grid, kThreads, kSmem, st>>>(maps, bias_table, tau, (bf16*)out, lse, p);
  HV_LAUNCH_OK("wattn_tc64_fwd_kernel");
#ifdef HV_TC_TRACE
  if (getenv("HV_TC_TRACE_DUMP")) {
    static long long h[64 * 16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, dtrace, sizeof(h), cudaMemcpyDeviceToHost);
    FILE* f = fopen(getenv("HV_TC_TRACE_DUMP"), "w");
    if (f) {
      for (int k = 0; k < 64; ++k) { for (int e = 0; e < 11; ++e) fprintf(f, "%lld ", h[k * 16 + e] ? h[k * 16 + e] - h[0] : -1LL); fprintf(f, "\n"); }
      fclose(f);
    }
  }
#endif
  return HV_OK;
}

}  // namespace hv
