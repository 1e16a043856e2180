// tcgen05 / TMEM / TMA backward kernel of the fused shifted-window scaled-cosine attention, window 8x8 (N = 64), head
// dim 32, bf16 (every stage of SwinV2-T).  Reference swinv2.py:210-263 differentiated; qkv / dqkv (B, H*W, 3C) and dout
// (B, H*W, C) in IMAGE token order; per-CTA partials of d(bias table), d(tau) and the dq column sums are folded by a
// second small kernel.  Machine mapping:
//
//   * the cyclic shift + window partition is the coordinate of 4-D TMA tile LOADS (q, k, v, dO tiles of a (window, head)
//     unit: 64 rows x 64 B, SWIZZLE_64B) and of 4-D TMA tile STORES: the epilogue writes dq / dk / dv over the q / k /
//     dO tiles of the stage (same swizzled layout) and one warp hands them to `cp.async.bulk.tensor` with the same box
//     geometry -- window_reverse + un-roll never exist as copies or as per-thread 64-byte global stores;
//   * nothing is recomputed that the forward already had: the forward kernels save, beside the row log-sum-exp, the
//     row scales r_i = 1 / |q_i| and c_j = tau log2e / |k_j| (three fp32 planes, 768 B per unit, fetched by 1-D bulk
//     copies onto the stage's mbarrier), and D_i = sum_j P_ij dP_ij is formed by the softmax threads themselves (the
//     reference's own softmax backward) -- no pre-pass over the tiles and the attention output `o` is never read;
//   * all products are tcgen05.mma with fp32 accumulators in tensor memory.  S = Q K^T, dP = dO V^T, dV = P^T dO,
//     dK^ = dS^T Q^, dQ^ = dS K^ are M = 64 MMAs, one chain per unit: the accumulator rows of an M = 64 MMA land in
//     lanes 32 (r / 16) + r % 16 and a lane offset of 16 moves them to the other half of every 32-lane quadrant, so the
//     two units of a pair interleave in the 128 lanes and share columns.  dBias += dS I accumulates over all windows
//     of the CTA in 64 columns;
//   * two softmax groups (8 warps each) work on alternate pairs.  A thread owns half a query row: pass 1 streams S and
//     dP from TMEM 16 keys at a time, forms P (staged to shared memory as bf16, and kept packed), the partial sums of
//     P dP, P dP t and P t (t = tau log2e cos); the two half-row owners swap their partial D through shared memory
//     (64-thread named barrier); pass 2 forms dS = P (dP - D) from the packed registers and stages it.  P / dS tiles
//     are [query][key] with 128-byte rows (SWIZZLE_128B): read MN-major they are the A operands P^T / dS^T, read
//     K-major dS -- no thread transposes anything.  sum_j dS_ij t_ij = A1 - D A2 is the d(tau) contribution and the
//     q^.M of the dQ epilogue;
//   * after its own arithmetic the group normalises the q / k tiles of its pair in place (q^ = r q, k^ = k / |k|):
//     the B operands of dK^ / dQ^ and the x^ of dx = scale (M - (x^.M) x^) in the epilogue;
//   * one token order per launch: a shifted layer loads / stores EVERY window as two column parts [0, 8-s) | [8-s, 8)
//     (2 boxes per tile, 4 on the bottom row); the position bias is a Toeplitz lookup (4 alignment copies of the
//     15 x 15 table per head, 10 KB);
//   * 28 warps: 0 TMA loads | 1 issuer of S, dP | 2 issuer of dV, dK^, dQ^, dBias | 3 TMA stores |
//     4-19 softmax / dS (2 groups) | 20-23 dV, dK epilogue | 24-27 dQ epilogue.  All hand-overs are mbarriers.
#include "hv_tc_win.cuh"
#include <atomic>

namespace hv {
namespace {
using namespace tc;

constexpr int kN = 64;
constexpr int kWs = 8;
constexpr int kTab = 225;
constexpr int kTile = kN * 64;         // one (window, head) q / k / v / dO tile: 64 rows x 64 B (SWIZZLE_64B)
// Two rings: v is dead as soon as dP exists, q, k and dO live until the stores of dq, dk, dv have left the stage.
constexpr int kStage = 6 * kTile;      // late ring:  q_a q_b k_a k_b g_a g_b   (g = dO)
constexpr int kStages = 5;
constexpr int kStageE = 2 * kTile;     // early ring: v_a v_b
constexpr int kStagesE = 2;
constexpr int kThreads = 896;          // 28 warps
constexpr int kPdTile = kN * 128;      // P or dS of one unit: 64 rows x 128 B (SWIZZLE_128B)
constexpr int kBiasRow = 20;           // floats per table row (15 + alignment slack)
constexpr int kBiasCopy = 328;         // floats per alignment copy: >= 15 * 20 and = 8 (mod 32) so 8 lanes hit 8 bank groups
constexpr int kStatBytes = kN * 4;     // one plane (lse | r | c) of one unit

// ---- shared memory map (dynamic, 1024-byte aligned base)
constexpr int kOffStage = 0;
constexpr int kOffStageE = kOffStage + kStages * kStage;
constexpr int kOffP = kOffStageE + kStagesE * kStageE;    // [2 buffers][2 units][64][128 B]
constexpr int kOffDS = kOffP + 4 * kPdTile;
constexpr int kOffEye = kOffDS + 4 * kPdTile;             // 16 rows x 128 B: 16 x 16 bf16 identity in K columns 0-15 (SWIZZLE_128B)
constexpr int kEyeBytes = 16 * 128;
constexpr int kOffBias = kOffEye + kEyeBytes;             // [2 units][4 copies][kBiasCopy] float
constexpr int kOffVec = kOffBias + 2 * 4 * kBiasCopy * 4; // [kStages][3: lse, r, c][2 units][64] float, SLOT order
// [kStages pairs in flight][2 halves][128] float: first the partial D of the half row (swapped between the two owners of
// a row), then sum_j dS_ij t_ij of the half row (read by the dQ epilogue)
constexpr int kOffDot = kOffVec + kStages * 3 * 128 * 4;
constexpr int kOffCol = kOffDot + kStages * 2 * 128 * 4;  // [2][32] float dq column sums, [2] d(tau)
constexpr int kOffBins = kOffP;                           // [2][256] float: d(bias) bins, after the main loop (aliases P)
constexpr int kOffGeo = kOffCol + (2 * 32 + 4) * 4;       // [8][2] UnitGeo
constexpr int kOffSlotMap = kOffGeo + 8 * 2 * 16;         // [64] bytes
constexpr int kOffBar = kOffSlotMap + 64;
constexpr int kNumBars = 5 * kStages + 2 * kStagesE + 10;
constexpr int kOffTmem = kOffBar + kNumBars * 8;
constexpr int kSmem = kOffTmem + 16;
static_assert(kOffP % 1024 == 0 && kOffEye % 1024 == 0 && kOffBias % 16 == 0 && kOffVec % 16 == 0 && kOffBar % 8 == 0, "shared-memory alignment");
static_assert(kSmem <= 227 * 1024, "shared memory budget");

// TMEM columns
// S and dP of a pair take 64 columns each: the two units are separate M = 64 MMAs whose accumulators interleave in the
// 128 lanes (rows 16q .. 16q + 15 of unit a in lanes 32q .. 32q + 15, of unit b in lanes 32q + 16 .. 32q + 31; measured with
// tools/probes/umma_m64_probe.cu) -- no wasted off-diagonal blocks, and room for two buffers
constexpr int kColS = 0, kColDP = 128;                                  // + 64 * buffer (one per softmax group)
constexpr int kColDV = 256, kColDK = 288, kColDQ = 320, kAccCols = 96;  // + 96 * buffer: M = 64 output MMAs, 32 columns each
constexpr int kColDB = 448;

struct BwdParams {
  Geom g;
  WinSchedule sched;  // CTAs per head group and window class (hv_tc_win.cuh)
  int64_t plane;      // floats per plane of the forward's statistics: B * nW * heads * 64
  int ko;             // HV_TC_TRACE builds only: knock-out bits (results are wrong) | traced CTA << 8
};
struct BwdMaps { CUtensorMap m[3][kNumWinMaps]; };  // qkv, dout, dqkv

#ifdef HV_TC_TRACE
__device__ long long* g_btrace = nullptr;  // [pairs][16 events] clock64 stamps of CTA 0
#define TRACE(k, ev) do { if (blockIdx.x == (p.ko >> 8) && lane == 0 && g_btrace && (k) < 64) g_btrace[(k) * 16 + (ev)] = clock64(); } while (0)
#define KO(bit) (p.ko & (bit))  // 1: no tile loads | 2: no tile stores
#else
#define TRACE(k, ev) do { } while (0)
#define KO(bit) false
#endif

// kMode 0: unshifted layer | 1: shifted layer, windows without column wrap (slot order) | 2: shifted layer, right-edge
// windows (column-split order)
template <int kMode>
__device__ __forceinline__ void wattn_tc64_bwd_body(const BwdMaps& maps, const float* __restrict__ stats,
                                                    const float* __restrict__ bias_table, const float* __restrict__ tau,
                                                    float* __restrict__ ws_dbias, float* __restrict__ ws_dtau, const BwdParams& p) {
  constexpr bool kSplit = kMode == 2;
  constexpr bool kMasked = kMode > 0;
  extern __shared__ __align__(1024) unsigned char smem[];
  const Geom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar0 = sb + kOffBar;
  auto bar_full = [&](int s) { return bar0 + 8 * s; };                      // q, k, dO tiles + statistics landed
  auto bar_empty = [&](int s) { return bar0 + 8 * (kStages + s); };         // the stores of dq, dk, dv have read the stage
  auto bar_written = [&](int s) { return bar0 + 8 * (2 * kStages + s); };   // epilogue wrote dq, dk, dv over the tiles
  auto bar_hat = [&](int s) { return bar0 + 8 * (3 * kStages + s); };       // q, k tiles normalised in place
  auto bar_sdp = [&](int s) { return bar0 + 8 * (4 * kStages + s); };       // S, dP accumulators complete
  auto bar_fullE = [&](int s) { return bar0 + 8 * (5 * kStages + s); };
  auto bar_emptyE = [&](int s) { return bar0 + 8 * (5 * kStages + kStagesE + s); };
  const uint32_t barx = bar0 + 8 * (5 * kStages + 2 * kStagesE);
  auto bar_staged = [&](int b) { return barx + 8 * b; };        // P / dS staging buffer b written
  auto bar_stfree = [&](int b) { return barx + 8 * (2 + b); };  // ... and read by the MMAs
  auto bar_sfree = [&](int b) { return barx + 8 * (4 + b); };   // S / dP buffer b read by its softmax group
  auto bar_acc = [&](int a) { return barx + 8 * (6 + a); };     // output MMAs of accumulator buffer a complete
  auto bar_accfree = [&](int a) { return barx + 8 * (8 + a); }; // epilogue has pulled the accumulators out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmem);
  const int nrows = g.B * g.nW;

  __shared__ CtaWork s_work;
  if (threadIdx.x == 0) {
    CtaWork w0;
    w0.init(p.g, p.sched, blockIdx.x);
    s_work = w0;
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);    // the store warp
      mbar_init(bar_written(s), 8);  // the eight epilogue warps
      mbar_init(bar_hat(s), 8);
      mbar_init(bar_sdp(s), 1);
    }
    for (int s = 0; s < kStagesE; ++s) {
      mbar_init(bar_fullE(s), 1);
      mbar_init(bar_emptyE(s), 1);  // the commit behind the dP MMAs
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_staged(b), 8);
      mbar_init(bar_stfree(b), 1);
      mbar_init(bar_sfree(b), 8);
      mbar_init(bar_acc(b), 1);
      mbar_init(bar_accfree(b), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // ---- one-time tables: identity tile, slot of every tile row, Toeplitz bias copies (log2 units), zeroed reduction bins
  unsigned char* slotmap = smem + kOffSlotMap;
  for (int idx = threadIdx.x; idx < kEyeBytes / 4; idx += kThreads) {
    // 32-bit word idx of the swizzled identity: row = byte / 128, physical chunk = (byte / 16) & 7, logical chunk = phys ^ (row & 7)
    const int byte = idx * 4, row = byte >> 7, chunk = ((byte >> 4) & 7) ^ (row & 7);
    const int col = chunk * 8 + ((byte & 15) >> 1);  // first of the two bf16 columns of this word
    uint32_t w = 0;
    if (col == row) w = 0x00003F80u;
    else if (col + 1 == row) w = 0x3F800000u;
    reinterpret_cast<uint32_t*>(smem + kOffEye)[idx] = w;
  }
  if (threadIdx.x < 64) slotmap[threadIdx.x] = (unsigned char)tile_row_slot(threadIdx.x, kSplit ? g.shift : 0);
  for (int idx = threadIdx.x; idx < 2 * 32 + 4; idx += kThreads) reinterpret_cast<float*>(smem + kOffCol)[idx] = 0.f;
  {
    CtaWork w0;
    w0.init(p.g, p.sched, blockIdx.x);
    float* bt = reinterpret_cast<float*>(smem + kOffBias);
    for (int idx = threadIdx.x; idx < 2 * 4 * kBiasCopy; idx += kThreads) {
      const int u = idx / (4 * kBiasCopy), rem = idx - u * 4 * kBiasCopy;
      const int c = rem / kBiasCopy, q = rem - c * kBiasCopy;
      const int dh = q / kBiasRow, pos = q - dh * kBiasRow;
      const int x = pos - 4 + c;  // x = 7 - iw + jw: reversed column difference
      float v = 0.f;
      if (dh < 15 && x >= 0 && x <= 14) v = kLog2e * __ldg(&bias_table[(dh * 15 + 14 - x) * g.heads + (u == 0 ? w0.head_a : w0.head_b)]);
      bt[idx] = v;
    }
  }
  fence_async_smem();  // identity tile is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const CtaWork work = s_work;
  const int npairs = work.npairs;
  UnitGeo* geo = reinterpret_cast<UnitGeo*>(smem + kOffGeo);
  float* vecs = reinterpret_cast<float*>(smem + kOffVec);

  if (warp < 4) {
    reg_dealloc<40>();
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer: lanes 0-7 load one tile each (q, k, v,
      // dO of the two units), lanes 8-13 one plane (lse, r, c) of the forward's statistics of one unit
      // the windows of this CTA: a cursor over its class's grid (cross group: two consecutive windows per pair)
      const int nWh = g.H / kWs;
      WinCursor cur;
      cur.init(work.cross ? 2 * work.first : work.first, work.cross ? 2 * work.stride : work.stride, work.wcls, work.hcls);
      for (int k = 0; k < npairs; ++k, cur.advance()) {
        const int s = k % kStages, se = k % kStagesE;
        mbar_wait_fast(bar_empty(s), ((k / kStages) & 1) ^ 1);
        mbar_wait_fast(bar_emptyE(se), ((k / kStagesE) & 1) ^ 1);
        TRACE(k, 0);
        // Every operand of the TMA instructions below is warp-uniform and ONE elected lane issues them all: a
        // `cp.async.bulk.tensor` whose operands differ per lane compiles into a loop over the active lanes
        int ur[2], ub[2], urow0[2], ucol0[2];
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
          ur[u] = r; ub[u] = b;
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
        // ONE elected lane issues every TMA instruction of the pair with warp-uniform operands (a `cp.async.bulk.tensor`
        // whose operands differ per lane compiles into a loop over the active lanes).  Unrolled where a tile is one box;
        // rolled on shifted layers, whose 8 tiles x up to 6 box variants would add ~10 KB to the ~35 KB of loop bodies the
        // 28 warps of the CTA run through (measured: the unrolled form made every role of such a CTA ~2x slower --
        // instruction-cache misses)
        if (elect_one()) {
          mbar_expect_tx(bar_full(s), (KO(1) ? 0 : kStage) + 6 * kStatBytes);
          mbar_expect_tx(bar_fullE(se), KO(1) ? 0 : kStageE);
          auto issue_tile = [&](int t) {
            const int u = t & 1, kind = t >> 1;  // kind: 0 q, 1 k, 2 v, 3 dO
            const int head = u == 0 ? work.head_a : work.head_b;
            const int c0 = (kind < 3 ? kind * g.C : 0) + head * 32;
            const bool early = kind == 2;
            const int tidx = (kind == 0 ? 0 : (kind == 1 ? 2 : (kind == 3 ? 4 : 0))) + u;
            const uint32_t dst = early ? sb + kOffStageE + se * kStageE + tidx * kTile : sb + kOffStage + s * kStage + tidx * kTile;
            const uint32_t bar = early ? bar_fullE(se) : bar_full(s);
            const CUtensorMap* mm = maps.m[kind == 3 ? 1 : 0];
            const int bb = u ? ub[1] : ub[0];
            if (!KO(1))
              for_each_box<kMode>(g, u ? ucol0[1] : ucol0[0], u ? urow0[1] : urow0[0], u ? ubottom[1] : ubottom[0],
                                  [&](int off, int mi, int col, int row) { tma_load_4d(dst + off, &mm[mi], bar, c0, col, row, bb); });
          };
          auto issue_stats = [&](int t) {
            const int u = t & 1, plane = t >> 1;
            const int head = u == 0 ? work.head_a : work.head_b;
            bulk_load(sb + kOffVec + ((s * 3 + plane) * 128 + u * 64) * 4,
                      stats + plane * p.plane + ((int64_t)(u ? ur[1] : ur[0]) * g.heads + head) * kN, kStatBytes, bar_full(s));
          };
          if (kMasked) {
#pragma unroll 1
            for (int t = 0; t < 8; ++t) issue_tile(t);
#pragma unroll 1
            for (int t = 0; t < 6; ++t) issue_stats(t);
          } else {
#pragma unroll
            for (int t = 0; t < 8; ++t) issue_tile(t);
#pragma unroll
            for (int t = 0; t < 6; ++t) issue_stats(t);
          }
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- issuer of S = Q K^T and dP = dO V^T
      const uint32_t id = idesc_bf16(64, 64, 0, 0);
      const uint64_t d_q = smem_desc(sb + kOffStage, 16, 512, 4), d_k = smem_desc(sb + kOffStage + 2 * kTile, 16, 512, 4);
      const uint64_t d_v = smem_desc(sb + kOffStageE, 16, 512, 4), d_g = smem_desc(sb + kOffStage + 4 * kTile, 16, 512, 4);
      for (int k = 0; k < npairs; ++k) {
        const int s = k % kStages, se = k % kStagesE, buf = k & 1;
        mbar_wait_fast(bar_full(s), (k / kStages) & 1);
        mbar_wait_fast(bar_fullE(se), (k / kStagesE) & 1);
        TRACE(k, 1);
        if (k > 1) mbar_wait_fast(bar_sfree(buf), ((k >> 1) - 1) & 1);  // the group of pair k-2 has read this S / dP buffer
        tc_fence_after();
        if (elect_one()) {
          const uint64_t so = (uint64_t)((s * kStage) >> 4), soe = (uint64_t)((se * kStageE) >> 4);
#pragma unroll
          for (int u = 0; u < 2; ++u) {  // unit u: tiles q_u, k_u (one tile = 4 KB further), accumulator lanes + 16 u
            const uint32_t dl = (uint32_t)(16 * u) << 16;
            const uint64_t uo = (uint64_t)(u * (kTile >> 4));
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
              umma_ss(tmem + dl + kColS + 64 * buf, d_q + so + uo + 2 * kk, d_k + so + uo + 2 * kk, id, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
              umma_ss(tmem + dl + kColDP + 64 * buf, d_g + so + uo + 2 * kk, d_v + soe + uo + 2 * kk, id, kk > 0);
          }
          umma_commit(bar_sdp(s));
          umma_commit(bar_emptyE(se));  // v is dead once dP exists
          TRACE(k, 2);
        }
        __syncwarp();
      }
    } else if (warp == 2) {
      // ---------------------------------------------------------------- issuer of dV, dK^, dQ^, dBias
      // dV, dK^, dQ^: one M = 64 MMA chain per unit (accumulator rows interleave in the lanes like S / dP: unit b at lane
      // offset 16), 32 columns each and two accumulator buffers, so the MMAs of pair k+1 never wait for the epilogue of
      // pair k.  dBias stays one stacked M = 128 chain accumulating over all pairs.
      const uint32_t id_t = idesc_bf16(64, 32, 1, 1);    // A = P^T / dS^T (MN-major), B = dO / q^ (MN-major)
      const uint32_t id_q = idesc_bf16(64, 32, 0, 1);    // A = dS (K-major), B = k^ (MN-major)
      // dBias += dS I, 16 keys at a time: A = dS of both units (K-major, 16 key columns), B = a 16 x 16 identity, D = the
      // same 16 columns of the accumulator
      const uint32_t id_b = idesc_bf16(128, 16, 0, 0);
      // A, MN-major view of a [query][key] tile: 64 keys = one 128-byte atom, 8 queries = 1 KB (SBO)
      const uint64_t a_pt = smem_desc(sb + kOffP, 16, 1024, 2), a_dst = smem_desc(sb + kOffDS, 16, 1024, 2);
      // A, K-major view: query rows of 128 B, 8-row groups 1 KB apart (unit b's tile follows unit a's: 128 rows for dBias)
      const uint64_t a_ds = smem_desc(sb + kOffDS, 16, 1024, 2);
      const uint64_t b_eye = smem_desc(sb + kOffEye, 16, 1024, 2);
      // B, MN-major view of a 64 x 64-byte tile: 32 channels = one 64-byte atom, 8 tokens = 512 B (SBO)
      const uint64_t b_q = smem_desc(sb + kOffStage, 16, 512, 4), b_k = smem_desc(sb + kOffStage + 2 * kTile, 16, 512, 4);
      const uint64_t b_g = smem_desc(sb + kOffStage + 4 * kTile, 16, 512, 4);
      for (int k = 0; k < npairs; ++k) {
        const int s = k % kStages;
        const int buf = k & 1;
        mbar_wait_fast(bar_staged(buf), (k >> 1) & 1);
        TRACE(k, 10);
        mbar_wait_fast(bar_hat(s), (k / kStages) & 1);
        if (k > 1) mbar_wait_fast(bar_accfree(buf), ((k >> 1) - 1) & 1);  // epilogue of pair k-2 has drained this buffer
        tc_fence_after();
        if (elect_one()) {
          const uint64_t so = (uint64_t)((s * kStage) >> 4), bo = (uint64_t)(buf * ((2 * kPdTile) >> 4));
          const uint32_t acc = tmem + kAccCols * buf;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const uint32_t dl = (uint32_t)(16 * u) << 16;
            const uint64_t ao = bo + (uint64_t)(u * (kPdTile >> 4)), to = so + (uint64_t)(u * (kTile >> 4));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)  // 16 queries per step: A += 2 KB, B += 1 KB
              umma_ss(acc + dl + kColDV, a_pt + ao + (uint64_t)(128 * ks), b_g + to + (uint64_t)(64 * ks), id_t, ks > 0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_ss(acc + dl + kColDK, a_dst + ao + (uint64_t)(128 * ks), b_q + to + (uint64_t)(64 * ks), id_t, ks > 0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)  // 16 keys per step: A += 32 B inside the swizzle atom, B += 1 KB
              umma_ss(acc + dl + kColDQ, a_ds + ao + (uint64_t)(2 * ks), b_k + to + (uint64_t)(64 * ks), id_q, ks > 0);
          }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ss(tmem + kColDB + 16 * ks, a_ds + bo + (uint64_t)(2 * ks), b_eye, id_b, k > 0 ? 1u : 0u);
          umma_commit(bar_acc(buf));
          umma_commit(bar_stfree(buf));
          TRACE(k, 12);
        }
        __syncwarp();
      }
    } else {
      // ---------------------------------------------------------------- warp 3: TMA stores of dq, dk, dv + dq column sums
      // The epilogue warps have written dq over the q tile, dk over the k tile and dv over the dO tile of the stage (same
      // swizzled layout the loads produced), so the boxes of the loads are the boxes of the stores.
      // (The gradient of q_bias -- the column sums of dq -- is a separate streaming kernel below: summed here, by this one
      // warp, it cost 30 % of the whole kernel.)
      const CUtensorMap* mm = maps.m[2];
      for (int k = 0; k < npairs; ++k) {
        const int s = k % kStages;
        mbar_wait_fast(bar_written(s), (k / kStages) & 1);
        TRACE(k, 14);
        const uint32_t st = sb + kOffStage + s * kStage;
        if (elect_one()) {  // warp-uniform operands, one issuing lane (see the producer)
          auto store_tile = [&](int t) {
            const int u = t & 1, kind = t >> 1;  // kind: 0 dq (q tile), 1 dk (k tile), 2 dv (dO tile)
            const UnitGeo ug = geo[(k & 7) * 2 + u];
            const int head = u == 0 ? work.head_a : work.head_b;
            if ((ug.rflags & 1) && !KO(2)) {
              const uint32_t src = st + (2 * kind + u) * kTile;
              const int c0 = kind * g.C + head * 32;
              for_each_box<kMode>(g, ug.col0, ug.row0, (ug.rflags & 2) != 0, [&](int off, int mi, int col, int row) {
                tma_store_4d(&mm[mi], src + off, c0, col, row, ug.b);
              });
            }
          };
          if (kMasked) {
#pragma unroll 1
            for (int t = 0; t < 6; ++t) store_tile(t);
          } else {
#pragma unroll
            for (int t = 0; t < 6; ++t) store_tile(t);
          }
        }
        __syncwarp();
        bulk_commit();
        bulk_wait_read0();
        __syncwarp();
        TRACE(k, 15);
        if (lane == 0) mbar_arrive(bar_empty(s));
      }
      bulk_wait0();
    }
  } else if (warp < 20) {
    // ------------------------------------------------------------------ softmax / dS threads: two groups (warps 4-11 even
    // pairs, 12-19 odd pairs) so that the hand-over latencies of one group hide behind the arithmetic of the other; a thread
    // owns half a logit row.  Lanes 0-15 of a warp are rows of unit a, lanes 16-31 of unit b (M = 64 layout).
    reg_alloc<80>();
    const int grp = (warp - 4) >> 3;
    const int half = ((warp - 4) >> 2) & 1;
    const int quad = warp & 3;
    const int u = lane >> 4, i = 16 * quad + (lane & 15);  // unit of the pair, tile row (query) inside the unit
    const int row = 64 * u + i;                            // row of the pair in the per-pair vectors
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16);
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
    const uint32_t p_row = sb + kOffP + grp * 2 * kPdTile + u * kPdTile + i * 128;
    const uint32_t ds_row = sb + kOffDS + grp * 2 * kPdTile + u * kPdTile + i * 128;
    const uint32_t tS = tl + kColS + 64 * grp + 32 * half, tP = tl + kColDP + 64 * grp + 32 * half;
    float acc_tau = 0.f;
    float* dots = reinterpret_cast<float*>(smem + kOffDot);
    // In-place normalisation of the q / k tiles of the group's own pair once S has been computed from the raw tiles
    // (q^ = q / |q|, k^ = k / |k|): half a tile (one row per lane) per warp, after the group's own arithmetic and under the
    // same fence.proxy.async as its staging stores -- independent of the accumulator / epilogue chain
    const int hw = 4 * half + quad, htile = hw >> 1, hpart = htile >> 1, hu = htile & 1, hrow = 32 * (hw & 1) + lane;
    const int hslot = slotmap[hrow];
    const float inv_mult = hpart == 0 ? 1.0f : 1.0f / (__ldg(&tau[hu == 0 ? work.head_a : work.head_b]) * kLog2e);
    auto hat = [&](int k) {
      const int s = k % kStages;
      const uint32_t tile = sb + kOffStage + s * kStage + (2 * hpart + hu) * kTile;
      const float* vec = vecs + (s * 3 + 1 + hpart) * 128 + 64 * hu;
      uint4 v[4];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) v[ch] = lds128(tile + hrow * 64 + ((ch ^ ((hrow >> 1) & 3)) << 4));
      const float sc = vec[hslot] * inv_mult;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint4 w = v[ch];
        w.x = pack_bf16x2(bf16lo_to_f32(w.x) * sc, bf16hi_to_f32(w.x) * sc);
        w.y = pack_bf16x2(bf16lo_to_f32(w.y) * sc, bf16hi_to_f32(w.y) * sc);
        w.z = pack_bf16x2(bf16lo_to_f32(w.z) * sc, bf16hi_to_f32(w.z) * sc);
        w.w = pack_bf16x2(bf16lo_to_f32(w.w) * sc, bf16hi_to_f32(w.w) * sc);
        sts128(tile + hrow * 64 + ((ch ^ ((hrow >> 1) & 3)) << 4), w);
      }
    };

    for (int k = grp; k < npairs; k += 2) {
      const int s = k % kStages;
      const uint32_t ph = (k / kStages) & 1;
      mbar_wait_fast(bar_full(s), ph);  // statistics (and tiles) of the pair have landed
      if (warp == 4) TRACE(k, 5);
      const int rflags = geo[(k & 7) * 2 + u].rflags;
      const float* vec = vecs + s * 3 * 128;
      // lse = 1e30 for the padding unit of an odd tail: P = dS = 0 there
      const float li = (rflags & 1) ? vec[64 * u + si] : 1e30f;
      const float ri = vec[128 + 64 * u + si];
      // c_j of keys 4q .. 4q + 3 of this half (slot order): slot order = tile order | split order: window row q, columns 4 half ..
      const float* cv = vec + 256 + 64 * u + (kSplit ? 4 * half : 32 * half);
      uint32_t m = 0u;
      if (kMasked) m = ((rflags & 2) ? mH : 0u) | ((rflags & 4) ? mW : 0u);
      if (k > 1) mbar_wait_fast(bar_stfree(grp), ((k >> 1) - 1) & 1);  // the MMAs of pair k-2 have read this group's staging tiles
      mbar_wait_fast(bar_sdp(s), ph);
      if (warp == 4) TRACE(k, 6);
      tc_fence_after();
      // pass 1: P (staged, and kept packed), dP kept packed (the reference's dP is a bf16 matmul output as well), partial
      // sums over this half row: Dp = sum P dP, A1 = sum P dP t, A2 = sum P t
      float Dp = 0.f, A1 = 0.f, A2 = 0.f;
      uint32_t pk[16], dk[16];
      auto chunk = [&](auto masked, auto ck_tag) {
        constexpr int ck = decltype(ck_tag)::value;  // keys 16 ck .. 16 ck + 15 of this half
        uint32_t sa[16], pa[16];
        HV_TMEM_LD16(tS + 16 * ck, sa);
        HV_TMEM_LD16(tP + 16 * ck, pa);
        tmem_wait_ld();
        if (ck == 1) {  // the whole half row has left TMEM: hand the buffer back to the S / dP issuer
          tc_fence_before();
          __syncwarp();
          if (warp == 4) TRACE(k, 7);
          if (lane == 0) mbar_arrive(bar_sfree(grp));
        }
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {  // eight keys = one 16-byte staging chunk of P
#pragma unroll
          for (int q2 = 0; q2 < 2; ++q2) {
            const int q = 4 * ck + 2 * h8 + q2;
            // keys 4q .. 4q + 3 of this half: slot order = window row 4 half + q / 2, columns 4 (q & 1) ..;
            // split order = window row q, columns 4 half ..
            const float4 b = *reinterpret_cast<const float4*>(bias_base + (kSplit ? -q * kBiasRow : -(q >> 1) * kBiasRow + 4 * (q & 1)));
            const float4 c = *reinterpret_cast<const float4*>(cv + (kSplit ? 8 * q : 4 * q));
            const float bb[4] = {b.x, b.y, b.z, b.w}, cc[4] = {c.x, c.y, c.z, c.w};
            float pv[4], dv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = 4 * q + e, jj = 8 * h8 + 4 * q2 + e;
              const float t = (__uint_as_float(sa[jj]) * ri) * cc[e];  // tau log2e cos(q_i, k_j)
              float x = (t + bb[e]) - li;
              if (decltype(masked)::value && ((m >> j) & 1u)) x += kNeg;
              const float pe = ex2(x);
              const float dp = __uint_as_float(pa[jj]);
              const float ge = pe * dp;
              Dp += ge;
              A1 = fmaf(ge, t, A1);
              A2 = fmaf(pe, t, A2);
              pv[e] = pe;
              dv[e] = dp;
            }
            const int w = 8 * ck + 4 * h8 + 2 * q2;
            pk[w] = pack_bf16x2(pv[0], pv[1]);
            pk[w + 1] = pack_bf16x2(pv[2], pv[3]);
            dk[w] = pack_bf16x2(dv[0], dv[1]);
            dk[w + 1] = pack_bf16x2(dv[2], dv[3]);
          }
          const int w0 = 8 * ck + 4 * h8;
          const uint32_t off = (uint32_t)(((4 * half + 2 * ck + h8) ^ (i & 7)) << 4);
          sts128(p_row + off, make_uint4(pk[w0], pk[w0 + 1], pk[w0 + 2], pk[w0 + 3]));
        }
      };
      // one instantiation per kernel mode: a second (unmasked) copy of this loop body for the interior windows of a shifted
      // layer costs more in instruction-cache misses than the two predicated instructions per logit it would save
      chunk(BoolTag<kMasked>{}, IntTag<0>{});
      chunk(BoolTag<kMasked>{}, IntTag<1>{});
      // D = both halves' partial sums: swap through shared memory with the warp that owns the other half of these rows
      float* dpp = dots + s * 256;
      dpp[half * 128 + row] = Dp;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + grp * 4 + quad) : "memory");
      const float Di = Dp + dpp[(half ^ 1) * 128 + row];
      if (warp == 4) TRACE(k, 8);
      // pass 2: dS = P (dP - D)
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        uint32_t dd[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const uint32_t pw = pk[4 * c8 + w], dw = dk[4 * c8 + w];
          dd[w] = pack_bf16x2(bf16lo_to_f32(pw) * (bf16lo_to_f32(dw) - Di), bf16hi_to_f32(pw) * (bf16hi_to_f32(dw) - Di));
        }
        const uint32_t off = (uint32_t)(((4 * half + c8) ^ (i & 7)) << 4);
        sts128(ds_row + off, make_uint4(dd[0], dd[1], dd[2], dd[3]));
      }
      const float racc = fmaf(-Di, A2, A1);  // sum_j dS_ij t_ij over this half row: d(tau) contribution, the dQ epilogue's q^.M
      acc_tau += racc;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + grp * 4 + quad) : "memory");  // the other owner has read this row's partial D
      dpp[half * 128 + row] = racc;
      hat(k);
      fence_async_smem();
      __syncwarp();
      if (warp == 4) TRACE(k, 9);
      if (lane == 0) {
        mbar_arrive(bar_staged(grp));
        mbar_arrive(bar_hat(s));
      }
    }
    // d(tau) = sum dS cos = sum dS t / (tau log2e); lanes 0-15 / 16-31 of a warp belong to unit a / b
    const float tot = group_sum<16>(acc_tau);
    if ((lane & 15) == 0) {
      const float tu = __ldg(&tau[u == 0 ? work.head_a : work.head_b]) * kLog2e;
      atomicAdd(reinterpret_cast<float*>(smem + kOffCol) + 64 + u, tot / tu);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps: 20-23 dV and dK, 24-27 dQ.  A thread
    // owns one token row: accumulators from TMEM, projection of the normalised-row gradient back to the raw row, result
    // written as bf16 over the row of the dO (dv), k (dk) or q (dq) tile -- the store warp sends the tiles out by TMA
    auto epilogue = [&](auto role_tag) {
    constexpr int role = decltype(role_tag)::value;  // 1 dV + dK, 2 dQ
    const int quad = warp & 3;
    const int u = lane >> 4, t = 16 * quad + (lane & 15);  // M = 64 accumulator layout: lanes 0-15 unit a, 16-31 unit b
    const int row = 64 * u + t;
    const int head = u == 0 ? work.head_a : work.head_b;
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16);
    const int sl = slotmap[t];
    const float tau_h = __ldg(&tau[head]);
    const float inv_tl = 1.0f / (tau_h * kLog2e);
    const uint32_t swz = (uint32_t)((t >> 1) & 3);

    for (int k = 0; k < npairs; ++k) {
      const int s = k % kStages;
      const int ab = k & 1;
      mbar_wait_fast(bar_acc(ab), (k >> 1) & 1);
      if (warp == 20) TRACE(k, 13);
      tc_fence_after();
      uint32_t a[32];
      const uint32_t st = sb + kOffStage + s * kStage;
      if (role == 1) {  // dV first: pack and write over the dO row, then the same registers take dK
        HV_TMEM_LD32(tl + kAccCols * ab + kColDV, a);
        tmem_wait_ld();
        HV_REG_FENCE32(a);
        const uint32_t grow = st + (4 + u) * kTile + t * 64;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(a[8 * q + 0]), __uint_as_float(a[8 * q + 1]));
          v.y = pack_bf16x2(__uint_as_float(a[8 * q + 2]), __uint_as_float(a[8 * q + 3]));
          v.z = pack_bf16x2(__uint_as_float(a[8 * q + 4]), __uint_as_float(a[8 * q + 5]));
          v.w = pack_bf16x2(__uint_as_float(a[8 * q + 6]), __uint_as_float(a[8 * q + 7]));
          sts128(grow + ((q ^ swz) << 4), v);
        }
      }
      HV_TMEM_LD32(tl + kAccCols * ab + (role == 1 ? kColDK : kColDQ), a);
      float qdot = 0.f;
      if (role == 2) {
        const float* dp = reinterpret_cast<const float*>(smem + kOffDot) + s * 256 + row;
        qdot = (dp[0] + dp[128]) * inv_tl;
      }
      tmem_wait_ld();
      HV_REG_FENCE32(a);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_accfree(ab));
      // projection of the gradient of the normalised row back to the raw row: d x = sc (M - (x^ . M) x^)
      const float* vec = vecs + s * 3 * 128;
      const uint32_t xrow = st + ((role == 2 ? 0 : 2) + u) * kTile + t * 64;
      const float sc = role == 2 ? vec[128 + 64 * u + sl] * tau_h : vec[256 + 64 * u + sl] * kLn2;  // tau / |q_i|  |  tau / |k_j| = c_j ln 2
      if (role == 1) {
        uint32_t xh[16];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const uint4 v = lds128(xrow + ((ch ^ swz) << 4));
          xh[4 * ch] = v.x; xh[4 * ch + 1] = v.y; xh[4 * ch + 2] = v.z; xh[4 * ch + 3] = v.w;
        }
        float dot = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          dot = fmaf(bf16lo_to_f32(xh[e]), __uint_as_float(a[2 * e]), dot);
          dot = fmaf(bf16hi_to_f32(xh[e]), __uint_as_float(a[2 * e + 1]), dot);
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v0 = sc * fmaf(-dot, bf16lo_to_f32(xh[4 * ch + e]), __uint_as_float(a[8 * ch + 2 * e]));
            const float v1 = sc * fmaf(-dot, bf16hi_to_f32(xh[4 * ch + e]), __uint_as_float(a[8 * ch + 2 * e + 1]));
            o[e] = pack_bf16x2(v0, v1);
          }
          sts128(xrow + ((ch ^ swz) << 4), make_uint4(o[0], o[1], o[2], o[3]));
        }
      } else {
        // q^_i . M_i = sum_j dS_ij cos_ij: the softmax threads already have it (two half-row sums of dS t, t = tau log2e cos)
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const uint4 v = lds128(xrow + ((ch ^ swz) << 4));
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v0 = sc * fmaf(-qdot, bf16lo_to_f32(w[e]), __uint_as_float(a[8 * ch + 2 * e]));
            const float v1 = sc * fmaf(-qdot, bf16hi_to_f32(w[e]), __uint_as_float(a[8 * ch + 2 * e + 1]));
            o[e] = pack_bf16x2(v0, v1);
          }
          sts128(xrow + ((ch ^ swz) << 4), make_uint4(o[0], o[1], o[2], o[3]));
        }
      }
      fence_async_smem();  // the tiles are read by the TMA engine (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_written(s));
    }
    };
    if (warp < 24) epilogue(IntTag<1>{}); else epilogue(IntTag<2>{});
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  for (int idx = threadIdx.x; idx < 2 * 256; idx += kThreads) reinterpret_cast<float*>(smem + kOffBins)[idx] = 0.f;
  __syncthreads();
  // ---- d(bias): fold the 64 x 64 accumulators of the two units into the 225 table bins
  if (warp >= 24 && warp < 28 && npairs > 0) {
    const int quad = warp & 3, row = quad * 32 + lane, u = row >> 6, t = row & 63;
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16);
    const int si = slotmap[t];
    float* bins = reinterpret_cast<float*>(smem + kOffBins) + u * 256;
#pragma unroll 1
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t a[32];
      HV_TMEM_LD32(tl + kColDB + 32 * hh, a);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int sj = slotmap[32 * hh + j];
        const int rel = ((si >> 3) - (sj >> 3) + kWs - 1) * (2 * kWs - 1) + ((si & 7) - (sj & 7) + kWs - 1);
        atomicAdd(&bins[rel], __uint_as_float(a[j]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  {
    const float* bins = reinterpret_cast<const float*>(smem + kOffBins);
    const float* col = reinterpret_cast<const float*>(smem + kOffCol);
    for (int idx = threadIdx.x; idx < 2 * kTab; idx += kThreads) {
      const int u = idx / kTab, r = idx - u * kTab;
      ws_dbias[((int64_t)blockIdx.x * 2 + u) * kTab + r] = bins[u * 256 + r];
    }
    if (threadIdx.x < 2) ws_dtau[blockIdx.x * 2 + threadIdx.x] = col[64 + threadIdx.x];
  }
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

template <bool kShift>
__global__ void __launch_bounds__(kThreads, 1)
wattn_tc64_bwd_kernel(const __grid_constant__ BwdMaps maps, const float* __restrict__ stats, const float* __restrict__ bias_table,
                      const float* __restrict__ tau, float* __restrict__ ws_dbias, float* __restrict__ ws_dtau,
                      const __grid_constant__ BwdParams p) {
  if (!kShift) {
    wattn_tc64_bwd_body<0>(maps, stats, bias_table, tau, ws_dbias, ws_dtau, p);
  } else {
    const int cls = cta_window_class(p.sched, blockIdx.x);  // a CTA serves one window class (hv_tc_win.cuh)
    if (cls == 0) wattn_tc64_bwd_body<0>(maps, stats, bias_table, tau, ws_dbias, ws_dtau, p);
    else if (cls == 1) wattn_tc64_bwd_body<1>(maps, stats, bias_table, tau, ws_dbias, ws_dtau, p);
    else wattn_tc64_bwd_body<2>(maps, stats, bias_table, tau, ws_dbias, ws_dtau, p);
  }
}

// d(q_bias) = column sums of the q third of dqkv (swinv2.py:193-195, 211-220): a streaming pass over C of every 3C
// channels, 16-byte loads, eight rows per thread in flight, one row of partial sums per CTA (folded, in a fixed order, by
// the reduce kernel below).  HBM-bound: tokens * C * 2 bytes.
constexpr int kCsThreads = 256;
__global__ void __launch_bounds__(kCsThreads) dq_colsum_kernel(const bf16* __restrict__ dqkv, int64_t tokens, int C,
                                                              float* __restrict__ partials) {
  __shared__ float red[kCsThreads * 8];
  const int vecs = C / 8, rpi = kCsThreads / vecs;  // 16-byte vectors per row, rows per block iteration
  const int rl = threadIdx.x / vecs, v = threadIdx.x - rl * vecs;
  const int64_t per = (tokens + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(r0 + per, tokens);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (rl < rpi) {
    const bf16* base = dqkv + v * 8;
    for (int64_t r = r0 + rl; r < r1; r += 8 * rpi) {
      uint4 q[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t rr = r + (int64_t)i * rpi;
        q[i] = rr < r1 ? __ldcs(reinterpret_cast<const uint4*>(base + rr * (3 * C))) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[0] += bf16lo_to_f32(q[i].x); acc[1] += bf16hi_to_f32(q[i].x);
        acc[2] += bf16lo_to_f32(q[i].y); acc[3] += bf16hi_to_f32(q[i].y);
        acc[4] += bf16lo_to_f32(q[i].z); acc[5] += bf16hi_to_f32(q[i].z);
        acc[6] += bf16lo_to_f32(q[i].w); acc[7] += bf16hi_to_f32(q[i].w);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[threadIdx.x * 8 + e] = acc[e];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kCsThreads) {
    const int cv = c >> 3, ce = c & 7;
    float sum = 0.f;
    for (int j = 0; j < rpi; ++j) sum += red[(j * vecs + cv) * 8 + ce];
    partials[(int64_t)blockIdx.x * C + c] = sum;
  }
}

__global__ void __launch_bounds__(256) dq_colsum_fold_kernel(const float* __restrict__ partials, int rows, int C, float* __restrict__ out) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  float s = 0.f;
  for (int b = lane; b < rows; b += 32) s += partials[(int64_t)b * C + c];
  s = warp_sum(s);
  if (lane == 0) out[c] = s;
}

// Sum the per-CTA partials.  Attention CTA c of a same-window group holds heads (2 grp, 2 grp + 1) in units 0 / 1; the
// CTAs of the cross group hold the last (odd) head in both units; the column-sum kernel left cs_rows rows of C partial
// sums.  One warp per output value: the lanes stride over the partial rows (a serial loop over hundreds of rows would
// cost more than the attention kernel's own tail).
__global__ void __launch_bounds__(256) wattn_tc64_bwd_reduce_kernel(const float* __restrict__ ws_dbias, const float* __restrict__ ws_dtau,
                                                                    const float* __restrict__ cs_partials, int cs_rows, BwdParams p,
                                                                    float* __restrict__ dbias_table, float* __restrict__ dtau,
                                                                    float* __restrict__ dq_colsum) {
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int heads = p.g.heads, per_head = kTab + 1, n_tab = heads * per_head;
  if (idx >= n_tab) {
    const int c = idx - n_tab;
    if (c >= p.g.C || dq_colsum == nullptr) return;
    float s = 0.f;
    for (int b = lane; b < cs_rows; b += 32) s += cs_partials[(int64_t)b * p.g.C + c];
    s = warp_sum(s);
    if (lane == 0) dq_colsum[c] = s;
    return;
  }
  const int head = idx / per_head, e = idx - head * per_head;
  const bool is_tau = e == kTab;
  int c0, c1, u0, u1;
  if (head < 2 * p.sched.n_same) {
    c0 = (head >> 1) * p.sched.ctas_same; c1 = c0 + p.sched.ctas_same; u0 = u1 = head & 1;
  } else {
    c0 = p.sched.n_same * p.sched.ctas_same; c1 = c0 + p.sched.ctas_cross; u0 = 0; u1 = 1;
  }
  float s = 0.f;
  for (int c = c0 + lane; c < c1; c += 32)
    for (int u = u0; u <= u1; ++u) s += is_tau ? ws_dtau[c * 2 + u] : ws_dbias[((int64_t)c * 2 + u) * kTab + e];
  s = warp_sum(s);
  if (lane != 0) return;
  if (is_tau) dtau[head] = s;
  else dbias_table[e * heads + head] = s;
}

}  // namespace

static std::atomic<int> g_bwd_variant{-1};  // -1: HV_ATTN_TCGEN05_BWD environment variable (default automatic), 0: mma.sync, 1: tcgen05

int wattn_bwd_variant_set(int v) {
  const int old = g_bwd_variant.exchange(v, std::memory_order_relaxed);
  return old;
}

bool wattn_tc64_bwd_supported(const Geom& g, int dtype) {
  // HV_ATTN_TCGEN05_BWD: unset = automatic (this kernel wherever it is valid), 0 = never (mma.sync backward), 1 = same as unset
  static const int env = []() { const char* e = getenv("HV_ATTN_TCGEN05_BWD"); return e == nullptr ? -1 : (atoi(e) != 0 ? 1 : 0); }();
  const int cur = g_bwd_variant.load(std::memory_order_relaxed);
  const int mode = cur < 0 ? env : cur;
  if (mode == 0) return false;
  // shifted layers: the bias lookup reads runs of four keys, i.e. the column split must sit at 4 (shift = ws / 2, the
  // only shift SwinV2 uses, swinv2.py:560)
  return dtype == HV_BF16 && g.ws == kWs && g.d == 32 && g.C % 32 == 0 && (g.shift == 0 || g.shift == 4) &&
         (int64_t)g.B * g.H * g.W < (int64_t(1) << 31) && g.W * g.C * 2 % 16 == 0;
}

size_t dq_colsum_workspace_bytes(int C) { return (size_t)(4 * num_sms()) * C * sizeof(float) + 256; }

// out[c] = sum over tokens of dqkv[token, c], c < C (dqkv: (tokens, 3C) bf16): the gradient of WindowAttention.q_bias
int dq_colsum(const void* dqkv, int64_t tokens, int C, float* out, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (!aligned16(dqkv) || C % 8 != 0) HV_FAIL(HV_ERR_ALIGN, "dq_colsum: dqkv must be 16-byte aligned and C a multiple of 8");
  if (C / 8 > kCsThreads) HV_FAIL(HV_ERR_SHAPE, "dq_colsum: C = %d > %d", C, 8 * kCsThreads);
  if (workspace == nullptr || workspace_bytes < dq_colsum_workspace_bytes(C))
    HV_FAIL(HV_ERR_WORKSPACE, "dq_colsum: workspace of %zu bytes required", dq_colsum_workspace_bytes(C));
  int cgrid = 4 * num_sms();
  if ((int64_t)cgrid > (tokens + 127) / 128) cgrid = (int)((tokens + 127) / 128);
  float* partials = static_cast<float*>(workspace);
  dq_colsum_kernel<<<cgrid, kCsThreads, 0, st>>>((const bf16*)dqkv, tokens, C, partials);
  HV_LAUNCH_OK("dq_colsum_kernel");
  dq_colsum_fold_kernel<<<(C + 7) / 8, 256, 0, st>>>(partials, cgrid, C, out);
  HV_LAUNCH_OK("dq_colsum_fold_kernel");
  return HV_OK;
}

size_t wattn_tc64_bwd_workspace_bytes(const Geom& g) {
  // one row of partials per attention CTA (grid <= 2 * SMs) | the column-sum kernel's partial rows (4 CTAs per SM)
  return (size_t)(2 * num_sms()) * (2 * kTab + 2) * sizeof(float) + (size_t)(4 * num_sms()) * g.C * sizeof(float) + 256;
}

// `stats`: the three planes (lse | r | c) written by the forward kernels of this geometry (hv_window_attn_stats_floats)
int wattn_tc64_bwd(const Geom& g, const void* qkv, const void* dout, const float* stats, const float* bias_table,
                   const float* tau, void* dqkv, float* dbias_table, float* dtau, float* dq_colsum, void* workspace,
                   size_t workspace_bytes, cudaStream_t st) {
  if (!aligned16(qkv) || !aligned16(dout) || !aligned16(dqkv) || !aligned16(stats))
    HV_FAIL(HV_ERR_ALIGN, "window_attn_bwd: tensors must be 16-byte aligned");
  if (workspace == nullptr || workspace_bytes < wattn_tc64_bwd_workspace_bytes(g))
    HV_FAIL(HV_ERR_WORKSPACE, "window_attn_bwd: workspace of %zu bytes required", wattn_tc64_bwd_workspace_bytes(g));
  struct MapKey { const void *qkv, *dout, *dqkv; int B, H, W, C, shift; };
  struct MapEntry { MapKey key; BwdMaps maps; };
  static thread_local MapEntry cache[32];
  static thread_local int cache_n = 0, cache_next = 0;
  const MapKey key = {qkv, dout, dqkv, g.B, g.H, g.W, g.C, g.shift};
  const BwdMaps* mp = nullptr;
  for (int i = 0; i < cache_n; ++i) {
    const MapKey& c = cache[i].key;
    if (c.qkv == key.qkv && c.dout == key.dout && c.dqkv == key.dqkv && c.B == key.B && c.H == key.H && c.W == key.W &&
        c.C == key.C && c.shift == key.shift) { mp = &cache[i].maps; break; }
  }
  if (!mp) {
    MapEntry& e = cache[cache_next];
    const void* base[3] = {qkv, dout, dqkv};
    const int row_elems[3] = {3 * g.C, g.C, 3 * g.C};
    for (int t = 0; t < 3; ++t) {
      const int rc = make_window_maps(e.maps.m[t], base[t], g, row_elems[t]);
      if (rc) return rc;
    }
    e.key = key;
    mp = &e.maps;
    cache_next = (cache_next + 1) % 32;
    if (cache_n < 32) ++cache_n;
  }
  BwdParams p;
  p.g = g;
  p.plane = (int64_t)g.B * g.nW * g.heads * kN;
  p.ko = 0;
#ifdef HV_TC_TRACE
  if (getenv("HV_TC_KO")) p.ko = atoi(getenv("HV_TC_KO"));
  if (getenv("HV_TC_TRACE_CTA")) p.ko |= atoi(getenv("HV_TC_TRACE_CTA")) << 8;
#endif
  const int nsm = num_sms();
  // CTAs per window class: windows x relative cost (a wrapped window needs twice the TMA boxes and the mask;
  // HV_CLASS_COST="bottom,edge" overrides for tuning)
  static double cost[3] = {1.0, 1.3, 1.4};
  static const bool cost_env = []() {
    const char* e = getenv("HV_CLASS_COST");
    if (e) sscanf(e, "%lf,%lf", &cost[1], &cost[2]);
    return e != nullptr;
  }();
  (void)cost_env;
  const int grid = plan_window_schedule(g, nsm, cost, p.sched);
  float* ws_dbias = static_cast<float*>(workspace);
  float* ws_dtau = ws_dbias + (size_t)grid * 2 * kTab;
  // behind the attention kernel's partials (sized for 2 * SMs CTAs): the column-sum kernel's partial rows
  float* cs_partials = static_cast<float*>(workspace) + (size_t)(2 * nsm) * (2 * kTab + 2);
  static thread_local int attr_dev = -1;
  int dev = 0;
  HV_CUDA_OK(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    HV_CUDA_OK(cudaFuncSetAttribute(wattn_tc64_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    HV_CUDA_OK(cudaFuncSetAttribute(wattn_tc64_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_dev = dev;
  }
#ifdef HV_TC_TRACE
  static long long* dtrace = nullptr;
  if (!dtrace) {
    cudaMalloc(&dtrace, 64 * 16 * sizeof(long long));
    cudaMemcpyToSymbol(g_btrace, &dtrace, sizeof(dtrace));
  }
  cudaMemsetAsync(dtrace, 0, 64 * 16 * sizeof(long long), st);
#endif
  if (g.shift > 0)
    wattn_tc64_bwd_kernel<true><<<grid, kThreads, kSmem, st>>>(*mp, stats, bias_table, tau, ws_dbias, ws_dtau, p);
  else
    wattn_tc64_bwd_kernel<false><<<grid, kThreads, kSmem, st>>>(*mp, stats, bias_table, tau, ws_dbias, ws_dtau, p);
  HV_LAUNCH_OK("wattn_tc64_bwd_kernel");
#ifdef HV_TC_TRACE
  if (getenv("HV_TC_BTRACE_DUMP")) {
    static long long h[64 * 16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, dtrace, sizeof(h), cudaMemcpyDeviceToHost);
    FILE* f = fopen(getenv("HV_TC_BTRACE_DUMP"), "w");
    if (f) {
      for (int k = 0; k < 64; ++k) { for (int e = 0; e < 16; ++e) fprintf(f, "%lld ", h[k * 16 + e] ? h[k * 16 + e] - h[0] : -1LL); fprintf(f, "\n"); }
      fclose(f);
    }
  }
#endif
  int cgrid = 0;
  if (dq_colsum != nullptr) {
    const int64_t tokens = (int64_t)g.B * g.H * g.W;
    cgrid = 4 * nsm;
    if ((int64_t)cgrid > (tokens + 127) / 128) cgrid = (int)((tokens + 127) / 128);
    dq_colsum_kernel<<<cgrid, kCsThreads, 0, st>>>((const bf16*)dqkv, tokens, g.C, cs_partials);
    HV_LAUNCH_OK("dq_colsum_kernel");
  }
  const int n = g.heads * (kTab + 1) + (dq_colsum != nullptr ? g.C : 0);
  wattn_tc64_bwd_reduce_kernel<<<(n + 7) / 8, 256, 0, st>>>(ws_dbias, ws_dtau, cs_partials, cgrid, p, dbias_table, dtau, dq_colsum);
  HV_LAUNCH_OK("wattn_tc64_bwd_reduce_kernel");
  return HV_OK;
}

}  // namespace hv
