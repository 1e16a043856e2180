// tcgen05 / TMEM / TMA backward kernel of the fused shifted-window scaled-cosine attention, window 8x8 (N = 64), head
// dim 32, bf16 (every stage of SwinV2-T).  Same contract as wattn_mma64_bwd (reference swinv2.py:210-263 differentiated;
// qkv / dqkv (B, H*W, 3C), out / dout (B, H*W, C) in IMAGE token order, lse in log2 units, per-CTA partials of
// d(bias table), d(tau) and the dq column sums folded by a second small kernel), different machine mapping:
//
//   * one thread per query row of a (window, head) unit, two units stacked into the 128 TMEM lanes.  The five GEMMs of
//     the backward are tcgen05.mma with fp32 accumulators in tensor memory:
//         S  = [Q_a; Q_b] [K_a; K_b]^T      dP = [dO_a; dO_b] [V_a; V_b]^T        (128 x 128 x 32, diagonal blocks used)
//         dV = P^T dO      dK^ = dS^T Q^      dQ^ = dS K^      dBias += dS I        (128 x 64 x 64)
//     P and dS = P o (dP - D) are written once to shared memory as a [query][key] bf16 tile (128-byte rows, SWIZZLE_128B):
//     read MN-major it is the A operand P^T / dS^T, read K-major it is dS.  Nothing is transposed by threads, no
//     fragment shuffles, and d(bias) accumulates over all windows of the CTA inside the tensor core (64 TMEM columns);
//   * cosine attention: S uses the raw q, k tiles and is scaled by 1/|q_i| * tau/|k_j| in fp32 exactly like the
//     forward kernels (so P matches the forward's lse); afterwards the q / k tiles are normalised IN PLACE (bf16) and
//     serve as the B operands of dK^ / dQ^ and for the projection dq = tau/|q| (M - (q^.M) q^) in the epilogue;
//   * cyclic shift + window partition = coordinates of 4-D TMA tile loads (as in wattn_tc64_fwd).  A shifted layer loads
//     EVERY window as two column parts ([0, 8-s) and [8-s, 8)): one token order for all windows of the launch, so bias
//     lookup, mask and the d(bias) accumulator are uniform and column-wrapped windows need no special case;
//   * the continuous position bias is looked up from a Toeplitz table in shared memory (4 alignment copies of the
//     15 x 15 table per head, 10 KB) instead of an expanded 64 x 64 matrix: 16-byte conflict-free loads, and the
//     shared memory goes to a 4-deep ring of 40 KB stages;
//   * warp roles (24 warps): 0 TMA producer | 1 issuer of S, dP | 2 issuer of dV, dK, dQ, dBias | 4-7 pre-pass (row
//     norms, D = dO.o by tensor-pipe self products, lse; later the in-place normalisation) | 8-15 softmax / dS threads
//     (half a logit row each) | 16-19 dV, dK epilogue | 20-23 dQ epilogue.  All hand-overs are mbarriers.
#include "hv_tc.cuh"

namespace hv {
namespace {
using namespace tc;

constexpr int kN = 64;
constexpr int kWs = 8;
constexpr int kTab = 225;
constexpr int kTile = kN * 64;         // one (window, head) q / k / v / dO / o tile: 64 rows x 64 B (SWIZZLE_64B)
// Two rings: v and o are dead as soon as dP and D exist (early in the life of a pair), q, k and dO live until the epilogue
// has read the normalised rows.  Splitting them lets the long-lived ring be four deep in the same shared memory.
constexpr int kStage = 6 * kTile;      // late ring:  q_a q_b k_a k_b g_a g_b   (g = dO)
constexpr int kStages = 4;
constexpr int kStageE = 4 * kTile;     // early ring: v_a v_b o_a o_b
constexpr int kStagesE = 2;
constexpr int kThreads = 1024;  // 32 warps
constexpr int kPdTile = kN * 128;      // P or dS of one unit: 64 rows x 128 B (SWIZZLE_128B)
constexpr int kBiasRow = 20;           // floats per table row (15 + alignment slack)
constexpr int kBiasCopy = 328;         // floats per alignment copy: >= 15 * 20 and = 8 (mod 32) so 8 lanes hit 8 bank groups

// ---- shared memory map (dynamic, 1024-byte aligned base)
constexpr int kOffStage = 0;
constexpr int kOffStageE = kOffStage + kStages * kStage;
constexpr int kOffP = kOffStageE + kStagesE * kStageE;    // [2 buffers][2 units][64][128 B]
constexpr int kOffDS = kOffP + 4 * kPdTile;
constexpr int kOffEye = kOffDS + 4 * kPdTile;             // 64 x 64 bf16 identity (SWIZZLE_128B)
constexpr int kOffBias = kOffEye + kPdTile;               // [2 units][4 copies][kBiasCopy] float
constexpr int kOffVec = kOffBias + 2 * 4 * kBiasCopy * 4; // [kStages][4: r, c, D, lse][128] float
constexpr int kOffDot = kOffVec + kStages * 4 * 128 * 4;  // [4 pairs in flight][2 halves][128] float: sum_j dS_ij t_ij per half row
constexpr int kOffCol = kOffDot + 4 * 2 * 128 * 4;        // [2][32] float dq column sums, [2] d(tau)
constexpr int kOffBins = kOffP;                           // [2][256] float: d(bias) bins, after the main loop (aliases P)
constexpr int kOffGeo = kOffCol + (2 * 32 + 4) * 4;       // [8][2] UnitGeo
constexpr int kOffSlotMap = kOffGeo + 8 * 2 * 16;         // [64] bytes
constexpr int kOffBar = kOffSlotMap + 64;
constexpr int kNumBars = 5 * kStages + 10 + 2 * kStagesE;
constexpr int kOffTmem = kOffBar + kNumBars * 8;
constexpr int kSmem = kOffTmem + 16;
static_assert(kOffP % 1024 == 0 && kOffBias % 16 == 0 && kOffBar % 8 == 0, "shared-memory alignment");
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
  int n_same, has_cross, ctas_same, ctas_cross;
  int ko;  // HV_TC_TRACE builds only: knock-out bits for bottleneck experiments (results are wrong)
};
// per tensor: [0] full (8, 8) | split order: [1] (wa, 8) [2] (s, 8) [3] (wa, wa) [4] (wa, s) [5] (s, wa) [6] (s, s)
struct BwdMaps { CUtensorMap m[3][7]; };  // qkv, dout, out

struct CtaWork {
  int head_a, head_b, cross, first, stride, npairs;
  __device__ __forceinline__ void init(const BwdParams& p, int cta) {
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
  __device__ __forceinline__ int row(int k, int which, int nrows, bool& valid) const {
    const int idx = first + k * stride;
    valid = true;
    if (!cross) return idx;
    const int r = 2 * idx + which;
    if (r >= nrows) { valid = false; return nrows - 1; }
    return r;
  }
};

struct UnitGeo { int b, row0, col0, rflags; };  // rflags = window row << 3 | right << 2 | bottom << 1 | valid

// window slot (ih, iw) of tile row t: slot order, or the two-column-part order of a shifted layer
__device__ __forceinline__ int tile_row_slot(int t, int shift) {
  if (shift == 0) return t;
  const int wa = kWs - shift;
  int ih, iw;
  if (t < kWs * wa) { ih = t / wa; iw = t - ih * wa; }
  else { const int t2 = t - kWs * wa; ih = t2 / shift; iw = wa + t2 - ih * shift; }
  return ih << 3 | iw;
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 |
// version 1 << 46 | layout type << 61 (2: SWIZZLE_128B, 4: SWIZZLE_64B)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t type) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)type << 61);
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// diagonal of X Y^T for 16 rows x 32 bf16 held as two A fragments each (rows g, g + 8 of the block)
__device__ __forceinline__ void rowdot_mma(const uint32_t (&x)[2][4], const uint32_t (&y)[2][4], float (&n0)[4], float (&n1)[4]) {
#pragma unroll
  for (int e = 0; e < 4; ++e) n0[e] = n1[e] = 0.f;
  mma_bf16(n0, x[0], y[0][0], y[0][2]);
  mma_bf16(n1, x[0], y[0][1], y[0][3]);
  mma_bf16(n0, x[1], y[1][0], y[1][2]);
  mma_bf16(n1, x[1], y[1][1], y[1][3]);
}

#ifdef HV_TC_TRACE
__device__ long long* g_btrace = nullptr;  // [pairs][16 events] clock64 stamps of CTA 0
#define TRACE(k, ev) do { if (blockIdx.x == 0 && lane == 0 && g_btrace && (k) < 64) g_btrace[(k) * 16 + (ev)] = clock64(); } while (0)
#define KO(bit) (p.ko & (bit))
#else
#define TRACE(k, ev) do { } while (0)
#define KO(bit) false
#endif

template <bool V> struct BoolTag { static constexpr bool value = V; };
template <int V> struct IntTag { static constexpr int value = V; };

template <bool kSplit>
__global__ void __launch_bounds__(kThreads, 1)
wattn_tc64_bwd_kernel(const __grid_constant__ BwdMaps maps, const float* __restrict__ lse, const float* __restrict__ bias_table,
                      const float* __restrict__ tau, bf16* __restrict__ dqkv, float* __restrict__ ws_dbias,
                      float* __restrict__ ws_dtau, float* __restrict__ ws_colsum, int want_colsum, BwdParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const Geom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar0 = sb + kOffBar;
  auto bar_full = [&](int s) { return bar0 + 8 * s; };
  auto bar_empty = [&](int s) { return bar0 + 8 * (kStages + s); };
  auto bar_pre = [&](int s) { return bar0 + 8 * (2 * kStages + s); };
  auto bar_hat = [&](int s) { return bar0 + 8 * (3 * kStages + s); };
  auto bar_sdp = [&](int s) { return bar0 + 8 * (4 * kStages + s); };
  // accumulator buffer a (= pair parity): output MMAs complete / epilogue has pulled the accumulators out of TMEM
  auto bar_acc = [&](int a) { return bar0 + 8 * (a ? 5 * kStages + 8 + 2 * kStagesE : 5 * kStages + 1); };
  auto bar_accfree = [&](int a) { return bar0 + 8 * (a ? 5 * kStages + 9 + 2 * kStagesE : 5 * kStages + 2); };
  auto bar_sfree = [&](int b) { return bar0 + 8 * (5 * kStages + (b ? 0 : 7)); };  // S / dP buffer b read by its softmax group
  auto bar_fullE = [&](int s) { return bar0 + 8 * (5 * kStages + 8 + s); };
  auto bar_emptyE = [&](int s) { return bar0 + 8 * (5 * kStages + 8 + kStagesE + s); };
  auto bar_staged = [&](int b) { return bar0 + 8 * (5 * kStages + 3 + b); };  // P / dS staging buffer b written
  auto bar_stfree = [&](int b) { return bar0 + 8 * (5 * kStages + 5 + b); };  // ... and read by the MMAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmem);
  const int nrows = g.B * g.nW;

  __shared__ CtaWork s_work;
  if (threadIdx.x == 0) {
    CtaWork w0;
    w0.init(p, blockIdx.x);
    s_work = w0;
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 8);  // the eight epilogue warps
      mbar_init(bar_pre(s), 4);
      mbar_init(bar_hat(s), 8);
      mbar_init(bar_sdp(s), 1);
    }
    for (int s = 0; s < kStagesE; ++s) {
      mbar_init(bar_fullE(s), 1);
      mbar_init(bar_emptyE(s), 5);  // the four pre-pass warps (o, for D) and the commit behind the dP MMAs (v)
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_staged(b), 8);
      mbar_init(bar_stfree(b), 1);
      mbar_init(bar_sfree(b), 8);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc(a), 1);
      mbar_init(bar_accfree(a), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // ---- one-time tables: identity tile, slot of every tile row, Toeplitz bias copies (log2 units), zeroed reduction bins
  unsigned char* slotmap = smem + kOffSlotMap;
  for (int idx = threadIdx.x; idx < kPdTile / 4; idx += kThreads) {
    // 32-bit word idx of the swizzled identity: row = byte / 128, physical chunk = (byte / 16) & 7, logical chunk = phys ^ (row & 7)
    const int byte = idx * 4, row = byte >> 7, chunk = ((byte >> 4) & 7) ^ (row & 7);
    const int col = chunk * 8 + ((byte & 15) >> 1);  // first of the two bf16 columns of this word
    uint32_t w = 0;
    if (col == row) w = 0x00003F80u;
    else if (col + 1 == row) w = 0x3F800000u;
    reinterpret_cast<uint32_t*>(smem + kOffEye)[idx] = w;
  }
  if (threadIdx.x < 64) slotmap[threadIdx.x] = (unsigned char)tile_row_slot(threadIdx.x, g.shift);
  for (int idx = threadIdx.x; idx < 2 * 32 + 4; idx += kThreads) reinterpret_cast<float*>(smem + kOffCol)[idx] = 0.f;
  {
    CtaWork w0;
    w0.init(p, blockIdx.x);
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
      // ---------------------------------------------------------------- TMA producer: lane t < 8 loads tile t of the stage
      for (int k = 0; k < npairs; ++k) {
        const int s = k % kStages, se = k % kStagesE;
        mbar_wait_fast(bar_empty(s), ((k / kStages) & 1) ^ 1);
        mbar_wait_fast(bar_emptyE(se), ((k / kStagesE) & 1) ^ 1);
        TRACE(k, 0);
        const int which = lane & 1, kind = lane >> 1;  // kind: 0 q, 1 k, 2 v, 3 dO, 4 o
        bool valid;
        const int r = work.row(k, which, nrows, valid);
        const int b = r / g.nW, win = r - b * g.nW;
        const int wh = win / g.nWw, ww = win - wh * g.nWw;
        const int row0 = wh * kWs + g.shift, col0 = ww * kWs + g.shift;
        const bool bottom = g.shift > 0 && wh == g.H / kWs - 1, right = g.shift > 0 && ww == g.nWw - 1;
        if (lane < 2) {
          UnitGeo ug;
          ug.b = b; ug.row0 = row0; ug.col0 = col0;
          ug.rflags = (r << 3) | (right ? 4 : 0) | (bottom ? 2 : 0) | (valid ? 1 : 0);
          geo[(k & 7) * 2 + which] = ug;
        }
        __syncwarp();
        if (lane == 0) {
          mbar_expect_tx(bar_full(s), KO(1) ? 0 : kStage);
          mbar_expect_tx(bar_fullE(se), KO(1) ? 0 : kStageE);
        }
        __syncwarp();
        if (lane < 10 && !KO(1)) {
          const int head = which == 0 ? work.head_a : work.head_b;
          const int tsr = kind < 3 ? 0 : (kind == 3 ? 1 : 2);
          const int c0 = (kind < 3 ? kind * g.C : 0) + head * 32;
          const bool early = kind == 2 || kind == 4;  // v, o
          const int tidx = (kind == 0 ? 0 : (kind == 1 ? 2 : (kind == 3 ? 4 : (kind == 2 ? 0 : 2)))) + which;
          const uint32_t dst = early ? sb + kOffStageE + se * kStageE + tidx * kTile : sb + kOffStage + s * kStage + tidx * kTile;
          const uint32_t bar = early ? bar_fullE(se) : bar_full(s);
          const CUtensorMap* mm = maps.m[tsr];
          if (!kSplit) {
            tma_load_4d(dst, &mm[0], bar, c0, col0, row0, b);
          } else {
            const int sh = g.shift, wa = kWs - g.shift;
            int colb = col0 + wa;
            if (colb >= g.W) colb -= g.W;
            const uint32_t dstb = dst + kWs * wa * 64;
            if (!bottom) {
              tma_load_4d(dst, &mm[1], bar, c0, col0, row0, b);
              tma_load_4d(dstb, &mm[2], bar, c0, colb, row0, b);
            } else {
              tma_load_4d(dst, &mm[3], bar, c0, col0, row0, b);
              tma_load_4d(dst + wa * wa * 64, &mm[4], bar, c0, col0, 0, b);
              tma_load_4d(dstb, &mm[5], bar, c0, colb, row0, b);
              tma_load_4d(dstb + sh * wa * 64, &mm[6], bar, c0, colb, 0, b);
            }
          }
        }
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
      const uint32_t id_b = idesc_bf16(128, 64, 0, 0);   // A = dS of both units (K-major), B = identity
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
          if (!KO(8)) {
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
              umma_ss(tmem + kColDB, a_ds + bo + (uint64_t)(2 * ks), b_eye + (uint64_t)(2 * ks), id_b, (k > 0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(bar_acc(buf));
          umma_commit(bar_stfree(buf));
          TRACE(k, 12);
        }
        __syncwarp();
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ pre-pass warps
    const int w4 = warp - 4;
    const int u = w4 & 1, part = w4 >> 1;  // norms: tile (part: q | k, unit u); D: unit u, rows 32 * part ..
    const int head_u = u == 0 ? work.head_a : work.head_b;
    const float tau_u = __ldg(&tau[head_u]);
    const float mult = part == 0 ? 1.0f : tau_u * kLog2e;
    const int g_ = lane >> 2, t_ = lane & 3;
    const int arow = (lane & 7) + 8 * ((lane >> 3) & 1), achunk = lane >> 4;
    const bool odd = (lane >> 2) & 1;
    const int src = (lane & ~3) | (lane >> 3);
    // lse gather: this warp fetches pair rows 32 * w4 + lane
    const int lrow = 32 * w4 + lane, lu = lrow >> 6, lslot = slotmap[lrow & 63];
    const int lhead = lu == 0 ? work.head_a : work.head_b;

    auto pre = [&](int k) {
      const int s = k % kStages, se = k % kStagesE;
      mbar_wait_fast(bar_full(s), (k / kStages) & 1);
      mbar_wait_fast(bar_fullE(se), (k / kStagesE) & 1);
      const uint32_t st = sb + kOffStage + s * kStage;
      float* vec = vecs + s * 4 * 128;
      // lse of the pair's rows (1e30 for the padding unit of an odd tail: P = dS = 0 there); the load is issued first and
      // consumed at the end of the pre-pass so that its latency hides behind the tensor-pipe work
      const int rf = geo[(k & 7) * 2 + lu].rflags;
      const float lse_v = (rf & 1) ? __ldg(&lse[((int64_t)(rf >> 3) * g.heads + lhead) * kN + lslot]) : 1e30f;
      const uint32_t tile = st + (2 * part + u) * kTile;
      const uint32_t gt = st + (4 + u) * kTile, ot = sb + kOffStageE + se * kStageE + (2 + u) * kTile;
#pragma unroll
      for (int bp = 0; bp < 2; ++bp) {
        uint32_t x[2][2][4];
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) {
          const int row = 16 * (2 * bp + b2) + arow;
          ldsm_x4(tile + row * 64 + (((achunk) ^ ((row >> 1) & 3)) << 4), x[b2][0]);
          ldsm_x4(tile + row * 64 + (((2 + achunk) ^ ((row >> 1) & 3)) << 4), x[b2][1]);
        }
        float n0[2][4], n1[2][4];
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) rowdot_mma(x[b2], x[b2], n0[b2], n1[b2]);
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) {
          const float s0 = __shfl_sync(0xffffffffu, odd ? n0[b2][1] : n0[b2][0], src);
          const float s1 = __shfl_sync(0xffffffffu, odd ? n1[b2][3] : n1[b2][2], src);
          if (t_ == 0) {
            vec[part * 128 + 64 * u + 16 * (2 * bp + b2) + g_] = mult * inv_norm(s0);
            vec[part * 128 + 64 * u + 16 * (2 * bp + b2) + g_ + 8] = mult * inv_norm(s1);
          }
        }
      }
      {  // D = rowsum(dO o o) for rows 32 * part .. + 32 of unit u
        uint32_t x[2][2][4], y[2][2][4];
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) {
          const int row = 32 * part + 16 * b2 + arow;
          const uint32_t o0 = row * 64 + (((achunk) ^ ((row >> 1) & 3)) << 4), o1 = row * 64 + (((2 + achunk) ^ ((row >> 1) & 3)) << 4);
          ldsm_x4(gt + o0, x[b2][0]);
          ldsm_x4(gt + o1, x[b2][1]);
          ldsm_x4(ot + o0, y[b2][0]);
          ldsm_x4(ot + o1, y[b2][1]);
        }
        float n0[2][4], n1[2][4];
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) rowdot_mma(x[b2], y[b2], n0[b2], n1[b2]);
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) {
          const float s0 = __shfl_sync(0xffffffffu, odd ? n0[b2][1] : n0[b2][0], src);
          const float s1 = __shfl_sync(0xffffffffu, odd ? n1[b2][3] : n1[b2][2], src);
          if (t_ == 0) {
            vec[2 * 128 + 64 * u + 32 * part + 16 * b2 + g_] = s0;
            vec[2 * 128 + 64 * u + 32 * part + 16 * b2 + g_ + 8] = s1;
          }
        }
      }
      vec[3 * 128 + lrow] = lse_v;
      __syncwarp();
      if (warp == 4) TRACE(k, 3);
      if (lane == 0) {
        mbar_arrive(bar_pre(s));
        mbar_arrive(bar_emptyE(se));  // o is dead once D exists
      }
    };
    for (int k = 0; k < npairs; ++k) pre(k);

  } else if (warp < 24) {
    // ------------------------------------------------------------------ softmax / dS threads: two groups (warps 8-15 even
    // pairs, 16-23 odd pairs) so that the hand-over latencies of one group hide behind the arithmetic of the other; a thread
    // owns half a logit row and streams it from TMEM 16 keys at a time (64 registers per thread: S / dP stay in TMEM, which
    // has room for one buffer per group).  Lanes 0-15 of a warp are rows of unit a, lanes 16-31 of unit b (M = 64 layout).
    const int grp = (warp - 8) >> 3;
    const int half = ((warp - 8) >> 2) & 1;
    const int quad = warp & 3;
    const int u = lane >> 4, i = 16 * quad + (lane & 15);  // unit of the pair, tile row (query) inside the unit
    const int row = 64 * u + i;                            // row of the pair in the per-stage vectors
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16);
    const float kNeg = kMaskValue * kLog2e;
    const int si = slotmap[i], ih = si >> 3, iw = si & 7;
    // Toeplitz bias: float index of (dh = ih + 7, x = 7 - iw) in the alignment copy that makes x a multiple of 4
    const int cpy = (7 - iw) & 3;
    const float* bias_base = reinterpret_cast<const float*>(smem + kOffBias) + u * 4 * kBiasCopy + cpy * kBiasCopy +
                             (ih + 7) * kBiasRow + (7 - iw - cpy) + 4 + (kSplit ? 4 * half : -(4 * half) * kBiasRow);
    // masks of a shifted layer: bit j set = key j of this thread's half sits on the other side of the wrap than the query
    uint32_t mH = 0u, mW = 0u;
    if (kSplit) {
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
    const float inv_mult = hpart == 0 ? 1.0f : 1.0f / (__ldg(&tau[hu == 0 ? work.head_a : work.head_b]) * kLog2e);
    auto hat = [&](int k) {
      const int s = k % kStages;
      const uint32_t tile = sb + kOffStage + s * kStage + (2 * hpart + hu) * kTile;
      const float* vec = vecs + s * 4 * 128 + hpart * 128 + 64 * hu;
      if (!KO(16)) {
        uint4 v[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) v[ch] = lds128(tile + hrow * 64 + ((ch ^ ((hrow >> 1) & 3)) << 4));
        const float sc = vec[hrow] * inv_mult;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint4 w = v[ch];
          w.x = pack_bf16x2(bf16lo_to_f32(w.x) * sc, bf16hi_to_f32(w.x) * sc);
          w.y = pack_bf16x2(bf16lo_to_f32(w.y) * sc, bf16hi_to_f32(w.y) * sc);
          w.z = pack_bf16x2(bf16lo_to_f32(w.z) * sc, bf16hi_to_f32(w.z) * sc);
          w.w = pack_bf16x2(bf16lo_to_f32(w.w) * sc, bf16hi_to_f32(w.w) * sc);
          sts128(tile + hrow * 64 + ((ch ^ ((hrow >> 1) & 3)) << 4), w);
        }
      }
    };

    for (int k = grp; k < npairs; k += 2) {
      const int s = k % kStages;
      const uint32_t ph = (k / kStages) & 1;
      mbar_wait_fast(bar_pre(s), ph);
      if (warp == 8) TRACE(k, 5);
      const int rflags = geo[(k & 7) * 2 + u].rflags;
      const float* vec = vecs + s * 4 * 128;
      const float ri = vec[row], Di = vec[2 * 128 + row], li = vec[3 * 128 + row];
      const float* cv = vec + 128 + 64 * u + 32 * half;
      uint32_t m = 0u;
      if (kSplit) m = ((rflags & 2) ? mH : 0u) | ((rflags & 4) ? mW : 0u);
      const bool any_mask = kSplit && __any_sync(0xffffffffu, m != 0u);
      if (k > 1) mbar_wait_fast(bar_stfree(grp), ((k >> 1) - 1) & 1);  // the MMAs of pair k-2 have read this group's staging tiles
      mbar_wait_fast(bar_sdp(s), ph);
      if (warp == 8) TRACE(k, 6);
      tc_fence_after();
      float racc = 0.f;  // sum_j dS_ij t_ij over this half row: d(tau) contribution and the dQ epilogue's q^.M
      auto chunk = [&](auto masked, auto ck_tag) {
        constexpr int ck = decltype(ck_tag)::value;  // keys 16 ck .. 16 ck + 15 of this half
        uint32_t sa[16], pa[16];
        HV_TMEM_LD16(tS + 16 * ck, sa);
        HV_TMEM_LD16(tP + 16 * ck, pa);
        tmem_wait_ld();
        if (ck == 1) {  // the whole half row has left TMEM: hand the buffer back to the S / dP issuer
          tc_fence_before();
          __syncwarp();
          if (warp == 8) TRACE(k, 7);
          if (lane == 0) mbar_arrive(bar_sfree(grp));
        }
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {  // eight keys = one 16-byte staging chunk of P and of dS
          uint32_t pp[4], dd[4];
#pragma unroll
          for (int q2 = 0; q2 < 2; ++q2) {
            const int q = 4 * ck + 2 * h8 + q2;
            // keys 4q .. 4q + 3 of this half: slot order = window row 4 half + q / 2, columns 4 (q & 1) ..;
            // split order = window row q, columns 4 half ..
            const float4 b = *reinterpret_cast<const float4*>(bias_base + (kSplit ? -q * kBiasRow : -(q >> 1) * kBiasRow + 4 * (q & 1)));
            const float4 c = *reinterpret_cast<const float4*>(cv + 4 * q);
            const float bb[4] = {b.x, b.y, b.z, b.w}, cc[4] = {c.x, c.y, c.z, c.w};
            float pv[4], dv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = 4 * q + e, jj = 8 * h8 + 4 * q2 + e;
              const float t = (__uint_as_float(sa[jj]) * ri) * cc[e];  // tau log2e cos(q_i, k_j)
              float x = (t + bb[e]) - li;
              if (decltype(masked)::value && ((m >> j) & 1u)) x += kNeg;
              const float pe = ex2(x);
              const float de = pe * (__uint_as_float(pa[jj]) - Di);
              racc = fmaf(de, t, racc);
              pv[e] = pe;
              dv[e] = de;
            }
            pp[2 * q2] = pack_bf16x2(pv[0], pv[1]);
            pp[2 * q2 + 1] = pack_bf16x2(pv[2], pv[3]);
            dd[2 * q2] = pack_bf16x2(dv[0], dv[1]);
            dd[2 * q2 + 1] = pack_bf16x2(dv[2], dv[3]);
          }
          const uint32_t off = (uint32_t)(((4 * half + 2 * ck + h8) ^ (i & 7)) << 4);
          sts128(p_row + off, make_uint4(pp[0], pp[1], pp[2], pp[3]));
          sts128(ds_row + off, make_uint4(dd[0], dd[1], dd[2], dd[3]));
        }
      };
      if (any_mask) { chunk(BoolTag<true>{}, IntTag<0>{}); chunk(BoolTag<true>{}, IntTag<1>{}); }
      else { chunk(BoolTag<false>{}, IntTag<0>{}); chunk(BoolTag<false>{}, IntTag<1>{}); }
      if (warp == 8) TRACE(k, 8);
      acc_tau += racc;
      dots[((k & 3) * 2 + half) * 128 + row] = racc;
      hat(k);
      fence_async_smem();
      __syncwarp();
      if (warp == 8) TRACE(k, 9);
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
    // ------------------------------------------------------------------ epilogue warps: 24-27 dV and dK, 28-31 dQ
    auto epilogue = [&](auto role_tag) {
    constexpr int role = decltype(role_tag)::value;  // 1 dV + dK, 2 dQ
    if (role == 2) reg_alloc<88>();
    const int quad = warp & 3;
    const int u = lane >> 4, t = 16 * quad + (lane & 15);  // M = 64 accumulator layout: lanes 0-15 unit a, 16-31 unit b
    const int row = 64 * u + t;
    const int head = u == 0 ? work.head_a : work.head_b;
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16);
    const int sl = slotmap[t], ih = sl >> 3, iw = sl & 7;
    const float tau_h = __ldg(&tau[head]);
    const float inv_tl = 1.0f / (tau_h * kLog2e);
    float csum[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) csum[e] = 0.f;

    for (int k = 0; k < npairs; ++k) {
      const int s = k % kStages;
      const int ab = k & 1;
      mbar_wait_fast(bar_acc(ab), (k >> 1) & 1);
      if (warp == 24) TRACE(k, 13);
      tc_fence_after();
      uint32_t a[32];
      const UnitGeo ug = geo[(k & 7) * 2 + u];
      int prow = ug.row0 + ih; if (prow >= g.H) prow -= g.H;
      int pcol = ug.col0 + iw; if (pcol >= g.W) pcol -= g.W;
      const int64_t tok = ((int64_t)ug.b * g.H + prow) * g.W + pcol;
      bf16* drow = dqkv + tok * (3 * g.C) + head * 32 + (role == 1 ? g.C : 0);
      const bool valid = (ug.rflags & 1) && !KO(2);
      if (role == 1) {  // dV first: pack and store, then the same registers take dK
        HV_TMEM_LD32(tl + kAccCols * ab + kColDV, a);
        tmem_wait_ld();
        HV_REG_FENCE32(a);
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(drow + g.C);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(a[8 * q + 0]), __uint_as_float(a[8 * q + 1]));
            v.y = pack_bf16x2(__uint_as_float(a[8 * q + 2]), __uint_as_float(a[8 * q + 3]));
            v.z = pack_bf16x2(__uint_as_float(a[8 * q + 4]), __uint_as_float(a[8 * q + 5]));
            v.w = pack_bf16x2(__uint_as_float(a[8 * q + 6]), __uint_as_float(a[8 * q + 7]));
            dst[q] = v;
          }
        }
      }
      HV_TMEM_LD32(tl + kAccCols * ab + (role == 1 ? kColDK : kColDQ), a);
      float qdot = 0.f;
      if (role == 2) {
        const float* dp = reinterpret_cast<const float*>(smem + kOffDot) + (k & 3) * 256 + row;
        qdot = (dp[0] + dp[128]) * inv_tl;
      }
      tmem_wait_ld();
      HV_REG_FENCE32(a);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_accfree(ab));
      // projection of the gradient of the normalised row back to the raw row: d x = sc (M - (x^ . M) x^)
      const float* vec = vecs + s * 4 * 128;
      const uint32_t tile = sb + kOffStage + s * kStage + ((role == 2 ? 0 : 2) + u) * kTile + t * 64;
      const float sc = role == 2 ? vec[row] * tau_h : vec[128 + row] * kLn2;  // tau / |q_i|  |  tau / |k_j| = c_j ln 2
      uint4* dst = reinterpret_cast<uint4*>(drow);
      if (role == 1) {
        uint32_t xh[16];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const uint4 v = lds128(tile + ((ch ^ ((t >> 1) & 3)) << 4));
          xh[4 * ch] = v.x; xh[4 * ch + 1] = v.y; xh[4 * ch + 2] = v.z; xh[4 * ch + 3] = v.w;
        }
        __syncwarp();
        if (warp == 24) TRACE(k, 15);
        if (lane == 0) mbar_arrive(bar_empty(s));  // last read of the stage by this warp
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
          if (valid) dst[ch] = make_uint4(o[0], o[1], o[2], o[3]);
        }
      } else {
        // q^_i . M_i = sum_j dS_ij cos_ij: the softmax threads already have it (two half-row sums of dS t, t = tau log2e cos)
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const uint4 v = lds128(tile + ((ch ^ ((t >> 1) & 3)) << 4));
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v0 = sc * fmaf(-qdot, bf16lo_to_f32(w[e]), __uint_as_float(a[8 * ch + 2 * e]));
            const float v1 = sc * fmaf(-qdot, bf16hi_to_f32(w[e]), __uint_as_float(a[8 * ch + 2 * e + 1]));
            if (valid) { csum[8 * ch + 2 * e] += v0; csum[8 * ch + 2 * e + 1] += v1; }
            o[e] = pack_bf16x2(v0, v1);
          }
          if (valid) dst[ch] = make_uint4(o[0], o[1], o[2], o[3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty(s));  // last read of the stage by this warp
      }
    }
    if (role == 2 && want_colsum) {
      float* col = reinterpret_cast<float*>(smem + kOffCol) + u * 32;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float v = group_sum<16>(csum[e]);  // lanes 0-15 / 16-31 are rows of unit a / b
        if ((lane & 15) == 0) atomicAdd(&col[e], v);
      }
    }
    };
    if (warp < 28) epilogue(IntTag<1>{}); else epilogue(IntTag<2>{});
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
    if (threadIdx.x < 64) ws_colsum[blockIdx.x * 64 + threadIdx.x] = col[threadIdx.x];
  }
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// Sum the per-CTA partials.  CTA c of a same-window group holds heads (2 grp, 2 grp + 1) in units 0 / 1; the CTAs of the
// cross group hold the last (odd) head in both units.  One warp per output value: the lanes stride over the CTAs (a serial
// loop over ~100 partial rows would cost more than the attention kernel's own tail).
__global__ void __launch_bounds__(256) wattn_tc64_bwd_reduce_kernel(const float* __restrict__ ws_dbias, const float* __restrict__ ws_dtau,
                                                                    const float* __restrict__ ws_colsum, BwdParams p,
                                                                    float* __restrict__ dbias_table, float* __restrict__ dtau,
                                                                    float* __restrict__ dq_colsum) {
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int heads = p.g.heads, per_head = kTab + 1, n_tab = heads * per_head;
  int head, what, e = 0;  // what: 0 bias bin, 1 tau, 2 column sum
  if (idx < n_tab) {
    head = idx / per_head;
    e = idx - head * per_head;
    what = e < kTab ? 0 : 1;
  } else if (idx < n_tab + p.g.C && dq_colsum != nullptr) {
    head = (idx - n_tab) / 32;
    e = (idx - n_tab) & 31;
    what = 2;
  } else {
    return;
  }
  int c0, c1, u0, u1;
  if (head < 2 * p.n_same) {
    c0 = (head >> 1) * p.ctas_same; c1 = c0 + p.ctas_same; u0 = u1 = head & 1;
  } else {
    c0 = p.n_same * p.ctas_same; c1 = c0 + p.ctas_cross; u0 = 0; u1 = 1;
  }
  float s = 0.f;
  for (int c = c0 + lane; c < c1; c += 32)
    for (int u = u0; u <= u1; ++u)
      s += what == 0 ? ws_dbias[((int64_t)c * 2 + u) * kTab + e] : (what == 1 ? ws_dtau[c * 2 + u] : ws_colsum[c * 64 + u * 32 + e]);
  s = warp_sum(s);
  if (lane != 0) return;
  if (what == 0) dbias_table[e * heads + head] = s;
  else if (what == 1) dtau[head] = s;
  else dq_colsum[idx - n_tab] = s;
}

}  // namespace

static int g_bwd_variant = -1;  // -1: HV_ATTN_TCGEN05_BWD environment variable (default automatic), 0: mma.sync, 1: tcgen05

int wattn_bwd_variant_set(int v) {
  const int old = g_bwd_variant;
  g_bwd_variant = v;
  return old;
}

bool wattn_tc64_bwd_supported(const Geom& g, int dtype) {
  static const int env = []() { const char* e = getenv("HV_ATTN_TCGEN05_BWD"); return e == nullptr ? -1 : (atoi(e) != 0 ? 1 : 0); }();
  const int mode = g_bwd_variant < 0 ? env : g_bwd_variant;
  if (mode == 0) return false;
  // shifted layers: the bias lookup reads runs of four keys, i.e. the column split must sit at 4 (shift = ws / 2, the
  // only shift SwinV2 uses, swinv2.py:560)
  const bool valid = dtype == HV_BF16 && g.ws == kWs && g.d == 32 && g.C % 32 == 0 && (g.shift == 0 || g.shift == 4) &&
                     (int64_t)g.B * g.H * g.W < (int64_t(1) << 31) && g.W * g.C * 2 % 16 == 0;
  if (!valid || mode == 1) return valid;
  return false;  // automatic: the mma.sync backward until this kernel is measured to be the faster one
}

size_t wattn_tc64_bwd_workspace_bytes(const Geom& g) {
  (void)g;
  return (size_t)num_sms() * (2 * kTab + 2 + 64) * sizeof(float) + 256;
}

int wattn_tc64_bwd(const Geom& g, const void* qkv, const void* out, const void* dout, const float* lse,
                   const float* bias_table, const float* tau, void* dqkv, float* dbias_table, float* dtau,
                   float* dq_colsum, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (!aligned16(qkv) || !aligned16(out) || !aligned16(dout) || !aligned16(dqkv) || !aligned16(lse))
    HV_FAIL(HV_ERR_ALIGN, "window_attn_bwd: tensors must be 16-byte aligned");
  if (workspace == nullptr || workspace_bytes < wattn_tc64_bwd_workspace_bytes(g))
    HV_FAIL(HV_ERR_WORKSPACE, "window_attn_bwd: workspace of %zu bytes required", wattn_tc64_bwd_workspace_bytes(g));
  struct MapKey { const void *qkv, *out, *dout; int B, H, W, C, shift; };
  struct MapEntry { MapKey key; BwdMaps maps; };
  static thread_local MapEntry cache[32];
  static thread_local int cache_n = 0, cache_next = 0;
  const MapKey key = {qkv, out, dout, g.B, g.H, g.W, g.C, g.shift};
  const BwdMaps* mp = nullptr;
  for (int i = 0; i < cache_n; ++i) {
    const MapKey& c = cache[i].key;
    if (c.qkv == key.qkv && c.out == key.out && c.dout == key.dout && c.B == key.B && c.H == key.H && c.W == key.W &&
        c.C == key.C && c.shift == key.shift) { mp = &cache[i].maps; break; }
  }
  if (!mp) {
    MapEntry& e = cache[cache_next];
    const int s = g.shift, wa = kWs - g.shift;
    const int bw[7] = {kWs, s ? wa : kWs, s ? s : kWs, s ? wa : kWs, s ? wa : kWs, s ? s : kWs, s ? s : kWs};
    const int bh[7] = {kWs, kWs, kWs, s ? wa : kWs, s ? s : kWs, s ? wa : kWs, s ? s : kWs};
    const void* base[3] = {qkv, dout, out};
    const int row_elems[3] = {3 * g.C, g.C, g.C};
    for (int t = 0; t < 3; ++t)
      for (int i = 0; i < 7; ++i) {
        const int rc = make_map(&e.maps.m[t][i], base[t], g, row_elems[t], bw[i], bh[i]);
        if (rc) return rc;
      }
    e.key = key;
    mp = &e.maps;
    cache_next = (cache_next + 1) % 32;
    if (cache_n < 32) ++cache_n;
  }
  BwdParams p;
  p.g = g;
  p.n_same = g.heads / 2;
  p.has_cross = g.heads & 1;
  p.ko = 0;
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
  float* ws_dbias = static_cast<float*>(workspace);
  float* ws_dtau = ws_dbias + (size_t)grid * 2 * kTab;
  float* ws_colsum = ws_dtau + (size_t)grid * 2;
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
    wattn_tc64_bwd_kernel<true><<<grid, kThreads, kSmem, st>>>(*mp, lse, bias_table, tau, (bf16*)dqkv, ws_dbias, ws_dtau, ws_colsum,
                                                               dq_colsum != nullptr, p);
  else
    wattn_tc64_bwd_kernel<false><<<grid, kThreads, kSmem, st>>>(*mp, lse, bias_table, tau, (bf16*)dqkv, ws_dbias, ws_dtau, ws_colsum,
                                                                dq_colsum != nullptr, p);
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
  const int n = g.heads * (kTab + 1) + g.C;
  wattn_tc64_bwd_reduce_kernel<<<(n + 7) / 8, 256, 0, st>>>(ws_dbias, ws_dtau, ws_colsum, p, dbias_table, dtau, dq_colsum);
  HV_LAUNCH_OK("wattn_tc64_bwd_reduce_kernel");
  return HV_OK;
}

}  // namespace hv
