// Window scheduling shared by the tcgen05 / TMEM / TMA attention kernels (8x8 windows, head dim 32, bf16): which
// (window, head) units a CTA walks, in which token order a window sits in its shared-memory tile, and the TMA boxes that
// move a tile between the image-ordered tensors and that tile (the cyclic shift + window partition of reference
// swinv2.py:399-412 / the reverse + un-roll of :420-429 are these coordinates).
#pragma once
#include <stdlib.h>

#include "hv_tc.cuh"

namespace hv {
namespace tc {

constexpr int kWinSide = 8;

// Head-pair groups (both units of a pair = heads 2g, 2g+1 of one window) and the cross group (the odd last head of two
// consecutive windows).  The ctas_same / ctas_cross CTAs of a group are split between the window classes of a shifted
// layer (cls_same / cls_cross CTAs per class, in class order; an unshifted layer has class 0 only).
struct WinSchedule {
  int n_same, has_cross, ctas_same, ctas_cross, cls_same[3], cls_cross[3];
};

// Window classes of a shifted layer: 0 = interior windows (no wrap, no mask: exactly the unshifted code path), 1 = bottom
// row of windows left of the last column (rows wrap: slot order, two row boxes, mask along h), 2 = right-edge windows
// (columns wrap: two column parts, permuted token order, mask along w and, in the corner, h).  A CTA serves one class,
// so that its tile order -- and with it bias lookup, mask and any per-CTA accumulator -- is uniform, and 49 of 64
// windows of a stage-0 layer run the cheap class-0 code.
__device__ __forceinline__ int cta_window_class(const WinSchedule& sc, int cta) {
  const int same_total = sc.n_same * sc.ctas_same;
  const int pos = cta < same_total ? cta % sc.ctas_same : cta - same_total;
  const int* cc = cta < same_total ? sc.cls_same : sc.cls_cross;
  return pos < cc[0] ? 0 : (pos < cc[0] + cc[1] ? 1 : 2);
}

struct CtaWork {
  int head_a, head_b, cross, cls, first, stride, npairs, ncls, wcls, hcls;
  __device__ __forceinline__ void init(const Geom& g, const WinSchedule& sc, int cta) {
    const int same_total = sc.n_same * sc.ctas_same;
    int pos;
    const int* cc;
    if (cta < same_total) {
      const int grp = cta / sc.ctas_same;
      cross = 0; head_a = 2 * grp; head_b = 2 * grp + 1;
      pos = cta - grp * sc.ctas_same; cc = sc.cls_same;
    } else {
      cross = 1; head_a = head_b = g.heads - 1;
      pos = cta - same_total; cc = sc.cls_cross;
    }
    cls = pos < cc[0] ? 0 : (pos < cc[0] + cc[1] ? 1 : 2);
    first = pos - (cls > 0 ? cc[0] : 0) - (cls > 1 ? cc[1] : 0);
    stride = cc[cls];
    const int nWh = g.H / kWinSide, nWw = g.nWw;
    if (g.shift == 0) { wcls = nWw; hcls = nWh; }
    else { wcls = cls == 2 ? 1 : nWw - 1; hcls = cls == 0 ? nWh - 1 : (cls == 1 ? 1 : nWh); }
    ncls = g.B * hcls * wcls;
    const int units = cross ? (ncls + 1) / 2 : ncls;
    npairs = first < units ? (units - first + stride - 1) / stride : 0;
  }
};

struct UnitGeo { int b, row0, col0, rflags; };  // rflags = window row << 3 | right << 2 | bottom << 1 | valid

// window slot (ih, iw) of tile row t: slot order, or the two-column-part order of a right-edge window
__device__ __forceinline__ int tile_row_slot(int t, int shift) {
  if (shift == 0) return t;
  const int wa = kWinSide - shift;
  int ih, iw;
  if (t < kWinSide * wa) { ih = t / wa; iw = t - ih * wa; }
  else { const int t2 = t - kWinSide * wa; ih = t2 / shift; iw = wa + t2 - ih * shift; }
  return ih << 3 | iw;
}

// Tensor maps per tensor, box (w, h): [0] full (8, 8) | column-split order: [1] (wa, 8) [2] (s, 8) [3] (wa, wa) [4] (wa, s)
// [5] (s, wa) [6] (s, s) | row wrap in slot order: [7] (8, wa) [8] (8, s)
constexpr int kNumWinMaps = 9;

// The TMA boxes of one tile: f(byte offset inside the tile, map index, image column, image row).  kMode = window class
// (0: one box | 1: slot order, the rows of a bottom window wrap | 2: two column parts [0, wa) | [wa, 8), both wrap along
// the rows in a bottom window).  The same list drives loads and stores.
template <int kMode, typename F>
__device__ __forceinline__ void for_each_box(const Geom& g, int col0, int row0, bool bottom, F&& f) {
  const int sh = g.shift, wa = kWinSide - g.shift;
  if (kMode == 0) {
    f(0, 0, col0, row0);
  } else if (kMode == 1) {
    if (!bottom) {
      f(0, 0, col0, row0);
    } else {
      f(0, 7, col0, row0);
      f(wa * kWinSide * 64, 8, col0, 0);
    }
  } else {
    int colb = col0 + wa;
    if (colb >= g.W) colb -= g.W;
    const int offb = kWinSide * wa * 64;
    if (!bottom) {
      f(0, 1, col0, row0);
      f(offb, 2, colb, row0);
    } else {
      f(0, 3, col0, row0);
      f(wa * wa * 64, 4, col0, 0);
      f(offb, 5, colb, row0);
      f(offb + sh * wa * 64, 6, colb, 0);
    }
  }
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

template <bool V> struct BoolTag { static constexpr bool value = V; };
template <int V> struct IntTag { static constexpr int value = V; };

// ---- host side ---------------------------------------------------------------------------------------------------
// The nine tensor maps of one (B, H*W, row_elems) bf16 tensor for geometry g
inline int make_window_maps(CUtensorMap* m, const void* base, const Geom& g, int row_elems) {
  const int s = g.shift, wa = kWinSide - g.shift, W8 = kWinSide;
  const int bw[kNumWinMaps] = {W8, s ? wa : W8, s ? s : W8, s ? wa : W8, s ? wa : W8, s ? s : W8, s ? s : W8, W8, W8};
  const int bh[kNumWinMaps] = {W8, W8, W8, s ? wa : W8, s ? s : W8, s ? wa : W8, s ? s : W8, s ? wa : W8, s ? s : W8};
  for (int i = 0; i < kNumWinMaps; ++i) {
    const int rc = make_map(&m[i], base, g, row_elems, bw[i], bh[i]);
    if (rc) return rc;
  }
  return HV_OK;
}

// CTAs per head group and window class for a launch on nsm SMs; returns the grid size.  Classes get CTAs in proportion to
// windows x relative cost per window (cost[1], cost[2]: a wrapped window needs twice the TMA boxes and the mask).
inline int plan_window_schedule(const Geom& g, int nsm, const double* cost, WinSchedule& sc) {
  const int nrows = g.B * g.nW;
  sc.n_same = g.heads / 2;
  sc.has_cross = g.heads & 1;
  if (sc.n_same == 0) {
    sc.ctas_same = 0;
    sc.ctas_cross = nsm;
  } else if (!sc.has_cross) {
    sc.ctas_same = nsm / sc.n_same;
    sc.ctas_cross = 0;
  } else {
    sc.ctas_cross = nsm / (2 * sc.n_same + 1);
    if (sc.ctas_cross < 1) sc.ctas_cross = 1;
    sc.ctas_same = (nsm - sc.ctas_cross) / sc.n_same;
  }
  if (sc.ctas_same < 1 && sc.n_same) sc.ctas_same = 1;
  if (sc.ctas_same > nrows) sc.ctas_same = nrows;
  if (sc.ctas_cross > (nrows + 1) / 2) sc.ctas_cross = (nrows + 1) / 2;
  const int nWh = g.H / kWinSide;
  int n[3] = {nrows, 0, 0};
  if (g.shift > 0) {
    n[0] = g.B * (nWh - 1) * (g.nWw - 1);
    n[1] = g.B * (g.nWw - 1);
    n[2] = g.B * nWh;
  }
  const double c[3] = {1.0, cost[1], cost[2]};
  auto split = [&](int& ctas, int* out) {
    out[0] = out[1] = out[2] = 0;
    if (ctas == 0) return;
    int nonempty = 0;
    double wsum = 0;
    for (int k = 0; k < 3; ++k) { nonempty += n[k] > 0; wsum += n[k] * c[k]; }
    if (ctas < nonempty) ctas = nonempty;
    int used = 0, big = -1;
    for (int k = 0; k < 3; ++k) {
      if (n[k] == 0) continue;
      out[k] = (int)(ctas * n[k] * c[k] / wsum + 0.5);
      if (out[k] < 1) out[k] = 1;
      used += out[k];
      if (big < 0 || n[k] * c[k] > n[big] * c[big]) big = k;
    }
    out[big] += ctas - used;  // rounding goes to the largest class
    if (out[big] < 1) { ctas += 1 - out[big]; out[big] = 1; }
  };
  split(sc.ctas_same, sc.cls_same);
  split(sc.ctas_cross, sc.cls_cross);
  return sc.n_same * sc.ctas_same + sc.ctas_cross;
}

}  // namespace tc
}  // namespace hv
