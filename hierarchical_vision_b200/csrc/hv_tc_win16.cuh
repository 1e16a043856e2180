// Window geometry shared by the tcgen05 / TMEM / TMA attention kernels for 16x16 windows (N = 256, head dim 32, bf16:
// SwinV2-B at window 16, reference swinv2.py:105-283 with window_size = 16).
//
// One token order for every window (TILE order): the window is held as two COLUMN PARTS of 8 columns x 16 rows, tile
// row t = 128 part + 8 ih + (iw & 7).  A part is one TMA box (8 w x 16 h tokens x 64 B of one head), so the cyclic shift
// by ws / 2 = 8 (the only shift SwinV2 uses, swinv2.py:560) never splits a box along the columns: the second part of a
// right-edge window simply starts at image column 0.  Only the bottom row of windows wraps along the rows and takes two
// 8 x 8 boxes per part.  window_partition / window_reverse and torch.roll (swinv2.py:399-429) are these coordinates.
//
// In tile order a 64-row block of the tile is an 8 x 8 spatial sub-block of the window and a 128-row block a whole column
// part, which is what makes the shift mask of a (128 query) x (64 key) block all-or-nothing per query row half, and the
// relative-position lookup of such a block a Toeplitz window of the 31 x 31 table.
#pragma once
#include "hv_tc_win.cuh"

namespace hv {
namespace tc {

constexpr int kWin16 = 16;
constexpr int kN16 = 256;
constexpr int kTab16 = 31;                // relative offsets per axis
constexpr int kTile16 = kN16 * 64;        // one (window, head) q / k / v / dO tile: 256 rows x 64 B (SWIZZLE_64B)
constexpr int kBiasStride16 = 40;         // floats per table row (a multiple of 4: the alignment of a run does not depend on the row)
// Four ALIGNMENT COPIES of the table, copy c shifted by c floats: a thread reads runs of 8 consecutive entries starting at
// a column whose residue mod 4 is fixed by its query column, picks the copy that makes the run 16-byte aligned, and loads
// it as two float4.  The 8 lanes of a quarter warp (one window row of queries, columns 0 .. 7) then read the runs
// (copy, start) = (1, U) (2, U) (3, U) (0, U - 4) (1, U - 4) (2, U - 4) (3, U - 4) (0, U - 8): with the copies based at 0, 4,
// 12, 20 (mod 32) floats the eight float4 fall into eight different 4-bank groups -- a uniform copy stride cannot do that
// (measured: 2-way conflicts on every load with a stride of 8 mod 32).
constexpr int kBiasCopyLen16 = kTab16 * kBiasStride16 + 4;  // entries of one copy, shift included
__host__ __device__ constexpr int bias_copy_base16(int c) { return c * 1248 + (c == 0 ? 0 : 8 * c - 4); }
constexpr int kBiasFloats16 = bias_copy_base16(3) + kBiasCopyLen16 + 12;
static_assert(kBiasFloats16 % 4 == 0, "bias copies end on a 16-byte boundary");

// fill the alignment copies: entry (dy, x) = scale * table[(dy, 30 - x), head] - off, dy = ih - jh + 15, x = 15 - iw + jw
// (REVERSED column, so the 8 keys of a window row are 8 consecutive floats)
__device__ __forceinline__ void fill_bias16(float* bt, const float* __restrict__ bias_table, int heads, int head, float scale,
                                            float off, int tid, int nthreads) {
  for (int idx = tid; idx < 4 * kBiasCopyLen16; idx += nthreads) {
    const int c = idx / kBiasCopyLen16, pos = idx - c * kBiasCopyLen16, j = pos - c;  // copy_c[j + c] = table entry j
    const int dy = j / kBiasStride16, x = j - dy * kBiasStride16;
    bt[bias_copy_base16(c) + pos] =
        (j >= 0 && dy < kTab16 && x < kTab16) ? scale * __ldg(&bias_table[(dy * kTab16 + 30 - x) * heads + head]) - off : 0.f;
  }
}
// pointer to the aligned run that holds table entries i0 .. i0 + 7 (i0 = dy * 40 + x0)
__device__ __forceinline__ const float* bias_run16(const float* bt, int i0) {
  const int c = (4 - (i0 & 3)) & 3;
  return bt + bias_copy_base16(c) + i0 + c;
}

// Wait with a real sleep between polls (ns = 0: the plain try_wait loop).  An experiment switch: a dozen waiting warps in
// tight try_wait loops execute a third of the kernel's instructions, but letting the off-path warps (producer, store warp,
// epilogues) sleep 64-1000 ns between polls made both N = 256 kernels 3-4 % SLOWER (late wake-ups), so the kernels pass 0
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, uint32_t ns) {
  if (ns == 0) { mbar_wait_fast(bar, parity); return; }
  uint32_t spins = 0;
  while (!mbar_test(bar, parity)) {
    asm volatile("nanosleep.u32 %0;" ::"r"(ns));
    if (++spins > (1u << HV_SPIN_LIMIT_LOG2)) __trap();
  }
}  // (mbarrier.test_wait is an acquire at CTA scope by default, like try_wait)

__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

struct UnitGeo16 { int b, row0, col0, flags; };  // flags = window row (b * nW + win) << 2 | right << 1 | bottom

// Units (windows) of one head walked by a CTA: heads x cph CTAs, CTA c serves head c / cph, windows first, first + cph, ...
struct Work16 {
  int head, first, stride, nunits;
  __device__ __forceinline__ void init(int cph, int total) {
    head = blockIdx.x / cph;
    first = blockIdx.x - head * cph;
    stride = cph;
    nunits = first < total ? (total - first + cph - 1) / cph : 0;
  }
};

HV_HD UnitGeo16 unit_geo16(const Geom& g, int widx) {
  const int b = widx / g.nW, win = widx - b * g.nW;
  const int wh = win / g.nWw, ww = win - wh * g.nWw;
  UnitGeo16 ug;
  ug.b = b;
  ug.row0 = wh * kWin16 + g.shift;
  ug.col0 = ww * kWin16 + g.shift;
  const int bottom = g.shift > 0 && wh == g.H / kWin16 - 1, right = g.shift > 0 && ww == g.nWw - 1;
  ug.flags = (widx << 2) | (right << 1) | bottom;
  return ug;
}

// The TMA boxes of one tile: f(byte offset inside the tile, map index (0: 8 x 16 box, 1: 8 x 8 box), image column, image
// row).  The same list drives loads and stores.
template <typename F>
__device__ __forceinline__ void for_each_box16(const Geom& g, const UnitGeo16& ug, F&& f) {
  int colb = ug.col0 + 8;
  if (colb >= g.W) colb -= g.W;
  if (!(ug.flags & 1)) {
    f(0, 0, ug.col0, ug.row0);
    f(8192, 0, colb, ug.row0);
  } else {  // rows [row0, row0 + 8) then the wrapped rows [0, 8)
    f(0, 1, ug.col0, ug.row0);
    f(4096, 1, ug.col0, 0);
    f(8192, 1, colb, ug.row0);
    f(12288, 1, colb, 0);
  }
}

// image (row, col) of tile row t of a window -- for the plain-load helper kernels
HV_HD void tile_row_rc16(const Geom& g, const UnitGeo16& ug, int t, int& row, int& col) {
  const int part = t >> 7, ih = (t & 127) >> 3, iw = 8 * part + (t & 7);
  row = ug.row0 + ih;
  col = ug.col0 + iw;
  if (row >= g.H) row -= g.H;
  if (col >= g.W) col -= g.W;
}

inline int make_window_maps16(CUtensorMap* m, const void* base, const Geom& g, int row_elems) {
  int rc = make_map(&m[0], base, g, row_elems, 8, 16);
  if (rc) return rc;
  return make_map(&m[1], base, g, row_elems, 8, 8);
}

}  // namespace tc
}  // namespace hv
