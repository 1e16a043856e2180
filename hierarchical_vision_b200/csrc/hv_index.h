// Closed-form integer maps of the SwinV2 window path, shared by host code and sm_100a kernels.
//
// Every function here is `__host__ __device__` so that the *same* arithmetic that the kernels
// use for their index-remapped loads/stores can be evaluated on the CPU through the C-ABI
// (hv_relative_position_index / hv_shift_window_mask / hv_window_token_index) and compared
// bit-for-bit with the buffers the reference builds:
//   relative_position_index  reference swinv2.py:175-190
//   attn_mask                reference swinv2.py:357-388
//   roll + window_partition  reference swinv2.py:69-83, 399-412 (inverse: 86-102, 420-429)
//   PatchMerging concat      reference swinv2.py:486-490
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HV_HD __host__ __device__ __forceinline__
#else
#define HV_HD inline
#endif

namespace hv {

struct Geom {
  int B, H, W, C, heads, ws, shift;
  int d;    // C / heads
  int N;    // ws * ws
  int nWw;  // W / ws   (windows per row)
  int nW;   // windows per image
};

HV_HD Geom make_geom(int B, int H, int W, int C, int heads, int ws, int shift) {
  Geom g;
  g.B = B; g.H = H; g.W = W; g.C = C; g.heads = heads; g.ws = ws; g.shift = shift;
  g.d = C / heads;
  g.N = ws * ws;
  g.nWw = W / ws;
  g.nW = (H / ws) * g.nWw;
  return g;
}

// Token (b, image row, image col) feeding slot i of window `win` (win = wh * nWw + ww inside
// image b): the cyclic shift by -shift composed with the window partition.
HV_HD void window_slot_to_rc(const Geom& g, int win, int slot, int& row, int& col) {
  const int wh = win / g.nWw, ww = win - wh * g.nWw;
  const int ih = slot / g.ws, iw = slot - ih * g.ws;
  row = wh * g.ws + ih + g.shift;
  col = ww * g.ws + iw + g.shift;
  if (row >= g.H) row -= g.H;
  if (col >= g.W) col -= g.W;
}

HV_HD int64_t window_slot_to_token(const Geom& g, int b, int win, int slot) {
  int row, col;
  window_slot_to_rc(g, win, slot, row, col);
  return ((int64_t)b * g.H + row) * g.W + col;
}

// Band of a *shifted* coordinate p along an axis of length L: 0 for [0, L-ws), 1 for
// [L-ws, L-shift), 2 for [L-shift, L).
HV_HD int shift_band(int p, int L, int ws, int shift) {
  return p < L - ws ? 0 : (p < L - shift ? 1 : 2);
}

// Region id (0..8) of slot `slot` of window `win`; two slots may attend to each other iff
// their ids are equal, otherwise the logit gets -100.
HV_HD int window_slot_region(const Geom& g, int win, int slot) {
  const int wh = win / g.nWw, ww = win - wh * g.nWw;
  const int ih = slot / g.ws, iw = slot - ih * g.ws;
  return 3 * shift_band(wh * g.ws + ih, g.H, g.ws, g.shift) + shift_band(ww * g.ws + iw, g.W, g.ws, g.shift);
}

// Index into the ((2ws-1)^2, heads) continuous-position-bias table for the pair (i, j).
HV_HD int rel_pos_index(int ws, int i, int j) {
  const int ih = i / ws, iw = i - ih * ws;
  const int jh = j / ws, jw = j - jh * ws;
  return (ih - jh + ws - 1) * (2 * ws - 1) + (iw - jw + ws - 1);
}

// PatchMerging: source token of channel block m (0..3) of output token (b, i, j).
HV_HD int64_t merge_src_token(int H, int W, int b, int i, int j, int m) {
  const int row = 2 * i + (m & 1);
  const int col = 2 * j + (m >> 1);
  return ((int64_t)b * H + row) * W + col;
}

}  // namespace hv
