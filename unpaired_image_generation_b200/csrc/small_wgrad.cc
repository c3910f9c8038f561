#include "small_wgrad.h"

#include "pointwise.h"

#include <cstdlib>

namespace cgb {

static bool virtual_rows_enabled() {
  static const bool on = !(std::getenv("CGB_VIRTUAL_COL") && std::atoi(std::getenv("CGB_VIRTUAL_COL")) == 0);
  return on;
}

int im2col4_width(int taps) { return taps * 4 <= 64 ? 64 : 256; }
bool small_wgrad_virtual(int k) { return virtual_rows_enabled() && k >= 5 && k <= 8; }
bool small_wgrad_virtual_in(int k) {
  static const bool on = !(std::getenv("CGB_VIRTUAL_COL_IN") && std::atoi(std::getenv("CGB_VIRTUAL_COL_IN")) == 0);
  return on && small_wgrad_virtual(k);
}

size_t small_wgrad_col_elems(const ConvSpec& s, const TensorDesc& x, const TensorDesc& dy) {
  const int Kp = im2col4_width(s.taps());
  if (s.Cin <= 16) return (size_t)dy.N * dy.H * dy.W * Kp;
  const int halo = s.reflect ? x.halo : 0;
  return (size_t)x.N * (x.H + 2 * halo) * (x.W + 2 * halo) * Kp;
}

SmallWgradPlan plan_wgrad_small(const ConvSpec& s, const TensorDesc& x, const TensorDesc& dy, float* g, bf16* colbuf,
                                size_t colbuf_elems, int sm_count, const TensorDesc* precomputed_col) {
  CGB_CHECK(!s.transposed, "small-channel transposed convs are not part of the model");
  CGB_CHECK(s.Cin <= 4 || s.Cout <= 4, "plan_wgrad_small is for 3-/1-channel layers");
  SmallWgradPlan p;
  const int T = s.taps();
  const int Kp = im2col4_width(T);
  CGB_CHECK(T * 4 <= Kp, "im2col row too long");
  p.flops = 2.0 * dy.N * (double)dy.H * dy.W * s.Cout * s.Cin * T;
  ConvSpec s1;
  s1.k = 1;
  s1.stride = 1;
  s1.pad = 0;
  p.k = s.k;
  if (s.Cin <= 4) {
    // im2col on the input side: g[co][(t, ci)] = sum_px dy[px][co] * col[px][t*4 + ci]
    // 7x7 stem: VIRTUAL im2col -- col is the row-expanded input [N][H + 8][W][64] (expand_rows4, reflect halo
    // included) and the GEMM's X-side tensor map reads rows h, h + 2, h + 4, h + 6 of it as the four 64-column atoms;
    // GEMM column j = r * 32 + s * 4 + ci.
    const bool virt = small_wgrad_virtual_in(s.k) && Kp == 256 && s.stride == 1;
    if (precomputed_col) {
      p.col = *precomputed_col;
      p.col_is_precomputed = true;
      if (virt)
        CGB_CHECK(p.col.N == dy.N && p.col.H == dy.H + 8 && p.col.W == dy.W && p.col.C == 64, "precomputed row-expanded shape");
      else
        CGB_CHECK(p.col.N == dy.N && p.col.H == dy.H && p.col.W == dy.W && p.col.C == Kp, "precomputed im2col shape");
    } else {
      p.col.ptr = colbuf; p.col.halo = 0;
      p.col.N = dy.N; p.col.H = virt ? dy.H + 8 : dy.H; p.col.W = dy.W; p.col.C = virt ? 64 : Kp;
      CGB_CHECK((size_t)p.col.elems() <= colbuf_elems, "im2col scratch too small");
    }
    p.src = x; p.stride = s.stride; p.sgn = +1; p.off = -s.pad; p.use_halo = s.reflect;
    s1.Cin = Kp; s1.CinS = Kp; s1.Cout = s.Cout; s1.CoutS = s.CoutS;
    TensorDesc colx = p.col;  // what the GEMM sees: [N][dy.H][dy.W][Kp]
    colx.H = dy.H;
    colx.C = Kp;
    p.gemm = plan_wgrad(s1, colx, dy, g, sm_count);
    p.gemm.args.Cin = T * s.Cin;  // row length of g
    p.gemm.args.ncols = Kp;
    p.col_map.assign(Kp, -1);
    if (virt) {
      p.virtual_rows = true;
      const int TWk = 1 << p.gemm.args.tw_shift, THk = 64 >> p.gemm.args.tw_shift;
      const int dims[5] = {64, dy.W, 4, dy.H, dy.N};
      const long long str[4] = {64, 2LL * dy.W * 64, (long long)dy.W * 64, (long long)p.col.H * dy.W * 64};
      p.gemm.tmX = make_tmap_act5d(p.col.ptr, dims, str, 64, TWk, THk, 128);
      p.gemm.args.b_virtual = 1;
      for (int r = 0; r < s.k; ++r)
        for (int sx = 0; sx < s.k; ++sx)
          for (int ci = 0; ci < s.Cin; ++ci) p.col_map[r * 32 + sx * 4 + ci] = (r * s.k + sx) * s.Cin + ci;
    } else {
      for (int j = 0; j < T * 4; ++j)
        if (j % 4 < s.Cin) p.col_map[j] = (j / 4) * s.Cin + j % 4;
    }
  } else {
    // im2col on the output-gradient side: g[(t, co)][ci] = sum_px col[px][t*4 + co] * x[px][ci]
    CGB_CHECK(s.stride == 1, "strided skinny-output conv");
    CGB_CHECK(small_wgrad_col_elems(s, x, dy) <= colbuf_elems, "im2col scratch too small");
    TensorDesc x1 = x;  // pixel domain = input pixels as the conv sees them (padded domain for reflect convs)
    if (s.reflect) {
      x1.H = x.H + 2 * x.halo; x1.W = x.W + 2 * x.halo; x1.halo = 0;
    }
    p.col.ptr = colbuf; p.col.halo = 0;
    p.col.N = x1.N; p.col.H = x1.H; p.col.W = x1.W; p.col.C = Kp;
    p.src = dy; p.stride = 1; p.sgn = -1; p.off = s.reflect ? 0 : s.pad; p.use_halo = false;
    s1.Cin = s.Cin; s1.CinS = s.CinS; s1.Cout = Kp; s1.CoutS = Kp;
    p.gemm = plan_wgrad(s1, x1, p.col, g, sm_count);
    p.row_map.assign(Kp, -1);
    // 7x7 (5 <= k <= 8): VIRTUAL im2col.  Only the k horizontal taps are materialised (32 columns per pixel instead
    // of 256); the k vertical taps are rows h .. h + k - 1 of that tensor, read through a tensor map whose row-tap
    // dimension overlaps its height dimension.  GEMM row m = r' * 32 + s * 4 + co with r' = k - 1 - r (the expansion
    // is shifted down by k - 1 rows so that every row tap is a non-negative offset).
    if (virtual_rows_enabled() && Kp == 256 && s.k >= 5 && s.k <= 8 &&
        (size_t)x1.N * (x1.H + 8) * x1.W * 64 <= colbuf_elems) {
      p.virtual_rows = true;
      p.col.H = x1.H + 8;  // k - 1 rows above, spare rows below; every element is written by expand_rows4
      p.col.C = 64;        // per pixel: the 32 expanded columns of its row and of the row below
      const int TWk = 1 << p.gemm.args.tw_shift, THk = 64 >> p.gemm.args.tw_shift;
      const int dims[5] = {64, x1.W, 4, x1.H, x1.N};
      const long long str[4] = {64, 2LL * x1.W * 64, (long long)x1.W * 64, (long long)p.col.H * x1.W * 64};
      p.gemm.tmDY = make_tmap_act5d(colbuf, dims, str, 64, TWk, THk, 128);
      p.gemm.args.a_virtual = 1;
      for (int rp = 0; rp < s.k; ++rp)
        for (int sx = 0; sx < s.k; ++sx)
          for (int co = 0; co < s.Cout; ++co) p.row_map[rp * 32 + sx * 4 + co] = co * T + (s.k - 1 - rp) * s.k + sx;
    } else {
      for (int j = 0; j < T * 4; ++j)
        if (j % 4 < s.Cout) p.row_map[j] = (j % 4) * T + j / 4;
    }
  }
  p.gemm.flops = p.flops;
  return p;
}

void run(const SmallWgradPlan& p, cudaStream_t stream) {
  if (p.virtual_rows && !p.col_is_precomputed) expand_rows4(p.src, p.k, p.sgn, p.off, p.use_halo, p.col, stream);
  else if (!p.virtual_rows && !p.col_is_precomputed) im2col4(p.src, p.k, p.stride, p.sgn, p.off, p.use_halo, p.col, stream);
  run(p.gemm, stream);
}

}  // namespace cgb
