// Weight gradients of the 3-/1-channel layers (generator stem and head, discriminator conv0 and conv4):
// the skinny operand is expanded by an explicit im2col (pointwise.h: im2col_small) so that the reduction
// over pixels becomes a plain tcgen05 GEMM (conv_tc.cu: wgrad_kernel) writing straight into the master
// gradient layout g[Cout][T][Cin].
#pragma once
#include <vector>

#include "conv_plan.h"

namespace cgb {

struct SmallWgradPlan {
  WgradPlan gemm;            // gemm.args.taps / row_map must point at device copies before run()
  std::vector<int> row_map;  // (tap*4 + cout) -> row of g   (im2col on the output-gradient side)
  std::vector<int> col_map;  // (tap*4 + cin)  -> column of g (im2col on the input side)
  // im2col parameters
  TensorDesc src, col;
  int k = 0, stride = 1, sgn = 1, off = 0;
  bool use_halo = false;
  bool col_is_precomputed = false;  // the caller keeps `col` up to date (shared stem im2col)
  // virtual im2col (7x7 head and stem): `col` is the ROW-EXPANDED gradient [N][Hp + 8][Wp][64] (pointwise.h expand_rows4) and the
  // GEMM's dY-side tensor map reads rows h, h + 2, h + 4, h + 6 of it as the four 64-column atoms (overlapping strides)
  bool virtual_rows = false;
  double flops = 0;
};

size_t small_wgrad_col_elems(const ConvSpec& s, const TensorDesc& x, const TensorDesc& dy);
// precomputed_col: an im2col4 matrix of x that the caller maintains (input-side layers only), else nullptr
SmallWgradPlan plan_wgrad_small(const ConvSpec& s, const TensorDesc& x, const TensorDesc& dy, float* g, bf16* colbuf,
                                size_t colbuf_elems, int sm_count, const TensorDesc* precomputed_col = nullptr);
int im2col4_width(int taps);  // stored columns of an im2col4 matrix: 64 or 256
// true: k x k layers with <= 4 channels on one side use the virtual im2col (row-expanded tensor [N][H + 8][W][64] read
// through an overlapping-stride tensor map) instead of a materialised im2col4 matrix; CGB_VIRTUAL_COL=0 disables
bool small_wgrad_virtual(int k);
bool small_wgrad_virtual_in(int k);  // ... on the input side (the stem: CGB_VIRTUAL_COL_IN=0 keeps its im2col4 matrix)
void run(const SmallWgradPlan& p, cudaStream_t stream);

}  // namespace cgb
