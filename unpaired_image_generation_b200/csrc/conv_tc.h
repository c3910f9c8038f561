// Host-side entry points of the tcgen05 convolution kernels (conv_tc.cu) and the TMA
// tensor-map builder (tmap.cc).
#pragma once
#include "common.h"

namespace cgb {

// 5-D activation view: dims (d0 channels, d1 width, d2 parity plane, d3 height, d4 image),
// strides in ELEMENTS for d1..d4 (d0 is contiguous), box = (box_c, box_w, 1, box_h, 1).
// swizzle_bytes must equal box_c * 2 (32, 64 or 128).  Strides may OVERLAP (two dimensions walking the same rows): that is
// how the weight gradient of the 7x7 head reads a row-expanded gradient tensor as a virtual im2col matrix (small_wgrad.cc).
CUtensorMap make_tmap_act5d(const bf16* base, const int dims[5], const long long strides_elems[4], int box_c,
                            int box_w, int box_h, int swizzle_bytes);

// 2-D K-major matrix [rows][cols] bf16 with row pitch `pitch_elems`; box = (box_cols, box_rows).
CUtensorMap make_tmap_2d(const bf16* base, long long rows, long long cols, long long pitch_elems, int box_cols,
                         int box_rows, int swizzle_bytes);

// CM x CN: thread-block cluster (M tiles x N blocks) sharing operands through TMA multicast; 1 x 1 = none.
// lite: 2-stage low-shared-memory instantiation (2-3 CTAs per SM) for launches of many short CTAs.
void launch_igemm(int BN, int BK, int CM, int CN, bool lite, const CUtensorMap& tmA, const CUtensorMap& tmB,
                  const IgemmArgs& args, int num_tiles, int n_blocks, int n_classes, cudaStream_t stream);


// Patch-resident variant for stride-1 convs / stride-1 input gradients with 64-channel chunks (conv_patch.cu).
// MT = M tiles (16 x 8 output pixels each, stacked along h) per CTA; KPS = filter taps per weight stage;
// CG = 2: CTA pairs (cta_group::2), tmB's box then carries BN / 2 rows and num_ctas_m is even.
void launch_igemm_patch(int BN, int MT, int CG, const CUtensorMap& tmA, const CUtensorMap& tmB, const IgemmArgs& args,
                        const PatchArgs& pa, int num_ctas_m, int n_blocks, cudaStream_t stream);
int igemm_patch_kps(int BN, int ka);  // weight boxes carried by one stage for this N tile / patch-row width
int igemm_patch_smem_budget(); // bytes available for the weight ring + patches

void launch_wgrad(int BNW, const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradArgs& args, int m_blocks,
                  cudaStream_t stream);

// 4-D view of a packed weight matrix W[rows >= 5][k * k * 64] for the taps-in-N kernel: dims (64 channels, 4 rows,
// 8 taps, k filter rows), one box = the whole filter (k x 4 KB: rows tx * 4 + row of filter row ty at ty * 4096).
CUtensorMap make_tmap_tapn_weights(const bf16* base, int k, long long pitch_elems);
// Taps-in-N conv (conv_tapn.cu).  tmA: box (64 channels, 16, 1, 16 + k - 1, 1) of the activation view.
void launch_conv_tapn(const CUtensorMap& tmA, const CUtensorMap& tmB, const TapNArgs& args, int num_ctas, cudaStream_t stream);

// CTA-pair weight gradient of stride-1 convs (wgrad_pair.cu).  tmDY: box (64 channels, 8, 1, 8, 1) of the dY view;
// tmX: box (64 channels, 8 + kw - 1, 1, 8, 1) of the (padded, for reflect convs) X view.
void launch_wgrad_pair(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradPairArgs& args, cudaStream_t stream);

}  // namespace cgb
