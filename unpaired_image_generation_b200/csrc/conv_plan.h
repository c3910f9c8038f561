// Host-side planning of the implicit-GEMM convolutions: K-iteration (tap) tables, TMA tensor
// maps, tile shapes, packed-weight layouts.  One plan is built once per (layer, pass, buffers)
// and replayed every step; all shapes are static so plans are CUDA-graph friendly.
//
// Layer semantics follow the stand-in oracle (oracle/cyclegan_standin.py): Conv2d with
// reflection or zero padding, stride 1 or 2, and ConvTranspose2d(k=3, s=2, p=1, op=1).
#pragma once
#include <vector>

#include "common.h"
#include "conv_tc.h"

namespace cgb {

// NHWC bf16 activation resident in HBM, optionally stored with a halo of `halo` pixels on each
// side (used for reflection-padded conv inputs: the producer writes the mirrored border).
struct TensorDesc {
  bf16* ptr = nullptr;  // storage base: [N][H + 2*halo][W + 2*halo][C]
  int N = 0, H = 0, W = 0, C = 0;
  int halo = 0;
  int esz = 2;          // bytes per element: 2 = bf16 (the product path), 4 = fp32 (validation mode, fp32_path.h);
                        // strides below stay in ELEMENTS, `ptr` keeps its bf16* type and is re-cast by the fp32 kernels
  long long sW() const { return C; }
  long long sH() const { return (long long)(W + 2 * halo) * C; }
  long long sN() const { return (long long)(H + 2 * halo) * sH(); }
  long long elems() const { return (long long)N * sN(); }
  // address of element offset `off` from the storage base (scaled by the element size)
  bf16* at(long long off) const { return reinterpret_cast<bf16*>(reinterpret_cast<char*>(ptr) + off * esz); }
  bf16* interior() const { return at(halo * sH() + halo * sW()); }
  size_t bytes() const { return (size_t)elems() * esz; }
  // view of n consecutive images starting at image `first`
  TensorDesc images(int first, int n) const {
    TensorDesc v = *this;
    v.ptr = at((long long)first * sN());
    v.N = n;
    return v;
  }
};

struct ConvSpec {
  int Cin = 0, Cout = 0;   // logical channels
  int CinS = 0, CoutS = 0; // stored channels of the input / output activations (>= logical, zero padded)
  int k = 3, stride = 1, pad = 1;
  bool reflect = false;    // reflection padding (input tensor carries a halo == pad)
  bool transposed = false; // ConvTranspose2d(k=3, stride=2, padding=1, output_padding=1)
  int taps() const { return k * k; }
};

// Packed bf16 weights.  Master fp32 layout is [Cout][T][Cin]; packs add zero padding:
//   Wf [rows = CoutP][T * CinS]  (fprop of conv and of transposed conv)
//   Wt [rows = CinP ][T * CoutS] (input gradients)
// rows are padded up to a multiple of the N tile.
long long packed_wf_elems(const ConvSpec& s);
long long packed_wt_elems(const ConvSpec& s);
int padded_rows(int c);  // 16 for c <= 16, else next multiple of 64

struct IgemmPlan {
  CUtensorMap tmA, tmB;
  IgemmArgs args;
  int BN = 0, BK = 0, num_tiles = 0, n_blocks = 0, n_classes = 1;
  int CM = 1, CN = 1;         // thread-block cluster (M tiles x N blocks) sharing operands by TMA multicast
  bool lite = false;          // tap-table kernel: low-shared-memory instantiation, several CTAs per SM
  std::vector<KIter> kiters;  // host copy; args.kiters must point at a device copy
  double flops = 0;           // algorithmic 2*MACs (for roofline accounting)
  // patch-resident variant (conv_patch.cu): stride-1 convs / input gradients with 64-channel chunks
  bool patch = false;
  PatchArgs pargs;
  int MT = 1;                 // stacked 16 x 8 M tiles per CTA
  int CG = 1;                 // 2: CTA pairs (cta_group::2) sharing one weight tile
  bool tapn = false;          // taps-in-N kernel (conv_tapn.cu): k x k, 64 -> <= 4 channels, stride 1
  TapNArgs targs;
  int num_ctas_m = 0;
};

struct WgradPlan {
  CUtensorMap tmDY, tmX;
  WgradArgs args;
  int BNW = 0, m_blocks = 0;
  std::vector<WTap> taps;
  double flops = 0;
  bool pair = false;     // stride-1 layers with Cout % 256 == 0, Cin % 128 == 0: CTA-pair kernel (wgrad_pair.cu)
  WgradPairArgs pargs;
};

// Forward of Conv2d (stride 1 or 2) or ConvTranspose2d.  x: input (with halo when reflect),
// y: output interior is written (y may itself carry a halo; only the interior is touched).
IgemmPlan plan_fprop(const ConvSpec& s, const TensorDesc& x, const bf16* wf, const TensorDesc& y, const float* bias,
                     int act, int sm_count);

// Input gradient.  dy: gradient w.r.t. the conv output (interior view used, zero outside).
// dx: for reflect convs this is the PADDED-DOMAIN gradient tensor (H+2p, W+2p, halo 0) that the
// consumer folds back; otherwise the plain input-shaped gradient.
IgemmPlan plan_dgrad(const ConvSpec& s, const TensorDesc& dy, const bf16* wt, const TensorDesc& dx, int sm_count);

// Weight gradient into g[Cout][T][Cin] (fp32, pre-zeroed, accumulated with atomics).
WgradPlan plan_wgrad(const ConvSpec& s, const TensorDesc& x, const TensorDesc& dy, float* g, int sm_count);

void run(const IgemmPlan& p, cudaStream_t stream);
void run(const WgradPlan& p, cudaStream_t stream);

// True when the tensor-core path supports the pass for this layer (channel counts).
bool tc_supports_fprop(const ConvSpec& s);
bool tc_supports_dgrad(const ConvSpec& s);
bool tc_supports_wgrad(const ConvSpec& s);

int out_extent(const ConvSpec& s, int in_extent);

}  // namespace cgb
