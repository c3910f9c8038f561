// Shared host/device definitions for the B200 CycleGAN hot path.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>

namespace cgb {

typedef __nv_bfloat16 bf16;

// Thread-local last-error string, surfaced through the C ABI (cgb_last_error()).
void set_last_error(const std::string& msg);

struct Error : public std::runtime_error {
  explicit Error(const std::string& m) : std::runtime_error(m) {}
};

#define CGB_CHECK(cond, msg)                                                                  \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      throw ::cgb::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + msg); \
    }                                                                                         \
  } while (0)

#define CGB_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      throw ::cgb::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": CUDA " +   \
                         cudaGetErrorString(_e) + " in " #expr);                               \
    }                                                                                          \
  } while (0)

// One K-iteration of the implicit GEMM: which 5-D box of the activation tensor to fetch
// (relative to the tile's output origin) and which K offset of the packed weights.
struct KIter {
  int16_t a_c;    // coordinate 0: channel (already includes the w-parity * C offset)
  int16_t a_dx;   // added to the tile's w origin -> coordinate 1
  int16_t a_par;  // coordinate 2: h-parity plane (0 for stride-1 views)
  int16_t a_dy;   // added to the tile's h origin -> coordinate 3
  int32_t b_k;    // K coordinate in the packed weight matrix [Cout_pad][Ktot]
  int32_t pad;
};
static_assert(sizeof(KIter) == 16, "KIter must be 16 bytes");

constexpr int kMaxClasses = 4;  // output-parity classes of a stride-2 transposed conv / dgrad

// Arguments of the tcgen05 implicit-GEMM kernel (fprop, dgrad and transposed fprop are
// all expressed through the K-iteration table).
struct IgemmArgs {
  const KIter* kiters;
  int k_begin[kMaxClasses];
  int k_count[kMaxClasses];
  long long out_off[kMaxClasses];  // element offset of the class's first output pixel
  int tiles_w, tiles_h;            // tiles per image (tile = TH x TW output pixels, TH*TW = 128)
  int N;                           // images (tile indices past N * tiles are padding of a cluster launch)
  int tw_shift;                    // TW = 1 << tw_shift
  int Ho, Wo;                      // valid output extent in class-local coordinates
  int Cout;                        // channels actually stored (<= n_blocks * BN)
  long long sN, sH, sW;            // output element strides (class-local pixel steps)
  bf16* out;
  const float* bias;               // nullptr = none (fp32, indexed by absolute channel)
  int bias_n;                      // number of valid bias entries
  int act;                         // 0 none, 1 LeakyReLU(0.2), 2 tanh
  float* stats;                    // optional [N][Cout][2] fp32 (sum, sumsq) accumulated with atomics
  long long* prof;                 // optional [ctas][8] clock64 timestamps (development profiling)
};

// Geometry of the patch-resident implicit GEMM (conv_patch.cu): the halo'd input patch of a 16 x 8 pixel output
// tile is fetched ONCE per 64-channel chunk and every filter tap reads it through a shifted UMMA descriptor.
struct PatchArgs {
  int k;            // filter taps per side
  int flip;         // 1: weight tap index = k*k-1 - (py*k + px)  (input gradients walk the filter backwards)
  int ox, oy;       // patch origin relative to the tile's output origin, in the TMA view's coordinates
  int chunks;       // 64-channel chunks of the contraction
  int tap_stride;   // K distance between consecutive filter taps in the packed weights (= stored channels)
  int PH;           // patch rows (16 + k - 1)
  int nbox;         // TMA boxes per patch: 1 (whole patch, PW = 8 + k - 1 columns) or k column-shifted 8-wide boxes
  int box_bytes;    // bytes of one box
  int patch_bytes;  // bytes reserved per patch in shared memory (multiple of 1024)
  int row_step;     // descriptor start-address step per patch row    (16-byte units)
  int col_step;     // descriptor start-address step per patch column (16-byte units)
  int sbo;          // bytes between consecutive 8-pixel row groups of the M tile
  int base_offset;  // 1: put (start address >> 7) & 7 into the descriptor's base-offset field
  int b_stages;     // weight-tile ring depth
  int ka;           // channels per patch row: 64 (128-byte rows) or 16 (32-byte rows of the image-like tensors)
  int b_resident;   // 1: the ring holds the whole filter; it is loaded once and never recycled
  int num_items;    // M-direction work items (MT stacked tiles each); a CTA loops over blockIdx.x + i * gridDim.x
};

// One filter tap of the weight-gradient GEMM: offsets for both operands.
struct WTap {
  int16_t a_c, a_dx, a_par, a_dy;  // dY side (M = Cout)
  int16_t b_c, b_dx, b_par, b_dy;  // X side  (N = Cin)
  int32_t out_tap;                 // tap index in g[Cout][T][Cin]
  int32_t pad;
};
static_assert(sizeof(WTap) == 24, "WTap must be 24 bytes");

struct WgradArgs {
  const WTap* taps;
  int num_taps;
  int T;                   // taps per filter (kh*kw) = middle dim of g
  int Cout, Cin;           // valid extents of g
  int tiles_w, tiles_h, N; // pixel chunks (64 pixels = THk x TWk) per image, images
  int tw_shift;            // TWk = 1 << tw_shift, THk = 64 >> tw_shift
  int split_k;             // number of K splits (gridDim.z)
  float* g;                // [Cout][T][Cin] fp32, accumulated with atomics (must be zeroed)
  const int* row_map;      // optional: GEMM row m -> row of g (row length Cin, T ignored); < 0 = skip
  const int* col_map;      // optional: GEMM column -> column of g (< 0 = skip); forces the scalar epilogue
  int ncols;               // GEMM N extent (0: same as Cin)
  int a_virtual;           // 1: the dY-side tensor map is a virtual im2col over a row-expanded tensor: dims (64 channels, w,
                           //    row pair, h, n); 64-channel atom a of the GEMM's M extent = row pair a (rows h + 2a, h + 2a + 1)
  int b_virtual;           // the same for the X side (atoms of the GEMM's N extent)
  int trigger;             // 1: let the next kernel of the lane start launching once the main loop is done (PDL)
};

// Weight gradient of a stride-1 conv on CTA pairs (wgrad_pair.cu): a pair owns one filter row, 256 output channels and
// 128 input channels; K (8 x 8 pixel chunks) is split over pairs.
struct WgradPairArgs {
  int kh, kw, T;                 // filter extent, taps per filter (= kh * kw)
  int Cout, Cin;                 // extents of g[Cout][T][Cin] (Cout % 256 == 0, Cin % 128 == 0)
  int tiles_w, tiles_h, N;       // 8 x 8 pixel chunks per image, images
  int split_k;                   // K splits
  int n_units, cin_blocks;       // units = kh x cin_blocks x (Cout / 256); pair q works on unit q % n_units, split q / n_units
  int x_ox, x_oy;                // origin of the X patch of tap (0, 0) relative to the dY chunk origin, in the X view's coordinates
  int trigger;                   // 1: let the next kernel of the lane start launching once the main loop is done (PDL)
  float* g;                      // fp32, accumulated with atomics (must be zeroed)
};

// "Taps in N" kernel (conv_tapn.cu): stride-1 k x k conv from 64 channels to <= 4 (generator head; stem input gradient).
struct TapNArgs {
  int k;                  // filter taps per side (<= 8)
  int flip;               // 1: input gradient (filter rows walked backwards, taps gathered backwards)
  int ox, oy;             // patch origin relative to the tile's output origin, in the activation view's coordinates
  int tiles_w, tiles_h;   // 8-wide, 16-high output tiles per image
  int N, num_items;       // images, tiles in total (persistent CTAs loop over blockIdx.x + i * gridDim.x)
  int Ho, Wo;             // valid output extent
  int cout;               // real output channels (<= 4); the tensor stores 16
  long long sN, sH, sW;   // output element strides
  bf16* out;
  const float* bias;      // nullptr = none
  int bias_n;
  int act;                // Act enum: none or tanh
};

// Launch with the programmatic-dependent-launch attribute (see ptx.cuh: pdl_wait / pdl_launch_dependents).
// Every kernel launched through this helper MUST call pdl_wait() before touching global memory.
bool pdl_enabled();  // CGB_PDL=0 disables (conv_plan.cc)
bool pair_pdl_enabled();  // programmatic dependent launch for the CTA-pair (cluster) kernels: CGB_PAIR_PDL (conv_plan.cc)
// Dynamic shared memory a TMEM-hungry kernel requests so that no other TMEM-using CTA fits beside it on the SM
// (the smallest such CTA, a lite tap-table instantiation, needs 54 KB): 0 when CGB_EXCL_SMEM=0 (conv_plan.cc)
int tmem_exclusive_smem();
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CGB_CUDA(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...));
}

enum Act { kActNone = 0, kActLeaky = 1, kActTanh = 2, kActRelu = 3 };

}  // namespace cgb
