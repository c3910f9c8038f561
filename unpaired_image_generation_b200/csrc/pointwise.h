// HBM-bound kernels of the CycleGAN step: layout conversion, InstanceNorm statistics / apply /
// backward (with reflection-halo writing and halo-gradient folding), activation backward,
// LSGAN / L1 losses with their seed gradients, bias gradients, weight packing and Adam.
// Stand-in counterparts (oracle/cyclegan_standin.py): _inorm, F.pad(mode="reflect"), F.relu,
// F.leaky_relu, torch.tanh, F.l1_loss, CycleGANTrainer._mse_to, torch.optim.Adam.
#pragma once
#include "common.h"
#include "conv_plan.h"

namespace cgb {

// ---- layout ------------------------------------------------------------------------------
// fp32 NCHW [N][C][H][W] -> bf16 NHWC (dst.C stored channels, zero padded) interior + reflect halo
void nchw_to_nhwc(const float* src, int C, const TensorDesc& dst, cudaStream_t st);
// bf16 NHWC interior -> fp32 NCHW (first C channels)
void nhwc_to_nchw(const TensorDesc& src, int C, float* dst, cudaStream_t st);
// mirror the interior into the halo (reflection padding) in place
void fill_reflect_halo(const TensorDesc& t, cudaStream_t st);

// ---- InstanceNorm --------------------------------------------------------------------------
// stats[n][c] = (sum, sumsq) over H*W of the bf16 values; buffer must be zero on entry.
void in_stats(const TensorDesc& y, float2* stats, cudaStream_t st);
// out(interior + reflect halo) = act(norm(y)) [+ residual.interior]
void in_apply(const TensorDesc& y, const float2* stats, int act, const TensorDesc* residual, const TensorDesc& out,
              cudaStream_t st);

// Gradient w.r.t. a layer's post-activation output, assembled on the fly from up to two sources:
//   g1: plain tensor (interior view), g2: gradient on the reflect-padded domain (H+2p, W+2p) whose
//   mirrored border is folded back onto the interior.
struct GradSrc {
  const TensorDesc* g1 = nullptr;
  const TensorDesc* g2 = nullptr;  // dims (H + 2*fold, W + 2*fold), halo 0
  int fold = 0;
};
// bstats[n][c] += (sum dz, sum dz*xhat); optionally stores the assembled (bf16-rounded) gradient.
void in_bwd_reduce(const TensorDesc& y, const float2* stats, const GradSrc& g, int act, const TensorDesc* da_out,
                   float2* bstats, cudaStream_t st);
// Both steps in ONE kernel (thread-block clusters; the map's per-image slice of 32 or 16 channels stays in the
// cluster's shared memory): same results as in_bwd_reduce + in_bwd_apply but deterministic (no atomics).  Returns false
// without launching anything when the map is too large for it (or CGB_IN_FUSED=0): the caller then uses the two kernels.
bool in_bwd_fused(const TensorDesc& y, const float2* stats, const GradSrc& g, int act, const TensorDesc* da_out,
                  const TensorDesc& dy, cudaStream_t st);
bool in_bwd_fused_supported(const TensorDesc& y);
// dy = rstd * (dz - mean(dz) - xhat * mean(dz * xhat))
void in_bwd_apply(const TensorDesc& y, const float2* stats, const float2* bstats, const GradSrc& g, int act,
                  const TensorDesc& dy, cudaStream_t st);

// ---- pointwise activation backward ----------------------------------------------------------
// generator head: dpre = (l1_scale * sign(out - target) + g1 + fold(g2)) * (1 - out^2); also
// accumulates loss_slot += l1_scale * sum|out - target| (only the first `C` channels count).
void tanh_bwd(const TensorDesc& out, const TensorDesc* target, float l1_scale, const GradSrc& g, int C,
              const TensorDesc& dpre, float* loss_slot, cudaStream_t st);
// L1 loss value only (forward-only paths): loss_slot += scale * sum|a - b| over first C channels
void l1_loss(const TensorDesc& a, const TensorDesc& b, int C, float scale, float* loss_slot, cudaStream_t st);
// discriminator conv0: dpre = g * (a > 0 ? 1 : 0.2)
void leaky_bwd(const TensorDesc& a, const TensorDesc& g, const TensorDesc& dpre, cudaStream_t st);
// LSGAN: loss_slot += w * mean((p - target)^2); dlogits(ch 0) = 2 w (p - target) / numel (may be null)
void mse_loss(const TensorDesc& logits, float target, float w, float* loss_slot, const TensorDesc* dlogits,
              cudaStream_t st);
// gbias[c] += sum over pixels of dy[..][c], c < C
void bias_grad(const TensorDesc& dy, int C, float* gbias, cudaStream_t st);

// ---- weights ---------------------------------------------------------------------------------
struct PackEntry {
  long long src_off;  // offset (elements) of the fp32 master weight [Cout][T][Cin] in the flat buffer
  long long wf_off;   // offset (elements) into the bf16 pack arena of Wf [CoutP][T][CinS]
  long long wt_off;   // offset of Wt [CinP][T][CoutS]
  long long wx_off;   // >= 0: extra pack Wx[Cout][wx_pitch] with column t*4 + ci (GEMM over an im2col4 matrix)
  int wx_pitch;
  int Cout, Cin, T, CinS, CoutS;
  int pad;
};
// one launch re-packs every layer of a parameter group (after Adam)
void pack_weights(const float* master, const PackEntry* entries_dev, int n_entries, int max_elems, bf16* arena,
                  cudaStream_t st);

// torch.optim.Adam semantics (no weight decay, no amsgrad); g is multiplied by grad_scale first.
// The step counter and the bias-correction scalars live in device memory (state: int step;
// hyper: {lr / (1 - beta1^t), 1 / sqrt(1 - beta2^t)}) so the launch is CUDA-graph replayable.
void adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
               float eps, int* step_dev, float* hyper_dev, float grad_scale, cudaStream_t st);

// the same update on a sub-range (data parallel: each gradient bucket is stepped right after its all-reduce);
// advance_step: first range of this optimiser step (increments the step counter, refreshes the bias corrections)
void adam_range(float* p, const float* g, float* m, float* v, long long n, float beta1, float beta2, float eps,
                int* step_dev, float* hyper_dev, float grad_scale, bool advance_step, cudaStream_t st);

// Image history pool exchange: d_in[n] = dec[n].ret >= 0 ? pool[ret] : fake[n]; pool[dec[n].store] = fake[n]
// (batch order; dec = [N][2] ints (store, ret) on the device, -1 = none)
void pool_exchange(const TensorDesc& fake, const TensorDesc& pool, const int* dec, const TensorDesc& d_in, cudaStream_t st);
// one-element store on the stream (graph- and stream-ordered scalar updates: the learning rate)
void set_device_float(float* dst, float value, cudaStream_t st);
// bf16 NHWC image -> uint8 interleaved [N][H][W][C], u8 = clamp(rint((x + 1) * 127.5), 0, 255)
void nhwc_to_u8hwc(const TensorDesc& src, int C, unsigned char* dst, cudaStream_t st);
// uint8 interleaved RGB [N][H][W][3] -> fp32 planar [N][3][H][W], x = u8 / 127.5 - 1
void u8hwc_to_nchw(const unsigned char* src, int N, int H, int W, float* dst, cudaStream_t st);

// ---- explicit im2col of a 16-stored-channel tensor with <= 4 real channels (the 3- and 1-channel sides of
// stem / head / conv0 / conv4), so that those layers become plain tensor-core GEMMs:
//   dst[n][h][w][t*4 + c] = src[n][h*stride + sgn*r + off][w*stride + sgn*s + off][c],  t = r*k + s, c < 4
// (zero outside the valid region; with use_halo the reflect halo of src counts as valid).  dst has
// dst.C >= 4*k*k stored channels; columns >= 4*k*k are never written (they must be zero-initialised once).
void im2col4(const TensorDesc& src, int k, int stride, int sgn, int off, bool use_halo, const TensorDesc& dst,
             cudaStream_t st);

// Row expansion of a 16-stored-channel tensor with <= 4 real channels: the k horizontal taps of a row AND of the row
// below it,
//   dst[n][hh][w][half*32 + s*4 + c] = src[n][hh + half + (sgn > 0 ? off : off - (k - 1))][w + sgn*s + off][c]
// for half < 2, s < k, c < 4 (zero outside src -- with use_halo the reflect halo of src counts as inside; dst.C == 64,
// columns with s >= k are written as zeros).  The k vertical taps are NOT materialised: a tensor map whose row-pair
// dimension (stride two rows) overlaps its height dimension reads rows hh, hh + 2, hh + 4, hh + 6 of dst as the four
// 64-column atoms of the im2col matrix (small_wgrad.cc): 4x fewer bytes than im2col4 for the 7x7 layers.
void expand_rows4(const TensorDesc& src, int k, int sgn, int off, bool use_halo, const TensorDesc& dst, cudaStream_t st);

}  // namespace cgb
