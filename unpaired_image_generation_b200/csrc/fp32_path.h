// fp32 VALIDATION MODE of the CycleGAN step (north_star: "1e-5 for an fp32 validation mode").
//
// Same engine, same recorded programs, lanes, buffers, halo / fold logic and schedule as the bf16 product path
// (engine.cc): only the element type of the activation tensors (TensorDesc::esz == 4) and the kernels behind
// each op change.  Every kernel here is a plain CUDA-core kernel that
//   * stores activations and activation gradients in fp32 (NHWC, same stored channel counts and halos),
//   * multiplies in fp32 and ACCUMULATES EVERY LONG SUM IN fp64 (convolution contractions, InstanceNorm statistics,
//     loss sums, bias gradients), so each stored value is the correctly rounded fp32 result of its layer, and
//   * is deterministic: no atomics, every reduction has a fixed order (per-thread strided partial sums in fp64, then
//     a fixed-order tree), weight gradients are written (not accumulated) into one buffer per pass and summed in a
//     fixed order at the end (cgb_engine::gslot).
// It exists to check the orchestration of the step (what the bf16 noise floor hides) against the fp32 stand-in
// (oracle/cyclegan_standin.py:312 forward_only, :352 backward_only, :374 train_step) at 1e-5; it is not a
// performance path and nothing in it runs in the default (bf16) mode.
#pragma once
#include "conv_plan.h"
#include "pointwise.h"

namespace cgb {
namespace f32 {

// Conv2d / ConvTranspose2d forward.  w: fp32 master weights [Cout][T][Cin]; bias may be null; act: Act enum.
// Channels >= Cout of y (stored padding) are written as zeros.
void conv_fprop(const ConvSpec& s, const TensorDesc& x, const float* w, const float* bias, int act, const TensorDesc& y,
                cudaStream_t st);
// Input gradient; for reflect convs dx is the padded-domain tensor (H + 2p, W + 2p, halo 0), as in plan_dgrad.
void conv_dgrad(const ConvSpec& s, const TensorDesc& dy, const float* w, const TensorDesc& dx, cudaStream_t st);
// Weight gradient, STORED (not accumulated) into g[Cout][T][Cin].
void conv_wgrad(const ConvSpec& s, const TensorDesc& x, const TensorDesc& dy, float* g, cudaStream_t st);

// InstanceNorm forward: statistics (written to stats as (mean, rstd)) + normalise + activation (+ residual) and the
// reflect halo of `out`, one kernel.
void in_forward(const TensorDesc& y, float2* stats, int act, const TensorDesc* residual, const TensorDesc& out,
                cudaStream_t st);
// InstanceNorm + activation backward (both reductions and the apply pass, one kernel).  stats = (mean, rstd).
void in_backward(const TensorDesc& y, const float2* stats, const GradSrc& g, int act, const TensorDesc* da_store,
                 const TensorDesc& dy, cudaStream_t st);

void nchw_to_nhwc(const float* src, int C, const TensorDesc& dst, cudaStream_t st);
void nhwc_to_nchw(const TensorDesc& src, int C, float* dst, cudaStream_t st);
void nhwc_to_u8hwc(const TensorDesc& src, int C, unsigned char* dst, cudaStream_t st);
void tanh_bwd(const TensorDesc& out, const TensorDesc* target, float l1_scale, const GradSrc& g, int C,
              const TensorDesc& dpre, float* loss_slot, cudaStream_t st);
void l1_loss(const TensorDesc& a, const TensorDesc& b, int C, float scale, float* loss_slot, cudaStream_t st);
void leaky_bwd(const TensorDesc& a, const TensorDesc& g, const TensorDesc& dpre, cudaStream_t st);
void mse_loss(const TensorDesc& logits, float target, float w, float* loss_slot, const TensorDesc* dlogits,
              cudaStream_t st);
// gbias[c] = sum over pixels (STORED), c < C
void bias_grad(const TensorDesc& dy, int C, float* gbias, cudaStream_t st);
// dst[i] = slots[0][i] + slots[1][i] (+ slots[2][i]), fixed order
void sum_slots(float* dst, const float* s0, const float* s1, const float* s2, long long n, cudaStream_t st);

}  // namespace f32
}  // namespace cgb
