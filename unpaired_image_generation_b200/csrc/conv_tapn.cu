// "Taps in N": stride-1 k x k convolutions (k <= 8) from 64 channels to at most 4 (the 7 x 7 generator head, and the
// input gradient of the 7 x 7 stem, which has the same shape), persistent tcgen05 kernel for sm_100a.
//
// The patch-resident kernel (conv_patch.cu) runs these layers with N = 16 (3 real output channels): 49 taps x 4
// K-steps = 196 MMAs of M128 x N16 x K16 per 128-pixel tile, each bound by the 4 KB A-operand read (~48 cycles), i.e.
// 143 us for the head at batch 8 on a layer whose HBM time is 13 us.  Here the k horizontal taps of a filter row
// become the N dimension:
//   D'[(r, c)][tx * 4 + co] = sum over ty, ci of patch[r + ty][c][ci] * W[co][ty][tx][ci]     (c: patch column, 0 .. 15)
//   out[r][w][co]           = sum over tx of D'[(r, w + tx)][tx * 4 + co]                      (w: output column, 0 .. 7)
// One 16 x 8 output tile takes 2 column groups x k filter rows x 4 K-steps = 56 MMAs of N = 32 instead of 196 of
// N = 16; the shift-and-add over tx is done by the epilogue through shared memory.  The B operand needs no special
// weight pack: filter row ty of Wf[co][(ty * k + tx) * 64 + ci] read through a 4-D tensor map (ci, co < 4, tx < 8, ty)
// lands in shared memory as rows tx * 4 + co.  `flip` (input gradients) walks the filter rows backwards in the MMA
// loop and the taps backwards in the epilogue gather.
//
// Stand-in counterpart: F.conv2d of the generator head (c7s1-3 + tanh) and the input gradient of the stem (c7s1-64)
// in oracle/cyclegan_standin.py (Generator.forward).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#include "common.h"
#include "conv_tc.h"
#include "ptx.cuh"

namespace cgb {

using namespace ptx;

namespace {
constexpr int kTnBBytes = 8 * 4096;            // up to 8 filter rows x (32 rows x 128 bytes)
constexpr int kTnStageFloats = 16 * 16 * 32;   // D' of one tile: 16 rows x 16 patch columns x 32 columns
constexpr int kTnSmemMax = 232448;
}  // namespace

__global__ void __launch_bounds__(192, 1)
conv_tapn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TapNArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int PH = 16 + a.k - 1;
  const int patch_bytes = PH * 16 * 128;  // 16 patch columns of 128 bytes per row: a multiple of 1024
  uint8_t* sB = smem;
  uint8_t* patches = sB + kTnBBytes;
  float* stage = reinterpret_cast<float*>(patches + 2 * patch_bytes);
  uint64_t* b_full = reinterpret_cast<uint64_t*>(stage + kTnStageFloats);
  uint64_t* a_full = b_full + 1;        // [2]
  uint64_t* a_empty = a_full + 2;       // [2]
  uint64_t* tmem_full = a_empty + 2;    // [2]
  uint64_t* tmem_empty = tmem_full + 2; // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int items = a.num_items;
  const int tiles_per_img = a.tiles_w * a.tiles_h;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    mbar_init(b_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 128);
    }
    fence_mbar_init();
  } else if (warp == 1) {
    tmem_alloc(tmem_ptr, 128);  // two accumulator sets of 2 column groups x 32 columns
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp == 0) {
    // ===================== producer: the filter once, then one patch per tile =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(b_full, (uint32_t)(a.k * 4096));
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
              smem_u32(sB)),
          "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(b_full)), "r"(0), "r"(0), "r"(0), "r"(0)
          : "memory");
    }
    __syncwarp();
    int buf = 0;
    uint32_t ph = 1;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int n = item / tiles_per_img, r = item - n * tiles_per_img;
      const int th = r / a.tiles_w, tw = r - th * a.tiles_w;
      mbar_wait(&a_empty[buf], ph);
      if (elect_one()) {
        mbar_arrive_expect_tx(&a_full[buf], (uint32_t)patch_bytes);
        tma_load_5d(patches + buf * patch_bytes, &tmA, &a_full[buf], 0, tw * 8 + a.ox, 0, th * 16 + a.oy, n);
      }
      __syncwarp();
      if (buf == 1) ph ^= 1;
      buf ^= 1;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(128, 32, 0, 0);
    constexpr uint32_t desc_hi_a = smem_desc_hi(2048, 2);  // 8-pixel groups (tile rows) are one patch row apart
    constexpr uint32_t desc_hi_b = smem_desc_hi(1024, 2);
    const uint32_t lo_b = smem_u32(sB) >> 4;
    const uint32_t lo_p = smem_u32(patches) >> 4;
    mbar_wait(b_full, 0);
    int buf = 0, acc = 0;
    uint32_t a_ph = 0, e_ph = 1;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      mbar_wait(&tmem_empty[acc], e_ph);
      mbar_wait(&a_full[buf], a_ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t pbase = lo_p + (uint32_t)(buf * patch_bytes >> 4);
        for (int ty = 0; ty < a.k; ++ty) {
          const uint32_t brow = lo_b + (uint32_t)((a.flip ? a.k - 1 - ty : ty) * (4096 >> 4));
#pragma unroll
          for (int j = 0; j < 2; ++j) {  // patch columns 8j .. 8j + 7
            const uint32_t arow = pbase + (uint32_t)((ty * 2048 + j * 1024) >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16(tmem_base + acc * 64 + j * 32, smem_desc_join(arow + 2 * kk, desc_hi_a),
                        smem_desc_join(brow + 2 * kk, desc_hi_b), idesc, (ty | kk) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&a_empty[buf]);
        umma_commit(&tmem_full[acc]);
      }
      __syncwarp();
      if (buf == 1) a_ph ^= 1;
      buf ^= 1;
      if (acc == 1) e_ph ^= 1;
      acc ^= 1;
    }
  } else {
    // ===================== epilogue (warps 2..5): shift-and-add over the taps of a row, bias, activation =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;       // TMEM lane = accumulator row = (tile row, column inside the group)
    const int tr = m >> 3, tc = m & 7;
    float bias[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.bias != nullptr)
      for (int c = 0; c < 4; ++c) bias[c] = c < a.bias_n ? __ldg(a.bias + c) : 0.f;
    int acc = 0;
    uint32_t f_ph = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int n = item / tiles_per_img, r = item - n * tiles_per_img;
      const int th = r / a.tiles_w, tw = r - th * a.tiles_w;
      mbar_wait_relaxed(&tmem_full[acc], f_ph);
      tc_fence_after();
      if (item + (int)gridDim.x >= items) pdl_launch_dependents();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t rr[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 64 + j * 32, rr);
        tmem_ld_wait();
        // D'[(tr, 8j + tc)][0 .. 31] -> stage[(tr * 16 + col)][32], 16-byte pieces XOR-swizzled by the column
        const int col = 8 * j + tc;
        float4* row = reinterpret_cast<float4*>(stage + (tr * 16 + col) * 32);
#pragma unroll
        for (int p = 0; p < 8; ++p)
          row[p ^ (col & 7)] = make_float4(__uint_as_float(rr[4 * p]), __uint_as_float(rr[4 * p + 1]),
                                           __uint_as_float(rr[4 * p + 2]), __uint_as_float(rr[4 * p + 3]));
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);  // this thread's TMEM reads are done: the MMA warp may refill the set
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // output pixel (tr, tc): sum over tx of D'[(tr, tc + tx')][tx], tx' = tx (forward) or k - 1 - tx (input gradients)
      float4 s = make_float4(bias[0], bias[1], bias[2], bias[3]);
      for (int tx = 0; tx < a.k; ++tx) {
        const int col = tc + (a.flip ? a.k - 1 - tx : tx);
        const float4 v = reinterpret_cast<const float4*>(stage + (tr * 16 + col) * 32)[tx ^ (col & 7)];
        s.x += v.x;
        s.y += v.y;
        s.z += v.z;
        s.w += v.w;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the stage is rewritten by the next tile
      if (a.act == kActTanh) {
        s.x = tanhf(s.x);
        s.y = tanhf(s.y);
        s.z = tanhf(s.z);
        s.w = tanhf(s.w);
      }
      const int ho = th * 16 + tr, wo = tw * 8 + tc;
      if (ho < a.Ho && wo < a.Wo) {
        uint4 lo, hi = make_uint4(0u, 0u, 0u, 0u);
        lo.x = pack_bf16x2(a.cout > 0 ? s.x : 0.f, a.cout > 1 ? s.y : 0.f);
        lo.y = pack_bf16x2(a.cout > 2 ? s.z : 0.f, a.cout > 3 ? s.w : 0.f);
        lo.z = lo.w = 0u;
        uint4* o = reinterpret_cast<uint4*>(a.out + (long long)n * a.sN + (long long)ho * a.sH + (long long)wo * a.sW);
        o[0] = lo;  // 16 stored channels: the real ones, then zeros
        o[1] = hi;
      }
      if (acc == 1) f_ph ^= 1;
      acc ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

void launch_conv_tapn(const CUtensorMap& tmA, const CUtensorMap& tmB, const TapNArgs& a, int num_ctas, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    CGB_CUDA(cudaFuncSetAttribute(conv_tapn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTnSmemMax));
    configured = true;
  }
  CGB_CHECK(a.k >= 2 && a.k <= 8, "taps-in-N conv: 2 <= k <= 8");
  CGB_CHECK(a.cout >= 1 && a.cout <= 4, "taps-in-N conv: at most 4 output channels");
  int smem = 1024 + kTnBBytes + 2 * (16 + a.k - 1) * 16 * 128 + kTnStageFloats * 4 + 256;
  CGB_CHECK(smem <= kTnSmemMax, "taps-in-N conv: shared memory budget exceeded");
  if (tmem_exclusive_smem() > smem) smem = tmem_exclusive_smem();  // own the SM (see wgrad_pair.cu)
  launch_pdl(conv_tapn_kernel, dim3(num_ctas), dim3(192), (size_t)smem, stream, tmA, tmB, a);
}

}  // namespace cgb
