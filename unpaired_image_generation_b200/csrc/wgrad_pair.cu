// Weight gradient of stride-1 convolutions on CTA pairs (cta_group::2), patch-resident (sm_100a).
//
//   g[cout][r * kw + c][cin] += sum over pixels (n, h, w) of dY[n, h, w, cout] * Xp[n, h + r, w + c, cin]
//
// wgrad_kernel (conv_tc.cu) gives every filter tap its own CTA: each tap-CTA streams its own copy of the dY tile and
// a shifted copy of the X tile, 48 KB of operands per four MMAs (88 B/clk per SM), and is bound by the L2 -> SM
// bandwidth (measured 780-850 TFLOP/s at batch 8, 18 CTAs at batch 1).  Here a CTA PAIR owns one filter ROW r, 256
// output channels (128 per CTA) and 128 input channels (64 staged per CTA):
//   * per 8 x 8 pixel chunk the pair fetches the dY tile once and ONE X patch of 8 x (8 + kw - 1) pixels; the kw taps
//     of the row read that patch in place through MN-major UMMA descriptors whose start address is shifted by c
//     pixels (128-byte rows, SWIZZLE_128B, stride between 8-pixel groups = the patch pitch: the same absolute-address
//     swizzle property the patch-resident forward kernel relies on, conv_patch.cu);
//   * the kw accumulators (kw x 128 TMEM columns) are filled by M = 256, N = 128, K = 16 MMAs issued by the leader;
//   * operands per chunk and CTA: 16 KB of dY + 10 KB of X for 3 x 4 MMAs = 34 B/clk per SM instead of 88.
// K (pixels) is split over pairs; partial tiles are reduced into the fp32 master gradient with coalesced vector
// atomics (the same transposing epilogue as wgrad_kernel).
//
// Stand-in counterpart: the weight gradient of F.conv2d (stride 1) in oracle/cyclegan_standin.py (ResnetBlock convs,
// discriminator conv3).
//
// Warp roles (192 threads): warp 0 = TMA producer (both CTAs), warp 1 = TMEM allocator (both) + MMA issuer (leader),
// warps 2..5 = epilogue (both CTAs: each drains its own 128 output-channel rows).
#include "common.h"
#include "conv_tc.h"
#include "ptx.cuh"

namespace cgb {

using namespace ptx;

namespace {
constexpr int kDyBytes = 2 * 64 * 128;  // two 64-channel atoms of 64 pixels x 128 bytes
constexpr int kWpStages = 6;
constexpr int kWpSmemMax = 232448;
}  // namespace

__global__ void __launch_bounds__(192, 1)
wgrad_pair_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                  const WgradPairArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int PW = 8 + a.kw - 1;             // patch pitch in pixels
  const int patch_bytes = PW * 8 * 128;    // 8 rows; a multiple of 1024
  const int stage_bytes = kDyBytes + patch_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kWpStages * stage_bytes);
  uint64_t* empty_bar = full_bar + kWpStages;
  uint64_t* tmem_full_bar = empty_bar + kWpStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int q = blockIdx.x >> 1;
  const int unit = q % a.n_units, split = q / a.n_units;
  const int r = unit % a.kh;                          // filter row
  const int nb = (unit / a.kh) % a.cin_blocks;        // block of 128 input channels
  const int mb = unit / (a.kh * a.cin_blocks);        // block of 256 output channels
  const int total = a.tiles_w * a.tiles_h * a.N;
  const int per = (total + a.split_k - 1) / a.split_k;
  const int cbeg = split * per;
  const int kcnt = max(0, min(total, cbeg + per) - cbeg);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmDY);
    prefetch_tmap(&tmX);
    for (int s = 0; s < kWpStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  } else if (warp == 1) {
    tmem_alloc_pair(tmem_ptr, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; bytes are credited to the leader's full barrier) =====================
    int s = 0;
    uint32_t ph = 1;
    for (int i = 0; i < kcnt; ++i) {
      mbar_wait(&empty_bar[s], ph);
      if (elect_one()) {
        int c = cbeg + i;
        const int tw = c % a.tiles_w;
        c /= a.tiles_w;
        const int th = c % a.tiles_h;
        const int n = c / a.tiles_h;
        const int w0 = tw * 8, h0 = th * 8;
        uint8_t* sa = smem + s * stage_bytes;
        uint8_t* sb = sa + kDyBytes;
        if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2u * (uint32_t)stage_bytes);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          tma_load_5d_pair(sa + j * (kDyBytes / 2), &tmDY, &full_bar[s], mb * 256 + rank * 128 + j * 64, w0, 0, h0, n);
        tma_load_5d_pair(sb, &tmX, &full_bar[s], nb * 128 + rank * 64, w0 + a.x_ox, 0, h0 + r + a.x_oy, n);
      }
      __syncwarp();
      if (++s == kWpStages) {
        s = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===================== MMA issuer (leader): MN-major operands, K = pixels =====================
      constexpr uint32_t idesc = make_idesc_bf16(256, 128, 1, 1);
      constexpr uint32_t desc_hi_a = smem_desc_hi(1024, 2);              // 8-pixel groups of the dY tile are 1 KB apart
      const uint32_t desc_hi_b = smem_desc_hi((uint32_t)(PW * 128), 2);  // ... of the patch one patch row apart
      constexpr uint32_t lbo_lo = ((kDyBytes / 2) >> 4) << 16;           // 64-channel atoms of dY are 8 KB apart
      const uint32_t lo0 = (smem_u32(smem) & 0x3FFFFu) >> 4;
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < kcnt; ++i) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = (lo0 + (uint32_t)(s * stage_bytes >> 4)) | lbo_lo;
          const uint32_t b_lo0 = lo0 + (uint32_t)((s * stage_bytes + kDyBytes) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // UMMA_K = 16 pixels = tile rows 2k, 2k + 1
            const uint32_t b_row = b_lo0 + (uint32_t)(2 * k * PW * 8);  // 8 x 16 bytes per patch pixel
            for (int c = 0; c < a.kw; ++c)
              umma_bf16_pair(tmem_base + c * 128, smem_desc_join(a_lo + k * 128, desc_hi_a),
                             smem_desc_join(b_row + c * 8, desc_hi_b) | ((uint64_t)lbo_lo), idesc, (i | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[s]);  // frees the slot in both CTAs once these MMAs retire
          if (i == kcnt - 1) umma_commit_pair(tmem_full_bar);
        }
        __syncwarp();
        if (++s == kWpStages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (kcnt > 0) {
    // ===================== epilogue (both CTAs): coalesced fp32 reductions into g =====================
    const int qw = warp & 3;
    const int co = mb * 256 + rank * 128 + qw * 32 + lane;
    const bool valid = co < a.Cout;
    mbar_wait_relaxed(tmem_full_bar, 0);
    tc_fence_after();
    if (a.trigger) pdl_launch_dependents();
    // a thread owns one gradient row (cout): the 32 x 32 fp32 block of each column chunk is transposed through shared
    // memory (the pipeline buffers are idle now) so that a warp instruction covers 4 rows x 128 contiguous bytes
    float* tstage = reinterpret_cast<float*>(smem) + qw * (32 * 33);
    for (int c = 0; c < a.kw; ++c) {
      const unsigned long long grow_u =
          reinterpret_cast<unsigned long long>(a.g + ((long long)co * a.T + r * a.kw + c) * a.Cin + nb * 128);
#pragma unroll 1
      for (int cc = 0; cc < 128; cc += 32) {
        uint32_t rr[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(qw * 32) << 16) + c * 128 + cc, rr);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) tstage[lane * 33 + j] = __uint_as_float(rr[j]);
        __syncwarp();
        const int sub = lane >> 3, piece = lane & 7;  // 4 rows per instruction, 8 x 16 bytes per row
#pragma unroll
        for (int r0 = 0; r0 < 32; r0 += 4) {
          const int row = r0 + sub;
          const unsigned long long gp = __shfl_sync(0xffffffffu, grow_u, row);
          const int ok = __shfl_sync(0xffffffffu, valid ? 1 : 0, row);
          const float* sp = tstage + row * 33 + piece * 4;
          if (ok) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gp + (unsigned long long)(cc + piece * 4) * 4),
                         "f"(sp[0]), "f"(sp[1]), "f"(sp[2]), "f"(sp[3])
                         : "memory");
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA may exit or free TMEM while the other can still read its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

void launch_wgrad_pair(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradPairArgs& a, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    CGB_CUDA(cudaFuncSetAttribute(wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWpSmemMax));
    configured = true;
  }
  CGB_CHECK(a.kw >= 1 && a.kw <= 4, "wgrad pairs: at most 4 taps per filter row (4 x 128 TMEM columns)");
  CGB_CHECK(a.Cout % 256 == 0 && a.Cin % 128 == 0, "wgrad pairs: Cout % 256 == 0 and Cin % 128 == 0 required");
  const int stage = kDyBytes + (8 + a.kw - 1) * 8 * 128;
  int smem = 1024 + kWpStages * stage + 256;
  CGB_CHECK(smem <= kWpSmemMax, "wgrad pairs: shared memory budget exceeded");
  // This kernel allocates all 512 TMEM columns: no other TMEM-using CTA may share its SM (a co-resident CTA of the
  // low-shared-memory tap-table instantiations would make tcgen05.alloc block).  Request enough shared memory to own the SM.
  if (tmem_exclusive_smem() > smem) smem = tmem_exclusive_smem();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2u * (unsigned)(a.n_units * a.split_k));
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pair_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  CGB_CUDA(cudaLaunchKernelEx(&cfg, wgrad_pair_kernel, tmDY, tmX, a));
}

}  // namespace cgb
