// tcgen05 / TMEM / TMA implicit-GEMM convolution kernels for sm_100a.
//
//  * igemm_conv_kernel : Y[pixel][cout] = sum_k A[pixel][k] * W[cout][k]
//      A tiles are 5-D TMA boxes of the NHWC activation tensor, one (filter tap, channel
//      chunk) per K iteration, described by a KIter table (common.h).  Stride-1 convs, stride-2
//      convs (through a parity-plane 5-D view), transposed convs and stride-2 dgrads (as four
//      output-parity classes), and stride-1 dgrads all run through this one kernel.
//      Stand-in counterpart: F.conv2d / F.conv_transpose2d and their input gradients in
//      oracle/cyclegan_standin.py (Generator.forward / Discriminator.forward).
//  * wgrad_kernel : g[cout][tap][cin] += sum_pixels dY[pixel][cout] * X[pixel+tap][cin]
//      both operands MN-major (channels contiguous), K = pixels, split-K with fp32 atomics.
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator +
// MMA issuer (one lane), warps 2..5 = epilogue (TMEM -> registers -> global).
#include "common.h"
#include "conv_tc.h"
#include "ptx.cuh"

namespace cgb {

using namespace ptx;

template <int BN, int BK, int STAGES>
struct IgemmCfg {
  static constexpr int kSwizzle = BK * 2;  // bytes per smem row == swizzle span
  static constexpr int kABytes = 128 * kSwizzle;
  static constexpr int kBBytesTx = BN * kSwizzle;
  static constexpr int kBBytes = (kBBytesTx + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  static constexpr int kSmemBytes = STAGES * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, int BK, int STAGES>
__global__ void __launch_bounds__(192, 1)
igemm_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const IgemmArgs args) {
  using Cfg = IgemmCfg<BN, BK, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cls = blockIdx.z;
  const int nblk = blockIdx.y;
  const int kbeg = args.k_begin[cls];
  const int kcnt = args.k_count[cls];

  // tile -> (image, tile row, tile col)
  const int TW = 1 << args.tw_shift;
  const int TH = 128 >> args.tw_shift;
  int t = blockIdx.x;
  const int tw = t % args.tiles_w;
  t /= args.tiles_w;
  const int th = t % args.tiles_h;
  const int n = t / args.tiles_h;
  const int wo0 = tw * TW, ho0 = th * TH;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  } else if (warp == 1) {
    tmem_alloc(tmem_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      for (int i = 0; i < kcnt; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const KIter kt = args.kiters[kbeg + i];
        uint8_t* sa = smem + s * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        mbar_arrive_expect_tx(&full_bar[s], Cfg::kABytes + Cfg::kBBytesTx);
        tma_load_5d(sa, &tmA, &full_bar[s], kt.a_c, wo0 + kt.a_dx, kt.a_par, ho0 + kt.a_dy, n);
        tma_load_2d(sb, &tmB, &full_bar[s], kt.b_k, nblk * BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = make_idesc_bf16(128, BN < 16 ? 16 : BN, 0, 0);
      constexpr uint32_t lt = swizzle_layout_type(Cfg::kSwizzle);
      for (int i = 0; i < kcnt; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
        const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t ad = make_smem_desc(a_addr + k * 32, 0, 8 * Cfg::kSwizzle, lt);
          const uint64_t bd = make_smem_desc(b_addr + k * 32, 0, 8 * Cfg::kSwizzle, lt);
          umma_bf16(tmem_base, ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs retire
      }
      umma_commit(tmem_full_bar);
    }
  } else if (kcnt > 0) {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int ho = ho0 + (row >> args.tw_shift);
    const int wo = wo0 + (row & (TW - 1));
    const bool valid = (ho < args.Ho) && (wo < args.Wo);
    bf16* orow = args.out + args.out_off[cls] + (long long)n * args.sN + (long long)ho * args.sH +
                 (long long)wo * args.sW;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    constexpr int CH = BN >= 32 ? 32 : 16;
#pragma unroll 1
    for (int c = 0; c < BN; c += CH) {
      float v[CH];
      {
        uint32_t r[CH];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c;
        if constexpr (CH == 32) {
          tmem_ld32(taddr, r);
        } else {
          tmem_ld16(taddr, r);
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]);
      }
      const int co0 = nblk * BN + c;
      if (args.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          if (co0 + j < args.bias_n) v[j] += __ldg(args.bias + co0 + j);
        }
      }
      if (args.act == kActLeaky) {
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
      } else if (args.act == kActTanh) {
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = tanhf(v[j]);
      } else if (args.act == kActRelu) {
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (valid) {
#pragma unroll
        for (int j = 0; j < CH; j += 8) {
          if (co0 + j + 8 <= args.Cout) {
            uint4 pk;
            pk.x = pack_bf16x2(v[j + 0], v[j + 1]);
            pk.y = pack_bf16x2(v[j + 2], v[j + 3]);
            pk.z = pack_bf16x2(v[j + 4], v[j + 5]);
            pk.w = pack_bf16x2(v[j + 6], v[j + 7]);
            *reinterpret_cast<uint4*>(orow + co0 + j) = pk;
          } else {
            for (int jj = j; jj < j + 8; ++jj) {
              if (co0 + jj < args.Cout) orow[co0 + jj] = __float2bfloat16_rn(v[jj]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// Weight gradient: D[cout 128][cin BNW] for one tap, K = pixels (64 per stage), MN-major operands.
// ------------------------------------------------------------------------------------------
template <int BNW, int STAGES>
struct WgradCfg {
  static constexpr int kAtomBytes = 64 * 128;       // one TMA box: 64 pixels x 64 channels (128 B)
  static constexpr int kABytes = 2 * kAtomBytes;    // M = 128 channels of dY
  static constexpr int kBBytes = (BNW / 64) * kAtomBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = STAGES * kStageBytes + 1024 + 256;
};

template <int BNW, int STAGES>
__global__ void __launch_bounds__(192, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
             const WgradArgs args) {
  using Cfg = WgradCfg<BNW, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_blocks = (args.Cin + BNW - 1) / BNW;
  const int mblk = blockIdx.x / n_blocks;
  const int nblk = blockIdx.x % n_blocks;
  const WTap tap = args.taps[blockIdx.y];

  const int chunks_per_img = args.tiles_w * args.tiles_h;
  const int total = chunks_per_img * args.N;
  const int per = (total + args.split_k - 1) / args.split_k;
  const int cbeg = blockIdx.z * per;
  const int cend = min(total, cbeg + per);
  const int kcnt = max(0, cend - cbeg);
  const int TWk = 1 << args.tw_shift;
  const int THk = 64 >> args.tw_shift;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmDY);
    prefetch_tmap(&tmX);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  } else if (warp == 1) {
    tmem_alloc(tmem_ptr, BNW < 32 ? 32 : BNW);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < kcnt; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        int c = cbeg + i;
        const int tw = c % args.tiles_w;
        c /= args.tiles_w;
        const int th = c % args.tiles_h;
        const int n = c / args.tiles_h;
        const int w0 = tw * TWk, h0 = th * THk;
        uint8_t* sa = smem + s * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          tma_load_5d(sa + j * Cfg::kAtomBytes, &tmDY, &full_bar[s], tap.a_c + mblk * 128 + j * 64, w0 + tap.a_dx,
                      tap.a_par, h0 + tap.a_dy, n);
#pragma unroll
        for (int j = 0; j < BNW / 64; ++j)
          tma_load_5d(sb + j * Cfg::kAtomBytes, &tmX, &full_bar[s], tap.b_c + nblk * BNW + j * 64, w0 + tap.b_dx,
                      tap.b_par, h0 + tap.b_dy, n);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BNW, 1, 1);
      for (int i = 0; i < kcnt; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
        const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 64 pixels per stage, UMMA_K = 16 pixels = 16 rows of 128 B
          const uint64_t ad = make_smem_desc(a_addr + k * 2048, Cfg::kAtomBytes, 1024, 2);
          const uint64_t bd = make_smem_desc(b_addr + k * 2048, Cfg::kAtomBytes, 1024, 2);
          umma_bf16(tmem_base, ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      if (kcnt > 0) umma_commit(tmem_full_bar);
    }
  } else if (kcnt > 0) {
    const int q = warp & 3;
    const int co = mblk * 128 + q * 32 + lane;
    bool valid = co < args.Cout;
    long long row = (long long)co * args.T + tap.out_tap;
    if (args.row_map != nullptr) {
      const int rm = valid ? __ldg(args.row_map + co) : -1;
      valid = rm >= 0;
      row = rm;
    }
    float* grow = args.g + row * args.Cin;
    // rows are 16-byte aligned when Cin % 4 == 0: use 4-wide vector reductions
    const bool vec_ok = (args.Cin & 3) == 0;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BNW; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      tmem_ld_wait();
      if (valid) {
        const int ci0 = nblk * BNW + c;
        if (vec_ok && ci0 + 32 <= args.Cin) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(grow + ci0 + j),
                         "f"(__uint_as_float(r[j])), "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])),
                         "f"(__uint_as_float(r[j + 3]))
                         : "memory");
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (ci0 + j < args.Cin) atomicAdd(grow + ci0 + j, __uint_as_float(r[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BNW < 32 ? 32 : BNW);
  }
}

// ------------------------------------------------------------------------------------------
// Host launchers
// ------------------------------------------------------------------------------------------
template <int BN, int BK, int STAGES>
static void launch_igemm_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const IgemmArgs& args, dim3 grid,
                           cudaStream_t stream) {
  using Cfg = IgemmCfg<BN, BK, STAGES>;
  static bool configured = false;
  auto kern = igemm_conv_kernel<BN, BK, STAGES>;
  if (!configured) {
    CGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  kern<<<grid, 192, Cfg::kSmemBytes, stream>>>(tmA, tmB, args);
  CGB_CUDA(cudaGetLastError());
}

void launch_igemm(int BN, int BK, const CUtensorMap& tmA, const CUtensorMap& tmB, const IgemmArgs& args,
                  int num_tiles, int n_blocks, int n_classes, cudaStream_t stream) {
  dim3 grid(num_tiles, n_blocks, n_classes);
  if (BK == 64) {
    switch (BN) {
      case 256: return launch_igemm_t<256, 64, 4>(tmA, tmB, args, grid, stream);
      case 128: return launch_igemm_t<128, 64, 4>(tmA, tmB, args, grid, stream);
      case 64: return launch_igemm_t<64, 64, 4>(tmA, tmB, args, grid, stream);
      case 16: return launch_igemm_t<16, 64, 6>(tmA, tmB, args, grid, stream);
      default: break;
    }
  } else if (BK == 16) {
    switch (BN) {
      case 64: return launch_igemm_t<64, 16, 8>(tmA, tmB, args, grid, stream);
      case 16: return launch_igemm_t<16, 16, 8>(tmA, tmB, args, grid, stream);
      default: break;
    }
  }
  CGB_CHECK(false, "launch_igemm: unsupported tile BN=" + std::to_string(BN) + " BK=" + std::to_string(BK));
}

template <int BNW, int STAGES>
static void launch_wgrad_t(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradArgs& args, dim3 grid,
                           cudaStream_t stream) {
  using Cfg = WgradCfg<BNW, STAGES>;
  static bool configured = false;
  auto kern = wgrad_kernel<BNW, STAGES>;
  if (!configured) {
    CGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  kern<<<grid, 192, Cfg::kSmemBytes, stream>>>(tmDY, tmX, args);
  CGB_CUDA(cudaGetLastError());
}

void launch_wgrad(int BNW, const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradArgs& args, int m_blocks,
                  cudaStream_t stream) {
  const int n_blocks = (args.Cin + BNW - 1) / BNW;
  dim3 grid(m_blocks * n_blocks, args.num_taps, args.split_k);
  switch (BNW) {
    case 256: return launch_wgrad_t<256, 4>(tmDY, tmX, args, grid, stream);
    case 128: return launch_wgrad_t<128, 4>(tmDY, tmX, args, grid, stream);
    case 64: return launch_wgrad_t<64, 4>(tmDY, tmX, args, grid, stream);
    default: break;
  }
  CGB_CHECK(false, "launch_wgrad: unsupported BNW=" + std::to_string(BNW));
}

}  // namespace cgb
