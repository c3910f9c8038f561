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
#include "conv_epilogue.cuh"
#include "conv_tc.h"
#include "ptx.cuh"

namespace cgb {

using namespace ptx;

// KPS = K iterations (tap x channel-chunk boxes) carried by one pipeline stage: one full/empty barrier round
// trip and one tcgen05.commit per stage instead of per K iteration (the round trip costs ~350 cycles, which
// dominates narrow tiles and the 49-tap 7x7 layers).
template <int BN, int BK, int STAGES, int KPS>
struct IgemmCfg {
  static constexpr int kSwizzle = BK * 2;  // bytes per smem row == swizzle span
  static constexpr int kABytes = 128 * kSwizzle;
  static constexpr int kBBytesTx = BN * kSwizzle;
  static constexpr int kBBytes = (kBBytesTx + 1023) / 1024 * 1024;
  static constexpr int kSubBytes = kABytes + kBBytes;  // one K iteration
  static constexpr int kStageBytes = KPS * kSubBytes;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  static constexpr int kMaxKIters = 192;  // K-iteration table staged in smem (largest layer: 128)
  static constexpr int kSmemBytes =
      STAGES * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + kMaxKIters * 16 + 1024 /*bias*/;
  static_assert(kSmemBytes <= 232448, "shared memory budget exceeded");
  static_assert(STAGES * kStageBytes >= 128 * BN * 2 || BN < 64, "epilogue staging does not fit");
};

// CM x CN thread-block cluster: the CN CTAs that share an M tile (same blockIdx.x, consecutive N blocks) each
// fetch 1/CN of the activation box and multicast it to the others; the CM CTAs that share an N block
// (consecutive M tiles) do the same with the weight tile.  L2 -> SM operand traffic drops by CN (A) and CM (B).
// MINB = CTAs per SM the register allocation must allow: the "lite" instantiations (2-stage rings, one K iteration per
// stage, 54-102 KB of shared memory) put 2-3 CTAs on an SM for launches of many short CTAs (the stride-2 / transposed
// layers on large maps: 2-9 K iterations per CTA), so one CTA's prologue and epilogue overlap another's loads and MMAs.
template <int BN, int BK, int STAGES, int KPS, int CM, int CN, int MINB = 1>
__global__ void __launch_bounds__(192, MINB)
igemm_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const IgemmArgs args) {
  using Cfg = IgemmCfg<BN, BK, STAGES, KPS>;
  constexpr bool kCluster = (CM * CN) > 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  KIter* s_kiters = reinterpret_cast<KIter*>(smem + STAGES * Cfg::kStageBytes + 256);
  float* s_bias = reinterpret_cast<float*>(smem + STAGES * Cfg::kStageBytes + 256 + Cfg::kMaxKIters * 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cls = blockIdx.z;
  const int nblk = blockIdx.y;
  long long* prof = args.prof ? args.prof + 16 * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) : nullptr;
  if (prof && threadIdx.x == 0) prof[0] = clock64();
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

  // cluster geometry: rank = x + y * CM (x: M-tile direction, y: N-block direction)
  uint32_t cx = 0, cy = 0;
  uint16_t a_mask = 1, b_mask = 1, peer_mask = 1;
  if constexpr (kCluster) {
    const uint32_t rank = cluster_ctarank();
    cx = rank % CM;
    cy = rank / CM;
    a_mask = 0;
    b_mask = 0;
    for (int y = 0; y < CN; ++y) a_mask |= (uint16_t)(1u << (cx + y * CM));  // CTAs sharing my M tile
    for (int x = 0; x < CM; ++x) b_mask |= (uint16_t)(1u << (x + cy * CM));  // CTAs sharing my N block
    peer_mask = a_mask | b_mask;
  }
  constexpr int kPeers = CM + CN - 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kPeers);  // every CTA that receives my slices must release the slot
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  } else if (warp == 1) {
    tmem_alloc(tmem_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kCluster) cluster_sync_all();  // peers' barriers are initialised before anyone multicasts
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();
  if (prof && threadIdx.x == 0) prof[1] = clock64();

  if (warp == 0) {
    // ===================== TMA producer (whole warp converged, one elected lane issues) =====================
    for (int i = lane; i < kcnt; i += 32) s_kiters[i] = args.kiters[kbeg + i];
    __syncwarp();
    // A slice: rows [cy * 128/CN, ...) of the 128-pixel tile = SH x SW pixels; B slice: BN/CM weight rows
    constexpr int kASliceRows = 128 / CN;
    constexpr int kBSliceRows = BN / CM;
    const int a_row0 = cy * kASliceRows;
    const int a_dh = a_row0 >> args.tw_shift, a_dw = a_row0 & (TW - 1);
    const int nstages = (kcnt + KPS - 1) / KPS;
    for (int i = 0; i < nstages; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (elect_one()) {
        const int nk = min(KPS, kcnt - i * KPS);
        mbar_arrive_expect_tx(&full_bar[s], nk * (Cfg::kABytes + Cfg::kBBytesTx));
#pragma unroll
        for (int j = 0; j < KPS; ++j) {
          if (j < nk) {
            const KIter kt = s_kiters[i * KPS + j];
            uint8_t* sa = smem + s * Cfg::kStageBytes + j * Cfg::kSubBytes;
            uint8_t* sb = sa + Cfg::kABytes;
            if constexpr (CN > 1) {
              tma_load_5d_mc(sa + a_row0 * Cfg::kSwizzle, &tmA, &full_bar[s], kt.a_c, wo0 + a_dw + kt.a_dx, kt.a_par,
                             ho0 + a_dh + kt.a_dy, n, a_mask);
            } else {
              tma_load_5d(sa, &tmA, &full_bar[s], kt.a_c, wo0 + kt.a_dx, kt.a_par, ho0 + kt.a_dy, n);
            }
            if constexpr (CM > 1) {
              tma_load_2d_mc(sb + cx * kBSliceRows * Cfg::kSwizzle, &tmB, &full_bar[s], kt.b_k,
                             nblk * BN + cx * kBSliceRows, b_mask);
            } else {
              tma_load_2d(sb, &tmB, &full_bar[s], kt.b_k, nblk * BN);
            }
          }
        }
        if (prof && i == 0) prof[2] = clock64();
      }
      __syncwarp();
    }
    if (prof && lane == 0) prof[3] = clock64();
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
    constexpr uint32_t idesc = make_idesc_bf16(128, BN < 16 ? 16 : BN, 0, 0);
    constexpr uint32_t desc_hi = smem_desc_hi(8 * Cfg::kSwizzle, swizzle_layout_type(Cfg::kSwizzle));
    const uint32_t lo0 = smem_u32(smem) >> 4;  // smem addresses are < 256 KB: 14 bits after the shift
    const int nstages = (kcnt + KPS - 1) / KPS;
    for (int i = 0; i < nstages; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const int nk = min(KPS, kcnt - i * KPS);
#pragma unroll
        for (int j = 0; j < KPS; ++j) {
          if (j < nk) {
            const uint32_t a_lo = lo0 + s * (Cfg::kStageBytes >> 4) + j * (Cfg::kSubBytes >> 4);
            const uint32_t b_lo = a_lo + (Cfg::kABytes >> 4);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {  // +32 bytes (2 units) per 16-element K step inside the swizzle atom
              umma_bf16(tmem_base, smem_desc_join(a_lo + 2 * k, desc_hi), smem_desc_join(b_lo + 2 * k, desc_hi), idesc,
                        (i | j | k) != 0 ? 1u : 0u);
            }
          }
        }
        // frees the smem slot once these MMAs retire (in every CTA that feeds this one)
        if constexpr (kCluster) {
          umma_commit_mc(&empty_bar[s], peer_mask);
        } else {
          umma_commit(&empty_bar[s]);
        }
        if (i == nstages - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
    if (prof && lane == 0) prof[4] = clock64();
  } else if (kcnt > 0) {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int ho = ho0 + (row >> args.tw_shift);
    const int wo = wo0 + (row & (TW - 1));
    // this CTA's slice of the bias vector -> shared memory (read back as warp-wide broadcasts)
    if (args.bias != nullptr) {
      for (int j = threadIdx.x - 64; j < BN; j += 128) {
        const int co = nblk * BN + j;
        s_bias[j] = co < args.bias_n ? __ldg(args.bias + co) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
    }
    mbar_wait_relaxed(tmem_full_bar, 0);
    tc_fence_after();
    pdl_launch_dependents();
    if (prof && threadIdx.x == 64) prof[5] = clock64();
    epilogue_tile<BN>(args, tmem_base, smem, s_bias, n, ho, wo, nblk, args.out_off[cls], q, lane, prof);
  }
  if (prof && threadIdx.x == 64) prof[6] = clock64();
  tc_fence_before();
  __syncthreads();
  if (prof && threadIdx.x == 0) prof[7] = clock64();
  if constexpr (kCluster) cluster_sync_all();  // no peer may still signal my barriers after I exit
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
  const int ncols = args.ncols > 0 ? args.ncols : args.Cin;
  const int n_blocks = (ncols + BNW - 1) / BNW;
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
  pdl_wait();

  if (warp == 0) {
    // TMA producer: whole warp converged, one elected lane issues
    for (int i = 0; i < kcnt; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (elect_one()) {
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
        for (int j = 0; j < 2; ++j) {
          if (args.a_virtual)  // (channel 0, w, row pair, h, n): atom = rows h + 2a and h + 2a + 1 of the expanded tensor
            tma_load_5d(sa + j * Cfg::kAtomBytes, &tmDY, &full_bar[s], 0, w0, (mblk * 128 + j * 64) >> 6, h0, n);
          else
            tma_load_5d(sa + j * Cfg::kAtomBytes, &tmDY, &full_bar[s], tap.a_c + mblk * 128 + j * 64, w0 + tap.a_dx,
                        tap.a_par, h0 + tap.a_dy, n);
        }
#pragma unroll
        for (int j = 0; j < BNW / 64; ++j) {
          if (args.b_virtual)
            tma_load_5d(sb + j * Cfg::kAtomBytes, &tmX, &full_bar[s], 0, w0, (nblk * BNW + j * 64) >> 6, h0, n);
          else
            tma_load_5d(sb + j * Cfg::kAtomBytes, &tmX, &full_bar[s], tap.b_c + nblk * BNW + j * 64, w0 + tap.b_dx,
                        tap.b_par, h0 + tap.b_dy, n);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // MMA issuer: MN-major SWIZZLE_128B operands, LBO = one 64-channel atom (8 KB), SBO = 8 pixel rows (1 KB)
    constexpr uint32_t idesc = make_idesc_bf16(128, BNW, 1, 1);
    constexpr uint32_t desc_hi = smem_desc_hi(1024, 2);
    constexpr uint32_t lbo_lo = (Cfg::kAtomBytes >> 4) << 16;
    const uint32_t lo0 = smem_u32(smem) >> 4;
    for (int i = 0; i < kcnt; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = (lo0 + s * (Cfg::kStageBytes >> 4)) | lbo_lo;
        const uint32_t b_lo = a_lo + (Cfg::kABytes >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 64 pixels per stage, UMMA_K = 16 pixels = 16 rows of 128 B = 2 KB
          umma_bf16(tmem_base, smem_desc_join(a_lo + k * 128, desc_hi), smem_desc_join(b_lo + k * 128, desc_hi), idesc,
                    (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
        if (i == kcnt - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
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
    const bool vec_ok = (args.Cin & 3) == 0 && args.col_map == nullptr;
    mbar_wait_relaxed(tmem_full_bar, 0);
    tc_fence_after();
    if (args.trigger) pdl_launch_dependents();
    // Coalesced reductions: a thread owns one gradient ROW (cout), so a direct red.v4 per thread touches 32
    // different rows per warp instruction (32 L2 transactions of 16 bytes).  The 32 x 32 fp32 block of each
    // chunk is transposed through shared memory (the pipeline buffers are idle now) so that every warp
    // instruction covers 4 rows x 128 contiguous bytes.
    float* tstage = reinterpret_cast<float*>(smem) + q * (32 * 33);
    const unsigned long long grow_u = reinterpret_cast<unsigned long long>(grow);
#pragma unroll 1
    for (int c = 0; c < BNW; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      tmem_ld_wait();
      const int ci0 = nblk * BNW + c;
      if (vec_ok && ci0 + 32 <= args.Cin) {
#pragma unroll
        for (int j = 0; j < 32; ++j) tstage[lane * 33 + j] = __uint_as_float(r[j]);
        __syncwarp();
        const int sub = lane >> 3, piece = lane & 7;  // 4 rows per instruction, 8 x 16 bytes per row
#pragma unroll
        for (int r0 = 0; r0 < 32; r0 += 4) {
          const int rr = r0 + sub;
          const unsigned long long gp = __shfl_sync(0xffffffffu, grow_u, rr);
          const int ok = __shfl_sync(0xffffffffu, valid ? 1 : 0, rr);
          const float* sp = tstage + rr * 33 + piece * 4;
          if (ok) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gp + (unsigned long long)(ci0 + piece * 4) * 4),
                         "f"(sp[0]), "f"(sp[1]), "f"(sp[2]), "f"(sp[3])
                         : "memory");
          }
        }
        __syncwarp();
      } else if (valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          int cc = ci0 + j;
          if (cc < ncols) {
            if (args.col_map != nullptr) cc = __ldg(args.col_map + cc);
            if (cc >= 0 && cc < args.Cin) atomicAdd(grow + cc, __uint_as_float(r[j]));
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
template <int BN, int BK, int STAGES, int KPS, int CM, int CN, int MINB = 1>
static void launch_igemm_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const IgemmArgs& args, dim3 grid,
                           cudaStream_t stream) {
  using Cfg = IgemmCfg<BN, BK, STAGES, KPS>;
  static bool configured = false;
  auto kern = igemm_conv_kernel<BN, BK, STAGES, KPS, CM, CN, MINB>;
  if (!configured) {
    CGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  if (CM * CN == 1) {
    launch_pdl(kern, grid, dim3(192), (size_t)Cfg::kSmemBytes, stream, tmA, tmB, args);
  } else {
    grid.x = (grid.x + CM - 1) / CM * CM;  // padded tiles compute masked-out pixels
    CGB_CHECK(grid.y % CN == 0, "cluster N extent must divide the number of N blocks");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CM;
    attr[0].val.clusterDim.y = CN;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CGB_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, args));
  }
  CGB_CUDA(cudaGetLastError());
}

void launch_igemm(int BN, int BK, int CM, int CN, bool lite, const CUtensorMap& tmA, const CUtensorMap& tmB,
                  const IgemmArgs& args, int num_tiles, int n_blocks, int n_classes, cudaStream_t stream) {
  dim3 grid(num_tiles, n_blocks, n_classes);
  if (lite && BK == 64 && CM * CN == 1) {
    switch (BN) {
      case 256: return launch_igemm_t<256, 64, 2, 1, 1, 1, 2>(tmA, tmB, args, grid, stream);
      case 128: return launch_igemm_t<128, 64, 2, 1, 1, 1, 3>(tmA, tmB, args, grid, stream);
      case 64: return launch_igemm_t<64, 64, 2, 1, 1, 1, 3>(tmA, tmB, args, grid, stream);
      default: break;
    }
  }
  const int key = BN * 1000000 + BK * 10000 + CM * 100 + CN;
  switch (key) {
    // single-CTA (BN, BK, stages, K iterations per stage)
    case 256 * 1000000 + 64 * 10000 + 101: return launch_igemm_t<256, 64, 4, 1, 1, 1>(tmA, tmB, args, grid, stream);
    case 128 * 1000000 + 64 * 10000 + 101: return launch_igemm_t<128, 64, 3, 2, 1, 1>(tmA, tmB, args, grid, stream);
    case 64 * 1000000 + 64 * 10000 + 101: return launch_igemm_t<64, 64, 4, 2, 1, 1>(tmA, tmB, args, grid, stream);
    case 16 * 1000000 + 64 * 10000 + 101: return launch_igemm_t<16, 64, 3, 4, 1, 1>(tmA, tmB, args, grid, stream);
    case 64 * 1000000 + 16 * 10000 + 101: return launch_igemm_t<64, 16, 4, 7, 1, 1>(tmA, tmB, args, grid, stream);
    case 16 * 1000000 + 16 * 10000 + 101: return launch_igemm_t<16, 16, 4, 7, 1, 1>(tmA, tmB, args, grid, stream);
    // clusters with TMA multicast (experimental, CGB_CLUSTER=1)
    case 64 * 1000000 + 64 * 10000 + 204: return launch_igemm_t<64, 64, 6, 1, 2, 4>(tmA, tmB, args, grid, stream);
    case 64 * 1000000 + 64 * 10000 + 402: return launch_igemm_t<64, 64, 6, 1, 4, 2>(tmA, tmB, args, grid, stream);
    case 64 * 1000000 + 64 * 10000 + 401: return launch_igemm_t<64, 64, 6, 1, 4, 1>(tmA, tmB, args, grid, stream);
    case 128 * 1000000 + 64 * 10000 + 402: return launch_igemm_t<128, 64, 4, 1, 4, 2>(tmA, tmB, args, grid, stream);
    case 128 * 1000000 + 64 * 10000 + 401: return launch_igemm_t<128, 64, 4, 1, 4, 1>(tmA, tmB, args, grid, stream);
    case 256 * 1000000 + 64 * 10000 + 401: return launch_igemm_t<256, 64, 4, 1, 4, 1>(tmA, tmB, args, grid, stream);
    default: break;
  }
  CGB_CHECK(false, "launch_igemm: unsupported config BN=" + std::to_string(BN) + " BK=" + std::to_string(BK) +
                       " cluster " + std::to_string(CM) + "x" + std::to_string(CN));
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
  launch_pdl(kern, grid, dim3(192), (size_t)Cfg::kSmemBytes, stream, tmDY, tmX, args);
}

void launch_wgrad(int BNW, const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradArgs& args, int m_blocks,
                  cudaStream_t stream) {
  const int n_blocks = ((args.ncols > 0 ? args.ncols : args.Cin) + BNW - 1) / BNW;
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
