// Patch-resident, persistent tcgen05 implicit GEMM for stride-1 convolutions and stride-1 input gradients (sm_100a).
//
// igemm_conv_kernel (conv_tc.cu) re-fetches the 128-pixel activation box for every filter tap: a 3x3 conv pulls 9x
// (a 7x7 conv 49x) the activation bytes through L2 and needs one TMA issue + barrier round trip per tap and
// channel chunk.  Here the halo'd input patch of a 16 x 8 pixel output tile ((16 + k - 1) x (8 + k - 1) pixels x
// 64 channels, SWIZZLE_128B rows of 128 bytes -- or x 16 channels, SWIZZLE_32B rows of 32 bytes) is fetched ONCE per
// channel chunk by one TMA box; every filter tap reads it in place through a UMMA shared-memory descriptor whose
// start address is shifted by (py * PW + px) rows and whose stride between 8-pixel row groups (SBO) is the patch
// pitch.  This works because the UMMA swizzle XOR is a function of the absolute shared-memory address, like TMA's
// (measured: profiles/r01_d_patch_descriptor_modes.txt); the descriptor's base-offset field stays 0.  Only the
// weights stream per tap (and stay resident when the whole filter fits the ring).
//
// CTAs are persistent (at most one per SM and N block): each loops over its work items (MT stacked tiles) with two
// TMEM accumulator sets, so the epilogue of item i overlaps the MMAs of item i + 1.  Measured on the residual conv
// at batch 8 (ncu, profiles/r01_f_ncu_full_res_conv.txt): L2 -> SM traffic 453 -> 326 MB per launch, DRAM traffic =
// the algorithmic 19 MB, 43.6 -> 32.3 us (1198 TFLOP/s); the 49-tap head dgrad 268 -> 89 us.
//
// Stand-in counterpart: F.conv2d (stride 1, reflection- or zero-padded) and its input gradient in
// oracle/cyclegan_standin.py (ResnetBlock convs, generator stem and head, discriminator conv3/conv4).
//
// Warp roles (224 threads): warp 0 = weight-tile TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue, warp 6 = patch TMA producer.
#include "common.h"
#include "conv_epilogue.cuh"
#include "conv_tc.h"
#include "ptx.cuh"

#include <cstdlib>

namespace cgb {

using namespace ptx;

namespace {
constexpr int kPatchStageBytes = 16384;  // epilogue staging: 128 rows x 64 columns x bf16
constexpr int kPatchMisc = kPatchStageBytes + 512 /*barriers*/ + 512 /*tap tables*/ + 1024 /*bias*/;
constexpr int kPatchMaxBStages = 16;
constexpr int kPatchSmemMax = 232448;
}  // namespace

// KA = channels per patch row: 64 (128-byte rows, SWIZZLE_128B, four K = 16 MMAs per tap) or 16 (the 16-stored-
// channel image-like tensors: 32-byte rows, SWIZZLE_32B, one MMA per tap).  Weight boxes are always 64 K-elements
// wide (SWIZZLE_128B): with KA = 16 one box carries four consecutive taps.
//
// CG = 2: CTA pairs (cta_group::2, 2-CTA clusters).  The pair computes two M tiles (one per CTA, M = 256 per MMA)
// against ONE weight tile of which each CTA stages half the rows (BN / 2): the weight bytes streamed from L2 per
// FLOP -- the limiter of the single-CTA kernel at 148 CTAs (measured: tensor pipe 66 % active, 60 B/clk/SM of weight
// boxes against the ~43 B/clk/SM the L2 delivers chip-wide) -- and the B-operand shared-memory reads per SM halve.
// Only the leader (cluster rank 0) issues MMAs; both CTAs run their own patch / weight producers (the peer's TMA
// bytes are credited to the leader's full barriers) and their own epilogue (each CTA's TMEM holds its own 128 rows).
template <int BN, int MT, int KPS, int KA, int CG>
__global__ void __launch_bounds__(224, 1)
igemm_patch_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const IgemmArgs args, const PatchArgs pa) {
  static_assert(KA == 64 || KA == 16, "patch rows carry 64 or 16 channels");
  static_assert(CG == 1 || (CG == 2 && BN >= 32), "CTA pairs split the N tile in two");
  constexpr int TPB = 64 / KA;   // taps per weight box
  constexpr int kKK = KA / 16;   // K = 16 MMAs per tap
  constexpr int BNL = BN / CG;   // weight rows staged by this CTA
  constexpr int kBBytesTx = BNL * 128;
  constexpr int kBBytes = (kBBytesTx + 1023) / 1024 * 1024;
  constexpr int kStageBytes = KPS * kBBytes;
  constexpr int kAcc = BN < 32 ? 32 : BN;  // TMEM columns per accumulator
  // two accumulator sets when they fit: the epilogue of tile i drains one while the MMAs of tile i+1 fill the other
  constexpr int NACC = (2 * MT * kAcc <= 512) ? 2 : 1;
  constexpr int kTmemCols = NACC * MT * kAcc;
  static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM allocation must be a power of two <= 512");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int ring_bytes = pa.b_stages * kStageBytes;
  uint8_t* patches = smem + ring_bytes;
  uint8_t* stage = patches + 2 * MT * pa.patch_bytes;  // epilogue staging: 4 warps x 32 rows x 128 bytes
  uint8_t* misc = stage + kPatchStageBytes;
  uint64_t* b_full = reinterpret_cast<uint64_t*>(misc);
  uint64_t* b_empty = b_full + kPatchMaxBStages;
  uint64_t* a_full = b_empty + kPatchMaxBStages;
  uint64_t* a_empty = a_full + 2;
  uint64_t* tmem_full_bar = a_empty + 2;    // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint32_t* s_aoff = reinterpret_cast<uint32_t*>(misc + 512);  // per tap: descriptor start offset (16 B units)
  int32_t* s_bk = reinterpret_cast<int32_t*>(misc + 512 + 256);  // per tap: K offset in the packed weights
  float* s_bias = reinterpret_cast<float*>(misc + 512 + 512);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nblk = blockIdx.y;
  const int rank = CG == 2 ? (int)cluster_ctarank() : 0;  // 0 = leader of the pair
  long long* prof = args.prof ? args.prof + 16 * (blockIdx.x + gridDim.x * blockIdx.y) : nullptr;
  if (prof && threadIdx.x == 0) prof[0] = clock64();

  const int T = pa.k * pa.k;
  const int nbox = (T + TPB - 1) / TPB;      // weight boxes per channel chunk
  const int nbs = (nbox + KPS - 1) / KPS;    // weight stages per channel chunk
  // persistent: this CTA owns the M-direction work items blockIdx.x, blockIdx.x + gridDim.x, ...
  // (a work item is MT stacked tiles of 16 x 8 output pixels of one image).  CTA pairs walk UNITS of two consecutive
  // items (item = 2 * unit + rank); the second item of the last unit may not exist: the peer then streams an
  // out-of-range image (TMA zero fill) and skips its epilogue.
  const int items = pa.num_items;
  const int tiles_per_img = args.tiles_w * args.tiles_h;
  const int unit0 = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int ustep = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int units = CG == 2 ? (items + 1) / 2 : items;

  // tables indexed by the WEIGHT tap w (taps are swept in weight order; input gradients read the patch backwards)
  for (int w = threadIdx.x; w < T; w += blockDim.x) {
    const int i = pa.flip ? T - 1 - w : w;
    const int py = i / pa.k, px = i - py * pa.k;
    s_aoff[w] = (uint32_t)(py * pa.row_step + px * pa.col_step);
    s_bk[w] = w * pa.tap_stride;
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < pa.b_stages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 128 * CG);  // every epilogue thread (of both CTAs) arrives once its TMEM reads are done
    }
    fence_mbar_init();
  } else if (warp == 1) {
    if constexpr (CG == 2) {
      tmem_alloc_pair(tmem_ptr, kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_ptr, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // everything above overlapped the previous kernel's tail; its outputs are visible from here on
  if (prof && threadIdx.x == 0) prof[1] = clock64();

  // The producer / MMA loops are executed by ONE thread each and their per-stage instruction latency is on the
  // critical path of narrow tiles: ring positions and phases are tracked incrementally (no runtime div / mod).
  if (warp == 6) {
    // ===================== patch producer =====================
    uint32_t ph = 1;  // parity to wait for on a_empty: the first pass over the two buffers does not block
    int ca = 0;
    for (int unit = unit0; unit < units; unit += ustep) {
      const int item = CG == 2 ? 2 * unit + rank : unit;
      int n = item / tiles_per_img;
      const int r = item - n * tiles_per_img;
      const int th = r / args.tiles_w, tw = r - th * args.tiles_w;
      const int wo0 = tw * 8, ho0 = th * 16 * MT;
      if (item >= items) n = args.N;  // missing second item of the last unit: out-of-range image, zero filled
      for (int c = 0; c < pa.chunks; ++c) {
        mbar_wait(&a_empty[ca], ph);
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&a_full[ca], CG * MT * pa.nbox * pa.box_bytes);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            uint8_t* dst = patches + (ca * MT + mt) * pa.patch_bytes;
            for (int b = 0; b < pa.nbox; ++b) {
              if constexpr (CG == 2)
                tma_load_5d_pair(dst + b * pa.box_bytes, &tmA, &a_full[ca], c * KA, wo0 + pa.ox + b, 0,
                                 ho0 + mt * 16 + pa.oy, n);
              else
                tma_load_5d(dst + b * pa.box_bytes, &tmA, &a_full[ca], c * KA, wo0 + pa.ox + b, 0,
                            ho0 + mt * 16 + pa.oy, n);
            }
          }
          if (prof && c == 0 && unit == unit0) prof[2] = clock64();
        }
        __syncwarp();
        if (ca == 1) ph ^= 1;
        ca ^= 1;
      }
    }
  } else if (warp == 0) {
    // ===================== weight-tile producer =====================
    int s = 0;
    uint32_t ph = 1;
    uint8_t* sb = smem;
    for (int unit = unit0; unit < units; unit += ustep) {
      if (pa.b_resident && unit != unit0) break;  // the whole filter stays in shared memory after the first item
      for (int c = 0; c < pa.chunks; ++c) {
        for (int bs = 0; bs < nbs; ++bs) {
          mbar_wait(&b_empty[s], ph);
          if (elect_one()) {
            const int nk = min(KPS, nbox - bs * KPS);
            if (rank == 0) mbar_arrive_expect_tx(&b_full[s], CG * nk * kBBytesTx);
#pragma unroll
            for (int j = 0; j < KPS; ++j) {
              if (j < nk) {
                const int box = bs * KPS + j;
                const int kcoord = (KA == 64) ? s_bk[box] + c * 64 : box * 64;
                if constexpr (CG == 2)
                  tma_load_2d_pair(sb + j * kBBytes, &tmB, &b_full[s], kcoord, nblk * BN + rank * BNL);
                else
                  tma_load_2d(sb + j * kBBytes, &tmB, &b_full[s], kcoord, nblk * BN);
              }
            }
          }
          __syncwarp();
          sb += kStageBytes;
          if (++s == pa.b_stages) {
            s = 0;
            ph ^= 1;
            sb = smem;
          }
        }
      }
    }
    if (prof && lane == 0) prof[3] = clock64();
  } else if (warp == 1) {
    // ===================== MMA issuer (the leader's when CTAs are paired) =====================
    if (rank == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128 * CG, BN < 16 ? 16 : BN, 0, 0);
    constexpr uint32_t desc_hi_b = smem_desc_hi(1024, 2);
    const uint32_t desc_hi_a = smem_desc_hi((uint32_t)pa.sbo, swizzle_layout_type(KA * 2));
    const uint32_t lo_ring = (smem_u32(smem) & 0x3FFFFu) >> 4;
    const uint32_t lo_patch = (smem_u32(patches) & 0x3FFFFu) >> 4;
    const uint32_t patch_units = (uint32_t)pa.patch_bytes >> 4;
    int s = 0, ca = 0, acc_i = 0;
    uint32_t ph = 0, a_ph = 0, e_ph = 1;  // e_ph: parity to wait for on tmem_empty (first use of each set: free)
    uint32_t b_lo0 = lo_ring;
    for (int unit = unit0; unit < units; unit += ustep) {
      mbar_wait(&tmem_empty_bar[acc_i], e_ph);  // the epilogue (of both CTAs) has drained this accumulator set
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + acc_i * (MT * kAcc);
      uint32_t accumulate = 0;
      for (int c = 0; c < pa.chunks; ++c) {
        mbar_wait(&a_full[ca], a_ph);
        const uint32_t a_base = lo_patch + (uint32_t)(ca * MT) * patch_units;
        for (int bs = 0; bs < nbs; ++bs) {
          mbar_wait(&b_full[s], pa.b_resident ? 0u : ph);  // resident filter: filled once, never recycled
          tc_fence_after();
          if (elect_one()) {
            const int nk = min(KPS, nbox - bs * KPS);
#pragma unroll
            for (int j = 0; j < KPS; ++j) {
              if (j < nk) {
                const int box = bs * KPS + j;
#pragma unroll
                for (int ts = 0; ts < TPB; ++ts) {
                  const int w = box * TPB + ts;  // weight tap
                  if (TPB > 1 && w >= T) break;
                  const uint32_t b_lo = b_lo0 + j * (kBBytes >> 4) + ts * (KA * 2 / 16);
                  const uint32_t aoff = s_aoff[w];
#pragma unroll
                  for (int mt = 0; mt < MT; ++mt) {
                    const uint32_t a_lo = a_base + mt * patch_units + aoff;
#pragma unroll
                    for (int kk = 0; kk < kKK; ++kk) {  // K = 16 channels = 32 bytes inside the swizzle atom
                      if constexpr (CG == 2)
                        umma_bf16_pair(tmem_acc + mt * kAcc, smem_desc_join(a_lo + 2 * kk, desc_hi_a),
                                       smem_desc_join(b_lo + 2 * kk, desc_hi_b), idesc,
                                       (kk == 0 && j == 0 && ts == 0) ? accumulate : 1u);
                      else
                        umma_bf16(tmem_acc + mt * kAcc, smem_desc_join(a_lo + 2 * kk, desc_hi_a),
                                  smem_desc_join(b_lo + 2 * kk, desc_hi_b), idesc,
                                  (kk == 0 && j == 0 && ts == 0) ? accumulate : 1u);
                    }
                  }
                }
              }
            }
            if constexpr (CG == 2) {  // the same barriers of BOTH CTAs
              if (!pa.b_resident) umma_commit_pair(&b_empty[s]);
              if (bs == nbs - 1) {
                umma_commit_pair(&a_empty[ca]);
                if (c == pa.chunks - 1) umma_commit_pair(&tmem_full_bar[acc_i]);
              }
            } else {
              if (!pa.b_resident) umma_commit(&b_empty[s]);  // weight slot free once these MMAs retire
              if (bs == nbs - 1) {
                umma_commit(&a_empty[ca]);  // ... and the patch buffer after the chunk's last tap
                if (c == pa.chunks - 1) umma_commit(&tmem_full_bar[acc_i]);
              }
            }
          }
          __syncwarp();
          accumulate = 1;
          b_lo0 += kStageBytes >> 4;
          if (++s == pa.b_stages) {
            s = 0;
            ph ^= 1;
            b_lo0 = lo_ring;
          }
        }
        if (ca == 1) a_ph ^= 1;
        ca ^= 1;
      }
      if (++acc_i == NACC) {
        acc_i = 0;
        e_ph ^= 1;
      }
    }
    if (prof && lane == 0) prof[4] = clock64();
    }  // rank == 0
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    if (args.bias != nullptr) {
      for (int j = threadIdx.x - 64; j < BN; j += 128) {
        const int co = nblk * BN + j;
        s_bias[j] = co < args.bias_n ? __ldg(args.bias + co) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
    }
    int acc_i = 0;
    uint32_t f_ph = 0;
    for (int unit = unit0; unit < units; unit += ustep) {
      const int item = CG == 2 ? 2 * unit + rank : unit;
      const int n = item / tiles_per_img, r = item - n * tiles_per_img;
      const int th = r / args.tiles_w, tw = r - th * args.tiles_w;
      const int wo0 = tw * 8, ho0 = th * 16 * MT;
      mbar_wait_relaxed(&tmem_full_bar[acc_i], f_ph);
      tc_fence_after();
      if (unit + ustep >= units) pdl_launch_dependents();  // last main loop done: the next kernel may launch
      if (prof && threadIdx.x == 64 && unit == unit0) prof[5] = clock64();
      const uint32_t tmem_acc = tmem_base + acc_i * (MT * kAcc);
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        if (item >= items || ho0 + mt * 16 >= args.Ho) break;  // missing item / ragged CTA row: nothing to write
        epilogue_tile<BN, (BN >= 64 ? 64 : BN)>(args, tmem_acc + mt * kAcc, stage, s_bias, n, ho0 + mt * 16 + (row >> 3),
                                                wo0 + (row & 7), nblk, args.out_off[0], q, lane, prof);
      }
      tc_fence_before();
      // this thread's TMEM reads of the set are complete (the MMA issuer that refills it lives in the leader)
      if constexpr (CG == 2) mbar_arrive_leader(&tmem_empty_bar[acc_i]);
      else mbar_arrive(&tmem_empty_bar[acc_i]);
      if (++acc_i == NACC) {
        acc_i = 0;
        f_ph ^= 1;
      }
    }
  }
  if (prof && threadIdx.x == 64) prof[6] = clock64();
  tc_fence_before();
  __syncthreads();
  if (prof && threadIdx.x == 0) prof[7] = clock64();
  // neither CTA of a pair may exit (or free its TMEM) while the other can still read its shared memory / signal its barriers
  if constexpr (CG == 2) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// Host launcher
// ------------------------------------------------------------------------------------------
// CGB_PATCH_SMEM_KB caps the dynamic shared memory of one CTA (default: all of it); at <= 112 KB two CTAs share an
// SM, so one CTA's prologue / epilogue overlaps the other's main loop when several convs are in flight.
static int patch_smem_cap() {
  static const int cap = std::getenv("CGB_PATCH_SMEM_KB") ? std::atoi(std::getenv("CGB_PATCH_SMEM_KB")) * 1024 : kPatchSmemMax;
  return cap < kPatchSmemMax ? cap : kPatchSmemMax;
}
static bool patch_kps1() {
  static const bool v = std::getenv("CGB_PATCH_KPS1") != nullptr;
  return v;
}
int igemm_patch_kps(int BN, int ka) { return (ka == 16 || BN >= 256) ? 1 : BN >= 64 ? (patch_kps1() ? 1 : 3) : 7; }
int igemm_patch_smem_budget() { return patch_smem_cap() - 1024 - kPatchMisc; }

// BN = N tile of the MMA; CG = CTAs per tile pair (1, or 2: each CTA stages BN / 2 weight rows and KPS follows that half)
template <int BN, int MT, int KPS, int KA, int CG>
static void launch_patch_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const IgemmArgs& args, const PatchArgs& pa,
                           dim3 grid, cudaStream_t stream) {
  constexpr int kBBytes = ((BN / CG) * 128 + 1023) / 1024 * 1024;
  static bool configured = false;
  auto kern = igemm_patch_kernel<BN, MT, KPS, KA, CG>;
  if (!configured) {
    CGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPatchSmemMax));
    configured = true;
  }
  CGB_CHECK(pa.b_stages >= 2 && pa.b_stages <= kPatchMaxBStages, "patch igemm: weight ring depth out of range");
  CGB_CHECK(pa.k * pa.k <= 64, "patch igemm: at most 64 filter taps");
  CGB_CHECK(pa.ka == KA && (KA == 64 || pa.chunks == 1), "patch igemm: channel-chunk width mismatch");
  const int ring = pa.b_stages * KPS * kBBytes;
  const int smem = 1024 + ring + 2 * MT * pa.patch_bytes + kPatchMisc;
  CGB_CHECK(smem <= kPatchSmemMax, "patch igemm: shared memory budget exceeded");
  if constexpr (CG == 1) {
    launch_pdl(kern, grid, dim3(224), (size_t)smem, stream, tmA, tmB, args, pa);
  } else {
    CGB_CHECK(grid.x % 2 == 0, "patch igemm: CTA pairs need an even grid");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(224);
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
    CGB_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, args, pa));
  }
}

void launch_igemm_patch(int BN, int MT, int CG, const CUtensorMap& tmA, const CUtensorMap& tmB, const IgemmArgs& args,
                        const PatchArgs& pa, int num_ctas_m, int n_blocks, cudaStream_t stream) {
  dim3 grid(num_ctas_m, n_blocks, 1);
  if (CG == 2) {
    CGB_CHECK(pa.ka == 64, "patch igemm: CTA pairs are built for 64-channel chunks");
    switch (BN * 10 + MT) {
      case 2561: return launch_patch_t<256, 1, 3, 64, 2>(tmA, tmB, args, pa, grid, stream);
      case 1282: return launch_patch_t<128, 2, 3, 64, 2>(tmA, tmB, args, pa, grid, stream);
      case 1281: return launch_patch_t<128, 1, 3, 64, 2>(tmA, tmB, args, pa, grid, stream);
      default: break;
    }
    CGB_CHECK(false, "launch_igemm_patch: unsupported CTA-pair config BN=" + std::to_string(BN) + " MT=" + std::to_string(MT));
  }
  if (pa.ka == 16) {
    switch (BN * 10 + MT) {
      case 2561: return launch_patch_t<256, 1, 1, 16, 1>(tmA, tmB, args, pa, grid, stream);
      case 1282: return launch_patch_t<128, 2, 1, 16, 1>(tmA, tmB, args, pa, grid, stream);
      case 1281: return launch_patch_t<128, 1, 1, 16, 1>(tmA, tmB, args, pa, grid, stream);
      case 642: return launch_patch_t<64, 2, 1, 16, 1>(tmA, tmB, args, pa, grid, stream);
      case 641: return launch_patch_t<64, 1, 1, 16, 1>(tmA, tmB, args, pa, grid, stream);
      default: break;
    }
    CGB_CHECK(false, "launch_igemm_patch: unsupported 16-channel config BN=" + std::to_string(BN) + " MT=" + std::to_string(MT));
  }
  switch (BN * 10 + MT) {
    case 2562: return launch_patch_t<256, 2, 1, 64, 1>(tmA, tmB, args, pa, grid, stream);
    case 2561: return launch_patch_t<256, 1, 1, 64, 1>(tmA, tmB, args, pa, grid, stream);
    case 1282: return patch_kps1() ? launch_patch_t<128, 2, 1, 64, 1>(tmA, tmB, args, pa, grid, stream)
                                   : launch_patch_t<128, 2, 3, 64, 1>(tmA, tmB, args, pa, grid, stream);
    case 1281: return patch_kps1() ? launch_patch_t<128, 1, 1, 64, 1>(tmA, tmB, args, pa, grid, stream)
                                   : launch_patch_t<128, 1, 3, 64, 1>(tmA, tmB, args, pa, grid, stream);
    case 642: return patch_kps1() ? launch_patch_t<64, 2, 1, 64, 1>(tmA, tmB, args, pa, grid, stream)
                                  : launch_patch_t<64, 2, 3, 64, 1>(tmA, tmB, args, pa, grid, stream);
    case 641: return patch_kps1() ? launch_patch_t<64, 1, 1, 64, 1>(tmA, tmB, args, pa, grid, stream)
                                  : launch_patch_t<64, 1, 3, 64, 1>(tmA, tmB, args, pa, grid, stream);
    case 162: return launch_patch_t<16, 2, 7, 64, 1>(tmA, tmB, args, pa, grid, stream);
    case 161: return launch_patch_t<16, 1, 7, 64, 1>(tmA, tmB, args, pa, grid, stream);
    default: break;
  }
  CGB_CHECK(false, "launch_igemm_patch: unsupported config BN=" + std::to_string(BN) + " MT=" + std::to_string(MT));
}

}  // namespace cgb
