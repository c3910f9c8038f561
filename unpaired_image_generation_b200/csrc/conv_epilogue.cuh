// Accumulator epilogue shared by the tcgen05 implicit-GEMM conv kernels (conv_tc.cu, conv_patch.cu):
// TMEM -> registers -> (+bias, activation) -> bf16 -> global, with the InstanceNorm statistics (sum, sum of
// squares of the ROUNDED outputs) reduced per column and accumulated with one vector atomic per column per warp.
// Stand-in counterpart: the bias add of F.conv2d, torch.tanh / leaky_relu, and the mean/var reductions of
// `_inorm` in oracle/cyclegan_standin.py.
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace cgb {

// One warp (TMEM lane quarter q) drains its 32 accumulator rows x BN columns.
//   tmem_acc : TMEM address of column 0 of this accumulator (lane field 0)
//   stage    : shared-memory staging area of the CTA, >= 128 * BN * 2 bytes, idle pipeline buffers (BN >= 64)
//   s_bias   : BN floats (this CTA's slice of the bias) or unused when args.bias == nullptr
//   ho, wo   : this thread's output pixel (row q*32 + lane of the tile); n : image; out_off : class offset
//   SC       : columns staged per write-out round (BN: whole rows through `stage` of 128*BN*2 bytes; 64: rounds of
//              128-byte row pieces through 16 KB, for kernels whose pipeline buffers stay busy during the epilogue)
template <int BN, int SC = BN>
__device__ __forceinline__ void epilogue_tile(const IgemmArgs& args, uint32_t tmem_acc, uint8_t* stage,
                                              const float* s_bias, int n, int ho, int wo, int nblk, long long out_off,
                                              int q, int lane, long long* prof) {
  using namespace ptx;
  const bool valid = (ho < args.Ho) && (wo < args.Wo) && (n < args.N);
  bf16* orow = args.out + out_off + (long long)n * args.sN + (long long)ho * args.sH + (long long)wo * args.sW;
  constexpr int CH = BN >= 32 ? 32 : 16;
  // BN >= 64: rows are staged in shared memory (the pipeline buffers are idle once the accumulator is
  // complete) and written out with every warp instruction covering whole 128-byte lines of one pixel.
  constexpr bool kStaged = BN >= 64;
  constexpr int kRowBytes = SC * 2;
  static_assert(!kStaged || (SC >= 64 && BN % SC == 0), "staging round must be a multiple of 64 columns dividing BN");
  uint8_t* stage_base = stage + (size_t)q * 32 * kRowBytes;  // this warp's 32 rows
  const unsigned long long optr0 = reinterpret_cast<unsigned long long>(orow + nblk * BN);
#pragma unroll 1
  for (int c = 0; c < BN; c += CH) {
    float v[CH];
    {
      uint32_t r[CH];
      const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + c;
      if constexpr (CH == 32) {
        tmem_ld32(taddr, r);
      } else {
        tmem_ld16(taddr, r);
      }
      tmem_ld_wait();
      if (prof && threadIdx.x == 64 && c == 0) prof[8] = clock64();
#pragma unroll
      for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]);
    }
    const int co0 = nblk * BN + c;
    if (args.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < CH; j += 4) {
        const float4 bv = *reinterpret_cast<const float4*>(s_bias + c + j);
        v[j] += bv.x;
        v[j + 1] += bv.y;
        v[j + 2] += bv.z;
        v[j + 3] += bv.w;
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
    if (args.stats != nullptr) {
      // InstanceNorm statistics of the bf16-rounded outputs, fused: per-column sums over this warp's 32
      // pixel rows by a transposing butterfly (31 shuffles per quantity), then one atomic per column.
      float a[CH], b[CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const float r = valid ? __bfloat162float(__float2bfloat16_rn(v[j])) : 0.f;
        a[j] = r;
        b[j] = r * r;
      }
#pragma unroll
      for (int off = CH / 2; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < off; ++j) {
          const float sa = up ? a[j] : a[j + off];
          const float sb = up ? b[j] : b[j + off];
          const float ka = up ? a[j + off] : a[j];
          const float kb = up ? b[j + off] : b[j];
          a[j] = ka + __shfl_xor_sync(0xffffffffu, sa, off);
          b[j] = kb + __shfl_xor_sync(0xffffffffu, sb, off);
        }
      }
      // lane l now holds column (l mod CH); for CH = 16 lanes l and l + 16 hold halves of the same column
      if (CH == 16) {
        a[0] += __shfl_xor_sync(0xffffffffu, a[0], 16);
        b[0] += __shfl_xor_sync(0xffffffffu, b[0], 16);
      }
      const int col = co0 + (lane & (CH - 1));
      if (n < args.N && col < args.Cout && (CH == 32 || lane < 16)) {
        float* st = args.stats + ((long long)n * args.Cout + col) * 2;
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(st), "f"(a[0]), "f"(b[0]) : "memory");
      }
    }
    if (prof && threadIdx.x == 64 && c == 0) prof[9] = clock64();
    if constexpr (kStaged) {
      // 16-byte pieces, XOR-swizzled by the row so that the 32 lanes of a store spread over all banks
      uint8_t* srow = stage_base + (size_t)lane * kRowBytes;
#pragma unroll
      for (int j = 0; j < CH; j += 8) {
        uint4 pk;
        pk.x = pack_bf16x2(v[j + 0], v[j + 1]);
        pk.y = pack_bf16x2(v[j + 2], v[j + 3]);
        pk.z = pack_bf16x2(v[j + 4], v[j + 5]);
        pk.w = pack_bf16x2(v[j + 6], v[j + 7]);
        const int piece = ((c % SC) + j) >> 3;  // 16-byte piece index within the staged row
        *reinterpret_cast<uint4*>(srow + (((piece & ~7) | ((piece ^ lane) & 7)) << 4)) = pk;
      }
    } else if (valid) {
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
    if constexpr (kStaged) {
      if ((c + CH) % SC == 0) {  // a staging round is complete: write it out
        __syncwarp();
        // each warp instruction writes kRowsPerInst pixel rows of SC channels (>= 128 contiguous bytes each)
        constexpr int kLanesPerRow = kRowBytes / 16;  // 8, 16 or 32
        constexpr int kRowsPerInst = 32 / kLanesPerRow;
        const int sub = lane / kLanesPerRow, piece = lane % kLanesPerRow;
        const unsigned long long optr = optr0 + (unsigned long long)(c + CH - SC) * 2;
#pragma unroll 4
        for (int r0 = 0; r0 < 32; r0 += kRowsPerInst) {
          const int r = r0 + sub;
          const unsigned long long p = __shfl_sync(0xffffffffu, optr, r);
          const int ok = __shfl_sync(0xffffffffu, valid ? 1 : 0, r);
          const uint4 val = *reinterpret_cast<const uint4*>(stage_base + (size_t)r * kRowBytes +
                                                            (((piece & ~7) | ((piece ^ r) & 7)) << 4));
          if (ok) *reinterpret_cast<uint4*>(p + (unsigned long long)piece * 16) = val;
        }
        __syncwarp();  // the staging rows are rewritten by the next round / the caller's next tile
      }
    }
  }
  if (prof && threadIdx.x == 64) prof[10] = clock64();
}

}  // namespace cgb
