// HBM-bound kernels of the CycleGAN step (see pointwise.h).  All activations are bf16 NHWC with
// the channel count a multiple of 8, so every thread moves 16-byte vectors.
#include "pointwise.h"

#include "fp32_path.h"
#include "ptx.cuh"

#include <cstdlib>
#include <set>

#ifndef CGB_PW_TRIGGER
#define CGB_PW_TRIGGER 0
#endif

namespace cgb {

namespace {

// 1: pointwise kernels let their successor start launching right away.  Measured: a conv CTA that is launched
// early then sits in griddepcontrol.wait holding ~200 KB of shared memory, which starves the other lanes.
constexpr bool kPwTrigger = CGB_PW_TRIGGER != 0;

struct DevTensor {
  bf16* p;  // interior origin
  long long sN, sH, sW;
  int N, H, W, C, halo;
};

DevTensor dev(const TensorDesc& t) {
  CGB_CHECK(t.esz == 2, "bf16 kernel called on an fp32 (validation mode) tensor");
  DevTensor d;
  d.p = t.interior();
  d.sN = t.sN();
  d.sH = t.sH();
  d.sW = t.sW();
  d.N = t.N;
  d.H = t.H;
  d.W = t.W;
  d.C = t.C;
  d.halo = t.halo;
  return d;
}
// pure copy kernels (halo fill, pool exchange) also serve the fp32 validation mode: an fp32 tensor [.., C] is
// viewed as a bf16 tensor [.., 2C] (strides are in bf16 elements; the base pointer is already byte-exact)
DevTensor dev_bytes(const TensorDesc& t) {
  TensorDesc b = t;
  b.esz = 2;
  DevTensor d = dev(b);
  d.p = t.interior();
  if (t.esz == 4) {
    d.sN *= 2;
    d.sH *= 2;
    d.sW *= 2;
    d.C *= 2;
  }
  return d;
}
DevTensor dev_null() {
  DevTensor d;
  d.p = nullptr;
  d.sN = d.sH = d.sW = 0;
  d.N = d.H = d.W = d.C = d.halo = 0;
  return d;
}

struct DevGrad {
  DevTensor g1, g2;
  int fold;
};
DevGrad dev(const GradSrc& g) {
  DevGrad d;
  d.g1 = g.g1 ? dev(*g.g1) : dev_null();
  d.g2 = g.g2 ? dev(*g.g2) : dev_null();
  d.fold = g.fold;
  return d;
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
// one 16-byte reduction instead of four scalar atomics (same-address fp32 atomics serialise in L2)
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float round_bf16(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// gradient w.r.t. the activation at interior pixel (n, h, w), channels [c0, c0+8)
__device__ __forceinline__ void load_grad8(const DevGrad& g, int n, int h, int w, int c0, int H, int W,
                                           float (&out)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] = 0.f;
  if (g.g1.p != nullptr) {
    float v[8];
    load8(g.g1.p + n * g.g1.sN + h * g.g1.sH + w * g.g1.sW + c0, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] += v[i];
  }
  if (g.g2.p != nullptr) {
    const int p = g.fold;
    int hs[2], ws[2], nh = 1, nw = 1;
    hs[0] = h + p;
    ws[0] = w + p;
    if (h >= 1 && h <= p) hs[nh++] = p - h;
    else if (h >= H - 1 - p && h <= H - 2) hs[nh++] = 2 * (H - 1) - h + p;
    if (w >= 1 && w <= p) ws[nw++] = p - w;
    else if (w >= W - 1 - p && w <= W - 2) ws[nw++] = 2 * (W - 1) - w + p;
    for (int a = 0; a < nh; ++a)
      for (int b = 0; b < nw; ++b) {
        float v[8];
        load8(g.g2.p + n * g.g2.sN + hs[a] * g.g2.sH + ws[b] * g.g2.sW + c0, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] += v[i];
      }
  }
}

__device__ __forceinline__ float act_fwd(float x, int act) {
  if (act == kActRelu) return fmaxf(x, 0.f);
  if (act == kActLeaky) return x > 0.f ? x : 0.2f * x;
  return x;
}
__device__ __forceinline__ float act_grad(float x, int act) {
  if (act == kActRelu) return x > 0.f ? 1.f : 0.f;
  if (act == kActLeaky) return x > 0.f ? 1.f : 0.2f;
  return 1.f;
}

inline unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------ layout
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int C, DevTensor dst) {
  const int HP = dst.H + 2 * dst.halo, WP = dst.W + 2 * dst.halo;
  const long long total = (long long)dst.N * HP * WP;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wp = idx % WP;
  const int hp = (idx / WP) % HP;
  const int n = idx / ((long long)WP * HP);
  const int h = reflect_idx(hp - dst.halo, dst.H), w = reflect_idx(wp - dst.halo, dst.W);
  bf16* o = dst.p + n * dst.sN + (hp - dst.halo) * dst.sH + (wp - dst.halo) * dst.sW;
  for (int c0 = 0; c0 < dst.C; c0 += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      v[i] = c < C ? src[(((long long)n * C + c) * dst.H + h) * dst.W + w] : 0.f;
    }
    store8(o + c0, v);
  }
}

__global__ void nhwc_to_nchw_kernel(DevTensor src, int C, float* __restrict__ dst) {
  const long long total = (long long)src.N * src.H * src.W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int w = idx % src.W;
  const int h = (idx / src.W) % src.H;
  const int n = idx / ((long long)src.W * src.H);
  const bf16* p = src.p + n * src.sN + h * src.sH + w * src.sW;
  for (int c = 0; c < C; ++c) dst[(((long long)n * C + c) * src.H + h) * src.W + w] = __bfloat162float(p[c]);
}

__global__ void fill_halo_kernel(DevTensor t) {
  ptx::pdl_wait();  // launched with programmatic stream serialization (see launch_pdl)
  if (kPwTrigger) ptx::pdl_launch_dependents();
  const int HP = t.H + 2 * t.halo, WP = t.W + 2 * t.halo, C8 = t.C / 8;
  const long long total = (long long)t.N * HP * WP * C8;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c0 = (idx % C8) * 8;
  long long r = idx / C8;
  const int wp = r % WP;
  r /= WP;
  const int hp = r % HP;
  const int n = r / HP;
  const int h = hp - t.halo, w = wp - t.halo;
  if (h >= 0 && h < t.H && w >= 0 && w < t.W) return;
  const int hs = reflect_idx(h, t.H), ws = reflect_idx(w, t.W);
  const uint4 v = *reinterpret_cast<const uint4*>(t.p + n * t.sN + hs * t.sH + ws * t.sW + c0);
  *reinterpret_cast<uint4*>(t.p + n * t.sN + h * t.sH + w * t.sW + c0) = v;
}

// ------------------------------------------------------------------------------------------ column sums
// Block = 256 threads = (C/8 channel lanes) x (rows of pixels).  MODE 0: (sum, sumsq) -> float2 stats;
// MODE 1: sum only of the first Cvalid channels -> float* (bias gradients).
template <int MODE>
__global__ void colsum_kernel(DevTensor y, float* __restrict__ out, int pix_per_block, int Cvalid) {
  ptx::pdl_wait();  // launched with programmatic stream serialization (see launch_pdl)
  if (kPwTrigger) ptx::pdl_launch_dependents();
  const int C8 = y.C / 8;
  const int lanes = C8 < 256 ? C8 : 256;
  const int rows = 256 / lanes;
  const int cl = threadIdx.x % lanes, pr = threadIdx.x / lanes;
  const int n = blockIdx.y;
  const int HW = y.H * y.W;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  __shared__ float red[256 * 16];
  for (int cb = cl; cb < C8; cb += lanes) {
    float s[8], ss[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = ss[i] = 0.f;
    if (pr < rows) {
      for (int p = p0 + pr; p < p1; p += rows) {
        const int h = p / y.W, w = p % y.W;
        float v[8];
        load8(y.p + n * y.sN + h * y.sH + w * y.sW + cb * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i] += v[i];
          if (MODE == 0) ss[i] += v[i] * v[i];
        }
      }
    }
    // reduce across pixel rows through shared memory
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      red[threadIdx.x * 16 + i] = s[i];
      red[threadIdx.x * 16 + 8 + i] = ss[i];
    }
    __syncthreads();
    if (pr == 0) {
      for (int r = 1; r < rows; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i] += red[(r * lanes + cl) * 16 + i];
          ss[i] += red[(r * lanes + cl) * 16 + 8 + i];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = cb * 8 + i;
        if (MODE == 0) {
          if ((i & 1) == 0) red_add_v4(out + ((long long)n * y.C + c) * 2, s[i], ss[i], s[i + 1], ss[i + 1]);
        } else if (c < Cvalid) {
          atomicAdd(out + c, s[i]);
        }
      }
    }
    __syncthreads();
  }
}

int pick_pix_per_block(int HW) {
  int ppb = (HW + 511) / 512;  // ~512 blocks per image: several CTAs per SM even at batch 1
  ppb = (ppb + 31) / 32 * 32;
  return ppb < 32 ? 32 : ppb;
}

// ------------------------------------------------------------------------------------------ IN apply
// Block = 256 threads = (C/8 channel lanes) x (pixel rows); a thread keeps ONE 8-channel chunk, so mean / rstd
// are computed once and reused over PPT consecutive pixels.  All PPT loads are issued before the first use
// (these kernels are latency-bound on small maps and bandwidth-bound on large ones: either way the number of
// 16-byte loads in flight per thread is what matters).
struct PixIter {  // walks consecutive pixels of a WP-wide raster without a division per pixel
  int hp, wp;
  __device__ __forceinline__ PixIter(int p, int WP) : hp(p / WP), wp(p - (p / WP) * WP) {}
  __device__ __forceinline__ void next(int WP) {
    if (++wp == WP) {
      wp = 0;
      ++hp;
    }
  }
};

// per-channel affine form of the normalisation: xhat = x * a + b with a = rstd, b = -mean * rstd
__device__ __forceinline__ void load_norm8(const float2* __restrict__ stats, long long idx, float inv, float (&a)[8],
                                           float (&b)[8]) {
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    const float4 st = __ldg(reinterpret_cast<const float4*>(stats + idx + i));
    const float m0 = st.x * inv, m1 = st.z * inv;
    a[i] = rsqrtf(fmaxf(st.y * inv - m0 * m0, 0.f) + 1e-5f);
    a[i + 1] = rsqrtf(fmaxf(st.w * inv - m1 * m1, 0.f) + 1e-5f);
    b[i] = -m0 * a[i];
    b[i + 1] = -m1 * a[i + 1];
  }
}

// Each block walks `ppb` consecutive pixels of one image: per iteration a thread issues PPT loads (x and, when
// present, the residual) before using any of them.
template <int PPT>
__global__ void __launch_bounds__(256)
in_apply_kernel(DevTensor y, const float2* __restrict__ stats, int act, DevTensor res, DevTensor out, int ppb) {
  ptx::pdl_wait();  // launched with programmatic stream serialization (see launch_pdl)
  if (kPwTrigger) ptx::pdl_launch_dependents();
  const int C8 = out.C / 8;
  const int lanes = C8 < 256 ? C8 : 256;
  const int rows = 256 / lanes;
  const int cl = threadIdx.x % lanes, pr = threadIdx.x / lanes;
  const int n = blockIdx.y;
  const int HP = out.H + 2 * out.halo, WP = out.W + 2 * out.halo;
  const int total = HP * WP;
  const float inv = 1.f / (float)(y.H * y.W);
  const int p0 = blockIdx.x * ppb, p1 = min(total, p0 + ppb);
  if (pr >= rows) return;
  for (int cb = cl; cb < C8; cb += lanes) {
    const int c0 = cb * 8;
    float a[8], b[8];
    load_norm8(stats, (long long)n * y.C + c0, inv, a, b);
#pragma unroll 1
    for (int pb = p0 + pr * PPT; pb < p1; pb += rows * PPT) {
      uint4 raw[PPT], rraw[PPT];
      long long ooff[PPT];
      bool ok[PPT];
      PixIter it(pb, WP);
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        ok[k] = pb + k < p1;
        const int h = reflect_idx(it.hp - out.halo, out.H), w = reflect_idx(it.wp - out.halo, out.W);
        ooff[k] = n * out.sN + (long long)(it.hp - out.halo) * out.sH + (long long)(it.wp - out.halo) * out.sW + c0;
        if (ok[k]) {
          raw[k] = *reinterpret_cast<const uint4*>(y.p + n * y.sN + h * y.sH + w * y.sW + c0);
          if (res.p != nullptr)
            rraw[k] = *reinterpret_cast<const uint4*>(res.p + n * res.sN + h * res.sH + w * res.sW + c0);
        }
        it.next(WP);
      }
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        if (!ok[k]) continue;
        float v[8];
        unpack8(raw[k], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = act_fwd(fmaf(v[i], a[i], b[i]), act);
        if (res.p != nullptr) {
          float rv[8];
          unpack8(rraw[k], rv);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] += rv[i];
        }
        store8(out.p + ooff[k], v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ IN backward
// Gradient gather shared by both backward kernels: the main source pixels are loaded for all PPT pixels first;
// the (rare) mirrored border contributions of a padded-domain gradient are added afterwards.
template <int PPT>
struct BwdLoads {
  uint4 yraw[PPT], g1raw[PPT], g2raw[PPT];
  int h[PPT], w[PPT];
  bool ok[PPT];
};

template <int PPT>
__device__ __forceinline__ void bwd_issue_loads(const DevTensor& y, const DevGrad& g, int n, int p_begin, int total,
                                                int c0, BwdLoads<PPT>& L) {
  PixIter it(p_begin, y.W);
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    L.ok[k] = p_begin + k < total;
    L.h[k] = it.hp;
    L.w[k] = it.wp;
    if (L.ok[k]) {
      L.yraw[k] = *reinterpret_cast<const uint4*>(y.p + n * y.sN + it.hp * y.sH + it.wp * y.sW + c0);
      if (g.g1.p != nullptr)
        L.g1raw[k] = *reinterpret_cast<const uint4*>(g.g1.p + n * g.g1.sN + it.hp * g.g1.sH + it.wp * g.g1.sW + c0);
      if (g.g2.p != nullptr)
        L.g2raw[k] = *reinterpret_cast<const uint4*>(g.g2.p + n * g.g2.sN + (it.hp + g.fold) * g.g2.sH +
                                                      (it.wp + g.fold) * g.g2.sW + c0);
    }
    it.next(y.W);
  }
}

// gradient w.r.t. the activation for pixel k (fp32), including the folded reflection-halo contributions
template <int PPT>
__device__ __forceinline__ void bwd_grad(const DevTensor& y, const DevGrad& g, int n, int c0, const BwdLoads<PPT>& L,
                                         int k, float (&out)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] = 0.f;
  if (g.g1.p != nullptr) {
    float v[8];
    unpack8(L.g1raw[k], v);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] += v[i];
  }
  if (g.g2.p != nullptr) {
    float v[8];
    unpack8(L.g2raw[k], v);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] += v[i];
    const int p = g.fold, h = L.h[k], w = L.w[k], H = y.H, W = y.W;
    const bool hb = (h >= 1 && h <= p) || (h >= H - 1 - p && h <= H - 2);
    const bool wb = (w >= 1 && w <= p) || (w >= W - 1 - p && w <= W - 2);
    if (hb || wb) {  // this pixel is the mirror image of one or three halo pixels
      const int hm = (h >= 1 && h <= p) ? p - h : 2 * (H - 1) - h + p;
      const int wm = (w >= 1 && w <= p) ? p - w : 2 * (W - 1) - w + p;
      const bf16* base = g.g2.p + n * g.g2.sN + c0;
      if (hb) {
        load8(base + hm * g.g2.sH + (w + p) * g.g2.sW, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] += v[i];
      }
      if (wb) {
        load8(base + (h + p) * g.g2.sH + wm * g.g2.sW, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] += v[i];
      }
      if (hb && wb) {
        load8(base + hm * g.g2.sH + wm * g.g2.sW, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] += v[i];
      }
    }
  }
}

template <int PPT>
__global__ void __launch_bounds__(256)
in_bwd_reduce_kernel(DevTensor y, const float2* __restrict__ stats, DevGrad g, int act, DevTensor da,
                     float* __restrict__ bstats, int ppb) {
  ptx::pdl_wait();  // launched with programmatic stream serialization (see launch_pdl)
  if (kPwTrigger) ptx::pdl_launch_dependents();
  const int C8 = y.C / 8;
  const int lanes = C8 < 256 ? C8 : 256;
  const int rows = 256 / lanes;
  const int cl = threadIdx.x % lanes, pr = threadIdx.x / lanes;
  const int n = blockIdx.y;
  const int HW = y.H * y.W;
  const float inv = 1.f / (float)HW;
  const int p0 = blockIdx.x * ppb, p1 = min(HW, p0 + ppb);
  __shared__ float red[256 * 16];
  for (int cb = cl; cb < C8; cb += lanes) {
    const int c0 = cb * 8;
    float s1[8], s2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
    if (pr < rows) {
      float a[8], b[8];
      load_norm8(stats, (long long)n * y.C + c0, inv, a, b);
#pragma unroll 1
      for (int pb = p0 + pr * PPT; pb < p1; pb += rows * PPT) {
        BwdLoads<PPT> L;
        bwd_issue_loads<PPT>(y, g, n, pb, p1, c0, L);
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
          if (!L.ok[k]) continue;
          float v[8], gr[8];
          unpack8(L.yraw[k], v);
          bwd_grad<PPT>(y, g, n, c0, L, k, gr);
          if (da.p != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) gr[i] = round_bf16(gr[i]);
            store8(da.p + n * da.sN + L.h[k] * da.sH + L.w[k] * da.sW + c0, gr);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float xh = fmaf(v[i], a[i], b[i]);
            const float dz = gr[i] * act_grad(xh, act);
            s1[i] += dz;
            s2[i] = fmaf(dz, xh, s2[i]);
          }
        }
      }
    }
    // reduce across the block's pixel rows through shared memory, then one vector atomic per channel pair
    // (the launcher keeps the number of blocks per image small: same-address atomics serialise in L2)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      red[threadIdx.x * 16 + i] = s1[i];
      red[threadIdx.x * 16 + 8 + i] = s2[i];
    }
    __syncthreads();
    if (pr == 0) {
      for (int r = 1; r < rows; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s1[i] += red[(r * lanes + cl) * 16 + i];
          s2[i] += red[(r * lanes + cl) * 16 + 8 + i];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const int c = cb * 8 + i;
        red_add_v4(bstats + ((long long)n * y.C + c) * 2, s1[i], s2[i], s1[i + 1], s2[i + 1]);
      }
    }
    __syncthreads();
  }
}

template <int PPT>
__global__ void __launch_bounds__(256)
in_bwd_apply_kernel(DevTensor y, const float2* __restrict__ stats, const float2* __restrict__ bstats, DevGrad g,
                    int act, DevTensor dy, int ppb) {
  ptx::pdl_wait();  // launched with programmatic stream serialization (see launch_pdl)
  if (kPwTrigger) ptx::pdl_launch_dependents();
  const int C8 = y.C / 8;
  const int lanes = C8 < 256 ? C8 : 256;
  const int rows = 256 / lanes;
  const int cl = threadIdx.x % lanes, pr = threadIdx.x / lanes;
  const int n = blockIdx.y;
  const int total = y.H * y.W;
  const float inv = 1.f / (float)total;
  const int p0 = blockIdx.x * ppb, p1 = min(total, p0 + ppb);
  if (pr >= rows) return;
  for (int cb = cl; cb < C8; cb += lanes) {
    const int c0 = cb * 8;
    // dy = rstd * (dz - mean(dz) - xhat * mean(dz * xhat)) = a * dz - c - d * xhat
    float a[8], b[8], c[8], d[8];
    load_norm8(stats, (long long)n * y.C + c0, inv, a, b);
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const float4 bs = __ldg(reinterpret_cast<const float4*>(bstats + (long long)n * y.C + c0 + i));
      c[i] = a[i] * bs.x * inv;
      d[i] = a[i] * bs.y * inv;
      c[i + 1] = a[i + 1] * bs.z * inv;
      d[i + 1] = a[i + 1] * bs.w * inv;
    }
#pragma unroll 1
    for (int pb = p0 + pr * PPT; pb < p1; pb += rows * PPT) {
      BwdLoads<PPT> L;
      bwd_issue_loads<PPT>(y, g, n, pb, p1, c0, L);
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        if (!L.ok[k]) continue;
        float v[8], gr[8];
        unpack8(L.yraw[k], v);
        bwd_grad<PPT>(y, g, n, c0, L, k, gr);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = fmaf(v[i], a[i], b[i]);
          const float dz = gr[i] * act_grad(xh, act);
          v[i] = fmaf(a[i], dz, -c[i]) - d[i] * xh;
        }
        store8(dy.p + n * dy.sN + L.h[k] * dy.sH + L.w[k] * dy.sW + c0, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ cp.async variants
// Large maps: the register versions above keep 4-6 16-byte loads per thread in flight at 2 blocks / SM (117-120
// registers), ~40 KB per SM, which tops out near 2.5 TB/s.  Here every thread prefetches its vectors for the next
// kAsyncStages-1 loop iterations into private shared-memory slots with cp.async (no registers held while in
// flight), so ~100 KB per SM are in flight.  Used when a thread has >= 3 iterations (pick_async()).
constexpr int kAsyncStages = 3;
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ptx::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// slot of (stage, vector kind v < NV, pixel k < PPT) of this thread: consecutive threads -> consecutive 16 bytes
template <int PPT, int NV>
__device__ __forceinline__ uint4* async_slot(uint4* ring, int stage, int v, int k) {
  return ring + ((stage * NV + v) * PPT + k) * 256 + threadIdx.x;
}

template <int PPT>
__device__ __forceinline__ void bwd_async_issue(const DevTensor& y, const DevGrad& g, int n, int pb, int p1, int c0,
                                                uint4* ring, int stage) {
  if (pb < p1) {
    PixIter it(pb, y.W);
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      if (pb + k < p1) {
        cp_async16(async_slot<PPT, 3>(ring, stage, 0, k), y.p + n * y.sN + it.hp * y.sH + it.wp * y.sW + c0);
        if (g.g1.p != nullptr)
          cp_async16(async_slot<PPT, 3>(ring, stage, 1, k), g.g1.p + n * g.g1.sN + it.hp * g.g1.sH + it.wp * g.g1.sW + c0);
        if (g.g2.p != nullptr)
          cp_async16(async_slot<PPT, 3>(ring, stage, 2, k),
                     g.g2.p + n * g.g2.sN + (it.hp + g.fold) * g.g2.sH + (it.wp + g.fold) * g.g2.sW + c0);
      }
      it.next(y.W);
    }
  }
  cp_async_commit();  // one group per iteration, empty past the end, so wait_group counts stay uniform
}

template <int PPT>
__device__ __forceinline__ void bwd_async_collect(const DevTensor& y, const DevGrad& g, int pb, int p1, uint4* ring,
                                                  int stage, BwdLoads<PPT>& L) {
  PixIter it(pb, y.W);
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    L.ok[k] = pb + k < p1;
    L.h[k] = it.hp;
    L.w[k] = it.wp;
    if (L.ok[k]) {
      L.yraw[k] = *async_slot<PPT, 3>(ring, stage, 0, k);
      if (g.g1.p != nullptr) L.g1raw[k] = *async_slot<PPT, 3>(ring, stage, 1, k);
      if (g.g2.p != nullptr) L.g2raw[k] = *async_slot<PPT, 3>(ring, stage, 2, k);
    }
    it.next(y.W);
  }
}

template <int PPT>
__global__ void __launch_bounds__(256)
in_bwd_apply_async_kernel(DevTensor y, const float2* __restrict__ stats, const float2* __restrict__ bstats, DevGrad g,
                          int act, DevTensor dy, int ppb) {
  ptx::pdl_wait();
  extern __shared__ uint4 ring[];
  const int C8 = y.C / 8;
  const int lanes = C8 < 256 ? C8 : 256;
  const int rows = 256 / lanes;
  const int cl = threadIdx.x % lanes, pr = threadIdx.x / lanes;
  const int n = blockIdx.y;
  const int total = y.H * y.W;
  const float inv = 1.f / (float)total;
  const int p0 = blockIdx.x * ppb, p1 = min(total, p0 + ppb);
  const int step = rows * PPT;
  if (pr >= rows) return;
  for (int cb = cl; cb < C8; cb += lanes) {
    const int c0 = cb * 8;
    const int pfirst = p0 + pr * PPT;
#pragma unroll
    for (int s = 0; s < kAsyncStages - 1; ++s) bwd_async_issue<PPT>(y, g, n, pfirst + s * step, p1, c0, ring, s);
    float a[8], b[8], c[8], d[8];
    load_norm8(stats, (long long)n * y.C + c0, inv, a, b);
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const float4 bs = __ldg(reinterpret_cast<const float4*>(bstats + (long long)n * y.C + c0 + i));
      c[i] = a[i] * bs.x * inv;
      d[i] = a[i] * bs.y * inv;
      c[i + 1] = a[i + 1] * bs.z * inv;
      d[i + 1] = a[i + 1] * bs.w * inv;
    }
    int stage = 0;
#pragma unroll 1
    for (int pb = pfirst; pb < p1; pb += step) {
      int pre = stage + kAsyncStages - 1;
      if (pre >= kAsyncStages) pre -= kAsyncStages;
      bwd_async_issue<PPT>(y, g, n, pb + (kAsyncStages - 1) * step, p1, c0, ring, pre);
      cp_async_wait<kAsyncStages - 1>();
      BwdLoads<PPT> L;
      bwd_async_collect<PPT>(y, g, pb, p1, ring, stage, L);
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        if (!L.ok[k]) continue;
        float v[8], gr[8];
        unpack8(L.yraw[k], v);
        bwd_grad<PPT>(y, g, n, c0, L, k, gr);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = fmaf(v[i], a[i], b[i]);
          const float dz = gr[i] * act_grad(xh, act);
          v[i] = fmaf(a[i], dz, -c[i]) - d[i] * xh;
        }
        store8(dy.p + n * dy.sN + L.h[k] * dy.sH + L.w[k] * dy.sW + c0, v);
      }
      if (++stage == kAsyncStages) stage = 0;
    }
    cp_async_wait<0>();
  }
}

template <int PPT>
__global__ void __launch_bounds__(256)
in_bwd_reduce_async_kernel(DevTensor y, const float2* __restrict__ stats, DevGrad g, int act, DevTensor da,
                           float* __restrict__ bstats, int ppb) {
  ptx::pdl_wait();
  extern __shared__ uint4 ring[];
  __shared__ float red[256 * 16];
  const int C8 = y.C / 8;
  const int lanes = C8 < 256 ? C8 : 256;
  const int rows = 256 / lanes;
  const int cl = threadIdx.x % lanes, pr = threadIdx.x / lanes;
  const int n = blockIdx.y;
  const int HW = y.H * y.W;
  const float inv = 1.f / (float)HW;
  const int p0 = blockIdx.x * ppb, p1 = min(HW, p0 + ppb);
  const int step = rows * PPT;
  for (int cb = cl; cb < C8; cb += lanes) {
    const int c0 = cb * 8;
    float s1[8], s2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
    if (pr < rows) {
      const int pfirst = p0 + pr * PPT;
#pragma unroll
      for (int s = 0; s < kAsyncStages - 1; ++s) bwd_async_issue<PPT>(y, g, n, pfirst + s * step, p1, c0, ring, s);
      float a[8], b[8];
      load_norm8(stats, (long long)n * y.C + c0, inv, a, b);
      int stage = 0;
#pragma unroll 1
      for (int pb = pfirst; pb < p1; pb += step) {
        int pre = stage + kAsyncStages - 1;
        if (pre >= kAsyncStages) pre -= kAsyncStages;
        bwd_async_issue<PPT>(y, g, n, pb + (kAsyncStages - 1) * step, p1, c0, ring, pre);
        cp_async_wait<kAsyncStages - 1>();
        BwdLoads<PPT> L;
        bwd_async_collect<PPT>(y, g, pb, p1, ring, stage, L);
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
          if (!L.ok[k]) continue;
          float v[8], gr[8];
          unpack8(L.yraw[k], v);
          bwd_grad<PPT>(y, g, n, c0, L, k, gr);
          if (da.p != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) gr[i] = round_bf16(gr[i]);
            store8(da.p + n * da.sN + L.h[k] * da.sH + L.w[k] * da.sW + c0, gr);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float xh = fmaf(v[i], a[i], b[i]);
            const float dz = gr[i] * act_grad(xh, act);
            s1[i] += dz;
            s2[i] = fmaf(dz, xh, s2[i]);
          }
        }
        if (++stage == kAsyncStages) stage = 0;
      }
      cp_async_wait<0>();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      red[threadIdx.x * 16 + i] = s1[i];
      red[threadIdx.x * 16 + 8 + i] = s2[i];
    }
    __syncthreads();
    if (pr == 0) {
      for (int r = 1; r < rows; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s1[i] += red[(r * lanes + cl) * 16 + i];
          s2[i] += red[(r * lanes + cl) * 16 + 8 + i];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const int c = cb * 8 + i;
        red_add_v4(bstats + ((long long)n * y.C + c) * 2, s1[i], s2[i], s1[i + 1], s2[i + 1]);
      }
    }
    __syncthreads();
  }
}

template <int PPT>
__global__ void __launch_bounds__(256)
in_apply_async_kernel(DevTensor y, const float2* __restrict__ stats, int act, DevTensor res, DevTensor out, int ppb) {
  ptx::pdl_wait();
  extern __shared__ uint4 ring[];
  const int C8 = out.C / 8;
  const int lanes = C8 < 256 ? C8 : 256;
  const int rows = 256 / lanes;
  const int cl = threadIdx.x % lanes, pr = threadIdx.x / lanes;
  const int n = blockIdx.y;
  const int HP = out.H + 2 * out.halo, WP = out.W + 2 * out.halo;
  const int total = HP * WP;
  const float inv = 1.f / (float)(y.H * y.W);
  const int p0 = blockIdx.x * ppb, p1 = min(total, p0 + ppb);
  const int step = rows * PPT;
  if (pr >= rows) return;
  auto issue = [&](int pb, int c0, int stage) {
    if (pb < p1) {
      PixIter it(pb, WP);
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        if (pb + k < p1) {
          const int h = reflect_idx(it.hp - out.halo, out.H), w = reflect_idx(it.wp - out.halo, out.W);
          cp_async16(async_slot<PPT, 2>(ring, stage, 0, k), y.p + n * y.sN + h * y.sH + w * y.sW + c0);
          if (res.p != nullptr)
            cp_async16(async_slot<PPT, 2>(ring, stage, 1, k), res.p + n * res.sN + h * res.sH + w * res.sW + c0);
        }
        it.next(WP);
      }
    }
    cp_async_commit();
  };
  for (int cb = cl; cb < C8; cb += lanes) {
    const int c0 = cb * 8;
    const int pfirst = p0 + pr * PPT;
#pragma unroll
    for (int s = 0; s < kAsyncStages - 1; ++s) issue(pfirst + s * step, c0, s);
    float a[8], b[8];
    load_norm8(stats, (long long)n * y.C + c0, inv, a, b);
    int stage = 0;
#pragma unroll 1
    for (int pb = pfirst; pb < p1; pb += step) {
      int pre = stage + kAsyncStages - 1;
      if (pre >= kAsyncStages) pre -= kAsyncStages;
      issue(pb + (kAsyncStages - 1) * step, c0, pre);
      cp_async_wait<kAsyncStages - 1>();
      PixIter it(pb, WP);
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        if (pb + k < p1) {
          float v[8];
          unpack8(*async_slot<PPT, 2>(ring, stage, 0, k), v);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = act_fwd(fmaf(v[i], a[i], b[i]), act);
          if (res.p != nullptr) {
            float rv[8];
            unpack8(*async_slot<PPT, 2>(ring, stage, 1, k), rv);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += rv[i];
          }
          store8(out.p + n * out.sN + (long long)(it.hp - out.halo) * out.sH + (long long)(it.wp - out.halo) * out.sW + c0, v);
        }
        it.next(WP);
      }
      if (++stage == kAsyncStages) stage = 0;
    }
    cp_async_wait<0>();
  }
}

// ------------------------------------------------------------------------------------------ fused IN backward
// One thread-block CLUSTER of kInCluster CTAs owns (image n, CG consecutive channels) of a map of at most 4096 pixels
// (the residual stream at 256x256, the discriminator's inner layers): every thread keeps its ITEMS (pixel, 8-channel
// vector) elements of y and of the assembled gradient in REGISTERS, the two reductions (sum dz, sum dz * xhat) go
// thread -> warp shuffles -> CTA -> cluster (partials read through distributed shared memory in rank order), and the
// apply pass runs from the registers.  Versus in_bwd_reduce + in_bwd_apply: one launch instead of two on the backward
// chain, 6 instead of 10 bytes per element of traffic, all loads of a thread in flight at once, no atomics (the
// result is deterministic).  A thread owns items i * 256 + tid, so its channel vector (tid % OCT) is fixed.
constexpr int kInCluster = 8;

template <int CG, int ITEMS>
__global__ void __launch_bounds__(256)
in_bwd_fused_kernel(DevTensor y, const float2* __restrict__ stats, DevGrad g, int act, DevTensor da, DevTensor dy,
                    int pix_per_cta) {
  constexpr int OCT = CG / 8;  // 8-channel vectors per pixel handled by this cluster
  __shared__ float warp_part[8][OCT][16];
  __shared__ float cta_part[2 * CG];
  __shared__ float total[2 * CG];
  ptx::pdl_wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = ptx::cluster_ctarank();
  const int group = blockIdx.x / kInCluster, n = blockIdx.y;
  const int oct = tid % OCT;
  const int c0 = group * CG + oct * 8;
  const int HW = y.H * y.W;
  const float inv = 1.f / (float)HW;
  const int p_begin = (int)rank * pix_per_cta, p_end = min(HW, p_begin + pix_per_cta);
  // ---- phase 1a: issue every load of this thread
  uint4 yraw[ITEMS], g1raw[ITEMS], g2raw[ITEMS];
  int ph[ITEMS], pw[ITEMS];
  bool ok[ITEMS];
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int p = p_begin + (it * 256 + tid) / OCT;
    ok[it] = p < p_end;
    ph[it] = p / y.W;
    pw[it] = p - ph[it] * y.W;
    if (ok[it]) {
      yraw[it] = *reinterpret_cast<const uint4*>(y.p + n * y.sN + ph[it] * y.sH + pw[it] * y.sW + c0);
      if (g.g1.p != nullptr)
        g1raw[it] = *reinterpret_cast<const uint4*>(g.g1.p + n * g.g1.sN + ph[it] * g.g1.sH + pw[it] * g.g1.sW + c0);
      if (g.g2.p != nullptr)
        g2raw[it] = *reinterpret_cast<const uint4*>(g.g2.p + n * g.g2.sN + (ph[it] + g.fold) * g.g2.sH +
                                                    (pw[it] + g.fold) * g.g2.sW + c0);
    }
  }
  float a[8], b[8];
  load_norm8(stats, (long long)n * y.C + c0, inv, a, b);
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
  // ---- phase 1b: assemble the gradient (mirrored halo contributions of a padded-domain source are rare), sums
  float gr[ITEMS][8];
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) gr[it][i] = 0.f;
    if (!ok[it]) continue;
    float v[8];
    if (g.g1.p != nullptr) {
      unpack8(g1raw[it], v);
#pragma unroll
      for (int i = 0; i < 8; ++i) gr[it][i] += v[i];
    }
    if (g.g2.p != nullptr) {
      unpack8(g2raw[it], v);
#pragma unroll
      for (int i = 0; i < 8; ++i) gr[it][i] += v[i];
      const int p = g.fold, h = ph[it], w = pw[it], H = y.H, W = y.W;
      const bool hb = (h >= 1 && h <= p) || (h >= H - 1 - p && h <= H - 2);
      const bool wb = (w >= 1 && w <= p) || (w >= W - 1 - p && w <= W - 2);
      if (hb || wb) {
        const int hm = (h >= 1 && h <= p) ? p - h : 2 * (H - 1) - h + p;
        const int wm = (w >= 1 && w <= p) ? p - w : 2 * (W - 1) - w + p;
        const bf16* base = g.g2.p + n * g.g2.sN + c0;
        if (hb) {
          load8(base + hm * g.g2.sH + (w + p) * g.g2.sW, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) gr[it][i] += v[i];
        }
        if (wb) {
          load8(base + (h + p) * g.g2.sH + wm * g.g2.sW, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) gr[it][i] += v[i];
        }
        if (hb && wb) {
          load8(base + hm * g.g2.sH + wm * g.g2.sW, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) gr[it][i] += v[i];
        }
      }
    }
    if (da.p != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) gr[it][i] = round_bf16(gr[it][i]);
      store8(da.p + n * da.sN + ph[it] * da.sH + pw[it] * da.sW + c0, gr[it]);
    }
    unpack8(yraw[it], v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = fmaf(v[i], a[i], b[i]);
      const float dz = gr[it][i] * act_grad(xh, act);
      s1[i] += dz;
      s2[i] = fmaf(dz, xh, s2[i]);
    }
  }
  // ---- reductions: lanes with the same channel vector (lane % OCT) inside a warp, the 8 warps, the cluster's CTAs
#pragma unroll
  for (int off = 16; off >= OCT; off >>= 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], off);
      s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], off);
    }
  }
  if (lane < OCT) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      warp_part[warp][lane][i] = s1[i];
      warp_part[warp][lane][8 + i] = s2[i];
    }
  }
  __syncthreads();
  if (tid < 2 * CG) {  // entry j: which = j / CG (0: sum dz, 1: sum dz * xhat), channel = j % CG
    const int which = tid / CG, ch = tid % CG;
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) t += warp_part[wv][ch / 8][which * 8 + (ch & 7)];
    cta_part[tid] = t;
  }
  ptx::cluster_sync_all();  // every CTA's partials are written and visible cluster-wide
  if (tid < 2 * CG) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < kInCluster; ++r) t += ptx::ld_dsmem_f32(&cta_part[tid], (uint32_t)r);
    total[tid] = t;
  }
  __syncthreads();
  // ---- phase 2: dy = rstd * (dz - mean(dz) - xhat * mean(dz * xhat)) = a * dz - c - d * xhat, from the registers
  float c[8], d[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    c[i] = a[i] * total[oct * 8 + i] * inv;
    d[i] = a[i] * total[CG + oct * 8 + i] * inv;
  }
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    if (!ok[it]) continue;
    float v[8];
    unpack8(yraw[it], v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = fmaf(v[i], a[i], b[i]);
      const float dz = gr[it][i] * act_grad(xh, act);
      v[i] = fmaf(a[i], dz, -c[i]) - d[i] * xh;
    }
    store8(dy.p + n * dy.sN + ph[it] * dy.sH + pw[it] * dy.sW + c0, v);
  }
  ptx::cluster_sync_all();  // no CTA may exit while a peer can still read its partial sums
}

// ------------------------------------------------------------------------------------------ row-streaming variants
// The kernels above walk a flat pixel range with 64-bit (n, h, w) addressing, runtime activation / gradient-source
// switches and per-pixel border tests: ncu (profiles/r02_o_ncu_pointwise_b8.txt) shows them ISSUE-bound, not memory-
// bound -- 55-62 % issue-slot utilisation at 25 % occupancy (117-120 registers), 1.6-3 TB/s, ~500 SASS instructions per
// 16-byte vector.  Here a block owns whole image rows: an interior row of any tensor is contiguous (W * C elements),
// thread t handles the vectors t, t + 256, ... of the row, so its channel vector is fixed (C divides 2048), the
// addresses inside a row are one 32-bit multiply-add, and activation / gradient sources / residual are template
// parameters.  The reflect-halo writer (apply) and the halo fold (backward) touch the few border pixels in a
// separate, rarely taken branch.
template <int ACT>
__device__ __forceinline__ float act_fwd_t(float x) {
  if (ACT == kActRelu) return fmaxf(x, 0.f);
  if (ACT == kActLeaky) return x > 0.f ? x : 0.2f * x;
  return x;
}
template <int ACT>
__device__ __forceinline__ float act_grad_t(float x) {
  if (ACT == kActRelu) return x > 0.f ? 1.f : 0.f;
  if (ACT == kActLeaky) return x > 0.f ? 1.f : 0.2f;
  return 1.f;
}
// mirror partner of interior index i inside a reflect halo of width p (interior extent n): -1 when i has none
__device__ __forceinline__ bool mirror_of(int i, int n, int p, int* m) {
  if (i >= 1 && i <= p) {
    *m = -i;
    return true;
  }
  if (i >= n - 1 - p && i <= n - 2) {
    *m = 2 * (n - 1) - i;
    return true;
  }
  return false;
}

struct RowSpan {  // rows [h0, h1) of image n and the item range [j0, j1) of this block inside every row
  int h0, h1, j0, j1;
};
// grid.x = ceil(H / rpb) * jsplit: a block takes rpb rows (jsplit == 1) or a 1 / jsplit slice of one row
__device__ __forceinline__ RowSpan row_span(int H, int vpr, int rpb, int jsplit) {
  RowSpan r;
  const int u = blockIdx.x / jsplit, js = blockIdx.x - u * jsplit;
  r.h0 = u * rpb;
  r.h1 = min(H, r.h0 + rpb);
  const int J = (vpr + 255) >> 8;
  const int jper = (J + jsplit - 1) / jsplit;
  r.j0 = js * jper;
  r.j1 = min(J, r.j0 + jper);
  return r;
}

template <int ACT, bool RES, int PPT>
__global__ void __launch_bounds__(256, 3)
in_apply_rows_kernel(DevTensor y, const float2* __restrict__ stats, DevTensor res, DevTensor out, int vpr, int rpb,
                     int jsplit) {
  ptx::pdl_wait();
  if (kPwTrigger) ptx::pdl_launch_dependents();
  const int t = threadIdx.x, n = blockIdx.y;
  const int C = out.C, H = out.H, W = out.W, p = out.halo;
  const int lanes = C >> 3, rows = 256 / lanes, prow = t / lanes;
  const int c0 = (t - prow * lanes) * 8;
  const RowSpan sp = row_span(H, vpr, rpb, jsplit);
  float a[8], b[8];
  load_norm8(stats, (long long)n * C + c0, 1.f / (float)(H * W), a, b);
  for (int h = sp.h0; h < sp.h1; ++h) {
    const bf16* yrow = y.p + n * y.sN + h * y.sH;
    const bf16* rrow = RES ? res.p + n * res.sN + h * res.sH : nullptr;
    bf16* orow = out.p + n * out.sN + h * out.sH;
    int hm = 0;
    const bool hb = p > 0 && mirror_of(h, H, p, &hm);
    bf16* mrow = out.p + n * out.sN + (long long)hm * out.sH;
#pragma unroll 1
    for (int j = sp.j0; j < sp.j1; j += PPT) {
      uint4 raw[PPT], rr[PPT];
      bool ok[PPT];
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const int v = t + ((j + k) << 8);
        ok[k] = (j + k < sp.j1) && v < vpr;
        if (ok[k]) {
          raw[k] = *reinterpret_cast<const uint4*>(yrow + v * 8);
          if (RES) rr[k] = *reinterpret_cast<const uint4*>(rrow + v * 8);
        }
      }
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        if (!ok[k]) continue;
        float v[8];
        unpack8(raw[k], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = act_fwd_t<ACT>(fmaf(v[i], a[i], b[i]));
        if (RES) {
          float rv[8];
          unpack8(rr[k], rv);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] += rv[i];
        }
        const int w = prow + (j + k) * rows;
        const int eo = (t + ((j + k) << 8)) * 8;  // == w * C + c0
        store8(orow + eo, v);
        if (p > 0) {  // this pixel is the mirror image of up to three halo pixels
          int wm = 0;
          const bool wb = mirror_of(w, W, p, &wm);
          if (wb) store8(orow + wm * C + c0, v);
          if (hb) {
            store8(mrow + eo, v);
            if (wb) store8(mrow + wm * C + c0, v);
          }
        }
      }
    }
  }
}

// gradient w.r.t. the activation of one item: g1 + g2 (+ the mirrored halo pixels of the padded-domain source g2)
template <bool G1, bool G2>
__device__ __forceinline__ void rows_grad(const uint4& g1raw, const uint4& g2raw, const bf16* g2img, long long g2sH, int C,
                                          int c0, int h, int w, int H, int W, int p, bool hb, int hm, float (&gr)[8]) {
  if (G1) {
    unpack8(g1raw, gr);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) gr[i] = 0.f;
  }
  if (G2) {
    float v[8];
    unpack8(g2raw, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) gr[i] += v[i];
    const bool wb = (w >= 1 && w <= p) || (w >= W - 1 - p && w <= W - 2);
    if (hb || wb) {
      const int wm = (w >= 1 && w <= p) ? p - w : 2 * (W - 1) - w + p;  // padded-domain column of the mirror
      if (hb) {
        load8(g2img + hm * g2sH + (w + p) * C + c0, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) gr[i] += v[i];
      }
      if (wb) {
        load8(g2img + (h + p) * g2sH + wm * C + c0, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) gr[i] += v[i];
      }
      if (hb && wb) {
        load8(g2img + hm * g2sH + wm * C + c0, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) gr[i] += v[i];
      }
    }
  }
}

template <int ACT, bool G1, bool G2, bool DA, int PPT>
__global__ void __launch_bounds__(256, 3)
in_bwd_reduce_rows_kernel(DevTensor y, const float2* __restrict__ stats, DevGrad g, DevTensor da,
                          float* __restrict__ bstats, int vpr, int rpb) {
  ptx::pdl_wait();
  if (kPwTrigger) ptx::pdl_launch_dependents();
  __shared__ float red[256 * 16];
  const int t = threadIdx.x, n = blockIdx.y;
  const int C = y.C, H = y.H, W = y.W, p = g.fold;
  const int lanes = C >> 3, rows = 256 / lanes, prow = t / lanes, cl = t - prow * lanes;
  const int c0 = cl * 8;
  const RowSpan sp = row_span(H, vpr, rpb, 1);
  float a[8], b[8], s1[8], s2[8];
  load_norm8(stats, (long long)n * C + c0, 1.f / (float)(H * W), a, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
  const bf16* g2img = G2 ? g.g2.p + n * g.g2.sN : nullptr;
  for (int h = sp.h0; h < sp.h1; ++h) {
    const bf16* yrow = y.p + n * y.sN + h * y.sH;
    const bf16* g1row = G1 ? g.g1.p + n * g.g1.sN + h * g.g1.sH : nullptr;
    const bf16* g2row = G2 ? g2img + (h + p) * g.g2.sH + p * C : nullptr;
    bf16* darow = DA ? da.p + n * da.sN + h * da.sH : nullptr;
    const bool hb = G2 && ((h >= 1 && h <= p) || (h >= H - 1 - p && h <= H - 2));
    const int hm = (h >= 1 && h <= p) ? p - h : 2 * (H - 1) - h + p;  // padded-domain row of the mirror
#pragma unroll 1
    for (int j = sp.j0; j < sp.j1; j += PPT) {
      uint4 yraw[PPT], g1raw[PPT], g2raw[PPT];
      bool ok[PPT];
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const int v = t + ((j + k) << 8);
        ok[k] = (j + k < sp.j1) && v < vpr;
        if (ok[k]) {
          yraw[k] = *reinterpret_cast<const uint4*>(yrow + v * 8);
          if (G1) g1raw[k] = *reinterpret_cast<const uint4*>(g1row + v * 8);
          if (G2) g2raw[k] = *reinterpret_cast<const uint4*>(g2row + v * 8);
        }
      }
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        if (!ok[k]) continue;
        float v[8], gr[8];
        rows_grad<G1, G2>(g1raw[k], g2raw[k], g2img, g.g2.sH, C, c0, h, prow + (j + k) * rows, H, W, p, hb, hm, gr);
        if (DA) {
#pragma unroll
          for (int i = 0; i < 8; ++i) gr[i] = round_bf16(gr[i]);
          store8(darow + (t + ((j + k) << 8)) * 8, gr);
        }
        unpack8(yraw[k], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = fmaf(v[i], a[i], b[i]);
          const float dz = gr[i] * act_grad_t<ACT>(xh);
          s1[i] += dz;
          s2[i] = fmaf(dz, xh, s2[i]);
        }
      }
    }
  }
  // the block's pixel slots that share a channel vector -> shared memory -> one vector atomic per channel pair
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[t * 16 + i] = s1[i];
    red[t * 16 + 8 + i] = s2[i];
  }
  __syncthreads();
  if (prow == 0) {
    for (int r = 1; r < rows; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s1[i] += red[(r * lanes + cl) * 16 + i];
        s2[i] += red[(r * lanes + cl) * 16 + 8 + i];
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i += 2)
      red_add_v4(bstats + ((long long)n * C + c0 + i) * 2, s1[i], s2[i], s1[i + 1], s2[i + 1]);
  }
}

template <int ACT, bool G1, bool G2, int PPT>
__global__ void __launch_bounds__(256, 3)
in_bwd_apply_rows_kernel(DevTensor y, const float2* __restrict__ stats, const float2* __restrict__ bstats, DevGrad g,
                         DevTensor dy, int vpr, int rpb, int jsplit) {
  ptx::pdl_wait();
  if (kPwTrigger) ptx::pdl_launch_dependents();
  const int t = threadIdx.x, n = blockIdx.y;
  const int C = y.C, H = y.H, W = y.W, p = g.fold;
  const int lanes = C >> 3, rows = 256 / lanes, prow = t / lanes;
  const int c0 = (t - prow * lanes) * 8;
  const RowSpan sp = row_span(H, vpr, rpb, jsplit);
  // dy = rstd * (dz - mean(dz) - xhat * mean(dz * xhat)) = a * dz - c - d * xhat
  float a[8], b[8], c[8], d[8];
  const float inv = 1.f / (float)(H * W);
  load_norm8(stats, (long long)n * C + c0, inv, a, b);
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    const float4 bs = __ldg(reinterpret_cast<const float4*>(bstats + (long long)n * C + c0 + i));
    c[i] = a[i] * bs.x * inv;
    d[i] = a[i] * bs.y * inv;
    c[i + 1] = a[i + 1] * bs.z * inv;
    d[i + 1] = a[i + 1] * bs.w * inv;
  }
  const bf16* g2img = G2 ? g.g2.p + n * g.g2.sN : nullptr;
  for (int h = sp.h0; h < sp.h1; ++h) {
    const bf16* yrow = y.p + n * y.sN + h * y.sH;
    const bf16* g1row = G1 ? g.g1.p + n * g.g1.sN + h * g.g1.sH : nullptr;
    const bf16* g2row = G2 ? g2img + (h + p) * g.g2.sH + p * C : nullptr;
    bf16* dyrow = dy.p + n * dy.sN + h * dy.sH;
    const bool hb = G2 && ((h >= 1 && h <= p) || (h >= H - 1 - p && h <= H - 2));
    const int hm = (h >= 1 && h <= p) ? p - h : 2 * (H - 1) - h + p;
#pragma unroll 1
    for (int j = sp.j0; j < sp.j1; j += PPT) {
      uint4 yraw[PPT], g1raw[PPT], g2raw[PPT];
      bool ok[PPT];
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const int v = t + ((j + k) << 8);
        ok[k] = (j + k < sp.j1) && v < vpr;
        if (ok[k]) {
          yraw[k] = *reinterpret_cast<const uint4*>(yrow + v * 8);
          if (G1) g1raw[k] = *reinterpret_cast<const uint4*>(g1row + v * 8);
          if (G2) g2raw[k] = *reinterpret_cast<const uint4*>(g2row + v * 8);
        }
      }
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        if (!ok[k]) continue;
        float v[8], gr[8];
        rows_grad<G1, G2>(g1raw[k], g2raw[k], g2img, g.g2.sH, C, c0, h, prow + (j + k) * rows, H, W, p, hb, hm, gr);
        unpack8(yraw[k], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = fmaf(v[i], a[i], b[i]);
          const float dz = gr[i] * act_grad_t<ACT>(xh);
          v[i] = fmaf(a[i], dz, -c[i]) - d[i] * xh;
        }
        store8(dyrow + (t + ((j + k) << 8)) * 8, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ bulk-copy variants
// Same decomposition as the row-streaming kernels, but the inputs of a segment (one image row, or a slice of `segv`
// <= kBulkVec vectors of it) arrive by ONE cp.async.bulk per tensor issued by one thread: all of the block's reads are
// in flight from its first instruction (the register kernels above issue them in 2-4 dependent batches per thread,
// so a short-lived block spends most of its life ramping up), nothing is held in registers while in flight, and three
// blocks per SM keep up to 3 x 32-96 KB outstanding.  Threads then read their vectors from shared memory.
constexpr int kBulkVec = 2048;  // at most this many 16-byte vectors per segment and tensor (32 KB); multiples of 256

struct Seg {  // segment s of this block -> (row, first vector, vector count)
  int h, v0, nv;
};
__device__ __forceinline__ Seg seg_of(int s, int nslice, int vpr, int segv) {
  Seg g;
  g.h = s / nslice;
  g.v0 = (s - g.h * nslice) * segv;
  g.nv = min(segv, vpr - g.v0);
  return g;
}

template <int ACT, bool RES>
__global__ void __launch_bounds__(256, 3)
in_apply_bulk_kernel(DevTensor y, const float2* __restrict__ stats, DevTensor res, DevTensor out, int vpr, int nslice, int segv) {
  extern __shared__ __align__(128) uint8_t bulk_smem[];
  __shared__ uint64_t bar;
  uint4* sy = reinterpret_cast<uint4*>(bulk_smem);
  uint4* sr = sy + segv;
  const int t = threadIdx.x, n = blockIdx.y;
  const int C = out.C, H = out.H, W = out.W, p = out.halo;
  const int lanes = C >> 3, lsh = 31 - __clz(lanes);
  const int c0 = (t & (lanes - 1)) * 8;
  const Seg sg = seg_of(blockIdx.x, nslice, vpr, segv);
  const int h = sg.h;
  if (t == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  ptx::pdl_wait();
  if (kPwTrigger) ptx::pdl_launch_dependents();
  if (t == 0) {
    ptx::mbar_arrive_expect_tx(&bar, (uint32_t)sg.nv * 16u * (RES ? 2u : 1u));
    ptx::bulk_load(sy, y.p + n * y.sN + h * y.sH + (long long)sg.v0 * 8, (uint32_t)sg.nv * 16u, &bar);
    if (RES) ptx::bulk_load(sr, res.p + n * res.sN + h * res.sH + (long long)sg.v0 * 8, (uint32_t)sg.nv * 16u, &bar);
  }
  float a[8], b[8];
  load_norm8(stats, (long long)n * C + c0, 1.f / (float)(H * W), a, b);
  bf16* orow = out.p + n * out.sN + h * out.sH;
  int hm = 0;
  const bool hb = p > 0 && mirror_of(h, H, p, &hm);
  bf16* mrow = out.p + n * out.sN + (long long)hm * out.sH;
  ptx::mbar_wait(&bar, 0);
#pragma unroll 2
  for (int vi = t; vi < sg.nv; vi += 256) {
    float v[8];
    unpack8(sy[vi], v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = act_fwd_t<ACT>(fmaf(v[i], a[i], b[i]));
    if (RES) {
      float rv[8];
      unpack8(sr[vi], rv);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += rv[i];
    }
    const int gv = sg.v0 + vi;   // vector index inside the row
    const int w = gv >> lsh;
    store8(orow + gv * 8, v);
    if (p > 0) {  // this pixel is the mirror image of up to three halo pixels
      int wm = 0;
      const bool wb = mirror_of(w, W, p, &wm);
      if (wb) store8(orow + wm * C + c0, v);
      if (hb) {
        store8(mrow + gv * 8, v);
        if (wb) store8(mrow + wm * C + c0, v);
      }
    }
  }
}

// issue the bulk copies of one backward segment (y, g1, g2 rows) -- one thread
template <bool G1, bool G2>
__device__ __forceinline__ void bwd_bulk_issue(const DevTensor& y, const DevGrad& g, int n, const Seg& sg, uint4* sy,
                                               uint4* sg1, uint4* sg2, uint64_t* bar) {
  const uint32_t bytes = (uint32_t)sg.nv * 16u;
  ptx::mbar_arrive_expect_tx(bar, bytes * (1u + (G1 ? 1u : 0u) + (G2 ? 1u : 0u)));
  ptx::bulk_load(sy, y.p + n * y.sN + sg.h * y.sH + (long long)sg.v0 * 8, bytes, bar);
  if (G1) ptx::bulk_load(sg1, g.g1.p + n * g.g1.sN + sg.h * g.g1.sH + (long long)sg.v0 * 8, bytes, bar);
  if (G2)
    ptx::bulk_load(sg2, g.g2.p + n * g.g2.sN + (sg.h + g.fold) * g.g2.sH + (long long)g.fold * y.C + (long long)sg.v0 * 8,
                   bytes, bar);
}

template <int ACT, bool G1, bool G2, bool DA>
__global__ void __launch_bounds__(256, 3)
in_bwd_reduce_bulk_kernel(DevTensor y, const float2* __restrict__ stats, DevGrad g, DevTensor da,
                          float* __restrict__ bstats, int vpr, int nslice, int segv, int spb) {
  extern __shared__ __align__(128) uint8_t bulk_smem[];
  __shared__ uint64_t bar;
  uint4* sy = reinterpret_cast<uint4*>(bulk_smem);
  uint4* sg1 = sy + segv;
  uint4* sg2 = sg1 + (G1 ? segv : 0);
  const int t = threadIdx.x, n = blockIdx.y;
  const int C = y.C, H = y.H, W = y.W, p = g.fold;
  const int lanes = C >> 3, lsh = 31 - __clz(lanes), rows = 256 >> lsh, prow = t >> lsh, cl = t & (lanes - 1);
  const int c0 = cl * 8;
  const int s0 = blockIdx.x * spb, s1e = min(H * nslice, s0 + spb);  // this block's segments
  if (t == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  ptx::pdl_wait();
  if (kPwTrigger) ptx::pdl_launch_dependents();
  if (t == 0 && s0 < s1e) bwd_bulk_issue<G1, G2>(y, g, n, seg_of(s0, nslice, vpr, segv), sy, sg1, sg2, &bar);
  float a[8], b[8], s1[8], s2[8];
  load_norm8(stats, (long long)n * C + c0, 1.f / (float)(H * W), a, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
  const bf16* g2img = G2 ? g.g2.p + n * g.g2.sN : nullptr;
  uint32_t phase = 0;
  for (int s = s0; s < s1e; ++s) {
    const Seg sg = seg_of(s, nslice, vpr, segv);
    const int h = sg.h;
    bf16* darow = DA ? da.p + n * da.sN + h * da.sH : nullptr;
    const bool hb = G2 && ((h >= 1 && h <= p) || (h >= H - 1 - p && h <= H - 2));
    const int hm = (h >= 1 && h <= p) ? p - h : 2 * (H - 1) - h + p;  // padded-domain row of the mirror
    ptx::mbar_wait(&bar, phase);
    phase ^= 1;
#pragma unroll 2
    for (int vi = t; vi < sg.nv; vi += 256) {
      const int gv = sg.v0 + vi;
      float v[8], gr[8];
      rows_grad<G1, G2>(G1 ? sg1[vi] : sy[vi], G2 ? sg2[vi] : sy[vi], g2img, g.g2.sH, C, c0, h, gv >> lsh, H, W, p, hb, hm, gr);
      if (DA) {
#pragma unroll
        for (int i = 0; i < 8; ++i) gr[i] = round_bf16(gr[i]);
        store8(darow + gv * 8, gr);
      }
      unpack8(sy[vi], v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = fmaf(v[i], a[i], b[i]);
        const float dz = gr[i] * act_grad_t<ACT>(xh);
        s1[i] += dz;
        s2[i] = fmaf(dz, xh, s2[i]);
      }
    }
    if (s + 1 < s1e) {
      __syncthreads();  // every thread is done reading the segment: its buffers may be refilled
      if (t == 0) bwd_bulk_issue<G1, G2>(y, g, n, seg_of(s + 1, nslice, vpr, segv), sy, sg1, sg2, &bar);
    }
  }
  // the block's pixel slots that share a channel vector -> shared memory -> one vector atomic per channel pair
  __syncthreads();
  float* red = reinterpret_cast<float*>(bulk_smem);  // 256 * 16 floats: the segment buffers are idle now
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[t * 16 + i] = s1[i];
    red[t * 16 + 8 + i] = s2[i];
  }
  __syncthreads();
  if (prow == 0) {
    for (int r = 1; r < rows; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s1[i] += red[(r * lanes + cl) * 16 + i];
        s2[i] += red[(r * lanes + cl) * 16 + 8 + i];
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i += 2)
      red_add_v4(bstats + ((long long)n * C + c0 + i) * 2, s1[i], s2[i], s1[i + 1], s2[i + 1]);
  }
}

template <int ACT, bool G1, bool G2>
__global__ void __launch_bounds__(256, 3)
in_bwd_apply_bulk_kernel(DevTensor y, const float2* __restrict__ stats, const float2* __restrict__ bstats, DevGrad g,
                         DevTensor dy, int vpr, int nslice, int segv) {
  extern __shared__ __align__(128) uint8_t bulk_smem[];
  __shared__ uint64_t bar;
  uint4* sy = reinterpret_cast<uint4*>(bulk_smem);
  uint4* sg1 = sy + segv;
  uint4* sg2 = sg1 + (G1 ? segv : 0);
  const int t = threadIdx.x, n = blockIdx.y;
  const int C = y.C, H = y.H, W = y.W, p = g.fold;
  const int lanes = C >> 3, lsh = 31 - __clz(lanes);
  const int c0 = (t & (lanes - 1)) * 8;
  const Seg sg = seg_of(blockIdx.x, nslice, vpr, segv);
  const int h = sg.h;
  if (t == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  ptx::pdl_wait();
  if (kPwTrigger) ptx::pdl_launch_dependents();
  if (t == 0) bwd_bulk_issue<G1, G2>(y, g, n, sg, sy, sg1, sg2, &bar);
  // dy = rstd * (dz - mean(dz) - xhat * mean(dz * xhat)) = a * dz - c - d * xhat
  float a[8], b[8], c[8], d[8];
  const float inv = 1.f / (float)(H * W);
  load_norm8(stats, (long long)n * C + c0, inv, a, b);
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    const float4 bs = __ldg(reinterpret_cast<const float4*>(bstats + (long long)n * C + c0 + i));
    c[i] = a[i] * bs.x * inv;
    d[i] = a[i] * bs.y * inv;
    c[i + 1] = a[i + 1] * bs.z * inv;
    d[i + 1] = a[i + 1] * bs.w * inv;
  }
  const bf16* g2img = G2 ? g.g2.p + n * g.g2.sN : nullptr;
  bf16* dyrow = dy.p + n * dy.sN + h * dy.sH;
  const bool hb = G2 && ((h >= 1 && h <= p) || (h >= H - 1 - p && h <= H - 2));
  const int hm = (h >= 1 && h <= p) ? p - h : 2 * (H - 1) - h + p;
  ptx::mbar_wait(&bar, 0);
#pragma unroll 2
  for (int vi = t; vi < sg.nv; vi += 256) {
    const int gv = sg.v0 + vi;
    float v[8], gr[8];
    rows_grad<G1, G2>(G1 ? sg1[vi] : sy[vi], G2 ? sg2[vi] : sy[vi], g2img, g.g2.sH, C, c0, h, gv >> lsh, H, W, p, hb, hm, gr);
    unpack8(sy[vi], v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = fmaf(v[i], a[i], b[i]);
      const float dz = gr[i] * act_grad_t<ACT>(xh);
      v[i] = fmaf(a[i], dz, -c[i]) - d[i] * xh;
    }
    store8(dyrow + gv * 8, v);
  }
}

// ------------------------------------------------------------------------------------------ head / losses
__device__ __forceinline__ float block_sum(float v) {
  __shared__ float sh[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x + 31) / 32 ? sh[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid in thread 0
}

// one thread per pixel; tensors have 16 stored channels, the first C (<= 8) are real
__global__ void tanh_bwd_kernel(DevTensor out, DevTensor target, float l1_scale, DevGrad g, int C, DevTensor dpre,
                                float* __restrict__ loss_slot) {
  ptx::pdl_wait();  // launched with programmatic stream serialization (see launch_pdl)
  if (kPwTrigger) ptx::pdl_launch_dependents();
  const long long total = (long long)out.N * out.H * out.W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  float lsum = 0.f;
  if (idx < total) {
    const int w = idx % out.W;
    const int h = (idx / out.W) % out.H;
    const int n = idx / ((long long)out.W * out.H);
    float o[8], gr[8], d[8];
    load8(out.p + n * out.sN + h * out.sH + w * out.sW, o);
    load_grad8(g, n, h, w, 0, out.H, out.W, gr);
    if (target.p != nullptr) {
      float t[8];
      load8(target.p + n * target.sN + h * target.sH + w * target.sW, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < C) {
          const float diff = o[i] - t[i];
          lsum += fabsf(diff);
          gr[i] += l1_scale * (float)((diff > 0.f) - (diff < 0.f));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = i < C ? gr[i] * (1.f - o[i] * o[i]) : 0.f;
    bf16* dp = dpre.p + n * dpre.sN + h * dpre.sH + w * dpre.sW;
    store8(dp, d);
    *reinterpret_cast<uint4*>(dp + 8) = make_uint4(0, 0, 0, 0);
  }
  if (loss_slot != nullptr && target.p != nullptr) {
    const float t = block_sum(lsum);
    if (threadIdx.x == 0) atomicAdd(loss_slot, t * l1_scale);
  }
}

__global__ void l1_loss_kernel(DevTensor a, DevTensor b, int C, float scale, float* __restrict__ loss_slot) {
  const long long total = (long long)a.N * a.H * a.W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  float lsum = 0.f;
  if (idx < total) {
    const int w = idx % a.W;
    const int h = (idx / a.W) % a.H;
    const int n = idx / ((long long)a.W * a.H);
    float x[8], y[8];
    load8(a.p + n * a.sN + h * a.sH + w * a.sW, x);
    load8(b.p + n * b.sN + h * b.sH + w * b.sW, y);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < C) lsum += fabsf(x[i] - y[i]);
  }
  const float t = block_sum(lsum);
  if (threadIdx.x == 0) atomicAdd(loss_slot, t * scale);
}

__global__ void leaky_bwd_kernel(DevTensor a, DevTensor g, DevTensor dpre) {
  const int C8 = a.C / 8;
  const long long total = (long long)a.N * a.H * a.W * C8;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c0 = (idx % C8) * 8;
  long long r = idx / C8;
  const int w = r % a.W;
  r /= a.W;
  const int h = r % a.H;
  const int n = r / a.H;
  float av[8], gv[8];
  load8(a.p + n * a.sN + h * a.sH + w * a.sW + c0, av);
  load8(g.p + n * g.sN + h * g.sH + w * g.sW + c0, gv);
#pragma unroll
  for (int i = 0; i < 8; ++i) gv[i] *= av[i] > 0.f ? 1.f : 0.2f;
  store8(dpre.p + n * dpre.sN + h * dpre.sH + w * dpre.sW + c0, gv);
}

__global__ void mse_loss_kernel(DevTensor logits, float target, float wgt, float* __restrict__ loss_slot,
                                DevTensor dl) {
  const long long total = (long long)logits.N * logits.H * logits.W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  float lsum = 0.f;
  const float inv = 1.f / (float)total;
  if (idx < total) {
    const int w = idx % logits.W;
    const int h = (idx / logits.W) % logits.H;
    const int n = idx / ((long long)logits.W * logits.H);
    const float p = __bfloat162float(logits.p[n * logits.sN + h * logits.sH + w * logits.sW]);
    const float d = p - target;
    lsum = d * d;
    if (dl.p != nullptr) {
      float v[8] = {2.f * wgt * d * inv, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      bf16* o = dl.p + n * dl.sN + h * dl.sH + w * dl.sW;
      store8(o, v);
      *reinterpret_cast<uint4*>(o + 8) = make_uint4(0, 0, 0, 0);
    }
  }
  const float t = block_sum(lsum);
  if (threadIdx.x == 0 && loss_slot != nullptr) atomicAdd(loss_slot, t * wgt * inv);
}

// ------------------------------------------------------------------------------------------ weights
// One block = one (tap, 32 x 32 tile of (cout, cin)) of one layer: the fp32 master [Cout][T][Cin] is read with cin
// fastest, Wf[co][t][ci] is written in the same orientation and Wt[ci][t][co] after a transpose through shared
// memory, so all three streams are coalesced (the scattered 2-byte writes of a direct transpose made this the
// slowest kernel of the optimiser phase: 176 us for the generator pair).
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ master, const PackEntry* __restrict__ entries, bf16* __restrict__ arena) {
  const PackEntry e = entries[blockIdx.y];
  const int tiles_ci = (e.Cin + 31) / 32, tiles_co = (e.Cout + 31) / 32;
  const int tiles = e.T * tiles_co * tiles_ci;
  __shared__ bf16 tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int tl = blockIdx.x; tl < tiles; tl += gridDim.x) {
    const int tci = tl % tiles_ci;
    const int tco = (tl / tiles_ci) % tiles_co;
    const int t = tl / (tiles_ci * tiles_co);
    const int ci = tci * 32 + tx;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int co = tco * 32 + r;
      bf16 v = __float2bfloat16_rn(0.f);
      if (co < e.Cout && ci < e.Cin) {
        v = __float2bfloat16_rn(master[e.src_off + ((long long)co * e.T + t) * e.Cin + ci]);
        arena[e.wf_off + ((long long)co * e.T + t) * e.CinS + ci] = v;
        if (e.wx_off >= 0) arena[e.wx_off + (long long)co * e.wx_pitch + t * 4 + ci] = v;
      }
      tile[r][tx] = v;
    }
    __syncthreads();
    const int co = tco * 32 + tx;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int ci2 = tci * 32 + r;
      if (co < e.Cout && ci2 < e.Cin) arena[e.wt_off + ((long long)ci2 * e.T + t) * e.CoutS + co] = tile[tx][r];
    }
    __syncthreads();
  }
}

// hyper[2] holds the learning rate in device memory (cgb_set_lr), so an LR schedule needs no graph re-capture
__global__ void adam_prep_kernel(int* step, float* hyper, float beta1, float beta2) {
  const int t = *step + 1;
  *step = t;
  const float lr = hyper[2];
  hyper[0] = (float)((double)lr / (1.0 - pow((double)beta1, (double)t)));
  hyper[1] = (float)(1.0 / sqrt(1.0 - pow((double)beta2, (double)t)));
}

// Image history pool (canonical ImagePool.query with the decisions made on the host): per image n, in batch
// order, d_in[n] = ret >= 0 ? pool[ret] : fake[n]; then pool[store] = fake[n] when store >= 0.  A thread owns one
// 8-channel vector of a pixel and walks the batch sequentially, so two images that pick the same slot behave as in
// the sequential stand-in.  dec = [N][2] = (store, ret).
__global__ void pool_exchange_kernel(DevTensor fake, DevTensor pool, const int* __restrict__ dec, DevTensor d_in) {
  const int C8 = fake.C / 8;
  const long long total = (long long)fake.H * fake.W * C8;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c0 = (int)(idx % C8) * 8;
  const long long p = idx / C8;
  const int w = (int)(p % fake.W), h = (int)(p / fake.W);
  for (int n = 0; n < fake.N; ++n) {
    const int store = dec[2 * n], ret = dec[2 * n + 1];
    const uint4 v = *reinterpret_cast<const uint4*>(fake.p + n * fake.sN + h * fake.sH + w * fake.sW + c0);
    uint4 r = v;
    if (ret >= 0 && ret < pool.N) r = *reinterpret_cast<const uint4*>(pool.p + ret * pool.sN + h * pool.sH + w * pool.sW + c0);
    *reinterpret_cast<uint4*>(d_in.p + n * d_in.sN + h * d_in.sH + w * d_in.sW + c0) = r;
    if (store >= 0 && store < pool.N) *reinterpret_cast<uint4*>(pool.p + store * pool.sN + h * pool.sH + w * pool.sW + c0) = v;
  }
}

// bf16 NHWC image (16 stored channels, C real) -> uint8 interleaved [N][H][W][C]: u8 = clamp(rint((x + 1) * 127.5))
// (the inverse of u8hwc_to_nchw; the stand-in's to_uint8)
__global__ void nhwc_to_u8hwc_kernel(DevTensor src, int C, unsigned char* __restrict__ dst) {
  const long long total = (long long)src.N * src.H * src.W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int w = idx % src.W;
  const int h = (idx / src.W) % src.H;
  const int n = idx / ((long long)src.W * src.H);
  const bf16* p = src.p + n * src.sN + h * src.sH + w * src.sW;
  for (int c = 0; c < C; ++c) {
    const float v = __fmul_rn(__fadd_rn(__bfloat162float(p[c]), 1.f), 127.5f);
    dst[idx * C + c] = (unsigned char)min(255, max(0, __float2int_rn(v)));
  }
}

__global__ void set_float_kernel(float* dst, float value) { *dst = value; }

// uint8 interleaved RGB [N][H][W][3] -> fp32 planar [N][3][H][W] in [-1, 1]: x = u8 / 127.5 - 1
// (the canonical ToTensor + Normalize(0.5, 0.5) of the CycleGAN input pipeline)
__global__ void u8hwc_to_nchw_kernel(const unsigned char* __restrict__ src, int N, int H, int W, float* __restrict__ dst) {
  const long long total = (long long)N * H * W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long hw = (long long)H * W;
  const long long n = idx / hw, p = idx - n * hw;
  const unsigned char* s = src + idx * 3;
#pragma unroll
  // two separately rounded operations (no FMA contraction): bit-identical to the stand-in's x * (1 / 127.5) - 1
  for (int c = 0; c < 3; ++c) dst[(n * 3 + c) * hw + p] = __fsub_rn(__fmul_rn((float)s[c], (float)(1.0 / 127.5)), 1.f);
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float beta1, float beta2, float eps,
                            const float* __restrict__ hyper, float grad_scale) {
  const float step_size = hyper[0], inv_sqrt_bc2 = hyper[1];
  const long long i4 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4;
  if (i4 + 4 <= n) {
    float4 pp = *reinterpret_cast<float4*>(p + i4);
    const float4 gg = *reinterpret_cast<const float4*>(g + i4);
    float4 mm = *reinterpret_cast<float4*>(m + i4);
    float4 vv = *reinterpret_cast<float4*>(v + i4);
    float* pa = &pp.x;
    const float* ga = &gg.x;
    float* ma = &mm.x;
    float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = ga[k] * grad_scale;
      ma[k] = beta1 * ma[k] + (1.f - beta1) * gr;
      va[k] = beta2 * va[k] + (1.f - beta2) * gr * gr;
      pa[k] -= step_size * ma[k] / (sqrtf(va[k]) * inv_sqrt_bc2 + eps);
    }
    *reinterpret_cast<float4*>(p + i4) = pp;
    *reinterpret_cast<float4*>(m + i4) = mm;
    *reinterpret_cast<float4*>(v + i4) = vv;
  } else {
    for (long long i = i4; i < n; ++i) {
      const float gr = g[i] * grad_scale;
      m[i] = beta1 * m[i] + (1.f - beta1) * gr;
      v[i] = beta2 * v[i] + (1.f - beta2) * gr * gr;
      p[i] -= step_size * m[i] / (sqrtf(v[i]) * inv_sqrt_bc2 + eps);
    }
  }
}

// ------------------------------------------------------------------------------------------ im2col (<= 4 channels)
// one thread per (pixel, tap): an 8-byte load (channels 0..3) and an 8-byte store; the taps of a pixel are
// consecutive threads, so stores are contiguous
__global__ void im2col4_kernel(DevTensor src, int k, int stride, int sgn, int off, int use_halo, DevTensor dst) {
  ptx::pdl_wait();  // launched with programmatic stream serialization (see launch_pdl)
  if (kPwTrigger) ptx::pdl_launch_dependents();
  const int T = k * k;
  const long long total = (long long)dst.N * dst.H * dst.W * T;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int t = idx % T;
  long long r = idx / T;
  const int w = r % dst.W;
  r /= dst.W;
  const int h = r % dst.H;
  const int n = r / dst.H;
  const int lo = use_halo ? -src.halo : 0;
  const int hiH = src.H + (use_halo ? src.halo : 0), hiW = src.W + (use_halo ? src.halo : 0);
  const int sh = h * stride + sgn * (t / k) + off, sw = w * stride + sgn * (t % k) + off;
  uint2 v = make_uint2(0u, 0u);
  if (sh >= lo && sh < hiH && sw >= lo && sw < hiW)
    v = *reinterpret_cast<const uint2*>(src.p + n * src.sN + sh * src.sH + sw * src.sW);
  *reinterpret_cast<uint2*>(dst.p + n * dst.sN + h * dst.sH + w * dst.sW + t * 4) = v;
}

// one thread per (pixel of dst, half, horizontal tap s < 8): an 8-byte load (channels 0..3) and an 8-byte store
__global__ void expand_rows4_kernel(DevTensor src, int k, int sgn, int off, int row_off, int lo, DevTensor dst) {
  ptx::pdl_wait();  // launched with programmatic stream serialization (see launch_pdl)
  const long long total = (long long)dst.N * dst.H * dst.W * 16;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int s = idx & 7, half = (idx >> 3) & 1;
  long long r = idx >> 4;
  const int w = r % dst.W;
  r /= dst.W;
  const int hh = r % dst.H;
  const int n = r / dst.H;
  const int sh = hh + half + row_off, sw = w + sgn * s + off;
  uint2 v = make_uint2(0u, 0u);
  if (s < k && sh >= lo && sh < src.H - lo && sw >= lo && sw < src.W - lo)
    v = *reinterpret_cast<const uint2*>(src.p + n * src.sN + sh * src.sH + sw * src.sW);
  *reinterpret_cast<uint2*>(dst.p + n * dst.sN + hh * dst.sH + w * dst.sW + half * 32 + s * 4) = v;
}

}  // namespace

// ================================================================================================ host API
void nchw_to_nhwc(const float* src, int C, const TensorDesc& dst, cudaStream_t st) {
  if (dst.esz == 4) return f32::nchw_to_nhwc(src, C, dst, st);
  CGB_CHECK(dst.C % 8 == 0 && C <= dst.C, "nchw_to_nhwc: bad channel counts");
  const long long total = (long long)dst.N * (dst.H + 2 * dst.halo) * (dst.W + 2 * dst.halo);
  nchw_to_nhwc_kernel<<<blocks_for(total, 256), 256, 0, st>>>(src, C, dev(dst));
  CGB_CUDA(cudaGetLastError());
}

void nhwc_to_nchw(const TensorDesc& src, int C, float* dst, cudaStream_t st) {
  if (src.esz == 4) return f32::nhwc_to_nchw(src, C, dst, st);
  const long long total = (long long)src.N * src.H * src.W;
  nhwc_to_nchw_kernel<<<blocks_for(total, 256), 256, 0, st>>>(dev(src), C, dst);
  CGB_CUDA(cudaGetLastError());
}

void fill_reflect_halo(const TensorDesc& t, cudaStream_t st) {
  if (t.halo == 0) return;
  const DevTensor d = dev_bytes(t);
  const long long total = (long long)t.N * (t.H + 2 * t.halo) * (t.W + 2 * t.halo) * (d.C / 8);
  launch_pdl(fill_halo_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, st, d);
}

void in_stats(const TensorDesc& y, float2* stats, cudaStream_t st) {
  CGB_CHECK(y.C % 8 == 0, "in_stats: channels must be a multiple of 8");
  const int HW = y.H * y.W;
  const int ppb = pick_pix_per_block(HW);
  dim3 grid((HW + ppb - 1) / ppb, y.N);
  launch_pdl(colsum_kernel<0>, grid, dim3(256), 0, st, dev(y), reinterpret_cast<float*>(stats), ppb, y.C);
}

void bias_grad(const TensorDesc& dy, int C, float* gbias, cudaStream_t st) {
  if (dy.esz == 4) return f32::bias_grad(dy, C, gbias, st);  // validation mode: STORES the sum (one buffer per pass)
  const int HW = dy.H * dy.W;
  // every block ends with one scalar atomic per channel on the SAME few addresses (all images share the bias):
  // ~148 blocks over the whole batch (512 per image made the 3-channel head bias gradient 73 us at batch 8)
  const int per_image = std::max(1, 148 / dy.N);  // ~one block per SM over the batch
  const int ppb = std::max(pick_pix_per_block(HW), (HW / per_image + 31) / 32 * 32);
  dim3 grid((HW + ppb - 1) / ppb, dy.N);
  launch_pdl(colsum_kernel<1>, grid, dim3(256), 0, st, dev(dy), gbias, ppb, C);
}

static bool async_enabled() {
  static const bool on = !(std::getenv("CGB_PW_ASYNC") && std::atoi(std::getenv("CGB_PW_ASYNC")) == 0);
  return on;
}
// cp.async variants: ~2 blocks per SM over the whole batch, used when that leaves every thread >= 3 iterations
static bool pick_async(long long pixels, int rows, int images, int /*unused*/, int* ppb_out) {
  if (!async_enabled()) return false;
  const int unit = rows * 2;
  const long long per_image = std::max(1LL, 2LL * 148 / images);
  long long ppb = (pixels + per_image - 1) / per_image;
  ppb = (ppb + unit - 1) / unit * unit;
  if (ppb / unit < 3) return false;
  *ppb_out = (int)ppb;
  return true;
}

// Streaming kernels: blocks of `ppb` pixels sized for ~8 CTAs per SM over the whole batch (but at least one
// PPT-batch per pixel row of the block, so small maps still spread over the machine).
static int stream_ppb(long long pixels, int rows, int ppt, int images) {
  const int unit = rows * ppt;
  const long long want_blocks = std::max(1LL, 8LL * 148 / images);
  long long ppb = (pixels + want_blocks - 1) / want_blocks;
  ppb = (ppb + unit - 1) / unit * unit;
  return (int)std::max<long long>(ppb, unit);
}

// Row-streaming kernels (CGB_PW_ROWS=0 falls back to the flat-range kernels): channel counts that divide 2048
// (8 .. 2048, powers of two) with a halo / fold narrower than the map.
static bool rows_enabled() {
  static const bool on = !(std::getenv("CGB_PW_ROWS") && std::atoi(std::getenv("CGB_PW_ROWS")) == 0);
  return on;
}
static bool rows_ok(const TensorDesc& y, int border) {
  return rows_enabled() && y.esz == 2 && y.C >= 8 && y.C <= 2048 && (y.C & (y.C - 1)) == 0 && 2 * border + 2 <= y.H &&
         2 * border + 2 <= y.W;
}
// rows per block (and, when the whole batch has fewer rows than the machine holds blocks, slices per row) for a
// launch that should fill 148 SMs x 3 resident blocks about `waves` times; max_bpi caps the blocks per image
static void rows_grid(int H, int N, int vpr, int max_bpi, int* rpb, int* jsplit) {
  const long long R = (long long)H * N;
  const int want = 148 * 3 * 2;
  int r = (int)std::max<long long>(1, R / want);
  if (max_bpi > 0) r = std::max(r, (H + max_bpi - 1) / max_bpi);
  int js = 1;
  if (max_bpi == 0 && r == 1) {
    const int J = (vpr + 255) / 256;
    while (js * 2 <= J && R * js * 2 <= 148 * 3) js *= 2;
  }
  *rpb = r;
  *jsplit = js;
}
#define CGB_ACT_SWITCH(act, CALL)                              \
  switch (act) {                                               \
    case kActNone: { constexpr int ACT = kActNone; CALL; } break;   \
    case kActLeaky: { constexpr int ACT = kActLeaky; CALL; } break; \
    case kActRelu: { constexpr int ACT = kActRelu; CALL; } break;   \
    default: CGB_CHECK(false, "InstanceNorm kernels: unsupported activation");  \
  }

// Bulk-copy kernels (CGB_PW_BULK=0 falls back to the register row-streaming kernels): same shape conditions.
static bool bulk_enabled() {
  static const bool on = !(std::getenv("CGB_PW_BULK") && std::atoi(std::getenv("CGB_PW_BULK")) == 0);
  return on;
}
// Segments are whole rows, or kBulkVec-vector slices of longer rows.  Small launches (fewer than two blocks per SM:
// batch 1 on the 64 x 64 maps) stay on the register kernels: they are latency-bound and the barrier set-up plus the
// single-thread issue cost more than they save (measured at batch 1: 4.35 ms/step with the register kernels, 4.41 with
// bulk copies everywhere; at batch 8: 24.1 vs 23.6).
static bool bulk_plan(int H, int N, int vpr, int* segv, int* nslice) {
  const int sv = std::min(kBulkVec, (vpr + 255) / 256 * 256);
  *segv = sv;
  *nslice = (vpr + sv - 1) / sv;
  static const int min_blocks = std::getenv("CGB_PW_BULK_MIN") ? std::atoi(std::getenv("CGB_PW_BULK_MIN")) : 2 * 148;
  return bulk_enabled() && (long long)H * N * *nslice >= min_blocks;
}
template <typename... KArgs, typename... Args>
static void launch_bulk(void (*kern)(KArgs...), dim3 grid, size_t smem, cudaStream_t st, Args&&... args) {
  static std::set<const void*> configured;  // kernels whose dynamic shared-memory limit has been raised
  const void* key = reinterpret_cast<const void*>(kern);
  if (!configured.count(key)) {
    CGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * kBulkVec * 16));
    configured.insert(key);
  }
  launch_pdl(kern, grid, dim3(256), smem, st, std::forward<Args>(args)...);
}

void in_apply(const TensorDesc& y, const float2* stats, int act, const TensorDesc* residual, const TensorDesc& out,
              cudaStream_t st) {
  CGB_CHECK(y.C == out.C && y.H == out.H && y.W == out.W && y.N == out.N, "in_apply: shape mismatch");
  int segv = 0, nslice = 0;
  if (rows_ok(out, out.halo) && (act == kActNone || act == kActLeaky || act == kActRelu) &&
      bulk_plan(out.H, out.N, out.W * out.C / 8, &segv, &nslice)) {
    const int vpr = out.W * out.C / 8;
    dim3 grid(out.H * nslice, out.N);
    const DevTensor r = residual ? dev(*residual) : dev_null();
    if (residual) {
      CGB_ACT_SWITCH(act, (launch_bulk(in_apply_bulk_kernel<ACT, true>, grid, (size_t)segv * 32, st, dev(y), stats, r, dev(out),
                                       vpr, nslice, segv)));
    } else {
      CGB_ACT_SWITCH(act, (launch_bulk(in_apply_bulk_kernel<ACT, false>, grid, (size_t)segv * 16, st, dev(y), stats, r, dev(out),
                                       vpr, nslice, segv)));
    }
    return;
  }
  if (rows_ok(out, out.halo) && (act == kActNone || act == kActLeaky || act == kActRelu)) {
    const int vpr = out.W * out.C / 8;
    int rpb, jsplit;
    rows_grid(out.H, out.N, vpr, 0, &rpb, &jsplit);
    dim3 grid((out.H + rpb - 1) / rpb * jsplit, out.N);
    const DevTensor r = residual ? dev(*residual) : dev_null();
    if (residual) {
      CGB_ACT_SWITCH(act, (launch_pdl(in_apply_rows_kernel<ACT, true, 4>, grid, dim3(256), 0, st, dev(y), stats, r, dev(out),
                                      vpr, rpb, jsplit)));
    } else {
      CGB_ACT_SWITCH(act, (launch_pdl(in_apply_rows_kernel<ACT, false, 4>, grid, dim3(256), 0, st, dev(y), stats, r, dev(out),
                                      vpr, rpb, jsplit)));
    }
    return;
  }
  const int total = (out.H + 2 * out.halo) * (out.W + 2 * out.halo);
  const int lanes = std::min(256, out.C / 8), rows = 256 / lanes;
  int appb;
  if (pick_async(total, rows, out.N, 0, &appb)) {
    static bool configured = false;
    constexpr int kSmem = kAsyncStages * 2 * 2 * 256 * 16;
    if (!configured) {
      CGB_CUDA(cudaFuncSetAttribute(in_apply_async_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
      configured = true;
    }
    dim3 grid((total + appb - 1) / appb, out.N);
    launch_pdl(in_apply_async_kernel<2>, grid, dim3(256), (size_t)kSmem, st, dev(y), stats, act,
               residual ? dev(*residual) : dev_null(), dev(out), appb);
    return;
  }
  const int ppb = stream_ppb(total, rows, 2, out.N);
  dim3 grid((total + ppb - 1) / ppb, out.N);
  launch_pdl(in_apply_kernel<2>, grid, dim3(256), 0, st, dev(y), stats, act, residual ? dev(*residual) : dev_null(), dev(out), ppb);
}

static void check_grad(const TensorDesc& y, const GradSrc& g) {
  CGB_CHECK(g.g1 || g.g2, "gradient source is empty");
  if (g.g1) CGB_CHECK(g.g1->H == y.H && g.g1->W == y.W && g.g1->C == y.C, "g1 shape mismatch");
  if (g.g2)
    CGB_CHECK(g.g2->H == y.H + 2 * g.fold && g.g2->W == y.W + 2 * g.fold && g.g2->C == y.C && g.g2->halo == 0,
              "g2 (padded-domain gradient) shape mismatch");
}

static bool in_fused_enabled() {
  static const bool on = !(std::getenv("CGB_IN_FUSED") && std::atoi(std::getenv("CGB_IN_FUSED")) == 0);
  return on;
}

// channel-group width (32, else 16) and register items per thread (1, 2 or 4); 0: the map has more than 4096 pixels
// per image (or an odd channel count) and keeps the two-kernel path
static int in_fused_plan(const TensorDesc& y, int* pix_per_cta, int* items) {
  const int HW = y.H * y.W;
  for (int cg : {32, 16}) {
    if (y.C % cg != 0) continue;
    const int ppc = (HW + kInCluster - 1) / kInCluster;
    int it = (ppc * (cg / 8) + 255) / 256;
    if (it > 4) continue;
    it = it <= 1 ? 1 : it <= 2 ? 2 : 4;
    // Measured on B200 (profiles/r02_c_ops_b*.txt): the cluster kernel halves the InstanceNorm-backward time of a
    // residual-stream tensor at batch 1 (16.5 -> 8-10 us) but its register-resident elements cap the occupancy, so
    // on grids of several waves (batch >= 4) the two streaming kernels are faster (33 vs 47 us at batch 8).
    static const int max_ctas = std::getenv("CGB_IN_FUSED_MAX_CTAS") ? std::atoi(std::getenv("CGB_IN_FUSED_MAX_CTAS")) : 2 * 148;
    if ((long long)kInCluster * (y.C / cg) * y.N > max_ctas) return 0;
    *pix_per_cta = ppc;
    *items = it;
    return cg;
  }
  return 0;
}

bool in_bwd_fused_supported(const TensorDesc& y) {
  int ppc = 0, items = 0;
  return in_fused_enabled() && y.esz == 2 && y.C % 8 == 0 && in_fused_plan(y, &ppc, &items) != 0;
}

template <int CG, int ITEMS>
static void launch_in_bwd_fused(const TensorDesc& y, const float2* stats, const GradSrc& g, int act, const TensorDesc* da_out,
                                const TensorDesc& dy, int ppc, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(kInCluster * (y.C / CG)), (unsigned)y.N);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kInCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  CGB_CUDA(cudaLaunchKernelEx(&cfg, in_bwd_fused_kernel<CG, ITEMS>, dev(y), stats, dev(g), act,
                              da_out ? dev(*da_out) : dev_null(), dev(dy), ppc));
}

bool in_bwd_fused(const TensorDesc& y, const float2* stats, const GradSrc& g, int act, const TensorDesc* da_out,
                  const TensorDesc& dy, cudaStream_t st) {
  if (!in_fused_enabled() || y.esz != 2) return false;
  int ppc = 0, items = 0;
  const int cg = in_fused_plan(y, &ppc, &items);
  if (cg == 0) return false;
  check_grad(y, g);
  switch (cg * 10 + items) {
    case 321: launch_in_bwd_fused<32, 1>(y, stats, g, act, da_out, dy, ppc, st); break;
    case 322: launch_in_bwd_fused<32, 2>(y, stats, g, act, da_out, dy, ppc, st); break;
    case 324: launch_in_bwd_fused<32, 4>(y, stats, g, act, da_out, dy, ppc, st); break;
    case 161: launch_in_bwd_fused<16, 1>(y, stats, g, act, da_out, dy, ppc, st); break;
    case 162: launch_in_bwd_fused<16, 2>(y, stats, g, act, da_out, dy, ppc, st); break;
    case 164: launch_in_bwd_fused<16, 4>(y, stats, g, act, da_out, dy, ppc, st); break;
    default: return false;
  }
  return true;
}

void in_bwd_reduce(const TensorDesc& y, const float2* stats, const GradSrc& g, int act, const TensorDesc* da_out,
                   float2* bstats, cudaStream_t st) {
  check_grad(y, g);
  int segv = 0, nslice = 0;
  if (rows_ok(y, g.g2 ? g.fold : 0) && (act == kActNone || act == kActLeaky || act == kActRelu) &&
      bulk_plan(y.H, y.N, y.W * y.C / 8, &segv, &nslice)) {
    const int vpr = y.W * y.C / 8;
    const int segs = y.H * nslice;
    const int spb = (segs + 63) / 64;  // <= 64 blocks per image end in atomics on the same statistics
    dim3 grid((segs + spb - 1) / spb, y.N);
    const DevTensor dat = da_out ? dev(*da_out) : dev_null();
    float* bs = reinterpret_cast<float*>(bstats);
    // (the final cross-slot reduction reuses the buffers: 256 x 16 floats)
    const size_t smem = std::max<size_t>(256 * 16 * sizeof(float), (size_t)segv * 16 * (1 + (g.g1 ? 1 : 0) + (g.g2 ? 1 : 0)));
#define CGB_RED_LAUNCH(G1, G2, DA) \
  CGB_ACT_SWITCH(act, (launch_bulk(in_bwd_reduce_bulk_kernel<ACT, G1, G2, DA>, grid, smem, st, dev(y), stats, dev(g), dat, bs, vpr, nslice, segv, spb)))
    const int key = (g.g1 ? 4 : 0) | (g.g2 ? 2 : 0) | (da_out ? 1 : 0);
    switch (key) {
      case 4: CGB_RED_LAUNCH(true, false, false); break;
      case 5: CGB_RED_LAUNCH(true, false, true); break;
      case 2: CGB_RED_LAUNCH(false, true, false); break;
      case 3: CGB_RED_LAUNCH(false, true, true); break;
      case 6: CGB_RED_LAUNCH(true, true, false); break;
      case 7: CGB_RED_LAUNCH(true, true, true); break;
      default: CGB_CHECK(false, "in_bwd_reduce: no gradient source");
    }
#undef CGB_RED_LAUNCH
    return;
  }
  if (rows_ok(y, g.g2 ? g.fold : 0) && (act == kActNone || act == kActLeaky || act == kActRelu)) {
    const int vpr = y.W * y.C / 8;
    int rpb, jsplit;
    rows_grid(y.H, y.N, vpr, 64, &rpb, &jsplit);  // <= 64 blocks per image end in atomics on the same statistics
    dim3 grid((y.H + rpb - 1) / rpb, y.N);
    const DevTensor dat = da_out ? dev(*da_out) : dev_null();
    float* bs = reinterpret_cast<float*>(bstats);
#define CGB_RED_LAUNCH(G1, G2, DA) \
  CGB_ACT_SWITCH(act, (launch_pdl(in_bwd_reduce_rows_kernel<ACT, G1, G2, DA, 2>, grid, dim3(256), 0, st, dev(y), stats, dev(g), dat, bs, vpr, rpb)))
    const int key = (g.g1 ? 4 : 0) | (g.g2 ? 2 : 0) | (da_out ? 1 : 0);
    switch (key) {
      case 4: CGB_RED_LAUNCH(true, false, false); break;
      case 5: CGB_RED_LAUNCH(true, false, true); break;
      case 2: CGB_RED_LAUNCH(false, true, false); break;
      case 3: CGB_RED_LAUNCH(false, true, true); break;
      case 6: CGB_RED_LAUNCH(true, true, false); break;
      case 7: CGB_RED_LAUNCH(true, true, true); break;
      default: CGB_CHECK(false, "in_bwd_reduce: no gradient source");
    }
#undef CGB_RED_LAUNCH
    return;
  }
  const int HW = y.H * y.W;
  const int lanes = std::min(256, y.C / 8), rows = 256 / lanes;
  // every block ends with one atomic per channel on the image's (sum dz, sum dz*xhat) pair: at most 64 blocks
  // per image (measured: 512 blocks per image made this kernel 19 us on a 2 MB tensor, all of it atomics)
  const int blocks_per_image = std::min(64, std::max(16, (2 * 148 + y.N - 1) / y.N));
  const int unit = rows * 2;
  int ppb = (HW + blocks_per_image - 1) / blocks_per_image;
  ppb = std::max(unit, (ppb + unit - 1) / unit * unit);
  dim3 grid((HW + ppb - 1) / ppb, y.N);
  if (async_enabled() && ppb / unit >= 3) {
    static bool configured = false;
    constexpr int kSmem = kAsyncStages * 3 * 2 * 256 * 16;
    if (!configured) {
      CGB_CUDA(cudaFuncSetAttribute(in_bwd_reduce_async_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
      configured = true;
    }
    launch_pdl(in_bwd_reduce_async_kernel<2>, grid, dim3(256), (size_t)kSmem, st, dev(y), stats, dev(g), act,
               da_out ? dev(*da_out) : dev_null(), reinterpret_cast<float*>(bstats), ppb);
    return;
  }
  launch_pdl(in_bwd_reduce_kernel<2>, grid, dim3(256), 0, st, dev(y), stats, dev(g), act,
             da_out ? dev(*da_out) : dev_null(), reinterpret_cast<float*>(bstats), ppb);
}

void in_bwd_apply(const TensorDesc& y, const float2* stats, const float2* bstats, const GradSrc& g, int act,
                  const TensorDesc& dy, cudaStream_t st) {
  check_grad(y, g);
  int segv = 0, nslice = 0;
  if (rows_ok(y, g.g2 ? g.fold : 0) && (act == kActNone || act == kActLeaky || act == kActRelu) &&
      bulk_plan(y.H, y.N, y.W * y.C / 8, &segv, &nslice)) {
    const int vpr = y.W * y.C / 8;
    dim3 grid(y.H * nslice, y.N);
    const size_t smem = (size_t)segv * 16 * (1 + (g.g1 ? 1 : 0) + (g.g2 ? 1 : 0));
#define CGB_APPB_LAUNCH(G1, G2) \
  CGB_ACT_SWITCH(act, (launch_bulk(in_bwd_apply_bulk_kernel<ACT, G1, G2>, grid, smem, st, dev(y), stats, bstats, dev(g), dev(dy), vpr, nslice, segv)))
    if (g.g1 && g.g2) {
      CGB_APPB_LAUNCH(true, true);
    } else if (g.g1) {
      CGB_APPB_LAUNCH(true, false);
    } else {
      CGB_APPB_LAUNCH(false, true);
    }
#undef CGB_APPB_LAUNCH
    return;
  }
  if (rows_ok(y, g.g2 ? g.fold : 0) && (act == kActNone || act == kActLeaky || act == kActRelu)) {
    const int vpr = y.W * y.C / 8;
    int rpb, jsplit;
    rows_grid(y.H, y.N, vpr, 0, &rpb, &jsplit);
    dim3 grid((y.H + rpb - 1) / rpb * jsplit, y.N);
#define CGB_APP_LAUNCH(G1, G2) \
  CGB_ACT_SWITCH(act, (launch_pdl(in_bwd_apply_rows_kernel<ACT, G1, G2, 2>, grid, dim3(256), 0, st, dev(y), stats, bstats, dev(g), dev(dy), vpr, rpb, jsplit)))
    if (g.g1 && g.g2) {
      CGB_APP_LAUNCH(true, true);
    } else if (g.g1) {
      CGB_APP_LAUNCH(true, false);
    } else {
      CGB_APP_LAUNCH(false, true);
    }
#undef CGB_APP_LAUNCH
    return;
  }
  const int total = y.H * y.W;
  const int lanes = std::min(256, y.C / 8), rows = 256 / lanes;
  int appb;
  if (pick_async(total, rows, y.N, 0, &appb)) {
    static bool configured = false;
    constexpr int kSmem = kAsyncStages * 3 * 2 * 256 * 16;
    if (!configured) {
      CGB_CUDA(cudaFuncSetAttribute(in_bwd_apply_async_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
      configured = true;
    }
    dim3 grid((total + appb - 1) / appb, y.N);
    launch_pdl(in_bwd_apply_async_kernel<2>, grid, dim3(256), (size_t)kSmem, st, dev(y), stats, bstats, dev(g), act,
               dev(dy), appb);
    return;
  }
  const int ppb = stream_ppb(total, rows, 2, y.N);
  dim3 grid((total + ppb - 1) / ppb, y.N);
  launch_pdl(in_bwd_apply_kernel<2>, grid, dim3(256), 0, st, dev(y), stats, bstats, dev(g), act, dev(dy), ppb);
}

void tanh_bwd(const TensorDesc& out, const TensorDesc* target, float l1_scale, const GradSrc& g, int C,
              const TensorDesc& dpre, float* loss_slot, cudaStream_t st) {
  if (out.esz == 4) return f32::tanh_bwd(out, target, l1_scale, g, C, dpre, loss_slot, st);
  CGB_CHECK(out.C == 16 && dpre.C == 16 && C <= 8, "tanh_bwd expects 16-channel image tensors");
  if (g.g1 || g.g2) check_grad(out, g);
  const long long total = (long long)out.N * out.H * out.W;
  launch_pdl(tanh_bwd_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, st, dev(out),
             target ? dev(*target) : dev_null(), l1_scale, dev(g), C, dev(dpre), loss_slot);
}

void l1_loss(const TensorDesc& a, const TensorDesc& b, int C, float scale, float* loss_slot, cudaStream_t st) {
  if (a.esz == 4) return f32::l1_loss(a, b, C, scale, loss_slot, st);
  const long long total = (long long)a.N * a.H * a.W;
  l1_loss_kernel<<<blocks_for(total, 256), 256, 0, st>>>(dev(a), dev(b), C, scale, loss_slot);
  CGB_CUDA(cudaGetLastError());
}

void leaky_bwd(const TensorDesc& a, const TensorDesc& g, const TensorDesc& dpre, cudaStream_t st) {
  if (a.esz == 4) return f32::leaky_bwd(a, g, dpre, st);
  const long long total = (long long)a.N * a.H * a.W * (a.C / 8);
  leaky_bwd_kernel<<<blocks_for(total, 256), 256, 0, st>>>(dev(a), dev(g), dev(dpre));
  CGB_CUDA(cudaGetLastError());
}

void mse_loss(const TensorDesc& logits, float target, float w, float* loss_slot, const TensorDesc* dlogits,
              cudaStream_t st) {
  if (logits.esz == 4) return f32::mse_loss(logits, target, w, loss_slot, dlogits, st);
  const long long total = (long long)logits.N * logits.H * logits.W;
  mse_loss_kernel<<<blocks_for(total, 256), 256, 0, st>>>(dev(logits), target, w, loss_slot,
                                                          dlogits ? dev(*dlogits) : dev_null());
  CGB_CUDA(cudaGetLastError());
}

void pack_weights(const float* master, const PackEntry* entries_dev, int n_entries, int max_elems, bf16* arena,
                  cudaStream_t st) {
  if (n_entries == 0) return;
  const int bx = std::max(1, std::min(96, (max_elems + 1023) / 1024));  // blocks per layer (grid-stride over its tiles)
  dim3 grid(bx, n_entries);
  pack_weights_kernel<<<grid, 256, 0, st>>>(master, entries_dev, arena);
  CGB_CUDA(cudaGetLastError());
}

void adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
               float eps, int* step_dev, float* hyper_dev, float grad_scale, cudaStream_t st) {
  (void)lr;  // the live value is hyper_dev[2] (set_device_float), initialised with the configured rate at bind time
  adam_range(p, g, m, v, n, beta1, beta2, eps, step_dev, hyper_dev, grad_scale, true, st);
}

void adam_range(float* p, const float* g, float* m, float* v, long long n, float beta1, float beta2, float eps,
                int* step_dev, float* hyper_dev, float grad_scale, bool advance_step, cudaStream_t st) {
  if (advance_step) {
    adam_prep_kernel<<<1, 1, 0, st>>>(step_dev, hyper_dev, beta1, beta2);
    CGB_CUDA(cudaGetLastError());
  }
  if (n <= 0) return;
  adam_kernel<<<blocks_for((n + 3) / 4, 256), 256, 0, st>>>(p, g, m, v, n, beta1, beta2, eps, hyper_dev, grad_scale);
  CGB_CUDA(cudaGetLastError());
}

void pool_exchange(const TensorDesc& fake, const TensorDesc& pool, const int* dec, const TensorDesc& d_in, cudaStream_t st) {
  CGB_CHECK(fake.C == pool.C && fake.C == d_in.C && fake.H == pool.H && fake.W == pool.W && d_in.N == fake.N &&
                d_in.H == fake.H && d_in.W == fake.W && fake.C % 8 == 0,
            "pool_exchange: shape mismatch");
  const DevTensor f = dev_bytes(fake);
  const long long total = (long long)fake.H * fake.W * (f.C / 8);
  pool_exchange_kernel<<<blocks_for(total, 256), 256, 0, st>>>(f, dev_bytes(pool), dec, dev_bytes(d_in));
  CGB_CUDA(cudaGetLastError());
}

void nhwc_to_u8hwc(const TensorDesc& src, int C, unsigned char* dst, cudaStream_t st) {
  if (src.esz == 4) return f32::nhwc_to_u8hwc(src, C, dst, st);
  const long long total = (long long)src.N * src.H * src.W;
  nhwc_to_u8hwc_kernel<<<blocks_for(total, 256), 256, 0, st>>>(dev(src), C, dst);
  CGB_CUDA(cudaGetLastError());
}

void set_device_float(float* dst, float value, cudaStream_t st) {
  set_float_kernel<<<1, 1, 0, st>>>(dst, value);
  CGB_CUDA(cudaGetLastError());
}

void u8hwc_to_nchw(const unsigned char* src, int N, int H, int W, float* dst, cudaStream_t st) {
  const long long total = (long long)N * H * W;
  u8hwc_to_nchw_kernel<<<blocks_for(total, 256), 256, 0, st>>>(src, N, H, W, dst);
  CGB_CUDA(cudaGetLastError());
}

void expand_rows4(const TensorDesc& src, int k, int sgn, int off, bool use_halo, const TensorDesc& dst, cudaStream_t st) {
  if (src.esz == 4) return;  // validation mode: the 3-channel weight gradients are computed directly (f32::conv_wgrad)
  CGB_CHECK(src.C >= 4 && dst.C == 64 && dst.halo == 0 && k <= 8 && src.N == dst.N, "expand_rows4: bad source / destination");
  const long long total = (long long)dst.N * dst.H * dst.W * 16;
  launch_pdl(expand_rows4_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, st, dev(src), k, sgn, off,
             sgn > 0 ? off : off - (k - 1), use_halo ? -src.halo : 0, dev(dst));
}

void im2col4(const TensorDesc& src, int k, int stride, int sgn, int off, bool use_halo, const TensorDesc& dst,
             cudaStream_t st) {
  if (src.esz == 4) return;  // validation mode: the 3-channel weight gradients are computed directly (f32::conv_wgrad)
  CGB_CHECK(src.C >= 4 && dst.C >= 4 * k * k && dst.halo == 0, "im2col4: bad source / destination");
  const long long total = (long long)dst.N * dst.H * dst.W * k * k;
  launch_pdl(im2col4_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, st, dev(src), k, stride, sgn, off, use_halo ? 1 : 0,
             dev(dst));
}

}  // namespace cgb
