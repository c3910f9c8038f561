// fp32 validation mode kernels (see fp32_path.h): CUDA-core, fp64 accumulation, deterministic.
// Stand-in counterparts (oracle/cyclegan_standin.py): F.conv2d / F.conv_transpose2d and their gradients,
// _inorm (:99), F.pad(mode="reflect"), F.relu / F.leaky_relu / torch.tanh, F.l1_loss, _mse_to (:324).
#include "fp32_path.h"

#include <algorithm>

namespace cgb {
namespace f32 {

namespace {

struct T32 {
  float* p;  // interior origin
  long long sN, sH, sW;
  int N, H, W, C, halo;
};

T32 dev(const TensorDesc& t) {
  CGB_CHECK(t.esz == 4, "fp32 path: tensor is not fp32");
  T32 d;
  d.p = reinterpret_cast<float*>(t.interior());
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
T32 dev_null() {
  T32 d;
  d.p = nullptr;
  d.sN = d.sH = d.sW = 0;
  d.N = d.H = d.W = d.C = d.halo = 0;
  return d;
}
struct G32 {
  T32 g1, g2;
  int fold;
};
G32 dev(const GradSrc& g) {
  G32 d;
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
__device__ __forceinline__ float act_fwd(float x, int act) {
  if (act == kActRelu) return fmaxf(x, 0.f);
  if (act == kActLeaky) return x > 0.f ? x : 0.2f * x;
  if (act == kActTanh) return tanhf(x);
  return x;
}
__device__ __forceinline__ float act_grad(float x, int act) {
  if (act == kActRelu) return x > 0.f ? 1.f : 0.f;
  if (act == kActLeaky) return x > 0.f ? 1.f : 0.2f;
  return 1.f;
}

// gradient w.r.t. an activation at interior pixel (n, h, w), channel c: g1 + the padded-domain gradient g2 folded
// back (the pixel itself plus the one or three halo pixels that mirror onto it); fixed summation order
__device__ __forceinline__ float load_grad(const G32& g, int n, int h, int w, int c, int H, int W) {
  float out = 0.f;
  if (g.g1.p != nullptr) out += g.g1.p[n * g.g1.sN + h * g.g1.sH + w * g.g1.sW + c];
  if (g.g2.p != nullptr) {
    const int p = g.fold;
    const float* base = g.g2.p + n * g.g2.sN + c;
    out += base[(h + p) * g.g2.sH + (w + p) * g.g2.sW];
    const bool hb = (h >= 1 && h <= p) || (h >= H - 1 - p && h <= H - 2);
    const bool wb = (w >= 1 && w <= p) || (w >= W - 1 - p && w <= W - 2);
    const int hm = (h >= 1 && h <= p) ? p - h : 2 * (H - 1) - h + p;
    const int wm = (w >= 1 && w <= p) ? p - w : 2 * (W - 1) - w + p;
    if (hb) out += base[hm * g.g2.sH + (w + p) * g.g2.sW];
    if (wb) out += base[(h + p) * g.g2.sH + wm * g.g2.sW];
    if (hb && wb) out += base[hm * g.g2.sH + wm * g.g2.sW];
  }
  return out;
}

// fixed-order block reduction of one fp64 value per thread (blockDim.x <= 1024); result valid in thread 0
__device__ __forceinline__ double block_sum_det(double v, double* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  return sh[0];
}

// ------------------------------------------------------------------------------------------ convolution
// out[n][oh][ow][co] = act(bias[co] + sum_{tap (r, c), ci} in[n][ih][iw][ci] * w[co * w_so + tap * w_st + ci * w_si])
// with ih = (oh * sn + off + sgn * r) / sd when divisible and inside [lo, hiH), likewise iw.  One CTA computes an
// 8 x 8 pixel tile x 64 output channels; K steps of 16 reduction channels per filter tap; fp64 accumulators.
struct ConvArgs {
  const float* in;
  long long isN, isH, isW;
  int lo, hiH, hiW;
  float* out;
  long long osN, osH, osW;
  int N, Ho, Wo, tiles_h, tiles_w;
  const float* w;
  long long w_so, w_st, w_si;
  int Co, CoS, Ci, k;
  int sn, sd, sgn, off;
  const float* bias;
  int act;
};

constexpr int kTP = 64, kTC = 64, kTK = 16;

__global__ void __launch_bounds__(256) conv_kernel(const ConvArgs a) {
  __shared__ float As[kTK][kTP + 1];
  __shared__ float Bs[kTK][kTC + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  int t = blockIdx.x;
  const int tw = t % a.tiles_w;
  t /= a.tiles_w;
  const int th = t % a.tiles_h;
  const int n = t / a.tiles_h;
  const int co0 = blockIdx.y * kTC;
  const int lp = tid >> 2, lq = tid & 3;  // loader role: pixel of the tile, quad of reduction channels
  const int loh = th * 8 + (lp >> 3), low = tw * 8 + (lp & 7);
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  const int T = a.k * a.k;
  for (int tap = 0; tap < T; ++tap) {
    const int r = tap / a.k, c = tap - r * a.k;
    int hn = loh * a.sn + a.off + a.sgn * r, wn = low * a.sn + a.off + a.sgn * c;
    bool ok = loh < a.Ho && low < a.Wo;
    if (a.sd == 2) {
      ok = ok && ((hn & 1) == 0) && ((wn & 1) == 0);
      hn >>= 1;
      wn >>= 1;
    }
    ok = ok && hn >= a.lo && hn < a.hiH && wn >= a.lo && wn < a.hiW;
    const float* src = a.in + n * a.isN + (long long)hn * a.isH + (long long)wn * a.isW;
    for (int c0 = 0; c0 < a.Ci; c0 += kTK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ci = c0 + lq * 4 + i;
        As[lq * 4 + i][lp] = (ok && ci < a.Ci) ? src[ci] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = tid + j * 256;
        int kk, cc;
        if (a.w_si == 1) {
          kk = idx & 15;
          cc = idx >> 4;
        } else {
          cc = idx & 63;
          kk = idx >> 6;
        }
        const int ci = c0 + kk, co = co0 + cc;
        Bs[kk][cc] = (ci < a.Ci && co < a.Co) ? a.w[co * a.w_so + tap * a.w_st + ci * a.w_si] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < kTK; ++kk) {
        double av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = (double)As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = (double)Bs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = ty * 4 + i;
    const int oh = th * 8 + (p >> 3), ow = tw * 8 + (p & 7);
    if (oh >= a.Ho || ow >= a.Wo) continue;
    float* o = a.out + n * a.osN + (long long)oh * a.osH + (long long)ow * a.osW;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co >= a.CoS) continue;
      float v = 0.f;
      if (co < a.Co) {
        double s = acc[i][j];
        if (a.bias != nullptr) s += (double)a.bias[co];
        v = act_fwd((float)s, a.act);
      }
      o[co] = v;
    }
  }
}

void launch_conv(ConvArgs a, cudaStream_t st) {
  a.tiles_h = (a.Ho + 7) / 8;
  a.tiles_w = (a.Wo + 7) / 8;
  dim3 grid((unsigned)(a.N * a.tiles_h * a.tiles_w), (unsigned)((a.CoS + kTC - 1) / kTC));
  conv_kernel<<<grid, 256, 0, st>>>(a);
  CGB_CUDA(cudaGetLastError());
}

// g[co][tap][ci] = sum over the pixel domain d of A[n][d * sa + offa + r * ga][co] * B[n][d * sb + offb + r * gb][ci]
// (A = dy, B = x; conv: the domain is the output, B walks the input; transposed conv: the domain is the input, A walks
// the output).  One CTA = 64 x 64 (co, ci) outputs of one filter tap over the WHOLE domain, in pixel order: deterministic.
struct WgradArgs32 {
  const float* a;
  long long asN, asH, asW;
  int aH, aW, aC;
  const float* b;
  long long bsN, bsH, bsW;
  int bLo, bHiH, bHiW, bC;
  int N, Hd, Wd;
  int sa, offa, ga, sb, offb, gb;
  int Co, Ci, T, k, ci_tiles;
  float* g;
};

__global__ void __launch_bounds__(256) wgrad_kernel32(const WgradArgs32 a) {
  __shared__ __align__(16) float As[kTK][kTC + 4];
  __shared__ __align__(16) float Bs[kTK][kTC + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int co0 = (blockIdx.x / a.ci_tiles) * kTC, ci0 = (blockIdx.x % a.ci_tiles) * kTC;
  const int tap = blockIdx.y;
  const int r = tap / a.k, c = tap - r * a.k;
  const int lk = tid >> 4, lq = tid & 15;  // loader role: pixel of the chunk, channel quad
  const long long HW = (long long)a.Hd * a.Wd, total = (long long)a.N * HW;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (long long p0 = 0; p0 < total; p0 += kTK) {
    const long long P = p0 + lk;
    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
    if (P < total) {
      const int n = (int)(P / HW);
      const int rem = (int)(P - (long long)n * HW);
      const int dh = rem / a.Wd, dw = rem - dh * a.Wd;
      const int ha = dh * a.sa + a.offa + r * a.ga, wa = dw * a.sa + a.offa + c * a.ga;
      const int hb = dh * a.sb + a.offb + r * a.gb, wb = dw * a.sb + a.offb + c * a.gb;
      const int ca = co0 + lq * 4, cb = ci0 + lq * 4;
      if (ha >= 0 && ha < a.aH && wa >= 0 && wa < a.aW && ca + 3 < a.aC)
        va = *reinterpret_cast<const float4*>(a.a + n * a.asN + (long long)ha * a.asH + (long long)wa * a.asW + ca);
      if (hb >= a.bLo && hb < a.bHiH && wb >= a.bLo && wb < a.bHiW && cb + 3 < a.bC)
        vb = *reinterpret_cast<const float4*>(a.b + n * a.bsN + (long long)hb * a.bsH + (long long)wb * a.bsW + cb);
    }
    *reinterpret_cast<float4*>(&As[lk][lq * 4]) = va;
    *reinterpret_cast<float4*>(&Bs[lk][lq * 4]) = vb;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kTK; ++kk) {
      double av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = (double)As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = (double)Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= a.Co) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < a.Ci) a.g[((long long)co * a.T + tap) * a.Ci + ci] = (float)acc[i][j];
    }
  }
}

// ------------------------------------------------------------------------------------------ InstanceNorm
// grid (C / 32, N), block 1024 = 32 channels x 32 pixel slices.  Phase 1: per-thread strided fp64 partial sums, then a
// fixed-order tree over the 32 slices; phase 2: normalise.  One CTA owns (image, 32 channels) entirely: deterministic.
__device__ __forceinline__ void slice_reduce2(double& s1, double& s2, double (*sh)[32][33]) {
  const int cx = threadIdx.x & 31, wy = threadIdx.x >> 5;
  sh[0][wy][cx] = s1;
  sh[1][wy][cx] = s2;
  __syncthreads();
  for (int s = 16; s > 0; s >>= 1) {
    if (wy < s) {
      sh[0][wy][cx] += sh[0][wy + s][cx];
      sh[1][wy][cx] += sh[1][wy + s][cx];
    }
    __syncthreads();
  }
  s1 = sh[0][0][cx];
  s2 = sh[1][0][cx];
  __syncthreads();
}

__global__ void __launch_bounds__(1024) in_forward_kernel(T32 y, float2* __restrict__ stats, int act, T32 res, T32 out) {
  __shared__ double sh[2][32][33];
  const int cx = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx, n = blockIdx.y;
  const bool cok = c < y.C;
  const int HW = y.H * y.W;
  double s1 = 0.0, s2 = 0.0;
  if (cok) {
    for (int p = wy; p < HW; p += 32) {
      const int h = p / y.W, w = p - h * y.W;
      const double v = (double)y.p[n * y.sN + h * y.sH + w * y.sW + c];
      s1 += v;
      s2 = fma(v, v, s2);
    }
  }
  slice_reduce2(s1, s2, sh);
  const double mean_d = s1 / (double)HW;
  const double var_d = fmax(s2 / (double)HW - mean_d * mean_d, 0.0);
  const float mean = (float)mean_d;
  const float rstd = (float)(1.0 / sqrt(var_d + 1e-5));
  if (!cok) return;
  if (wy == 0) stats[(long long)n * y.C + c] = make_float2(mean, rstd);
  const int HP = out.H + 2 * out.halo, WP = out.W + 2 * out.halo;
  for (int p = wy; p < HP * WP; p += 32) {
    const int hp = p / WP, wp = p - hp * WP;
    const int h = reflect_idx(hp - out.halo, out.H), w = reflect_idx(wp - out.halo, out.W);
    float v = act_fwd((y.p[n * y.sN + h * y.sH + w * y.sW + c] - mean) * rstd, act);
    if (res.p != nullptr) v += res.p[n * res.sN + h * res.sH + w * res.sW + c];
    out.p[n * out.sN + (long long)(hp - out.halo) * out.sH + (long long)(wp - out.halo) * out.sW + c] = v;
  }
}

__global__ void __launch_bounds__(1024) in_backward_kernel(T32 y, const float2* __restrict__ stats, G32 g, int act, T32 da,
                                                            T32 dy) {
  __shared__ double sh[2][32][33];
  const int cx = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx, n = blockIdx.y;
  const bool cok = c < y.C;
  const int HW = y.H * y.W;
  float mean = 0.f, rstd = 0.f;
  double s1 = 0.0, s2 = 0.0;
  if (cok) {
    const float2 st = stats[(long long)n * y.C + c];
    mean = st.x;
    rstd = st.y;
    for (int p = wy; p < HW; p += 32) {
      const int h = p / y.W, w = p - h * y.W;
      const float gr = load_grad(g, n, h, w, c, y.H, y.W);
      if (da.p != nullptr) da.p[n * da.sN + h * da.sH + w * da.sW + c] = gr;
      const float xh = (y.p[n * y.sN + h * y.sH + w * y.sW + c] - mean) * rstd;
      const float dz = gr * act_grad(xh, act);
      s1 += (double)dz;
      s2 = fma((double)dz, (double)xh, s2);
    }
  }
  slice_reduce2(s1, s2, sh);
  if (!cok) return;
  const float m1 = (float)(s1 / (double)HW), m2 = (float)(s2 / (double)HW);
  for (int p = wy; p < HW; p += 32) {
    const int h = p / y.W, w = p - h * y.W;
    const float gr = load_grad(g, n, h, w, c, y.H, y.W);
    const float xh = (y.p[n * y.sN + h * y.sH + w * y.sW + c] - mean) * rstd;
    const float dz = gr * act_grad(xh, act);
    dy.p[n * dy.sN + h * dy.sH + w * dy.sW + c] = rstd * (dz - m1 - xh * m2);
  }
}

// ------------------------------------------------------------------------------------------ layout
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int C, T32 dst) {
  const int HP = dst.H + 2 * dst.halo, WP = dst.W + 2 * dst.halo;
  const long long total = (long long)dst.N * HP * WP;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wp = idx % WP;
  const int hp = (idx / WP) % HP;
  const int n = idx / ((long long)WP * HP);
  const int h = reflect_idx(hp - dst.halo, dst.H), w = reflect_idx(wp - dst.halo, dst.W);
  float* o = dst.p + n * dst.sN + (long long)(hp - dst.halo) * dst.sH + (long long)(wp - dst.halo) * dst.sW;
  for (int c = 0; c < dst.C; ++c) o[c] = c < C ? src[(((long long)n * C + c) * dst.H + h) * dst.W + w] : 0.f;
}

__global__ void nhwc_to_nchw_kernel(T32 src, int C, float* __restrict__ dst) {
  const long long total = (long long)src.N * src.H * src.W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int w = idx % src.W;
  const int h = (idx / src.W) % src.H;
  const int n = idx / ((long long)src.W * src.H);
  const float* p = src.p + n * src.sN + h * src.sH + w * src.sW;
  for (int c = 0; c < C; ++c) dst[(((long long)n * C + c) * src.H + h) * src.W + w] = p[c];
}

__global__ void nhwc_to_u8hwc_kernel(T32 src, int C, unsigned char* __restrict__ dst) {
  const long long total = (long long)src.N * src.H * src.W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int w = idx % src.W;
  const int h = (idx / src.W) % src.H;
  const int n = idx / ((long long)src.W * src.H);
  const float* p = src.p + n * src.sN + h * src.sH + w * src.sW;
  for (int c = 0; c < C; ++c) {
    const float v = __fmul_rn(__fadd_rn(p[c], 1.f), 127.5f);
    dst[idx * C + c] = (unsigned char)min(255, max(0, __float2int_rn(v)));
  }
}

// ------------------------------------------------------------------------------------------ head / losses
// ONE block walks every pixel (deterministic loss sums; the tensors involved are the 3-channel images / 1-channel
// logits, a few hundred thousand elements)
__global__ void __launch_bounds__(1024) tanh_bwd_kernel(T32 out, T32 target, float l1_scale, G32 g, int C, T32 dpre,
                                                         float* __restrict__ loss_slot) {
  __shared__ double sh[1024];
  const long long total = (long long)out.N * out.H * out.W;
  double lsum = 0.0;
  for (long long idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int w = idx % out.W;
    const int h = (idx / out.W) % out.H;
    const int n = idx / ((long long)out.W * out.H);
    float* dp = dpre.p + n * dpre.sN + h * dpre.sH + w * dpre.sW;
    for (int c = 0; c < dpre.C; ++c) {
      float d = 0.f;
      if (c < C) {
        const float o = out.p[n * out.sN + h * out.sH + w * out.sW + c];
        float gr = load_grad(g, n, h, w, c, out.H, out.W);
        if (target.p != nullptr) {
          const float diff = o - target.p[n * target.sN + h * target.sH + w * target.sW + c];
          lsum += (double)fabsf(diff);
          gr += l1_scale * (float)((diff > 0.f) - (diff < 0.f));
        }
        d = gr * (1.f - o * o);
      }
      dp[c] = d;
    }
  }
  const double t = block_sum_det(lsum, sh);
  if (threadIdx.x == 0 && loss_slot != nullptr && target.p != nullptr) *loss_slot += (float)(t * (double)l1_scale);
}

__global__ void __launch_bounds__(1024) l1_loss_kernel(T32 a, T32 b, int C, float scale, float* __restrict__ loss_slot) {
  __shared__ double sh[1024];
  const long long total = (long long)a.N * a.H * a.W;
  double lsum = 0.0;
  for (long long idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int w = idx % a.W;
    const int h = (idx / a.W) % a.H;
    const int n = idx / ((long long)a.W * a.H);
    for (int c = 0; c < C; ++c)
      lsum += (double)fabsf(a.p[n * a.sN + h * a.sH + w * a.sW + c] - b.p[n * b.sN + h * b.sH + w * b.sW + c]);
  }
  const double t = block_sum_det(lsum, sh);
  if (threadIdx.x == 0) *loss_slot += (float)(t * (double)scale);
}

__global__ void leaky_bwd_kernel(T32 a, T32 g, T32 dpre) {
  const long long total = (long long)a.N * a.H * a.W * a.C;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = idx % a.C;
  long long r = idx / a.C;
  const int w = r % a.W;
  r /= a.W;
  const int h = r % a.H;
  const int n = r / a.H;
  const float av = a.p[n * a.sN + h * a.sH + w * a.sW + c];
  const float gv = g.p[n * g.sN + h * g.sH + w * g.sW + c];
  dpre.p[n * dpre.sN + h * dpre.sH + w * dpre.sW + c] = gv * (av > 0.f ? 1.f : 0.2f);
}

__global__ void __launch_bounds__(1024) mse_loss_kernel(T32 logits, float target, float wgt, float* __restrict__ loss_slot,
                                                         T32 dl) {
  __shared__ double sh[1024];
  const long long total = (long long)logits.N * logits.H * logits.W;
  const float inv = 1.f / (float)total;
  double lsum = 0.0;
  for (long long idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int w = idx % logits.W;
    const int h = (idx / logits.W) % logits.H;
    const int n = idx / ((long long)logits.W * logits.H);
    const float d = logits.p[n * logits.sN + h * logits.sH + w * logits.sW] - target;
    lsum = fma((double)d, (double)d, lsum);
    if (dl.p != nullptr) {
      float* o = dl.p + n * dl.sN + h * dl.sH + w * dl.sW;
      o[0] = 2.f * wgt * d * inv;
      for (int c = 1; c < dl.C; ++c) o[c] = 0.f;
    }
  }
  const double t = block_sum_det(lsum, sh);
  if (threadIdx.x == 0 && loss_slot != nullptr) *loss_slot += (float)(t * (double)wgt / (double)total);
}

// grid (ceil(C / 32)), block 1024 = 32 channels x 32 pixel slices over ALL images: gbias[c] = sum
__global__ void __launch_bounds__(1024) bias_grad_kernel(T32 dy, int C, float* __restrict__ gbias) {
  __shared__ double sh[2][32][33];
  const int cx = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const long long HW = (long long)dy.H * dy.W, total = (long long)dy.N * HW;
  double s1 = 0.0, s2 = 0.0;
  if (c < C) {
    for (long long p = wy; p < total; p += 32) {
      const int n = (int)(p / HW);
      const int rem = (int)(p - (long long)n * HW);
      const int h = rem / dy.W, w = rem - h * dy.W;
      s1 += (double)dy.p[n * dy.sN + h * dy.sH + w * dy.sW + c];
    }
  }
  slice_reduce2(s1, s2, sh);
  if (c < C && wy == 0) gbias[c] = (float)s1;
}

__global__ void sum_slots_kernel(float* __restrict__ dst, const float* __restrict__ s0, const float* __restrict__ s1,
                                 const float* __restrict__ s2, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = s0[i] + s1[i];
  if (s2 != nullptr) v += s2[i];
  dst[i] = v;
}

inline unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

// ================================================================================================ host API
void conv_fprop(const ConvSpec& s, const TensorDesc& x, const float* w, const float* bias, int act, const TensorDesc& y,
                cudaStream_t st) {
  CGB_CHECK(x.C == s.CinS && y.C == s.CoutS, "fp32 fprop: tensor channels do not match the spec");
  CGB_CHECK(y.H == out_extent(s, x.H) && y.W == out_extent(s, x.W) && y.N == x.N, "fp32 fprop: output extent mismatch");
  const T32 X = dev(x), Y = dev(y);
  ConvArgs a{};
  a.in = X.p;
  a.isN = X.sN;
  a.isH = X.sH;
  a.isW = X.sW;
  a.out = Y.p;
  a.osN = Y.sN;
  a.osH = Y.sH;
  a.osW = Y.sW;
  a.N = x.N;
  a.Ho = y.H;
  a.Wo = y.W;
  a.w = w;
  a.w_so = (long long)s.taps() * s.Cin;
  a.w_st = s.Cin;
  a.w_si = 1;
  a.Co = s.Cout;
  a.CoS = s.CoutS;
  a.Ci = s.Cin;
  a.k = s.k;
  a.bias = bias;
  a.act = act;
  if (!s.transposed) {
    if (s.reflect) CGB_CHECK(x.halo == s.pad && s.stride == 1, "reflect conv input must carry a halo equal to the padding");
    const int halo = s.reflect ? x.halo : 0;
    a.lo = -halo;
    a.hiH = x.H + halo;
    a.hiW = x.W + halo;
    a.sn = s.stride;
    a.sd = 1;
    a.sgn = 1;
    a.off = -s.pad;
  } else {
    a.lo = 0;
    a.hiH = x.H;
    a.hiW = x.W;
    a.sn = 1;
    a.sd = 2;
    a.sgn = -1;
    a.off = s.pad;
  }
  launch_conv(a, st);
}

void conv_dgrad(const ConvSpec& s, const TensorDesc& dy, const float* w, const TensorDesc& dx, cudaStream_t st) {
  CGB_CHECK(dy.C == s.CoutS && dx.C == s.CinS, "fp32 dgrad: tensor channels do not match the spec");
  const T32 DY = dev(dy), DX = dev(dx);
  ConvArgs a{};
  a.in = DY.p;
  a.isN = DY.sN;
  a.isH = DY.sH;
  a.isW = DY.sW;
  a.lo = 0;
  a.hiH = dy.H;
  a.hiW = dy.W;
  a.out = DX.p;
  a.osN = DX.sN;
  a.osH = DX.sH;
  a.osW = DX.sW;
  a.N = dx.N;
  a.Ho = dx.H;
  a.Wo = dx.W;
  a.w = w;
  a.w_so = 1;
  a.w_st = s.Cin;
  a.w_si = (long long)s.taps() * s.Cin;
  a.Co = s.Cin;
  a.CoS = s.CinS;
  a.Ci = s.Cout;
  a.k = s.k;
  a.bias = nullptr;
  a.act = kActNone;
  if (!s.transposed) {
    // dx[i] = sum_r dy[(i + pad - r) / stride] w[r]; reflect: dx is the padded-domain tensor, index hp = i + pad
    a.sn = 1;
    a.sd = s.stride;
    a.sgn = -1;
    a.off = s.reflect ? 0 : s.pad;
    CGB_CHECK(s.stride == 1 || s.stride == 2, "fp32 dgrad: stride must be 1 or 2");
  } else {
    // transposed conv: dx[i] = sum_r dy[2 i - pad + r] w[r]
    a.sn = 2;
    a.sd = 1;
    a.sgn = 1;
    a.off = -s.pad;
  }
  launch_conv(a, st);
}

void conv_wgrad(const ConvSpec& s, const TensorDesc& x, const TensorDesc& dy, float* g, cudaStream_t st) {
  CGB_CHECK(x.C == s.CinS && dy.C == s.CoutS, "fp32 wgrad: tensor channels do not match the spec");
  const T32 X = dev(x), DY = dev(dy);
  WgradArgs32 a{};
  a.a = DY.p;
  a.asN = DY.sN;
  a.asH = DY.sH;
  a.asW = DY.sW;
  a.aH = dy.H;
  a.aW = dy.W;
  a.aC = dy.C;
  a.b = X.p;
  a.bsN = X.sN;
  a.bsH = X.sH;
  a.bsW = X.sW;
  a.bC = x.C;
  a.N = x.N;
  a.Co = s.Cout;
  a.Ci = s.Cin;
  a.T = s.taps();
  a.k = s.k;
  a.g = g;
  if (!s.transposed) {
    const int halo = s.reflect ? x.halo : 0;
    if (s.reflect) CGB_CHECK(x.halo == s.pad, "reflect conv input must carry a halo equal to the padding");
    a.bLo = -halo;
    a.bHiH = x.H + halo;
    a.bHiW = x.W + halo;
    a.Hd = dy.H;
    a.Wd = dy.W;
    a.sa = 1;
    a.offa = 0;
    a.ga = 0;
    a.sb = s.stride;
    a.offb = -s.pad;
    a.gb = 1;
  } else {
    a.bLo = 0;
    a.bHiH = x.H;
    a.bHiW = x.W;
    a.Hd = x.H;
    a.Wd = x.W;
    a.sa = 2;
    a.offa = -s.pad;
    a.ga = 1;
    a.sb = 1;
    a.offb = 0;
    a.gb = 0;
  }
  a.ci_tiles = (s.Cin + kTC - 1) / kTC;
  dim3 grid((unsigned)(((s.Cout + kTC - 1) / kTC) * a.ci_tiles), (unsigned)a.T);
  wgrad_kernel32<<<grid, 256, 0, st>>>(a);
  CGB_CUDA(cudaGetLastError());
}

void in_forward(const TensorDesc& y, float2* stats, int act, const TensorDesc* residual, const TensorDesc& out,
                cudaStream_t st) {
  CGB_CHECK(y.C == out.C && y.H == out.H && y.W == out.W && y.N == out.N, "fp32 in_forward: shape mismatch");
  dim3 grid((unsigned)((y.C + 31) / 32), (unsigned)y.N);
  in_forward_kernel<<<grid, 1024, 0, st>>>(dev(y), stats, act, residual ? dev(*residual) : dev_null(), dev(out));
  CGB_CUDA(cudaGetLastError());
}

void in_backward(const TensorDesc& y, const float2* stats, const GradSrc& g, int act, const TensorDesc* da_store,
                 const TensorDesc& dy, cudaStream_t st) {
  CGB_CHECK(g.g1 || g.g2, "fp32 in_backward: gradient source is empty");
  if (g.g1) CGB_CHECK(g.g1->H == y.H && g.g1->W == y.W && g.g1->C == y.C, "g1 shape mismatch");
  if (g.g2)
    CGB_CHECK(g.g2->H == y.H + 2 * g.fold && g.g2->W == y.W + 2 * g.fold && g.g2->C == y.C && g.g2->halo == 0,
              "g2 (padded-domain gradient) shape mismatch");
  dim3 grid((unsigned)((y.C + 31) / 32), (unsigned)y.N);
  in_backward_kernel<<<grid, 1024, 0, st>>>(dev(y), stats, dev(g), act, da_store ? dev(*da_store) : dev_null(), dev(dy));
  CGB_CUDA(cudaGetLastError());
}

void nchw_to_nhwc(const float* src, int C, const TensorDesc& dst, cudaStream_t st) {
  const long long total = (long long)dst.N * (dst.H + 2 * dst.halo) * (dst.W + 2 * dst.halo);
  nchw_to_nhwc_kernel<<<blocks_for(total, 256), 256, 0, st>>>(src, C, dev(dst));
  CGB_CUDA(cudaGetLastError());
}

void nhwc_to_nchw(const TensorDesc& src, int C, float* dst, cudaStream_t st) {
  const long long total = (long long)src.N * src.H * src.W;
  nhwc_to_nchw_kernel<<<blocks_for(total, 256), 256, 0, st>>>(dev(src), C, dst);
  CGB_CUDA(cudaGetLastError());
}

void nhwc_to_u8hwc(const TensorDesc& src, int C, unsigned char* dst, cudaStream_t st) {
  const long long total = (long long)src.N * src.H * src.W;
  nhwc_to_u8hwc_kernel<<<blocks_for(total, 256), 256, 0, st>>>(dev(src), C, dst);
  CGB_CUDA(cudaGetLastError());
}

void tanh_bwd(const TensorDesc& out, const TensorDesc* target, float l1_scale, const GradSrc& g, int C,
              const TensorDesc& dpre, float* loss_slot, cudaStream_t st) {
  tanh_bwd_kernel<<<1, 1024, 0, st>>>(dev(out), target ? dev(*target) : dev_null(), l1_scale, dev(g), C, dev(dpre),
                                      loss_slot);
  CGB_CUDA(cudaGetLastError());
}

void l1_loss(const TensorDesc& a, const TensorDesc& b, int C, float scale, float* loss_slot, cudaStream_t st) {
  l1_loss_kernel<<<1, 1024, 0, st>>>(dev(a), dev(b), C, scale, loss_slot);
  CGB_CUDA(cudaGetLastError());
}

void leaky_bwd(const TensorDesc& a, const TensorDesc& g, const TensorDesc& dpre, cudaStream_t st) {
  const long long total = (long long)a.N * a.H * a.W * a.C;
  leaky_bwd_kernel<<<blocks_for(total, 256), 256, 0, st>>>(dev(a), dev(g), dev(dpre));
  CGB_CUDA(cudaGetLastError());
}

void mse_loss(const TensorDesc& logits, float target, float w, float* loss_slot, const TensorDesc* dlogits,
              cudaStream_t st) {
  mse_loss_kernel<<<1, 1024, 0, st>>>(dev(logits), target, w, loss_slot, dlogits ? dev(*dlogits) : dev_null());
  CGB_CUDA(cudaGetLastError());
}

void bias_grad(const TensorDesc& dy, int C, float* gbias, cudaStream_t st) {
  bias_grad_kernel<<<(unsigned)((C + 31) / 32), 1024, 0, st>>>(dev(dy), C, gbias);
  CGB_CUDA(cudaGetLastError());
}

void sum_slots(float* dst, const float* s0, const float* s1, const float* s2, long long n, cudaStream_t st) {
  sum_slots_kernel<<<blocks_for(n, 256), 256, 0, st>>>(dst, s0, s1, s2, n);
  CGB_CUDA(cudaGetLastError());
}

}  // namespace f32
}  // namespace cgb
