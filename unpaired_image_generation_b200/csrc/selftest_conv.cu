// Stand-alone GPU self-test of the tcgen05 convolution kernels against a plain CPU loop nest
// (development tool; the graded parity tests live in tests/ and go through the C ABI).
//   selftest_conv <case> [N] [H]
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "conv_plan.h"

using namespace cgb;

namespace cgb {
void set_last_error(const std::string&) {}
}  // namespace cgb

static float bfr(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

static int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

struct Host {
  ConvSpec s;
  int N, H, W, Ho, Wo;
  std::vector<float> x;   // [N][H][W][Cin] logical
  std::vector<float> w;   // [Cout][T][Cin] master layout
  std::vector<float> b;   // [Cout]
  std::vector<float> dy;  // [N][Ho][Wo][Cout]
};

// y = conv(x) (+bias); logical channels
static void ref_fprop(const Host& h, std::vector<float>& y) {
  const ConvSpec& s = h.s;
  const int k = s.k, T = k * k;
  y.assign((size_t)h.N * h.Ho * h.Wo * s.Cout, 0.f);
#pragma omp parallel for collapse(2)
  for (int n = 0; n < h.N; ++n)
    for (int oh = 0; oh < h.Ho; ++oh)
      for (int ow = 0; ow < h.Wo; ++ow)
        for (int co = 0; co < s.Cout; ++co) {
          double acc = 0;
          for (int r = 0; r < k; ++r)
            for (int c = 0; c < k; ++c) {
              int ih, iw;
              if (!s.transposed) {
                ih = oh * s.stride + r - s.pad;
                iw = ow * s.stride + c - s.pad;
                if (s.reflect) {
                  ih = reflect_idx(ih, h.H);
                  iw = reflect_idx(iw, h.W);
                }
              } else {
                const int th = oh + s.pad - r, tw = ow + s.pad - c;
                if (th % 2 != 0 || tw % 2 != 0 || th < 0 || tw < 0) continue;
                ih = th / 2;
                iw = tw / 2;
              }
              if (ih < 0 || ih >= h.H || iw < 0 || iw >= h.W) continue;
              const float* xp = &h.x[(((size_t)n * h.H + ih) * h.W + iw) * s.Cin];
              const float* wp = &h.w[((size_t)co * T + r * k + c) * s.Cin];
              for (int ci = 0; ci < s.Cin; ++ci) acc += (double)xp[ci] * wp[ci];
            }
          y[(((size_t)n * h.Ho + oh) * h.Wo + ow) * s.Cout + co] = (float)acc + h.b[co];
        }
}

// exact adjoint w.r.t. the (padded, when reflect) input
static void ref_dgrad(const Host& h, std::vector<float>& dx, int* DH, int* DW) {
  const ConvSpec& s = h.s;
  const int k = s.k, T = k * k;
  const int p = s.reflect ? s.pad : 0;
  *DH = h.H + 2 * p;
  *DW = h.W + 2 * p;
  std::vector<double> acc((size_t)h.N * *DH * *DW * s.Cin, 0.0);
  for (int n = 0; n < h.N; ++n)
    for (int oh = 0; oh < h.Ho; ++oh)
      for (int ow = 0; ow < h.Wo; ++ow)
        for (int r = 0; r < k; ++r)
          for (int c = 0; c < k; ++c) {
            int ih, iw;
            if (!s.transposed) {
              ih = oh * s.stride + r - (s.reflect ? 0 : s.pad);
              iw = ow * s.stride + c - (s.reflect ? 0 : s.pad);
            } else {
              const int th = oh + s.pad - r, tw = ow + s.pad - c;
              if (th % 2 != 0 || tw % 2 != 0 || th < 0 || tw < 0) continue;
              ih = th / 2;
              iw = tw / 2;
            }
            if (ih < 0 || ih >= *DH || iw < 0 || iw >= *DW) continue;
            const float* dyp = &h.dy[(((size_t)n * h.Ho + oh) * h.Wo + ow) * s.Cout];
            double* dxp = &acc[(((size_t)n * *DH + ih) * *DW + iw) * s.Cin];
            for (int co = 0; co < s.Cout; ++co) {
              const float* wp = &h.w[((size_t)co * T + r * k + c) * s.Cin];
              const double g = dyp[co];
              for (int ci = 0; ci < s.Cin; ++ci) dxp[ci] += g * wp[ci];
            }
          }
  dx.resize(acc.size());
  for (size_t i = 0; i < acc.size(); ++i) dx[i] = (float)acc[i];
}

static void ref_wgrad(const Host& h, std::vector<float>& g) {
  const ConvSpec& s = h.s;
  const int k = s.k, T = k * k;
  g.assign((size_t)s.Cout * T * s.Cin, 0.f);
#pragma omp parallel for
  for (int co = 0; co < s.Cout; ++co) {
    std::vector<double> acc((size_t)T * s.Cin, 0.0);
    for (int n = 0; n < h.N; ++n)
      for (int oh = 0; oh < h.Ho; ++oh)
        for (int ow = 0; ow < h.Wo; ++ow) {
          const double d = h.dy[(((size_t)n * h.Ho + oh) * h.Wo + ow) * s.Cout + co];
          for (int r = 0; r < k; ++r)
            for (int c = 0; c < k; ++c) {
              int ih, iw;
              if (!s.transposed) {
                ih = oh * s.stride + r - s.pad;
                iw = ow * s.stride + c - s.pad;
                if (s.reflect) {
                  ih = reflect_idx(ih, h.H);
                  iw = reflect_idx(iw, h.W);
                }
              } else {
                const int th = oh + s.pad - r, tw = ow + s.pad - c;
                if (th % 2 != 0 || tw % 2 != 0 || th < 0 || tw < 0) continue;
                ih = th / 2;
                iw = tw / 2;
              }
              if (ih < 0 || ih >= h.H || iw < 0 || iw >= h.W) continue;
              const float* xp = &h.x[(((size_t)n * h.H + ih) * h.W + iw) * s.Cin];
              double* ap = &acc[(size_t)(r * k + c) * s.Cin];
              for (int ci = 0; ci < s.Cin; ++ci) ap[ci] += d * xp[ci];
            }
        }
    for (size_t i = 0; i < acc.size(); ++i) g[(size_t)co * T * s.Cin + i] = (float)acc[i];
  }
}

template <typename T>
static T* dev_upload(const std::vector<T>& v) {
  T* d = nullptr;
  CGB_CUDA(cudaMalloc(&d, v.size() * sizeof(T)));
  CGB_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

// logical NHWC float -> stored bf16 with channel padding and (optional) reflect halo
static std::vector<bf16> to_stored(const std::vector<float>& v, int N, int H, int W, int C, int CS, int halo,
                                   bool reflect_fill) {
  const int HP = H + 2 * halo, WP = W + 2 * halo;
  std::vector<bf16> out((size_t)N * HP * WP * CS, __float2bfloat16_rn(0.f));
  for (int n = 0; n < N; ++n)
    for (int hp = 0; hp < HP; ++hp)
      for (int wp = 0; wp < WP; ++wp) {
        int h = hp - halo, w = wp - halo;
        if (h < 0 || h >= H || w < 0 || w >= W) {
          if (!reflect_fill) continue;
          h = reflect_idx(h, H);
          w = reflect_idx(w, W);
        }
        for (int c = 0; c < C; ++c)
          out[(((size_t)n * HP + hp) * WP + wp) * CS + c] = __float2bfloat16_rn(v[(((size_t)n * H + h) * W + w) * C + c]);
      }
  return out;
}

static double compare(const char* what, const std::vector<float>& ref, const std::vector<float>& got) {
  double maxref = 0, maxerr = 0, se = 0, sr = 0;
  size_t worst = 0;
  for (size_t i = 0; i < ref.size(); ++i) {
    maxref = std::max(maxref, (double)std::fabs(ref[i]));
    const double e = std::fabs((double)ref[i] - got[i]);
    if (e > maxerr) {
      maxerr = e;
      worst = i;
    }
    se += e * e;
    sr += (double)ref[i] * ref[i];
  }
  const double rel = std::sqrt(se / std::max(sr, 1e-30));
  printf("  %-8s max|ref| %.4f  max|err| %.5f (at %zu: ref %.5f got %.5f)  rel-l2 %.3e  -> %s\n", what, maxref, maxerr,
         worst, ref.empty() ? 0.f : ref[worst], got.empty() ? 0.f : got[worst], rel, rel < 1e-2 ? "OK" : "FAIL");
  return rel;
}

static int run_case(const std::string& name, ConvSpec s, int N, int H, int W, int act, int passes, int force_bn) {
  // CGB_PASSES: bit mask (1 fprop, 2 dgrad, 4 wgrad) restricting the passes; CGB_TIMING_ONLY: skip the CPU loops
  if (getenv("CGB_PASSES")) passes &= atoi(getenv("CGB_PASSES"));
  const bool timing_only = getenv("CGB_TIMING_ONLY") != nullptr;
  printf("case %s: N=%d H=%d W=%d Cin=%d(%d) Cout=%d(%d) k=%d s=%d p=%d reflect=%d transposed=%d\n", name.c_str(), N,
         H, W, s.Cin, s.CinS, s.Cout, s.CoutS, s.k, s.stride, s.pad, (int)s.reflect, (int)s.transposed);
  Host h;
  h.s = s;
  h.N = N;
  h.H = H;
  h.W = W;
  h.Ho = out_extent(s, H);
  h.Wo = out_extent(s, W);
  const int T = s.taps();
  std::mt19937 rng(1234);
  std::normal_distribution<float> nd(0.f, 1.f);
  h.x.resize((size_t)N * H * W * s.Cin);
  for (auto& v : h.x) v = bfr(nd(rng));
  h.w.resize((size_t)s.Cout * T * s.Cin);
  for (auto& v : h.w) v = bfr(0.05f * nd(rng));
  h.b.resize(s.Cout);
  for (auto& v : h.b) v = 0.1f * nd(rng);
  h.dy.resize((size_t)N * h.Ho * h.Wo * s.Cout);
  for (auto& v : h.dy) v = bfr(nd(rng));

  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, 0);
  int fails = 0;
  const int halo = s.reflect ? s.pad : 0;

  // device tensors
  TensorDesc x{nullptr, N, H, W, s.CinS, halo};
  x.ptr = dev_upload(to_stored(h.x, N, H, W, s.Cin, s.CinS, halo, true));
  TensorDesc y{nullptr, N, h.Ho, h.Wo, s.CoutS, 0};
  CGB_CUDA(cudaMalloc(&y.ptr, y.elems() * sizeof(bf16)));
  CGB_CUDA(cudaMemset(y.ptr, 0xFF, y.elems() * sizeof(bf16)));
  TensorDesc dy{nullptr, N, h.Ho, h.Wo, s.CoutS, 0};
  dy.ptr = dev_upload(to_stored(h.dy, N, h.Ho, h.Wo, s.Cout, s.CoutS, 0, false));

  // packed weights
  std::vector<bf16> wf((size_t)packed_wf_elems(s), __float2bfloat16_rn(0.f));
  std::vector<bf16> wt((size_t)packed_wt_elems(s), __float2bfloat16_rn(0.f));
  for (int co = 0; co < s.Cout; ++co)
    for (int t = 0; t < T; ++t)
      for (int ci = 0; ci < s.Cin; ++ci) {
        const bf16 v = __float2bfloat16_rn(h.w[((size_t)co * T + t) * s.Cin + ci]);
        wf[((size_t)co * T + t) * s.CinS + ci] = v;
        wt[((size_t)ci * T + t) * s.CoutS + co] = v;
      }
  bf16* d_wf = dev_upload(wf);
  bf16* d_wt = dev_upload(wt);
  float* d_bias = dev_upload(h.b);

  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);

  if (passes & 1) {
    IgemmPlan p = plan_fprop(s, x, d_wf, y, d_bias, act, sm_count);
    (void)force_bn;
    p.args.kiters = dev_upload(p.kiters);
    printf("  fprop: BN=%d BK=%d cluster %dx%d tiles=%d nblk=%d classes=%d kiters=%zu patch=%d MT=%d CG=%d ctas_m=%d bstages=%d lite=%d tapn=%d\n", p.BN,
           p.BK, p.CM, p.CN, p.num_tiles, p.n_blocks, p.n_classes, p.kiters.size(), (int)p.patch, p.MT, p.CG, p.num_ctas_m,
           p.patch ? p.pargs.b_stages : 0, (int)p.lite, (int)p.tapn);
    run(p, 0);
    CGB_CUDA(cudaDeviceSynchronize());
    std::vector<bf16> yb((size_t)y.elems());
    CGB_CUDA(cudaMemcpy(yb.data(), y.ptr, yb.size() * sizeof(bf16), cudaMemcpyDeviceToHost));
    std::vector<float> ref, got;
    if (!timing_only) ref_fprop(h, ref);
    else ref.assign((size_t)N * h.Ho * h.Wo * s.Cout, 0.f);
    for (auto& v : ref) {
      if (act == kActLeaky) v = v > 0 ? v : 0.2f * v;
      if (act == kActTanh) v = std::tanh(v);
    }
    got.resize(ref.size());
    bool pad_zero = true;
    for (size_t px = 0; px < (size_t)N * h.Ho * h.Wo; ++px)
      for (int c = 0; c < s.CoutS; ++c) {
        const float v = __bfloat162float(yb[px * s.CoutS + c]);
        if (c < s.Cout)
          got[px * s.Cout + c] = v;
        else if (v != 0.f)
          pad_zero = false;
      }
    if (!timing_only && (compare("fprop", ref, got) >= 1e-2 || !pad_zero)) ++fails;
    if (!pad_zero) printf("  fprop: padded channels are not zero -> FAIL\n");
    // timing
    for (int i = 0; i < 3; ++i) run(p, 0);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) run(p, 0);
    cudaEventRecord(e1);
    CGB_CUDA(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("  fprop: %.2f us/launch, %.1f TFLOP/s (algorithmic)\n", ms / reps * 1e3, p.flops / (ms / reps * 1e-3) * 1e-12);
    if (getenv("CGB_PROF")) {
      const int ctas = p.num_tiles * p.n_blocks * p.n_classes;
      long long* d_prof;
      CGB_CUDA(cudaMalloc(&d_prof, (size_t)ctas * 16 * sizeof(long long)));
      CGB_CUDA(cudaMemset(d_prof, 0, (size_t)ctas * 16 * sizeof(long long)));
      p.args.prof = d_prof;
      run(p, 0);
      CGB_CUDA(cudaDeviceSynchronize());
      std::vector<long long> hp((size_t)ctas * 16);
      CGB_CUDA(cudaMemcpy(hp.data(), d_prof, hp.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      double acc[16] = {0};
      for (int c = 0; c < ctas; ++c)
        for (int j = 1; j < 16; ++j) acc[j] += (double)(hp[c * 16 + j] - hp[c * 16]);
      printf("  fprop phases (mean cycles from CTA start over %d CTAs): setup %.0f | first TMA issued %.0f | producer done %.0f | "
             "MMA issue done %.0f | accumulator ready %.0f | epilogue done %.0f | teardown %.0f\n",
             ctas, acc[1] / ctas, acc[2] / ctas, acc[3] / ctas, acc[4] / ctas, acc[5] / ctas, acc[6] / ctas, acc[7] / ctas);
      printf("  epilogue detail: first tmem.ld done %.0f | first chunk math done %.0f | all chunks staged %.0f\n", acc[8] / ctas,
             acc[9] / ctas, acc[10] / ctas);
      p.args.prof = nullptr;
    }
  }
  if (passes & 2) {
    int DH, DW;
    std::vector<float> ref;
    if (!timing_only) ref_dgrad(h, ref, &DH, &DW);
    else {
      DH = H + 2 * (s.reflect ? s.pad : 0);
      DW = W + 2 * (s.reflect ? s.pad : 0);
      ref.assign((size_t)N * DH * DW * s.Cin, 0.f);
    }
    TensorDesc dx{nullptr, N, DH, DW, s.CinS, 0};
    CGB_CUDA(cudaMalloc(&dx.ptr, dx.elems() * sizeof(bf16)));
    CGB_CUDA(cudaMemset(dx.ptr, 0xFF, dx.elems() * sizeof(bf16)));
    IgemmPlan p = plan_dgrad(s, dy, d_wt, dx, sm_count);
    p.args.kiters = dev_upload(p.kiters);
    printf("  dgrad: BN=%d BK=%d cluster %dx%d tiles=%d nblk=%d classes=%d kiters=%zu out %dx%d patch=%d MT=%d CG=%d ctas_m=%d bstages=%d lite=%d tapn=%d\n",
           p.BN, p.BK, p.CM, p.CN, p.num_tiles, p.n_blocks, p.n_classes, p.kiters.size(), DH, DW, (int)p.patch, p.MT, p.CG,
           p.num_ctas_m, p.patch ? p.pargs.b_stages : 0, (int)p.lite, (int)p.tapn);
    run(p, 0);
    CGB_CUDA(cudaDeviceSynchronize());
    std::vector<bf16> db((size_t)dx.elems());
    CGB_CUDA(cudaMemcpy(db.data(), dx.ptr, db.size() * sizeof(bf16), cudaMemcpyDeviceToHost));
    std::vector<float> got(ref.size());
    for (size_t px = 0; px < (size_t)N * DH * DW; ++px)
      for (int c = 0; c < s.Cin; ++c) got[px * s.Cin + c] = __bfloat162float(db[px * s.CinS + c]);
    if (!timing_only && compare("dgrad", ref, got) >= 1e-2) ++fails;
    for (int i = 0; i < 3; ++i) run(p, 0);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) run(p, 0);
    cudaEventRecord(e1);
    CGB_CUDA(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("  dgrad: %.2f us/launch, %.1f TFLOP/s (algorithmic)\n", ms / reps * 1e3, p.flops / (ms / reps * 1e-3) * 1e-12);
  }
  if (passes & 4) {
    std::vector<float> ref;
    if (!timing_only) ref_wgrad(h, ref);
    else ref.assign((size_t)s.Cout * T * s.Cin, 0.f);
    float* d_g = nullptr;
    CGB_CUDA(cudaMalloc(&d_g, ref.size() * sizeof(float)));
    CGB_CUDA(cudaMemset(d_g, 0, ref.size() * sizeof(float)));
    WgradPlan p = plan_wgrad(s, x, dy, d_g, sm_count);
    p.args.taps = dev_upload(p.taps);
    printf("  wgrad: BNW=%d m_blocks=%d taps=%d split_k=%d chunks=%dx%dx%d pair=%d units=%d\n", p.BNW, p.m_blocks,
           p.args.num_taps, p.args.split_k, p.args.N, p.args.tiles_h, p.args.tiles_w, (int)p.pair, p.pair ? p.pargs.n_units : 0);
    run(p, 0);
    CGB_CUDA(cudaDeviceSynchronize());
    std::vector<float> got(ref.size());
    CGB_CUDA(cudaMemcpy(got.data(), d_g, got.size() * sizeof(float), cudaMemcpyDeviceToHost));
    if (!timing_only && compare("wgrad", ref, got) >= 1e-2) ++fails;
    for (int i = 0; i < 3; ++i) run(p, 0);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) run(p, 0);
    cudaEventRecord(e1);
    CGB_CUDA(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("  wgrad: %.2f us/launch, %.1f TFLOP/s (algorithmic)\n", ms / reps * 1e3, p.flops / (ms / reps * 1e-3) * 1e-12);
  }
  printf("case %s: %s\n", name.c_str(), fails ? "FAILED" : "PASSED");
  fflush(stdout);
  return fails;
}

static ConvSpec spec(int Cin, int Cout, int k, int stride, int pad, bool reflect, bool transposed) {
  ConvSpec s;
  s.Cin = Cin;
  s.Cout = Cout;
  s.CinS = Cin % 64 == 0 ? Cin : 16;
  s.CoutS = Cout % 64 == 0 ? Cout : 16;
  s.k = k;
  s.stride = stride;
  s.pad = pad;
  s.reflect = reflect;
  s.transposed = transposed;
  return s;
}

int main(int argc, char** argv) {
  const std::string name = argc > 1 ? argv[1] : "res_small";
  const int N = argc > 2 ? atoi(argv[2]) : 1;
  const int force_bn = argc > 3 ? atoi(argv[3]) : 0;
  const int Hov = argc > 4 ? atoi(argv[4]) : 0;  // spatial extent override (0: the case's default)
  auto HH = [&](int d) { return Hov ? Hov : d; };
  try {
    if (name == "res_small") return run_case(name, spec(64, 64, 3, 1, 1, true, false), N, HH(16), HH(16), 0, 7, force_bn);
    if (name == "res") return run_case(name, spec(256, 256, 3, 1, 1, true, false), N, HH(64), HH(64), 0, 7, force_bn);
    if (name == "dconv3") return run_case(name, spec(256, 512, 4, 1, 1, false, false), N, HH(32), HH(32), 0, 7, force_bn);
    if (name == "down") return run_case(name, spec(64, 128, 3, 2, 1, false, false), N, HH(32), HH(32), 0, 7, force_bn);
    if (name == "down2") return run_case(name, spec(128, 256, 3, 2, 1, false, false), N, HH(128), HH(128), 0, 7, force_bn);
    if (name == "dconv1") return run_case(name, spec(64, 128, 4, 2, 1, false, false), N, HH(32), HH(32), 0, 7, force_bn);
    if (name == "up") return run_case(name, spec(256, 128, 3, 2, 1, false, true), N, HH(16), HH(16), 0, 7, force_bn);
    if (name == "up2") return run_case(name, spec(128, 64, 3, 2, 1, false, true), N, HH(32), HH(32), 0, 7, force_bn);
    if (name == "stem") return run_case(name, spec(3, 64, 7, 1, 3, true, false), N, HH(32), HH(32), 0, 3, force_bn);
    if (name == "head") return run_case(name, spec(64, 3, 7, 1, 3, true, false), N, HH(32), HH(32), kActTanh, 3, force_bn);
    if (name == "dconv0") return run_case(name, spec(3, 64, 4, 2, 1, false, false), N, HH(32), HH(32), kActLeaky, 3, force_bn);
    if (name == "dconv4") return run_case(name, spec(512, 1, 4, 1, 1, false, false), N, HH(31), HH(31), 0, 3, force_bn);
    printf("unknown case %s\n", name.c_str());
    return 2;
  } catch (const std::exception& e) {
    printf("case %s: EXCEPTION %s\n", name.c_str(), e.what());
    return 3;
  }
}
