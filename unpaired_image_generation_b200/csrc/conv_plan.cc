// Host-side planning of the implicit-GEMM convolutions (see conv_plan.h).
#include "conv_plan.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace cgb {

bool pdl_enabled() {
  static const bool on = !(std::getenv("CGB_PDL") && std::atoi(std::getenv("CGB_PDL")) == 0);
  return on;
}

// The CTA-pair (2-CTA cluster) kernels are launched as ORDINARY graph nodes, without the programmatic-dependent-launch
// attribute.  With it the training step hung about once per 20 000 replays at batch 1 (scripts/gpu_stress.py: stalls at
// replays 3 949, 14 849, 15 999, 30 699 and 38 849 of five runs; 70 000 replays without a stall once the attribute was
// dropped, 80 000 on the commit before the pair kernels existed).  No bounded mbarrier wait tripped, so the stall is in
// the launch machinery (cluster kernel as primary / dependent of a programmatic edge), not in the kernels' protocols;
// the root cause was not isolated.  CGB_PAIR_PDL=1 restores the attribute for experiments.
bool pair_pdl_enabled() {
  static const bool on = std::getenv("CGB_PAIR_PDL") ? std::atoi(std::getenv("CGB_PAIR_PDL")) != 0 : false;
  return on && pdl_enabled();
}

int tmem_exclusive_smem() {
  static const bool on = !(std::getenv("CGB_EXCL_SMEM") && std::atoi(std::getenv("CGB_EXCL_SMEM")) == 0);
  return on ? 184 * 1024 : 0;  // 184 KB + 54 KB (+ 1 KB reserved per CTA) > 227 KB
}

int padded_rows(int c) { return c <= 16 ? 16 : (c + 63) / 64 * 64; }
long long packed_wf_elems(const ConvSpec& s) { return (long long)padded_rows(s.CoutS) * s.taps() * s.CinS; }
long long packed_wt_elems(const ConvSpec& s) { return (long long)padded_rows(s.CinS) * s.taps() * s.CoutS; }

int out_extent(const ConvSpec& s, int in) {
  if (s.transposed) return in * 2;
  return (in + 2 * s.pad - s.k) / s.stride + 1;
}

static bool chan_ok(int c) { return c == 16 || (c > 0 && c % 64 == 0); }
bool tc_supports_fprop(const ConvSpec& s) { return chan_ok(s.CinS) && chan_ok(s.CoutS); }
bool tc_supports_dgrad(const ConvSpec& s) { return chan_ok(s.CinS) && chan_ok(s.CoutS); }
bool tc_supports_wgrad(const ConvSpec& s) { return s.CinS % 64 == 0 && s.CoutS % 64 == 0; }

static int ilog2_ceil(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// floor-division split of a tap offset q into (parity, half-resolution offset): q = 2*off + par
static void split_parity(int q, int* par, int* off) {
  *par = ((q % 2) + 2) % 2;
  *off = (q - *par) / 2;
}

static CUtensorMap view_s1(const TensorDesc& t, bool padded_view, int box_c, int box_w, int box_h) {
  const int h = padded_view ? t.halo : 0;
  const int dims[5] = {t.C, t.W + 2 * h, 1, t.H + 2 * h, t.N};
  const long long str[4] = {t.sW(), t.sH(), t.sH(), t.sN()};
  return make_tmap_act5d(padded_view ? t.ptr : t.interior(), dims, str, box_c, box_w, box_h, box_c * 2);
}

static CUtensorMap view_s2(const TensorDesc& t, int box_c, int box_w, int box_h) {
  CGB_CHECK(t.H % 2 == 0 && t.W % 2 == 0, "stride-2 parity view needs even extents");
  const int dims[5] = {2 * t.C, t.W / 2, 2, t.H / 2, t.N};
  const long long str[4] = {2 * t.sW(), t.sH(), 2 * t.sH(), t.sN()};
  return make_tmap_act5d(t.interior(), dims, str, box_c, box_w, box_h, box_c * 2);
}

static int choose_bn(int rows, int bk, long long ctas_per_nblock, int sm_count) {
  if (rows <= 16) return 16;
  if (bk == 16) return 64;
  if (const char* f = std::getenv("CGB_FORCE_BN")) {
    const int bn = std::atoi(f);
    if ((bn == 64 || bn == 128 || bn == 256) && rows % bn == 0) return bn;
  }
  // The widest N tile that still yields `target` CTAs.  Wide tiles cost the least SM time per FLOP (the MMA
  // issue loop is paid per K block regardless of N), and the step graph keeps several independent convs in
  // flight, so the target is a fraction of the machine rather than all of it.
  static const double frac = std::getenv("CGB_CTA_FRAC") ? std::atof(std::getenv("CGB_CTA_FRAC")) : 0.3;
  const long long target = (long long)(sm_count * frac);
  const int cands[3] = {256, 128, 64};
  for (int bn : cands) {
    if (rows % bn != 0) continue;
    if (ctas_per_nblock * (rows / bn) >= target) return bn;
  }
  return 64;
}

// Cluster shape for a (BN, BK, n_blocks, tiles) problem; must match an instantiation in conv_tc.cu.
static void choose_cluster(int BN, int BK, int n_blocks, int num_tiles, int* CM, int* CN) {
  *CM = 1;
  *CN = 1;
  // Measured on B200 (profiles/r01_c_cluster_multicast.txt): with <= 8-CTA clusters TMA multicast does not
  // reduce L2 -> SM time (the microarchitecture notes say the same: multicast ~ unicast at cluster size <= 4),
  // and the cross-CTA slot release couples the pipelines: res-block conv 18.6 us -> 33.5 us at batch 1,
  // 53.6 -> 49.4 us at batch 8.  Kept behind CGB_CLUSTER=1 for experiments.
  static const bool enabled = std::getenv("CGB_CLUSTER") != nullptr;
  if (!enabled || BK != 64 || BN < 64) return;
  int cn = 1;
  if (BN == 64) cn = n_blocks % 4 == 0 ? 4 : (n_blocks % 2 == 0 ? 2 : 1);
  else if (BN == 128) cn = n_blocks % 2 == 0 ? 2 : 1;
  int cm = cn == 4 ? 2 : 4;
  if (num_tiles < 2 * cm) return;
  // instantiated: BN 64: 2x4, 4x2, 4x1; BN 128: 4x2, 4x1; BN 256: 4x1
  if (BN == 256) cn = 1;
  *CM = cm;
  *CN = cn;
}

enum ViewKind { kViewS1Interior, kViewS1Padded, kViewS2 };

// choose the N tile and the cluster, then build both tensor maps with per-CTA slice boxes
static void finalize(IgemmPlan& p, ViewKind vk, const TensorDesc& act, const bf16* w, int rows, int Kw, int sm_count) {
  p.BN = choose_bn(rows, p.BK, (long long)p.num_tiles * p.n_classes, sm_count);
  p.n_blocks = (padded_rows(rows) + p.BN - 1) / p.BN;
  choose_cluster(p.BN, p.BK, p.n_blocks, p.num_tiles, &p.CM, &p.CN);
  const int TW = 1 << p.args.tw_shift, TH = 128 >> p.args.tw_shift;
  // A slice of a CTA: 128 / CN consecutive tile pixels = SH rows x SW columns
  const int SH = TH >= p.CN ? TH / p.CN : 1;
  const int SW = TH >= p.CN ? TW : TW * TH / p.CN;
  CGB_CHECK(SH * SW * p.CN == 128 && SW >= 8, "cluster slice does not tile the 128-pixel M tile");
  if (vk == kViewS2) p.tmA = view_s2(act, p.BK, SW, SH);
  else p.tmA = view_s1(act, vk == kViewS1Padded, p.BK, SW, SH);
  p.tmB = make_tmap_2d(w, padded_rows(rows), Kw, Kw, p.BK, p.BN / p.CM, p.BK * 2);
  // many short CTAs (the stride-2 / transposed layers on large maps: 2-9 K iterations each, several waves): trade the
  // deep ring for 2-3 co-resident CTAs per SM.  CGB_IGEMM_LITE=0 disables, =N sets the minimum number of waves (-1: always),
  // CGB_IGEMM_LITE_MAXK the longest K loop (iterations of 64 channels) that still qualifies.
  static const int lite_waves = std::getenv("CGB_IGEMM_LITE") ? std::atoi(std::getenv("CGB_IGEMM_LITE")) : 2;
  static const int lite_maxk = std::getenv("CGB_IGEMM_LITE_MAXK") ? std::atoi(std::getenv("CGB_IGEMM_LITE_MAXK")) : 16;
  int max_k = 0;
  for (int z = 0; z < p.n_classes; ++z) max_k = std::max(max_k, p.args.k_count[z]);
  const long long ctas = (long long)p.num_tiles * p.n_blocks * p.n_classes;
  p.lite = lite_waves != 0 && p.BK == 64 && p.BN >= 64 && p.CM * p.CN == 1 &&
           (lite_waves < 0 || (max_k <= lite_maxk && ctas >= (long long)lite_waves * sm_count));
}

static void set_tiles(IgemmPlan& p, int N, int Ho, int Wo) {
  p.args.N = N;
  // tile = TH x TW output pixels with TH * TW = 128: pick the shape that wastes the fewest pixels
  // (e.g. the 66x66 padded-domain dgrads: 8x16 tiles -> 45 tiles instead of 66 of 1x128)
  int shift = 7;
  long long best = -1;
  for (int sh = 7; sh >= 3; --sh) {
    const int tw = 1 << sh, th = 128 >> sh;
    const long long tiles = (long long)((Wo + tw - 1) / tw) * ((Ho + th - 1) / th);
    if (best < 0 || tiles < best) {
      best = tiles;
      shift = sh;
    }
  }
  const int TW = 1 << shift, TH = 128 >> shift;
  p.args.tw_shift = shift;
  p.args.tiles_w = (Wo + TW - 1) / TW;
  p.args.tiles_h = (Ho + TH - 1) / TH;
  p.args.Ho = Ho;
  p.args.Wo = Wo;
  p.num_tiles = N * p.args.tiles_w * p.args.tiles_h;
}

// Patch-resident kernel (conv_patch.cu) for a stride-1 conv-like pass over `act` (64-channel chunks):
// output extent Ho x Wo, patch origin (ox, oy) relative to the output pixel, `flip` walks the filter backwards.
// Returns false (plan untouched) when the geometry does not fit; the tap-table kernel is used instead.
static int patch_mode() {
  static const int mode = std::getenv("CGB_PATCH_MODE") ? std::atoi(std::getenv("CGB_PATCH_MODE")) : 1;
  return mode;
}

static bool try_patch(IgemmPlan& p, const TensorDesc& act, bool padded_view, const bf16* w, int rows, int Kw, int k,
                      int chans, int Ho, int Wo, int ox, int oy, bool flip, int sm_count) {
  const int mode = patch_mode();
  // patch rows carry 64 channels (128 bytes) or, for the 16-stored-channel image-like tensors, 16 channels (32 bytes)
  const int ka = chans % 64 == 0 ? 64 : 16;
  static const bool no16 = std::getenv("CGB_PATCH_NO16") != nullptr;
  if (mode <= 0 || k < 2 || k > 8 || (ka == 16 && (chans != 16 || mode != 1 || no16 || padded_rows(rows) < 64))) return false;
  const int rb = ka * 2;  // bytes per patch row
  PatchArgs pa;
  std::memset(&pa, 0, sizeof(pa));
  pa.k = k;
  pa.ka = ka;
  pa.flip = flip ? 1 : 0;
  pa.ox = ox;
  pa.oy = oy;
  pa.chunks = chans / ka;
  pa.tap_stride = chans;
  pa.PH = 16 + k - 1;
  int box_w;
  if (mode == 3) {         // k column-shifted 8-wide boxes: every descriptor start is 1024-byte aligned
    box_w = 8;
    pa.nbox = k;
    pa.box_bytes = pa.PH * 8 * 128;
    pa.patch_bytes = k * pa.box_bytes;
    pa.row_step = 64;
    pa.col_step = pa.PH * 64;
    pa.sbo = 1024;
  } else {                 // one box; taps start at arbitrary rows of it
    box_w = (mode == 5) ? 16 : 8 + k - 1;  // mode 5: 2048-byte pitch, so SBO is a multiple of the 1024-byte swizzle atom
    pa.nbox = 1;
    pa.box_bytes = pa.PH * box_w * rb;
    pa.patch_bytes = (pa.box_bytes + 1023) / 1024 * 1024;
    pa.row_step = box_w * rb / 16;
    pa.col_step = rb / 16;
    pa.sbo = box_w * rb;
    pa.base_offset = 0;  // measured on B200: the swizzle XOR uses absolute smem address bits; the base-offset field must stay 0
  }
  const int tiles_w = (Wo + 7) / 8, tile_rows = (Ho + 15) / 16;
  const int prow = padded_rows(rows);
  static const double frac_f = std::getenv("CGB_CTA_FRAC") ? std::atof(std::getenv("CGB_CTA_FRAC")) : 0.3;
  static const double frac_d = std::getenv("CGB_CTA_FRAC_DGRAD") ? std::atof(std::getenv("CGB_CTA_FRAC_DGRAD")) : frac_f;
  const long long target = (long long)(sm_count * (flip ? frac_d : frac_f));
  int forced_bn = 0, forced_mt = 0;
  if (const char* f = std::getenv("CGB_FORCE_BN")) forced_bn = std::atoi(f);
  if (const char* f = std::getenv("CGB_FORCE_MT")) forced_mt = std::atoi(f);
  const int bn_cands[3] = {256, 128, 64};
  int best_bn = 0, best_mt = 0, best_stages = 0;
  const int T = k * k;
  bool resident = false;
  // CTA pairs (cta_group::2, conv_patch.cu): two M tiles share one weight tile, each CTA stages half of its rows.
  // CGB_PATCH_CG=0 keeps every layer on single CTAs.
  static const int cg_mode = std::getenv("CGB_PATCH_CG") ? std::atoi(std::getenv("CGB_PATCH_CG")) : 1;
  auto pair_ok = [&](int bn, int mt) {
    return cg_mode != 0 && ka == 64 && ((bn == 256 && mt == 1) || bn == 128);
  };
  auto fits = [&](int bn, int mt, int* stages) {
    const bool pr = pair_ok(bn, mt);
    const int kps = pr ? 3 : igemm_patch_kps(bn, ka);
    const int stage = kps * (((pr ? bn / 2 : bn) * 128 + 1023) / 1024 * 1024);
    const int budget = igemm_patch_smem_budget() - 2 * mt * pa.patch_bytes;
    int st = budget / stage;
    if (st > 16) st = 16;
    if (st < 2) return false;
    // the whole filter fits the ring: load it once per CTA and keep it (persistent CTAs sweep many work items)
    const int boxes = ka == 64 ? T : (T * ka + 63) / 64;
    const int per_item = (boxes + kps - 1) / kps * pa.chunks;
    resident = per_item >= 2 && per_item <= st;
    if (resident) st = per_item;
    *stages = st;
    return true;
  };
  if (prow <= 16) {
    for (int mt : {2, 1}) {
      int st;
      if (forced_mt && mt != forced_mt) continue;
      if (!fits(16, mt, &st)) continue;
      const long long ctas = (long long)act.N * tiles_w * ((tile_rows + mt - 1) / mt);
      best_bn = 16, best_mt = mt, best_stages = st;
      // short per-tile work: only stack tiles when there are plenty of them (measured: stacking beats a resident
      // filter on the 64 -> 3 head, 143 vs 161 us at batch 8: N = 16 MMAs are bound by the A-operand reads)
      if (ctas >= 4 * target) break;
    }
  } else {
    bool done = false;
    for (int bn : bn_cands) {
      if (done || prow % bn != 0 || (forced_bn && bn != forced_bn)) continue;
      for (int mt : {2, 1}) {
        int st;
        if (done || (forced_mt && mt != forced_mt)) continue;
        // stacking wastes a whole tile when the tile-row count is odd (the 66-row padded-domain gradients: 5 rows)
        if (mt == 2 && !forced_mt && tile_rows % 2 != 0 && tile_rows < 9) continue;
        // BN = 256: two stacked tiles would take all 512 TMEM columns and leave no second accumulator set for the
        // epilogue / main-loop overlap of the persistent kernel
        if (mt == 2 && !forced_mt && bn == 256) continue;
        if (!fits(bn, mt, &st)) continue;
        const long long ctas = (long long)act.N * tiles_w * ((tile_rows + mt - 1) / mt) * (prow / bn);
        best_bn = bn, best_mt = mt, best_stages = st;  // the last (smallest) candidate wins if none reaches the target
        if (ctas >= target) done = true;
      }
    }
  }
  if (best_bn == 0) return false;
  pa.b_stages = best_stages;
  {
    int st;
    fits(best_bn, best_mt, &st);  // re-evaluate `resident` for the chosen shape
    pa.b_resident = resident ? 1 : 0;
  }
  p.patch = true;
  p.pargs = pa;
  p.BN = best_bn;
  p.MT = best_mt;
  p.CG = pair_ok(best_bn, best_mt) ? 2 : 1;
  p.CM = p.CN = 1;
  p.n_blocks = (prow + p.BN - 1) / p.BN;
  p.n_classes = 1;
  p.args.tw_shift = 3;
  p.args.tiles_w = tiles_w;
  p.args.tiles_h = (tile_rows + p.MT - 1) / p.MT;  // CTA rows
  p.args.Ho = Ho;
  p.args.Wo = Wo;
  // persistent CTAs: at most one per SM (per N block); each loops over its share of the work items, overlapping
  // the epilogue of one item with the main loop of the next (two TMEM accumulator sets)
  pa.num_items = act.N * tiles_w * p.args.tiles_h;
  p.pargs.num_items = pa.num_items;
  if (p.CG == 2)  // pairs walk units of two items; one pair per TPC (two SMs)
    p.num_ctas_m = 2 * std::min((pa.num_items + 1) / 2, std::max(1, sm_count / p.n_blocks / 2));
  else
    p.num_ctas_m = std::min(pa.num_items, std::max(1, sm_count / p.n_blocks));
  p.num_tiles = pa.num_items;
  p.tmA = view_s1(act, padded_view, ka, box_w, pa.PH);
  p.tmB = make_tmap_2d(w, prow, Kw, Kw, 64, p.BN / p.CG, 128);
  return true;
}

// Taps-in-N kernel (conv_tapn.cu) for a stride-1 k x k conv-like pass from 64 stored channels to <= 4 real ones
// (16 stored): the 7 x 7 generator head and the input gradient of the 7 x 7 stem.  CGB_TAPN=0 disables.
static bool try_tapn(IgemmPlan& p, const TensorDesc& act, bool padded_view, const bf16* w, long long Kw, int k, int cout,
                     const TensorDesc& out, int Ho, int Wo, int ox, int oy, bool flip, const float* bias, int bias_n,
                     int actfn, int sm_count) {
  static const bool on = !(std::getenv("CGB_TAPN") && std::atoi(std::getenv("CGB_TAPN")) == 0);
  if (!on || k < 5 || k > 8 || act.C != 64 || out.C != 16 || cout > 4 || (actfn != kActNone && actfn != kActTanh)) return false;
  TapNArgs a;
  std::memset(&a, 0, sizeof(a));
  a.k = k;
  a.flip = flip ? 1 : 0;
  a.ox = ox;
  a.oy = oy;
  a.tiles_w = (Wo + 7) / 8;
  a.tiles_h = (Ho + 15) / 16;
  a.N = act.N;
  a.num_items = a.N * a.tiles_w * a.tiles_h;
  a.Ho = Ho;
  a.Wo = Wo;
  a.cout = cout;
  a.sN = out.sN();
  a.sH = out.sH();
  a.sW = out.sW();
  a.out = out.interior();
  a.bias = bias;
  a.bias_n = bias_n;
  a.act = actfn;
  p.tapn = true;
  p.targs = a;
  p.patch = false;
  p.BN = 32;
  p.MT = 1;
  p.n_blocks = 1;
  p.n_classes = 1;
  p.num_tiles = a.num_items;
  p.num_ctas_m = std::min(a.num_items, sm_count);
  p.tmA = view_s1(act, padded_view, 64, 16, 16 + k - 1);
  p.tmB = make_tmap_tapn_weights(w, k, Kw);
  return true;
}

IgemmPlan plan_fprop(const ConvSpec& s, const TensorDesc& x, const bf16* wf, const TensorDesc& y, const float* bias,
                     int act, int sm_count) {
  CGB_CHECK(tc_supports_fprop(s), "fprop: channel counts not supported by the tensor-core path");
  CGB_CHECK(x.C == s.CinS && y.C == s.CoutS, "fprop: tensor channels do not match the spec");
  IgemmPlan p;
  std::memset(&p.args, 0, sizeof(p.args));
  const int k = s.k, T = s.taps();
  const int BK = (s.CinS % 64 == 0) ? 64 : 16;
  p.BK = BK;
  const int Ho = out_extent(s, x.H), Wo = out_extent(s, x.W);
  CGB_CHECK(y.H == Ho && y.W == Wo && y.N == x.N, "fprop: output extent mismatch");
  const int Kw = T * s.CinS;  // packed weight row length
  p.args.out = y.interior();
  p.args.sN = y.sN();
  p.args.bias = bias;
  p.args.bias_n = bias ? s.Cout : 0;
  p.args.act = act;
  p.args.Cout = s.CoutS;
  p.args.stats = nullptr;
  ViewKind vk = kViewS1Interior;

  if (!s.transposed) {
    set_tiles(p, x.N, Ho, Wo);
    p.n_classes = 1;
    p.args.sH = y.sH();
    p.args.sW = y.sW();
    p.args.out_off[0] = 0;
    p.args.k_begin[0] = 0;
    if (s.stride == 1) {
      if (s.reflect) CGB_CHECK(x.halo == s.pad, "reflect conv input must carry a halo equal to the padding");
      vk = s.reflect ? kViewS1Padded : kViewS1Interior;
      for (int r = 0; r < k; ++r)
        for (int c = 0; c < k; ++c)
          for (int c0 = 0; c0 < s.CinS; c0 += BK) {
            KIter it{};
            it.a_c = (int16_t)c0;
            it.a_dx = (int16_t)(s.reflect ? c : c - s.pad);
            it.a_par = 0;
            it.a_dy = (int16_t)(s.reflect ? r : r - s.pad);
            it.b_k = (r * k + c) * s.CinS + c0;
            p.kiters.push_back(it);
          }
    } else {
      CGB_CHECK(s.stride == 2 && !s.reflect, "only zero-padded stride-2 convs are supported");
      vk = kViewS2;
      for (int r = 0; r < k; ++r)
        for (int c = 0; c < k; ++c) {
          int hp, hy, wp, wx;
          split_parity(r - s.pad, &hp, &hy);
          split_parity(c - s.pad, &wp, &wx);
          for (int c0 = 0; c0 < s.CinS; c0 += BK) {
            KIter it{};
            it.a_c = (int16_t)(wp * s.CinS + c0);
            it.a_dx = (int16_t)wx;
            it.a_par = (int16_t)hp;
            it.a_dy = (int16_t)hy;
            it.b_k = (r * k + c) * s.CinS + c0;
            p.kiters.push_back(it);
          }
        }
    }
    p.args.k_count[0] = (int)p.kiters.size();
  } else {
    CGB_CHECK(s.k == 3 && s.stride == 2 && s.pad == 1, "transposed conv: only k=3, s=2, p=1, op=1");
    // four output-parity classes, each a small stride-1 conv over the input
    set_tiles(p, x.N, x.H, x.W);
    p.n_classes = 4;
    p.args.sH = 2 * y.sH();
    p.args.sW = 2 * y.sW();
    vk = kViewS1Interior;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        const int z = a * 2 + b;
        p.args.out_off[z] = a * y.sH() + b * y.sW();
        p.args.k_begin[z] = (int)p.kiters.size();
        for (int r = 0; r < k; ++r) {
          if ((a + s.pad - r) % 2 != 0) continue;
          for (int c = 0; c < k; ++c) {
            if ((b + s.pad - c) % 2 != 0) continue;
            for (int c0 = 0; c0 < s.CinS; c0 += BK) {
              KIter it{};
              it.a_c = (int16_t)c0;
              it.a_dx = (int16_t)((b + s.pad - c) / 2);
              it.a_par = 0;
              it.a_dy = (int16_t)((a + s.pad - r) / 2);
              it.b_k = (r * k + c) * s.CinS + c0;
              p.kiters.push_back(it);
            }
          }
        }
        p.args.k_count[z] = (int)p.kiters.size() - p.args.k_begin[z];
      }
  }
  bool patched = false;
  if (!s.transposed && s.stride == 1)
    patched = try_tapn(p, x, s.reflect, wf, Kw, k, s.Cout, y, Ho, Wo, s.reflect ? 0 : -s.pad, s.reflect ? 0 : -s.pad, false,
                       bias, bias ? s.Cout : 0, act, sm_count);
  if (!patched && !s.transposed && s.stride == 1)
    patched = try_patch(p, x, s.reflect, wf, s.CoutS, Kw, k, s.CinS, Ho, Wo, s.reflect ? 0 : -s.pad,
                        s.reflect ? 0 : -s.pad, false, sm_count);
  if (!patched) finalize(p, vk, x, wf, s.CoutS, Kw, sm_count);
  p.flops = 2.0 * x.N * (s.transposed ? (double)x.H * x.W : (double)Ho * Wo) * s.Cout * s.Cin * T;
  return p;
}

IgemmPlan plan_dgrad(const ConvSpec& s, const TensorDesc& dy, const bf16* wt, const TensorDesc& dx, int sm_count) {
  CGB_CHECK(tc_supports_dgrad(s), "dgrad: channel counts not supported by the tensor-core path");
  CGB_CHECK(dy.C == s.CoutS && dx.C == s.CinS, "dgrad: tensor channels do not match the spec");
  IgemmPlan p;
  std::memset(&p.args, 0, sizeof(p.args));
  const int k = s.k, T = s.taps();
  const int BK = (s.CoutS % 64 == 0) ? 64 : 16;
  p.BK = BK;
  const int Kw = T * s.CoutS;
  p.args.out = dx.interior();
  p.args.sN = dx.sN();
  p.args.bias = nullptr;
  p.args.bias_n = 0;
  p.args.act = kActNone;
  p.args.Cout = s.CinS;
  p.args.stats = nullptr;
  ViewKind vk = kViewS1Interior;

  if (!s.transposed && s.stride == 1) {
    // zero pad:   dx[h]   = sum_r dy[h + p - r] w[r]
    // reflect:    dxp[hp] = sum_r dy[hp - r]    w[r]   on the padded domain (dx.H == H + 2p)
    const int off = s.reflect ? 0 : s.pad;
    set_tiles(p, dx.N, dx.H, dx.W);
    p.n_classes = 1;
    p.args.sH = dx.sH();
    p.args.sW = dx.sW();
    for (int r = 0; r < k; ++r)
      for (int c = 0; c < k; ++c)
        for (int c0 = 0; c0 < s.CoutS; c0 += BK) {
          KIter it{};
          it.a_c = (int16_t)c0;
          it.a_dx = (int16_t)(off - c);
          it.a_par = 0;
          it.a_dy = (int16_t)(off - r);
          it.b_k = (r * k + c) * s.CoutS + c0;
          p.kiters.push_back(it);
        }
    p.args.k_begin[0] = 0;
    p.args.k_count[0] = (int)p.kiters.size();
  } else if (!s.transposed) {
    // stride 2: ih = 2*oh + r - p.  Four input-parity classes (a, b); oh = i + (a + p - r)/2.
    CGB_CHECK(s.stride == 2 && !s.reflect, "dgrad: only zero-padded stride-2 convs are supported");
    CGB_CHECK(dx.H % 2 == 0 && dx.W % 2 == 0, "dgrad stride 2: input extents must be even");
    set_tiles(p, dx.N, dx.H / 2, dx.W / 2);
    p.n_classes = 4;
    p.args.sH = 2 * dx.sH();
    p.args.sW = 2 * dx.sW();
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        const int z = a * 2 + b;
        p.args.out_off[z] = a * dx.sH() + b * dx.sW();
        p.args.k_begin[z] = (int)p.kiters.size();
        for (int r = 0; r < k; ++r) {
          if (((a + s.pad - r) % 2 + 2) % 2 != 0) continue;
          for (int c = 0; c < k; ++c) {
            if (((b + s.pad - c) % 2 + 2) % 2 != 0) continue;
            int py, oy, px, ox;
            split_parity(a + s.pad - r, &py, &oy);
            split_parity(b + s.pad - c, &px, &ox);
            for (int c0 = 0; c0 < s.CoutS; c0 += BK) {
              KIter it{};
              it.a_c = (int16_t)c0;
              it.a_dx = (int16_t)ox;
              it.a_par = 0;
              it.a_dy = (int16_t)oy;
              it.b_k = (r * k + c) * s.CoutS + c0;
              p.kiters.push_back(it);
            }
          }
        }
        p.args.k_count[z] = (int)p.kiters.size() - p.args.k_begin[z];
      }
  } else {
    // transposed conv: dx[i] = sum_r dy[2i - p + r] w[r]  == a stride-2 conv over dy
    set_tiles(p, dx.N, dx.H, dx.W);
    p.n_classes = 1;
    p.args.sH = dx.sH();
    p.args.sW = dx.sW();
    vk = kViewS2;
    for (int r = 0; r < k; ++r)
      for (int c = 0; c < k; ++c) {
        int hp, hy, wp, wx;
        split_parity(r - s.pad, &hp, &hy);
        split_parity(c - s.pad, &wp, &wx);
        for (int c0 = 0; c0 < s.CoutS; c0 += BK) {
          KIter it{};
          it.a_c = (int16_t)(wp * s.CoutS + c0);
          it.a_dx = (int16_t)wx;
          it.a_par = (int16_t)hp;
          it.a_dy = (int16_t)hy;
          it.b_k = (r * k + c) * s.CoutS + c0;
          p.kiters.push_back(it);
        }
      }
    p.args.k_begin[0] = 0;
    p.args.k_count[0] = (int)p.kiters.size();
  }
  bool patched = false;
  if (!s.transposed && s.stride == 1) {
    // dx[h] = sum_r dy[h + off - r] w[r]: the patch starts at off - (k - 1) and the filter is walked backwards
    const int off = s.reflect ? 0 : s.pad;
    patched = try_tapn(p, dy, false, wt, Kw, k, s.Cin, dx, dx.H, dx.W, off - (k - 1), off - (k - 1), true, nullptr, 0,
                       kActNone, sm_count);
    if (!patched)
      patched = try_patch(p, dy, false, wt, s.CinS, Kw, k, s.CoutS, dx.H, dx.W, off - (k - 1), off - (k - 1), true,
                          sm_count);
  }
  if (!patched) finalize(p, vk, dy, wt, s.CinS, Kw, sm_count);
  const double out_px = s.transposed ? (double)dx.H * dx.W : (double)dy.H * dy.W;
  p.flops = 2.0 * dy.N * out_px * s.Cout * s.Cin * T;
  return p;
}

WgradPlan plan_wgrad(const ConvSpec& s, const TensorDesc& x, const TensorDesc& dy, float* g, int sm_count) {
  CGB_CHECK(tc_supports_wgrad(s), "wgrad: channel counts not supported by the tensor-core path");
  CGB_CHECK(x.C == s.CinS && dy.C == s.CoutS, "wgrad: tensor channels do not match the spec");
  WgradPlan p;
  std::memset(&p.args, 0, sizeof(p.args));
  const int k = s.k, T = s.taps();
  // pixel domain of the reduction: conv -> output pixels; transposed conv -> input pixels
  const int PH = s.transposed ? x.H : dy.H;
  const int PW = s.transposed ? x.W : dy.W;
  const int shift = std::max(3, std::min(6, ilog2_ceil(std::min(PW, 64))));
  const int TWk = 1 << shift, THk = 64 >> shift;
  p.args.tw_shift = shift;
  p.args.tiles_w = (PW + TWk - 1) / TWk;
  p.args.tiles_h = (PH + THk - 1) / THk;
  p.args.N = x.N;
  p.args.T = T;
  p.args.Cout = s.Cout;
  p.args.Cin = s.Cin;
  p.args.g = g;
  p.BNW = std::min(256, s.CinS);
  CGB_CHECK(p.BNW == 64 || p.BNW == 128 || p.BNW == 256, "wgrad: unsupported Cin tile");
  p.m_blocks = (s.CoutS + 127) / 128;

  if (!s.transposed) {
    p.tmDY = view_s1(dy, false, 64, TWk, THk);
    if (s.stride == 1) {
      if (s.reflect) CGB_CHECK(x.halo == s.pad, "reflect conv input must carry a halo equal to the padding");
      p.tmX = view_s1(x, s.reflect, 64, TWk, THk);
    } else {
      p.tmX = view_s2(x, 64, TWk, THk);
    }
    for (int r = 0; r < k; ++r)
      for (int c = 0; c < k; ++c) {
        WTap t{};
        if (s.stride == 1) {
          t.b_dx = (int16_t)(s.reflect ? c : c - s.pad);
          t.b_dy = (int16_t)(s.reflect ? r : r - s.pad);
        } else {
          int hp, hy, wp, wx;
          split_parity(r - s.pad, &hp, &hy);
          split_parity(c - s.pad, &wp, &wx);
          t.b_c = (int16_t)(wp * s.CinS);
          t.b_dx = (int16_t)wx;
          t.b_par = (int16_t)hp;
          t.b_dy = (int16_t)hy;
        }
        t.out_tap = r * k + c;
        p.taps.push_back(t);
      }
  } else {
    p.tmDY = view_s2(dy, 64, TWk, THk);
    p.tmX = view_s1(x, false, 64, TWk, THk);
    for (int r = 0; r < k; ++r)
      for (int c = 0; c < k; ++c) {
        WTap t{};
        int hp, hy, wp, wx;
        split_parity(r - s.pad, &hp, &hy);
        split_parity(c - s.pad, &wp, &wx);
        t.a_c = (int16_t)(wp * s.CoutS);
        t.a_dx = (int16_t)wx;
        t.a_par = (int16_t)hp;
        t.a_dy = (int16_t)hy;
        t.out_tap = r * k + c;
        p.taps.push_back(t);
      }
  }
  p.args.num_taps = (int)p.taps.size();
  p.flops = 2.0 * x.N * (double)PH * PW * s.Cout * s.Cin * T;
  // Weight gradients run back to back on side lanes.  An early-launched successor (programmatic dependent launch)
  // would sit in griddepcontrol.wait holding an SM's worth of shared memory and TMEM that the backward chain's
  // kernels need: no early trigger by default (CGB_WGRAD_TRIGGER=1 restores it).
  static const int wtrigger = std::getenv("CGB_WGRAD_TRIGGER") ? std::atoi(std::getenv("CGB_WGRAD_TRIGGER")) : 0;
  p.args.trigger = wtrigger;
  // stride-1 layers with wide channels: CTA pairs, one filter row per pair (wgrad_pair.cu).  CGB_WGRAD_PAIR=0 disables.
  static const bool pair_on = !(std::getenv("CGB_WGRAD_PAIR") && std::atoi(std::getenv("CGB_WGRAD_PAIR")) == 0);
  // (short reductions stay on the tap-per-CTA kernel: the discriminator's 31 x 31 layer at batch 1 has 16 chunks; measured
  // 15.0 us there against 22.0 us for pairs)
  static const long long pair_min_total = std::getenv("CGB_WGRADP_MIN_TOTAL") ? std::atoll(std::getenv("CGB_WGRADP_MIN_TOTAL")) : 32;
  if (pair_on && !s.transposed && s.stride == 1 && s.k <= 4 && s.Cout == s.CoutS && s.Cin == s.CinS && s.Cout % 256 == 0 &&
      s.Cin % 128 == 0 && (long long)x.N * ((PW + 7) / 8) * ((PH + 7) / 8) >= pair_min_total) {
    WgradPairArgs& a = p.pargs;
    std::memset(&a, 0, sizeof(a));
    a.kh = a.kw = k;
    a.T = T;
    a.Cout = s.Cout;
    a.Cin = s.Cin;
    a.tiles_w = (PW + 7) / 8;
    a.tiles_h = (PH + 7) / 8;
    a.N = x.N;
    a.cin_blocks = s.Cin / 128;
    a.n_units = k * a.cin_blocks * (s.Cout / 256);
    a.x_ox = a.x_oy = s.reflect ? 0 : -s.pad;
    a.g = g;
    a.trigger = wtrigger;
    // K split: about `frac` of the machine's TPCs in pairs, but at least `min_chunks` chunks per pair (the fp32
    // reduction of a pair's 256 x (kw * 128) tile costs about as much as a dozen chunks of MMAs).  The weight gradients
    // run on side lanes next to the backward chain, so the FASTEST launch is not the best one: measured at batch 8
    // (profiles/r02_y_sweep_wgrad_pairs_b8.txt) frac 1.0 / 0.7 / 0.5 / 0.35 / 0.25: residual wgrad 1164 / 885 / 782 / 574 /
    // 456 TFLOP/s but step 22.26 / 21.93 / 21.89 / 21.76 / 21.85 ms; at batch 1 min_chunks 8 / 16 / 32 / 64: 4.75 /
    // 4.34 / 4.19 / 4.23 ms (the tap-per-CTA kernel: 4.28).
    static const double pfrac = std::getenv("CGB_WGRADP_FRAC") ? std::atof(std::getenv("CGB_WGRADP_FRAC")) : 0.4;
    static const long long pmin = std::getenv("CGB_WGRADP_MIN_CHUNKS") ? std::atoll(std::getenv("CGB_WGRADP_MIN_CHUNKS")) : 32;
    const long long chunks = (long long)a.N * a.tiles_w * a.tiles_h;
    long long split = (long long)(sm_count / 2 * pfrac) / a.n_units;
    split = std::max(1LL, std::min(split, std::max(1LL, chunks / pmin)));
    a.split_k = (int)split;
    p.tmDY = view_s1(dy, false, 64, 8, 8);
    p.tmX = view_s1(x, s.reflect, 64, 8 + k - 1, 8);
    p.pair = true;
    p.args.split_k = a.split_k;
    return p;
  }
  const long long total_chunks = (long long)p.args.N * p.args.tiles_w * p.args.tiles_h;
  const long long base_ctas = (long long)p.m_blocks * ((s.CinS + p.BNW - 1) / p.BNW) * T;
  // split-K so that the launch has about `frac` of a wave of CTAs: long K per CTA amortises the prologue and
  // the fp32 reduction epilogue, and leaves room for the other lanes of the step graph
  static const double frac = std::getenv("CGB_WGRAD_FRAC") ? std::atof(std::getenv("CGB_WGRAD_FRAC")) : 0.85;
  // ... but never shorter than `min_chunks` (64) 64-pixel chunks per CTA: below that the prologue and the fp32
  // reduction of the 128 x BNW tile dominate (measured at batch 1: split 6 -> 2 on the residual layers is
  // 7 % faster end to end although the isolated kernel is slower)
  static const long long min_chunks = std::getenv("CGB_WGRAD_MIN_CHUNKS") ? std::atoll(std::getenv("CGB_WGRAD_MIN_CHUNKS")) : 64;
  long long split = (long long)(sm_count * frac) / base_ctas;
  split = std::max(1LL, std::min(split, std::max(1LL, total_chunks / min_chunks)));
  p.args.split_k = (int)split;
  p.flops = 2.0 * x.N * (double)PH * PW * s.Cout * s.Cin * T;
  return p;
}

void run(const IgemmPlan& p, cudaStream_t stream) {
  if (p.tapn) {
    launch_conv_tapn(p.tmA, p.tmB, p.targs, p.num_ctas_m, stream);
    return;
  }
  CGB_CHECK(p.args.kiters != nullptr, "igemm plan has no device K-iteration table");
  if (p.patch) {
    launch_igemm_patch(p.BN, p.MT, p.CG, p.tmA, p.tmB, p.args, p.pargs, p.num_ctas_m, p.n_blocks, stream);
    return;
  }
  for (int z = 0; z < p.n_classes; ++z) CGB_CHECK(p.args.k_count[z] <= 192, "K-iteration table exceeds the smem staging area");
  launch_igemm(p.BN, p.BK, p.CM, p.CN, p.lite, p.tmA, p.tmB, p.args, p.num_tiles, p.n_blocks, p.n_classes, stream);
}

void run(const WgradPlan& p, cudaStream_t stream) {
  if (p.pair) {
    launch_wgrad_pair(p.tmDY, p.tmX, p.pargs, stream);
    return;
  }
  CGB_CHECK(p.args.taps != nullptr, "wgrad plan has no device tap table");
  launch_wgrad(p.BNW, p.tmDY, p.tmX, p.args, p.m_blocks, stream);
}

}  // namespace cgb
