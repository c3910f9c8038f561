// CycleGAN step engine (see engine.h).  Host-side only: every device operation goes through
// conv_plan.h (tcgen05 convolutions) and pointwise.h (HBM-bound kernels).
#include "engine.h"

#include <chrono>

#include <cstdlib>
#include <cstring>

using namespace cgb;

namespace {

ConvSpec make_spec(int cin, int cout, int k, int stride, int pad, bool reflect, bool transposed) {
  ConvSpec s;
  s.Cin = cin;
  s.Cout = cout;
  s.CinS = cin % 64 == 0 ? cin : 16;
  s.CoutS = cout % 64 == 0 ? cout : 16;
  s.k = k;
  s.stride = stride;
  s.pad = pad;
  s.reflect = reflect;
  s.transposed = transposed;
  return s;
}

long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

// The stem's input-side im2col operand of one image set, shared by the stem weight gradients of every pass that reads
// those images: the row-expanded form [N][S + 8][S][64] read as a virtual im2col matrix (small_wgrad.h), or -- with
// CGB_VIRTUAL_COL=0 -- the materialised im2col4 matrix [N][S][S][256]
TensorDesc stem_col(Arena& A, int N, int S) {
  return small_wgrad_virtual_in(7) ? A.tensor(N, S + 8, S, 64, 0) : A.tensor(N, S, S, 256, 0);
}
void build_stem_col(const TensorDesc& image, const TensorDesc& xc, cudaStream_t s) {
  if (xc.C == 64) expand_rows4(image, 7, +1, -3, true, xc, s);
  else im2col4(image, 7, 1, +1, -3, true, xc, s);
}

}  // namespace

cgb_engine::~cgb_engine() {
  drop_graphs();
  for (cudaEvent_t ev : events) cudaEventDestroy(ev);
  for (auto& v : seg_events)
    for (cudaEvent_t ev : v) cudaEventDestroy(ev);
  for (cudaEvent_t ev : grad_events) cudaEventDestroy(ev);
  for (int l = 1; l < kLanes; ++l)
    if (lane_streams[l]) cudaStreamDestroy(lane_streams[l]);
  if (meta) cudaFree(meta);
}

// ------------------------------------------------------------------------------------------------
// Parameter inventory: same tensors, order and names as the stand-in's state_dict
// (oracle/cyclegan_standin.py Generator.__init__ / Discriminator.__init__).
// ------------------------------------------------------------------------------------------------
void cgb_engine::build_inventory() {
  const int nb = cfg.n_blocks;
  for (int net = 0; net < 4; ++net) {
    std::vector<LayerParam>& L = layers[net];
    L.clear();
    auto add = [&](const std::string& name, ConvSpec s, bool has_in) {
      LayerParam p;
      p.name = name;
      p.spec = s;
      p.has_in = has_in;
      p.net = net;
      p.group = net < 2 ? CGB_GROUP_G : CGB_GROUP_D;
      L.push_back(p);
    };
    if (net < 2) {
      add("stem", make_spec(3, 64, 7, 1, 3, true, false), true);
      add("down1", make_spec(64, 128, 3, 2, 1, false, false), true);
      add("down2", make_spec(128, 256, 3, 2, 1, false, false), true);
      for (int i = 0; i < nb; ++i) {
        add("res." + std::to_string(i) + ".conv1", make_spec(256, 256, 3, 1, 1, true, false), true);
        add("res." + std::to_string(i) + ".conv2", make_spec(256, 256, 3, 1, 1, true, false), true);
      }
      add("up1", make_spec(256, 128, 3, 2, 1, false, true), true);
      add("up2", make_spec(128, 64, 3, 2, 1, false, true), true);
      add("head", make_spec(64, 3, 7, 1, 3, true, false), false);
    } else {
      add("conv0", make_spec(3, 64, 4, 2, 1, false, false), false);
      add("conv1", make_spec(64, 128, 4, 2, 1, false, false), true);
      add("conv2", make_spec(128, 256, 4, 2, 1, false, false), true);
      add("conv3", make_spec(256, 512, 4, 1, 1, false, false), true);
      add("conv4", make_spec(512, 1, 4, 1, 1, false, false), false);
    }
  }
  for (int g = 0; g < 2; ++g) {
    long long off = 0, poff = 0;
    for (int net = g * 2; net < g * 2 + 2; ++net)
      for (LayerParam& p : layers[net]) {
        p.w_off = off;
        off = align_up(off + (long long)p.spec.Cout * p.spec.taps() * p.spec.Cin, 4);
        p.b_off = off;
        off = align_up(off + p.spec.Cout, 4);
        p.wf_off = poff;
        poff = align_up(poff + packed_wf_elems(p.spec), 512);
        p.wt_off = poff;
        poff = align_up(poff + packed_wt_elems(p.spec), 512);
        if (p.spec.Cin <= 4 && p.spec.reflect && p.spec.k == 7) {  // generator stem
          p.wx_off = poff;
          poff = align_up(poff + (long long)padded_rows(p.spec.CoutS) * 256, 512);
        }
      }
    group_numel[g] = off;
    pack_elems[g] = poff;
  }
}

// ------------------------------------------------------------------------------------------------
// HBM workspace layout (bf16 NHWC activations; image tensors carry 16 stored channels and a
// 3-pixel reflect halo; residual-stream tensors carry a 1-pixel reflect halo).
// ------------------------------------------------------------------------------------------------
void cgb_engine::layout(Arena& A) {
  // inference-only engines (cgb_engine_create_ex, CGB_FLAG_INFERENCE) carry the packed weights and the
  // module-forward pass; every training tensor is allocated with batch 0
  const int N = infer_only ? 0 : cfg.batch, S = cfg.size, nb = cfg.n_blocks;
  const int H2 = S / 2, H4 = S / 4, H8 = S / 8;
  A.esz = fp32 ? 4 : 2;
  for (int g = 0; g < 2; ++g) pack[g] = static_cast<bf16*>(A.alloc((size_t)pack_elems[g] * sizeof(bf16)));
  auto images = [](const TensorDesc& t, int first, int n) { return t.images(first, n); };
  // The paired schedule does 20 % less kernel work but puts the identity passes on the critical chain.  Measured
  // on B200 with the final kernels: batch 1: 4.74 ms paired vs 4.65 ms unpaired; batch 8: 25.7 ms paired vs 27.2 ms
  // unpaired.  Default: paired from 2 image pairs per GPU up; CGB_PAIR=0 / 1 overrides.
  pair = std::getenv("CGB_PAIR") ? std::atoi(std::getenv("CGB_PAIR")) != 0 : cfg.batch >= 2;  // measured (profiles/r02_aa_sweep_sched_b*.txt): batch 1 4.21 vs 4.22 ms, batch 2 7.27 vs 6.58, batch 8 22.3 vs 21.5
  if (fp32) pair = false;  // validation mode: one pass per image set, one gradient buffer per pass kind
  if (fp32 && !infer_only) {
    for (int k = 0; k < 3; ++k) gslot[0][k] = static_cast<float*>(A.alloc((size_t)group_numel[0] * sizeof(float)));
    for (int k = 0; k < 2; ++k) gslot[1][k] = static_cast<float*>(A.alloc((size_t)group_numel[1] * sizeof(float)));
  }
  if (pair) {
    reals3 = A.tensor(3 * N, S, S, 16, 3);
    pair_out[0] = A.tensor(2 * N, S, S, 16, 3);  // [fake_B; idt_A]
    pair_out[1] = A.tensor(2 * N, S, S, 16, 3);  // [fake_A; idt_B]
    img[CGB_IMG_REAL_A] = images(reals3, 0, N);
    img[CGB_IMG_REAL_B] = images(reals3, N, N);
    pair_in[0] = images(reals3, 0, 2 * N);       // [real_A; real_B] -> G_AB
    pair_in[1] = images(reals3, N, 2 * N);       // [real_B; real_A] -> G_BA
    img[CGB_IMG_FAKE_B] = images(pair_out[0], 0, N);
    img[CGB_IMG_IDT_A] = images(pair_out[0], N, N);
    img[CGB_IMG_FAKE_A] = images(pair_out[1], 0, N);
    img[CGB_IMG_IDT_B] = images(pair_out[1], N, N);
    img[CGB_IMG_REC_A] = A.tensor(N, S, S, 16, 3);
    img[CGB_IMG_REC_B] = A.tensor(N, S, S, 16, 3);
  } else {
    for (int i = 0; i < 8; ++i) img[i] = A.tensor(N, S, S, 16, 3);
  }
  mod_in = A.tensor(cfg.batch, S, S, 16, 3);
  mod_out = A.tensor(cfg.batch, S, S, 16, 3);
  for (int i = 0; i < 2; ++i) staging[i] = static_cast<float*>(A.alloc((size_t)N * 3 * S * S * sizeof(float)));
  for (int i = 0; i < 2; ++i) staging_u8[i] = static_cast<unsigned char*>(A.alloc((size_t)N * 3 * S * S));
  losses = static_cast<float*>(A.alloc(64 * sizeof(float)));
  for (int g = 0; g < 2; ++g) {
    adam_step[g] = static_cast<int*>(A.alloc(64));
    adam_hyper[g] = static_cast<float*>(A.alloc(64));
  }

  gen.resize(7);
  for (size_t gi = 0; gi < gen.size(); ++gi) {
    GenPass& P = gen[gi];
    // paired schedule: passes 0 / 2 carry 2N images, the identity passes 4 / 5 do not exist
    const int N = gi == 6 ? cfg.batch : infer_only ? 0 : !pair ? cfg.batch : (gi == 0 || gi == 2) ? 2 * cfg.batch : (gi == 4 || gi == 5) ? 0 : cfg.batch;
    P.y_stem = A.tensor(N, S, S, 64, 0);
    P.a_stem = A.tensor(N, S, S, 64, 0);
    P.y_d1 = A.tensor(N, H2, H2, 128, 0);
    P.a_d1 = A.tensor(N, H2, H2, 128, 0);
    P.y_d2 = A.tensor(N, H4, H4, 256, 0);
    P.xp.resize(nb + 1);
    P.y1.resize(nb);
    P.bp.resize(nb);
    P.y2.resize(nb);
    P.xp[0] = A.tensor(N, H4, H4, 256, 1);
    for (int k = 0; k < nb; ++k) {
      P.y1[k] = A.tensor(N, H4, H4, 256, 0);
      P.bp[k] = A.tensor(N, H4, H4, 256, 1);
      P.y2[k] = A.tensor(N, H4, H4, 256, 0);
      P.xp[k + 1] = A.tensor(N, H4, H4, 256, 1);
    }
    P.y_u1 = A.tensor(N, H2, H2, 128, 0);
    P.a_u1 = A.tensor(N, H2, H2, 128, 0);
    P.y_u2 = A.tensor(N, S, S, 64, 0);
    P.a_u2p = A.tensor(N, S, S, 64, 3);
    // statistics: IN layer l has layers[0][l].spec.Cout channels
    P.stat_off.assign(5 + 2 * nb, 0);
    long long off = 0;
    for (int l = 0; l < 5 + 2 * nb; ++l) {
      P.stat_off[l] = off;
      off += (long long)N * layers[0][l].spec.Cout;
    }
    P.stats_bytes = (size_t)off * sizeof(float2);
    P.stats = static_cast<float2*>(A.alloc(P.stats_bytes));
    P.bstats = static_cast<float2*>(A.alloc(P.stats_bytes));
  }
  dis.resize(7);
  for (size_t di = 0; di < dis.size(); ++di) {
    DisPass& D = dis[di];
    const int N = di >= 5 ? ((pool_size > 0 && !infer_only) ? cfg.batch : 0) : (di == 4 || !infer_only) ? cfg.batch : 0;
    D.l0 = A.tensor(N, H2, H2, 64, 0);
    D.y1 = A.tensor(N, H4, H4, 128, 0);
    D.a1 = A.tensor(N, H4, H4, 128, 0);
    D.y2 = A.tensor(N, H8, H8, 256, 0);
    D.a2 = A.tensor(N, H8, H8, 256, 0);
    D.y3 = A.tensor(N, H8 - 1, H8 - 1, 512, 0);
    D.a3 = A.tensor(N, H8 - 1, H8 - 1, 512, 0);
    D.logits = A.tensor(N, H8 - 2, H8 - 2, 16, 0);
    D.stat_off[0] = 0;
    D.stat_off[1] = (long long)N * 128;
    D.stat_off[2] = (long long)N * (128 + 256);
    D.stats_bytes = (size_t)N * (128 + 256 + 512) * sizeof(float2);
    D.stats = static_cast<float2*>(A.alloc(D.stats_bytes));
    D.bstats = static_cast<float2*>(A.alloc(D.stats_bytes));
  }
  for (int gi = 0; gi < kPassLanes; ++gi) {  // one backward scratch set per lane
    GenScratch& g = gs[gi];
    // paired schedule: sets 0 / 1 serve the cycle passes (N images), sets 2 / 3 the paired passes (2N)
    const int N = infer_only ? 0 : (pair && gi >= 2) ? 2 * cfg.batch : cfg.batch;
    g.dpre_head = A.tensor(N, S, S, 16, 0);
    g.dxp_head = A.tensor(N, S + 6, S + 6, 64, 0);
    g.dyF = A.tensor(N, S, S, 64, 0);
    g.dxF = A.tensor(N, S, S, 64, 0);
    g.dyH = A.tensor(N, H2, H2, 128, 0);
    g.dxH = A.tensor(N, H2, H2, 128, 0);
    g.dyQ = A.tensor(N, H4, H4, 256, 0);
    g.dyQ2 = A.tensor(N, H4, H4, 256, 0);
    g.GQ[0] = A.tensor(N, H4, H4, 256, 0);
    g.GQ[1] = A.tensor(N, H4, H4, 256, 0);
    g.dbpQ = A.tensor(N, H4 + 2, H4 + 2, 256, 0);
    g.dxpQ = A.tensor(N, H4 + 2, H4 + 2, 256, 0);
    g.colbuf_elems = (size_t)N * (S + 6) * (S + 6) * 256;
    g.colbuf = static_cast<bf16*>(A.alloc(g.colbuf_elems * sizeof(bf16)));
  }
  for (int i = 0; i < 2; ++i) {
    dxp_img[i] = A.tensor(N, S + 6, S + 6, 16, 0);
    dx_D0[i] = A.tensor(N, S, S, 16, 0);
  }
  if (pair) {
    xcol3 = stem_col(A, 3 * N, S);
    xcol[0] = images(xcol3, 0, N);
    xcol[1] = images(xcol3, N, N);
    pair_xcol[0] = images(xcol3, 0, 2 * N);
    pair_xcol[1] = images(xcol3, N, 2 * N);
    for (int i = 2; i < 4; ++i) xcol[i] = stem_col(A, N, S);
  } else {
    for (int i = 0; i < 4; ++i) xcol[i] = stem_col(A, N, S);
  }
  for (DisScratch& d : ds) {
    d.dlogits = A.tensor(N, H8 - 2, H8 - 2, 16, 0);
    d.dx3 = A.tensor(N, H8 - 1, H8 - 1, 512, 0);
    d.dy3 = A.tensor(N, H8 - 1, H8 - 1, 512, 0);
    d.dx2 = A.tensor(N, H8, H8, 256, 0);
    d.dy2 = A.tensor(N, H8, H8, 256, 0);
    d.dx1 = A.tensor(N, H4, H4, 128, 0);
    d.dy1 = A.tensor(N, H4, H4, 128, 0);
    d.dx0 = A.tensor(N, H2, H2, 64, 0);
    d.dpre0 = A.tensor(N, H2, H2, 64, 0);
    d.colbuf_elems = (size_t)N * H2 * H2 * 64;
    d.colbuf = static_cast<bf16*>(A.alloc(d.colbuf_elems * sizeof(bf16)));
  }
  if (pool_size > 0 && !infer_only) {
    for (int i = 0; i < 2; ++i) {
      pool_img[i] = A.tensor(pool_size, S, S, 16, 0);
      pool_din[i] = A.tensor(cfg.batch, S, S, 16, 0);
    }
    pool_dec = static_cast<int*>(A.alloc((size_t)2 * cfg.batch * 2 * sizeof(int)));
  }
  A.alloc(1024);  // tail guard
}

void cgb_engine::drop_graphs() {
  for (Segment& s : segments) {
    if (s.exec) cudaGraphExecDestroy(s.exec);
    s.exec = nullptr;
  }
}

// First call: eager on `st` (configures kernel attributes, validates).  Second call: stream-capture the
// sequence with the lanes mapped to distinct streams, instantiate, launch.  Later calls: replay.
// The legacy default stream cannot be captured (eager then).
void cgb_engine::run_segment(int seg, cudaStream_t st) {
  Segment& S = segments[seg];
  static const bool no_graph = std::getenv("CGB_NO_GRAPH") != nullptr;
  const bool can_graph = !no_graph && st != nullptr && !S.failed;
  if (can_graph && S.exec == nullptr && S.calls >= 1) {
    lane_streams[0] = st;
    for (int l = 1; l < kLanes; ++l)
      if (!lane_streams[l]) CGB_CUDA(cudaStreamCreateWithFlags(&lane_streams[l], cudaStreamNonBlocking));
    cudaGraph_t g = nullptr;
    cudaError_t err = cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed);
    if (err == cudaSuccess) {
      try {
        size_t next_event = 0;
        std::vector<cudaEvent_t>& evs = seg_events[seg];
        for (const Program* p : S.seq) p->run_lanes(lane_streams, evs, &next_event);
      } catch (...) {
        cudaStreamEndCapture(st, &g);
        if (g) cudaGraphDestroy(g);
        S.failed = true;
        throw;
      }
      err = cudaStreamEndCapture(st, &g);
      if (err == cudaSuccess) err = cudaGraphInstantiate(&S.exec, g, 0);
      if (g) cudaGraphDestroy(g);
    }
    if (err != cudaSuccess) {
      S.exec = nullptr;
      S.failed = true;
      cudaGetLastError();
      const std::string msg = std::string("CUDA graph capture failed: ") + cudaGetErrorString(err);
      set_last_error(msg);
      if (std::getenv("CGB_REQUIRE_GRAPH")) throw Error(msg);
    }
  }
  ++S.calls;
  if (S.exec) {
    CGB_CUDA(cudaGraphLaunch(S.exec, st));
  } else {
    for (const Program* p : S.seq) p->run(st);
  }
}

// Captures the whole step with external event-record nodes at the program markers, replays it, and returns
// "label lane ms" lines (ms since the first marker).  Development profiling only.
std::string cgb_engine::timeline(cudaStream_t st) {
  CGB_CHECK(st != nullptr, "timeline needs a non-default stream");
  lane_streams[0] = st;
  for (int l = 1; l < kLanes; ++l)
    if (!lane_streams[l]) CGB_CUDA(cudaStreamCreateWithFlags(&lane_streams[l], cudaStreamNonBlocking));
  cudaStream_t mapped[kLanes];
  const int max_lanes = std::getenv("CGB_MAX_LANES") ? std::atoi(std::getenv("CGB_MAX_LANES")) : kLanes;
  for (int l = 0; l < kLanes; ++l) mapped[l] = lane_streams[l % std::max(1, max_lanes)];
  for (const Program* p : segments[CGB_SEG_STEP].seq) p->run(st);  // eager warm-up
  CGB_CUDA(cudaStreamSynchronize(st));
  std::vector<Program::Mark> marks;
  std::vector<cudaEvent_t> evs;
  size_t next_event = 0;
  cudaGraph_t g = nullptr;
  cudaGraphExec_t exec = nullptr;
  CGB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
  for (const Program* p : segments[CGB_SEG_STEP].seq) p->run_lanes(mapped, evs, &next_event, &marks);
  CGB_CUDA(cudaStreamEndCapture(st, &g));
  CGB_CUDA(cudaGraphInstantiate(&exec, g, 0));
  for (int i = 0; i < 3; ++i) CGB_CUDA(cudaGraphLaunch(exec, st));
  CGB_CUDA(cudaStreamSynchronize(st));
  std::string out;
  for (const Program::Mark& m : marks) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, marks[0].ev, m.ev);
    char line[160];
    snprintf(line, sizeof(line), "%-32s lane %d  %8.3f ms\n", m.label.c_str(), m.lane, ms);
    out += line;
  }
  for (const Program::Mark& m : marks) cudaEventDestroy(m.ev);
  for (cudaEvent_t ev : evs) cudaEventDestroy(ev);
  cudaGraphExecDestroy(exec);
  cudaGraphDestroy(g);
  return out;
}

// Hang hunt (development): replays the training step from a graph that carries an external event after every phase
// marker (fine == 0) or after every op (fine != 0) until a replay does not finish within stall_ms; then reports, per
// lane, the last marker that completed and the first that did not.  Returns "" when all `steps` replays finished.
std::string cgb_engine::hang_probe(cudaStream_t st, int steps, int stall_ms, int fine) {
  CGB_CHECK(st != nullptr, "hang probe needs a non-default stream");
  lane_streams[0] = st;
  for (int l = 1; l < kLanes; ++l)
    if (!lane_streams[l]) CGB_CUDA(cudaStreamCreateWithFlags(&lane_streams[l], cudaStreamNonBlocking));
  for (const Program* p : segments[CGB_SEG_STEP].seq) p->run(st);  // eager warm-up
  CGB_CUDA(cudaStreamSynchronize(st));
  std::vector<Program::Mark> marks;
  std::vector<cudaEvent_t> evs;
  size_t next_event = 0;
  cudaGraph_t g = nullptr;
  cudaGraphExec_t exec = nullptr;
  CGB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
  for (const Program* p : segments[CGB_SEG_STEP].seq) p->run_lanes(lane_streams, evs, &next_event, &marks, fine != 0);
  CGB_CUDA(cudaStreamEndCapture(st, &g));
  CGB_CUDA(cudaGraphInstantiate(&exec, g, 0));
  cudaEvent_t done;
  CGB_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
  std::string out;
  for (int i = 0; i < steps && out.empty(); ++i) {
    CGB_CUDA(cudaGraphLaunch(exec, st));
    if (i % 16 != 15 && i != steps - 1) continue;  // keep a few replays in flight, like a training loop
    CGB_CUDA(cudaEventRecord(done, st));
    const auto t0 = std::chrono::steady_clock::now();
    while (cudaEventQuery(done) == cudaErrorNotReady) {
      if (std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count() > stall_ms) {
        out = "HANG near replay " + std::to_string(i) + " (" + std::to_string(marks.size()) + " markers)\n";
        int last_done[kLanes], first_open[kLanes];
        for (int l = 0; l < kLanes; ++l) last_done[l] = first_open[l] = -1;
        for (int m = 0; m < (int)marks.size(); ++m) {
          const bool ok = cudaEventQuery(marks[m].ev) == cudaSuccess;
          if (ok) last_done[marks[m].lane] = m;
          else if (first_open[marks[m].lane] < 0) first_open[marks[m].lane] = m;
        }
        for (int l = 0; l < kLanes; ++l) {
          out += "lane " + std::to_string(l) + ": last finished [" + (last_done[l] >= 0 ? marks[last_done[l]].label : std::string("-")) +
                 "]  first unfinished [" + (first_open[l] >= 0 ? marks[first_open[l]].label : std::string("-")) + "]\n";
        }
        break;
      }
    }
  }
  if (out.empty()) {
    CGB_CUDA(cudaStreamSynchronize(st));
    for (const Program::Mark& m : marks) cudaEventDestroy(m.ev);
    for (cudaEvent_t ev : evs) cudaEventDestroy(ev);
    cudaEventDestroy(done);
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(g);
  }  // (after a hang nothing can be released: the caller exits the process)
  return out;
}

// Times every convolution launch of ONE generator pass and one discriminator pass individually (CUDA events,
// `reps` back-to-back launches each) and returns "name us TFLOP/s" lines.  Development profiling only.
std::string cgb_engine::profile_ops(cudaStream_t st, int reps) {
  std::string out;
  cudaEvent_t e0, e1;
  CGB_CUDA(cudaEventCreate(&e0));
  CGB_CUDA(cudaEventCreate(&e1));
  const Program* progs[3] = {&prog_cycle, &prog_G, &prog_D};
  std::vector<std::string> seen;
  for (const Program* p : progs)
    for (size_t i = 0; i < p->ops.size(); ++i) {
      if (p->names[i].empty()) continue;
      bool dup = false;
      for (const std::string& s : seen) dup = dup || s == p->names[i];
      if (dup) continue;
      seen.push_back(p->names[i]);
      p->ops[i](st);
      CGB_CUDA(cudaEventRecord(e0, st));
      for (int r = 0; r < reps; ++r) p->ops[i](st);
      CGB_CUDA(cudaEventRecord(e1, st));
      CGB_CUDA(cudaStreamSynchronize(st));
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      const double us = ms * 1e3 / reps;
      char line[200];
      snprintf(line, sizeof(line), "%-44s %9.2f us %8.1f TFLOP/s\n", p->names[i].c_str(), us,
               p->flops[i] / (us * 1e-6) * 1e-12);
      out += line;
    }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return out;
}

void* cgb_engine::meta_upload(const void* src, size_t bytes) {
  meta_off = (meta_off + 255) & ~size_t(255);
  CGB_CHECK(meta_off + bytes <= meta_cap, "engine meta buffer exhausted");
  void* dst = static_cast<uint8_t*>(meta) + meta_off;
  CGB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  meta_off += bytes;
  return dst;
}

// ------------------------------------------------------------------------------------------------
// Program recording
// ------------------------------------------------------------------------------------------------
void cgb_engine::record_programs() {
  const int N = cfg.batch, S = cfg.size, nb = cfg.n_blocks;
  cgb_engine* E = this;
  double* flops = &conv_flops;
  double dummy_flops = 0;

  // stats != nullptr: InstanceNorm (sum, sumsq) accumulated by the conv epilogue
  auto add_fprop = [E](Program& pr, double* fl, const LayerParam& L, const TensorDesc& x, const TensorDesc& y,
                       int act, float2* stats) {
    const float* bias = L.has_in ? nullptr : E->P[L.group] + L.b_off;
    if (E->fp32) {
      const ConvSpec sp = L.spec;
      const float* w = E->P[L.group] + L.w_off;
      const double f = 2.0 * x.N * (sp.transposed ? (double)x.H * x.W : (double)y.H * y.W) * sp.Cout * sp.Cin * sp.taps();
      pr.cur_name = "fprop(fp32) " + L.name;
      pr.add([sp, x, w, bias, act, y](cudaStream_t st) { f32::conv_fprop(sp, x, w, bias, act, y, st); }, 1, kOpIgemm, f);
      pr.cur_name.clear();
      *fl += f;
      return;
    }
    IgemmPlan p = plan_fprop(L.spec, x, E->pack[L.group] + L.wf_off, y, bias, act, E->sm_count);
    p.args.stats = reinterpret_cast<float*>(stats);
    p.args.kiters = static_cast<const KIter*>(E->meta_upload(p.kiters.data(), p.kiters.size() * sizeof(KIter)));
    E->igemm_plans.push_back(p);
    const IgemmPlan* pp = &E->igemm_plans.back();
    pr.cur_name = "fprop " + L.name + " BN" + std::to_string(p.BN) + " ctas" + std::to_string(p.num_tiles * p.n_blocks * p.n_classes);
    pr.add([pp](cudaStream_t st) { run(*pp, st); }, 1, kOpIgemm, p.flops);
    pr.cur_name.clear();
    *fl += p.flops;
  };
  auto add_dgrad = [E](Program& pr, double* fl, const LayerParam& L, const TensorDesc& dy, const TensorDesc& dx) {
    if (E->fp32) {
      const ConvSpec sp = L.spec;
      const float* w = E->P[L.group] + L.w_off;
      const double f = 2.0 * dy.N * (sp.transposed ? (double)dx.H * dx.W : (double)dy.H * dy.W) * sp.Cout * sp.Cin * sp.taps();
      pr.cur_name = "dgrad(fp32) " + L.name;
      pr.add([sp, dy, w, dx](cudaStream_t st) { f32::conv_dgrad(sp, dy, w, dx, st); }, 1, kOpIgemm, f);
      pr.cur_name.clear();
      *fl += f;
      return;
    }
    IgemmPlan p = plan_dgrad(L.spec, dy, E->pack[L.group] + L.wt_off, dx, E->sm_count);
    p.args.kiters = static_cast<const KIter*>(E->meta_upload(p.kiters.data(), p.kiters.size() * sizeof(KIter)));
    E->igemm_plans.push_back(p);
    const IgemmPlan* pp = &E->igemm_plans.back();
    pr.cur_name = "dgrad " + L.name + " BN" + std::to_string(p.BN) + " ctas" + std::to_string(p.num_tiles * p.n_blocks * p.n_classes);
    pr.add([pp](cudaStream_t st) { run(*pp, st); }, 1, kOpIgemm, p.flops);
    pr.cur_name.clear();
    *fl += p.flops;
  };
  // gbase: the flat gradient buffer of the layer's group this pass writes to (cgb_engine::grad_base)
  auto add_wgrad_on_lane = [E](Program& pr, double* fl, const LayerParam& L, const TensorDesc& x, const TensorDesc& dy,
                               bf16* colbuf, size_t colbuf_elems, const TensorDesc* precol, float* gbase) {
    float* g = gbase + L.w_off;
    if (E->fp32) {
      const ConvSpec sp = L.spec;
      const double f = 2.0 * x.N * (sp.transposed ? (double)x.H * x.W : (double)dy.H * dy.W) * sp.Cout * sp.Cin * sp.taps();
      pr.cur_name = "wgrad(fp32) " + L.name;
      pr.add([sp, x, dy, g](cudaStream_t st) { f32::conv_wgrad(sp, x, dy, g, st); }, 1, kOpWgradTc, f);
      pr.cur_name.clear();
      *fl += f;
      return;
    }
    if (tc_supports_wgrad(L.spec)) {
      WgradPlan p = plan_wgrad(L.spec, x, dy, g, E->sm_count);
      p.args.taps = static_cast<const WTap*>(E->meta_upload(p.taps.data(), p.taps.size() * sizeof(WTap)));
      E->wgrad_plans.push_back(p);
      const WgradPlan* pp = &E->wgrad_plans.back();
      pr.cur_name = "wgrad " + L.name + " split" + std::to_string(p.args.split_k);
      pr.add([pp](cudaStream_t st) { run(*pp, st); }, 1, kOpWgradTc, p.flops);
      pr.cur_name.clear();
      *fl += p.flops;
    } else {
      // 3-/1-channel layers: explicit im2col of the skinny operand, then a plain tensor-core GEMM
      SmallWgradPlan p = plan_wgrad_small(L.spec, x, dy, g, colbuf, colbuf_elems, E->sm_count, precol);
      p.gemm.args.taps = static_cast<const WTap*>(E->meta_upload(p.gemm.taps.data(), p.gemm.taps.size() * sizeof(WTap)));
      if (!p.row_map.empty())
        p.gemm.args.row_map = static_cast<const int*>(E->meta_upload(p.row_map.data(), p.row_map.size() * sizeof(int)));
      if (!p.col_map.empty())
        p.gemm.args.col_map = static_cast<const int*>(E->meta_upload(p.col_map.data(), p.col_map.size() * sizeof(int)));
      E->small_wgrad_plans.push_back(p);
      const SmallWgradPlan* pp = &E->small_wgrad_plans.back();
      pr.cur_name = "wgrad(im2col) " + L.name;
      pr.add([pp](cudaStream_t st) { run(*pp, st); }, 2, kOpWgradSmall, p.flops);
      pr.cur_name.clear();
      *fl += p.flops;
    }
  };
  auto add_norm = [E](Program& pr, const TensorDesc& y, float2* stats, int act, const TensorDesc* residual,
                      const TensorDesc& out) {
    if (E->fp32) {  // statistics (mean, rstd) + apply in one deterministic kernel
      pr.cur_name = "in_forward(fp32) C" + std::to_string(y.C) + " " + std::to_string(y.H) + "x" + std::to_string(y.W);
      const bool has_res = residual != nullptr;
      const TensorDesc r = has_res ? *residual : TensorDesc();
      pr.add([y, stats, act, has_res, r, out](cudaStream_t st) { f32::in_forward(y, stats, act, has_res ? &r : nullptr, out, st); },
             1, kOpNorm);
      pr.cur_name.clear();
      return;
    }
    // statistics were accumulated by the producing conv's epilogue
    pr.cur_name = "in_apply C" + std::to_string(y.C) + " " + std::to_string(y.H) + "x" + std::to_string(y.W) +
                  " halo" + std::to_string(out.halo) + (residual ? " +res" : "");
    if (residual) {
      const TensorDesc r = *residual;
      pr.add([y, stats, act, r, out](cudaStream_t st) { in_apply(y, stats, act, &r, out, st); }, 1, kOpNorm);
    } else {
      pr.add([y, stats, act, out](cudaStream_t st) { in_apply(y, stats, act, nullptr, out, st); }, 1, kOpNorm);
    }
    pr.cur_name.clear();
  };
  // InstanceNorm + activation backward; da_store (engine-owned tensor) receives the assembled gradient
  auto add_in_bwd_raw = [E](Program& pr, const TensorDesc& y, const float2* stats, float2* bstats, GradSrc g, int act,
                            const TensorDesc* da_store, const TensorDesc& dy) {
    if (E->fp32) {
      pr.cur_name = "in_backward(fp32) C" + std::to_string(y.C) + " " + std::to_string(y.H) + "x" + std::to_string(y.W);
      pr.add([y, stats, g, act, da_store, dy](cudaStream_t st) { f32::in_backward(y, stats, g, act, da_store, dy, st); }, 1,
             kOpNorm);
      pr.cur_name.clear();
      return;
    }
    const std::string shape = " C" + std::to_string(y.C) + " " + std::to_string(y.H) + "x" + std::to_string(y.W) +
                              (g.g1 ? " g1" : "") + (g.g2 ? " g2fold" + std::to_string(g.fold) : "") +
                              (da_store ? " +da" : "");
    if (in_bwd_fused_supported(y)) {  // one cluster kernel: reductions + apply from shared memory
      pr.cur_name = "in_bwd_fused" + shape;
      pr.add([y, stats, g, act, da_store, dy](cudaStream_t st) {
               CGB_CHECK(in_bwd_fused(y, stats, g, act, da_store, dy, st), "fused InstanceNorm backward refused a planned shape");
             }, 1, kOpNorm);
      pr.cur_name.clear();
      return;
    }
    pr.cur_name = "in_bwd_reduce" + shape;
    pr.add([y, stats, bstats, g, act, da_store](cudaStream_t st) { in_bwd_reduce(y, stats, g, act, da_store, bstats, st); },
           1, kOpNorm);
    pr.cur_name = "in_bwd_apply" + shape;
    GradSrc g2 = g;
    if (da_store) {
      g2 = GradSrc();
      g2.g1 = da_store;
    }
    pr.add([y, stats, bstats, g2, act, dy](cudaStream_t st) { in_bwd_apply(y, stats, bstats, g2, act, dy, st); }, 1,
           kOpNorm);
    pr.cur_name.clear();
  };

  // ---------------------------------------------------------------- generator forward
  // xcol_in: shared im2col4 of `in` (stem as a GEMM) or nullptr (stem as a 49-tap implicit GEMM);
  // xcol_out: when non-null the im2col4 of the generated image is produced for the pass that consumes it
  // xcol_lane >= 0: the im2col4 of the output (only the stem WEIGHT gradient of the consuming pass reads it) is
  // issued on that side lane instead of the chain
  auto emit_gen_forward = [&](Program& pr, double* fl, GenPass& P, int net, const TensorDesc& in,
                              const TensorDesc& out, bool fill_out_halo, const TensorDesc* xcol_in,
                              const TensorDesc* xcol_out, int xcol_lane = -1) {
    P.net = net;
    P.in = in;
    P.out = out;
    P.xcol = xcol_in;
    const std::vector<LayerParam>& L = E->layers[net];
    float2* st = P.stats;
    pr.add([st, bytes = P.stats_bytes](cudaStream_t s) { CGB_CUDA(cudaMemsetAsync(st, 0, bytes, s)); }, 0, kOpMemset);
    static const bool stem_gemm = std::getenv("CGB_STEM_GEMM") != nullptr;
    if (xcol_in && stem_gemm && !E->fp32 && xcol_in->C == 256) {
      // stem as a plain GEMM over the im2col4 matrix: K = 256 (49 taps x 4 channels, zero padded); superseded by the
      // 16-channel patch-resident conv (the im2col4 matrix still feeds the stem weight gradient)
      LayerParam Lx = L[0];
      Lx.name = "stem(gemm)";
      Lx.spec.Cin = L[0].spec.Cin * L[0].spec.taps(); Lx.spec.CinS = 256; Lx.spec.k = 1; Lx.spec.stride = 1; Lx.spec.pad = 0; Lx.spec.reflect = false;
      Lx.wf_off = L[0].wx_off;
      add_fprop(pr, fl, Lx, *xcol_in, P.y_stem, kActNone, st + P.stat_off[0]);
    } else {
      add_fprop(pr, fl, L[0], in, P.y_stem, kActNone, st + P.stat_off[0]);
    }
    add_norm(pr, P.y_stem, st + P.stat_off[0], kActRelu, nullptr, P.a_stem);
    add_fprop(pr, fl, L[1], P.a_stem, P.y_d1, kActNone, st + P.stat_off[1]);
    add_norm(pr, P.y_d1, st + P.stat_off[1], kActRelu, nullptr, P.a_d1);
    add_fprop(pr, fl, L[2], P.a_d1, P.y_d2, kActNone, st + P.stat_off[2]);
    add_norm(pr, P.y_d2, st + P.stat_off[2], kActRelu, nullptr, P.xp[0]);
    for (int k = 0; k < nb; ++k) {
      add_fprop(pr, fl, L[3 + 2 * k], P.xp[k], P.y1[k], kActNone, st + P.stat_off[3 + 2 * k]);
      add_norm(pr, P.y1[k], st + P.stat_off[3 + 2 * k], kActRelu, nullptr, P.bp[k]);
      add_fprop(pr, fl, L[4 + 2 * k], P.bp[k], P.y2[k], kActNone, st + P.stat_off[4 + 2 * k]);
      add_norm(pr, P.y2[k], st + P.stat_off[4 + 2 * k], kActNone, &P.xp[k], P.xp[k + 1]);
    }
    add_fprop(pr, fl, L[3 + 2 * nb], P.xp[nb], P.y_u1, kActNone, st + P.stat_off[3 + 2 * nb]);
    add_norm(pr, P.y_u1, st + P.stat_off[3 + 2 * nb], kActRelu, nullptr, P.a_u1);
    add_fprop(pr, fl, L[4 + 2 * nb], P.a_u1, P.y_u2, kActNone, st + P.stat_off[4 + 2 * nb]);
    add_norm(pr, P.y_u2, st + P.stat_off[4 + 2 * nb], kActRelu, nullptr, P.a_u2p);
    add_fprop(pr, fl, L[5 + 2 * nb], P.a_u2p, out, kActTanh, nullptr);
    if (fill_out_halo) pr.add([out](cudaStream_t s) { fill_reflect_halo(out, s); });
    if (xcol_out) {
      const TensorDesc xc = *xcol_out;
      const int main_lane = pr.cur_lane;
      if (xcol_lane >= 0) {
        pr.dep(main_lane, xcol_lane);
        pr.cur_lane = xcol_lane;
      }
      pr.add([out, xc](cudaStream_t s) { build_stem_col(out, xc, s); }, 1, kOpNorm);
      pr.cur_lane = main_lane;
    }
  };

  // ---------------------------------------------------------------- generator backward
  // Seeds of the head gradient: L1 against `target` (value into `loss_slot`) and / or the external gradient
  // `gsrc`.  Paired pass (target2 != nullptr): the first half of the batch is seeded by (target, gsrc), the
  // second half by the L1 term against target2.
  auto emit_gen_backward = [&](Program& pr, double* fl, GenPass& P, GenScratch& S, const TensorDesc* target,
                               float l1_scale, int loss_slot, GradSrc gsrc, const TensorDesc* dxp_img_out,
                               const TensorDesc* target2 = nullptr, float l1_scale2 = 0.f, int loss_slot2 = -1,
                               const std::function<void(Program&, int, int)>* after_wgrad = nullptr) {
    const std::vector<LayerParam>& L = E->layers[P.net];
    float* Gg = E->grad_base(CGB_GROUP_G, P.gslot);
    float2* st = P.stats;
    float2* bs = P.bstats;
    const int main_lane = pr.cur_lane, wlane = pr.cur_lane + kPassLanes;
    // Second weight-gradient lane (CGB_WGRAD_LANES=2, off by default): consecutive layers' weight gradients alternate
    // between two side lanes and run concurrently.  At batch 1 a residual weight gradient (18 CTAs, ~29 us) takes
    // longer than the chain spends on a layer once the InstanceNorm backward is one kernel (~27 us); measured on B200
    // (profiles/r02_d_sweep_b1.txt) the second lane changes nothing: 4.643 vs 4.650 ms at batch 1, 26.90 vs 26.81 ms at
    // batch 8 -- the step is bound by SM time, not by this chain.
    static const int n_wlanes = std::getenv("CGB_WGRAD_LANES") ? std::atoi(std::getenv("CGB_WGRAD_LANES")) : 1;
    const int wlane2 = n_wlanes >= 2 ? pr.cur_lane + 2 * kPassLanes : wlane;
    int n_wgrads = 0;
    // weight gradient on the side lane, beside the input gradient of the same layer
    std::vector<int> wgrad_done;  // record ids on the side lanes, one per weight gradient issued so far
    auto add_wgrad = [&](Program& pr_, double* fl_, const LayerParam& Lp, const TensorDesc& x, const TensorDesc& dy,
                         bf16* colbuf, size_t colbuf_elems, const TensorDesc* precol = nullptr) {
      // (a gradient that reads a precomputed im2col matrix stays on the first side lane: that is the lane the
      // matrix was built on / that waits for it, see record_step)
      const int wl = ((n_wgrads++ & 1) && precol == nullptr) ? wlane2 : wlane, other = wl == wlane ? wlane2 : wlane;
      pr_.dep(main_lane, wl);
      pr_.cur_lane = wl;
      add_wgrad_on_lane(pr_, fl_, Lp, x, dy, colbuf, colbuf_elems, precol, Gg);
      wgrad_done.push_back(pr_.record(wl));
      // data-parallel hook (still on the weight-gradient lane): this layer's gradient is final once the other passes
      // of the same generator are done too -- the callback adds those waits and the bucket's external event
      if (after_wgrad) (*after_wgrad)(pr_, (int)(&Lp - L.data()), other != wl ? other : -1);
      pr_.cur_lane = main_lane;
    };
    // A layer's dy buffer is read by its weight gradient on the side lane.  Consecutive layers never share a dy
    // buffer (dyQ / dyQ2 alternate along the residual chain), layers two apart may: before overwriting, wait for
    // the weight gradient issued two layers back -- the previous one keeps running beside this layer's chain.
    auto add_in_bwd = [&](Program& pr_, const TensorDesc& y, const float2* stats, float2* bstats, GradSrc g, int act,
                          const TensorDesc* da_store, const TensorDesc& dy) {
      if (wgrad_done.size() >= 2) pr_.wait(main_lane, wgrad_done[wgrad_done.size() - 2]);
      add_in_bwd_raw(pr_, y, stats, bstats, g, act, da_store, dy);
    };
    pr.add([bs, bytes = P.stats_bytes](cudaStream_t s) { CGB_CUDA(cudaMemsetAsync(bs, 0, bytes, s)); }, 0, kOpMemset);
    const LayerParam& head = L[5 + 2 * nb];
    if (target2) {
      const int half = P.out.N / 2;
      auto images = [](const TensorDesc& t, int first, int n) { return t.images(first, n); };
      const TensorDesc outA = images(P.out, 0, half), outB = images(P.out, half, half);
      const TensorDesc dpreA = images(S.dpre_head, 0, half), dpreB = images(S.dpre_head, half, half);
      const TensorDesc tg2 = *target2;
      float* slot2 = loss_slot2 >= 0 ? E->losses + loss_slot2 : nullptr;
      CGB_CHECK(target == nullptr, "paired backward: the first half is seeded by an external gradient only");
      pr.add([outA, gsrc, dpreA](cudaStream_t s) { tanh_bwd(outA, nullptr, 0.f, gsrc, 3, dpreA, nullptr, s); });
      pr.add([outB, tg2, l1_scale2, dpreB, slot2](cudaStream_t s) { tanh_bwd(outB, &tg2, l1_scale2, GradSrc(), 3, dpreB, slot2, s); });
      const TensorDesc dpre = S.dpre_head;
      float* gb = Gg + head.b_off;
      pr.add([dpre, gb](cudaStream_t s) { bias_grad(dpre, 3, gb, s); });
    } else {
      const TensorDesc out = P.out, dpre = S.dpre_head;
      float* slot = loss_slot >= 0 ? E->losses + loss_slot : nullptr;
      if (target) {
        const TensorDesc tg = *target;
        pr.add([out, tg, l1_scale, gsrc, dpre, slot](cudaStream_t s) { tanh_bwd(out, &tg, l1_scale, gsrc, 3, dpre, slot, s); });
      } else {
        pr.add([out, gsrc, dpre](cudaStream_t s) { tanh_bwd(out, nullptr, 0.f, gsrc, 3, dpre, nullptr, s); });
      }
      float* gb = Gg + head.b_off;
      pr.add([dpre, gb](cudaStream_t s) { bias_grad(dpre, 3, gb, s); });
    }
    add_wgrad(pr, fl, head, P.a_u2p, S.dpre_head, S.colbuf, S.colbuf_elems);
    add_dgrad(pr, fl, head, S.dpre_head, S.dxp_head);
    GradSrc g;
    g.g2 = &S.dxp_head;
    g.fold = 3;
    add_in_bwd(pr, P.y_u2, st + P.stat_off[4 + 2 * nb], bs + P.stat_off[4 + 2 * nb], g, kActRelu, nullptr, S.dyF);
    add_wgrad(pr, fl, L[4 + 2 * nb], P.a_u1, S.dyF, S.colbuf, S.colbuf_elems);
    add_dgrad(pr, fl, L[4 + 2 * nb], S.dyF, S.dxH);
    g = GradSrc();
    g.g1 = &S.dxH;
    add_in_bwd(pr, P.y_u1, st + P.stat_off[3 + 2 * nb], bs + P.stat_off[3 + 2 * nb], g, kActRelu, nullptr, S.dyH);
    add_wgrad(pr, fl, L[3 + 2 * nb], P.xp[nb], S.dyH, S.colbuf, S.colbuf_elems);
    int cur = 0;
    add_dgrad(pr, fl, L[3 + 2 * nb], S.dyH, S.GQ[cur]);  // G_nb: gradient w.r.t. the residual stream output
    for (int k = nb - 1; k >= 0; --k) {
      // gradient w.r.t. x_{k+1}: G_{k+1} = G_{k+2} + fold(dxp_{k+1})   (k == nb-1: G_nb as is)
      g = GradSrc();
      g.g1 = &S.GQ[cur];
      const TensorDesc* store = nullptr;
      if (k != nb - 1) {
        g.g2 = &S.dxpQ;
        g.fold = 1;
        store = &S.GQ[cur ^ 1];
      }
      add_in_bwd(pr, P.y2[k], st + P.stat_off[4 + 2 * k], bs + P.stat_off[4 + 2 * k], g, kActNone, store, S.dyQ);
      if (store) cur ^= 1;
      add_wgrad(pr, fl, L[4 + 2 * k], P.bp[k], S.dyQ, S.colbuf, S.colbuf_elems);
      add_dgrad(pr, fl, L[4 + 2 * k], S.dyQ, S.dbpQ);
      g = GradSrc();
      g.g2 = &S.dbpQ;
      g.fold = 1;
      add_in_bwd(pr, P.y1[k], st + P.stat_off[3 + 2 * k], bs + P.stat_off[3 + 2 * k], g, kActRelu, nullptr, S.dyQ2);
      add_wgrad(pr, fl, L[3 + 2 * k], P.xp[k], S.dyQ2, S.colbuf, S.colbuf_elems);
      add_dgrad(pr, fl, L[3 + 2 * k], S.dyQ2, S.dxpQ);
    }
    // down2 output feeds block 0 twice (conv path + skip): G_0 = G_1 + fold(dxp_0)
    g = GradSrc();
    g.g1 = &S.GQ[cur];
    g.g2 = &S.dxpQ;
    g.fold = 1;
    add_in_bwd(pr, P.y_d2, st + P.stat_off[2], bs + P.stat_off[2], g, kActRelu, &S.GQ[cur ^ 1], S.dyQ);
    add_wgrad(pr, fl, L[2], P.a_d1, S.dyQ, S.colbuf, S.colbuf_elems);
    add_dgrad(pr, fl, L[2], S.dyQ, S.dxH);
    g = GradSrc();
    g.g1 = &S.dxH;
    add_in_bwd(pr, P.y_d1, st + P.stat_off[1], bs + P.stat_off[1], g, kActRelu, nullptr, S.dyH);
    add_wgrad(pr, fl, L[1], P.a_stem, S.dyH, S.colbuf, S.colbuf_elems);
    add_dgrad(pr, fl, L[1], S.dyH, S.dxF);
    g = GradSrc();
    g.g1 = &S.dxF;
    add_in_bwd(pr, P.y_stem, st + P.stat_off[0], bs + P.stat_off[0], g, kActRelu, nullptr, S.dyF);
    add_wgrad(pr, fl, L[0], P.in, S.dyF, S.colbuf, S.colbuf_elems, P.xcol);
    if (dxp_img_out) add_dgrad(pr, fl, L[0], S.dyF, *dxp_img_out);
    pr.dep(wlane, main_lane);
    if (wlane2 != wlane) pr.dep(wlane2, main_lane);
  };

  // ---------------------------------------------------------------- discriminator
  auto emit_dis_forward = [&](Program& pr, double* fl, DisPass& D, int net, const TensorDesc& in) {
    D.net = net;
    D.in = in;
    const std::vector<LayerParam>& L = E->layers[net];
    float2* st = D.stats;
    pr.add([st, bytes = D.stats_bytes](cudaStream_t s) { CGB_CUDA(cudaMemsetAsync(st, 0, bytes, s)); }, 0, kOpMemset);
    add_fprop(pr, fl, L[0], in, D.l0, kActLeaky, nullptr);
    add_fprop(pr, fl, L[1], D.l0, D.y1, kActNone, st + D.stat_off[0]);
    add_norm(pr, D.y1, st + D.stat_off[0], kActLeaky, nullptr, D.a1);
    add_fprop(pr, fl, L[2], D.a1, D.y2, kActNone, st + D.stat_off[1]);
    add_norm(pr, D.y2, st + D.stat_off[1], kActLeaky, nullptr, D.a2);
    add_fprop(pr, fl, L[3], D.a2, D.y3, kActNone, st + D.stat_off[2]);
    add_norm(pr, D.y3, st + D.stat_off[2], kActLeaky, nullptr, D.a3);
    add_fprop(pr, fl, L[4], D.a3, D.logits, kActNone, nullptr);
  };
  auto emit_dis_backward = [&](Program& pr, double* fl, DisPass& D, DisScratch& S, float target, float w,
                               int loss_slot, bool weight_grads, const TensorDesc* dx_img_out) {
    const std::vector<LayerParam>& L = E->layers[D.net];
    float* Gd = E->grad_base(CGB_GROUP_D, D.gslot);
    float2* st = D.stats;
    float2* bs = D.bstats;
    const int main_lane = pr.cur_lane, wlane = pr.cur_lane + kPassLanes;
    auto add_wgrad = [&](Program& pr_, double* fl_, const LayerParam& Lp, const TensorDesc& x, const TensorDesc& dy,
                         bf16* colbuf, size_t colbuf_elems, const TensorDesc* precol = nullptr) {
      pr_.dep(main_lane, wlane);
      pr_.cur_lane = wlane;
      add_wgrad_on_lane(pr_, fl_, Lp, x, dy, colbuf, colbuf_elems, precol, Gd);
      pr_.cur_lane = main_lane;
    };
    auto add_in_bwd = [&](Program& pr_, const TensorDesc& y, const float2* stats, float2* bstats, GradSrc g, int act,
                          const TensorDesc* da_store, const TensorDesc& dy) {
      pr_.dep(wlane, main_lane);
      add_in_bwd_raw(pr_, y, stats, bstats, g, act, da_store, dy);
    };
    pr.add([bs, bytes = D.stats_bytes](cudaStream_t s) { CGB_CUDA(cudaMemsetAsync(bs, 0, bytes, s)); }, 0, kOpMemset);
    {
      const TensorDesc lg = D.logits;
      const TensorDesc* dl = &S.dlogits;
      float* slot = E->losses + loss_slot;
      pr.add([lg, target, w, slot, dl](cudaStream_t s) { mse_loss(lg, target, w, slot, dl, s); });
    }
    if (weight_grads) {
      const TensorDesc dl = S.dlogits;
      float* gb = Gd + L[4].b_off;
      pr.add([dl, gb](cudaStream_t s) { bias_grad(dl, 1, gb, s); });
      add_wgrad(pr, fl, L[4], D.a3, S.dlogits, S.colbuf, S.colbuf_elems);
    }
    add_dgrad(pr, fl, L[4], S.dlogits, S.dx3);
    GradSrc g;
    g.g1 = &S.dx3;
    add_in_bwd(pr, D.y3, st + D.stat_off[2], bs + D.stat_off[2], g, kActLeaky, nullptr, S.dy3);
    if (weight_grads) add_wgrad(pr, fl, L[3], D.a2, S.dy3, S.colbuf, S.colbuf_elems);
    add_dgrad(pr, fl, L[3], S.dy3, S.dx2);
    g.g1 = &S.dx2;
    add_in_bwd(pr, D.y2, st + D.stat_off[1], bs + D.stat_off[1], g, kActLeaky, nullptr, S.dy2);
    if (weight_grads) add_wgrad(pr, fl, L[2], D.a1, S.dy2, S.colbuf, S.colbuf_elems);
    add_dgrad(pr, fl, L[2], S.dy2, S.dx1);
    g.g1 = &S.dx1;
    add_in_bwd(pr, D.y1, st + D.stat_off[0], bs + D.stat_off[0], g, kActLeaky, nullptr, S.dy1);
    if (weight_grads) add_wgrad(pr, fl, L[1], D.l0, S.dy1, S.colbuf, S.colbuf_elems);
    add_dgrad(pr, fl, L[1], S.dy1, S.dx0);
    {
      const TensorDesc l0 = D.l0, dx0_ = S.dx0, dp = S.dpre0;
      pr.add([l0, dx0_, dp](cudaStream_t s) { leaky_bwd(l0, dx0_, dp, s); });
      if (weight_grads) {
        float* gb = Gd + L[0].b_off;
        pr.add([dp, gb](cudaStream_t s) { bias_grad(dp, 64, gb, s); });
        add_wgrad(pr, fl, L[0], D.in, S.dpre0, S.colbuf, S.colbuf_elems);
      }
    }
    if (dx_img_out) add_dgrad(pr, fl, L[0], S.dpre0, *dx_img_out);
    pr.dep(wlane, main_lane);
  };

  // ---- optimiser + bf16 weight refresh
  for (int g = 0; g < 2; ++g) {
    std::vector<PackEntry> table;
    int max_elems = 0;
    for (int net = g * 2; net < g * 2 + 2; ++net)
      for (const LayerParam& p : layers[net]) {
        PackEntry e{};
        e.src_off = p.w_off;
        e.wf_off = p.wf_off;
        e.wt_off = p.wt_off;
        e.Cout = p.spec.Cout;
        e.Cin = p.spec.Cin;
        e.T = p.spec.taps();
        e.CinS = p.spec.CinS;
        e.CoutS = p.spec.CoutS;
        e.wx_off = p.wx_off;
        e.wx_pitch = 256;
        table.push_back(e);
        max_elems = std::max(max_elems, e.Cout * e.Cin * e.T);
      }
    pack_table[g] = static_cast<PackEntry*>(meta_upload(table.data(), table.size() * sizeof(PackEntry)));
    pack_count[g] = (int)table.size();
    pack_max[g] = max_elems;
    const PackEntry* tb = pack_table[g];
    const int cnt = pack_count[g], mx = pack_max[g];
    float* p = P[g];
    bf16* arena = pack[g];
    prog_refresh[g].add([p, tb, cnt, mx, arena](cudaStream_t s) { pack_weights(p, tb, cnt, mx, arena, s); });
    float *gg = G[g], *mm = M[g], *vv = V[g];
    const long long n = group_numel[g];
    int* stp = adam_step[g];
    float* hyp = adam_hyper[g];
    prog_adam[g].add(
        [E, p, gg, mm, vv, n, stp, hyp](cudaStream_t s) {
          cgb::adam_step(p, gg, mm, vv, n, E->cfg.lr, E->cfg.beta1, E->cfg.beta2, E->cfg.eps, stp, hyp, E->grad_scale, s);
        },
        2);
    prog_adam[g].add([p, tb, cnt, mx, arena](cudaStream_t s) { pack_weights(p, tb, cnt, mx, arena, s); });
  }
  if (!infer_only) {
  // fp32 validation mode: per-pass gradient buffers (generators: fake / rec / idt pass; discriminators: real / fake)
  gen[0].gslot = gen[2].gslot = 0;
  gen[1].gslot = gen[3].gslot = 1;
  gen[4].gslot = gen[5].gslot = 2;
  dis[2].gslot = dis[3].gslot = 0;
  dis[0].gslot = dis[1].gslot = dis[5].gslot = dis[6].gslot = 1;
  auto add_sum_slots = [E](Program& pr, int group) {
    if (!E->fp32) return;
    float* dst = E->G[group];
    const float *s0 = E->gslot[group][0], *s1 = E->gslot[group][1], *s2 = group == CGB_GROUP_G ? E->gslot[group][2] : nullptr;
    const long long n = E->group_numel[group];
    pr.add([dst, s0, s1, s2, n](cudaStream_t st) { f32::sum_slots(dst, s0, s1, s2, n, st); });
  };
  // ---------------------------------------------------------------- step programs
  const TensorDesc &fake_B = img[CGB_IMG_FAKE_B], &rec_A = img[CGB_IMG_REC_A], &fake_A = img[CGB_IMG_FAKE_A],
                   &rec_B = img[CGB_IMG_REC_B], &idt_A = img[CGB_IMG_IDT_A], &idt_B = img[CGB_IMG_IDT_B],
                   &real_A = img[CGB_IMG_REAL_A], &real_B = img[CGB_IMG_REAL_B];
  (void)rec_A; (void)rec_B; (void)idt_A; (void)idt_B;

  {  // staging (fp32 NCHW) -> bf16 NHWC images with reflect halo
    const float* sa = staging[0];
    const float* sb = staging[1];
    const TensorDesc ra = real_A, rb = real_B;
    prog_set_inputs.add([sa, ra](cudaStream_t s) { nchw_to_nhwc(sa, 3, ra, s); });
    prog_set_inputs.add([sb, rb](cudaStream_t s) { nchw_to_nhwc(sb, 3, rb, s); });
    prog_set_inputs_lite.add([sa, ra](cudaStream_t s) { nchw_to_nhwc(sa, 3, ra, s); });
    prog_set_inputs_lite.add([sb, rb](cudaStream_t s) { nchw_to_nhwc(sb, 3, rb, s); });
    if (pair) {
      // [real_A; real_B; real_A]: images 0..2N feed G_AB, images N..3N feed G_BA; one im2col over all three
      const TensorDesc ra2 = reals3.images(2 * N, N);
      prog_set_inputs.add([sa, ra2](cudaStream_t s) { nchw_to_nhwc(sa, 3, ra2, s); });
      prog_set_inputs_lite.add([sa, ra2](cudaStream_t s) { nchw_to_nhwc(sa, 3, ra2, s); });
      const TensorDesc r3 = reals3, x3 = xcol3;
      prog_set_inputs.add([r3, x3](cudaStream_t s) { build_stem_col(r3, x3, s); }, 1, kOpNorm);
    } else {
      // one im2col per distinct input image serves the stem forward and the stem weight gradient of every pass
      const TensorDesc xa = xcol[0], xb = xcol[1];
      prog_set_inputs.add([ra, xa](cudaStream_t s) { build_stem_col(ra, xa, s); }, 1, kOpNorm);
      prog_set_inputs.add([rb, xb](cudaStream_t s) { build_stem_col(rb, xb, s); }, 1, kOpNorm);
    }
  }
  // im2col4 of a generated image for the pass that consumes it (paired schedule: the fake half of the pair output)
  auto add_xcol = [](Program& pr, const TensorDesc& image, const TensorDesc& xc) {
    pr.add([image, xc](cudaStream_t s) { build_stem_col(image, xc, s); }, 1, kOpNorm);
  };
  // pass order: 0 fake_B = G_AB(real_A), 1 rec_A = G_BA(fake_B), 2 fake_A = G_BA(real_B),
  //             3 rec_B = G_AB(fake_A), 4 idt_A = G_AB(real_B), 5 idt_B = G_BA(real_A)
  // Independent passes are recorded on different lanes (parallel branches of the step graph).
  {
    Program& pr = prog_cycle;
    pr.mark("cycle begin");
    pr.fork();
    if (pair) {
      pr.cur_lane = 0;
      emit_gen_forward(pr, flops, gen[0], CGB_NET_G_AB, pair_in[0], pair_out[0], true, &pair_xcol[0], nullptr);
      add_xcol(pr, fake_B, xcol[2]);
      pr.mark("fwd fake_B + idt_A done");
      pr.cur_lane = 1;
      emit_gen_forward(pr, flops, gen[2], CGB_NET_G_BA, pair_in[1], pair_out[1], true, &pair_xcol[1], nullptr);
      add_xcol(pr, fake_A, xcol[3]);
      pr.mark("fwd fake_A + idt_B done");
    } else {
    pr.cur_lane = 0;
    emit_gen_forward(pr, flops, gen[0], CGB_NET_G_AB, real_A, img[CGB_IMG_FAKE_B], true, &xcol[0], &xcol[2]);
    pr.mark("fwd fake_B done");
    pr.cur_lane = 1;
    emit_gen_forward(pr, flops, gen[2], CGB_NET_G_BA, real_B, img[CGB_IMG_FAKE_A], true, &xcol[1], &xcol[3]);
    pr.mark("fwd fake_A done");
    pr.cur_lane = 2;
    emit_gen_forward(pr, flops, gen[4], CGB_NET_G_AB, real_B, img[CGB_IMG_IDT_A], false, &xcol[1], nullptr);
    pr.mark("fwd idt_A done");
    pr.cur_lane = 3;
    emit_gen_forward(pr, flops, gen[5], CGB_NET_G_BA, real_A, img[CGB_IMG_IDT_B], false, &xcol[0], nullptr);
    pr.mark("fwd idt_B done");
    }
    pr.cur_lane = 0;
    emit_gen_forward(pr, flops, gen[1], CGB_NET_G_BA, fake_B, img[CGB_IMG_REC_A], false, &xcol[2], nullptr);
    pr.mark("fwd rec_A done");
    pr.cur_lane = 1;
    emit_gen_forward(pr, flops, gen[3], CGB_NET_G_AB, fake_A, img[CGB_IMG_REC_B], false, &xcol[3], nullptr);
    pr.mark("fwd rec_B done");
    pr.join();
    pr.cur_lane = 0;
    pr.mark("cycle end");
  }

  const float numel_img = (float)N * 3.f * S * S;
  {  // ---- G phase (backward part; the six forwards are prog_cycle)
    Program& pr = prog_G;
    float* gG = G[CGB_GROUP_G];
    const size_t gbytes = (size_t)group_numel[CGB_GROUP_G] * sizeof(float);
    float* ls = losses;
    pr.cur_lane = 0;
    pr.add([gG, gbytes](cudaStream_t s) { CGB_CUDA(cudaMemsetAsync(gG, 0, gbytes, s)); }, 0, kOpMemset);
    pr.add([ls](cudaStream_t s) { CGB_CUDA(cudaMemsetAsync(ls, 0, 64 * sizeof(float), s)); }, 0, kOpMemset);
    pr.fork();
    // identity passes (L1 seeds), then the adversarial terms (D frozen: input gradients only)
    pr.cur_lane = 2;
    if (!pair) {
      emit_gen_backward(pr, flops, gen[4], gs[2], &real_B, cfg.lambda_B * cfg.lambda_idt / numel_img, CGB_LOSS_IDT_A, GradSrc(), nullptr);
      pr.mark("bwd idt_A done");
    }
    emit_dis_forward(pr, flops, dis[0], CGB_NET_D_A, fake_B);
    emit_dis_backward(pr, flops, dis[0], ds[0], 1.f, 1.f, CGB_LOSS_G_A, false, &dx_D0[0]);
    pr.mark("D_A(fake_B) fwd+dgrad done");
    pr.cur_lane = 3;
    if (!pair) {
      emit_gen_backward(pr, flops, gen[5], gs[3], &real_A, cfg.lambda_A * cfg.lambda_idt / numel_img, CGB_LOSS_IDT_B, GradSrc(), nullptr);
      pr.mark("bwd idt_B done");
    }
    emit_dis_forward(pr, flops, dis[1], CGB_NET_D_B, fake_A);
    emit_dis_backward(pr, flops, dis[1], ds[1], 1.f, 1.f, CGB_LOSS_G_B, false, &dx_D0[1]);
    pr.mark("D_B(fake_A) fwd+dgrad done");
    // cycle passes: also produce the gradient w.r.t. the fake images (padded domain of the next stem)
    pr.cur_lane = 0;
    emit_gen_backward(pr, flops, gen[1], gs[0], &real_A, cfg.lambda_A / numel_img, CGB_LOSS_CYCLE_A, GradSrc(), &dxp_img[0]);
    pr.mark("bwd rec_A done");
    pr.cur_lane = 1;
    emit_gen_backward(pr, flops, gen[3], gs[1], &real_B, cfg.lambda_B / numel_img, CGB_LOSS_CYCLE_B, GradSrc(), &dxp_img[1]);
    pr.mark("bwd rec_B done");
    // the passes that produced the fakes: gradient = D's input gradient + folded stem gradient
    pr.dep(2, 0);
    pr.dep(3, 1);
    GradSrc g;
    g.g1 = &dx_D0[0];
    g.g2 = &dxp_img[0];
    g.fold = 3;
    pr.cur_lane = 0;
    if (pair)
      emit_gen_backward(pr, flops, gen[0], gs[2], nullptr, 0.f, -1, g, nullptr, &real_B,
                        cfg.lambda_B * cfg.lambda_idt / numel_img, CGB_LOSS_IDT_A);
    else
      emit_gen_backward(pr, flops, gen[0], gs[0], nullptr, 0.f, -1, g, nullptr);
    pr.mark("bwd fake_B done");
    g.g1 = &dx_D0[1];
    g.g2 = &dxp_img[1];
    pr.cur_lane = 1;
    if (pair)
      emit_gen_backward(pr, flops, gen[2], gs[3], nullptr, 0.f, -1, g, nullptr, &real_A,
                        cfg.lambda_A * cfg.lambda_idt / numel_img, CGB_LOSS_IDT_B);
    else
      emit_gen_backward(pr, flops, gen[2], gs[1], nullptr, 0.f, -1, g, nullptr);
    pr.mark("bwd fake_A done");
    pr.join();
    pr.cur_lane = 0;
    add_sum_slots(pr, CGB_GROUP_G);
    pr.mark("G phase end");
  }
  // D-phase pass on the fakes.  Without a pool the G-phase forward of D on the fake is reused (D unchanged since).
  // With the image history pool the discriminator sees pool.query(fake): exchange kernel, then a forward of its own.
  auto emit_dis_fake_phase = [&](Program& pr, double* fl, int side) {
    const int dnet = side == 0 ? CGB_NET_D_A : CGB_NET_D_B;
    const int slot = side == 0 ? CGB_LOSS_D_A : CGB_LOSS_D_B;
    if (pool_size > 0) {
      const TensorDesc fk = side == 0 ? fake_B : fake_A, pl = pool_img[side], din = pool_din[side];
      const int* dec = pool_dec + (size_t)side * cfg.batch * 2;
      pr.add([fk, pl, dec, din](cudaStream_t s) { pool_exchange(fk, pl, dec, din, s); });
      emit_dis_forward(pr, fl, dis[5 + side], dnet, pool_din[side]);
      emit_dis_backward(pr, fl, dis[5 + side], ds[side], 0.f, 0.5f, slot, true, nullptr);
    } else {
      emit_dis_backward(pr, fl, dis[side], ds[side], 0.f, 0.5f, slot, true, nullptr);
    }
  };
  {  // ---- D phase: real passes are new; the fake passes reuse the G-phase forward (D unchanged since)
    Program& pr = prog_D;
    float* gD = G[CGB_GROUP_D];
    const size_t gbytes = (size_t)group_numel[CGB_GROUP_D] * sizeof(float);
    pr.cur_lane = 0;
    pr.add([gD, gbytes](cudaStream_t s) { CGB_CUDA(cudaMemsetAsync(gD, 0, gbytes, s)); }, 0, kOpMemset);
    pr.fork();
    pr.cur_lane = 0;
    emit_dis_forward(pr, flops, dis[2], CGB_NET_D_A, real_B);
    emit_dis_backward(pr, flops, dis[2], ds[0], 1.f, 0.5f, CGB_LOSS_D_A, true, nullptr);
    emit_dis_fake_phase(pr, flops, 0);
    pr.cur_lane = 1;
    emit_dis_forward(pr, flops, dis[3], CGB_NET_D_B, real_A);
    emit_dis_backward(pr, flops, dis[3], ds[1], 1.f, 0.5f, CGB_LOSS_D_B, true, nullptr);
    emit_dis_fake_phase(pr, flops, 1);
    pr.join();
    pr.cur_lane = 0;
    add_sum_slots(pr, CGB_GROUP_D);
    pr.mark("D phase end");
  }
  // ---- the whole step as one schedule, recorded twice: with Adam(D) in the shadow of the generator backward
  //      chains (single GPU) and without any optimiser (data parallel: the gradients are all-reduced first).
  const bool xcol_side = std::getenv("CGB_STEM_GEMM") == nullptr && std::getenv("CGB_XCOL_ON_CHAIN") == nullptr;
  auto record_step = [&](Program& pr, bool with_adam_d) {  // Lanes 0/1 carry the critical chains
     //      fwd fake -> fwd rec -> bwd rec -> bwd fake; lanes 2/3 do the identity passes, the frozen-D input
     //      gradients and then the entire D phase in the shadow of those chains.
    double sink = 0;  // FLOPs are already accounted by prog_cycle / prog_G / prog_D
    float* gG = G[CGB_GROUP_G];
    float* gD = G[CGB_GROUP_D];
    const size_t gGb = (size_t)group_numel[CGB_GROUP_G] * sizeof(float), gDb = (size_t)group_numel[CGB_GROUP_D] * sizeof(float);
    float* ls = losses;
    // Data-parallel program (no optimiser inside): gradient ranges become final while the step is still running and
    // are announced through external events (cgb_engine::grad_buckets), so the caller's all-reduce of a bucket
    // overlaps the rest of the backward pass.  A generator's gradient gets contributions from three passes (fake,
    // rec, idt); the fake pass is the last one, its weight gradients run head -> stem on the weight-gradient lane.
    const bool dp = !with_adam_d;
    int ev_rec_done[2] = {-1, -1};  // [g]: the rec pass through generator g (G_AB: rec_B, G_BA: rec_A) is complete
    int ev_idt_done[2] = {-1, -1};  // [g]: the identity pass through generator g is complete
    int ev_gstep = -1;              // in-step optimiser: the generators' step counter has been advanced (lane 2)
    bool gstep_waited = false;
    int pend_lo[2] = {-1, -1}, pend_hi[2] = {-1, -1};  // bucket of each generator whose bf16 packs are not refreshed yet
    auto add_pack_layers = [&](Program& p, int gnet, int lo, int hi) {
      if (lo < 0 || hi <= lo) return;
      const int nl = (int)layers[gnet].size();
      const PackEntry* tb = pack_table[CGB_GROUP_G] + gnet * nl + lo;
      const int cnt = hi - lo, mx = pack_max[CGB_GROUP_G];
      float* pm = P[CGB_GROUP_G];
      bf16* arena = pack[CGB_GROUP_G];
      p.add([pm, tb, cnt, mx, arena](cudaStream_t st) { pack_weights(pm, tb, cnt, mx, arena, st); });
    };
    // Buckets per generator.  Data parallel: 2 -- every bucket is one more collective on the communication stream and at
    // 8 GPUs their latencies, not their bytes, are what is left exposed (measured on 8 x B200, batch 1 per GPU,
    // profiles/r02_oo_dp8_sweep.txt: 4 buckets 4.683 ms/step, 2 buckets 4.588, one bucket after the step 5.07).
    // Single GPU (in-step Adam): 4.  CGB_DP_BUCKETS / CGB_ADAM_BUCKETS override.
    static const int dp_buckets = std::getenv("CGB_DP_BUCKETS") ? std::max(1, std::atoi(std::getenv("CGB_DP_BUCKETS"))) : 2;
    static const int adam_buckets = std::getenv("CGB_ADAM_BUCKETS") ? std::max(1, std::atoi(std::getenv("CGB_ADAM_BUCKETS"))) : 4;
    const int n_dp_buckets = dp ? dp_buckets : adam_buckets;
    // Single-GPU program: the same bucket boundaries drive the generators' optimiser INSIDE the step.  Once a bucket is
    // final, Adam on that range runs on lane 2 (G_AB) / 3 (G_BA) -- idle after the D phase -- while the backward chains
    // continue; the bf16 packs of a bucket are refreshed when the NEXT bucket of the same generator is final (the chain
    // has left those layers, nothing else reads them), the last bucket's after the join.  Only that last, smallest
    // bucket remains exposed instead of the whole 0.25 ms Adam + refresh.  CGB_ADAM_OVERLAP=0: optimiser after the step.
    static const bool adam_overlap_on = !(std::getenv("CGB_ADAM_OVERLAP") && std::atoi(std::getenv("CGB_ADAM_OVERLAP")) == 0);
    const bool adam_overlap = !dp && !fp32 && n_dp_buckets > 1 && adam_overlap_on;
    if (!dp) step_has_adam_g = adam_overlap;
    auto bucket_hook = [&](int gnet) {
      return std::function<void(Program&, int, int)>([&, gnet](Program& p, int layer, int other_wlane) {
        if (!(dp || adam_overlap) || fp32 || n_dp_buckets <= 1) return;  // (validation mode: gradients are final after the slot sum)
        const std::vector<LayerParam>& L = layers[gnet];
        // bucket j covers layers [lo_j, lo_{j-1}); lo_0 = end, residual blocks split evenly, the last bucket ends at the stem
        int lo = -1, hi = (int)L.size();
        for (int j = 1; j <= n_dp_buckets; ++j) {
          // (boundaries on residual-block starts, spaced so that the LAST bucket -- whose all-reduce cannot overlap
          // anything -- is the smallest: 4 buckets of a 9-block generator hold 3.9 / 2.4 / 3.5 / 1.6 M parameters)
          const int l = j == n_dp_buckets ? 0 : 3 + 2 * std::max(0, nb - (j * (nb + 1) + n_dp_buckets / 2) / n_dp_buckets);
          if (l == layer) lo = l;
          if (lo < 0) hi = l;
          if (lo >= 0) break;
        }
        if (lo < 0 || lo >= hi) return;
        if (other_wlane >= 0) p.dep(other_wlane, p.cur_lane);  // the previous layer's gradient ran on the other side lane
        if (ev_rec_done[gnet] >= 0) p.wait(p.cur_lane, ev_rec_done[gnet]);
        if (ev_idt_done[gnet] >= 0) p.wait(p.cur_lane, ev_idt_done[gnet]);
        const long long net_end = gnet == 0 ? layers[1][0].w_off : group_numel[CGB_GROUP_G];
        const long long off = L[lo].w_off, end = hi < (int)L.size() ? L[hi].w_off : net_end;
        if (adam_overlap) {
          const int ev_final = p.record(p.cur_lane);
          const int back = p.cur_lane, al = 2 + gnet;
          p.cur_lane = al;
          p.wait(al, ev_final);
          if (al == 3 && !gstep_waited) {  // the step counter / bias corrections were advanced on lane 2
            p.wait(3, ev_gstep);
            gstep_waited = true;
          }
          float *pp = P[CGB_GROUP_G] + off, *gg = G[CGB_GROUP_G] + off, *mm = M[CGB_GROUP_G] + off, *vv = V[CGB_GROUP_G] + off;
          const long long n = end - off;
          int* stp = adam_step[CGB_GROUP_G];
          float* hyp = adam_hyper[CGB_GROUP_G];
          p.add([E, pp, gg, mm, vv, n, stp, hyp](cudaStream_t st) {
            cgb::adam_range(pp, gg, mm, vv, n, E->cfg.beta1, E->cfg.beta2, E->cfg.eps, stp, hyp, E->grad_scale, false, st);
          });
          add_pack_layers(p, gnet, pend_lo[gnet], pend_hi[gnet]);  // the previous bucket of this generator
          pend_lo[gnet] = lo;
          pend_hi[gnet] = hi;
          p.cur_lane = back;
          return;
        }
        int nth = 0;  // how many buckets of this generator came before
        for (const GradBucket& b : grad_buckets) nth += b.net == gnet;
        GradBucket gb{CGB_GROUP_G, off, end - off};
        gb.net = gnet;
        gb.layer_lo = lo;
        gb.layer_hi = hi;
        gb.order = 1 + 2 * nth + gnet;
        grad_buckets.push_back(gb);
        p.ext_event((int)grad_buckets.size() - 1);
      });
    };
    pr.cur_lane = 0;
    pr.mark("step begin");
    pr.add([gG, gGb](cudaStream_t s) { CGB_CUDA(cudaMemsetAsync(gG, 0, gGb, s)); }, 0, kOpMemset);
    pr.add([gD, gDb](cudaStream_t s) { CGB_CUDA(cudaMemsetAsync(gD, 0, gDb, s)); }, 0, kOpMemset);
    pr.add([ls](cudaStream_t s) { CGB_CUDA(cudaMemsetAsync(ls, 0, 64 * sizeof(float), s)); }, 0, kOpMemset);
    pr.fork();
    // The im2col4 matrices feed only the stem WEIGHT gradients (the stem forward is a patch-resident conv), so
    // they are built on the weight-gradient lanes, off the critical chains (xcol_side; the im2col-GEMM stem of
    // CGB_STEM_GEMM=1 needs them on the chain and keeps the full prog_set_inputs).
    if (xcol_side) {
      if (pair) {
        pr.cur_lane = 4;
        add_xcol(pr, reals3, xcol3);
        const int ev_x = pr.record(4);
        pr.wait(5, ev_x);
      } else {
        pr.cur_lane = 4;
        add_xcol(pr, real_A, xcol[0]);
        const int ev_xa = pr.record(4);
        pr.cur_lane = 5;
        add_xcol(pr, real_B, xcol[1]);
        const int ev_xb = pr.record(5);
        pr.wait(6, ev_xb);  // idt_A = G_AB(real_B): its stem weight gradient runs on lane 6
        pr.wait(7, ev_xa);  // idt_B = G_BA(real_A)
      }
    }
    const int xl0 = xcol_side ? 4 : -1, xl1 = xcol_side ? 5 : -1;
    int ev_fake_B, ev_fake_A;
    if (pair) {
      pr.cur_lane = 0;
      emit_gen_forward(pr, &sink, gen[0], CGB_NET_G_AB, pair_in[0], pair_out[0], true, &pair_xcol[0], nullptr);
      ev_fake_B = pr.record(0);
      if (xcol_side) pr.dep(0, 4), pr.cur_lane = 4;
      add_xcol(pr, fake_B, xcol[2]);
      pr.cur_lane = 0;
      pr.mark("fwd fake_B + idt_A done");
      pr.cur_lane = 1;
      emit_gen_forward(pr, &sink, gen[2], CGB_NET_G_BA, pair_in[1], pair_out[1], true, &pair_xcol[1], nullptr);
      ev_fake_A = pr.record(1);
      if (xcol_side) pr.dep(1, 5), pr.cur_lane = 5;
      add_xcol(pr, fake_A, xcol[3]);
      pr.cur_lane = 1;
      pr.mark("fwd fake_A + idt_B done");
    } else {
    pr.cur_lane = 0;
    emit_gen_forward(pr, &sink, gen[0], CGB_NET_G_AB, real_A, img[CGB_IMG_FAKE_B], true, &xcol[0], &xcol[2], xl0);
    ev_fake_B = pr.record(0);
    pr.mark("fwd fake_B done");
    pr.cur_lane = 1;
    emit_gen_forward(pr, &sink, gen[2], CGB_NET_G_BA, real_B, img[CGB_IMG_FAKE_A], true, &xcol[1], &xcol[3], xl1);
    ev_fake_A = pr.record(1);
    pr.mark("fwd fake_A done");
    // (identity passes: emitted below together with the rest of lanes 2 / 3)
    }
    // cycle passes
    pr.cur_lane = 0;
    emit_gen_forward(pr, &sink, gen[1], CGB_NET_G_BA, fake_B, img[CGB_IMG_REC_A], false, &xcol[2], nullptr);
    emit_gen_backward(pr, &sink, gen[1], gs[0], &real_A, cfg.lambda_A / numel_img, CGB_LOSS_CYCLE_A, GradSrc(), &dxp_img[0]);
    if (dp || adam_overlap) ev_rec_done[CGB_NET_G_BA] = pr.record(0);
    pr.mark("rec_A fwd+bwd done");
    pr.cur_lane = 1;
    emit_gen_forward(pr, &sink, gen[3], CGB_NET_G_AB, fake_A, img[CGB_IMG_REC_B], false, &xcol[3], nullptr);
    emit_gen_backward(pr, &sink, gen[3], gs[1], &real_B, cfg.lambda_B / numel_img, CGB_LOSS_CYCLE_B, GradSrc(), &dxp_img[1]);
    if (dp || adam_overlap) ev_rec_done[CGB_NET_G_AB] = pr.record(1);
    pr.mark("rec_B fwd+bwd done");
    // Lanes 2 / 3: the identity pass of one generator, the adversarial term on the fake (D frozen: input gradient
    // only, needed by the fake's backward on lane 0 / 1) and the whole D phase of one discriminator.
    // CGB_SCHED orders them: 1 = identity fwd+bwd first (default); 2 = adversarial term first, identity pass after
    // it; 3 = identity forward, adversarial term, identity backward.
    static const int sched = std::getenv("CGB_SCHED") ? std::atoi(std::getenv("CGB_SCHED")) : 1;
    auto side_lane = [&](int lane, int side, int ev_fake) {
      GenPass& GP = gen[4 + side];
      GenScratch& GS = gs[2 + side];
      const int gnet = side == 0 ? CGB_NET_G_AB : CGB_NET_G_BA;
      const TensorDesc& real_in = side == 0 ? real_B : real_A;   // idt_A = G_AB(real_B), idt_B = G_BA(real_A)
      const TensorDesc& idt_out = img[side == 0 ? CGB_IMG_IDT_A : CGB_IMG_IDT_B];
      const TensorDesc* xc = &xcol[side == 0 ? 1 : 0];
      const float idt_scale = (side == 0 ? cfg.lambda_B : cfg.lambda_A) * cfg.lambda_idt / numel_img;
      const int idt_slot = side == 0 ? CGB_LOSS_IDT_A : CGB_LOSS_IDT_B;
      const int dnet = side == 0 ? CGB_NET_D_A : CGB_NET_D_B;
      const TensorDesc& fake = side == 0 ? fake_B : fake_A;      // D_A judges domain-B images
      auto idt_fwd = [&]() { if (!pair) emit_gen_forward(pr, &sink, GP, gnet, real_in, idt_out, false, xc, nullptr); };
      auto idt_bwd = [&]() {
        if (!pair) {
          emit_gen_backward(pr, &sink, GP, GS, &real_in, idt_scale, idt_slot, GradSrc(), nullptr);
          if (dp || adam_overlap) ev_idt_done[gnet] = pr.record(lane);
          pr.mark(side == 0 ? "idt_A fwd+bwd done" : "idt_B fwd+bwd done");
        }
      };
      int ev = -1;
      auto adv = [&]() {
        pr.wait(lane, ev_fake);
        emit_dis_forward(pr, &sink, dis[side], dnet, fake);
        emit_dis_backward(pr, &sink, dis[side], ds[side], 1.f, 1.f, side == 0 ? CGB_LOSS_G_A : CGB_LOSS_G_B, false, &dx_D0[side]);
        ev = pr.record(lane);
      };
      pr.cur_lane = lane;
      if (sched == 2) {
        adv();
        idt_fwd();
        idt_bwd();
      } else if (sched == 3) {
        idt_fwd();
        adv();
        idt_bwd();
      } else {
        idt_fwd();
        idt_bwd();
        adv();
      }
      emit_dis_forward(pr, &sink, dis[2 + side], dnet, real_in);
      emit_dis_backward(pr, &sink, dis[2 + side], ds[side], 1.f, 0.5f, side == 0 ? CGB_LOSS_D_A : CGB_LOSS_D_B, true, nullptr);
      emit_dis_fake_phase(pr, &sink, side);
      pr.mark(side == 0 ? "D_A all done" : "D_B all done");
      return ev;
    };
    const int ev_dD_A = side_lane(2, 0, ev_fake_B);
    const int ev_dD_B = side_lane(3, 1, ev_fake_A);
    // both discriminators' gradients are complete: Adam(D) + bf16 refresh run here, in the shadow of the generator
    // backward chains (nothing later in the step reads the discriminator weights)
    pr.dep(3, 2);
    pr.cur_lane = 2;
    add_sum_slots(pr, CGB_GROUP_D);
    if (dp) {  // the discriminator gradients are final: first bucket
      grad_buckets.push_back({CGB_GROUP_D, 0, group_numel[CGB_GROUP_D]});
      pr.ext_event((int)grad_buckets.size() - 1);
    }
    if (with_adam_d) {
      const long long before = pr.launches;
      for (size_t i = 0; i < prog_adam[CGB_GROUP_D].ops.size(); ++i)
        pr.add(prog_adam[CGB_GROUP_D].ops[i], 0, prog_adam[CGB_GROUP_D].kinds[i]);
      pr.launches = before + prog_adam[CGB_GROUP_D].launches;
    }
    if (adam_overlap) {  // advance the generators' step counter / bias corrections once, before any bucket's Adam
      int* stp = adam_step[CGB_GROUP_G];
      float* hyp = adam_hyper[CGB_GROUP_G];
      pr.add([E, stp, hyp](cudaStream_t st) {
        cgb::adam_range(nullptr, nullptr, nullptr, nullptr, 0, E->cfg.beta1, E->cfg.beta2, E->cfg.eps, stp, hyp, E->grad_scale, true, st);
      });
      ev_gstep = pr.record(2);
    }
    pr.mark("Adam(D) done");
    // the passes that produced the fakes
    GradSrc g;
    g.fold = 3;
    g.g1 = &dx_D0[0];
    g.g2 = &dxp_img[0];
    pr.cur_lane = 0;
    pr.wait(0, ev_dD_A);
    const std::function<void(Program&, int, int)> hook_AB = bucket_hook(CGB_NET_G_AB), hook_BA = bucket_hook(CGB_NET_G_BA);
    if (pair)
      emit_gen_backward(pr, &sink, gen[0], gs[2], nullptr, 0.f, -1, g, nullptr, &real_B,
                        cfg.lambda_B * cfg.lambda_idt / numel_img, CGB_LOSS_IDT_A, &hook_AB);
    else
      emit_gen_backward(pr, &sink, gen[0], gs[0], nullptr, 0.f, -1, g, nullptr, nullptr, 0.f, -1, &hook_AB);
    pr.mark("bwd fake_B done");
    g.g1 = &dx_D0[1];
    g.g2 = &dxp_img[1];
    pr.cur_lane = 1;
    pr.wait(1, ev_dD_B);
    if (pair)
      emit_gen_backward(pr, &sink, gen[2], gs[3], nullptr, 0.f, -1, g, nullptr, &real_A,
                        cfg.lambda_A * cfg.lambda_idt / numel_img, CGB_LOSS_IDT_B, &hook_BA);
    else
      emit_gen_backward(pr, &sink, gen[2], gs[1], nullptr, 0.f, -1, g, nullptr, nullptr, 0.f, -1, &hook_BA);
    pr.mark("bwd fake_A done");
    pr.join();
    pr.cur_lane = 0;
    if (adam_overlap) {  // the chains are done with the lowest layers: refresh the last bucket of each generator
      add_pack_layers(pr, CGB_NET_G_AB, pend_lo[0], pend_hi[0]);
      add_pack_layers(pr, CGB_NET_G_BA, pend_lo[1], pend_hi[1]);
    }
    add_sum_slots(pr, CGB_GROUP_G);
    if (dp && (fp32 || n_dp_buckets <= 1)) {  // one bucket: the whole generator group, at the end
      GradBucket gb{CGB_GROUP_G, 0, group_numel[CGB_GROUP_G]};
      gb.order = 1;
      grad_buckets.push_back(gb);
      pr.ext_event((int)grad_buckets.size() - 1);
    }
    pr.mark("step end (before Adam)");
  };
  record_step(prog_step, true);
  record_step(prog_step_dp, false);
  grad_events.resize(grad_buckets.size());
  for (cudaEvent_t& ev : grad_events) CGB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  prog_step_dp.ext_events = grad_events.data();
  }  // !infer_only
  {  // both optimisers side by side
    Program& pr = prog_adams;
    pr.fork(2);
    for (int g = 0; g < 2; ++g) {
      pr.cur_lane = g;
      for (size_t i = 0; i < prog_adam[g].ops.size(); ++i) pr.add(prog_adam[g].ops[i], 0, prog_adam[g].kinds[i]);
    }
    pr.launches = prog_adam[0].launches + prog_adam[1].launches;
    pr.join(2);
    pr.cur_lane = 0;
  }
  const bool lite = std::getenv("CGB_STEM_GEMM") == nullptr && std::getenv("CGB_XCOL_ON_CHAIN") == nullptr;
  segments[CGB_SEG_STEP].seq = {lite ? &prog_set_inputs_lite : &prog_set_inputs, &prog_step};
  if (!step_has_adam_g) segments[CGB_SEG_STEP].seq.push_back(&prog_adam[CGB_GROUP_G]);
  segments[CGB_SEG_G].seq = {&prog_set_inputs, &prog_cycle, &prog_G};
  segments[CGB_SEG_D].seq = {&prog_D};
  segments[CGB_SEG_ADAM_G].seq = {&prog_adam[0]};
  segments[CGB_SEG_ADAM_D].seq = {&prog_adam[1]};
  segments[CGB_SEG_FORWARD].seq = {&prog_set_inputs, &prog_cycle};
  segments[CGB_SEG_STEP_NOOPT].seq = {lite ? &prog_set_inputs_lite : &prog_set_inputs, &prog_step_dp};
  // ---- module-level forward programs (Generator.forward / Discriminator.forward)
  for (int net = 0; net < 2; ++net) {
    emit_gen_forward(prog_mod_gen[net], &dummy_flops, gen[6], net, mod_in, mod_out, false, nullptr, nullptr);
    emit_dis_forward(prog_mod_dis[net], &dummy_flops, dis[4], 2 + net, mod_in);
  }
}
