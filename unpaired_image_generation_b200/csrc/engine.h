// The CycleGAN step engine: parameter inventory, HBM workspace layout, convolution plans and the
// recorded launch programs for the generator / discriminator phases.
// Mirrors CycleGANTrainer in the stand-in (oracle/cyclegan_standin.py:278-400).
#pragma once
#include <deque>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../../include/cyclegan_b200.h"
#include "conv_plan.h"
#include "fp32_path.h"
#include "pointwise.h"
#include "small_wgrad.h"

namespace cgb {

struct LayerParam {
  std::string name;  // "stem", "res.3.conv1", "conv0", ...
  ConvSpec spec;
  bool has_in = false;  // followed by InstanceNorm (its bias is mathematically dead and is skipped)
  int net = 0, group = 0;
  long long w_off = 0, b_off = 0;    // element offsets in the group's flat fp32 buffers
  long long wf_off = 0, wt_off = 0;  // element offsets in the group's bf16 pack arena
  long long wx_off = -1;             // stem only: Wx[Cout][256] for the GEMM over the shared im2col4 matrix
};

typedef std::function<void(cudaStream_t)> Op;
enum OpKind { kOpOther = 0, kOpIgemm = 1, kOpWgradTc = 2, kOpWgradSmall = 3, kOpNorm = 4, kOpMemset = 5, kOpDep = 6, kOpMarker = 7, kOpRecord = 8, kOpWait = 9, kOpExtEvent = 10, kNumOpKinds = 11 };
constexpr int kLanes = 12;  // parallel graph branches: lanes 0-3 carry independent passes, lanes l+4 and l+8 the weight
                           // gradients of the pass on lane l (wgrad runs beside the dgrad of the same layer)
constexpr int kPassLanes = 4;

// A recorded launch sequence.  Each op belongs to a lane; `dep(a, b)` makes everything issued later on lane b
// wait for everything issued so far on lane a.  Run on ONE stream the recorded order is already a valid
// serialisation (eager mode); captured into a CUDA graph, lanes become parallel branches.
struct Program {
  std::vector<Op> ops;
  std::vector<int> kinds, lanes, dep_from;
  std::vector<double> flops;
  std::vector<std::string> names;  // per op (profiling)
  std::string cur_name;            // label attached to the ops added next
  long long launches = 0;
  int cur_lane = 0;
  void add(Op op, int n_launches = 1, int kind = kOpOther, double fl = 0.0) {
    ops.push_back(std::move(op));
    kinds.push_back(kind);
    lanes.push_back(cur_lane);
    dep_from.push_back(-1);
    flops.push_back(fl);
    names.push_back(cur_name);
    launches += n_launches;
  }
  void dep(int from, int to) {
    ops.push_back(Op());
    kinds.push_back(kOpDep);
    lanes.push_back(to);
    dep_from.push_back(from);
    flops.push_back(0.0);
    names.push_back("");
  }
  // timeline marker (profiling): an external event record on the current lane when `timeline` is set
  std::vector<std::string> labels;
  void mark(const std::string& label) {
    ops.push_back(Op());
    kinds.push_back(kOpMarker);
    lanes.push_back(cur_lane);
    dep_from.push_back((int)labels.size());
    flops.push_back(0.0);
    names.push_back("");
    labels.push_back(label);
  }
  // split form of dep(): `int e = record(a)` ... later ... `wait(b, e)`: lane b waits only for what lane a had
  // issued at the record point
  int n_records = 0;
  int record(int lane) {
    ops.push_back(Op());
    kinds.push_back(kOpRecord);
    lanes.push_back(lane);
    dep_from.push_back(n_records);
    flops.push_back(0.0);
    names.push_back("");
    return n_records++;
  }
  void wait(int lane, int record_id) {
    ops.push_back(Op());
    kinds.push_back(kOpWait);
    lanes.push_back(lane);
    dep_from.push_back(record_id);
    flops.push_back(0.0);
    names.push_back("");
  }
  // external event: recorded on the current lane at this point of the program; streams OUTSIDE the program (the
  // data-parallel communication stream) wait on it while the rest of the program keeps running.  In a captured graph
  // it becomes an event-record node (cudaEventRecordExternal), in eager mode a plain cudaEventRecord.
  cudaEvent_t* ext_events = nullptr;  // owned by the engine
  void ext_event(int id) {
    ops.push_back(Op());
    kinds.push_back(kOpExtEvent);
    lanes.push_back(cur_lane);
    dep_from.push_back(id);
    flops.push_back(0.0);
    names.push_back("");
  }
  void fork(int n = kLanes) {
    for (int l = 1; l < n; ++l) dep(0, l);
  }
  void join(int n = kLanes) {
    for (int l = 1; l < n; ++l) dep(l, 0);
  }
  void run(cudaStream_t st) const {
    for (size_t i = 0; i < ops.size(); ++i) {
      if (kinds[i] < kOpDep) ops[i](st);
      else if (kinds[i] == kOpExtEvent && ext_events) CGB_CUDA(cudaEventRecord(ext_events[dep_from[i]], st));
    }
  }
  // lanes[l] are distinct streams (lane 0 = the caller's); events: one per dep op, created by the caller
  struct Mark {
    std::string label;
    int lane;
    cudaEvent_t ev;
  };
  void run_lanes(cudaStream_t* lane_streams, std::vector<cudaEvent_t>& events, size_t* next_event,
                 std::vector<Mark>* timeline = nullptr, bool mark_ops = false) const {
    std::vector<cudaEvent_t> rec(n_records, nullptr);
    for (size_t i = 0; i < ops.size(); ++i) {
      if (kinds[i] == kOpRecord || kinds[i] == kOpWait) {
        if (kinds[i] == kOpRecord) {
          if (*next_event >= events.size()) {
            cudaEvent_t ev;
            CGB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            events.push_back(ev);
          }
          rec[dep_from[i]] = events[(*next_event)++];
          CGB_CUDA(cudaEventRecord(rec[dep_from[i]], lane_streams[lanes[i]]));
        } else {
          CGB_CUDA(cudaStreamWaitEvent(lane_streams[lanes[i]], rec[dep_from[i]], 0));
        }
      } else if (kinds[i] == kOpExtEvent) {
        if (ext_events)
          CGB_CUDA(cudaEventRecordWithFlags(ext_events[dep_from[i]], lane_streams[lanes[i]], cudaEventRecordExternal));
      } else if (kinds[i] == kOpMarker) {
        if (timeline) {
          Mark m;
          m.label = labels[dep_from[i]];
          m.lane = lanes[i];
          CGB_CUDA(cudaEventCreate(&m.ev));
          CGB_CUDA(cudaEventRecordWithFlags(m.ev, lane_streams[lanes[i]], cudaEventRecordExternal));
          timeline->push_back(m);
        }
      } else if (kinds[i] == kOpDep) {
        if (*next_event >= events.size()) {
          cudaEvent_t ev;
          CGB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
          events.push_back(ev);
        }
        cudaEvent_t ev = events[(*next_event)++];
        CGB_CUDA(cudaEventRecord(ev, lane_streams[dep_from[i]]));
        CGB_CUDA(cudaStreamWaitEvent(lane_streams[lanes[i]], ev, 0));
      } else {
        ops[i](lane_streams[lanes[i]]);
        if (timeline && mark_ops) {  // hang probe: an external event after every op tells which op never finished
          Mark m;
          m.label = "#" + std::to_string(i) + " " + names[i];
          m.lane = lanes[i];
          CGB_CUDA(cudaEventCreateWithFlags(&m.ev, cudaEventDisableTiming));
          CGB_CUDA(cudaEventRecordWithFlags(m.ev, lane_streams[lanes[i]], cudaEventRecordExternal));
          timeline->push_back(m);
        }
      }
    }
  }
  // replay only the ops of one kind (profiling; data dependencies are ignored on purpose)
  void run_kind(int kind, cudaStream_t st, long long* count, double* fl) const {
    for (size_t i = 0; i < ops.size(); ++i)
      if (kinds[i] == kind && kind < kOpDep) {
        ops[i](st);
        if (count) ++*count;
        if (fl) *fl += flops[i];
      }
  }
};

// bump allocator over the caller's workspace (base == nullptr: measuring pass)
struct Arena {
  uint8_t* base = nullptr;
  size_t off = 0;
  int esz = 2;  // element size of the activation tensors carved from this arena (4: fp32 validation mode)
  void* alloc(size_t bytes) {
    off = (off + 1023) & ~size_t(1023);
    void* p = reinterpret_cast<void*>(reinterpret_cast<uintptr_t>(base) + off);
    off += bytes;
    return p;
  }
  TensorDesc tensor(int N, int H, int W, int C, int halo) {
    TensorDesc t;
    t.N = N;
    t.H = H;
    t.W = W;
    t.C = C;
    t.halo = halo;
    t.esz = esz;
    t.ptr = static_cast<bf16*>(alloc(t.bytes()));
    return t;
  }
};

struct GenPass {  // activations of one generator forward pass, kept for its backward
  int net = 0;
  int gslot = 0;  // fp32 validation mode: which per-pass weight-gradient buffer this pass writes (cgb_engine::gslot)
  TensorDesc in, out;  // image tensors (not owned)
  const TensorDesc* xcol = nullptr;  // shared im2col4 matrix of `in` (stem fprop / wgrad as GEMMs), or null
  TensorDesc y_stem, a_stem, y_d1, a_d1, y_d2, y_u1, a_u1, y_u2, a_u2p;
  std::vector<TensorDesc> xp, y1, bp, y2;
  float2* stats = nullptr;   // forward IN statistics of all 5 + 2*nb layers, [layer][N][C]
  float2* bstats = nullptr;  // backward reductions, same layout
  size_t stats_bytes = 0;
  std::vector<long long> stat_off;  // float2 offset per IN layer (0 stem, 1 d1, 2 d2, 3+2k, 4+2k, u1, u2)
};

struct DisPass {
  int net = 0;
  int gslot = 0;
  TensorDesc in;  // image (not owned)
  TensorDesc l0, y1, a1, y2, a2, y3, a3, logits;
  float2* stats = nullptr;
  float2* bstats = nullptr;
  size_t stats_bytes = 0;
  long long stat_off[3] = {0, 0, 0};
};

struct GenScratch {  // backward scratch of one generator pass (one set per lane)
  TensorDesc dpre_head, dxp_head, dyF, dxF, dyH, dxH, dyQ, dyQ2, GQ[2], dbpQ, dxpQ;  // dyQ / dyQ2 alternate along the residual chain
  bf16* colbuf = nullptr;  // im2col scratch of the 3-channel operands (stem / head weight gradients)
  size_t colbuf_elems = 0;
};
struct DisScratch {
  TensorDesc dlogits, dx3, dy3, dx2, dy2, dx1, dy1, dx0, dpre0;
  bf16* colbuf = nullptr;  // conv0 / conv4 weight gradients
  size_t colbuf_elems = 0;
};

}  // namespace cgb

struct cgb_engine {
  cgb_config_t cfg;
  int sm_count = 148;
  bool bound = false;
  bool infer_only = false;  // CGB_FLAG_INFERENCE: module forwards only
  // CGB_FLAG_FP32_VALIDATE: fp32 activations, fp64 accumulation, deterministic kernels (fp32_path.h); same programs.
  // Weight gradients are WRITTEN into one flat buffer per pass kind (generators: fake / rec / idt pass; discriminators:
  // real / fake pass) and summed in a fixed order at the end of each phase, so no float accumulation order depends
  // on how the lanes interleave.
  bool fp32 = false;
  float* gslot[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  float* grad_base(int group, int slot) const { return fp32 ? gslot[group][slot] : G[group]; }
  int pool_size = 0;        // image history pool (cgb_engine_set_image_pool): 0 = off (D sees the current fakes)
  cgb::TensorDesc pool_img[2], pool_din[2];  // [0]: fake_B history (for D_A), [1]: fake_A history (for D_B)
  int* pool_dec = nullptr;  // device [2][batch][2]: (store, ret) per image, written by cgb_set_pool_decisions
  float grad_scale = 1.f;

  std::vector<cgb::LayerParam> layers[4];
  long long group_numel[2] = {0, 0};
  long long pack_elems[2] = {0, 0};
  size_t workspace_bytes = 0;

  // caller-owned flat buffers
  float* P[2] = {nullptr, nullptr};
  float* G[2] = {nullptr, nullptr};
  float* M[2] = {nullptr, nullptr};
  float* V[2] = {nullptr, nullptr};

  // workspace carve-up
  cgb::bf16* pack[2] = {nullptr, nullptr};
  cgb::TensorDesc img[8];
  cgb::TensorDesc mod_in, mod_out;
  float* staging[2] = {nullptr, nullptr};  // fp32 NCHW inputs
  unsigned char* staging_u8[2] = {nullptr, nullptr};  // uint8 HWC inputs (cgb_stage_inputs_u8)
  float* losses = nullptr;                 // CGB_NUM_LOSSES + padding
  int* adam_step[2] = {nullptr, nullptr};
  float* adam_hyper[2] = {nullptr, nullptr};
  std::vector<cgb::GenPass> gen;  // 6 training passes + 1 module-forward pass
  std::vector<cgb::DisPass> dis;  // 4 training passes + 1 module-forward pass + 2 passes on the pool's output
  cgb::GenScratch gs[cgb::kPassLanes];
  cgb::DisScratch ds[2];
  cgb::TensorDesc dxp_img[2], dx_D0[2];  // gradients w.r.t. the fake images (from the cycle passes / from D)
  cgb::TensorDesc xcol[4];               // im2col4 (7x7 taps x 4 channels -> 256 columns) of real_A, real_B, fake_B, fake_A
  // paired schedule (default: on from batch 4 up; CGB_PAIR=0 / 1 overrides): the two passes that share a generator AND
  // only need the real images -- fake_B = G_AB(real_A) with idt_A = G_AB(real_B), fake_A = G_BA(real_B) with
  // idt_B = G_BA(real_A) -- run as ONE pass of batch 2N.  [0] = G_AB pair, [1] = G_BA pair.
  bool pair = false;
  cgb::TensorDesc reals3, xcol3;          // [real_A; real_B; real_A] and its im2col4
  cgb::TensorDesc pair_in[2], pair_out[2], pair_xcol[2];

  // library-owned small tables
  void* meta = nullptr;
  size_t meta_cap = 0, meta_off = 0;
  cgb::PackEntry* pack_table[2] = {nullptr, nullptr};
  int pack_count[2] = {0, 0};
  int pack_max[2] = {0, 0};

  std::deque<cgb::IgemmPlan> igemm_plans;
  std::deque<cgb::WgradPlan> wgrad_plans;
  std::deque<cgb::SmallWgradPlan> small_wgrad_plans;
  cgb::Program prog_set_inputs_lite;  // staging -> images only (the merged step builds the im2col4 matrices on side lanes)
  cgb::Program prog_set_inputs, prog_cycle, prog_G, prog_D, prog_adam[2], prog_refresh[2];
  cgb::Program prog_step;   // forward + G phase + D phase as ONE schedule (no joins between the phases)
  bool step_has_adam_g = false;  // prog_step runs the generators' Adam + bf16 refresh itself, bucket by bucket
  cgb::Program prog_step_dp;  // the same without Adam(D): data-parallel callers all-reduce the gradients first
  cgb::Program prog_adams;  // both optimisers side by side
  cgb::Program prog_mod_gen[2], prog_mod_dis[2];
  double conv_flops = 0;  // accumulated while recording prog_cycle/prog_G/prog_D

  // Data parallel: the gradient buffers become final bucket by bucket while the merged step (prog_step_dp) is still
  // running; each bucket has an external event the caller's communication stream waits on (cgb_wait_grad_bucket).
  struct GradBucket {
    int group;
    long long offset, numel;  // range of the group's flat gradient buffer
    int net = -1, layer_lo = 0, layer_hi = 0;  // the layers [lo, hi) of one network it covers (net < 0: the whole group)
    int order = 0;            // sort key: buckets of the two generators alternate (their backward chains run side by side)
  };
  std::vector<GradBucket> grad_buckets;  // in the order they become ready
  std::vector<cudaEvent_t> grad_events;  // one per bucket
  cudaStream_t lane_streams[cgb::kLanes] = {};  // [0] = caller's stream
  std::vector<cudaEvent_t> events;
  std::vector<cudaEvent_t> seg_events[CGB_NUM_SEGMENTS];
  struct Segment {  // a graph-replayable sequence of programs
    std::vector<const cgb::Program*> seq;
    cudaGraphExec_t exec = nullptr;
    bool failed = false;
    int calls = 0;
  };
  Segment segments[CGB_NUM_SEGMENTS];
  void run_segment(int seg, cudaStream_t st);
  void drop_graphs();
  std::string timeline(cudaStream_t st);
  std::string hang_probe(cudaStream_t st, int steps, int stall_ms, int fine);
  std::string profile_ops(cudaStream_t st, int reps);

  ~cgb_engine();
  void build_inventory();
  void layout(cgb::Arena& A);
  void record_programs();
  void* meta_upload(const void* src, size_t bytes);
};
