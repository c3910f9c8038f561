// The CycleGAN step engine: parameter inventory, HBM workspace layout, convolution plans and the
// recorded launch programs for the generator / discriminator phases.
// Mirrors CycleGANTrainer in the stand-in (oracle/cyclegan_standin.py:195-330).
#pragma once
#include <deque>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../../include/cyclegan_b200.h"
#include "conv_plan.h"
#include "pointwise.h"

namespace cgb {

struct LayerParam {
  std::string name;  // "stem", "res.3.conv1", "conv0", ...
  ConvSpec spec;
  bool has_in = false;  // followed by InstanceNorm (its bias is mathematically dead and is skipped)
  int net = 0, group = 0;
  long long w_off = 0, b_off = 0;    // element offsets in the group's flat fp32 buffers
  long long wf_off = 0, wt_off = 0;  // element offsets in the group's bf16 pack arena
};

typedef std::function<void(cudaStream_t)> Op;
enum OpKind { kOpOther = 0, kOpIgemm = 1, kOpWgradTc = 2, kOpWgradDirect = 3, kOpNorm = 4, kOpMemset = 5, kNumOpKinds = 6 };
struct Program {
  std::vector<Op> ops;
  std::vector<int> kinds;
  std::vector<double> flops;
  long long launches = 0;
  void add(Op op, int n_launches = 1, int kind = kOpOther, double fl = 0.0) {
    ops.push_back(std::move(op));
    kinds.push_back(kind);
    flops.push_back(fl);
    launches += n_launches;
  }
  void run(cudaStream_t st) const {
    for (const Op& op : ops) op(st);
  }
  // replay only the ops of one kind (profiling; data dependencies are ignored on purpose)
  void run_kind(int kind, cudaStream_t st, long long* count, double* fl) const {
    for (size_t i = 0; i < ops.size(); ++i)
      if (kinds[i] == kind) {
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
    t.ptr = static_cast<bf16*>(alloc((size_t)t.elems() * sizeof(bf16)));
    return t;
  }
};

struct GenPass {  // activations of one generator forward pass, kept for its backward
  int net = 0;
  TensorDesc in, out;  // image tensors (not owned)
  TensorDesc y_stem, a_stem, y_d1, a_d1, y_d2, y_u1, a_u1, y_u2, a_u2p;
  std::vector<TensorDesc> xp, y1, bp, y2;
  float2* stats = nullptr;   // forward IN statistics of all 5 + 2*nb layers, [layer][N][C]
  float2* bstats = nullptr;  // backward reductions, same layout
  size_t stats_bytes = 0;
  std::vector<long long> stat_off;  // float2 offset per IN layer (0 stem, 1 d1, 2 d2, 3+2k, 4+2k, u1, u2)
};

struct DisPass {
  int net = 0;
  TensorDesc in;  // image (not owned)
  TensorDesc l0, y1, a1, y2, a2, y3, a3, logits;
  float2* stats = nullptr;
  float2* bstats = nullptr;
  size_t stats_bytes = 0;
  long long stat_off[3] = {0, 0, 0};
};

}  // namespace cgb

struct cgb_engine {
  cgb_config_t cfg;
  int sm_count = 148;
  bool bound = false;
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
  float* losses = nullptr;                 // CGB_NUM_LOSSES + padding
  int* adam_step[2] = {nullptr, nullptr};
  float* adam_hyper[2] = {nullptr, nullptr};
  std::vector<cgb::GenPass> gen;  // 6 training passes + 1 module-forward pass
  std::vector<cgb::DisPass> dis;  // 4 training passes + 1 module-forward pass
  // generator backward scratch
  cgb::TensorDesc dpre_head, dxp_head, dyF, dxF, dyH, dxH, dyQ, GQ[2], dbpQ, dxpQ, dxp_img[2], dx_D0[2];
  // discriminator backward scratch
  cgb::TensorDesc dlogits, dx3, dy3, dx2, dy2, dx1, dy1, dx0, dpre0;

  // library-owned small tables
  void* meta = nullptr;
  size_t meta_cap = 0, meta_off = 0;
  cgb::PackEntry* pack_table[2] = {nullptr, nullptr};
  int pack_count[2] = {0, 0};
  int pack_max[2] = {0, 0};

  std::deque<cgb::IgemmPlan> igemm_plans;
  std::deque<cgb::WgradPlan> wgrad_plans;
  cgb::Program prog_set_inputs, prog_cycle, prog_G, prog_D, prog_adam[2], prog_refresh[2];
  cgb::Program prog_mod_gen[2], prog_mod_dis[2];
  double conv_flops = 0;  // accumulated while recording prog_cycle/prog_G/prog_D

  cudaGraphExec_t graph = nullptr;
  bool graph_failed = false;
  int step_calls = 0;

  ~cgb_engine();
  void build_inventory();
  void layout(cgb::Arena& A);
  void record_programs();
  void* meta_upload(const void* src, size_t bytes);
};
