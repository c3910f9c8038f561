// extern "C" surface of libcyclegan_b200.so (declared in include/cyclegan_b200.h).
#include <cstdio>
#include <cstring>
#include <vector>

#include "engine.h"

using namespace cgb;

namespace cgb {
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
}  // namespace cgb

#define CGB_API_BEGIN try {
#define CGB_API_END                       \
  }                                       \
  catch (const std::exception& ex) {      \
    cgb::set_last_error(ex.what());       \
    return 1;                             \
  }                                       \
  catch (...) {                           \
    cgb::set_last_error("unknown error"); \
    return 2;                             \
  }                                       \
  return 0;

static cudaStream_t S(void* stream) { return static_cast<cudaStream_t>(stream); }

extern "C" {

const char* cgb_last_error(void) { return cgb::g_last_error.c_str(); }
int cgb_version(void) { return 100; }

int cgb_engine_create(const cgb_config_t* cfg, cgb_engine_t** out) { return cgb_engine_create_ex(cfg, 0, out); }

int cgb_engine_create_ex(const cgb_config_t* cfg, int flags, cgb_engine_t** out) {
  CGB_API_BEGIN
  CGB_CHECK(cfg && out, "null argument");
  CGB_CHECK(cfg->batch >= 1, "batch must be >= 1");
  CGB_CHECK(cfg->size >= 32 && cfg->size % 8 == 0, "size must be a multiple of 8 and >= 32");
  CGB_CHECK(cfg->n_blocks >= 1 && cfg->n_blocks <= 32, "n_blocks must be in [1, 32]");
  cgb_engine* e = new cgb_engine();
  e->cfg = *cfg;
  e->infer_only = (flags & CGB_FLAG_INFERENCE) != 0;
  e->fp32 = (flags & CGB_FLAG_FP32_VALIDATE) != 0;
  e->build_inventory();
  Arena A;
  e->layout(A);
  e->workspace_bytes = A.off;
  *out = e;
  CGB_API_END
}

void cgb_engine_destroy(cgb_engine_t* e) { delete e; }

int cgb_num_params(const cgb_engine_t* e, int net) {
  if (!e || net < 0 || net > 3) return -1;
  return (int)e->layers[net].size() * 2;
}

int cgb_param_info(const cgb_engine_t* e, int net, int index, cgb_param_info_t* out) {
  CGB_API_BEGIN
  CGB_CHECK(e && out && net >= 0 && net < 4, "bad argument");
  CGB_CHECK(index >= 0 && index < (int)e->layers[net].size() * 2, "parameter index out of range");
  const LayerParam& p = e->layers[net][index / 2];
  std::memset(out, 0, sizeof(*out));
  const bool is_bias = index % 2 == 1;
  const std::string nm = p.name + (is_bias ? ".bias" : ".weight");
  std::strncpy(out->name, nm.c_str(), sizeof(out->name) - 1);
  out->is_bias = is_bias;
  out->transposed = p.spec.transposed;
  out->cout = p.spec.Cout;
  out->cin = p.spec.Cin;
  out->k = p.spec.k;
  out->offset = is_bias ? p.b_off : p.w_off;
  out->numel = is_bias ? p.spec.Cout : (long long)p.spec.Cout * p.spec.taps() * p.spec.Cin;
  CGB_API_END
}

long long cgb_group_numel(const cgb_engine_t* e, int group) { return (e && group >= 0 && group < 2) ? e->group_numel[group] : -1; }
long long cgb_workspace_bytes(const cgb_engine_t* e) { return e ? (long long)e->workspace_bytes : -1; }

int cgb_engine_bind(cgb_engine_t* e, float* pG, float* gG, float* mG, float* vG, float* pD, float* gD, float* mD,
                    float* vD, void* workspace, long long workspace_bytes) {
  CGB_API_BEGIN
  CGB_CHECK(e, "null engine");
  CGB_CHECK(!e->bound, "engine is already bound");
  CGB_CHECK(pG && gG && mG && vG && pD && gD && mD && vD && workspace, "null buffer");
  CGB_CHECK(workspace_bytes >= (long long)e->workspace_bytes, "workspace too small");
  CGB_CHECK((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "workspace must be 1024-byte aligned");
  int dev = 0;
  CGB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CGB_CUDA(cudaGetDeviceProperties(&prop, dev));
  CGB_CHECK(prop.major == 10, "this library is built for sm_100a (B200) only; found compute capability " +
                                  std::to_string(prop.major) + "." + std::to_string(prop.minor));
  e->sm_count = prop.multiProcessorCount;
  e->P[0] = pG; e->G[0] = gG; e->M[0] = mG; e->V[0] = vG;
  e->P[1] = pD; e->G[1] = gD; e->M[1] = mD; e->V[1] = vD;
  Arena A;
  A.base = static_cast<uint8_t*>(workspace);
  e->layout(A);
  CGB_CUDA(cudaMemset(workspace, 0, e->workspace_bytes));
  e->meta_cap = 32u << 20;
  CGB_CUDA(cudaMalloc(&e->meta, e->meta_cap));
  if (e->pool_dec) CGB_CUDA(cudaMemset(e->pool_dec, 0xFF, (size_t)2 * e->cfg.batch * 2 * sizeof(int)));  // (-1, -1): pass-through
  e->record_programs();
  for (int g = 0; g < 2; ++g) CGB_CUDA(cudaMemcpy(e->adam_hyper[g] + 2, &e->cfg.lr, sizeof(float), cudaMemcpyHostToDevice));
  e->bound = true;
  e->prog_refresh[0].run(0);
  e->prog_refresh[1].run(0);
  CGB_CUDA(cudaDeviceSynchronize());
  CGB_API_END
}

int cgb_refresh_weights(cgb_engine_t* e, int group, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && group >= 0 && group < 2, "bad argument / engine not bound");
  e->prog_refresh[group].run(S(stream));
  CGB_API_END
}

int cgb_set_grad_scale(cgb_engine_t* e, float scale) {
  if (!e) return 1;
  e->grad_scale = scale;
  e->drop_graphs();  // the scale is a baked kernel argument: re-capture
  return 0;
}

int cgb_set_step_count(cgb_engine_t* e, int group, int step) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && group >= 0 && group < 2, "bad argument / engine not bound");
  CGB_CUDA(cudaMemcpy(e->adam_step[group], &step, sizeof(int), cudaMemcpyHostToDevice));
  CGB_API_END
}

int cgb_get_step_count(cgb_engine_t* e, int group, int* step_out) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && group >= 0 && group < 2 && step_out, "bad argument / engine not bound");
  CGB_CUDA(cudaDeviceSynchronize());
  CGB_CUDA(cudaMemcpy(step_out, e->adam_step[group], sizeof(int), cudaMemcpyDeviceToHost));
  CGB_API_END
}

int cgb_generator_forward(cgb_engine_t* e, int net, const float* x, float* y, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && (net == 0 || net == 1) && x && y, "bad argument / engine not bound");
  nchw_to_nhwc(x, 3, e->mod_in, S(stream));
  e->prog_mod_gen[net].run(S(stream));
  nhwc_to_nchw(e->mod_out, 3, y, S(stream));
  CGB_API_END
}

int cgb_discriminator_forward(cgb_engine_t* e, int net, const float* x, float* logits, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && (net == 2 || net == 3) && x && logits, "bad argument / engine not bound");
  nchw_to_nhwc(x, 3, e->mod_in, S(stream));
  e->prog_mod_dis[net - 2].run(S(stream));
  nhwc_to_nchw(e->dis[4].logits, 1, logits, S(stream));
  CGB_API_END
}

int cgb_set_inputs(cgb_engine_t* e, const float* real_A, const float* real_B, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && real_A && real_B, "bad argument / engine not bound");
  CGB_CHECK(!e->infer_only, "inference-only engine: training entry points are unavailable");
  const size_t bytes = (size_t)e->cfg.batch * 3 * e->cfg.size * e->cfg.size * sizeof(float);
  CGB_CUDA(cudaMemcpyAsync(e->staging[0], real_A, bytes, cudaMemcpyDefault, S(stream)));
  CGB_CUDA(cudaMemcpyAsync(e->staging[1], real_B, bytes, cudaMemcpyDefault, S(stream)));
  e->prog_set_inputs.run(S(stream));
  CGB_API_END
}

int cgb_forward_cycle(cgb_engine_t* e, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound, "engine not bound");
  CGB_CHECK(!e->infer_only, "inference-only engine: training entry points are unavailable");
  e->prog_cycle.run(S(stream));
  CGB_API_END
}

int cgb_get_image(cgb_engine_t* e, int which, float* out, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && which >= 0 && which < 10 && out, "bad argument / engine not bound");
  CGB_CHECK(!e->infer_only, "inference-only engine: training entry points are unavailable");
  if (which >= CGB_IMG_POOL_FAKE_B) {
    CGB_CHECK(e->pool_size > 0, "the image pool is not enabled (cgb_engine_set_image_pool)");
    nhwc_to_nchw(e->pool_din[which - CGB_IMG_POOL_FAKE_B], 3, out, S(stream));
  } else {
    nhwc_to_nchw(e->img[which], 3, out, S(stream));
  }
  CGB_API_END
}

int cgb_get_image_u8(cgb_engine_t* e, int which, unsigned char* out, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && which >= 0 && which < 10 && out, "bad argument / engine not bound");
  CGB_CHECK(!e->infer_only, "inference-only engine: training entry points are unavailable");
  if (which >= CGB_IMG_POOL_FAKE_B) {
    CGB_CHECK(e->pool_size > 0, "the image pool is not enabled (cgb_engine_set_image_pool)");
    nhwc_to_u8hwc(e->pool_din[which - CGB_IMG_POOL_FAKE_B], 3, out, S(stream));
  } else {
    nhwc_to_u8hwc(e->img[which], 3, out, S(stream));
  }
  CGB_API_END
}

int cgb_phase_generators(cgb_engine_t* e, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound, "engine not bound");
  CGB_CHECK(!e->infer_only, "inference-only engine: training entry points are unavailable");
  e->prog_cycle.run(S(stream));
  e->prog_G.run(S(stream));
  CGB_API_END
}

int cgb_phase_discriminators(cgb_engine_t* e, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound, "engine not bound");
  CGB_CHECK(!e->infer_only, "inference-only engine: training entry points are unavailable");
  e->prog_D.run(S(stream));
  CGB_API_END
}

int cgb_adam(cgb_engine_t* e, int group, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && group >= 0 && group < 2, "bad argument / engine not bound");
  CGB_CHECK(!e->infer_only, "inference-only engine: training entry points are unavailable");
  e->prog_adam[group].run(S(stream));
  CGB_API_END
}

int cgb_adam_range(cgb_engine_t* e, int group, long long offset, long long numel, int advance_step, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && group >= 0 && group < 2, "bad argument / engine not bound");
  CGB_CHECK(!e->infer_only, "inference-only engine: training entry points are unavailable");
  CGB_CHECK(offset >= 0 && numel >= 0 && offset % 4 == 0 && offset + numel <= e->group_numel[group],
            "range must lie inside the group and start at a multiple of 4 elements");
  adam_range(e->P[group] + offset, e->G[group] + offset, e->M[group] + offset, e->V[group] + offset, numel, e->cfg.beta1,
             e->cfg.beta2, e->cfg.eps, e->adam_step[group], e->adam_hyper[group], e->grad_scale, advance_step != 0,
             S(stream));
  CGB_API_END
}

int cgb_train_step(cgb_engine_t* e, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound, "engine not bound");
  CGB_CHECK(!e->infer_only, "inference-only engine: training entry points are unavailable");
  e->run_segment(CGB_SEG_STEP, S(stream));
  CGB_API_END
}

int cgb_run_segment(cgb_engine_t* e, int segment, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && segment >= 0 && segment < CGB_NUM_SEGMENTS, "bad argument / engine not bound");
  CGB_CHECK(!e->infer_only, "inference-only engine: training entry points are unavailable");
  e->run_segment(segment, S(stream));
  CGB_API_END
}

int cgb_num_grad_buckets(const cgb_engine_t* e) { return (e && e->bound) ? (int)e->grad_buckets.size() : -1; }

int cgb_grad_bucket_info(const cgb_engine_t* e, int index, int* group, long long* offset, long long* numel) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && index >= 0 && index < (int)e->grad_buckets.size() && group && offset && numel,
            "bad argument / engine not bound");
  *group = e->grad_buckets[index].group;
  *offset = e->grad_buckets[index].offset;
  *numel = e->grad_buckets[index].numel;
  CGB_API_END
}

int cgb_grad_bucket_layers(const cgb_engine_t* e, int index, int* net, int* layer_lo, int* layer_hi, int* order) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && index >= 0 && index < (int)e->grad_buckets.size() && net && layer_lo && layer_hi && order,
            "bad argument / engine not bound");
  *net = e->grad_buckets[index].net;
  *layer_lo = e->grad_buckets[index].layer_lo;
  *layer_hi = e->grad_buckets[index].layer_hi;
  *order = e->grad_buckets[index].order;
  CGB_API_END
}

int cgb_refresh_weights_layers(cgb_engine_t* e, int net, int layer_lo, int layer_hi, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && net >= 0 && net < 4, "bad argument / engine not bound");
  const int nl = (int)e->layers[net].size();
  CGB_CHECK(layer_lo >= 0 && layer_lo <= layer_hi && layer_hi <= nl, "layer range out of bounds");
  const int g = net < 2 ? CGB_GROUP_G : CGB_GROUP_D;
  const int first = (net - 2 * g) * nl + layer_lo, count = layer_hi - layer_lo;
  if (count > 0) pack_weights(e->P[g], e->pack_table[g] + first, count, e->pack_max[g], e->pack[g], S(stream));
  CGB_API_END
}

int cgb_wait_grad_bucket(cgb_engine_t* e, int index, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && index >= 0 && index < (int)e->grad_buckets.size(), "bad argument / engine not bound");
  CGB_CUDA(cudaStreamWaitEvent(S(stream), e->grad_events[index], 0));
  CGB_API_END
}

int cgb_stage_inputs(cgb_engine_t* e, const float* real_A, const float* real_B, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && real_A && real_B, "bad argument / engine not bound");
  CGB_CHECK(!e->infer_only, "inference-only engine: training entry points are unavailable");
  const size_t bytes = (size_t)e->cfg.batch * 3 * e->cfg.size * e->cfg.size * sizeof(float);
  CGB_CUDA(cudaMemcpyAsync(e->staging[0], real_A, bytes, cudaMemcpyDefault, S(stream)));
  CGB_CUDA(cudaMemcpyAsync(e->staging[1], real_B, bytes, cudaMemcpyDefault, S(stream)));
  CGB_API_END
}

int cgb_stage_inputs_u8(cgb_engine_t* e, const unsigned char* real_A, const unsigned char* real_B, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && real_A && real_B, "bad argument / engine not bound");
  CGB_CHECK(!e->infer_only, "inference-only engine: training entry points are unavailable");
  const int N = e->cfg.batch, Sz = e->cfg.size;
  const size_t bytes = (size_t)N * 3 * Sz * Sz;
  CGB_CUDA(cudaMemcpyAsync(e->staging_u8[0], real_A, bytes, cudaMemcpyDefault, S(stream)));
  CGB_CUDA(cudaMemcpyAsync(e->staging_u8[1], real_B, bytes, cudaMemcpyDefault, S(stream)));
  u8hwc_to_nchw(e->staging_u8[0], N, Sz, Sz, e->staging[0], S(stream));
  u8hwc_to_nchw(e->staging_u8[1], N, Sz, Sz, e->staging[1], S(stream));
  CGB_API_END
}

int cgb_engine_set_image_pool(cgb_engine_t* e, int pool_size) {
  CGB_API_BEGIN
  CGB_CHECK(e && !e->bound, "cgb_engine_set_image_pool must be called before cgb_engine_bind");
  CGB_CHECK(pool_size >= 0 && pool_size <= 4096, "pool_size out of range");
  e->pool_size = pool_size;
  Arena A;  // re-plan the workspace
  e->layout(A);
  e->workspace_bytes = A.off;
  CGB_API_END
}

int cgb_set_pool_decisions(cgb_engine_t* e, const int* decisions, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && decisions, "bad argument / engine not bound");
  CGB_CHECK(e->pool_size > 0, "the image pool is not enabled (cgb_engine_set_image_pool)");
  CGB_CUDA(cudaMemcpyAsync(e->pool_dec, decisions, (size_t)2 * e->cfg.batch * 2 * sizeof(int), cudaMemcpyDefault, S(stream)));
  CGB_API_END
}

int cgb_set_lr(cgb_engine_t* e, int group, float lr, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && group >= 0 && group < 2 && lr >= 0.f, "bad argument / engine not bound");
  set_device_float(e->adam_hyper[group] + 2, lr, S(stream));
  CGB_API_END
}

int cgb_get_losses_host(cgb_engine_t* e, float* losses_host, void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && losses_host, "bad argument / engine not bound");
  CGB_CHECK(!e->infer_only, "inference-only engine: training entry points are unavailable");
  CGB_CUDA(cudaMemcpyAsync(losses_host, e->losses, CGB_NUM_LOSSES * sizeof(float), cudaMemcpyDeviceToHost, S(stream)));
  CGB_CUDA(cudaStreamSynchronize(S(stream)));
  losses_host[CGB_LOSS_G] = 0.f;
  for (int i = CGB_LOSS_G_A; i <= CGB_LOSS_IDT_B; ++i) losses_host[CGB_LOSS_G] += losses_host[i];
  CGB_API_END
}

int cgb_train_step_host(cgb_engine_t* e, const float* real_A_host, const float* real_B_host, float* losses_host,
                        void* stream) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && real_A_host && real_B_host && losses_host, "bad argument / engine not bound");
  CGB_CHECK(!e->infer_only, "inference-only engine: training entry points are unavailable");
  const size_t bytes = (size_t)e->cfg.batch * 3 * e->cfg.size * e->cfg.size * sizeof(float);
  CGB_CUDA(cudaMemcpyAsync(e->staging[0], real_A_host, bytes, cudaMemcpyHostToDevice, S(stream)));
  CGB_CUDA(cudaMemcpyAsync(e->staging[1], real_B_host, bytes, cudaMemcpyHostToDevice, S(stream)));
  int rc = cgb_train_step(e, stream);
  if (rc != 0) return rc;
  rc = cgb_get_losses_host(e, losses_host, stream);
  if (rc != 0) return rc;
  CGB_API_END
}

long long cgb_launches_per_step(const cgb_engine_t* e) {
  if (!e || !e->bound) return -1;
  long long n = 0;
  for (const Program* p : e->segments[CGB_SEG_STEP].seq) n += p->launches;
  return n;
}

int cgb_profile_kind(cgb_engine_t* e, int kind, int reps, void* stream, float* ms_per_step, long long* launches,
                     double* flops) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && kind > 0 && kind < kNumOpKinds && reps > 0, "bad argument / engine not bound");
  CGB_CHECK(!e->infer_only, "inference-only engine: training entry points are unavailable");
  cudaStream_t st = S(stream);
  const Program* progs[3] = {&e->prog_cycle, &e->prog_G, &e->prog_D};
  long long count = 0;
  double fl = 0;
  for (const Program* p : progs) p->run_kind(kind, st, &count, &fl);  // warm-up + accounting
  cudaEvent_t e0, e1;
  CGB_CUDA(cudaEventCreate(&e0));
  CGB_CUDA(cudaEventCreate(&e1));
  CGB_CUDA(cudaEventRecord(e0, st));
  for (int r = 0; r < reps; ++r)
    for (const Program* p : progs) p->run_kind(kind, st, nullptr, nullptr);
  CGB_CUDA(cudaEventRecord(e1, st));
  CGB_CUDA(cudaStreamSynchronize(st));
  float ms = 0.f;
  CGB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ms_per_step) *ms_per_step = ms / reps;
  if (launches) *launches = count;
  if (flops) *flops = fl;
  CGB_API_END
}

int cgb_profile_timeline(cgb_engine_t* e, void* stream, char* buf, int buf_cap) {
  CGB_API_BEGIN
  CGB_CHECK(e && e->bound && buf && buf_cap > 0, "bad argument / engine not bound");
  CGB_CHECK(!e->infer_only, "inference-only engine: training entry points are unavailable");
  std::string t;
  if (const char* hp = std::getenv("CGB_HANG_PROBE")) {  // "steps,stall_ms,fine": development hang hunt (engine.cc hang_probe)
    int steps = 1000, stall_ms = 5000, fine = 0;
    std::sscanf(hp, "%d,%d,%d", &steps, &stall_ms, &fine);
    t = e->hang_probe(S(stream), steps, stall_ms, fine);
    if (t.empty()) t = "no hang in " + std::to_string(steps) + " replays\n";
  } else {
    t = std::getenv("CGB_PROFILE_OPS") ? e->profile_ops(S(stream), 5) : e->timeline(S(stream));
  }
  std::strncpy(buf, t.c_str(), buf_cap - 1);
  buf[buf_cap - 1] = 0;
  CGB_API_END
}

double cgb_conv_flops_per_step(const cgb_engine_t* e) { return (e && e->bound) ? e->conv_flops : -1.0; }

// ------------------------------------------------------------------------------------------------
// Single-layer harnesses for the parity tests
// ------------------------------------------------------------------------------------------------
namespace {
struct Scratch {
  std::vector<void*> ptrs;
  ~Scratch() {
    for (void* p : ptrs) cudaFree(p);
  }
  void* alloc(size_t bytes) {
    void* p = nullptr;
    CGB_CUDA(cudaMalloc(&p, bytes));
    CGB_CUDA(cudaMemset(p, 0, bytes));
    ptrs.push_back(p);
    return p;
  }
  TensorDesc tensor(int N, int H, int W, int C, int halo, int esz = 2) {
    TensorDesc t;
    t.N = N; t.H = H; t.W = W; t.C = C; t.halo = halo; t.esz = esz;
    t.ptr = static_cast<bf16*>(alloc(t.bytes()));
    return t;
  }
};

// torch layout -> master [Cout][T][Cin]   (conv: [Cout][Cin][k][k]; transposed: [Cin][Cout][k][k])
std::vector<float> to_master(const std::vector<float>& w, int cout, int cin, int k, bool transposed) {
  const int T = k * k;
  std::vector<float> m((size_t)cout * T * cin);
  for (int co = 0; co < cout; ++co)
    for (int t = 0; t < T; ++t)
      for (int ci = 0; ci < cin; ++ci) {
        const size_t src = transposed ? ((size_t)ci * cout + co) * T + t : ((size_t)co * cin + ci) * T + t;
        m[((size_t)co * T + t) * cin + ci] = w[src];
      }
  return m;
}
}  // namespace

int cgb_conv_layer_test(int n, int h, int w, int cin, int cout, int k, int stride, int pad, int reflect,
                        int transposed, int act, const float* x, const float* weight, const float* bias,
                        const float* dy, float* y, float* dx, float* dw, float* db) {
  CGB_API_BEGIN
  CGB_CHECK(x && weight, "x and weight are required");
  int sm = 148;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  ConvSpec s;
  s.Cin = cin; s.Cout = cout;
  s.CinS = cin % 64 == 0 ? cin : 16;
  s.CoutS = cout % 64 == 0 ? cout : 16;
  s.k = k; s.stride = stride; s.pad = pad; s.reflect = reflect != 0; s.transposed = transposed != 0;
  CGB_CHECK(cin % 64 == 0 || cin <= 16, "cin must be a multiple of 64 or <= 16");
  CGB_CHECK(cout % 64 == 0 || cout <= 16, "cout must be a multiple of 64 or <= 16");
  const int T = k * k, ho = out_extent(s, h), wo = out_extent(s, w);
  Scratch sc;
  const int halo = s.reflect ? pad : 0;
  TensorDesc X = sc.tensor(n, h, w, s.CinS, halo);
  TensorDesc Y = sc.tensor(n, ho, wo, s.CoutS, 0);
  nchw_to_nhwc(x, cin, X, 0);
  // weights: torch layout (device) -> master -> packs via the product pack kernel
  const size_t wn = (size_t)cout * T * cin;
  std::vector<float> wh(wn);
  CGB_CUDA(cudaMemcpy(wh.data(), weight, wn * sizeof(float), cudaMemcpyDeviceToHost));
  std::vector<float> wm = to_master(wh, cout, cin, k, s.transposed);
  float* d_master = static_cast<float*>(sc.alloc(wn * sizeof(float)));
  CGB_CUDA(cudaMemcpy(d_master, wm.data(), wn * sizeof(float), cudaMemcpyHostToDevice));
  const long long wf_elems = packed_wf_elems(s), wt_elems = packed_wt_elems(s);
  bf16* arena = static_cast<bf16*>(sc.alloc((size_t)(wf_elems + wt_elems + 1024) * sizeof(bf16)));
  PackEntry pe{};
  pe.src_off = 0; pe.wf_off = 0; pe.wt_off = (wf_elems + 511) / 512 * 512;
  pe.Cout = cout; pe.Cin = cin; pe.T = T; pe.CinS = s.CinS; pe.CoutS = s.CoutS; pe.wx_off = -1; pe.wx_pitch = 0;
  PackEntry* d_pe = static_cast<PackEntry*>(sc.alloc(sizeof(PackEntry)));
  CGB_CUDA(cudaMemcpy(d_pe, &pe, sizeof(pe), cudaMemcpyHostToDevice));
  pack_weights(d_master, d_pe, 1, (int)wn, arena, 0);

  IgemmPlan pf = plan_fprop(s, X, arena + pe.wf_off, Y, bias, act, sm);
  KIter* kt = static_cast<KIter*>(sc.alloc(pf.kiters.size() * sizeof(KIter)));
  CGB_CUDA(cudaMemcpy(kt, pf.kiters.data(), pf.kiters.size() * sizeof(KIter), cudaMemcpyHostToDevice));
  pf.args.kiters = kt;
  run(pf, 0);
  if (y) nhwc_to_nchw(Y, cout, y, 0);

  if (dy) {
    TensorDesc DY = sc.tensor(n, ho, wo, s.CoutS, 0);
    nchw_to_nhwc(dy, cout, DY, 0);
    if (dx) {
      TensorDesc DX = sc.tensor(n, h + 2 * halo, w + 2 * halo, s.CinS, 0);
      IgemmPlan pd = plan_dgrad(s, DY, arena + pe.wt_off, DX, sm);
      KIter* kd = static_cast<KIter*>(sc.alloc(pd.kiters.size() * sizeof(KIter)));
      CGB_CUDA(cudaMemcpy(kd, pd.kiters.data(), pd.kiters.size() * sizeof(KIter), cudaMemcpyHostToDevice));
      pd.args.kiters = kd;
      run(pd, 0);
      if (halo == 0) {
        nhwc_to_nchw(DX, cin, dx, 0);
      } else {
        // fold the padded-domain gradient back onto the interior with the product fold path:
        // tanh_bwd-free route: use in_bwd machinery's loader through a plain fold (act none, no norm)
        // -> implemented by a zero-initialised interior tensor + GradSrc fold via l1-free tanh? Keep it
        // simple: copy the padded-domain tensor out and fold on the host.
        std::vector<bf16> hb((size_t)DX.elems());
        CGB_CUDA(cudaMemcpy(hb.data(), DX.ptr, hb.size() * sizeof(bf16), cudaMemcpyDeviceToHost));
        std::vector<float> out((size_t)n * cin * h * w, 0.f);
        const int HP = h + 2 * halo, WP = w + 2 * halo;
        auto refl = [](int i, int nn) { if (i < 0) i = -i; if (i >= nn) i = 2 * (nn - 1) - i; return i; };
        for (int b = 0; b < n; ++b)
          for (int hp = 0; hp < HP; ++hp)
            for (int wp = 0; wp < WP; ++wp) {
              const int hs = refl(hp - halo, h), ws = refl(wp - halo, w);
              for (int c = 0; c < cin; ++c)
                out[(((size_t)b * cin + c) * h + hs) * w + ws] +=
                    __bfloat162float(hb[(((size_t)b * HP + hp) * WP + wp) * s.CinS + c]);
            }
        CGB_CUDA(cudaMemcpy(dx, out.data(), out.size() * sizeof(float), cudaMemcpyHostToDevice));
      }
    }
    if (dw) {
      float* g = static_cast<float*>(sc.alloc(wn * sizeof(float)));
      if (tc_supports_wgrad(s)) {
        WgradPlan pw = plan_wgrad(s, X, DY, g, sm);
        WTap* tp = static_cast<WTap*>(sc.alloc(pw.taps.size() * sizeof(WTap)));
        CGB_CUDA(cudaMemcpy(tp, pw.taps.data(), pw.taps.size() * sizeof(WTap), cudaMemcpyHostToDevice));
        pw.args.taps = tp;
        run(pw, 0);
      } else {
        const size_t ce = small_wgrad_col_elems(s, X, DY);
        bf16* colbuf = static_cast<bf16*>(sc.alloc(ce * sizeof(bf16)));
        SmallWgradPlan ps = plan_wgrad_small(s, X, DY, g, colbuf, ce, sm);
        WTap* tp = static_cast<WTap*>(sc.alloc(ps.gemm.taps.size() * sizeof(WTap)));
        CGB_CUDA(cudaMemcpy(tp, ps.gemm.taps.data(), ps.gemm.taps.size() * sizeof(WTap), cudaMemcpyHostToDevice));
        ps.gemm.args.taps = tp;
        if (!ps.row_map.empty()) {
          int* rm = static_cast<int*>(sc.alloc(ps.row_map.size() * sizeof(int)));
          CGB_CUDA(cudaMemcpy(rm, ps.row_map.data(), ps.row_map.size() * sizeof(int), cudaMemcpyHostToDevice));
          ps.gemm.args.row_map = rm;
        }
        if (!ps.col_map.empty()) {
          int* cm = static_cast<int*>(sc.alloc(ps.col_map.size() * sizeof(int)));
          CGB_CUDA(cudaMemcpy(cm, ps.col_map.data(), ps.col_map.size() * sizeof(int), cudaMemcpyHostToDevice));
          ps.gemm.args.col_map = cm;
        }
        run(ps, 0);
      }
      std::vector<float> gm(wn), gt(wn);
      CGB_CUDA(cudaMemcpy(gm.data(), g, wn * sizeof(float), cudaMemcpyDeviceToHost));
      for (int co = 0; co < cout; ++co)
        for (int t = 0; t < T; ++t)
          for (int ci = 0; ci < cin; ++ci) {
            const size_t dst = s.transposed ? ((size_t)ci * cout + co) * T + t : ((size_t)co * cin + ci) * T + t;
            gt[dst] = gm[((size_t)co * T + t) * cin + ci];
          }
      CGB_CUDA(cudaMemcpy(dw, gt.data(), wn * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (db) {
      CGB_CUDA(cudaMemset(db, 0, cout * sizeof(float)));
      bias_grad(DY, cout, db, 0);
    }
  }
  CGB_CUDA(cudaDeviceSynchronize());
  CGB_API_END
}

int cgb_instnorm_test(int n, int c, int h, int w, int act, const float* y, const float* residual, const float* da,
                      float* out, float* dy_out) {
  CGB_API_BEGIN
  CGB_CHECK(y && out, "y and out are required");
  CGB_CHECK(c % 8 == 0, "channels must be a multiple of 8");
  Scratch sc;
  TensorDesc Y = sc.tensor(n, h, w, c, 0);
  TensorDesc O = sc.tensor(n, h, w, c, 1);  // exercises the reflect-halo writer
  float2* stats = static_cast<float2*>(sc.alloc((size_t)n * c * sizeof(float2)));
  float2* bstats = static_cast<float2*>(sc.alloc((size_t)n * c * sizeof(float2)));
  nchw_to_nhwc(y, c, Y, 0);
  in_stats(Y, stats, 0);
  TensorDesc R;
  if (residual) {
    R = sc.tensor(n, h, w, c, 1);
    nchw_to_nhwc(residual, c, R, 0);
  }
  in_apply(Y, stats, act, residual ? &R : nullptr, O, 0);
  nhwc_to_nchw(O, c, out, 0);
  if (da && dy_out) {
    TensorDesc DA = sc.tensor(n, h, w, c, 0);
    TensorDesc DY = sc.tensor(n, h, w, c, 0);
    nchw_to_nhwc(da, c, DA, 0);
    GradSrc g;
    g.g1 = &DA;
    if (!in_bwd_fused(Y, stats, g, act, nullptr, DY, 0)) {  // the engine makes the same choice (engine.cc add_in_bwd_raw)
      in_bwd_reduce(Y, stats, g, act, nullptr, bstats, 0);
      in_bwd_apply(Y, stats, bstats, g, act, DY, 0);
    }
    nhwc_to_nchw(DY, c, dy_out, 0);
  }
  CGB_CUDA(cudaDeviceSynchronize());
  CGB_API_END
}

int cgb_instnorm_bwd_test(int n, int c, int h, int w, int act, int fold, int force_two_pass, const float* y,
                          const float* g1, const float* g2, float* dy_out, float* da_out) {
  CGB_API_BEGIN
  CGB_CHECK(y && dy_out && (g1 || g2), "y, dy_out and at least one gradient source are required");
  CGB_CHECK(c % 8 == 0, "channels must be a multiple of 8");
  Scratch sc;
  TensorDesc Y = sc.tensor(n, h, w, c, 0);
  TensorDesc DY = sc.tensor(n, h, w, c, 0);
  float2* stats = static_cast<float2*>(sc.alloc((size_t)n * c * sizeof(float2)));
  float2* bstats = static_cast<float2*>(sc.alloc((size_t)n * c * sizeof(float2)));
  nchw_to_nhwc(y, c, Y, 0);
  in_stats(Y, stats, 0);
  TensorDesc G1, G2, DA;
  GradSrc g;
  if (g1) {
    G1 = sc.tensor(n, h, w, c, 0);
    nchw_to_nhwc(g1, c, G1, 0);
    g.g1 = &G1;
  }
  if (g2) {
    G2 = sc.tensor(n, h + 2 * fold, w + 2 * fold, c, 0);
    nchw_to_nhwc(g2, c, G2, 0);
    g.g2 = &G2;
    g.fold = fold;
  }
  if (da_out) DA = sc.tensor(n, h, w, c, 0);
  if (force_two_pass || !in_bwd_fused(Y, stats, g, act, da_out ? &DA : nullptr, DY, 0)) {
    CGB_CUDA(cudaMemsetAsync(bstats, 0, (size_t)n * c * sizeof(float2), 0));
    in_bwd_reduce(Y, stats, g, act, da_out ? &DA : nullptr, bstats, 0);
    in_bwd_apply(Y, stats, bstats, g, act, DY, 0);
  }
  nhwc_to_nchw(DY, c, dy_out, 0);
  if (da_out) nhwc_to_nchw(DA, c, da_out, 0);
  CGB_CUDA(cudaDeviceSynchronize());
  CGB_API_END
}

int cgb_conv_layer_test_f32(int n, int h, int w, int cin, int cout, int k, int stride, int pad, int reflect,
                            int transposed, int act, const float* x, const float* weight, const float* bias,
                            const float* dy, float* y, float* dx, float* dw, float* db) {
  CGB_API_BEGIN
  CGB_CHECK(x && weight, "x and weight are required");
  ConvSpec s;
  s.Cin = cin; s.Cout = cout;
  s.CinS = cin % 64 == 0 ? cin : 16;
  s.CoutS = cout % 64 == 0 ? cout : 16;
  s.k = k; s.stride = stride; s.pad = pad; s.reflect = reflect != 0; s.transposed = transposed != 0;
  CGB_CHECK(cin % 64 == 0 || cin <= 16, "cin must be a multiple of 64 or <= 16");
  CGB_CHECK(cout % 64 == 0 || cout <= 16, "cout must be a multiple of 64 or <= 16");
  const int T = k * k, ho = out_extent(s, h), wo = out_extent(s, w);
  Scratch sc;
  const int halo = s.reflect ? pad : 0;
  TensorDesc X = sc.tensor(n, h, w, s.CinS, halo, 4);
  TensorDesc Y = sc.tensor(n, ho, wo, s.CoutS, 0, 4);
  nchw_to_nhwc(x, cin, X, 0);
  const size_t wn = (size_t)cout * T * cin;
  std::vector<float> wh(wn);
  CGB_CUDA(cudaMemcpy(wh.data(), weight, wn * sizeof(float), cudaMemcpyDeviceToHost));
  std::vector<float> wm = to_master(wh, cout, cin, k, s.transposed);
  float* d_master = static_cast<float*>(sc.alloc(wn * sizeof(float)));
  CGB_CUDA(cudaMemcpy(d_master, wm.data(), wn * sizeof(float), cudaMemcpyHostToDevice));
  f32::conv_fprop(s, X, d_master, bias, act, Y, 0);
  if (y) nhwc_to_nchw(Y, cout, y, 0);
  if (dy) {
    TensorDesc DY = sc.tensor(n, ho, wo, s.CoutS, 0, 4);
    nchw_to_nhwc(dy, cout, DY, 0);
    if (dx) {
      TensorDesc DX = sc.tensor(n, h + 2 * halo, w + 2 * halo, s.CinS, 0, 4);
      f32::conv_dgrad(s, DY, d_master, DX, 0);
      if (halo == 0) {
        nhwc_to_nchw(DX, cin, dx, 0);
      } else {  // fold the padded-domain gradient on the host (the engine folds it inside the InstanceNorm backward)
        std::vector<float> hb((size_t)DX.elems());
        CGB_CUDA(cudaMemcpy(hb.data(), DX.ptr, hb.size() * sizeof(float), cudaMemcpyDeviceToHost));
        std::vector<double> out((size_t)n * cin * h * w, 0.0);
        const int HP = h + 2 * halo, WP = w + 2 * halo;
        auto refl = [](int i, int nn) { if (i < 0) i = -i; if (i >= nn) i = 2 * (nn - 1) - i; return i; };
        for (int b = 0; b < n; ++b)
          for (int hp = 0; hp < HP; ++hp)
            for (int wp = 0; wp < WP; ++wp) {
              const int hs = refl(hp - halo, h), ws = refl(wp - halo, w);
              for (int c = 0; c < cin; ++c)
                out[(((size_t)b * cin + c) * h + hs) * w + ws] += hb[(((size_t)b * HP + hp) * WP + wp) * s.CinS + c];
            }
        std::vector<float> outf(out.begin(), out.end());
        CGB_CUDA(cudaMemcpy(dx, outf.data(), outf.size() * sizeof(float), cudaMemcpyHostToDevice));
      }
    }
    if (dw) {
      float* g = static_cast<float*>(sc.alloc(wn * sizeof(float)));
      f32::conv_wgrad(s, X, DY, g, 0);
      std::vector<float> gm(wn), gt(wn);
      CGB_CUDA(cudaMemcpy(gm.data(), g, wn * sizeof(float), cudaMemcpyDeviceToHost));
      for (int co = 0; co < cout; ++co)
        for (int t = 0; t < T; ++t)
          for (int ci = 0; ci < cin; ++ci) {
            const size_t dst = s.transposed ? ((size_t)ci * cout + co) * T + t : ((size_t)co * cin + ci) * T + t;
            gt[dst] = gm[((size_t)co * T + t) * cin + ci];
          }
      CGB_CUDA(cudaMemcpy(dw, gt.data(), wn * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (db) f32::bias_grad(DY, cout, db, 0);
  }
  CGB_CUDA(cudaDeviceSynchronize());
  CGB_API_END
}

int cgb_instnorm_test_f32(int n, int c, int h, int w, int act, const float* y, const float* residual, const float* da,
                          float* out, float* dy_out) {
  CGB_API_BEGIN
  CGB_CHECK(y && out, "y and out are required");
  Scratch sc;
  TensorDesc Y = sc.tensor(n, h, w, c, 0, 4);
  TensorDesc O = sc.tensor(n, h, w, c, 1, 4);
  float2* stats = static_cast<float2*>(sc.alloc((size_t)n * c * sizeof(float2)));
  nchw_to_nhwc(y, c, Y, 0);
  TensorDesc R;
  if (residual) {
    R = sc.tensor(n, h, w, c, 1, 4);
    nchw_to_nhwc(residual, c, R, 0);
  }
  f32::in_forward(Y, stats, act, residual ? &R : nullptr, O, 0);
  nhwc_to_nchw(O, c, out, 0);
  if (da && dy_out) {
    TensorDesc DA = sc.tensor(n, h, w, c, 0, 4);
    TensorDesc DY = sc.tensor(n, h, w, c, 0, 4);
    nchw_to_nhwc(da, c, DA, 0);
    GradSrc g;
    g.g1 = &DA;
    f32::in_backward(Y, stats, g, act, nullptr, DY, 0);
    nhwc_to_nchw(DY, c, dy_out, 0);
  }
  CGB_CUDA(cudaDeviceSynchronize());
  CGB_API_END
}

}  // extern "C"
