/* C ABI of the B200-native CycleGAN training-step library (libcyclegan_b200.so).
 *
 * The nominal reference (EleutherAI/Unpaired-Image-Generation) has no FFI and no code
 * (/root/reference/README.md:1 is the whole repository); the interface replaced here is the
 * public surface of the committed stand-in, oracle/cyclegan_standin.py:
 *   Generator.forward            (oracle/cyclegan_standin.py:144)  -> cgb_generator_forward
 *   Discriminator.forward        (oracle/cyclegan_standin.py:186)  -> cgb_discriminator_forward
 *   CycleGANTrainer.forward_only (oracle/cyclegan_standin.py:312)  -> cgb_forward_cycle + cgb_get_image
 *   CycleGANTrainer.train_step   (oracle/cyclegan_standin.py:374)  -> cgb_phase_generators, cgb_adam,
 *                                                                     cgb_phase_discriminators, cgb_get_losses
 *   CycleGANTrainer.backward_only(oracle/cyclegan_standin.py:352)  -> the two phase calls without cgb_adam
 *   module.state_dict()/load_state_dict()                          -> cgb_param_info + caller-owned flat buffers
 *
 * Conventions: every function returns 0 on success and a non-zero code on failure, in which case
 * cgb_last_error() returns a thread-local message.  All pointers are plain device pointers unless the
 * name says `host`.  No function synchronises the device except the *_host variants and
 * cgb_get_losses_host.  Streams are passed as void* (cudaStream_t).  The library owns only small
 * lookup tables (tap tables, pack tables); every large buffer is supplied by the caller.
 */
#ifndef CYCLEGAN_B200_H_
#define CYCLEGAN_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cgb_engine cgb_engine_t;

typedef struct cgb_config {
  int batch;        /* images per domain per step on this GPU */
  int size;         /* square image edge, multiple of 8 (canonical: 256) */
  int n_blocks;     /* residual blocks in each generator (canonical: 9) */
  float lambda_A;   /* cycle loss weight A (10) */
  float lambda_B;   /* cycle loss weight B (10) */
  float lambda_idt; /* identity loss weight (0.5) */
  float lr;         /* Adam learning rate (2e-4) */
  float beta1;      /* 0.5 */
  float beta2;      /* 0.999 */
  float eps;        /* 1e-8 */
} cgb_config_t;

enum { CGB_NET_G_AB = 0, CGB_NET_G_BA = 1, CGB_NET_D_A = 2, CGB_NET_D_B = 3 };
enum { CGB_GROUP_G = 0, CGB_GROUP_D = 1 };
enum { CGB_IMG_FAKE_B = 0, CGB_IMG_REC_A = 1, CGB_IMG_FAKE_A = 2, CGB_IMG_REC_B = 3, CGB_IMG_IDT_A = 4,
       CGB_IMG_IDT_B = 5, CGB_IMG_REAL_A = 6, CGB_IMG_REAL_B = 7,
       /* with the image pool: what the discriminators saw as fakes in the last D phase (pool.query output) */
       CGB_IMG_POOL_FAKE_B = 8, CGB_IMG_POOL_FAKE_A = 9 };
/* loss slots, same order as CycleGANTrainer.LOSS_KEYS in the stand-in */
enum { CGB_LOSS_G = 0, CGB_LOSS_G_A, CGB_LOSS_G_B, CGB_LOSS_CYCLE_A, CGB_LOSS_CYCLE_B, CGB_LOSS_IDT_A,
       CGB_LOSS_IDT_B, CGB_LOSS_D_A, CGB_LOSS_D_B, CGB_NUM_LOSSES };

typedef struct cgb_param_info {
  char name[64];        /* stand-in state_dict key, e.g. "res.3.conv1.weight" */
  int is_bias;
  int transposed;       /* 1: ConvTranspose2d (torch layout [Cin][Cout][k][k]) */
  int cout, cin, k;     /* bias: cout entries */
  long long offset;     /* element offset in the group's flat fp32 buffers */
  long long numel;
  /* weights are stored internally as [cout][k*k][cin] (fp32); a torch view of the flat buffer is
   * flat[offset:offset+numel].view(cout,k,k,cin).permute(0,3,1,2)  (conv)  or .permute(3,0,1,2) (transposed) */
} cgb_param_info_t;

const char* cgb_last_error(void);
int cgb_version(void);

/* ---- construction (host only; works without a GPU) -------------------------------------------------- */
int cgb_engine_create(const cgb_config_t* cfg, cgb_engine_t** out);
/* flags: CGB_FLAG_INFERENCE = engine for Generator.forward / Discriminator.forward only (BASELINE.json configs[4],
 * the generator-only inference sweep): the workspace holds the packed weights and ONE forward pass, none of
 * the training passes; every training entry point fails with "inference-only engine". */
/*        CGB_FLAG_FP32_VALIDATE = the fp32 VALIDATION MODE of north_star ("1e-5 for an fp32 validation mode"): the
 * same engine, programs and schedule with fp32 activations, fp64 accumulation and deterministic CUDA-core kernels
 * (csrc/fp32_path.h); bit-identical run to run; about 30x slower than the bf16 product path.  It checks the
 * orchestration against the fp32 stand-in (oracle/cyclegan_standin.py:374 train_step) below the bf16 noise floor. */
enum { CGB_FLAG_INFERENCE = 1, CGB_FLAG_FP32_VALIDATE = 2 };
int cgb_engine_create_ex(const cgb_config_t* cfg, int flags, cgb_engine_t** out);
void cgb_engine_destroy(cgb_engine_t* e);
int cgb_num_params(const cgb_engine_t* e, int net);                       /* tensors in a network */
int cgb_param_info(const cgb_engine_t* e, int net, int index, cgb_param_info_t* out);
long long cgb_group_numel(const cgb_engine_t* e, int group);              /* floats in the flat buffers */
long long cgb_workspace_bytes(const cgb_engine_t* e);

/* Image history pool of the canonical recipe (ImagePool, 50 images per domain): call BEFORE cgb_engine_bind (the
 * workspace grows by 2 * pool_size images).  With a pool the D phase sees pool.query(fake) instead of the current
 * fake.  The random decisions are made by the caller (host), one (store, ret) pair of ints per image and step:
 * d_in = ret >= 0 ? pool[ret] : fake; then pool[store] = fake when store >= 0 -- see PoolDecisions in the stand-in
 * (oracle/cyclegan_standin.py).  decisions: [2][batch][2] ints, [0] = fake_B pool (D_A), [1] = fake_A pool (D_B);
 * device or pinned host; (-1, -1) until first set. */
int cgb_engine_set_image_pool(cgb_engine_t* e, int pool_size);
int cgb_set_pool_decisions(cgb_engine_t* e, const int* decisions, void* stream);

/* ---- binding (needs the GPU) ------------------------------------------------------------------------- */
/* params/grads/m/v: fp32 flat buffers of cgb_group_numel(group) elements; workspace: cgb_workspace_bytes. */
int cgb_engine_bind(cgb_engine_t* e, float* params_G, float* grads_G, float* m_G, float* v_G, float* params_D,
                    float* grads_D, float* m_D, float* v_D, void* workspace, long long workspace_bytes);
/* re-derive the bf16 packed weights from the fp32 masters (after load_state_dict / Adam) */
int cgb_refresh_weights(cgb_engine_t* e, int group, void* stream);
int cgb_set_grad_scale(cgb_engine_t* e, float scale); /* 1/world_size for data parallel */
/* Adam step counters (the t of the bias corrections), for checkpoint / resume: together with the four flat buffers
 * of a group (params, exp_avg, exp_avg_sq; grads are scratch) they are the whole optimiser state
 * (stand-in: torch.optim.Adam.state_dict(), oracle/cyclegan_standin.py:300). */
int cgb_set_step_count(cgb_engine_t* e, int group, int step);
int cgb_get_step_count(cgb_engine_t* e, int group, int* step_out);

/* ---- modules -------------------------------------------------------------------------------------------- */
/* x, y: fp32 NCHW [batch][3][size][size] device tensors */
int cgb_generator_forward(cgb_engine_t* e, int net, const float* x, float* y, void* stream);
/* logits: fp32 [batch][1][size/8-2][size/8-2] */
int cgb_discriminator_forward(cgb_engine_t* e, int net, const float* x, float* logits, void* stream);

/* ---- training step ----------------------------------------------------------------------------------------- */
int cgb_set_inputs(cgb_engine_t* e, const float* real_A, const float* real_B, void* stream);
int cgb_forward_cycle(cgb_engine_t* e, void* stream);              /* the six generator passes */
int cgb_get_image(cgb_engine_t* e, int which, float* out, void* stream);
/* the same image as uint8 interleaved RGB [batch][size][size][3]: u8 = clamp(rint((x + 1) * 127.5), 0, 255) -- the
 * stand-in's to_uint8 (oracle/cyclegan_standin.py), inverse of cgb_stage_inputs_u8's normalisation */
int cgb_get_image_u8(cgb_engine_t* e, int which, unsigned char* out, void* stream);
/* forward + G-phase backward: grads_G = d loss_G / d(G_AB, G_BA); zeroes grads_G first */
int cgb_phase_generators(cgb_engine_t* e, void* stream);
/* D-phase forward/backward on real images and the pre-update fakes: grads_D; zeroes grads_D first */
int cgb_phase_discriminators(cgb_engine_t* e, void* stream);
/* Adam on a group (grads multiplied by grad_scale), then refreshes that group's bf16 weights */
int cgb_adam(cgb_engine_t* e, int group, void* stream);
/* Adam on the sub-range [offset, offset + numel) of a group (data parallel: step each gradient bucket right after its
 * all-reduce, overlapped with the rest of the backward pass).  advance_step != 0 for the FIRST range of an optimiser
 * step of that group (increments the step counter and refreshes the bias corrections).  Does not refresh the bf16
 * weights: call cgb_refresh_weights(group) after the last range. */
int cgb_adam_range(cgb_engine_t* e, int group, long long offset, long long numel, int advance_step, void* stream);
/* whole step on one stream, replayed from a CUDA graph after the first call */
int cgb_train_step(cgb_engine_t* e, void* stream);
/* copies fp32 NCHW inputs (device or pinned host) into the engine's staging buffers, nothing else */
int cgb_stage_inputs(cgb_engine_t* e, const float* real_A, const float* real_B, void* stream);
/* same for uint8 interleaved RGB images [batch][size][size][3] (device or pinned host): 4x fewer bytes over PCIe;
 * converted on the device with x = u8 / 127.5 - 1 (the stand-in's `from_uint8`, oracle/cyclegan_standin.py) */
int cgb_stage_inputs_u8(cgb_engine_t* e, const unsigned char* real_A, const unsigned char* real_B, void* stream);
/* learning rate of a parameter group, stream-ordered and held in device memory: an LR schedule
 * (stand-in: CycleGANTrainer.set_lr / linear_decay_lr) needs no re-capture of the step graph */
int cgb_set_lr(cgb_engine_t* e, int group, float lr, void* stream);
/* graph-replayed segments of the step, for callers that interleave their own collectives (data parallel):
 * CGB_SEG_STEP = whole step; CGB_SEG_G = staged inputs -> images, six forwards, G-phase backward;
 * CGB_SEG_D = D-phase forward/backward; CGB_SEG_ADAM_G / CGB_SEG_ADAM_D = optimiser + bf16 weight refresh.
 * First call of a segment runs eagerly, the second captures it (independent passes on parallel branches). */
enum { CGB_SEG_STEP = 0, CGB_SEG_G = 1, CGB_SEG_D = 2, CGB_SEG_ADAM_G = 3, CGB_SEG_ADAM_D = 4,
       CGB_SEG_FORWARD = 5 /* staged inputs -> images + the six generator forwards */,
       CGB_SEG_STEP_NOOPT = 6 /* staged inputs -> forward + G-phase + D-phase backward in the merged schedule, no
                                 optimiser: grads_G and grads_D are complete afterwards (data parallel) */,
       CGB_NUM_SEGMENTS = 7 };
int cgb_run_segment(cgb_engine_t* e, int segment, void* stream);
/* Data parallel, gradient all-reduce overlapped with the backward pass: while CGB_SEG_STEP_NOOPT is still running,
 * ranges of the flat gradient buffers become final bucket by bucket (discriminators first, then the generators from
 * the head towards the stem; CGB_DP_BUCKETS, default 3 per generator).  Buckets are listed in the order they become
 * ready.  cgb_wait_grad_bucket makes `stream` (the caller's communication stream) wait until bucket `index` of the
 * most recently launched CGB_SEG_STEP_NOOPT is final -- call it AFTER cgb_run_segment -- then all-reduce
 * grads[group][offset : offset + numel] on that stream.  The optimiser segments must wait for the collectives. */
int cgb_num_grad_buckets(const cgb_engine_t* e);
int cgb_grad_bucket_info(const cgb_engine_t* e, int index, int* group, long long* offset, long long* numel);
int cgb_wait_grad_bucket(cgb_engine_t* e, int index, void* stream);
/* which layers [layer_lo, layer_hi) of which network a bucket covers (net < 0: a whole parameter group) and its place
 * in the order the buckets become final (the two generators' buckets alternate: their chains run side by side) */
int cgb_grad_bucket_layers(const cgb_engine_t* e, int index, int* net, int* layer_lo, int* layer_hi, int* order);
/* bf16 weight refresh of the layers [layer_lo, layer_hi) of one network only (cgb_refresh_weights does a whole group).
 * Data parallel: the layers of generator bucket j may be refreshed as soon as bucket j + 1 of the SAME generator is
 * final (every kernel of the step that reads them has completed by then); the last bucket after the step. */
int cgb_refresh_weights_layers(cgb_engine_t* e, int net, int layer_lo, int layer_hi, void* stream);
/* copies the CGB_NUM_LOSSES loss values to host memory (synchronises the stream) */
int cgb_get_losses_host(cgb_engine_t* e, float* losses_host, void* stream);
/* end-to-end convenience: pinned/pageable HOST inputs in, losses out (H2D + step + D2H, synchronous) */
int cgb_train_step_host(cgb_engine_t* e, const float* real_A_host, const float* real_B_host, float* losses_host,
                        void* stream);

/* ---- accounting -------------------------------------------------------------------------------------------- */
long long cgb_launches_per_step(const cgb_engine_t* e);  /* kernels launched by one cgb_train_step */
double cgb_conv_flops_per_step(const cgb_engine_t* e);    /* algorithmic 2*MACs of all conv passes */

/* Replays, `reps` times between two CUDA events on `stream`, every launch of one kind of the recorded step
 * (kind: 1 = tcgen05 implicit-GEMM fprop/dgrad, 2 = tcgen05 wgrad, 3 = CUDA-core wgrad of the 3-channel
 * layers, 4 = InstanceNorm / pointwise).  Data dependencies are ignored (profiling only: gradients and
 * activations are garbage afterwards).  Returns milliseconds per step-equivalent, launches and algorithmic
 * FLOPs per step-equivalent.  Synchronises the stream. */
int cgb_profile_kind(cgb_engine_t* e, int kind, int reps, void* stream, float* ms_per_step, long long* launches,
                     double* flops);

/* Development profiling: replays the whole step as a CUDA graph with event-record nodes at the pass boundaries and
 * writes "label lane ms" lines (time since the step began) into buf.  Synchronises the stream. */
int cgb_profile_timeline(cgb_engine_t* e, void* stream, char* buf, int buf_cap);

/* ---- single-layer harness used by the parity tests (allocates its own scratch with cudaMalloc) -------------- */
/* Runs fprop (+bias, +act), and when dy != NULL also dgrad and wgrad, of one layer through the same plans the
 * engine uses.  x:[n][cin][h][w], w: torch layout, y/dy:[n][cout][ho][wo], dx like x, dw like w, db:[cout].
 * act: 0 none, 1 LeakyReLU(0.2), 2 tanh.  Any output pointer may be NULL. */
int cgb_conv_layer_test(int n, int h, int w, int cin, int cout, int k, int stride, int pad, int reflect,
                        int transposed, int act, const float* x, const float* weight, const float* bias,
                        const float* dy, float* y, float* dx, float* dw, float* db);
/* InstanceNorm(+act, +residual) forward and backward on NCHW fp32 tensors through the bf16 NHWC kernels.
 * act: 0 none, 3 ReLU, 1 LeakyReLU.  The output is produced into a tensor with a 1-pixel reflect halo (exercising
 * the halo writer); `out` receives its interior.  da: gradient w.r.t. out; dy_out: gradient w.r.t. y. */
int cgb_instnorm_test(int n, int c, int h, int w, int act, const float* y, const float* residual, const float* da,
                      float* out, float* dy_out);
/* InstanceNorm(+act) backward with the gradient sources the step uses: g1 (gradient w.r.t. the activation,
 * [n][c][h][w], may be NULL) and g2 (gradient w.r.t. the REFLECT-PADDED activation, [n][c][h+2*fold][w+2*fold],
 * may be NULL; its halo pixels are folded onto their mirror images).  dy_out: gradient w.r.t. y;
 * da_out (may be NULL): the assembled, bf16-rounded gradient w.r.t. the activation (the residual skip path).
 * force_two_pass != 0 skips the cluster-fused kernel so that the reduce + apply pair is exercised. */
int cgb_instnorm_bwd_test(int n, int c, int h, int w, int act, int fold, int force_two_pass, const float* y,
                          const float* g1, const float* g2, float* dy_out, float* da_out);
/* The same two harnesses through the fp32 validation kernels (csrc/fp32_path.h). */
int cgb_conv_layer_test_f32(int n, int h, int w, int cin, int cout, int k, int stride, int pad, int reflect,
                            int transposed, int act, const float* x, const float* weight, const float* bias,
                            const float* dy, float* y, float* dx, float* dw, float* db);
int cgb_instnorm_test_f32(int n, int c, int h, int w, int act, const float* y, const float* residual, const float* da,
                          float* out, float* dy_out);

#ifdef __cplusplus
}
#endif
#endif /* CYCLEGAN_B200_H_ */
