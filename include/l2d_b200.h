/*
 * l2d_b200.h -- C ABI of libl2d_b200.so: the B200 (sm_100a) implementation of Live2Diff's per-frame
 * streaming UNet step.  Plain pointers and sizes only; every pointer is a DEVICE pointer to fp16
 * (IEEE half) data unless stated otherwise; `stream` is a cudaStream_t passed as void*.
 *
 * Every function returns L2D_OK (0) or a negative code; l2d_last_error() gives the message of the
 * last failure on the calling thread (the Python shims raise it as RuntimeError, mirroring the
 * reference, which reports errors as Python exceptions only).  No function allocates device memory
 * per call except the *_create functions; no function synchronises the device.
 *
 * Reference interfaces replaced (paths relative to the Live2Diff tree):
 *   B1  stream.unet(...)                      live2diff/pipeline_stream_animation_depth.py:456-466
 *       (same swap point the TensorRT engine object uses: live2diff/utils/wrapper.py:613,
 *        live2diff/acceleration/tensorrt/engine.py:142-185)            -> l2d_unet_*
 *   B2  TemporalTransformer3DModel.forward    live2diff/animatediff/models/motion_module.py:256-299  -> l2d_tt_*
 *   B3  StreamTemporalAttention.forward core  live2diff/animatediff/models/stream_motion_module.py:99-194
 *                                                                        -> l2d_kv_attn
 *   a4  scheduler_step_batch + stream-batch shift  pipeline_stream_animation_depth.py:387-401,589-601
 *                                                                        -> l2d_lcm_step
 *   f4  predict_x0_batch state machine (a2 + a5 + a4 around B1), device-resident
 *                                              pipeline_stream_animation_depth.py:403-438, 573-601 -> l2d_stream_*
 *   f1  stream.unet_warmup(...) (warm-up pass)  pipeline_stream_animation_depth.py:315-338; the bidirectional
 *       VersatileAttention + cache fill          live2diff/animatediff/models/motion_module.py:469-530
 *                                                -> l2d_unet_* with cfg.warmup_frames > 0, l2d_warmup_attn
 * The op-level entry points (l2d_gemm, l2d_layernorm, ...) are the kernels those are built from;
 * they are exported so each kernel can be parity-tested through the same ABI.
 */
#ifndef L2D_B200_H
#define L2D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define L2D_ABI_VERSION 3

#define L2D_OK 0
#define L2D_ERR_INVALID (-1)     /* bad argument / unsupported shape */
#define L2D_ERR_CUDA (-2)        /* a CUDA runtime/driver call failed */
#define L2D_ERR_MISSING (-3)     /* a required weight tensor was not supplied */

int l2d_abi_version(void);
/* sha1 of the CUDA sources the library was compiled from (live2diff_b200/csrc/build.py compares it with the tree). */
const char* l2d_build_hash(void);
const char* l2d_last_error(void);
/* Number of kernels this library has launched in the calling process (monotonic counter). */
int64_t l2d_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * B3: temporal KV-cache attention (K1).  One call = steps 2-4 of SURVEY.md Appendix D:
 *   cache[n,0,p,u_n,:] = k_new[n,p,:]; cache[n,1,p,u_n,:] = v_new[n,p,:]          (PE-free append)
 *   out[n,p,head] = softmax_j( <q+Qpe[pi_n[u_n]], cache_k[j]+Kpe[pi_n[j]]> / sqrt(hd) + mask[n,j] )
 *                   . (cache_v[j] + Vpe[pi_n[j]])
 * q/k_new/v_new: rows n*hw+p, `qkv_ld` elements between rows (so they may alias a fused [M,3C] buffer).
 * kv_cache [N,2,hw,L,C] contiguous, mutated in place.  q_pe/k_pe/v_pe [L,C].  mask [N,L] fp16 additive.
 * pe_idx [N,L] int64, update_idx [N] int64 (device).  out [N*hw, C] contiguous.
 * Constraints: C % 8 == 0, (C/heads) % 8 == 0, L <= 32.
 * ------------------------------------------------------------------------------------------- */
int l2d_kv_attn(const void* q, const void* k_new, const void* v_new, int64_t qkv_ld, void* kv_cache,
                const void* q_pe, const void* k_pe, const void* v_pe, const void* mask,
                const int64_t* pe_idx, const int64_t* update_idx, void* out,
                int n_rows, int hw, int window, int channels, int heads, void* stream);

/* f1: warm-up temporal attention of ONE clip of `frames` frames (VersatileAttention.forward core, motion_module.py:488-516):
 *   kv_cache_row[0,p,f,:] = k[f,p,:]; kv_cache_row[1,p,f,:] = v[f,p,:]            (PE-free fill of slots 0..frames-1)
 *   out[f,p,head] = softmax_j( <q[f]+Qpe[f], k[j]+Kpe[j]> / sqrt(hd) ) . (v[j]+Vpe[j]),  j over the frames, no mask
 * q/k/v: rows f*hw+p, `qkv_ld` elements between rows.  kv_cache_row = cache[idx] of one denoise row: [2,hw,L,C].
 * q_pe/k_pe/v_pe: rows 0..frames-1 with pitch pe_ld.  out rows f*hw+p with pitch out_ld.  frames <= min(16, L). */
int l2d_warmup_attn(const void* q, const void* k, const void* v, int64_t qkv_ld, void* kv_cache_row,
                    const void* q_pe, const void* k_pe, const void* v_pe, int64_t pe_ld, void* out, int64_t out_ld,
                    int frames, int hw, int window, int channels, int heads, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Op-level kernels (row-major fp16 activations, "tokens x channels" = NHWC).
 * ------------------------------------------------------------------------------------------- */

/* y[m,:] = LayerNorm(x[m,:]) * gamma + beta, eps 1e-5 default of nn.LayerNorm. C % 8 == 0. */
int l2d_layernorm(const void* x, const void* gamma, const void* beta, void* y, int rows, int channels,
                  float eps, void* stream);

/* GroupNorm over NHWC x = concat_channels(x1[N,hw,C1], x2[N,hw,C2]) (x2 may be NULL, C2 = 0), `groups`
 * groups, optional SiLU.  `workspace` >= l2d_groupnorm_workspace_bytes(...) bytes, zero-filled ONCE by the caller
 * (it holds arrival counters that every launch leaves at zero); launches that share a workspace must be ordered on one
 * stream.  mode 0 runs as ONE kernel whose CTAs (<= #SMs, all resident) meet at a barrier in the workspace, so it must
 * not be launched concurrently with a kernel that never yields its SMs.
 * mode 0: y[N,hw,C] ; mode 1: y = im2col3x3(pad 1, stride `stride`) of the normalised tensor,
 * rows = N*(h/stride)*(w/stride), cols = 9*C ordered (tap, channel). */
int64_t l2d_groupnorm_workspace_bytes(int n_img, int groups);
int l2d_groupnorm(const void* x1, int c1, const void* x2, int c2, const void* gamma, const void* beta, void* y,
                  void* workspace, int n_img, int h, int w, int groups, float eps, int silu, int mode,
                  int stride, void* stream);

/* im2col3x3 (pad 1) of a raw NHWC tensor [N,h,w,C] with optional nearest x2 upsample before the
 * window (Upsample3D) or stride 2 (Downsample3D), optional SiLU on the source.  C % 8 == 0.
 * Output rows = N*ho*wo, cols = 9*C. */
int l2d_im2col3x3(const void* x, void* y, int n_img, int h, int w, int channels, int stride, int upsample2x,
                  int silu, void* stream);
/* Same for the 4-channel NCHW latent [N,4,h,w]: output [N*h*w, 64] with cols (tap,channel) in 0..35, rest 0. */
int l2d_im2col3x3_nchw4(const void* x, void* y, int n_img, int h, int w, void* stream);

/* NCHW <-> NHWC for [N,C,h,w] fp16.  nhwc_to_nchw optionally adds a residual given in NCHW. */
int l2d_nchw_to_nhwc(const void* x, void* y, int n_img, int channels, int hw, void* stream);
int l2d_nhwc_to_nchw(const void* x, const void* residual_nchw, void* y, int n_img, int channels, int hw, void* stream);

/* out[M,N] = epilogue( A[M,K] . W[N,K]^T )  -- tcgen05 tensor-core GEMM, fp32 accumulate.
 * A rows have `lda` elements (K <= lda), W is [N,K] contiguous (nn.Linear layout; a conv3x3 weight
 * repacked to [Cout, 9*Cin] (tap,cin) order), out rows have `ldo` elements.
 * Epilogue, in order:  + bias[n]  + rowgroup_bias[m / rows_per_group, n]  -> act  -> + residual[m,n].
 * act: 0 none, 1 SiLU, 2 GEGLU (W rows must be tile-interleaved by l2d_geglu_interleave; out has N/2 cols), 3 ReLU,
 *      4 ReLU applied after the residual add.
 * Any of bias / rowgroup_bias / residual may be NULL.  K % 8 == 0, N % 8 == 0. */
#define L2D_ACT_NONE 0
#define L2D_ACT_SILU 1
#define L2D_ACT_GEGLU 2
#define L2D_ACT_RELU 3        /* ReLU in the activation slot (before the residual) */
#define L2D_ACT_RELU_POST 4   /* no activation before the residual, ReLU after it: ReLU(conv(x) + skip) (AutoencoderTinyBlock.fuse) */
int l2d_gemm(const void* a, int64_t lda, const void* w, void* out, int64_t ldo, int m, int n, int k,
             const void* bias, const void* rowgroup_bias, int rows_per_group, const void* residual,
             int64_t ldr, int act, void* stream);
/* Implicit-GEMM 3x3 convolution, pad 1, stride 1 (no im2col matrix): x [N,h,w,Cin] channels-last contiguous,
 * weight [Cout, 9*Cin] with columns ordered (tap = ky*3+kx, cin), out rows = pixels (n*h*w + y*w + x), `ldo` elements
 * per row.  Epilogue as l2d_gemm (rowgroup_bias [N, Cout] is per image).  Needs Cin % 64 == 0 and h, w whose
 * power-of-two divisors tile 128 pixels (true for every stride-1 conv of the UNet at 512x512 / 768x512). */
int l2d_conv3x3(const void* x, int n_img, int h, int w, int cin, const void* weight, void* out, int64_t ldo, int cout,
                const void* bias, const void* rowgroup_bias, const void* residual, int64_t ldr, int act, void* stream);
/* Row permutation applied to a GEGLU projection weight [2F,K] (+ bias [2F]) so that every GEMM N-tile
 * holds matching value/gate columns.  `tile_n` = the N tile l2d_gemm will use for n=2F (l2d_gemm_tile_n). */
int l2d_gemm_tile_n(int m, int n, int k);
int l2d_geglu_interleave(const void* w_in, const void* b_in, void* w_out, void* b_out, int two_f, int k,
                         int tile_n, void* stream);

/* out[m,n] = act_out( sum_k act_in(x[m,k]) W[n,k] + b[n] ) for tiny m (<= 8): time MLP, time_emb_proj. */
int l2d_small_linear(const void* x, const void* w, const void* b, void* out, int m, int n, int k, int silu_in,
                     int silu_out, void* stream);
/* diffusers Timesteps(flip_sin_to_cos=True, freq_shift=0): t int64[N] -> fp16 [N, dim] = [cos | sin]. */
int l2d_timestep_embedding(const int64_t* t, void* out, int n, int dim, void* stream);

/* softmax(Q K^T / sqrt(hd)) V, multi-head, no mask (spatial self-/cross-attention).
 * q rows: (b*sq + i), head h at columns [q_off + h*hd, +hd), row stride ldq; k/v likewise with skv rows
 * per batch; out [b*sq, heads*hd] with row stride ldo.  hd % 8 == 0, hd <= 160. */
int l2d_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                  int64_t ldo, int batch, int heads, int sq, int skv, int hd, void* stream);

/* a4 + a2 (K9): x0 = c_out*(x - b*eps)/a + c_skip*x per row; out_last = x0[N-1];
 * next_buf[i] = a[i+1]*x0[i] + b[i+1]*noise[i] (i < N-1).  consts = fp32 [4,N] = (a, b, c_skip, c_out) on device.
 * All tensors [N,4,h,w] fp16 contiguous (rows = stream-batch rows); noise/next_buf have N-1 rows (may be NULL if N==1). */
int l2d_lcm_step(const void* x_t, const void* eps, const float* consts, const void* noise, void* x0_all,
                 void* out_last, void* next_buf, int n_rows, int elems_per_row, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Weights hand-over for the module/engine objects: the reference's state_dict, one entry per tensor.
 * `name` is the reference key relative to the object (e.g. "proj_in.weight"); data = device fp16,
 * contiguous, shape as in the reference.  The object copies/repacks what it needs; the caller keeps
 * ownership of its tensors.
 * ------------------------------------------------------------------------------------------- */
typedef struct l2d_tensor {
  const char* name;
  const void* data;
  int32_t ndim;
  int64_t shape[5];
} l2d_tensor;

/* B2: one TemporalTransformer3DModel (motion_module.py:153-299), streaming mode, 1 transformer block
 * with 2 StreamTemporalAttention blocks. */
typedef struct l2d_tt l2d_tt;
int l2d_tt_create(l2d_tt** out, const l2d_tensor* weights, int n_weights, int channels, int heads, int groups,
                  int window, int n_rows, int h, int w);
/* x, y: [N,C,1,h,w] (NCHW, f == 1).  kv_cache0/1: the two attention blocks' caches [N,2,hw,L,C]. */
int l2d_tt_forward(l2d_tt* tt, const void* x_nchw, void* y_nchw, void* kv_cache0, void* kv_cache1, const void* mask,
                   const int64_t* pe_idx, const int64_t* update_idx, void* stream);
void l2d_tt_destroy(l2d_tt* tt);

/* B1: the whole streaming UNet step. */
typedef struct l2d_unet l2d_unet;
typedef struct l2d_unet_config {
  int32_t n_levels;                 /* 4 */
  int32_t block_out_channels[8];    /* 320,640,1280,1280 */
  int32_t layers_per_block;         /* 2 */
  int32_t heads;                    /* 8 */
  int32_t cross_attention_dim;      /* 768 */
  int32_t ctx_len;                  /* 77 */
  int32_t groups;                   /* 32 */
  int32_t window;                   /* L */
  int32_t n_rows;                   /* N = denoising steps in the stream batch */
  int32_t latent_h, latent_w;       /* 64, 64 */
  int32_t mapping_channels[8];      /* 16,32,96,256 */
  int32_t n_mapping;                /* 4 */
  int32_t down_has_attn[8];         /* 1,1,1,0 */
  int32_t up_has_attn[8];           /* 0,1,1,1 */
  float norm_eps;                   /* 1e-5 */
  int32_t use_cuda_graph;           /* capture the step into a CUDA graph on first call and replay it */
  int32_t warmup_frames;            /* 0: streaming step.  F > 0: warm-up engine (UNet3DConditionWarmupModel): n_rows must
                                       equal F (the frames of one clip ride on the batch axis), the motion modules run the
                                       bidirectional attention over the F frames and fill cache slots 0..F-1; step args:
                                       kv_cache[i] = row `idx` of cache i ([2,hw,L,C]); mask / pe_idx / update_idx unused
                                       (may be NULL); timestep / encoder_hidden_states repeated per frame by the caller */
} l2d_unet_config;

typedef struct l2d_unet_step_args {
  const void* sample;               /* [N,4,1,h,w] fp16 */
  const int64_t* timestep;          /* [N] int64 (device) */
  const void* encoder_hidden_states;/* [N,ctx_len,cross_attention_dim] fp16 */
  const void* temporal_attention_mask; /* [N,L] fp16 additive */
  const void* depth_sample;         /* [N,4,1,h,w] fp16 */
  void* const* kv_cache;            /* host array of n_kv device pointers, motion_module_idx order */
  int32_t n_kv;                     /* 2 * (#motion modules) = 40 */
  const int64_t* pe_idx;            /* [N,L] int64 (device) */
  const int64_t* update_idx;        /* [N] int64 (device) */
  void* out_sample;                 /* [N,4,1,h,w] fp16 */
  int32_t reuse_constants;          /* 0: timestep / encoder_hidden_states may have changed -- their projections (time MLP,
                                       22 x time_emb_proj, 16 x cross-attention K|V) are recomputed before the step.
                                       1: both are unchanged since this engine's previous step: the projections are reused
                                       (the reference recomputes them every frame although both are constant per prompt,
                                       pipeline_stream_animation_depth.py:231-246).  They are never part of the CUDA graph. */
} l2d_unet_step_args;

int l2d_unet_create(l2d_unet** out, const l2d_unet_config* cfg, const l2d_tensor* weights, int n_weights);
/* A second engine over the SAME repacked weights (no copy: it owns only its workspace): same topology as `base`, its own
 * n_rows / latent size / warmup_frames / use_cuda_graph.  `base` must outlive it.  This is how the warm-up engine
 * (cfg.warmup_frames > 0) sits next to the streaming engine without a second 2.6 GiB of weights. */
int l2d_unet_create_shared(l2d_unet** out, const l2d_unet_config* cfg, const l2d_unet* base);
int l2d_unet_step(l2d_unet* u, const l2d_unet_step_args* args, void* stream);
/* Per-family kernel time of one step: the step is run eagerly once per family with CUDA events around that family's
 * launches only (so the host stays ahead of the GPU and the stream runs back to back, as in the graph replay), and
 * the event times are summed per family on return (synchronises the stream).  Families: 0 temporal KV-cache
 * attention (K1), 1 tcgen05 GEMM (linears + convs), 2 spatial attention, 3 LayerNorm/GroupNorm, 4 im2col (incl.
 * GroupNorm+SiLU+im2col), 5 other.  ms_by_family / launches_by_family: host arrays of 6.  The passes are real,
 * idempotent steps on the same inputs (the KV cache advances once: the same slot is rewritten with the same k/v). */
#define L2D_N_FAMILIES 6
int l2d_unet_profile_step(l2d_unet* u, const l2d_unet_step_args* args, void* stream, float* ms_by_family,
                          int32_t* launches_by_family);
/* Counter bumped every time the engine recomputed its (timestep, prompt) projections: a caller that passes
 * reuse_constants = 1 while sharing the engine with another caller (e.g. a device-resident stream) compares it with the
 * value it saw after its own last step and passes 0 when someone else has recomputed them in between. */
int64_t l2d_unet_constants_epoch(const l2d_unet* u);
/* Bytes of device memory owned by the engine (weights + workspace). */
int64_t l2d_unet_device_bytes(const l2d_unet* u);
/* Kernel launches per step (counted on the most recent step). */
int64_t l2d_unet_launches_per_step(const l2d_unet* u);
void l2d_unet_destroy(l2d_unet* u);

/* ---------------------------------------------------------------------------------------------
 * f4: device-resident stream = the state machine of predict_x0_batch (pipeline_stream_animation_depth.py:573-601) with
 * its state in HBM: latent / depth buffers [(N-1),4,h,w], the KV ring schedule (attn_bias, pe_idx, update_idx: :403-438),
 * LCM constants (:260-301), prompt embedding, frame counter.  One frame = one CUDA graph (assembly -> UNet step ->
 * x0 prediction + re-noise + shift -> schedule advance); the re-noise is a counter-based Philox4x32-10 normal keyed by
 * (seed, frame, row, element) instead of torch's global generator.  The KV caches stay caller-owned tensors (B1).
 *   timesteps [N] int64 and consts [4,N] fp32 = (sqrt(abar), sqrt(1-abar), c_skip, c_out): HOST arrays (copied).
 *   x_t_latent / depth_latent / out_x0: [1,4,1,h,w] fp16, device OR pinned-host pointers (copied in/out on the stream).
 *   noise: NULL (internal generator) or [(N-1),4,1,h,w] fp16 overriding it for this frame (parity tests).
 * ------------------------------------------------------------------------------------------- */
typedef struct l2d_stream l2d_stream;
int l2d_stream_create(l2d_stream** out, l2d_unet* unet, const int64_t* timesteps, const float* consts, int warmup_slots,
                      uint64_t seed, int do_add_noise, int use_cuda_graph);
void l2d_stream_destroy(l2d_stream* s);
/* prepare(): zero buffers, schedule back to its initial state, frame counter 0 (:171-214, 403-414). */
int l2d_stream_reset(l2d_stream* s, void* stream);
/* prompt_embeds [rows,ctx_len,cross_attention_dim] fp16, rows = 1 (repeated to N, :231/:376) or N. */
int l2d_stream_set_prompt(l2d_stream* s, const void* prompt_embeds, int rows, void* stream);
/* host array of n_kv device pointers ([N,2,hw,L,C] each, motion_module_idx order); re-captures the graph if they change. */
int l2d_stream_set_cache(l2d_stream* s, void* const* kv_cache, int n_kv);
int l2d_stream_frame(l2d_stream* s, const void* x_t_latent, const void* depth_latent, const void* noise, void* out_x0,
                     void* stream);
int64_t l2d_stream_launches_per_frame(const l2d_stream* s);
/* Read-back of the schedule (synchronises): valid [N] int32 (unmasked leading slots), pe_idx [N,L], update_idx [N]. */
int l2d_stream_get_schedule(l2d_stream* s, int32_t* valid, int64_t* pe_idx, int64_t* update_idx, uint64_t* frame);
/* Stream migration: buffers + schedule + frame counter + seed as one host blob (synchronises; KV caches not included). */
int64_t l2d_stream_state_bytes(const l2d_stream* s);
int l2d_stream_save_state(l2d_stream* s, void* host_buf, int64_t bytes);
int l2d_stream_load_state(l2d_stream* s, const void* host_buf, int64_t bytes);
/* The schedule transition and the generator evaluated on the HOST by the same code the kernels run (CPU-side tests):
 * init != 0 writes the initial state first, then `advance_frames` transitions are applied to the host arrays. */
int l2d_ring_schedule_host(int32_t* valid, int64_t* pe_idx, int64_t* update_idx, int n_rows, int window, int warmup,
                           int init, int advance_frames);
int l2d_stream_randn_host(uint64_t seed, uint64_t frame, uint32_t row, float* out, int count);
void l2d_philox4x32_10_host(const uint32_t* counter4, const uint32_t* key2, uint32_t* out4);

/* ---------------------------------------------------------------------------------------------
 * f3: the tiny VAE around the step -- `stream.vae` = diffusers AutoencoderTiny (live2diff/utils/wrapper.py:468-470) as
 * called by encode_image / encode_depth / decode_image (pipeline_stream_animation_depth.py:517-542, 565-571), and the
 * uint8 <-> [-1,1] conversions of __call__ (:630; live2diff/image_utils.py:9-30).  Weights: AutoencoderTiny.state_dict()
 * (keys encoder.layers.N..., decoder.layers.N...; fp16, device).  scaling_factor is 1.0 and applied by the caller.
 *   encode: image [n,3,H,W] fp16 in [-1,1] -> latents [n,4,H/8,W/8]           (vae.encode(x).latents)
 *   decode: latents [n,4,H/8,W/8] -> image [n,3,H,W] fp16; clip != 0 also applies .clip(-1,1)   (vae.decode(z)[0])
 * H, W multiples of 8 whose power-of-two divisors tile 128 pixels at every scale (512x512, 768x512, ...).
 * ------------------------------------------------------------------------------------------- */
typedef struct l2d_taesd l2d_taesd;
int l2d_taesd_create(l2d_taesd** out, const l2d_tensor* weights, int n_weights, int max_batch, int height, int width);
int l2d_taesd_encode(l2d_taesd* t, const void* image_nchw, void* latents, int n, void* stream);
int l2d_taesd_decode(l2d_taesd* t, const void* latents, void* image_nchw, int n, int clip, void* stream);
int64_t l2d_taesd_device_bytes(const l2d_taesd* t);
void l2d_taesd_destroy(l2d_taesd* t);
/* uint8 [n,H,W,3] -> fp16 [n,3,H,W] = x/255*2-1 (VaeImageProcessor.preprocess), and back: (x/2+0.5).clamp(0,1)*255 rounded
 * half-to-even (image_utils.denormalize + numpy_to_pil). */
int l2d_image_u8_to_f16(const void* u8_nhwc, void* f16_nchw, int n, int h, int w, void* stream);
int l2d_image_f16_to_u8(const void* f16_nchw, void* u8_nhwc, int n, int h, int w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* L2D_B200_H */
