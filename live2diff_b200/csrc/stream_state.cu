// Device-resident stream state: everything `StreamAnimateDiffusionDepth.predict_x0_batch`
// (live2diff/pipeline_stream_animation_depth.py:573-601) keeps between frames lives in HBM and one frame is ONE CUDA
// graph: stream-batch assembly (:579-581) -> UNet step (:583, engine.cu) -> LCM x0 prediction + re-noise + buffer shift
// (:589-601) -> ring-schedule advance (:585-587 -> update_attn_bias :416-438).  The host enqueues two small copies (the
// new latent pair in, x0 out) and a graph launch per frame: no host-side tensors, no device->host reads, no generator.
#include <cstring>
#include <memory>
#include <vector>

#include "../../include/l2d_b200_debug.h"
#include "ops.cuh"
#include "stream_state.cuh"

using namespace l2d;

namespace l2d {
namespace {

constexpr uint32_t STREAM_MAGIC = 0x4c324453u;   // "L2DS"
constexpr int STREAM_MAX_L = 32, STREAM_MAX_ROWS = 8;

struct StreamHeader {   // serialised state header (host layout, little endian)
  uint32_t magic, version;
  int32_t n_rows, window, warmup, per_row;
  uint64_t frame, seed;
};

// x_cat[0] = x_new, x_cat[1:] = x_buf; d_cat likewise (:579-581)
__global__ void __launch_bounds__(256) stream_assemble_kernel(const uint4* __restrict__ x_new, const uint4* __restrict__ d_new,
                                                              const uint4* __restrict__ x_buf, const uint4* __restrict__ d_buf,
                                                              uint4* __restrict__ x_cat, uint4* __restrict__ d_cat, int n_rows,
                                                              int per_row16) {
  const int total = n_rows * per_row16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < per_row16) {
      x_cat[i] = x_new[i];
      d_cat[i] = d_new[i];
    } else {
      x_cat[i] = x_buf[i - per_row16];
      d_cat[i] = d_buf[i - per_row16];
    }
  }
}

// scheduler_step_batch (:387-401) + output/buffer shift (:589-601), the reference's fp16 evaluation order (every
// intermediate an fp16 tensor, like lcm_step_kernel in pointwise.cu); the re-noise comes from `noise_in` when
// *noise_mode == 1 (parity tests inject it), from the counter-based generator when 0, and is dropped when 2
// (do_add_noise = False).  Also d_buf = d_cat[:-1] (:601).
__global__ void __launch_bounds__(256) stream_lcm_kernel(const __half* __restrict__ x_cat, const __half* __restrict__ eps,
                                                         const float* __restrict__ consts, const __half* __restrict__ noise_in,
                                                         const int* __restrict__ noise_mode, const uint64_t* __restrict__ frame,
                                                         uint64_t seed, __half* __restrict__ out_last, __half* __restrict__ x_buf,
                                                         const __half* __restrict__ d_cat, __half* __restrict__ d_buf,
                                                         int n_rows, int per_row) {
  const size_t total = (size_t)n_rows * per_row;
  const int mode = *noise_mode;
  const uint64_t fr = *frame;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / per_row);
    const size_t e = i - (size_t)row * per_row;
    const float a = consts[row], b = consts[n_rows + row], cs = consts[2 * n_rows + row], co = consts[3 * n_rows + row];
    const float x = __half2float(x_cat[i]);
    const __half t1 = __float2half_rn(__half2float(__float2half_rn(b)) * __half2float(eps[i]));
    const __half t2 = __float2half_rn(x - __half2float(t1));
    const __half f = __float2half_rn(__half2float(t2) / __half2float(__float2half_rn(a)));
    const __half u1 = __float2half_rn(__half2float(__float2half_rn(co)) * __half2float(f));
    const __half u2 = __float2half_rn(__half2float(__float2half_rn(cs)) * x);
    const __half x0 = __float2half_rn(__half2float(u1) + __half2float(u2));
    if (row == n_rows - 1) {
      out_last[e] = x0;
    } else {
      const float an = __half2float(__float2half_rn(consts[row + 1]));
      const float bn = __half2float(__float2half_rn(consts[n_rows + row + 1]));
      const __half v1 = __float2half_rn(an * __half2float(x0));
      __half nz = __float2half_rn(0.f);
      if (mode == 1) nz = noise_in[i];
      else if (mode == 0) nz = __float2half_rn(stream_randn(seed, fr, (uint32_t)row, (uint32_t)e));   // randn -> fp16 tensor
      const __half v2 = __float2half_rn(bn * __half2float(nz));
      x_buf[i] = mode == 2 ? v1 : __float2half_rn(__half2float(v1) + __half2float(v2));
      d_buf[i] = d_cat[i];
    }
  }
}

// one thread per denoise row: schedule transition + the additive mask row the K1 kernel reads; thread 0 bumps the frame
__global__ void stream_advance_kernel(int32_t* valid, int64_t* pe_idx, int64_t* update_idx, __half* mask, uint64_t* frame,
                                      int n_rows, int window, int warmup) {
  const int r = threadIdx.x;
  if (r < n_rows) {
    ring_advance_row(valid + r, pe_idx + (size_t)r * window, update_idx + r, window, warmup);
    const int v = valid[r];
    for (int j = 0; j < window; ++j) mask[(size_t)r * window + j] = __float2half_rn(j < v ? 0.f : -INFINITY);
  }
  if (r == 0) *frame += 1;
}

__global__ void stream_mask_kernel(const int32_t* valid, __half* mask, int n_rows, int window) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_rows * window) mask[i] = __float2half_rn((i % window) < valid[i / window] ? 0.f : -INFINITY);
}

}  // namespace
}  // namespace l2d

struct l2d_stream {
  l2d_unet* unet = nullptr;
  int n = 0, h = 0, w = 0, L = 0, W0 = 0, n_kv = 0, ctx_len = 0, ctx_dim = 0, per_row = 0;
  uint64_t seed = 0;
  int do_add_noise = 1;
  // device state (one allocation)
  uint8_t* blob = nullptr;
  size_t blob_bytes = 0;
  __half *x_in = nullptr, *d_in = nullptr, *x_cat = nullptr, *d_cat = nullptr, *x_buf = nullptr, *d_buf = nullptr,
         *eps = nullptr, *out_last = nullptr, *noise_in = nullptr, *mask = nullptr, *ctx = nullptr;
  float* consts = nullptr;
  int64_t *timesteps = nullptr, *pe_idx = nullptr, *update_idx = nullptr;
  int32_t* valid = nullptr;
  int* noise_mode = nullptr;
  uint64_t* frame = nullptr;
  std::vector<void*> kv;
  bool have_prompt = false;
  int64_t consts_epoch = -1;
  bool consts_dirty = true;   // prompt changed since the engine's (timestep, prompt) projections were computed
  // whole-frame graph on an owned stream (the caller's stream may be the legacy default stream: not capturable)
  cudaStream_t own_st = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  int use_graph = 1;
  int64_t frames_done = 0, launches_per_frame = 0;
  ~l2d_stream() {
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    if (ev_in) cudaEventDestroy(ev_in);
    if (ev_out) cudaEventDestroy(ev_out);
    if (own_st) cudaStreamDestroy(own_st);
    if (blob) cudaFree(blob);
  }
};

namespace {

int enqueue_frame(l2d_stream* s, cudaStream_t st) {
  const int per16 = s->per_row / 8;
  stream_assemble_kernel<<<ceil_div(s->n * per16, 256), 256, 0, st>>>(
      reinterpret_cast<const uint4*>(s->x_in), reinterpret_cast<const uint4*>(s->d_in),
      reinterpret_cast<const uint4*>(s->x_buf), reinterpret_cast<const uint4*>(s->d_buf),
      reinterpret_cast<uint4*>(s->x_cat), reinterpret_cast<uint4*>(s->d_cat), s->n, per16);
  L2D_LAUNCH_CHECK();
  l2d_unet_step_args a{};
  a.sample = s->x_cat; a.timestep = s->timesteps; a.encoder_hidden_states = s->ctx; a.temporal_attention_mask = s->mask;
  a.depth_sample = s->d_cat; a.kv_cache = s->kv.data(); a.n_kv = s->n_kv; a.pe_idx = s->pe_idx; a.update_idx = s->update_idx;
  a.out_sample = s->eps;
  const int rc = unet_run_eager(s->unet, &a, st);
  if (rc != L2D_OK) return rc;
  const size_t total = (size_t)s->n * s->per_row;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 4);
  stream_lcm_kernel<<<blocks, 256, 0, st>>>(s->x_cat, s->eps, s->consts, s->noise_in, s->noise_mode, s->frame, s->seed,
                                            s->out_last, s->x_buf, s->d_cat, s->d_buf, s->n, s->per_row);
  L2D_LAUNCH_CHECK();
  stream_advance_kernel<<<1, 32, 0, st>>>(s->valid, s->pe_idx, s->update_idx, s->mask, s->frame, s->n, s->L, s->W0);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

int reset_state(l2d_stream* s, cudaStream_t st) {
  std::vector<int32_t> valid(s->n);
  std::vector<int64_t> pe((size_t)s->n * s->L), up(s->n);
  for (int r = 0; r < s->n; ++r) ring_init_row(r, &valid[r], &pe[(size_t)r * s->L], &up[r], s->L, s->W0);
  L2D_CUDA(cudaMemcpyAsync(s->valid, valid.data(), valid.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  L2D_CUDA(cudaMemcpyAsync(s->pe_idx, pe.data(), pe.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  L2D_CUDA(cudaMemcpyAsync(s->update_idx, up.data(), up.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  L2D_CUDA(cudaStreamSynchronize(st));   // the host vectors die at return
  const size_t buf = (size_t)(s->n - 1) * s->per_row * sizeof(__half);
  if (buf) {
    L2D_CUDA(cudaMemsetAsync(s->x_buf, 0, buf, st));   // prepare(): zero latent / depth buffers (:186-203)
    L2D_CUDA(cudaMemsetAsync(s->d_buf, 0, buf, st));
  }
  L2D_CUDA(cudaMemsetAsync(s->frame, 0, sizeof(uint64_t), st));
  stream_mask_kernel<<<ceil_div(s->n * s->L, 128), 128, 0, st>>>(s->valid, s->mask, s->n, s->L);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

}  // namespace

extern "C" int l2d_stream_create(l2d_stream** out, l2d_unet* unet, const int64_t* timesteps, const float* consts,
                                 int warmup_slots, uint64_t seed, int do_add_noise, int use_cuda_graph) {
  L2D_CHECK_ARG(out && unet && timesteps && consts, "null arguments");
  std::unique_ptr<l2d_stream> s(new l2d_stream());
  s->unet = unet;
  int warm = 0;
  unet_geometry(unet, &s->n, &s->h, &s->w, &s->L, &s->n_kv, &s->ctx_len, &s->ctx_dim, &warm);
  L2D_CHECK_ARG(warm == 0, "a stream needs a streaming engine (warmup_frames == 0)");
  L2D_CHECK_ARG(s->n >= 1 && s->n <= STREAM_MAX_ROWS && s->L <= STREAM_MAX_L, "unsupported stream-batch geometry");
  L2D_CHECK_ARG(warmup_slots > 0 && warmup_slots < s->L, "need 0 < warmup < window");
  L2D_CHECK_ARG(s->n >= 2 || warmup_slots + 1 <= s->L, "window too small");
  s->W0 = warmup_slots;
  s->seed = seed;
  s->do_add_noise = do_add_noise;
  s->use_graph = use_cuda_graph;
  s->per_row = 4 * s->h * s->w;
  L2D_CHECK_ARG(s->per_row % 8 == 0, "latent h*w must be even");
  // ---- one device allocation, 256-byte aligned sub-buffers ----
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) / 256 * 256;
    return o;
  };
  const size_t row_b = (size_t)s->per_row * sizeof(__half), nm1 = (size_t)std::max(s->n - 1, 1);
  const size_t o_xin = take(row_b), o_din = take(row_b), o_xcat = take(row_b * s->n), o_dcat = take(row_b * s->n),
               o_xbuf = take(row_b * nm1), o_dbuf = take(row_b * nm1), o_eps = take(row_b * s->n), o_out = take(row_b),
               o_noise = take(row_b * s->n), o_mask = take((size_t)s->n * s->L * sizeof(__half)),
               o_ctx = take((size_t)s->n * s->ctx_len * s->ctx_dim * sizeof(__half)), o_consts = take((size_t)4 * s->n * sizeof(float)),
               o_ts = take((size_t)s->n * sizeof(int64_t)), o_pe = take((size_t)s->n * s->L * sizeof(int64_t)),
               o_up = take((size_t)s->n * sizeof(int64_t)), o_valid = take((size_t)s->n * sizeof(int32_t)),
               o_mode = take(sizeof(int)), o_frame = take(sizeof(uint64_t));
  s->blob_bytes = off;
  L2D_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->blob), off));
  L2D_CUDA(cudaMemset(s->blob, 0, off));
  uint8_t* b = s->blob;
  s->x_in = (__half*)(b + o_xin); s->d_in = (__half*)(b + o_din); s->x_cat = (__half*)(b + o_xcat);
  s->d_cat = (__half*)(b + o_dcat); s->x_buf = (__half*)(b + o_xbuf); s->d_buf = (__half*)(b + o_dbuf);
  s->eps = (__half*)(b + o_eps); s->out_last = (__half*)(b + o_out); s->noise_in = (__half*)(b + o_noise);
  s->mask = (__half*)(b + o_mask); s->ctx = (__half*)(b + o_ctx); s->consts = (float*)(b + o_consts);
  s->timesteps = (int64_t*)(b + o_ts); s->pe_idx = (int64_t*)(b + o_pe); s->update_idx = (int64_t*)(b + o_up);
  s->valid = (int32_t*)(b + o_valid); s->noise_mode = (int*)(b + o_mode); s->frame = (uint64_t*)(b + o_frame);
  L2D_CUDA(cudaMemcpy(s->timesteps, timesteps, (size_t)s->n * sizeof(int64_t), cudaMemcpyHostToDevice));
  L2D_CUDA(cudaMemcpy(s->consts, consts, (size_t)4 * s->n * sizeof(float), cudaMemcpyHostToDevice));
  const int mode = do_add_noise ? 0 : 2;
  L2D_CUDA(cudaMemcpy(s->noise_mode, &mode, sizeof(int), cudaMemcpyHostToDevice));
  L2D_CUDA(cudaStreamCreateWithFlags(&s->own_st, cudaStreamNonBlocking));
  L2D_CUDA(cudaEventCreateWithFlags(&s->ev_in, cudaEventDisableTiming));
  L2D_CUDA(cudaEventCreateWithFlags(&s->ev_out, cudaEventDisableTiming));
  const int rc = reset_state(s.get(), s->own_st);
  if (rc != L2D_OK) return rc;
  L2D_CUDA(cudaStreamSynchronize(s->own_st));
  *out = s.release();
  return L2D_OK;
}

extern "C" void l2d_stream_destroy(l2d_stream* s) { delete s; }

extern "C" int l2d_stream_reset(l2d_stream* s, void* stream) {
  L2D_CHECK_ARG(s, "null stream");
  return reset_state(s, (cudaStream_t)stream);
}

extern "C" int l2d_stream_set_prompt(l2d_stream* s, const void* prompt_embeds, int rows, void* stream) {
  L2D_CHECK_ARG(s && prompt_embeds, "null arguments");
  L2D_CHECK_ARG(rows == 1 || rows == s->n, "prompt_embeds must have 1 or N rows");
  const size_t row_b = (size_t)s->ctx_len * s->ctx_dim * sizeof(__half);
  for (int r = 0; r < s->n; ++r)   // encoder_output.repeat(batch_size, 1, 1)  (:231, :376)
    L2D_CUDA(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(s->ctx) + r * row_b,
                             static_cast<const uint8_t*>(prompt_embeds) + (rows == 1 ? 0 : r) * row_b, row_b,
                             cudaMemcpyDefault, (cudaStream_t)stream));
  s->have_prompt = true;
  s->consts_dirty = true;
  return L2D_OK;
}

extern "C" int l2d_stream_set_cache(l2d_stream* s, void* const* kv_cache, int n_kv) {
  L2D_CHECK_ARG(s && kv_cache, "null arguments");
  L2D_CHECK_ARG(n_kv == s->n_kv, "expected " + std::to_string(s->n_kv) + " kv-cache tensors");
  for (int i = 0; i < n_kv; ++i) L2D_CHECK_ARG(kv_cache[i] != nullptr, "null kv-cache pointer");
  const bool same = s->kv.size() == (size_t)n_kv && std::memcmp(s->kv.data(), kv_cache, n_kv * sizeof(void*)) == 0;
  if (!same) {
    s->kv.assign(kv_cache, kv_cache + n_kv);
    if (s->graph_exec) {   // the graph holds the old pointers
      cudaGraphExecDestroy(s->graph_exec);
      s->graph_exec = nullptr;
    }
  }
  return L2D_OK;
}

extern "C" int l2d_stream_frame(l2d_stream* s, const void* x_t_latent, const void* depth_latent, const void* noise,
                                void* out_x0, void* stream) {
  L2D_CHECK_ARG(s && x_t_latent && depth_latent && out_x0, "null arguments");
  L2D_CHECK_ARG(s->have_prompt, "l2d_stream_set_prompt has not been called");
  L2D_CHECK_ARG((int)s->kv.size() == s->n_kv, "l2d_stream_set_cache has not been called");
  cudaStream_t caller = (cudaStream_t)stream;
  const size_t row_b = (size_t)s->per_row * sizeof(__half);
  const int64_t l0 = l2d_launch_count();
  // inputs may live in (pinned) host memory or on the device; everything else of the frame is device-resident
  L2D_CUDA(cudaEventRecord(s->ev_in, caller));
  L2D_CUDA(cudaStreamWaitEvent(s->own_st, s->ev_in, 0));
  cudaStream_t st = s->own_st;
  L2D_CUDA(cudaMemcpyAsync(s->x_in, x_t_latent, row_b, cudaMemcpyDefault, st));
  L2D_CUDA(cudaMemcpyAsync(s->d_in, depth_latent, row_b, cudaMemcpyDefault, st));
  if (s->do_add_noise && s->n > 1) {
    const int mode = noise ? 1 : 0;
    if (noise) L2D_CUDA(cudaMemcpyAsync(s->noise_in, noise, row_b * (s->n - 1), cudaMemcpyDefault, st));
    L2D_CUDA(cudaMemsetAsync(s->noise_mode, 0, sizeof(int), st));
    if (mode) L2D_CUDA(cudaMemsetAsync(s->noise_mode, 1, 1, st));   // little endian: int 1
  }
  if (s->consts_dirty || unet_consts_epoch(s->unet) != s->consts_epoch) {   // once per prompt: time-embedding / cross-attention K|V projections, outside the frame graph
    const int rc = unet_prepare_constants(s->unet, s->timesteps, s->ctx, st);
    if (rc != L2D_OK) return rc;
    s->consts_dirty = false;
    s->consts_epoch = unet_consts_epoch(s->unet);
  }
  if (!s->use_graph || s->frames_done == 0) {
    // the first frame is always eager: it sizes shared-memory attributes and fills the tensor-map cache
    const int rc = enqueue_frame(s, st);
    if (rc != L2D_OK) return rc;
    s->launches_per_frame = l2d_launch_count() - l0;
  } else {
    if (!s->graph_exec) {
      cudaGraph_t graph = nullptr;
      L2D_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      const int rc = enqueue_frame(s, st);
      cudaError_t e = cudaStreamEndCapture(st, &graph);
      if (rc != L2D_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
      }
      if (e != cudaSuccess) return fail(L2D_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
      e = cudaGraphInstantiate(&s->graph_exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) return fail(L2D_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
      s->launches_per_frame = l2d_launch_count() - l0;
      count_launch(-(int)s->launches_per_frame);   // capture enqueued nothing; the replay below is what runs
    }
    L2D_CUDA(cudaGraphLaunch(s->graph_exec, st));
    count_launch((int)s->launches_per_frame);
  }
  L2D_CUDA(cudaMemcpyAsync(out_x0, s->out_last, row_b, cudaMemcpyDefault, st));
  L2D_CUDA(cudaEventRecord(s->ev_out, st));
  L2D_CUDA(cudaStreamWaitEvent(caller, s->ev_out, 0));
  ++s->frames_done;
  return L2D_OK;
}

// developer hook (include/l2d_b200_debug.h): drop the captured frame graph so the next frame re-captures it
extern "C" void l2d_stream_invalidate_graph(l2d_stream* s) {
  if (s && s->graph_exec) {
    cudaStreamSynchronize(s->own_st);
    cudaGraphExecDestroy(s->graph_exec);
    s->graph_exec = nullptr;
  }
}

extern "C" int64_t l2d_stream_launches_per_frame(const l2d_stream* s) { return s ? s->launches_per_frame : 0; }

// ---- schedule read-back / state (de)serialisation (stream migration; tests) -- these synchronise -----------------
extern "C" int l2d_stream_get_schedule(l2d_stream* s, int32_t* valid, int64_t* pe_idx, int64_t* update_idx, uint64_t* frame) {
  L2D_CHECK_ARG(s && valid && pe_idx && update_idx, "null arguments");
  L2D_CUDA(cudaStreamSynchronize(s->own_st));
  L2D_CUDA(cudaMemcpy(valid, s->valid, (size_t)s->n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  L2D_CUDA(cudaMemcpy(pe_idx, s->pe_idx, (size_t)s->n * s->L * sizeof(int64_t), cudaMemcpyDeviceToHost));
  L2D_CUDA(cudaMemcpy(update_idx, s->update_idx, (size_t)s->n * sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (frame) L2D_CUDA(cudaMemcpy(frame, s->frame, sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return L2D_OK;
}

extern "C" int64_t l2d_stream_state_bytes(const l2d_stream* s) {
  if (!s) return 0;
  return (int64_t)(sizeof(StreamHeader) + (size_t)s->n * sizeof(int32_t) + (size_t)s->n * s->L * sizeof(int64_t) +
                   (size_t)s->n * sizeof(int64_t) + 2 * (size_t)(s->n - 1) * s->per_row * sizeof(__half));
}

extern "C" int l2d_stream_save_state(l2d_stream* s, void* host_buf, int64_t bytes) {
  L2D_CHECK_ARG(s && host_buf, "null arguments");
  L2D_CHECK_ARG(bytes >= l2d_stream_state_bytes(s), "buffer too small");
  L2D_CUDA(cudaStreamSynchronize(s->own_st));
  uint8_t* p = static_cast<uint8_t*>(host_buf);
  StreamHeader hd{STREAM_MAGIC, 1, s->n, s->L, s->W0, s->per_row, 0, s->seed};
  L2D_CUDA(cudaMemcpy(&hd.frame, s->frame, sizeof(uint64_t), cudaMemcpyDeviceToHost));
  std::memcpy(p, &hd, sizeof(hd));
  p += sizeof(hd);
  auto pull = [&](const void* src, size_t n) -> cudaError_t {
    cudaError_t e = n ? cudaMemcpy(p, src, n, cudaMemcpyDeviceToHost) : cudaSuccess;
    p += n;
    return e;
  };
  L2D_CUDA(pull(s->valid, (size_t)s->n * sizeof(int32_t)));
  L2D_CUDA(pull(s->pe_idx, (size_t)s->n * s->L * sizeof(int64_t)));
  L2D_CUDA(pull(s->update_idx, (size_t)s->n * sizeof(int64_t)));
  L2D_CUDA(pull(s->x_buf, (size_t)(s->n - 1) * s->per_row * sizeof(__half)));
  L2D_CUDA(pull(s->d_buf, (size_t)(s->n - 1) * s->per_row * sizeof(__half)));
  return L2D_OK;
}

extern "C" int l2d_stream_load_state(l2d_stream* s, const void* host_buf, int64_t bytes) {
  L2D_CHECK_ARG(s && host_buf, "null arguments");
  L2D_CHECK_ARG(bytes >= l2d_stream_state_bytes(s), "buffer too small");
  const uint8_t* p = static_cast<const uint8_t*>(host_buf);
  StreamHeader hd;
  std::memcpy(&hd, p, sizeof(hd));
  p += sizeof(hd);
  L2D_CHECK_ARG(hd.magic == STREAM_MAGIC && hd.version == 1, "not a serialised l2d stream state");
  L2D_CHECK_ARG(hd.n_rows == s->n && hd.window == s->L && hd.warmup == s->W0 && hd.per_row == s->per_row,
                "serialised state belongs to a stream of a different geometry");
  L2D_CUDA(cudaStreamSynchronize(s->own_st));
  s->seed = hd.seed;
  if (s->graph_exec) {   // the seed is a kernel argument baked into the graph
    cudaGraphExecDestroy(s->graph_exec);
    s->graph_exec = nullptr;
  }
  L2D_CUDA(cudaMemcpy(s->frame, &hd.frame, sizeof(uint64_t), cudaMemcpyHostToDevice));
  auto push = [&](void* dst, size_t n) -> cudaError_t {
    cudaError_t e = n ? cudaMemcpy(dst, p, n, cudaMemcpyHostToDevice) : cudaSuccess;
    p += n;
    return e;
  };
  L2D_CUDA(push(s->valid, (size_t)s->n * sizeof(int32_t)));
  L2D_CUDA(push(s->pe_idx, (size_t)s->n * s->L * sizeof(int64_t)));
  L2D_CUDA(push(s->update_idx, (size_t)s->n * sizeof(int64_t)));
  L2D_CUDA(push(s->x_buf, (size_t)(s->n - 1) * s->per_row * sizeof(__half)));
  L2D_CUDA(push(s->d_buf, (size_t)(s->n - 1) * s->per_row * sizeof(__half)));
  stream_mask_kernel<<<ceil_div(s->n * s->L, 128), 128, 0, s->own_st>>>(s->valid, s->mask, s->n, s->L);
  L2D_LAUNCH_CHECK();
  L2D_CUDA(cudaStreamSynchronize(s->own_st));
  return L2D_OK;
}

// ---- host evaluation of the same schedule / generator code (no GPU needed: CPU tests pin them) --------------------
extern "C" int l2d_ring_schedule_host(int32_t* valid, int64_t* pe_idx, int64_t* update_idx, int n_rows, int window,
                                      int warmup, int init, int advance_frames) {
  L2D_CHECK_ARG(valid && pe_idx && update_idx && n_rows >= 1 && window >= 2 && warmup > 0 && warmup < window, "bad arguments");
  if (init)
    for (int r = 0; r < n_rows; ++r) ring_init_row(r, valid + r, pe_idx + (size_t)r * window, update_idx + r, window, warmup);
  for (int f = 0; f < advance_frames; ++f)
    for (int r = 0; r < n_rows; ++r) ring_advance_row(valid + r, pe_idx + (size_t)r * window, update_idx + r, window, warmup);
  return L2D_OK;
}

extern "C" int l2d_stream_randn_host(uint64_t seed, uint64_t frame, uint32_t row, float* out, int count) {
  L2D_CHECK_ARG(out && count >= 0, "bad arguments");
  for (int i = 0; i < count; ++i) out[i] = stream_randn(seed, frame, row, (uint32_t)i);
  return L2D_OK;
}

extern "C" void l2d_philox4x32_10_host(const uint32_t* counter4, const uint32_t* key2, uint32_t* out4) {
  uint32_t o[4];
  philox4x32_10(counter4[0], counter4[1], counter4[2], counter4[3], key2[0], key2[1], o);
  for (int i = 0; i < 4; ++i) out4[i] = o[i];
}
