// Internal launcher interfaces shared between the kernel files and the engine.
#pragma once
#include "common.cuh"

namespace l2d {

struct KvAttnParams {
  const __half* q;
  const __half* k_new;
  const __half* v_new;
  int64_t ld;
  __half* cache;
  const __half* q_pe;
  const __half* k_pe;
  const __half* v_pe;
  const __half* mask;
  const int64_t* pe_idx;
  const int64_t* update_idx;
  __half* out;
  int n_rows, hw, L, C, heads;
  int64_t pe_ld;  // row pitch of the q_pe/k_pe/v_pe tables (elements)
  int T;    // C / 8 chunk-threads per pixel
  int P;    // pixels per block
  int hd8;  // chunks per head
  float scale;
  int c_slices;  // tensor-core kernel: channel slices per row (1 = whole rows); a slice holds whole heads and is scheduled like
                 // a row of its own (own PE windows); C / heads / T then describe ONE slice and c_full the row pitch
  int c_full;
  int eager_planes;  // tensor-core kernel: planes the producer requests before the row's schedule / PE loads are queued
  int pdl;  // 1: launched behind the QKV GEMM inside the engine (reads of the cache / PE tables may precede the PDL wait);
            // 0 (stand-alone l2d_kv_attn): fully serialised launch, the caller may have just written the cache
};
int kv_attn_launch(const KvAttnParams& p, cudaStream_t stream);
bool kv_attn_mma_supported(const KvAttnParams& p);
int kv_attn_mma_launch(const KvAttnParams& p, cudaStream_t stream);

// warm-up (bidirectional) temporal attention of one clip + fill of cache slots 0..frames-1 (kv_warmup.cu)
struct WarmupAttnParams {
  const __half* q;
  const __half* k;
  const __half* v;
  int64_t ld;          // rows f*hw + pixel
  __half* cache_row;   // [2, hw, L, C]: one denoise row of the module's cache
  const __half* q_pe;
  const __half* k_pe;
  const __half* v_pe;
  int64_t pe_ld;
  __half* out;
  int64_t ldo;
  int frames, hw, L, C, heads;
  float scale;
};
int warmup_attn_launch(const WarmupAttnParams& p, cudaStream_t stream);

int groupnorm_launch(const __half* x1, int c1, const __half* x2, int c2, const __half* gamma, const __half* beta,
                     __half* y, float* ws, int n_img, int h, int w, int G, float eps, int silu, int mode, int stride,
                     cudaStream_t st);

int gemm_pick_tile_n(int m, int n, int k);
void gemm_weights_constant(bool on);   // thread-local: the following GEMM launches read engine-owned (constant) weights
struct GemmConstWeights {
  GemmConstWeights() { gemm_weights_constant(true); }
  ~GemmConstWeights() { gemm_weights_constant(false); }
};
// LayerNorm folded into the GEMMs around it (attention.py:243-268, motion_module.py:401-435: every LayerNorm's input is
// the output of a C x C projection and its output feeds a Linear):
//   producer  (stats_out != null): the epilogue also writes, per output row, (sum, sum of squares) of the fp16 values it
//             stores -- one float2 per (N tile, column half) = `gemm_stats_slots(m, n, k)` slots per row, summed by the
//             consumer in slot order (deterministic, no atomics)
//   consumer  (ln_stats != null): A holds the UN-normalised rows; W is gamma-scaled (w'[n,k] = gamma[k] w[n,k]) and
//             y[r,n] = rstd_r * (acc[r,n] - mean_r * ln_s[n]) + ln_b[n],  ln_s[n] = sum_k w'[n,k],  ln_b[n] = b[n] + sum_k beta[k] w[n,k]
//             which equals Linear(LayerNorm(x)) without materialising the normalised tensor (-1 launch, -2 passes over [M,C])
struct GemmFusion {
  float2* stats_out = nullptr;
  const float2* ln_stats = nullptr;
  int ln_slots = 0;
  const float* ln_s = nullptr;
  const float* ln_b = nullptr;
  int ln_c = 0;
  float ln_eps = 1e-5f;
};
int gemm_stats_slots(int m, int n, int k);
int gemm_launch(const __half* a, int64_t lda, const __half* w, int64_t ldw, __half* out, int64_t ldo, int m, int n, int k,
                const __half* bias, const __half* rowgroup_bias, int64_t rg_ld, int rows_per_group,
                const __half* residual, int64_t ldr, int act, int force_bn, cudaStream_t st, const GemmFusion* fx = nullptr);
// in place: w[n,k] *= gamma[k] (fp16 rounding), ln_s[n] = sum_k w'[n,k], ln_b[n] = bias[n] + sum_k beta[k] w[n,k]
int ln_fold_weights(__half* w, const __half* bias, const __half* gamma, const __half* beta, float* ln_s, float* ln_b, int n, int k,
                    cudaStream_t st);

bool conv3x3_implicit_supported(int n_img, int h, int w, int cin);
int conv3x3_launch(const __half* x, int n_img, int h, int w, int cin, const __half* weight, __half* out, int64_t ldo, int cout,
                   const __half* bias, const __half* rowgroup_bias, int64_t rg_ld, int rows_per_group,
                   const __half* residual, int64_t ldr, int act, cudaStream_t st);

int attention_launch(const __half* q, int64_t ldq, const __half* k, int64_t ldk, const __half* v, int64_t ldv, __half* o,
                     int64_t ldo, int batch, int heads, int sq, int skv, int hd, cudaStream_t st);

// tcgen05 / TMEM / TMA flash attention (flash_tcgen05.cu): hd 40 / 80, sq % 128 == 0; everything else takes the legacy
// mma.sync kernel in flash_attn.cu
bool attention_tcgen05_supported(const void* q, const void* k, const void* v, const void* o, int64_t ldq, int64_t ldk, int64_t ldv,
                                 int64_t ldo, int sq, int skv, int hd);
int attention_tcgen05_launch(const __half* q, int64_t ldq, const __half* k, int64_t ldk, const __half* v, int64_t ldv, __half* o,
                             int64_t ldo, int batch, int heads, int sq, int skv, int hd, cudaStream_t st);

// engine.cu internals used by the device-resident stream (stream_state.cu)
int unet_run_eager(::l2d_unet* u, const l2d_unet_step_args* a, cudaStream_t st);   // enqueue one step on st (capturable)
// time-embedding MLP + stacked time_emb_proj + stacked cross-attention K|V projection: once per (timestep, prompt)
int64_t unet_consts_epoch(const ::l2d_unet* u);
int unet_prepare_constants(::l2d_unet* u, const int64_t* timestep, const void* encoder_hidden_states, cudaStream_t st);
void unet_geometry(const ::l2d_unet* u, int* n_rows, int* h, int* w, int* window, int* n_kv, int* ctx_len, int* ctx_dim,
                   int* warmup_frames);

}  // namespace l2d
