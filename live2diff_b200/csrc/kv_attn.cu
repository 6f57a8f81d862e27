// K1 -- temporal KV-cache attention (StreamTemporalAttention core).
//
// Reference semantics: live2diff/animatediff/models/stream_motion_module.py:117-147 (slot append,
// PE-by-index add, rounding of q+pe / K+pe / V+pe to fp16) and :172-194 (per-row additive mask,
// q_len == 1 SDPA over the L-slot window).  Math spec: SURVEY.md Appendix D.
//
// This is an HBM-bandwidth kernel (~1 flop/byte, q_len = 1): per (row n, pixel p) it streams the
// pixel's contiguous K window [L,C] and V window [L,C] exactly once with 128-bit loads that bypass
// L1, writes the new k/v into slot u_n, and never materialises K+pe / V+pe / the repeated mask.
// Thread <-> one 8-channel (16 B) column chunk of one pixel; hd % 8 == 0 so a chunk never straddles
// heads.  Per-head reduction of the q.k partials goes through shared memory; softmax and the P.V
// accumulation are per-thread in fp32.
#include "ops.cuh"

namespace l2d {

constexpr int KV_CH = 16;  // slots handled per register chunk


__global__ void __launch_bounds__(320, 1) kv_attn_kernel(const KvAttnParams p) {
  extern __shared__ float smem[];
  const int L = p.L, T = p.T, P = p.P;
  float* s_part = smem;                          // [P][L][T]
  float* s_sc = s_part + (size_t)P * L * T;      // [P][heads][L]
  float* s_mask = s_sc + (size_t)P * p.heads * L;  // [L]
  int* s_pi = reinterpret_cast<int*>(s_mask + L);  // [L]
  __shared__ int s_u;

  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  if (tid < L) {
    s_pi[tid] = static_cast<int>(p.pe_idx[(size_t)n * L + tid]);
    s_mask[tid] = __half2float(p.mask[(size_t)n * L + tid]);
  }
  if (tid == 0) s_u = static_cast<int>(p.update_idx[n]);
  __syncthreads();
  const int u = s_u;

  const int pl = tid / T;
  const int c = tid - pl * T;
  const int pixel = blockIdx.x * P + pl;
  const bool active = (pl < P) && (pixel < p.hw);

  uint4 vreg[KV_CH];
  uint4 vnew = make_uint4(0, 0, 0, 0);
  const __half* vbase = nullptr;
  if (active) {
    const size_t row = (size_t)n * p.hw + pixel;
    const size_t win = (size_t)L * p.C;
    __half* kbase = p.cache + (((size_t)n * 2) * p.hw + pixel) * win + (size_t)c * 8;
    __half* vb = kbase + (size_t)p.hw * win;
    vbase = vb;

    const uint4 knew = ldg_cached(p.k_new + row * p.ld + (size_t)c * 8);
    vnew = ldg_cached(p.v_new + row * p.ld + (size_t)c * 8);
    uint4 qv = ldg_cached(p.q + row * p.ld + (size_t)c * 8);
    qv = hadd8(qv, ldg_cached(p.q_pe + (size_t)s_pi[u] * p.pe_ld + (size_t)c * 8));   // q + Q_pe[pi[u]] -> fp16
    float qf[8];
    unpack8(qv, qf);

    // ---- K window: q.k partials for this 8-channel chunk --------------------------------------
    for (int j0 = 0; j0 < L; j0 += KV_CH) {
      uint4 kreg[KV_CH];
#pragma unroll
      for (int jj = 0; jj < KV_CH; ++jj) {
        const int j = j0 + jj;
        if (j < L && j != u && s_mask[j] > -INFINITY) kreg[jj] = ldg_stream(kbase + (size_t)j * p.C);
      }
      if (j0 == 0) {
        // V loads of the first chunk are issued now so their latency overlaps the reductions below
#pragma unroll
        for (int jj = 0; jj < KV_CH; ++jj) {
          const int j = jj;
          if (j < L && j != u && s_mask[j] > -INFINITY) vreg[jj] = ldg_stream(vb + (size_t)j * p.C);
        }
      }
#pragma unroll
      for (int jj = 0; jj < KV_CH; ++jj) {
        const int j = j0 + jj;
        if (j < L) {
          float part = 0.f;
          if (s_mask[j] > -INFINITY) {
            uint4 kk = (j == u) ? knew : kreg[jj];
            kk = hadd8(kk, ldg_cached(p.k_pe + (size_t)s_pi[j] * p.pe_ld + (size_t)c * 8));  // K + K_pe -> fp16
            float kf[8];
            unpack8(kk, kf);
#pragma unroll
            for (int e = 0; e < 8; ++e) part = fmaf(qf[e], kf[e], part);
          }
          s_part[((size_t)pl * L + j) * T + c] = part;
        }
      }
    }
    // PE-free append (stream_motion_module.py:117-119)
    *reinterpret_cast<uint4*>(kbase + (size_t)u * p.C) = knew;
    *reinterpret_cast<uint4*>(vb + (size_t)u * p.C) = vnew;
  }
  __syncthreads();

  // ---- per-head scores: sum the hd/8 chunk partials, scale, add mask ---------------------------
  {
    const int total = P * p.heads * L;
    for (int idx = tid; idx < total; idx += blockDim.x) {
      const int j = idx % L;
      const int h = (idx / L) % p.heads;
      const int pl2 = idx / (L * p.heads);
      const float* src = s_part + ((size_t)pl2 * L + j) * T + h * p.hd8;
      float s = 0.f;
      for (int i = 0; i < p.hd8; ++i) s += src[i];
      s_sc[idx] = s * p.scale + s_mask[j];
    }
  }
  __syncthreads();
  if (!active) return;

  // ---- softmax over the window (fp32) + P.V ----------------------------------------------------
  const int h = c / p.hd8;
  const float* sc = s_sc + ((size_t)pl * p.heads + h) * L;
  float mx = -INFINITY;
  for (int j = 0; j < L; ++j) mx = fmaxf(mx, sc[j]);
  float denom = 0.f;
  for (int j = 0; j < L; ++j) denom += __expf(sc[j] - mx);
  const float inv = 1.f / denom;

  float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j0 = 0; j0 < L; j0 += KV_CH) {
    if (j0 > 0) {
#pragma unroll
      for (int jj = 0; jj < KV_CH; ++jj) {
        const int j = j0 + jj;
        if (j < L && j != u && s_mask[j] > -INFINITY) vreg[jj] = ldg_stream(vbase + (size_t)j * p.C);
      }
    }
#pragma unroll
    for (int jj = 0; jj < KV_CH; ++jj) {
      const int j = j0 + jj;
      if (j < L && s_mask[j] > -INFINITY) {
        const float pj = __expf(sc[j] - mx) * inv;
        uint4 vv = (j == u) ? vnew : vreg[jj];
        vv = hadd8(vv, ldg_cached(p.v_pe + (size_t)s_pi[j] * p.pe_ld + (size_t)c * 8));  // V + V_pe -> fp16
        float vf[8];
        unpack8(vv, vf);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(pj, vf[e], o[e]);
      }
    }
  }
  const size_t row = (size_t)n * p.hw + pixel;
  *reinterpret_cast<uint4*>(p.out + row * p.C + (size_t)c * 8) = pack8(o);
}

int kv_attn_launch(const KvAttnParams& p0, cudaStream_t stream) {
  KvAttnParams p = p0;
  p.T = p.C / 8;
  p.hd8 = (p.C / p.heads) / 8;
  int P = 320 / p.T;
  if (P < 1) P = 1;
  if (P > p.hw) P = p.hw;
  p.P = P;
  int threads = ((P * p.T + 31) / 32) * 32;
  if (threads > 320) return fail(L2D_ERR_INVALID, "kv_attn: channels too large for one block (C <= 2560)");
  p.scale = 1.0f / sqrtf((float)(p.C / p.heads));
  size_t smem = ((size_t)P * p.L * p.T + (size_t)P * p.heads * p.L + 2 * p.L) * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    L2D_CUDA(cudaFuncSetAttribute(kv_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid(ceil_div(p.hw, P), p.n_rows);
  kv_attn_kernel<<<grid, threads, smem, stream>>>(p);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

}  // namespace l2d

extern "C" int l2d_kv_attn(const void* q, const void* k_new, const void* v_new, int64_t qkv_ld, void* kv_cache,
                           const void* q_pe, const void* k_pe, const void* v_pe, const void* mask,
                           const int64_t* pe_idx, const int64_t* update_idx, void* out, int n_rows, int hw,
                           int window, int channels, int heads, void* stream) {
  using namespace l2d;
  L2D_CHECK_ARG(q && k_new && v_new && kv_cache && q_pe && k_pe && v_pe && mask && pe_idx && update_idx && out,
                "null pointer");
  L2D_CHECK_ARG(n_rows > 0 && hw > 0 && window > 0 && window <= 32, "need 0 < L <= 32");
  L2D_CHECK_ARG(heads > 0 && channels % heads == 0, "channels % heads != 0");
  L2D_CHECK_ARG(channels % 8 == 0 && (channels / heads) % 8 == 0, "need C % 8 == 0 and head_dim % 8 == 0");
  L2D_CHECK_ARG(qkv_ld % 8 == 0 && qkv_ld >= channels, "qkv_ld must be >= C and a multiple of 8");
  KvAttnParams p{};
  p.q = (const __half*)q; p.k_new = (const __half*)k_new; p.v_new = (const __half*)v_new; p.ld = qkv_ld;
  p.cache = (__half*)kv_cache; p.q_pe = (const __half*)q_pe; p.k_pe = (const __half*)k_pe;
  p.v_pe = (const __half*)v_pe; p.mask = (const __half*)mask; p.pe_idx = pe_idx; p.update_idx = update_idx;
  p.out = (__half*)out; p.pe_ld = channels; p.n_rows = n_rows; p.hw = hw; p.L = window; p.C = channels; p.heads = heads;
  return kv_attn_launch(p, (cudaStream_t)stream);
}
