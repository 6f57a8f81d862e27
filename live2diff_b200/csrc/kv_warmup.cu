// Warm-up temporal attention: the bidirectional attention of VersatileAttention.forward
// (live2diff/animatediff/models/motion_module.py:469-530) over the F warm-up frames of one clip, fused with the
// fill of the KV-cache sink slots that the streaming kernel K1 (kv_attn*.cu) reads afterwards.
//
// Per pixel d (sequence = the F frames of that pixel), reference order of operations:
//   cache[0, d, f, :] = k[f]        cache[1, d, f, :] = v[f]          (:488-489, PE-free)
//   q~ = q + to_q(pe[f])   k~ = k + to_k(pe[f])   v~ = v + to_v(pe[f])   (:491-499, fp16 adds)
//   out[f] = softmax_j(q~[f] . k~[j] / sqrt(hd)) v~[j]   per head, no mask  (:516)
// q/k/v are the bias-free projections already computed by the fused QKV GEMM: rows f*hw + d of [F*hw, ld].
// The PE tables are rows 0..F-1 of the engine's precomputed pe[:L] . [Wq;Wk;Wv]^T (an fp16 Linear like the reference).
//
// One CTA per pixel: stage q~/k~/v~ [F, C] in shared memory (writing the PE-free k/v to the cache on the way, 128-bit
// coalesced), F*F*heads dot products, fp32 softmax, P.V~.  This runs once per stream and denoise row, on 8 frames: it
// is HBM/latency-bound trivia next to the F-frame UNet around it, so it stays a plain CUDA-core kernel.
#include "ops.cuh"

namespace l2d {

namespace {

constexpr int WU_THREADS = 256;
constexpr int WU_MAX_F = 16;

__global__ void __launch_bounds__(WU_THREADS) warmup_attn_kernel(const WarmupAttnParams p) {
  extern __shared__ __align__(16) uint8_t wu_smem[];
  const int F = p.frames, C = p.C, T = C >> 3, hd = C / p.heads;
  __half* s_q = reinterpret_cast<__half*>(wu_smem);        // [F][C]
  __half* s_k = s_q + (size_t)F * C;
  __half* s_v = s_k + (size_t)F * C;
  float* s_p = reinterpret_cast<float*>(s_v + (size_t)F * C);   // [heads][F][F]
  const int pix = blockIdx.x, tid = threadIdx.x;

  // ---- stage: thread <-> one 16-byte chunk of one frame ----
  for (int i = tid; i < F * T; i += WU_THREADS) {
    const int f = i / T, c = (i - f * T) * 8;
    const size_t row = (size_t)f * p.hw + pix;
    const uint4 q = ldg_cached(p.q + row * p.ld + c);
    const uint4 k = ldg_cached(p.k + row * p.ld + c);
    const uint4 v = ldg_cached(p.v + row * p.ld + c);
    __half* kc = p.cache_row + (((size_t)0 * p.hw + pix) * p.L + f) * C + c;
    __half* vc = p.cache_row + (((size_t)1 * p.hw + pix) * p.L + f) * C + c;
    *reinterpret_cast<uint4*>(kc) = k;
    *reinterpret_cast<uint4*>(vc) = v;
    *reinterpret_cast<uint4*>(s_q + (size_t)f * C + c) = hadd8(q, ldg_cached(p.q_pe + (size_t)f * p.pe_ld + c));
    *reinterpret_cast<uint4*>(s_k + (size_t)f * C + c) = hadd8(k, ldg_cached(p.k_pe + (size_t)f * p.pe_ld + c));
    *reinterpret_cast<uint4*>(s_v + (size_t)f * C + c) = hadd8(v, ldg_cached(p.v_pe + (size_t)f * p.pe_ld + c));
  }
  __syncthreads();

  // ---- scores: one (head, i, j) dot product of length hd per thread-iteration ----
  const int n_dots = p.heads * F * F;
  for (int d = tid; d < n_dots; d += WU_THREADS) {
    const int h = d / (F * F), r = d - h * F * F, i = r / F, j = r - i * F;
    const __half2* a = reinterpret_cast<const __half2*>(s_q + (size_t)i * C + h * hd);
    const __half2* b = reinterpret_cast<const __half2*>(s_k + (size_t)j * C + h * hd);
    float acc = 0.f;
    for (int e = 0; e < hd / 2; ++e) {
      const float2 x = __half22float2(a[e]), y = __half22float2(b[e]);
      acc = fmaf(x.x, y.x, acc);
      acc = fmaf(x.y, y.y, acc);
    }
    s_p[d] = acc * p.scale;
  }
  __syncthreads();
  // ---- softmax over j (fp32), one (head, i) row per thread ----
  for (int r = tid; r < p.heads * F; r += WU_THREADS) {
    float* row = s_p + (size_t)r * F;
    float m = -INFINITY;
    for (int j = 0; j < F; ++j) m = fmaxf(m, row[j]);
    float s = 0.f;
    for (int j = 0; j < F; ++j) {
      const float e = __expf(row[j] - m);
      row[j] = e;
      s += e;
    }
    const float inv = 1.f / s;
    for (int j = 0; j < F; ++j) row[j] *= inv;
  }
  __syncthreads();
  // ---- out[i][c] = sum_j P[head(c)][i][j] v~[j][c]; thread <-> (frame i, channel pair) ----
  const int half_c = C >> 1;
  for (int o = tid; o < F * half_c; o += WU_THREADS) {
    const int i = o / half_c, c = (o - i * half_c) * 2, h = c / hd;
    const float* pr = s_p + ((size_t)h * F + i) * F;
    float ax = 0.f, ay = 0.f;
    for (int j = 0; j < F; ++j) {
      const float2 vv = __half22float2(*reinterpret_cast<const __half2*>(s_v + (size_t)j * C + c));
      ax = fmaf(pr[j], vv.x, ax);
      ay = fmaf(pr[j], vv.y, ay);
    }
    *reinterpret_cast<__half2*>(p.out + ((size_t)i * p.hw + pix) * p.ldo + c) = __floats2half2_rn(ax, ay);
  }
}

}  // namespace

int warmup_attn_launch(const WarmupAttnParams& p0, cudaStream_t stream) {
  WarmupAttnParams p = p0;
  if (p.frames < 1 || p.frames > WU_MAX_F || p.frames > p.L)
    return fail(L2D_ERR_INVALID, "warmup_attn: need 1 <= frames <= min(16, window)");
  if (p.C % 8 != 0 || p.heads <= 0 || p.C % p.heads != 0 || (p.C / p.heads) % 2 != 0)
    return fail(L2D_ERR_INVALID, "warmup_attn: need C % 8 == 0 and an even head_dim");
  p.scale = 1.0f / sqrtf((float)(p.C / p.heads));
  const size_t smem = (size_t)3 * p.frames * p.C * sizeof(__half) + (size_t)p.heads * p.frames * p.frames * sizeof(float);
  if (smem > 227 * 1024) return fail(L2D_ERR_INVALID, "warmup_attn: frames * C too large for shared memory");
  static size_t configured = 0;
  if (smem > configured) {
    L2D_CUDA(cudaFuncSetAttribute(warmup_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  warmup_attn_kernel<<<p.hw, WU_THREADS, smem, stream>>>(p);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

}  // namespace l2d

// C ABI: replaces VersatileAttention's q/k/v-to-output core + cache fill (motion_module.py:488-516) for one clip.
extern "C" int l2d_warmup_attn(const void* q, const void* k, const void* v, int64_t qkv_ld, void* kv_cache_row,
                               const void* q_pe, const void* k_pe, const void* v_pe, int64_t pe_ld, void* out,
                               int64_t out_ld, int frames, int hw, int window, int channels, int heads, void* stream) {
  using namespace l2d;
  L2D_CHECK_ARG(q && k && v && kv_cache_row && q_pe && k_pe && v_pe && out, "null pointer");
  L2D_CHECK_ARG(hw > 0 && window > 0 && frames > 0, "bad sizes");
  L2D_CHECK_ARG(qkv_ld % 8 == 0 && qkv_ld >= channels && out_ld % 2 == 0 && out_ld >= channels && pe_ld % 8 == 0,
                "leading dimensions must be >= C and 16-byte aligned");
  L2D_CHECK_ARG((uintptr_t)kv_cache_row % 16 == 0, "kv_cache_row must be 16-byte aligned");
  WarmupAttnParams p{};
  p.q = (const __half*)q; p.k = (const __half*)k; p.v = (const __half*)v; p.ld = qkv_ld;
  p.cache_row = (__half*)kv_cache_row; p.q_pe = (const __half*)q_pe; p.k_pe = (const __half*)k_pe;
  p.v_pe = (const __half*)v_pe; p.pe_ld = pe_ld; p.out = (__half*)out; p.ldo = out_ld;
  p.frames = frames; p.hw = hw; p.L = window; p.C = channels; p.heads = heads;
  return warmup_attn_launch(p, (cudaStream_t)stream);
}
