// K5/K6 -- GroupNorm(+SiLU)(+3x3 im2col) and LayerNorm on channels-last fp16 activations, plus the
// raw im2col / layout kernels.  All HBM/L2-bandwidth kernels: 128-bit accesses, fp32 statistics,
// warp-shuffle / fixed-order shared-memory reductions (no atomics on data: results are bit-reproducible).
//
// Reference semantics: nn.GroupNorm / InflatedGroupNorm per frame (resnet.py:68-76; eps 1e-5 in
// resnets/conv_norm_out, 1e-6 at transformer entries -- SURVEY A-9), nn.LayerNorm(eps 1e-5),
// F.silu, 3x3 / stride-2 / nearest-x2 convolutions of resnet.py:57-153 (the im2col here feeds the
// tcgen05 GEMM that performs the convolution's contraction).
#include "ops.cuh"

namespace l2d {

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, values kept in registers (C <= 8*32*MAXCH)
// ---------------------------------------------------------------------------------------------
constexpr int LN_MAXCH = 10;  // 16B chunks per lane -> C <= 2560

__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma,
                                                        const __half* __restrict__ beta, __half* __restrict__ y,
                                                        int rows, int C, float eps) {
  pdl_launch();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nch = C >> 3;
  const __half* xr = x + (size_t)warp * C;
  float v[LN_MAXCH][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXCH; ++i) {
    const int ch = lane + i * 32;
    if (ch < nch) {
      unpack8(ldg_act(xr + ch * 8), v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[i][e];
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXCH; ++i) {
    const int ch = lane + i * 32;
    if (ch < nch) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[i][e] - mean;
        q = fmaf(d, d, q);
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  __half* yr = y + (size_t)warp * C;
#pragma unroll
  for (int i = 0; i < LN_MAXCH; ++i) {
    const int ch = lane + i * 32;
    if (ch < nch) {
      float g[8], b[8], o[8];
      unpack8(ldg_cached(gamma + ch * 8), g);
      unpack8(ldg_cached(beta + ch * 8), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[i][e] - mean) * rstd * g[e] + b[e];
      *reinterpret_cast<uint4*>(yr + ch * 8) = pack8(o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm statistics: grid (chunks, N).  Each block reduces a slab of pixels x all channels to
// per-group (mean, M2) partials; the apply kernel merges the partials (Chan's formula).
// ---------------------------------------------------------------------------------------------
constexpr int GN_MAX_SPLIT = 128;

struct GnSrc {
  const __half* x1;
  const __half* x2;
  int c1, c2;  // channels of each source (c2 may be 0)
};

__device__ __forceinline__ uint4 gn_load_chunk(const GnSrc& s, size_t pix, int ch8) {
  const int c = ch8 * 8;
  if (c < s.c1) return ldg_act(s.x1 + pix * s.c1 + c);
  return ldg_act(s.x2 + pix * s.c2 + (c - s.c1));
}

// workspace layout (floats): partial [N][GN_MAX_SPLIT][G][2] (mean, M2 of each slab) | ab [N][2][GN_MAX_C] (per-channel
// scale, shift) | counters [N] (ints, zero between launches).  The last slab-block of an image to finish merges the
// slabs (Chan's formula, done as two parallel weighted sums) and writes the per-channel scale/shift, so the apply
// kernel is a pure streaming pass.
constexpr int GN_MAX_C = 4096;

__device__ __forceinline__ float* gn_ab(float* ws, int n_img, int G, int n) {
  return ws + (size_t)n_img * GN_MAX_SPLIT * G * 2 + (size_t)n * 2 * GN_MAX_C;
}
__device__ __forceinline__ int* gn_counters(float* ws, int n_img, int G) {
  return reinterpret_cast<int*>(ws + (size_t)n_img * GN_MAX_SPLIT * G * 2 + (size_t)n_img * 2 * GN_MAX_C);
}

// Merge the slab partials of image n (Chan's formula as two weighted sums over the slabs, slab order fixed) and write the
// per-channel scale / shift to ab[c] / ab[ab_stride + c] (global or shared).  Latency-bound: every lane first issues ALL
// of its partial loads (<= 8 independent 8-byte loads), so the merge costs one L2 round trip instead of one per slab batch.
// A group is handled by `lpg` lanes (16 when there are more groups than warps: two groups per warp, one pass).
__device__ __forceinline__ void gn_merge_groups(const float* partial, int n, int G, int split, int per, int hw, int cpg, float eps,
                                                const __half* __restrict__ gamma, const __half* __restrict__ beta, float* ab,
                                                int ab_stride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int lpg = G > nwarps ? 16 : 32, gpw = 32 / lpg;
  const int sub = lane / lpg, l = lane - sub * lpg;
  constexpr int MAXI = GN_MAX_SPLIT / 16;
  for (int gb = warp * gpw; gb < G; gb += nwarps * gpw) {     // warp-uniform trip count
    const int g = gb + sub;
    const bool gv = g < G;
    float cnt_s[MAXI], mu_s[MAXI], m2_s[MAXI];
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int s2 = l + lpg * i;
      const bool v = gv && s2 < split;
      const int q0 = s2 * per, q1 = min(hw, q0 + per);
      cnt_s[i] = v ? (float)max(q1 - q0, 0) * (float)cpg : 0.f;
      float2 pm = make_float2(0.f, 0.f);
      if (v) pm = __ldcg(reinterpret_cast<const float2*>(partial + (((size_t)n * GN_MAX_SPLIT + s2) * G + g) * 2));
      mu_s[i] = pm.x;
      m2_s[i] = pm.y;
    }
    float wsum = 0.f, wmean = 0.f;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      wsum += cnt_s[i];
      wmean = fmaf(cnt_s[i], mu_s[i], wmean);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {     // offsets < 16 stay inside a 16-lane half; the 32-lane case adds offset 16 below
      wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
      wmean += __shfl_xor_sync(0xffffffffu, wmean, o);
    }
    if (lpg == 32) {
      wsum += __shfl_xor_sync(0xffffffffu, wsum, 16);
      wmean += __shfl_xor_sync(0xffffffffu, wmean, 16);
    }
    const float mean = wsum > 0.f ? wmean / wsum : 0.f;
    float m2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const float d = mu_s[i] - mean;
      m2 += m2_s[i] + cnt_s[i] * d * d;
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    if (lpg == 32) m2 += __shfl_xor_sync(0xffffffffu, m2, 16);
    if (gv) {
      const float rstd = rsqrtf(m2 / wsum + eps);
      for (int i = l; i < cpg; i += lpg) {
        const int c = g * cpg + i;
        const float ga = __half2float(gamma[c]) * rstd;
        ab[c] = ga;
        ab[ab_stride + c] = __half2float(beta[c]) - mean * ga;
      }
    }
  }
}

__global__ void __launch_bounds__(512) groupnorm_stats_kernel(GnSrc src, const __half* __restrict__ gamma,
                                                               const __half* __restrict__ beta, float* __restrict__ ws,
                                                               int hw, int G, int split, float eps) {
  extern __shared__ float sm[];  // [2][R][C]: per pixel-lane, per channel (sum | sumsq); lane 0's rows end up holding the totals
  __shared__ int s_last;
  const int C = src.c1 + src.c2;
  const int nch = C >> 3;
  const int n = blockIdx.y, sp = blockIdx.x, n_img = gridDim.y;
  const int per = (hw + split - 1) / split;
  const int p0 = sp * per, p1 = min(hw, p0 + per);
  const int R = blockDim.x / nch;  // pixel lanes
  float* s_sum = sm;
  float* s_sq = sm + (size_t)R * C;
  pdl_launch();
  pdl_wait();
  const int r = threadIdx.x / nch, ch = threadIdx.x - r * nch;
  if (r < R) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = p0 + r; p < p1; p += 4 * R) {
      uint4 raw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)   // 4 independent loads in flight
        raw[i] = (p + i * R < p1) ? gn_load_chunk(src, (size_t)n * hw + p + i * R, ch) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v[8];
        unpack8(raw[i], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          a[e] += v[e];
          b[e] = fmaf(v[e], v[e], b[e]);
        }
      }
    }
    // fixed-order reduction (no atomics): every pixel lane parks its partials, then one thread per channel adds the
    // lanes in lane order -- the same sums bit for bit on every run
    float* ds = s_sum + (size_t)r * C + ch * 8;
    float* dq = s_sq + (size_t)r * C + ch * 8;
    *reinterpret_cast<float4*>(ds) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(ds + 4) = make_float4(a[4], a[5], a[6], a[7]);
    *reinterpret_cast<float4*>(dq) = make_float4(b[0], b[1], b[2], b[3]);
    *reinterpret_cast<float4*>(dq + 4) = make_float4(b[4], b[5], b[6], b[7]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float ts = s_sum[c], tq = s_sq[c];
    for (int r2 = 1; r2 < R; ++r2) {
      ts += s_sum[(size_t)r2 * C + c];
      tq += s_sq[(size_t)r2 * C + c];
    }
    s_sum[c] = ts;
    s_sq[c] = tq;
  }
  __syncthreads();
  const int cpg = C / G;
  const float cnt = (float)(p1 - p0) * (float)cpg;
  float* partial = ws;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int i = 0; i < cpg; ++i) {
      s += s_sum[g * cpg + i];
      q += s_sq[g * cpg + i];
    }
    const float mean = cnt > 0.f ? s / cnt : 0.f;
    const float m2 = fmaxf(q - s * mean, 0.f);
    float* o = partial + (((size_t)n * GN_MAX_SPLIT + sp) * G + g) * 2;
    o[0] = mean;
    o[1] = m2;
  }
  // ---- last slab of image n merges all slabs and publishes per-channel scale/shift ----
  __threadfence();
  __syncthreads();
  int* counters = gn_counters(ws, n_img, G);
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(&counters[n], 1);
    s_last = prev == split - 1;
    if (s_last) counters[n] = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  gn_merge_groups(partial, n, G, split, per, hw, cpg, eps, gamma, beta, gn_ab(ws, n_img, G, n), GN_MAX_C);
}

// Apply: y = act(x * scale_c + shift_c); mode 0 writes NHWC, mode 1 writes the 3x3 im2col matrix (pad 1, given
// stride) of the normalised tensor.
struct GnApplyParams {
  GnSrc src;
  const float* ab;   // [N][2][GN_MAX_C]
  __half* y;
  int h, w;
  int silu, mode, stride;
};

__device__ __forceinline__ void gn_norm8(float (&v)[8], const float* __restrict__ ab, int c0, int silu) {
  // (written by the statistics kernel right before this one: coherent L2 reads, see ldg_act)
  const float4 a0 = __ldcg(reinterpret_cast<const float4*>(ab + c0)), a1 = __ldcg(reinterpret_cast<const float4*>(ab + c0 + 4));
  const float4 b0 = __ldcg(reinterpret_cast<const float4*>(ab + GN_MAX_C + c0));
  const float4 b1 = __ldcg(reinterpret_cast<const float4*>(ab + GN_MAX_C + c0 + 4));
  const float A[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  const float B[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float t = fmaf(v[e], A[e], B[e]);
    v[e] = silu ? silu_f(t) : t;
  }
}

__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const GnApplyParams p) {
  pdl_launch();
  pdl_wait();
  const int C = p.src.c1 + p.src.c2;
  const int hw = p.h * p.w;
  const int n = blockIdx.y;
  const float* ab = p.ab + (size_t)n * 2 * GN_MAX_C;
  const int nch = C >> 3;
  if (p.mode == 0) {
    const size_t total = (size_t)hw * nch;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
      const int ch = (int)(i % nch);
      const size_t pix = (size_t)n * hw + i / nch;
      float v[8];
      unpack8(gn_load_chunk(p.src, pix, ch), v);
      gn_norm8(v, ab, ch * 8, p.silu);
      *reinterpret_cast<uint4*>(p.y + pix * C + ch * 8) = pack8(v);
    }
  } else {
    const int ho = p.h / p.stride, wo = p.w / p.stride;
    const size_t total = (size_t)ho * wo * 9 * nch;
    const size_t K = (size_t)9 * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
      const int ch = (int)(i % nch);
      const int tap = (int)((i / nch) % 9);
      const size_t opix = i / ((size_t)nch * 9);
      const int oy = (int)(opix / wo), ox = (int)(opix % wo);
      const int iy = oy * p.stride + tap / 3 - 1, ix = ox * p.stride + tap % 3 - 1;
      uint4 o = make_uint4(0, 0, 0, 0);
      if (iy >= 0 && iy < p.h && ix >= 0 && ix < p.w) {
        float v[8];
        unpack8(gn_load_chunk(p.src, (size_t)n * hw + (size_t)iy * p.w + ix, ch), v);
        gn_norm8(v, ab, ch * 8, p.silu);
        o = pack8(v);
      }
      *reinterpret_cast<uint4*>(p.y + ((size_t)n * ho * wo + opix) * K + (size_t)tap * C + ch * 8) = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// One-kernel GroupNorm (mode 0): grid (split, N) with split * N <= #SMs, so every CTA is resident at once and the grid can
// meet at a barrier in global memory.  Each CTA pulls its slab of pixels x all channels into shared memory with ONE bulk
// copy per source tensor (cp.async.bulk, the slab is contiguous in NHWC), reduces it to per-group (mean, M2) partials in
// the same fixed order as the statistics kernel, publishes them, waits until the other slabs of its image have done the
// same, merges the partials itself (every CTA computes the identical scale / shift -- no broadcast step) and normalises
// its slab out of shared memory.  The activation is read from HBM/L2 once, and one launch plus the scale/shift round trip
// through global memory are gone.  Bit-identical to the two-kernel path when `split` is the same; deterministic always.
// ---------------------------------------------------------------------------------------------
struct GnFusedParams {
  GnSrc src;
  const __half* gamma;
  const __half* beta;
  __half* y;
  float* ws;
  int hw, G, split, per, silu;
  float eps;
};

__device__ __forceinline__ uint32_t gn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512) groupnorm_fused_kernel(const GnFusedParams p) {
  extern __shared__ __align__(128) uint8_t gsm[];
  __shared__ __align__(8) uint64_t s_bar;
  const int c1 = p.src.c1, c2 = p.src.c2, C = c1 + c2, nch = C >> 3, nch1 = c1 >> 3;
  const int n = blockIdx.y, sp = blockIdx.x, n_img = gridDim.y;
  const int p0 = sp * p.per, p1 = min(p.hw, p0 + p.per), npix = p1 - p0;
  const int R = blockDim.x / nch;  // pixel lanes
  __half* slab1 = reinterpret_cast<__half*>(gsm);
  __half* slab2 = slab1 + (size_t)p.per * c1;
  float* s_sum = reinterpret_cast<float*>(gsm + (((size_t)p.per * C * 2 + 127) & ~(size_t)127));   // [R][C]
  float* s_sq = s_sum + (size_t)R * C;                                                             // [R][C]
  const uint32_t bar = gn_smem_u32(&s_bar);
  pdl_launch();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();   // barrier initialised before anyone arms or polls it (polls before the expect_tx below see the phase pending)
  pdl_wait();        // the activation belongs to the previous kernel until here
  if (threadIdx.x == 0) {
    const uint32_t b1 = (uint32_t)npix * c1 * 2, b2 = (uint32_t)npix * c2 * 2;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b1 + b2) : "memory");
    constexpr uint32_t PIECE = 32768;   // bulk copies in <= 32 KB pieces
    const uint8_t* g1 = reinterpret_cast<const uint8_t*>(p.src.x1 + ((size_t)n * p.hw + p0) * c1);
    for (uint32_t o = 0; o < b1; o += PIECE) {
      const uint32_t sz = min(PIECE, b1 - o);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       gn_smem_u32(slab1) + o),
                   "l"(g1 + o), "r"(sz), "r"(bar)
                   : "memory");
    }
    if (c2 > 0) {
      const uint8_t* g2 = reinterpret_cast<const uint8_t*>(p.src.x2 + ((size_t)n * p.hw + p0) * c2);
      for (uint32_t o = 0; o < b2; o += PIECE) {
        const uint32_t sz = min(PIECE, b2 - o);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         gn_smem_u32(slab2) + o),
                     "l"(g2 + o), "r"(sz), "r"(bar)
                     : "memory");
      }
    }
  }
  {
    uint32_t done = 0, spins = 0;
    while (true) {
      asm volatile(
          "{\n\t.reg .pred q;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, q;\n\t}"
          : "=r"(done)
          : "r"(bar)
          : "memory");
      if (done) break;
      if (++spins > (1u << 26)) __trap();   // never hang the GPU on a protocol bug
    }
  }
  auto slab_chunk = [&](int pix, int ch) -> uint4 {   // 8 channels of pixel `pix` of the slab
    return ch < nch1 ? *reinterpret_cast<const uint4*>(slab1 + (size_t)pix * c1 + ch * 8)
                     : *reinterpret_cast<const uint4*>(slab2 + (size_t)pix * c2 + (ch - nch1) * 8);
  };
  const int r = threadIdx.x / nch, ch = threadIdx.x - r * nch;
  if (r < R) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // same per-thread pixel order as groupnorm_stats_kernel (r, r + R, r + 2R, ...): same sums bit for bit
    for (int q = r; q < npix; q += R) {
      float v[8];
      unpack8(slab_chunk(q, ch), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        a[e] += v[e];
        b[e] = fmaf(v[e], v[e], b[e]);
      }
    }
    float* ds = s_sum + (size_t)r * C + ch * 8;
    float* dq = s_sq + (size_t)r * C + ch * 8;
    *reinterpret_cast<float4*>(ds) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(ds + 4) = make_float4(a[4], a[5], a[6], a[7]);
    *reinterpret_cast<float4*>(dq) = make_float4(b[0], b[1], b[2], b[3]);
    *reinterpret_cast<float4*>(dq + 4) = make_float4(b[4], b[5], b[6], b[7]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float ts = s_sum[c], tq = s_sq[c];
    for (int r2 = 1; r2 < R; ++r2) {
      ts += s_sum[(size_t)r2 * C + c];
      tq += s_sq[(size_t)r2 * C + c];
    }
    s_sum[c] = ts;
    s_sq[c] = tq;
  }
  __syncthreads();
  const int cpg = C / p.G;
  const float cnt = (float)npix * (float)cpg;
  float* partial = p.ws;
  for (int g = threadIdx.x; g < p.G; g += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int i = 0; i < cpg; ++i) {
      s += s_sum[g * cpg + i];
      q += s_sq[g * cpg + i];
    }
    const float mean = cnt > 0.f ? s / cnt : 0.f;
    const float m2 = fmaxf(q - s * mean, 0.f);
    float* o = partial + (((size_t)n * GN_MAX_SPLIT + sp) * p.G + g) * 2;
    o[0] = mean;
    o[1] = m2;
  }
  // ---- grid barrier over the slabs of image n: arrive, then wait for the others' partials ----
  int* arrive = gn_counters(p.ws, n_img, p.G) + n;
  int* depart = gn_counters(p.ws, n_img, p.G) + n_img + n;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(arrive, 1);
    uint32_t spins = 0;
    while (true) {
      int seen;
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(arrive) : "memory");
      if (seen >= p.split) break;
      __nanosleep(32);
      if (++spins > (1u << 24)) __trap();   // a slab that is not resident would be a launch-geometry bug: fail, do not hang
    }
  }
  __syncthreads();
  // every CTA merges the partials of its image itself (scale / shift land in shared memory, over the reduction scratch)
  float* ab = s_sum;
  gn_merge_groups(partial, n, p.G, p.split, p.per, p.hw, cpg, p.eps, p.gamma, p.beta, ab, C);
  __syncthreads();
  if (threadIdx.x == 0) {   // the last CTA of the image to get here re-arms the two counters for the next launch
    const int prev = atomicAdd(depart, 1);
    if (prev == p.split - 1) {
      *arrive = 0;
      *depart = 0;
    }
  }
  const int total = npix * nch;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int q = i / nch, c8 = i - q * nch;
    float v[8];
    unpack8(slab_chunk(q, c8), v);
    const float4 a0 = *reinterpret_cast<const float4*>(ab + c8 * 8), a1 = *reinterpret_cast<const float4*>(ab + c8 * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(ab + C + c8 * 8), b1 = *reinterpret_cast<const float4*>(ab + C + c8 * 8 + 4);
    const float A[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float B[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float t = fmaf(v[e], A[e], B[e]);
      v[e] = p.silu ? silu_f(t) : t;
    }
    *reinterpret_cast<uint4*>(p.y + ((size_t)n * p.hw + p0 + q) * C + c8 * 8) = pack8(v);
  }
}

// ---------------------------------------------------------------------------------------------
// raw im2col (no normalisation): stride 1/2, optional nearest x2 upsample, optional SiLU
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col3x3_kernel(const __half* __restrict__ x, __half* __restrict__ y, int n_img,
                                                        int h, int w, int C, int stride, int up, int silu) {
  const int hs = up ? 2 * h : h, ws = up ? 2 * w : w;  // size of the (virtual) conv input
  const int ho = hs / stride, wo = ws / stride;
  const int nch = C >> 3;
  const size_t total = (size_t)n_img * ho * wo * 9 * nch;
  const size_t K = (size_t)9 * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % nch);
    const int tap = (int)((i / nch) % 9);
    const size_t opix = i / ((size_t)nch * 9);
    const int n = (int)(opix / ((size_t)ho * wo));
    const int rem = (int)(opix % ((size_t)ho * wo));
    const int oy = rem / wo, ox = rem % wo;
    int iy = oy * stride + tap / 3 - 1, ix = ox * stride + tap % 3 - 1;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < hs && ix >= 0 && ix < ws) {
      if (up) {
        iy >>= 1;
        ix >>= 1;
      }
      o = ldg_act(x + (((size_t)n * h + iy) * w + ix) * C + ch * 8);
      if (silu) {
        float v[8];
        unpack8(o, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = silu_f(v[e]);
        o = pack8(v);
      }
    }
    *reinterpret_cast<uint4*>(y + opix * K + (size_t)tap * C + ch * 8) = o;
  }
}

// 4-channel NCHW latent -> [N*h*w, 64]; cols 0..35 = (tap, channel), 36..63 = 0
__global__ void __launch_bounds__(256) im2col3x3_nchw4_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                              int n_img, int h, int w) {
  const size_t total = (size_t)n_img * h * w * 8;  // 8 chunks of 8 columns per row
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int chunk = (int)(i & 7);
    const size_t pix = i >> 3;
    const int n = (int)(pix / ((size_t)h * w));
    const int rem = (int)(pix % ((size_t)h * w));
    const int oy = rem / w, ox = rem % w;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = chunk * 8 + e;
      float t = 0.f;
      if (col < 36) {
        const int tap = col >> 2, c = col & 3;
        const int iy = oy + tap / 3 - 1, ix = ox + tap % 3 - 1;
        if (iy >= 0 && iy < h && ix >= 0 && ix < w) t = __half2float(x[(((size_t)n * 4 + c) * h + iy) * w + ix]);
      }
      v[e] = t;
    }
    *reinterpret_cast<uint4*>(y + pix * 64 + chunk * 8) = pack8(v);
  }
}

// ---------------------------------------------------------------------------------------------
// NCHW <-> NHWC (32x32 smem tile transpose)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_kernel(const __half* __restrict__ x, const __half* __restrict__ res,
                                                        __half* __restrict__ y, int rows, int cols) {
  // x: [batch][rows][cols] -> y: [batch][cols][rows]; res (optional) has y's layout
  __shared__ __half tile[32][33];
  const size_t base = (size_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    if (r < rows && c < cols) tile[i][tx] = x[base + (size_t)r * cols + c];
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) {
      const size_t o = base + (size_t)c * rows + r;
      __half v = tile[tx][i];
      if (res) v = __hadd(v, res[o]);
      y[o] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
// The statistics pass is latency-bound (a launch moves <= 10 MB): pick the slab so that every thread issues ONE batch
// of <= 4 independent 16-byte loads -- pixels per block = 4 x (pixel lanes per block) -- instead of looping.
static int pick_split(int hw, int nch, int threads) {
  const int lanes = threads / nch > 0 ? threads / nch : 1;
  const int per_block = 4 * lanes;
  int s = (hw + per_block - 1) / per_block;
  if (s < 1) s = 1;
  if (s > GN_MAX_SPLIT) s = GN_MAX_SPLIT;
  return s;
}

int groupnorm_launch(const __half* x1, int c1, const __half* x2, int c2, const __half* gamma, const __half* beta,
                     __half* y, float* ws, int n_img, int h, int w, int G, float eps, int silu, int mode, int stride,
                     cudaStream_t st) {
  const int C = c1 + c2, hw = h * w, nch = C / 8;
  const int threads = 512;
  const int split = pick_split(hw, nch > 0 ? nch : 1, threads);
  GnSrc src{x1, x2, c1, c2};
  if (nch > threads || C > GN_MAX_C) return fail(L2D_ERR_INVALID, "groupnorm: C > 4096");
  // ---- one-kernel path: every slab of every image resident at once (grid <= #SMs), slab + scratch in shared memory ----
  static const int fused_on = [] { const char* e = getenv("L2D_GN_FUSED"); return e ? atoi(e) : 1; }();
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    L2D_CUDA(cudaGetDevice(&dev));
    L2D_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (fused_on && mode == 0 && nch > 0 && n_img <= num_sms && ((c1 | c2) & 7) == 0) {
    int fs = std::min(std::min(num_sms / n_img, GN_MAX_SPLIT), hw);
    const int per = (hw + fs - 1) / fs;
    fs = (hw + per - 1) / per;   // no empty slabs
    const int lanes_f = threads / nch;
    const size_t slab = ((size_t)per * C * 2 + 127) & ~(size_t)127;
    const size_t fsmem = slab + std::max((size_t)2 * lanes_f * C, (size_t)2 * C) * sizeof(float);
    if (fsmem <= 200 * 1024) {
      static size_t cfg_fused = 0;
      if (fsmem > 48 * 1024 && fsmem > cfg_fused) {
        L2D_CUDA(cudaFuncSetAttribute(groupnorm_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        cfg_fused = 200 * 1024;
      }
      GnFusedParams fp{src, gamma, beta, y, ws, hw, G, fs, per, silu, eps};
      launch_pdl_if(pdl_family(4), groupnorm_fused_kernel, dim3(fs, n_img), dim3(threads), fsmem, st, fp);
      L2D_LAUNCH_CHECK();
      return L2D_OK;
    }
  }
  static size_t cfg_stats = 0;
  const int lanes = threads / (nch > 0 ? nch : 1) > 0 ? threads / (nch > 0 ? nch : 1) : 1;
  const size_t smem = (size_t)2 * lanes * C * sizeof(float);   // <= 32 KB: lanes * C <= 8 * threads
  if (smem > 48 * 1024 && smem > cfg_stats) {
    L2D_CUDA(cudaFuncSetAttribute(groupnorm_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cfg_stats = smem;
  }
  launch_pdl_if(pdl_family(4), groupnorm_stats_kernel, dim3(split, n_img), dim3(threads), smem, st, src, gamma, beta, ws, hw, G, split, eps);
  L2D_LAUNCH_CHECK();
  GnApplyParams p{src, ws + (size_t)n_img * GN_MAX_SPLIT * G * 2, y, h, w, silu, mode, stride};
  size_t work = mode == 0 ? (size_t)hw * nch : (size_t)(h / stride) * (w / stride) * 9 * nch;
  int blocks = (int)std::min<size_t>((work + 255) / 256, 148 * 8);
  if (blocks < 1) blocks = 1;
  launch_pdl_if(pdl_family(4), groupnorm_apply_kernel, dim3(blocks, n_img), dim3(256), 0, st, p);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

}  // namespace l2d

using namespace l2d;

extern "C" int l2d_layernorm(const void* x, const void* gamma, const void* beta, void* y, int rows, int channels,
                             float eps, void* stream) {
  L2D_CHECK_ARG(x && gamma && beta && y, "null pointer");
  L2D_CHECK_ARG(channels % 8 == 0 && channels <= 8 * 32 * LN_MAXCH, "need C % 8 == 0 and C <= 2560");
  if (rows <= 0) return L2D_OK;
  const int blocks = ceil_div(rows, 8);
  launch_pdl_if(pdl_family(5), layernorm_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (const __half*)x, (const __half*)gamma,
             (const __half*)beta, (__half*)y, rows, channels, eps);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

extern "C" int64_t l2d_groupnorm_workspace_bytes(int n_img, int groups) {
  return ((int64_t)n_img * GN_MAX_SPLIT * groups * 2 + (int64_t)n_img * 2 * GN_MAX_C) * sizeof(float) +
         (int64_t)2 * n_img * sizeof(int) + 64;   // counters: arrive [N] | depart [N]
}

extern "C" int l2d_groupnorm(const void* x1, int c1, const void* x2, int c2, const void* gamma, const void* beta,
                             void* y, void* workspace, int n_img, int h, int w, int groups, float eps, int silu,
                             int mode, int stride, void* stream) {
  L2D_CHECK_ARG(x1 && gamma && beta && y && workspace, "null pointer");
  L2D_CHECK_ARG(c2 == 0 || x2, "x2 is null but c2 > 0");
  L2D_CHECK_ARG(c1 % 8 == 0 && c2 % 8 == 0, "channel counts must be multiples of 8");
  L2D_CHECK_ARG(groups > 0 && (c1 + c2) % groups == 0, "channels % groups != 0");
  L2D_CHECK_ARG(mode == 0 || mode == 1, "mode must be 0 or 1");
  L2D_CHECK_ARG(stride == 1 || (mode == 1 && stride == 2 && h % 2 == 0 && w % 2 == 0), "bad stride");
  return groupnorm_launch((const __half*)x1, c1, (const __half*)x2, c2, (const __half*)gamma, (const __half*)beta,
                          (__half*)y, (float*)workspace, n_img, h, w, groups, eps, silu, mode, stride,
                          (cudaStream_t)stream);
}

extern "C" int l2d_im2col3x3(const void* x, void* y, int n_img, int h, int w, int channels, int stride, int upsample2x,
                             int silu, void* stream) {
  L2D_CHECK_ARG(x && y, "null pointer");
  L2D_CHECK_ARG(channels % 8 == 0, "C % 8 != 0");
  L2D_CHECK_ARG(stride == 1 || stride == 2, "stride must be 1 or 2");
  L2D_CHECK_ARG(!(upsample2x && stride != 1), "upsample with stride 2 is not a reference op");
  const int hs = upsample2x ? 2 * h : h, ws = upsample2x ? 2 * w : w;
  L2D_CHECK_ARG(hs % stride == 0 && ws % stride == 0, "odd size with stride 2");
  const size_t total = (size_t)n_img * (hs / stride) * (ws / stride) * 9 * (channels / 8);
  int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  if (blocks < 1) blocks = 1;
  im2col3x3_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)x, (__half*)y, n_img, h, w, channels,
                                                             stride, upsample2x, silu);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

extern "C" int l2d_im2col3x3_nchw4(const void* x, void* y, int n_img, int h, int w, void* stream) {
  L2D_CHECK_ARG(x && y, "null pointer");
  const size_t total = (size_t)n_img * h * w * 8;
  int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  if (blocks < 1) blocks = 1;
  im2col3x3_nchw4_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)x, (__half*)y, n_img, h, w);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

extern "C" int l2d_nchw_to_nhwc(const void* x, void* y, int n_img, int channels, int hw, void* stream) {
  L2D_CHECK_ARG(x && y, "null pointer");
  dim3 grid(ceil_div(hw, 32), ceil_div(channels, 32), n_img);
  transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)x, nullptr, (__half*)y, channels, hw);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

extern "C" int l2d_nhwc_to_nchw(const void* x, const void* residual_nchw, void* y, int n_img, int channels, int hw,
                                void* stream) {
  L2D_CHECK_ARG(x && y, "null pointer");
  dim3 grid(ceil_div(channels, 32), ceil_div(hw, 32), n_img);
  transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)x, (const __half*)residual_nchw, (__half*)y,
                                                           hw, channels);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}
