// K8/K9 and small helpers: timestep sinusoid, tiny-M linears (time MLP, time_emb_proj), the LCM
// scheduler step + stream-batch shift, GEGLU weight interleave.
//
// Reference semantics: diffusers 0.25.0 Timesteps/TimestepEmbedding as used at
// unet_depth_streaming.py:499-505; ResnetBlock3D.time_emb_proj(SiLU(temb)) resnet.py:237-238;
// scheduler_step_batch pipeline_stream_animation_depth.py:387-401 and the buffer update :589-601.
#include "common.cuh"

namespace l2d {

// [cos | sin](t * exp(-ln(1e4) * i / half)), computed in fp32 then rounded to fp16 (":504 .to(dtype)")
__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, __half* __restrict__ out, int n, int dim) {
  const int half_dim = dim >> 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half_dim) return;
  const int row = i / half_dim, k = i - row * half_dim;
  const float freq = expf(-9.210340371976184f * (float)k / (float)half_dim);
  const float ang = (float)t[row] * freq;
  out[(size_t)row * dim + k] = __float2half_rn(cosf(ang));
  out[(size_t)row * dim + half_dim + k] = __float2half_rn(sinf(ang));
}

// out[m,n] = act_out(sum_k act_in(x[m,k]) * W[n,k] + b[n]); one warp per output column n, M <= 8.
// Weight-bandwidth bound (each W row read once, 128-bit loads); x rows staged in shared memory.
constexpr int SL_MAXM = 8;
__global__ void __launch_bounds__(256) small_linear_kernel(const __half* __restrict__ x, const __half* __restrict__ w,
                                                           const __half* __restrict__ b, __half* __restrict__ out,
                                                           int M, int N, int K, int silu_in, int silu_out) {
  extern __shared__ __half sx[];  // [M][K]
  for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
    float v = __half2float(x[i]);
    // the reference applies SiLU to the fp16 temb and rounds (resnet.py:238) before the Linear
    sx[i] = silu_in ? __float2half_rn(silu_f(v)) : x[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + warp;
  if (n >= N) return;
  float acc[SL_MAXM];
#pragma unroll
  for (int m = 0; m < SL_MAXM; ++m) acc[m] = 0.f;
  const __half* wr = w + (size_t)n * K;
  for (int k = lane * 8; k < K; k += 256) {
    float wf[8];
    unpack8(ldg_stream(wr + k), wf);
#pragma unroll
    for (int m = 0; m < SL_MAXM; ++m) {
      if (m < M) {
        float xf[8];
        unpack8(*reinterpret_cast<const uint4*>(sx + (size_t)m * K + k), xf);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[m] = fmaf(xf[e], wf[e], acc[m]);
      }
    }
  }
  const float bias = b ? __half2float(b[n]) : 0.f;
#pragma unroll
  for (int m = 0; m < SL_MAXM; ++m) {
    if (m < M) {
      float v = warp_sum(acc[m]) + bias;
      if (silu_out) v = silu_f(v);
      if (lane == 0) out[(size_t)m * N + n] = __float2half_rn(v);
    }
  }
}

// LCM step for the whole stream batch + shift (fp16 storage, fp32 math per element).
__global__ void __launch_bounds__(256) lcm_step_kernel(const __half* __restrict__ x_t, const __half* __restrict__ eps,
                                                       const float* __restrict__ consts, const __half* __restrict__ noise,
                                                       __half* __restrict__ x0_all, __half* __restrict__ out_last,
                                                       __half* __restrict__ next_buf, int n_rows, int per_row) {
  const size_t total = (size_t)n_rows * per_row;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / per_row);
    const size_t e = i - (size_t)row * per_row;
    const float a = consts[row], b = consts[n_rows + row], cs = consts[2 * n_rows + row], co = consts[3 * n_rows + row];
    const float x = __half2float(x_t[i]);
    // mirror the reference's fp16 evaluation order: every intermediate is an fp16 tensor
    const __half t1 = __float2half_rn(__half2float(__float2half_rn(b)) * __half2float(eps[i]));
    const __half t2 = __float2half_rn(x - __half2float(t1));
    const __half f = __float2half_rn(__half2float(t2) / __half2float(__float2half_rn(a)));
    const __half u1 = __float2half_rn(__half2float(__float2half_rn(co)) * __half2float(f));
    const __half u2 = __float2half_rn(__half2float(__float2half_rn(cs)) * x);
    const __half x0 = __float2half_rn(__half2float(u1) + __half2float(u2));
    if (x0_all) x0_all[i] = x0;
    if (row == n_rows - 1) {
      out_last[e] = x0;
    } else if (next_buf) {
      const float an = __half2float(__float2half_rn(consts[row + 1]));
      const float bn = __half2float(__float2half_rn(consts[n_rows + row + 1]));
      const __half v1 = __float2half_rn(an * __half2float(x0));
      const __half v2 = __float2half_rn(bn * (noise ? __half2float(noise[i]) : 0.f));
      next_buf[i] = __float2half_rn(__half2float(v1) + __half2float(v2));
    }
  }
}

// GEGLU interleave: out row (t*tile + r) = in row (t*tile/2 + r) for r < tile/2 (value half),
//                                          in row (F + t*tile/2 + r - tile/2) otherwise (gate half).
__global__ void geglu_interleave_kernel(const __half* __restrict__ w_in, const __half* __restrict__ b_in,
                                        __half* __restrict__ w_out, __half* __restrict__ b_out, int two_f, int K,
                                        int tile) {
  const int row = blockIdx.x;
  const int F = two_f >> 1, half_t = tile >> 1;
  const int t = row / tile, r = row - t * tile;
  const int src = r < half_t ? t * half_t + r : F + t * half_t + (r - half_t);
  for (int k = threadIdx.x; k < K; k += blockDim.x) w_out[(size_t)row * K + k] = w_in[(size_t)src * K + k];
  if (b_in && threadIdx.x == 0) b_out[row] = b_in[src];
}

}  // namespace l2d

using namespace l2d;

extern "C" int l2d_timestep_embedding(const int64_t* t, void* out, int n, int dim, void* stream) {
  L2D_CHECK_ARG(t && out && n > 0 && dim > 0 && dim % 2 == 0, "bad arguments");
  const int total = n * (dim / 2);
  timestep_embedding_kernel<<<ceil_div(total, 128), 128, 0, (cudaStream_t)stream>>>(t, (__half*)out, n, dim);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

extern "C" int l2d_small_linear(const void* x, const void* w, const void* b, void* out, int m, int n, int k,
                                int silu_in, int silu_out, void* stream) {
  L2D_CHECK_ARG(x && w && out, "null pointer");
  L2D_CHECK_ARG(m >= 1 && m <= SL_MAXM, "m must be in 1..8");
  L2D_CHECK_ARG(k % 8 == 0 && (size_t)m * k * 2 <= 48 * 1024, "k % 8 != 0 or m*k too large");
  small_linear_kernel<<<ceil_div(n, 8), 256, (size_t)m * k * sizeof(__half), (cudaStream_t)stream>>>(
      (const __half*)x, (const __half*)w, (const __half*)b, (__half*)out, m, n, k, silu_in, silu_out);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

extern "C" int l2d_lcm_step(const void* x_t, const void* eps, const float* consts, const void* noise, void* x0_all,
                            void* out_last, void* next_buf, int n_rows, int elems_per_row, void* stream) {
  L2D_CHECK_ARG(x_t && eps && consts && out_last, "null pointer");
  L2D_CHECK_ARG(n_rows >= 1 && elems_per_row > 0, "bad sizes");
  const size_t total = (size_t)n_rows * elems_per_row;
  int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 4);
  lcm_step_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)x_t, (const __half*)eps, consts,
                                                            (const __half*)noise, (__half*)x0_all, (__half*)out_last,
                                                            (__half*)next_buf, n_rows, elems_per_row);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

extern "C" int l2d_geglu_interleave(const void* w_in, const void* b_in, void* w_out, void* b_out, int two_f, int k,
                                    int tile_n, void* stream) {
  L2D_CHECK_ARG(w_in && w_out, "null pointer");
  L2D_CHECK_ARG(tile_n % 2 == 0 && two_f % tile_n == 0, "2F must be a multiple of tile_n");
  geglu_interleave_kernel<<<two_f, 128, 0, (cudaStream_t)stream>>>((const __half*)w_in, (const __half*)b_in,
                                                                   (__half*)w_out, (__half*)b_out, two_f, k, tile_n);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}
