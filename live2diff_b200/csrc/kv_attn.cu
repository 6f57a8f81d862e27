// K1 -- temporal KV-cache attention (StreamTemporalAttention core).
//
// Reference semantics: live2diff/animatediff/models/stream_motion_module.py:117-147 (slot append,
// PE-by-index add, rounding of q+pe / K+pe / V+pe to fp16) and :172-194 (per-row additive mask,
// q_len == 1 SDPA over the L-slot window).  Math spec: SURVEY.md Appendix D.
//
// This is an HBM-bandwidth kernel (~1 flop/byte, q_len = 1).  The cache layout [N,2,hw,L,C] makes the K
// (or V) windows of P consecutive pixels ONE contiguous P*L*C*2-byte span, so the kernel is a
// persistent streaming pipeline over the HBM ring buffer:
//   * grid = #SMs; each CTA walks tiles of P pixels (tile -> CTA round-robin)
//   * one elected thread streams each tile's K span and V span into a ring of shared-memory slots with
//     cp.async.bulk (TMA bulk copy, SASS UBLKCP) completing on mbarriers; up to S-1 spans (40 KB each at
//     L=16) are in flight per SM while the current one is consumed
//   * thread <-> one 8-channel (16 B) chunk of one pixel: conflict-free 128-bit shared loads, K+pe / V+pe
//     formed in registers with the reference's fp16 rounding, q.k partials reduced per head through shared
//     memory, fp32 softmax and P.V per thread, 128-bit stores of the output and of the appended k/v
// Nothing is materialised in HBM (no K+pe / V+pe / repeated mask tensors), masked slots never enter the
// math, and slot update_idx[n] is taken from the freshly projected k/v rather than from the ring.
#include "ops.cuh"

namespace l2d {

__device__ __forceinline__ uint32_t kv_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kv_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void kv_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void kv_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if ((++spins & 1023u) == 0) {  // never hang the GPU on a protocol bug
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();
    }
  }
}
// global -> shared bulk copy (non-tensor TMA), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

constexpr int KV_MAX_L = 32;
constexpr int KV_MAX_SLOTS = 4;

// dynamic shared memory: [ring: S * slot_bytes][s_part: P*L*T f32][s_sc: P*heads*L f32][s_mask: 32 f32][s_pi: 32 i32]
//                        [bars: S u64, 8-byte aligned]
// PE_REGS (L <= 16, <= 192 threads): the thread's K_pe / V_pe chunks of all L slots stay in registers for the whole
// row n (every pixel of a row shares pe_idx[n]), so the slot loops are LDS.128 + 4 HADD2 + 8 cvt + 8 FFMA per slot
// with no global loads and no index arithmetic.
template <bool PE_REGS>
__global__ void __launch_bounds__(PE_REGS ? 192 : 320, 1)
kv_attn_kernel(const KvAttnParams p, const int n_slots, const int slot_bytes, const int tiles_per_row) {
  constexpr int LR = 16;   // slots covered by the unrolled register path
  extern __shared__ __align__(128) uint8_t kv_smem[];
  const int L = p.L, T = p.T, P = p.P, C = p.C;
  uint8_t* ring = kv_smem;
  float* s_part = reinterpret_cast<float*>(ring + (size_t)n_slots * slot_bytes);
  float* s_sc = s_part + (size_t)P * L * T;
  float* s_mask = s_sc + (size_t)P * p.heads * L;
  int* s_pi = reinterpret_cast<int*>(s_mask + KV_MAX_L);
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_pi + KV_MAX_L) + 7) & ~uintptr_t(7));
  __shared__ int s_u;

  pdl_launch();
  pdl_wait();
  const int tid = threadIdx.x;
  const int total_tiles = tiles_per_row * p.n_rows;
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles b, b+G, ...
  if (my_tiles <= 0) return;
  const size_t win = (size_t)L * C;            // elements per pixel window
  const size_t kv_plane = (size_t)p.hw * win;  // elements between the K and V planes of one row

  // item q = 2*i + {0: K span, 1: V span} of this CTA's i-th tile; slot = q % S, phase parity = (q / S) & 1
  auto issue = [&](int q) {
    const int i = q >> 1;
    if (i >= my_tiles) return;
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const int n = tile / tiles_per_row, p0 = (tile - n * tiles_per_row) * P;
    const int np = min(P, p.hw - p0);
    const __half* src = p.cache + ((size_t)n * 2 + (q & 1)) * kv_plane + (size_t)p0 * win;
    const uint32_t bytes = (uint32_t)((size_t)np * win * sizeof(__half));
    const int slot = q % n_slots;
    const uint32_t bar = kv_smem_u32(&bars[slot]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic accesses to the slot precede the async refill
    kv_mbar_expect_tx(bar, bytes);
    bulk_g2s(kv_smem_u32(ring + (size_t)slot * slot_bytes), src, bytes, bar);
  };

  if (tid == 0) {
    for (int s = 0; s < n_slots; ++s) kv_mbar_init(kv_smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int q = 0; q < n_slots; ++q) issue(q);
  }

  const int pl = tid / T;
  const int c = tid - pl * T;
  const bool lane_ok = pl < P;
  const size_t row_bytes = (size_t)C * sizeof(__half);
  int cur_n = -1;
  uint32_t vbits = 0;               // bit j set <=> slot j is visible to row n (mask[n,j] > -inf)
  uint4 kpe[LR], vpe[LR];           // PE_REGS only

  for (int i = 0; i < my_tiles; ++i) {
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const int n = tile / tiles_per_row, p0 = (tile - n * tiles_per_row) * P;
    const bool new_row = n != cur_n;   // block-uniform
    if (new_row) {  // per-row schedule: pe_idx[n,:], mask[n,:], update_idx[n]
      __syncthreads();
      if (tid < L) {
        s_pi[tid] = static_cast<int>(p.pe_idx[(size_t)n * L + tid]);
        s_mask[tid] = __half2float(p.mask[(size_t)n * L + tid]);
      }
      if (tid == 0) s_u = static_cast<int>(p.update_idx[n]);
      cur_n = n;
    }
    __syncthreads();   // schedule visible; every thread has finished tile i-1 (its V slot may now be refilled)
    const int u = s_u;
    if (new_row) {
      vbits = 0;
      for (int j = 0; j < L; ++j) vbits |= (s_mask[j] > -INFINITY ? 1u : 0u) << j;
      if (PE_REGS && lane_ok) {
#pragma unroll
        for (int j = 0; j < LR; ++j) {
          if (j < L) {
            kpe[j] = ldg_cached(p.k_pe + (size_t)s_pi[j] * p.pe_ld + (size_t)c * 8);
            vpe[j] = ldg_cached(p.v_pe + (size_t)s_pi[j] * p.pe_ld + (size_t)c * 8);
          }
        }
      }
    }
    const int pixel = p0 + pl;
    const bool active = lane_ok && pixel < p.hw;
    const int qk = 2 * i, qv = 2 * i + 1;
    if (tid == 0 && i > 0) issue(2 * (i - 1) + 1 + n_slots);   // refill the slot V_{i-1} occupied

    uint4 knew = make_uint4(0, 0, 0, 0), vnew = make_uint4(0, 0, 0, 0);
    float qf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    size_t row = 0;
    if (active) {
      row = (size_t)n * p.hw + pixel;
      knew = ldg_act(p.k_new + row * p.ld + (size_t)c * 8);
      vnew = ldg_act(p.v_new + row * p.ld + (size_t)c * 8);
      uint4 qv4 = ldg_act(p.q + row * p.ld + (size_t)c * 8);
      qv4 = hadd8(qv4, ldg_cached(p.q_pe + (size_t)s_pi[u] * p.pe_ld + (size_t)c * 8));   // q + Q_pe[pi[u]] -> fp16
      unpack8(qv4, qf);
      // PE-free append to HBM (stream_motion_module.py:117-119)
      __half* kdst = p.cache + ((size_t)n * 2) * kv_plane + (size_t)pixel * win + (size_t)u * C + (size_t)c * 8;
      *reinterpret_cast<uint4*>(kdst) = knew;
      *reinterpret_cast<uint4*>(kdst + kv_plane) = vnew;
    }

    // ---- K span: q.k partials of this thread's 8-channel chunk -----------------------------------
    kv_mbar_wait(kv_smem_u32(&bars[qk % n_slots]), (uint32_t)(qk / n_slots) & 1u);
    if (active) {
      uint8_t* kb = ring + (size_t)(qk % n_slots) * slot_bytes + ((size_t)pl * win + (size_t)c * 8) * sizeof(__half);
      // slot u of the window is the freshly projected k: patch this thread's own chunk of the landed span
      *reinterpret_cast<uint4*>(kb + (size_t)u * row_bytes) = knew;
      float* part_dst = s_part + (size_t)pl * L * T + c;
      if (PE_REGS) {
#pragma unroll
        for (int j = 0; j < LR; ++j) {
          if (j < L) {
            float part = 0.f;
            if ((vbits >> j) & 1u) {
              const uint4 kk = hadd8(*reinterpret_cast<const uint4*>(kb + (size_t)j * row_bytes), kpe[j]);   // K + K_pe -> fp16
              float kf[8];
              unpack8(kk, kf);
#pragma unroll
              for (int e = 0; e < 8; ++e) part = fmaf(qf[e], kf[e], part);
            }
            part_dst[(size_t)j * T] = part;
          }
        }
      } else {
#pragma unroll 4
        for (int j = 0; j < L; ++j) {
          float part = 0.f;
          if ((vbits >> j) & 1u) {
            const uint4 kk = hadd8(*reinterpret_cast<const uint4*>(kb + (size_t)j * row_bytes),
                                   ldg_cached(p.k_pe + (size_t)s_pi[j] * p.pe_ld + (size_t)c * 8));
            float kf[8];
            unpack8(kk, kf);
#pragma unroll
            for (int e = 0; e < 8; ++e) part = fmaf(qf[e], kf[e], part);
          }
          part_dst[(size_t)j * T] = part;
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my patch of slot u precedes the async refill
    __syncthreads();   // K slot consumed by everyone -> refill it; partials visible
    if (tid == 0) issue(qk + n_slots);

    // ---- per-head scores: sum the hd/8 chunk partials, scale, add mask -----------------------------
    {
      const int total = P * p.heads * L;
      for (int idx = tid; idx < total; idx += blockDim.x) {
        const int j = idx % L;
        const int h = (idx / L) % p.heads;
        const int pl2 = idx / (L * p.heads);
        const float* src = s_part + ((size_t)pl2 * L + j) * T + h * p.hd8;
        float s = 0.f;
        for (int t = 0; t < p.hd8; ++t) s += src[t];
        s_sc[idx] = s * p.scale + s_mask[j];
      }
    }
    __syncthreads();

    // ---- softmax over the window (fp32) + P.V -----------------------------------------------------
    kv_mbar_wait(kv_smem_u32(&bars[qv % n_slots]), (uint32_t)(qv / n_slots) & 1u);
    if (active) {
      const int h = c / p.hd8;
      const float* sc = s_sc + ((size_t)pl * p.heads + h) * L;
      uint8_t* vb = ring + (size_t)(qv % n_slots) * slot_bytes + ((size_t)pl * win + (size_t)c * 8) * sizeof(__half);
      *reinterpret_cast<uint4*>(vb + (size_t)u * row_bytes) = vnew;
      float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      float mx = -INFINITY;
      for (int j = 0; j < L; ++j) mx = fmaxf(mx, sc[j]);
      if (PE_REGS) {
        float ex[LR];
        float denom = 0.f;
#pragma unroll
        for (int j = 0; j < LR; ++j) {
          ex[j] = (j < L && ((vbits >> j) & 1u)) ? __expf(sc[j] - mx) : 0.f;
          denom += ex[j];
        }
        const float inv = 1.f / denom;
#pragma unroll
        for (int j = 0; j < LR; ++j) {
          if (j < L && ((vbits >> j) & 1u)) {
            const float pj = ex[j] * inv;
            const uint4 vv = hadd8(*reinterpret_cast<const uint4*>(vb + (size_t)j * row_bytes), vpe[j]);   // V + V_pe -> fp16
            float vf[8];
            unpack8(vv, vf);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = fmaf(pj, vf[e], o[e]);
          }
        }
      } else {
        float denom = 0.f;
        for (int j = 0; j < L; ++j) denom += __expf(sc[j] - mx);
        const float inv = 1.f / denom;
#pragma unroll 4
        for (int j = 0; j < L; ++j) {
          if ((vbits >> j) & 1u) {
            const float pj = __expf(sc[j] - mx) * inv;
            const uint4 vv = hadd8(*reinterpret_cast<const uint4*>(vb + (size_t)j * row_bytes),
                                   ldg_cached(p.v_pe + (size_t)s_pi[j] * p.pe_ld + (size_t)c * 8));
            float vf[8];
            unpack8(vv, vf);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = fmaf(pj, vf[e], o[e]);
          }
        }
      }
      *reinterpret_cast<uint4*>(p.out + row * C + (size_t)c * 8) = pack8(o);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // (V slot is refilled after the next tile's barrier)
  }
}

static int g_num_sms = 0;

int kv_attn_launch(const KvAttnParams& p0, cudaStream_t stream) {
  // window of 16 slots and 64-aligned channels (every level of the UNet at the default config): tensor-core
  // formulation in kv_attn_mma.cu; other geometries (L = 4 / 32, unusual widths) take the scalar kernel below
  if (kv_attn_mma_supported(p0)) return kv_attn_mma_launch(p0, stream);
  KvAttnParams p = p0;
  p.T = p.C / 8;
  p.hd8 = (p.C / p.heads) / 8;
  if (p.T > 320) return fail(L2D_ERR_INVALID, "kv_attn: channels too large for one block (C <= 2560)");
  int P = 160 / p.T;                 // 160 chunk-threads per CTA: 4/2/1 pixels at C = 320/640/1280
  if (P < 1) P = 1;
  if (P > p.hw) P = p.hw;
  p.P = P;
  const int threads = ((P * p.T + 31) / 32) * 32;
  p.scale = 1.0f / sqrtf((float)(p.C / p.heads));
  const int slot_bytes = (int)((size_t)P * p.L * p.C * sizeof(__half));
  const size_t fixed = ((size_t)P * p.L * p.T + (size_t)P * p.heads * p.L + KV_MAX_L) * sizeof(float) +
                       KV_MAX_L * sizeof(int) + KV_MAX_SLOTS * sizeof(uint64_t) + 128;
  int n_slots = (int)((200 * 1024 - fixed) / (size_t)slot_bytes);
  if (n_slots > KV_MAX_SLOTS) n_slots = KV_MAX_SLOTS;
  if (n_slots < 2) return fail(L2D_ERR_INVALID, "kv_attn: window L*C too large for the shared-memory ring");
  const size_t smem = (size_t)n_slots * slot_bytes + fixed;
  const bool pe_regs = p.L <= 16 && threads <= 192;
  static size_t configured[2] = {0, 0};
  if (smem > configured[pe_regs]) {
    if (pe_regs)
      L2D_CUDA(cudaFuncSetAttribute(kv_attn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
      L2D_CUDA(cudaFuncSetAttribute(kv_attn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[pe_regs] = smem;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    L2D_CUDA(cudaGetDevice(&dev));
    L2D_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int tiles_per_row = ceil_div(p.hw, P);
  const int total = tiles_per_row * p.n_rows;
  const int grid = total < g_num_sms ? total : g_num_sms;
  if (pe_regs)
    launch_pdl_if(p.pdl != 0, kv_attn_kernel<true>, dim3(grid), dim3(threads), smem, stream, p, n_slots, slot_bytes, tiles_per_row);
  else
    launch_pdl_if(p.pdl != 0, kv_attn_kernel<false>, dim3(grid), dim3(threads), smem, stream, p, n_slots, slot_bytes, tiles_per_row);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

}  // namespace l2d

namespace l2d { void kv_attn_set_debug(long long* ptr); }
extern "C" void l2d_kv_attn_set_debug(void* timeline) { l2d::kv_attn_set_debug(static_cast<long long*>(timeline)); }

extern "C" int l2d_kv_attn(const void* q, const void* k_new, const void* v_new, int64_t qkv_ld, void* kv_cache,
                           const void* q_pe, const void* k_pe, const void* v_pe, const void* mask,
                           const int64_t* pe_idx, const int64_t* update_idx, void* out, int n_rows, int hw,
                           int window, int channels, int heads, void* stream) {
  using namespace l2d;
  L2D_CHECK_ARG(q && k_new && v_new && kv_cache && q_pe && k_pe && v_pe && mask && pe_idx && update_idx && out,
                "null pointer");
  L2D_CHECK_ARG(n_rows > 0 && hw > 0 && window > 0 && window <= KV_MAX_L, "need 0 < L <= 32");
  L2D_CHECK_ARG(heads > 0 && channels % heads == 0, "channels % heads != 0");
  L2D_CHECK_ARG(channels % 8 == 0 && (channels / heads) % 8 == 0, "need C % 8 == 0 and head_dim % 8 == 0");
  L2D_CHECK_ARG(qkv_ld % 8 == 0 && qkv_ld >= channels, "qkv_ld must be >= C and a multiple of 8");
  L2D_CHECK_ARG((uintptr_t)kv_cache % 16 == 0, "kv_cache must be 16-byte aligned");
  KvAttnParams p{};
  p.q = (const __half*)q; p.k_new = (const __half*)k_new; p.v_new = (const __half*)v_new; p.ld = qkv_ld;
  p.cache = (__half*)kv_cache; p.q_pe = (const __half*)q_pe; p.k_pe = (const __half*)k_pe;
  p.v_pe = (const __half*)v_pe; p.mask = (const __half*)mask; p.pe_idx = pe_idx; p.update_idx = update_idx;
  p.out = (__half*)out; p.pe_ld = channels; p.n_rows = n_rows; p.hw = hw; p.L = window; p.C = channels; p.heads = heads;
  return kv_attn_launch(p, (cudaStream_t)stream);
}
