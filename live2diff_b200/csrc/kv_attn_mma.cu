// K1, tensor-core formulation (window L == 16, C % 64 == 0, heads <= 8): same semantics as kv_attn.cu
// (stream_motion_module.py:117-194), ~6x fewer issue slots per byte.
//
// Why: the scalar kernel needs ~2000 warp-instructions per 80 KB tile (fp16->fp32 converts + FMAs per element), which
// with the 5-8 warps an SM can hold next to a 160 KB ring caps it near 25 % of HBM bandwidth (profiles/).  Here one
// mma.sync.m16n8k16 evaluates 16 slots x 16 channels of q.K (or of P.V) at once:
//   scores[16 slots] = K~[16 x hd] . q~[hd]      A = K~ tile (ldmatrix from the ring), B = q~ in column 0
//   out[hd]          = V~^T[hd x 16] . p[16]     A = V~^T tile (ldmatrix.trans),       B = p  in column 0
// (7 of the 8 B columns are zero -- q_len is 1 -- the tensor pipe is not the limiter, issue slots are.)
// K+pe / V+pe are formed on the A fragments with HADD2 against per-row PE fragments kept in registers, i.e. with the
// reference's fp16 rounding.  P is rounded to fp16 for the P.V product like the fused SDPA kernels the reference uses.
//
// Data movement: the cache [N,2,hw,L,C] viewed as a matrix [N*2*hw*L rows, C cols]; a tile of P pixels is
// R = P*16 consecutive rows of the K plane and of the V plane, fetched as 64-column TMA boxes with the 128B swizzle
// (cp.async.bulk.tensor.2d, conflict-free ldmatrix) into a 2-stage ring, completion on mbarriers.  One warp per head.
// Precondition (as in the reference, whose cache is zero-initialised): masked slots hold finite values.
#include <cuda.h>

#include "ops.cuh"

namespace l2d {

int get_tmap_2d(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out);   // gemm_tcgen05.cu

namespace {

__device__ __forceinline__ uint32_t km_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void km_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void km_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void km_mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void km_tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void km_ldsm(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void km_ldsm_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void km_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t km_hadd2(uint32_t a, uint32_t b) { return h2_as_u32(__hadd2(u32_as_h2(a), u32_as_h2(b))); }

constexpr int KM_L = 16;
constexpr int KM_STAGES = 2;

// swizzled byte offset of 16-byte chunk `chunk` (= channel / 8) of tile row `row` inside one plane of a stage:
// 64-column blocks of R rows x 128 B, 128B swizzle = chunk-in-block XOR (row % 8)
__device__ __forceinline__ uint32_t km_off(int row, int chunk, int R) {
  return (uint32_t)((chunk >> 3) * (R * 128) + row * 128 + (((chunk & 7) ^ (row & 7)) << 4));
}

// KS = ceil(head_dim / 16): k-steps of q.K per head = d-tiles of P.V per head
template <int KS>
__global__ void __launch_bounds__(256, 1)
kv_attn_mma_kernel(const __grid_constant__ CUtensorMap tmap, const KvAttnParams p, const int tiles_per_row, const int ncb) {
  extern __shared__ __align__(1024) uint8_t km_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(km_smem_raw) + 1023) & ~uintptr_t(1023));
  const int C = p.C, P = p.P, R = P * KM_L, T = C >> 3;
  const int hd = C / p.heads;
  const uint32_t plane_bytes = (uint32_t)ncb * R * 128;          // K (or V) rows of one tile
  const uint32_t stage_bytes = 2 * plane_bytes;
  uint8_t* ring = smem;
  __half* s_q = reinterpret_cast<__half*>(ring + (size_t)KM_STAGES * stage_bytes);   // [P][C]  q + Q_pe
  __half* s_out = s_q + (size_t)P * C;                                                // [P][C]
  float* s_mask = reinterpret_cast<float*>(s_out + (size_t)P * C);                    // [16]
  int* s_pi = reinterpret_cast<int*>(s_mask + KM_L);                                  // [16]
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_pi + KM_L) + 7) & ~uintptr_t(7));
  __shared__ int s_u;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int total_tiles = tiles_per_row * p.n_rows;
  const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  if (my_tiles == 0) return;

  auto issue = [&](int i) {   // tile i of this CTA -> stage i % 2: ncb K boxes + ncb V boxes
    if (i >= my_tiles) return;
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const int n = tile / tiles_per_row, p0 = (tile - n * tiles_per_row) * P;
    const int st = i % KM_STAGES;
    const uint32_t bar = km_smem_u32(&bars[st]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic accesses to the stage precede the refill
    km_mbar_expect_tx(bar, stage_bytes);
    const int krow = ((n * 2) * p.hw + p0) * KM_L, vrow = ((n * 2 + 1) * p.hw + p0) * KM_L;
    const uint32_t base = km_smem_u32(ring + (size_t)st * stage_bytes);
    for (int b = 0; b < ncb; ++b) {
      km_tma_2d(base + b * (R * 128), &tmap, bar, b * 64, krow);
      km_tma_2d(base + plane_bytes + b * (R * 128), &tmap, bar, b * 64, vrow);
    }
  };
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    for (int s = 0; s < KM_STAGES; ++s) km_mbar_init(km_smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < KM_STAGES; ++i) issue(i);
  }

  const bool head_warp = warp < p.heads;
  const int h = warp;
  const int ch0 = h * hd;                      // first channel of this warp's head
  uint32_t kpe[KS][4], vpe[KS][4];             // PE fragments of the current row n
  int cur_n = -1;

  for (int i = 0; i < my_tiles; ++i) {
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const int n = tile / tiles_per_row, p0 = (tile - n * tiles_per_row) * P;
    const int np = min(P, p.hw - p0);
    const bool new_row = n != cur_n;   // block-uniform
    if (new_row) {
      __syncthreads();
      if (tid < KM_L) {
        s_pi[tid] = static_cast<int>(p.pe_idx[(size_t)n * KM_L + tid]);
        s_mask[tid] = __half2float(p.mask[(size_t)n * KM_L + tid]);
      }
      if (tid == 0) s_u = static_cast<int>(p.update_idx[n]);
      cur_n = n;
      __syncthreads();
      if (head_warp) {
        // A-fragment register r of a 16x16 tile holds (row g + 8*(r&1), cols 2t,2t+1 of 8-column group r>>1)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            // K~ = K + K_pe[pi[slot]]: rows = slots, cols = channels
            const int slot = g + (r & 1) * 8;
            const int cc = ks * 16 + (r >> 1) * 8 + 2 * t;          // channel within the head
            kpe[ks][r] = cc < hd ? *reinterpret_cast<const uint32_t*>(p.k_pe + (size_t)s_pi[slot] * p.pe_ld + ch0 + cc) : 0u;
            // V~^T: rows = channels (d), cols = slots; register = (V_pe[pi[2t+..]][d], V_pe[pi[2t+1+..]][d])
            const int d = ks * 16 + g + (r & 1) * 8;
            const int s0 = 2 * t + (r >> 1) * 8;
            uint32_t v = 0u;
            if (d < hd) {
              const __half lo = p.v_pe[(size_t)s_pi[s0] * p.pe_ld + ch0 + d];
              const __half hi = p.v_pe[(size_t)s_pi[s0 + 1] * p.pe_ld + ch0 + d];
              v = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
            }
            vpe[ks][r] = v;
          }
        }
      }
    }
    const int u = s_u;
    const int st = i % KM_STAGES;
    uint8_t* kplane = ring + (size_t)st * stage_bytes;
    uint8_t* vplane = kplane + plane_bytes;
    km_mbar_wait(km_smem_u32(&bars[st]), (uint32_t)(i / KM_STAGES) & 1u);

    // ---- append (HBM + ring patch) and q~ staging: thread <-> one 16-byte chunk of one pixel ----
    for (int idx = tid; idx < np * T; idx += blockDim.x) {
      const int pl = idx / T, c = idx - pl * T;
      const size_t row = (size_t)n * p.hw + p0 + pl;
      const uint4 knew = ldg_cached(p.k_new + row * p.ld + (size_t)c * 8);
      const uint4 vnew = ldg_cached(p.v_new + row * p.ld + (size_t)c * 8);
      uint4 qv = ldg_cached(p.q + row * p.ld + (size_t)c * 8);
      qv = hadd8(qv, ldg_cached(p.q_pe + (size_t)s_pi[u] * p.pe_ld + (size_t)c * 8));     // q + Q_pe[pi[u]] -> fp16
      __half* kdst = p.cache + ((((size_t)n * 2) * p.hw + p0 + pl) * KM_L + u) * C + (size_t)c * 8;
      *reinterpret_cast<uint4*>(kdst) = knew;                                               // PE-free append (:117-119)
      *reinterpret_cast<uint4*>(kdst + (size_t)p.hw * KM_L * C) = vnew;
      const uint32_t off = km_off(pl * KM_L + u, c, R);
      *reinterpret_cast<uint4*>(kplane + off) = knew;                                       // the window sees the new slot
      *reinterpret_cast<uint4*>(vplane + off) = vnew;
      *reinterpret_cast<uint4*>(s_q + (size_t)pl * C + (size_t)c * 8) = qv;
    }
    __syncthreads();

    // ---- one warp per head: scores, softmax, P.V for every pixel of the tile ----
    if (head_warp) {
      const uint32_t kbase = km_smem_u32(kplane), vbase = km_smem_u32(vplane);
      const int mi = lane >> 3, r8 = lane & 7;
      const float m_lo = s_mask[g], m_hi = s_mask[g + 8];
      for (int pl = 0; pl < np; ++pl) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const __half* qrow = s_q + (size_t)pl * C + ch0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t a[4];
          // matrices: (slots 0-7, cols 0-7) (slots 8-15, cols 0-7) (slots 0-7, cols 8-15) (slots 8-15, cols 8-15)
          const int row = pl * KM_L + (mi & 1) * 8 + r8;
          int chunk = ((ch0 + ks * 16) >> 3) + (mi >> 1);
          if (chunk >= T) chunk = T - 1;                      // beyond the last head: any valid address (zeroed below)
          km_ldsm(kbase + km_off(row, chunk, R), a);
#pragma unroll
          for (int r = 0; r < 4; ++r) a[r] = km_hadd2(a[r], kpe[ks][r]);       // K + K_pe -> fp16
          if (ks * 16 + 8 >= hd) a[2] = a[3] = 0u;                             // columns past the head never enter
          uint32_t b0 = 0u, b1 = 0u;                                           // q~ lives in column 0 (lanes g == 0)
          if (g == 0) {
            b0 = *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 2 * t);
            if (ks * 16 + 8 < hd) b1 = *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 8 + 2 * t);
          }
          km_mma(acc, a, b0, b1);
        }
        // column 0 of the score tile: lanes with t == 0 hold slots g (acc[0]) and g+8 (acc[2])
        float s_lo = t == 0 ? fmaf(acc[0], p.scale, m_lo) : -INFINITY;
        float s_hi = t == 0 ? fmaf(acc[2], p.scale, m_hi) : -INFINITY;
        float mx = fmaxf(s_lo, s_hi);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float e_lo = t == 0 ? __expf(s_lo - mx) : 0.f;
        const float e_hi = t == 0 ? __expf(s_hi - mx) : 0.f;
        float sum = e_lo + e_hi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.f / sum;
        const float p_lo = e_lo * inv, p_hi = e_hi * inv;
        // B fragment of P.V: column 0 <-> lanes g == 0: (p[2t], p[2t+1]) and (p[2t+8], p[2t+9]);
        // p[j] for j < 8 sits in lane 4j (p_lo), p[j+8] in lane 4j (p_hi)
        const int src = 8 * t;
        const float x0 = __shfl_sync(0xffffffffu, p_lo, src), x1 = __shfl_sync(0xffffffffu, p_lo, src + 4);
        const float y0 = __shfl_sync(0xffffffffu, p_hi, src), y1 = __shfl_sync(0xffffffffu, p_hi, src + 4);
        uint32_t pb0 = 0u, pb1 = 0u;
        if (g == 0) {
          pb0 = h2_as_u32(__floats2half2_rn(x0, x1));
          pb1 = h2_as_u32(__floats2half2_rn(y0, y1));
        }
        __half* orow = s_out + (size_t)pl * C + ch0;
#pragma unroll
        for (int dt = 0; dt < KS; ++dt) {
          uint32_t a[4];
          // V~^T tile: matrices (slots 0-7, d 0-7) (slots 0-7, d 8-15) (slots 8-15, d 0-7) (slots 8-15, d 8-15), transposed
          const int row = pl * KM_L + (mi >> 1) * 8 + r8;
          int chunk = ((ch0 + dt * 16) >> 3) + (mi & 1);
          if (chunk >= T) chunk = T - 1;
          km_ldsm_t(vbase + km_off(row, chunk, R), a);
#pragma unroll
          for (int r = 0; r < 4; ++r) a[r] = km_hadd2(a[r], vpe[dt][r]);       // V + V_pe -> fp16
          float o[4] = {0.f, 0.f, 0.f, 0.f};
          km_mma(o, a, pb0, pb1);
          if (t == 0) {                                                         // column 0: d = dt*16 + g, + 8
            const int d0 = dt * 16 + g;
            if (d0 < hd) orow[d0] = __float2half_rn(o[0]);
            if (d0 + 8 < hd) orow[d0 + 8] = __float2half_rn(o[2]);
          }
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // ring patches precede the async refill
    __syncthreads();
    if (tid == 0) issue(i + KM_STAGES);                            // stage consumed: fetch the tile after next
    // ---- coalesced 128-bit stores of the tile's output rows ----
    for (int idx = tid; idx < np * T; idx += blockDim.x) {
      const int pl = idx / T, c = idx - pl * T;
      const size_t row = (size_t)n * p.hw + p0 + pl;
      *reinterpret_cast<uint4*>(p.out + row * C + (size_t)c * 8) = *reinterpret_cast<const uint4*>(s_out + (size_t)pl * C + (size_t)c * 8);
    }
  }
}

int g_km_sms = 0;

template <int KS>
int km_launch(const CUtensorMap& tm, const KvAttnParams& p, int tiles_per_row, int ncb, int grid, size_t smem, cudaStream_t st) {
  static size_t configured = 0;
  if (smem > configured) {
    L2D_CUDA(cudaFuncSetAttribute(kv_attn_mma_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  kv_attn_mma_kernel<KS><<<grid, 256, smem, st>>>(tm, p, tiles_per_row, ncb);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

}  // namespace

bool kv_attn_mma_supported(const KvAttnParams& p) {
  const int hd = p.heads > 0 ? p.C / p.heads : 0;
  return p.L == KM_L && p.C % 64 == 0 && p.heads >= 1 && p.heads <= 8 && hd % 8 == 0 && hd <= 160 && p.pe_ld % 2 == 0;
}

int kv_attn_mma_launch(const KvAttnParams& p0, cudaStream_t stream) {
  KvAttnParams p = p0;
  const int hd = p.C / p.heads;
  p.T = p.C / 8;
  p.hd8 = hd / 8;
  int P = 1280 / p.C;                        // 80 KB of K+V per tile: 4 / 2 / 1 pixels at C = 320 / 640 / 1280
  if (P < 1) P = 1;
  if (P * KM_L > 256) P = 256 / KM_L;        // TMA box rows <= 256
  if (P > p.hw) P = p.hw;
  p.P = P;
  p.scale = 1.0f / sqrtf((float)hd);
  const int ncb = p.C / 64;
  const size_t stage_bytes = (size_t)2 * ncb * P * KM_L * 128;
  const size_t smem = KM_STAGES * stage_bytes + (size_t)2 * P * p.C * sizeof(__half) + KM_L * 8 + KM_STAGES * 8 + 64 + 1024;
  if (smem > 227 * 1024) return fail(L2D_ERR_INVALID, "kv_attn(mma): tile does not fit in shared memory");
  CUtensorMap tm;
  const int64_t rows = (int64_t)p.n_rows * 2 * p.hw * KM_L;
  int rc = get_tmap_2d(p.cache, rows, p.C, p.C, P * KM_L, &tm);
  if (rc != L2D_OK) return rc;
  if (g_km_sms == 0) {
    int dev = 0;
    L2D_CUDA(cudaGetDevice(&dev));
    L2D_CUDA(cudaDeviceGetAttribute(&g_km_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int tiles_per_row = ceil_div(p.hw, P);
  const int total = tiles_per_row * p.n_rows;
  const int grid = total < g_km_sms ? total : g_km_sms;
  switch ((hd + 15) / 16) {
    case 1: return km_launch<1>(tm, p, tiles_per_row, ncb, grid, smem, stream);
    case 2: return km_launch<2>(tm, p, tiles_per_row, ncb, grid, smem, stream);
    case 3: return km_launch<3>(tm, p, tiles_per_row, ncb, grid, smem, stream);
    case 4: return km_launch<4>(tm, p, tiles_per_row, ncb, grid, smem, stream);
    case 5: return km_launch<5>(tm, p, tiles_per_row, ncb, grid, smem, stream);
    case 10: return km_launch<10>(tm, p, tiles_per_row, ncb, grid, smem, stream);
    default: return fail(L2D_ERR_INVALID, "kv_attn(mma): unsupported head_dim " + std::to_string(hd));
  }
}

}  // namespace l2d
