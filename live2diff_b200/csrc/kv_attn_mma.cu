// K1, tensor-core formulation (window L == 16, C % 64 == 0, C <= 640, heads <= 8): same semantics as kv_attn.cu
// (stream_motion_module.py:117-194) at a fraction of the issue slots per byte.
//
// Why: the scalar kernel needs ~2000 warp-instructions per 80 KB tile (fp16->fp32 converts + FMAs per element); with
// the few warps an SM can hold next to a 160 KB ring that caps it near 25 % of HBM bandwidth (profiles/).  Here a
// warp owns a PIXEL and one mma.sync.m16n8k16 evaluates 16 slots x 16 channels for all 8 heads at once (heads on the
// N dimension through a block-diagonal q matrix; see the kernel comment).  K+pe / V+pe are formed on the A fragments
// with HADD2, i.e. with the reference's fp16 rounding; P is rounded to fp16 for P.V like the fused SDPA kernels the
// reference dispatches to.
//
// Data movement: the cache [N,2,hw,L,C] viewed as a matrix [N*2*hw*L rows, C cols]; a tile of P pixels is
// R = P*16 consecutive rows of the K plane and of the V plane, fetched as 64-column TMA boxes with the 128B swizzle
// (cp.async.bulk.tensor.2d -> conflict-free ldmatrix) into a 2-stage ring, completion on mbarriers; the freshly
// projected k/v/q of the next tile are prefetched into registers during the current tile's math.
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
// (not volatile / no memory clobber: within the compute phase the ring is read-only, and leaving the scheduler free
// to interleave the independent per-pixel chains is what hides the ldmatrix -> HADD2 -> mma -> shuffle latencies)
__device__ __forceinline__ void km_ldsm(uint32_t addr, uint32_t (&r)[4]) {
  asm("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void km_ldsm_t(uint32_t addr, uint32_t (&r)[4]) {
  asm("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void km_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void km_st_pred(uint32_t addr, __half v, bool pred) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %2, 0; @p st.shared.b16 [%0], %1; }" ::"r"(addr),
               "h"(*reinterpret_cast<const unsigned short*>(&v)), "r"((int)pred)
               : "memory");
}
__device__ __forceinline__ uint32_t km_hadd2(uint32_t a, uint32_t b) { return h2_as_u32(__hadd2(u32_as_h2(a), u32_as_h2(b))); }

constexpr int KM_L = 16;
constexpr int KM_STAGES = 2;

// swizzled byte offset of 16-byte chunk `chunk` (= channel / 8) of tile row `row` inside one plane of a stage:
// 64-column blocks of R rows x 128 B, 128B swizzle = chunk-in-block XOR (row % 8)
__device__ __forceinline__ uint32_t km_off(int row, int chunk, int R) {
  return (uint32_t)((chunk >> 3) * (R * 128) + row * 128 + (((chunk & 7) ^ (row & 7)) << 4));
}

// One warp per PIXEL, the 8 heads on the N = 8 dimension of the MMA:
//   S[16 slots x 8 heads] = K~[16 x C] . Qblk[C x 8]     Qblk[c][h] = q~[c] if channel c belongs to head h else 0
//   O^T[C x 8 heads]      = V~^T[C x 16] . P[16 x 8]      row c of O^T is read at column head(c)
// so every B column does useful work, the softmax of all heads runs in one set of registers (scores of slot g /
// g+8 and heads 2t / 2t+1 per lane; the reduction over slots is 3 xor-shuffles for all 8 heads at once), and one pixel
// costs C/16 k-steps of (2 ldmatrix + 4 HADD2 + 1 mma) for the scores and again for P.V.
// PE: the row's K_pe[pi[j]] / V_pe[pi[j]] windows are staged once per row n in shared memory in the same swizzled
// [16 x C] layout as a pixel's K / V window, so the PE fragments come from the same ldmatrix addresses.
__global__ void __launch_bounds__(288, 1)
kv_attn_mma_kernel(const __grid_constant__ CUtensorMap tmap, const KvAttnParams p, const int tiles_per_row, const int ncb,
                   long long* dbg) {
  extern __shared__ __align__(1024) uint8_t km_smem_raw[];
  // align by pointer arithmetic (an integer round trip would turn every later access into a generic LD/ST)
  uint8_t* smem = km_smem_raw + ((1024u - (km_smem_u32(km_smem_raw) & 1023u)) & 1023u);
  const int C = p.C, P = p.P, R = P * KM_L, T = C >> 3;
  const int hd = C / p.heads;
  const uint32_t plane_bytes = (uint32_t)ncb * R * 128;          // K (or V) rows of one tile
  const uint32_t stage_bytes = 2 * plane_bytes;
  const uint32_t pe_plane = (uint32_t)ncb * KM_L * 128;          // one [16 x C] PE window
  uint8_t* ring = smem;
  uint8_t* pek = ring + (size_t)KM_STAGES * stage_bytes;         // K_pe[pi[j]] window of the current row n
  uint8_t* pev = pek + pe_plane;
  __half* s_q = reinterpret_cast<__half*>(pev + pe_plane);                            // [P][C]  q + Q_pe
  __half* s_out = s_q + (size_t)P * C;                                                // [P][C]
  float* s_mask = reinterpret_cast<float*>(s_out + (size_t)P * C);                    // [16]
  int* s_pi = reinterpret_cast<int*>(s_mask + KM_L);                                  // [16]
  uint8_t* s_head = reinterpret_cast<uint8_t*>(s_pi + KM_L);                          // [T] head of each 8-channel chunk
  float* s_part = reinterpret_cast<float*>(s_head + 192);                             // [8 warps][32 lanes][4] partial scores
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + 8 * 32 * 4);   // 8-byte aligned: every size above is a multiple of 8
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
    for (int b2 = 0; b2 < ncb; ++b2) {
      km_tma_2d(base + b2 * (R * 128), &tmap, bar, b2 * 64, krow);
      km_tma_2d(base + plane_bytes + b2 * (R * 128), &tmap, bar, b2 * 64, vrow);
    }
  };
  constexpr int PRODUCER_TID = 256;             // warp 8 never does math, so TMA issue does not delay it
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    for (int s2 = 0; s2 < KM_STAGES; ++s2) km_mbar_init(km_smem_u32(&bars[s2]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < KM_STAGES; ++i) issue(i);
  }
  for (int c = tid; c < T; c += blockDim.x) s_head[c] = (uint8_t)((c * 8) / hd);

  // k/v/q chunk of the NEXT tile, requested one tile ahead (thread <-> chunk, as in the append phase)
  uint4 pf_k = make_uint4(0, 0, 0, 0), pf_v = pf_k, pf_q = pf_k;
  auto load_qkv = [&](int n, int p0, int np) {
    if (tid < np * T) {
      const int pl = tid / T, c = tid - pl * T;
      const size_t row = (size_t)n * p.hw + p0 + pl;
      pf_k = ldg_cached(p.k_new + row * p.ld + (size_t)c * 8);
      pf_v = ldg_cached(p.v_new + row * p.ld + (size_t)c * 8);
      pf_q = ldg_cached(p.q + row * p.ld + (size_t)c * 8);
    }
  };
  int cur_n = -1;
  const int nks = C >> 4;                       // 16-channel k-steps (= d-tiles) per pixel
  const int mi = lane >> 3, r8 = lane & 7;

  long long d_wait = 0, d_patch = 0, d_comp = 0, d_store = 0, d_t0 = 0, d_t1 = 0, d_qk = 0, d_sm = 0, d_pv = 0;   // developer timeline (thread 0)
  for (int i = 0; i < my_tiles; ++i) {
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const int n = tile / tiles_per_row, p0 = (tile - n * tiles_per_row) * P;
    const int np = min(P, p.hw - p0);
    if (n != cur_n) {   // block-uniform: per-row schedule + PE windows
      __syncthreads();
      if (tid < KM_L) {
        s_pi[tid] = static_cast<int>(p.pe_idx[(size_t)n * KM_L + tid]);
        s_mask[tid] = __half2float(p.mask[(size_t)n * KM_L + tid]);
      }
      if (tid == 0) s_u = static_cast<int>(p.update_idx[n]);
      cur_n = n;
      __syncthreads();
      for (int idx = tid; idx < KM_L * T; idx += blockDim.x) {
        const int j = idx / T, c = idx - j * T;
        const uint32_t off = km_off(j, c, KM_L);
        *reinterpret_cast<uint4*>(pek + off) = ldg_cached(p.k_pe + (size_t)s_pi[j] * p.pe_ld + (size_t)c * 8);
        *reinterpret_cast<uint4*>(pev + off) = ldg_cached(p.v_pe + (size_t)s_pi[j] * p.pe_ld + (size_t)c * 8);
      }
    }
    const int u = s_u;
    const int st = i % KM_STAGES;
    uint8_t* kplane = ring + (size_t)st * stage_bytes;
    uint8_t* vplane = kplane + plane_bytes;
    if (dbg && tid == 0) d_t0 = clock64();
    km_mbar_wait(km_smem_u32(&bars[st]), (uint32_t)(i / KM_STAGES) & 1u);
    if (dbg && tid == 0) { const long long c = clock64(); d_wait += c - d_t0; d_t0 = c; }

    // ---- append (HBM + ring patch) and q~ staging: thread <-> one 16-byte chunk of one pixel.  The chunk's k/v/q were
    //      requested during the previous tile's compute phase (registers pf_*), so no global latency is exposed here ----
    if (i == 0) load_qkv(n, p0, np);
    if (tid < np * T) {
      const int pl = tid / T, c = tid - pl * T;
      const uint4 qv = hadd8(pf_q, ldg_cached(p.q_pe + (size_t)s_pi[u] * p.pe_ld + (size_t)c * 8));   // q + Q_pe[pi[u]] -> fp16
      __half* kdst = p.cache + ((((size_t)n * 2) * p.hw + p0 + pl) * KM_L + u) * C + (size_t)c * 8;
      *reinterpret_cast<uint4*>(kdst) = pf_k;                                              // PE-free append (:117-119)
      *reinterpret_cast<uint4*>(kdst + (size_t)p.hw * KM_L * C) = pf_v;
      const uint32_t off = km_off(pl * KM_L + u, c, R);
      *reinterpret_cast<uint4*>(kplane + off) = pf_k;                                      // the window sees the new slot
      *reinterpret_cast<uint4*>(vplane + off) = pf_v;
      *reinterpret_cast<uint4*>(s_q + (size_t)pl * C + (size_t)c * 8) = qv;
    }
    __syncthreads();
    if (dbg && tid == 0) { const long long c = clock64(); d_patch += c - d_t0; d_t0 = c; }
    if (i + 1 < my_tiles) {   // request the next tile's k/v/q now; they land while this tile is computed
      const int tile2 = tile + (int)gridDim.x;
      const int n2 = tile2 / tiles_per_row, q0 = (tile2 - n2 * tiles_per_row) * P;
      load_qkv(n2, q0, min(P, p.hw - q0));
    }

    // ---- W = 8 / P warps per pixel: the k-steps of the score MMA chain and the d-tiles of P.V are dealt round-robin to
    //      the pixel's warps (shorter dependent chains, more warps in flight); partial scores meet in shared memory ----
    const int W = 8 / P;                         // P is a power of two <= 8
    if (warp < 8 && warp / W < np) {
      const int pl = warp / W, sub = warp - pl * W;
      const uint32_t kbase = km_smem_u32(kplane), vbase = km_smem_u32(vplane);
      const uint32_t pkb = km_smem_u32(pek), pvb = km_smem_u32(pev);
      const __half* qrow = s_q + (size_t)pl * C;
      // scores: rows = slots, columns = heads; two accumulators halve the dependent mma chain
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
      const int rowk = (mi & 1) * 8 + r8;              // slot addressed by this lane for the K~ tile
      auto qk_step = [&](int ks, float (&dst)[4]) {
        uint32_t a[4], pe[4];
        // matrices: (slots 0-7, ch 0-7) (slots 8-15, ch 0-7) (slots 0-7, ch 8-15) (slots 8-15, ch 8-15)
        const int chunk = 2 * ks + (mi >> 1);
        km_ldsm(kbase + km_off(pl * KM_L + rowk, chunk, R), a);
        km_ldsm(pkb + km_off(rowk, chunk, KM_L), pe);
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = km_hadd2(a[r], pe[r]);             // K + K_pe -> fp16
        // B = Qblk: lane (g, t) holds rows (channels) ks*16 + 2t, +1 (b0) and + 8 (b1) of column (head) g
        const uint32_t q0 = *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 2 * t);
        const uint32_t q1 = *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 8 + 2 * t);
        const uint32_t b0 = (int)s_head[2 * ks] == g ? q0 : 0u;
        const uint32_t b1 = (int)s_head[2 * ks + 1] == g ? q1 : 0u;
        km_mma(dst, a, b0, b1);
      };
      int ks = sub;
#pragma unroll 2
      for (; ks + W < nks; ks += 2 * W) {
        qk_step(ks, acc);
        qk_step(ks + W, acc2);
      }
      if (ks < nks) qk_step(ks, acc);
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] += acc2[r];
      if (dbg && tid == 0) { const long long c = clock64(); d_qk += c - d_t0; d_t1 = c; }
      if (W > 1) {   // sum the partial scores of this pixel's warps
        float* mine = s_part + ((size_t)warp * 32 + lane) * 4;
        *reinterpret_cast<float4*>(mine) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + pl), "r"(W * 32) : "memory");
        acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
        for (int w2 = 0; w2 < W; ++w2) {
          const float4 v4 = *reinterpret_cast<const float4*>(s_part + ((size_t)(pl * W + w2) * 32 + lane) * 4);
          acc[0] += v4.x; acc[1] += v4.y; acc[2] += v4.z; acc[3] += v4.w;
        }
      }
      // lane (g, t): acc[0] = S[slot g][head 2t], acc[1] = S[g][2t+1], acc[2] = S[g+8][2t], acc[3] = S[g+8][2t+1]
      const float m_lo = s_mask[g], m_hi = s_mask[g + 8];
      float s0 = fmaf(acc[0], p.scale, m_lo), s1 = fmaf(acc[1], p.scale, m_lo);
      float s2 = fmaf(acc[2], p.scale, m_hi), s3 = fmaf(acc[3], p.scale, m_hi);
      float mxa = fmaxf(s0, s2), mxb = fmaxf(s1, s3);                        // heads 2t and 2t+1
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {                                    // reduce over g (the slots)
        mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, o));
        mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, o));
      }
      s0 = __expf(s0 - mxa); s2 = __expf(s2 - mxa);
      s1 = __expf(s1 - mxb); s3 = __expf(s3 - mxb);
      float suma = s0 + s2, sumb = s1 + s3;
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {
        suma += __shfl_xor_sync(0xffffffffu, suma, o);
        sumb += __shfl_xor_sync(0xffffffffu, sumb, o);
      }
      const float inva = 1.f / suma, invb = 1.f / sumb;
      const float pr0 = s0 * inva, pr1 = s1 * invb, pr2 = s2 * inva, pr3 = s3 * invb;   // P[slot g | g+8][head 2t | 2t+1]
      // B = P[16 slots x 8 heads]: lane (g, t) needs head g, slots 2t, 2t+1 (b0) and 2t+8, 2t+9 (b1).
      // P[slot s][head h] lives in lane (s % 8) * 4 + h / 2, register (s / 8) * 2 + h % 2.
      const int srcA = 8 * t + (g >> 1), srcB = srcA + 4;
      const float a0 = __shfl_sync(0xffffffffu, pr0, srcA), a1 = __shfl_sync(0xffffffffu, pr1, srcA);
      const float b0f = __shfl_sync(0xffffffffu, pr0, srcB), b1f = __shfl_sync(0xffffffffu, pr1, srcB);
      const float c0 = __shfl_sync(0xffffffffu, pr2, srcA), c1 = __shfl_sync(0xffffffffu, pr3, srcA);
      const float d0f = __shfl_sync(0xffffffffu, pr2, srcB), d1f = __shfl_sync(0xffffffffu, pr3, srcB);
      const bool odd = g & 1;
      const uint32_t pb0 = h2_as_u32(__floats2half2_rn(odd ? a1 : a0, odd ? b1f : b0f));   // slots 2t, 2t+1
      const uint32_t pb1 = h2_as_u32(__floats2half2_rn(odd ? c1 : c0, odd ? d1f : d0f));   // slots 2t+8, 2t+9
      if (dbg && tid == 0) { const long long c = clock64(); d_sm += c - d_t1; d_t1 = c; }
      // O^T: rows = channels, columns = heads; row c is taken at column head(c)
      __half* orow = s_out + (size_t)pl * C;
      const int rowv = (mi >> 1) * 8 + r8;             // slot addressed by this lane for the V~^T tile
#pragma unroll 4
      for (int dt = sub; dt < nks; dt += W) {
        uint32_t a[4], pe[4];
        // V~^T tile: matrices (slots 0-7, d 0-7) (slots 0-7, d 8-15) (slots 8-15, d 0-7) (slots 8-15, d 8-15), transposed
        const int chunk = 2 * dt + (mi & 1);
        km_ldsm_t(vbase + km_off(pl * KM_L + rowv, chunk, R), a);
        km_ldsm_t(pvb + km_off(rowv, chunk, KM_L), pe);
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = km_hadd2(a[r], pe[r]);             // V + V_pe -> fp16
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        km_mma(o, a, pb0, pb1);
        // lane (g, t): o[0] = O^T[dt*16+g][2t], o[1] = [..][2t+1], o[2] = O^T[dt*16+8+g][2t], o[3] = [..][2t+1]
        const int h0 = s_head[2 * dt], h1 = s_head[2 * dt + 1];
        const __half v0 = __float2half_rn((h0 & 1) ? o[1] : o[0]), v1 = __float2half_rn((h1 & 1) ? o[3] : o[2]);
        km_st_pred(km_smem_u32(orow + dt * 16 + g), v0, (h0 >> 1) == t);       // predicated, no divergent branch
        km_st_pred(km_smem_u32(orow + dt * 16 + 8 + g), v1, (h1 >> 1) == t);
      }
      if (dbg && tid == 0) { const long long c = clock64(); d_pv += c - d_t1; }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // ring patches precede the async refill
    __syncthreads();
    if (dbg && tid == 0) { const long long c = clock64(); d_comp += c - d_t0; d_t0 = c; }
    if (tid == PRODUCER_TID) issue(i + KM_STAGES);                 // stage consumed: fetch the tile after next
    // ---- coalesced 128-bit stores of the tile's output rows ----
    for (int idx = tid; idx < np * T; idx += blockDim.x) {
      const int pl = idx / T, c = idx - pl * T;
      const size_t row = (size_t)n * p.hw + p0 + pl;
      *reinterpret_cast<uint4*>(p.out + row * C + (size_t)c * 8) = *reinterpret_cast<const uint4*>(s_out + (size_t)pl * C + (size_t)c * 8);
    }
    if (dbg && tid == 0) d_store += clock64() - d_t0;
  }
  if (dbg && tid == 0) {
    long long* o = dbg + (size_t)blockIdx.x * 8;
    o[0] = d_wait; o[1] = d_patch; o[2] = d_comp; o[3] = d_store; o[4] = my_tiles; o[5] = d_qk; o[6] = d_sm; o[7] = d_pv;
  }
}

int g_km_sms = 0;
long long* g_km_dbg = nullptr;

}  // namespace

void kv_attn_set_debug(long long* ptr) { g_km_dbg = ptr; }

bool kv_attn_mma_supported(const KvAttnParams& p) {
  const int hd = p.heads > 0 ? p.C / p.heads : 0;
  // C <= 640: ring (2 x 80 KB) + the row's two PE windows (<= 40 KB) fit one SM; wider levels use the scalar kernel
  return p.L == KM_L && p.C % 64 == 0 && p.C <= 640 && p.heads >= 1 && p.heads <= 8 && p.C % p.heads == 0 &&
         hd % 8 == 0 && p.pe_ld % 8 == 0;
}

int kv_attn_mma_launch(const KvAttnParams& p0, cudaStream_t stream) {
  KvAttnParams p = p0;
  const int hd = p.C / p.heads;
  p.T = p.C / 8;
  p.hd8 = hd / 8;
  int P = 1280 / p.C;                        // 80 KB of K+V per tile: 4 / 2 pixels at C = 320 / 640
  if (P >= 8) P = 8; else if (P >= 4) P = 4; else if (P >= 2) P = 2; else P = 1;
  if (P < 1) P = 1;
  if (P > 8) P = 8;                          // 8 math warps, 8 / P of them per pixel
  while (P > p.hw) P >>= 1;                  // (stays a power of two)
  if (P < 1) P = 1;
  p.P = P;
  p.scale = 1.0f / sqrtf((float)hd);
  const int ncb = p.C / 64;
  const size_t stage_bytes = (size_t)2 * ncb * P * KM_L * 128;
  const size_t smem = KM_STAGES * stage_bytes + (size_t)2 * ncb * KM_L * 128 + (size_t)2 * P * p.C * sizeof(__half) +
                      KM_L * 8 + 192 + 8 * 32 * 16 + KM_STAGES * 8 + 64 + 1024;
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
  static size_t configured = 0;
  if (smem > configured) {
    L2D_CUDA(cudaFuncSetAttribute(kv_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int tiles_per_row = ceil_div(p.hw, P);
  const int total = tiles_per_row * p.n_rows;
  const int grid = total < g_km_sms ? total : g_km_sms;
  kv_attn_mma_kernel<<<grid, 288, smem, stream>>>(tm, p, tiles_per_row, ncb, g_km_dbg);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

}  // namespace l2d
