// K1, tensor-core formulation (window L == 16, C % 64 == 0, heads <= 8): same semantics as kv_attn.cu
// (stream_motion_module.py:117-194) at a fraction of the issue slots per byte.
//
// Why: the scalar kernel needs ~2000 warp-instructions per 80 KB tile (fp16->fp32 converts + FMAs per element); with
// the few warps an SM can hold next to a 160 KB ring that caps it near 25 % of HBM bandwidth (profiles/).  Here one
// mma.sync.m16n8k16 evaluates 16 slots x 16 channels for all 8 heads at once (heads on the N dimension through a
// block-diagonal q matrix; see the kernel comment).  K+pe / V+pe are formed on the A fragments with HADD2, i.e. with
// the reference's fp16 rounding; P is rounded to fp16 for P.V like the fused SDPA kernels the reference dispatches to.
//
// Data movement: the cache [N,2,hw,L,C] viewed as a matrix [N*2*hw*L rows, C cols]; a tile is P pixels (P*C = 1280 at
// the UNet's widths); its K plane and its V plane are each R = P*16 consecutive rows, fetched as 64-column TMA boxes
// with the 128B swizzle (cp.async.bulk.tensor.2d -> conflict-free ldmatrix) into a ring of NB plane buffers.
// Warp-specialised pipeline: warp 8 is the TMA producer (full/empty mbarriers per buffer); the 8 math warps are split
// into P pixel groups of W = 8/P warps that only synchronise among themselves (named barriers), so groups drift apart
// and overlap each other's latencies.  The freshly projected k/v/q of the next tile are prefetched into registers.
// Precondition (as in the reference, whose cache is zero-initialised): masked slots hold finite values.
//
// Window L == 32 (BASELINE config 4, LB = 2): a pixel's 32 slots are two consecutive 16-row blocks of the same matrix, so
// a tile holds "half-pixels" -- pixel group pl works on block kb = pl % 2 of real pixel pl / 2 with the L == 16 code path
// unchanged (own K/V rows, own PE-window block, own 16 mask bits); the two groups of a pixel meet twice per tile on a
// named barrier to exchange the per-head max and sum of their halves (so P is the softmax over all 32 slots), park their
// partial O^T in their own V windows, and split the gather of the summed output row between them.
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
__device__ __forceinline__ void km_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void km_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// named barrier among `count` threads (a pixel group, or the 8 math warps)
__device__ __forceinline__ void km_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// (volatile keeps the program order against the barriers in NVVM; ptxas schedules the loads freely inside a phase)
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
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void km_mma0(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {   // C = 0
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ uint32_t km_hadd2(uint32_t a, uint32_t b) { return h2_as_u32(__hadd2(u32_as_h2(a), u32_as_h2(b))); }

constexpr int KM_L = 16;
constexpr int KM_THREADS = 288;     // 8 math warps + the TMA warp

// swizzled byte offset of 16-byte chunk `chunk` (= channel / 8) of tile row `row` inside one plane:
// 64-column blocks of R rows x 128 B, 128B swizzle = chunk-in-block XOR (row % 8)
__device__ __forceinline__ uint32_t km_off(int row, int chunk, int R) {
  return (uint32_t)((chunk >> 3) * (R * 128) + row * 128 + (((chunk & 7) ^ (row & 7)) << 4));
}

// Fragment addressing of the specialised kernels.  Warp `sub` of a pixel group takes k-steps ks = sub + W*i, i.e. the
// 16-byte chunks c0 + 2*W*i with c0 = 2*sub + j (j = which half of the k-step the lane's ldmatrix row addresses), so the
// offset of step i is a per-lane base plus a compile-time constant:
//   W = 2: chunk = c0 + 4i  -> block i/2, chunk-in-block c0 (+4 for odd i: one xor on the swizzled chunk)
//   W = 4: chunk = c0 + 8i  -> block i,   chunk-in-block c0
//   W = 8: chunk = c0 + 16i -> block 2i + c0/8, chunk-in-block c0 % 8
struct KmFrag {
  uint32_t e, o;
};
template <int W>
__device__ __forceinline__ KmFrag km_frag(int row, int c0, int R) {
  KmFrag f;
  const int r7 = row & 7;
  if (W == 8) {
    f.e = (uint32_t)((c0 >> 3) * (R * 128) + row * 128 + (((c0 & 7) ^ r7) << 4));
    f.o = f.e;
  } else {
    f.e = (uint32_t)(row * 128 + ((c0 ^ r7) << 4));
    f.o = (uint32_t)(row * 128 + (((c0 ^ r7) ^ 4) << 4));
  }
  return f;
}
template <int W>
__device__ __forceinline__ uint32_t km_frag_off(const KmFrag& f, int i, int R) {
  if (W == 2) return (uint32_t)((i >> 1) * (R * 128)) + ((i & 1) ? f.o : f.e);
  if (W == 4) return (uint32_t)(i * (R * 128)) + f.e;
  return (uint32_t)(2 * i * (R * 128)) + f.e;
}

// The 8 heads sit on the N = 8 dimension of the MMA:
//   S[16 slots x 8 heads] = K~[16 x C] . Qblk[C x 8]     Qblk[c][h] = q~[c] if channel c belongs to head h else 0
//   O^T[C x 8 heads]      = V~^T[C x 16] . P[16 x 8]      row c of O^T is read at column head(c)
// so every B column does useful work, the softmax of all heads runs in one set of registers (scores of slot g / g+8
// and heads 2t / 2t+1 per lane; the reduction over slots is 3 xor-shuffles for all 8 heads at once), and one pixel
// costs C/16 k-steps of (2 ldmatrix + 4 HADD2 + 1 mma) for the scores and again for P.V.  The k-steps of a pixel are
// dealt round-robin to the W warps of its group; partial scores meet in shared memory.
// PE: the row's K_pe[pi[j]] / V_pe[pi[j]] windows are staged once per row n in shared memory in the same swizzled
// [16 x C] layout as a pixel's K / V window, so the PE fragments come from the same ldmatrix addresses.
// O^T staging: the 16 bytes (8 heads) of O^T row c = 16*ks + r overwrite the V window at (slot r, channels 16*ks..+7),
// bytes only this warp reads and has already consumed for that very k-step; swizzled like the window, so the writes and
// the gather are conflict-free.  The K plane is therefore free right after the scores and is released early.
// CT/PT: compile-time C and pixels per tile (0 = run-time geometry, any C <= 640); NB: plane buffers in the ring.
template <int CT, int PT, int NB, int LB = 1>
__global__ void __launch_bounds__(KM_THREADS, 1)
kv_attn_mma_kernel(const __grid_constant__ CUtensorMap tmap, const KvAttnParams p, const int tiles_per_row, long long* dbg) {
  static_assert(LB == 1 || (CT != 0 && PT % 2 == 0), "the 32-slot path exists for the specialised geometries only");
  extern __shared__ __align__(1024) uint8_t km_smem_raw[];
  // align by pointer arithmetic (an integer round trip would turn every later access into a generic LD/ST)
  uint8_t* smem = km_smem_raw + ((1024u - (km_smem_u32(km_smem_raw) & 1023u)) & 1023u);
  constexpr bool SPEC = CT != 0;
  constexpr int WT = SPEC ? 8 / (PT ? PT : 1) : 1;
  constexpr int NKW = SPEC ? (CT / 16) / WT : 1;             // k-steps per warp (10 at every UNet width)
  const int C = SPEC ? CT : p.C, P = SPEC ? PT : p.P;
  const int W = 8 / P, R = P * KM_L, T = C >> 3, ncb = C >> 6, nks = C >> 4;
  const uint32_t plane_bytes = (uint32_t)ncb * R * 128;          // K (or V) rows of one tile
  const uint32_t pe_plane = (uint32_t)ncb * KM_L * 128;          // one [16 x C] PE window
  const int hwp = p.hw * LB;                                     // (half-)pixels per row: the unit tiles are made of
  const int LW = KM_L * LB;                                      // window length
  uint8_t* ring = smem;
  uint8_t* pek = ring + (size_t)NB * plane_bytes;                // K_pe[pi[j]] windows of the current row n: LB blocks of [16 x C]
  uint8_t* pev = pek + LB * pe_plane;
  __half* s_q = reinterpret_cast<__half*>(pev + LB * pe_plane);  // [P][C]  q + Q_pe
  __half* s_qpe = s_q + (size_t)P * C;                           // [C]     Q_pe[pi[u]] of the current row n
  __half* s_vn = s_qpe + C;                                      // [2][P][C] freshly projected v of this / the next tile
  float* s_part = reinterpret_cast<float*>(s_vn + (size_t)2 * P * C);   // [8 warps][32 lanes][4] partial scores
  float* s_mask = s_part + 8 * 32 * 4;                           // [32]
  int* s_pi = reinterpret_cast<int*>(s_mask + 2 * KM_L);         // [32]
  int* s_misc = s_pi + 2 * KM_L;                                 // [0] = write slot u of the current row n
  float* s_xm = reinterpret_cast<float*>(s_misc + 4);            // [8 groups][8 heads] half-window max  (LB == 2)
  float* s_xs = s_xm + 64;                                       // [8][8] half-window sum
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_xs + 64);       // full[NB], empty[NB], go
  uint8_t* s_head = reinterpret_cast<uint8_t*>(bars + 2 * NB + 1);   // [T] head of each 8-channel chunk

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long d_entry = 0, d_gt0 = 0;
  if (dbg != nullptr && tid == 0) {
    d_entry = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(d_gt0));
  }
  const int S = p.c_slices;                                       // channel slices: "virtual row" nv = n * S + slice
  const int CF = p.c_full;                                       // row pitch (channels) of the cache / q / k / v / out rows
  const int nrv = p.n_rows * S;
  const int total_tiles = tiles_per_row * nrv;
  // contiguous tile range per CTA; when the grid divides into the rows no CTA straddles a row change (one PE staging)
  int t_begin, t_end;
  if ((int)gridDim.x % nrv == 0) {
    const int cpr = (int)gridDim.x / nrv, n0 = (int)blockIdx.x / cpr, r0 = (int)blockIdx.x - n0 * cpr;
    t_begin = n0 * tiles_per_row + (int)(((long long)tiles_per_row * r0) / cpr);
    t_end = n0 * tiles_per_row + (int)(((long long)tiles_per_row * (r0 + 1)) / cpr);
  } else {
    t_begin = (int)(((long long)total_tiles * blockIdx.x) / gridDim.x);
    t_end = (int)(((long long)total_tiles * (blockIdx.x + 1)) / gridDim.x);
  }
  const int my_tiles = t_end - t_begin;
  const uint32_t bar0 = km_smem_u32(bars);
  auto full_bar = [&](int b) { return bar0 + 8u * b; };
  auto empty_bar = [&](int b) { return bar0 + 8u * (NB + b); };
  const uint32_t go_bar = bar0 + 8u * (2 * NB);                  // math warps -> producer: the row's PE copies are queued

  // the first row's schedule is requested at kernel entry so its latency hides behind the prologue
  int e_pi = 0, e_u = 0;
  float e_mask = 0.f;
  if (tid < LW && my_tiles > 0) {
    const int n0 = (t_begin / tiles_per_row) / S;
    e_pi = static_cast<int>(p.pe_idx[(size_t)n0 * LW + tid]);
    e_mask = __half2float(p.mask[(size_t)n0 * LW + tid]);
    e_u = static_cast<int>(p.update_idx[n0]);
  }
  pdl_launch();
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    for (int b = 0; b < NB; ++b) {
      km_mbar_init(full_bar(b), 1);
      km_mbar_init(empty_bar(b), 8);                             // one arrival per math warp
    }
    km_mbar_init(go_bar, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    const int hd = C / p.heads;
    for (int c = tid; c < T; c += blockDim.x) s_head[c] = (uint8_t)((c * 8) / hd);
  }
  __syncthreads();
  if (my_tiles <= 0) return;

  if (warp == 8) {   // ---- TMA producer: planes in order K(0) V(0) K(1) V(1) ...; plane s lives in buffer s % NB ----
    if (lane == 0) {
      const int planes = 2 * my_tiles;
      int plane_col = 0;                                         // first channel of the tile's slice
      auto plane_row = [&](int s) {
        const int tile = t_begin + (s >> 1);
        const int nv = tile / tiles_per_row, p0 = (tile - nv * tiles_per_row) * P;
        plane_col = (nv % S) * C;
        return (((nv / S) * 2 + (s & 1)) * hwp + p0) * KM_L;
      };
      // The K plane of the first tile is needed before anything else: it is requested at once.  The rest of the first burst
      // (NB planes from every SM, ~20 MB) would sit in front of the math warps' small dependent loads (index tensors -> PE
      // rows), so it waits until those are queued.  (L2D_K1_EAGER_PLANES, developer A/B: planes requested before that.)
      const int eager = p.eager_planes;
      for (int s = 0; s < planes; ++s) {
        if (s == eager) km_mbar_wait(go_bar, 0);
        const int b = s % NB, use = s / NB;
        if (use > 0) km_mbar_wait(empty_bar(b), (uint32_t)(use - 1) & 1u);
        const int row0 = plane_row(s);
        km_mbar_expect_tx(full_bar(b), plane_bytes);
        const uint32_t dst = km_smem_u32(ring + (size_t)b * plane_bytes);
        for (int b2 = 0; b2 < ncb; ++b2) km_tma_2d(dst + b2 * (R * 128), &tmap, full_bar(b), plane_col + b2 * 64, row0);
      }
    }
    return;
  }

  // ---- math warps: group pl = warp / W owns pixel pl of every tile ----
  const int pl = warp / W, sub = warp - pl * W;
  const int GT = W * 32;                                         // threads of a pixel group
  const int gbar = 1 + pl;                                       // the group's named barrier
  const int kb = LB == 2 ? (pl & 1) : 0;                         // which 16-slot block of the pixel's window this group owns
  const int pbar = 9 + (pl >> 1);                                // LB == 2: named barrier of the two groups of one pixel
  const int cg = sub * 32 + lane;                                // the 16-byte chunk this thread appends / stages
  const bool has_chunk = cg < T;
  const int g = lane >> 2, t = lane & 3, mi = lane >> 3, r8 = lane & 7;
  const int rowk = (mi & 1) * 8 + r8, jq = mi >> 1;              // K~ tile: matrices (slots lo/hi) x (channels lo/hi)
  const int rowv = (mi >> 1) * 8 + r8, jv = mi & 1;              // V~^T tile (transposed load)
  const int nkw = SPEC ? NKW : (nks - sub + W - 1) / W;

  // per-lane addressing constants of the specialised kernels (see km_frag)
  KmFrag fk, fpk, fv, fpv;
  KmFrag fo;                                                     // O^T staging: (slot g, chunk 2*ks) of the V window
  uint32_t qm[2 * NKW];                                          // Qblk column masks of this lane's head g
  if constexpr (SPEC) {
    constexpr int RT = PT * KM_L;
    fk = km_frag<WT>(pl * KM_L + rowk, 2 * sub + jq, RT);
    fpk = km_frag<WT>(rowk, 2 * sub + jq, KM_L);
    fv = km_frag<WT>(pl * KM_L + rowv, 2 * sub + jv, RT);
    fpv = km_frag<WT>(rowv, 2 * sub + jv, KM_L);
    fo = km_frag<WT>(pl * KM_L + g, 2 * sub, RT);
#pragma unroll
    for (int i = 0; i < NKW; ++i) {
      qm[2 * i] = (int)s_head[2 * (sub + WT * i)] == g ? 0xffffffffu : 0u;
      qm[2 * i + 1] = (int)s_head[2 * (sub + WT * i) + 1] == g ? 0xffffffffu : 0u;
    }
  }
  auto ot_off = [&](int i) -> uint32_t {                         // byte offset of O^T row (16*ks + g), heads 2t, 2t+1
    if constexpr (SPEC) {
      constexpr int RT = PT * KM_L;
      return km_frag_off<WT>(fo, i, RT) + 4 * t;
    } else {
      return km_off(pl * KM_L + g, 2 * (sub + W * i), R) + 4 * t;
    }
  };

  // k/v/q chunk of the NEXT tile, requested one tile ahead
  uint4 pf_k = make_uint4(0, 0, 0, 0), pf_q = pf_k;
  // pixel row of this group in tile `tile2` (-1: past the ragged end of a row); no division when P divides hw
  const bool flat = S == 1 && (hwp % P) == 0;
  auto pixel_row = [&](int tile2) -> long long {   // row of this group's REAL pixel in q / k_new / v_new / out
    if (flat) return ((long long)tile2 * P + pl) / LB;
    const int n2 = tile2 / tiles_per_row, q0 = (tile2 - n2 * tiles_per_row) * P;
    return q0 + pl < hwp ? (long long)(n2 / S) * p.hw + (q0 + pl) / LB : -1;
  };
  auto slice_col = [&](int tile2) -> size_t { return (size_t)((tile2 / tiles_per_row) % S) * C; };
  auto load_kq = [&](int tile2) {
    const long long row = pixel_row(tile2);
    if (has_chunk && row >= 0) {
      const size_t col = (S > 1 ? slice_col(tile2) : 0) + (size_t)cg * 8;
      pf_k = ldg_act(p.k_new + (size_t)row * p.ld + col);
      pf_q = ldg_act(p.q + (size_t)row * p.ld + col);
    }
  };
  // v is needed half a tile later than k / q and would pin four more registers for a whole tile: it is staged through
  // shared memory with cp.async instead (each thread copies and later reads its own chunk: no barrier involved)
  auto load_v = [&](int tile2, int buf) {
    const long long row = tile2 < t_end ? pixel_row(tile2) : -1;
    if (has_chunk && row >= 0) {
      km_cp_async16(km_smem_u32(s_vn + ((size_t)buf * P + pl) * C + (size_t)cg * 8),
                    p.v_new + (size_t)row * p.ld + (S > 1 ? slice_col(tile2) : 0) + (size_t)cg * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // one group per tile, also when empty: wait_group 1 below counts them
  };
  // PDL: everything above -- and the producer warp's KV-plane TMA loads plus the PE-window staging below -- reads only
  // memory that no neighbouring kernel writes (this module's cache, PE tables, schedule tensors) and overlaps the tail of
  // the QKV GEMM; the freshly projected q / k / v, the cache slot append and the output row come after this wait.
  // per-row schedule + PE windows (block-uniform; all 8 math warps): Q_pe row and the K_pe / V_pe windows by cp.async
  int cur_n = -1;
  bool pe_pending = false;
  auto stage_row = [&](int nv) {                             // nv = virtual row: (denoise row n, channel slice)
    const int n = nv / S;
    const size_t coff = (size_t)(nv % S) * C;
    if (cur_n >= 0) km_bar(15, 256);                         // every group is done with the previous row's windows
    if (tid < LW) {
      if (cur_n >= 0) {
        e_pi = static_cast<int>(p.pe_idx[(size_t)n * LW + tid]);
        e_mask = __half2float(p.mask[(size_t)n * LW + tid]);
        e_u = static_cast<int>(p.update_idx[n]);
      }
      s_pi[tid] = e_pi;
      s_mask[tid] = e_mask;
      if (tid == 0) s_misc[0] = e_u;
    }
    km_bar(15, 256);
    // window slot j <- table row pe_idx[n][j], copied asynchronously (no registers, no wait here): the copies are
    // queued ahead of the producer's first burst and land while the first K plane is in flight
    const int u0 = s_misc[0];
    for (int idx = tid; idx < LW * T; idx += 256) {
      const int j = idx / T, c = idx - j * T;
      const uint32_t off = (uint32_t)(j >> 4) * pe_plane + km_off(j & 15, c, KM_L);   // block j / 16, slot j % 16
      km_cp_async16(km_smem_u32(pek + off), p.k_pe + (size_t)s_pi[j] * p.pe_ld + coff + (size_t)c * 8);
      km_cp_async16(km_smem_u32(pev + off), p.v_pe + (size_t)s_pi[j] * p.pe_ld + coff + (size_t)c * 8);
    }
    for (int c = tid; c < T; c += 256)
      km_cp_async16(km_smem_u32(s_qpe + (size_t)c * 8), p.q_pe + (size_t)s_pi[u0] * p.pe_ld + coff + (size_t)c * 8);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (cur_n < 0) {
      __syncwarp();
      if (lane == 0) km_mbar_arrive(go_bar);
    }
    pe_pending = true;
    cur_n = nv;
  };
  stage_row(t_begin / tiles_per_row);   // first row: queued before the PDL wait (and ahead of the producer's first burst)
  pdl_wait();
  load_kq(t_begin);
  load_v(t_begin, 0);

  // gather addresses of this thread's (up to 3) channel pairs inside the V plane, incl. the head column: fixed per kernel.
  // LB == 2: the 2*GT threads of a pixel's two groups share the row; every thread adds the two groups' partial O^T, which
  // sit at the same swizzled offset of their windows, 16 rows (2 KB) apart
  const int cgx = LB == 2 ? kb * GT + cg : cg;
  const int GS = LB * GT;
  const int partner = LB == 2 ? ((pl & 1) ? -16 * 128 : 16 * 128) : 0;
  uint32_t g_lo[3], g_hi[3];
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    const int c = 2 * (cgx + GS * m);
    g_lo[m] = g_hi[m] = 0xffffffffu;
    if (c < C) {
      const int hb = 2 * (int)s_head[c >> 3], row = pl * KM_L + (c & 15), ch = 2 * (c >> 4);
      g_lo[m] = km_off(row, ch, R) + hb;
      g_hi[m] = km_off(row + 1, ch, R) + hb;
    }
  }
  const bool tl = dbg != nullptr && tid == 0;
  long long d_wait = 0, d_patch = 0, d_qk = 0, d_mid = 0, d_pv = 0, d_store = 0, d_t = 0, d_wv = 0, d_bar2 = 0, d_gather = 0;   // developer timeline
  auto stamp = [&](long long& acc) {
    if (tl) {
      const long long c = clock64();
      acc += c - d_t;
      d_t = c;
    }
  };

  const long long d_loop0 = tl ? clock64() : 0;
  for (int i = 0; i < my_tiles; ++i) {
    const int tile = t_begin + i;
    const int nv = tile / tiles_per_row, p0 = (tile - nv * tiles_per_row) * P;
    const int n = nv / S;
    const size_t coff = (size_t)(nv % S) * C;                // first channel of this tile's slice
    const bool active = p0 + pl < hwp;
    if (nv != cur_n) stage_row(nv);
    const int u_full = s_misc[0];
    const int u = u_full & (KM_L - 1);                      // slot inside this group's 16-slot block
    const bool owner = LB == 1 || (u_full >> 4) == kb;      // the group whose block receives the appended k / v
    const int sK = 2 * i, sV = 2 * i + 1;
    const int bK = sK % NB, bV = sV % NB;
    uint8_t* kplane = ring + (size_t)bK * plane_bytes;
    uint8_t* vplane = ring + (size_t)bV * plane_bytes;
    if (tl) d_t = clock64();
    km_mbar_wait(full_bar(bK), (uint32_t)(sK / NB) & 1u);
    if (pe_pending) {   // first tile of a row: this thread's PE copies have landed; meet so everyone's are visible
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      km_bar(15, 256);
      pe_pending = false;
    }
    stamp(d_wait);
    if (active) {
      // ---- K append (HBM + window patch) and q~ staging: thread <-> one 16-byte chunk of the group's pixel ----
      if (has_chunk) {
        if (owner) {
          __half* kdst = p.cache + ((((size_t)n * 2) * hwp + p0 + pl) * KM_L + u) * CF + coff + (size_t)cg * 8;
          *reinterpret_cast<uint4*>(kdst) = pf_k;                                            // PE-free append (:117-119)
          *reinterpret_cast<uint4*>(kplane + km_off(pl * KM_L + u, cg, R)) = pf_k;           // the window sees the new slot
        }
        *reinterpret_cast<uint4*>(s_q + (size_t)pl * C + (size_t)cg * 8) =
            hadd8(pf_q, *reinterpret_cast<const uint4*>(s_qpe + (size_t)cg * 8));            // q + Q_pe[pi[u]] -> fp16
      }
      km_bar(gbar, GT);
      stamp(d_patch);
      if (i + 1 < my_tiles) load_kq(tile + 1);
      load_v(tile + 1, (i + 1) & 1);

      // ---- scores: rows = slots, columns = heads; two accumulators halve the dependent mma chain ----
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
      {
        const uint32_t kbase = km_smem_u32(kplane), pkb = km_smem_u32(pek) + (uint32_t)kb * pe_plane;
        const __half* qrow = s_q + (size_t)pl * C + sub * 16 + 2 * t;
        auto qk_step = [&](int ii, uint32_t koff, uint32_t pkoff, uint32_t m0, uint32_t m1, float (&dst)[4]) {
          uint32_t a[4], pe[4];
          // matrices: (slots 0-7, ch 0-7) (slots 8-15, ch 0-7) (slots 0-7, ch 8-15) (slots 8-15, ch 8-15)
          km_ldsm(kbase + koff, a);
          km_ldsm(pkb + pkoff, pe);
#pragma unroll
          for (int r = 0; r < 4; ++r) a[r] = km_hadd2(a[r], pe[r]);             // K + K_pe -> fp16
          // B = Qblk: lane (g, t) holds rows (channels) ks*16 + 2t, +1 (b0) and + 8 (b1) of column (head) g
          const uint32_t q0 = *reinterpret_cast<const uint32_t*>(qrow + ii * (W * 16));
          const uint32_t q1 = *reinterpret_cast<const uint32_t*>(qrow + ii * (W * 16) + 8);
          km_mma(dst, a, q0 & m0, q1 & m1);
        };
        if constexpr (SPEC) {
          constexpr int RT = PT * KM_L;
#pragma unroll
          for (int ii = 0; ii < NKW; ++ii) {
            if (ii & 1) qk_step(ii, km_frag_off<WT>(fk, ii, RT), km_frag_off<WT>(fpk, ii, KM_L), qm[2 * ii], qm[2 * ii + 1], acc2);
            else qk_step(ii, km_frag_off<WT>(fk, ii, RT), km_frag_off<WT>(fpk, ii, KM_L), qm[2 * ii], qm[2 * ii + 1], acc);
          }
        } else {
#pragma unroll 2
          for (int ii = 0; ii < nkw; ++ii) {
            const int ks = sub + W * ii;
            const uint32_t m0 = (int)s_head[2 * ks] == g ? 0xffffffffu : 0u;
            const uint32_t m1 = (int)s_head[2 * ks + 1] == g ? 0xffffffffu : 0u;
            qk_step(ii, km_off(pl * KM_L + rowk, 2 * ks + jq, R), km_off(rowk, 2 * ks + jq, KM_L), m0, m1, acc);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] += acc2[r];
      if (W > 1) *reinterpret_cast<float4*>(s_part + ((size_t)warp * 32 + lane) * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      stamp(d_qk);

      // ---- V append + patch (the V plane was requested one plane after K), then the group meets ----
      km_mbar_wait(full_bar(bV), (uint32_t)(sV / NB) & 1u);
      asm volatile("cp.async.wait_group 1;" ::: "memory");          // this tile's v (the next tile's may be in flight)
      if (has_chunk && owner) {
        const uint4 pf_v = *reinterpret_cast<const uint4*>(s_vn + ((size_t)(i & 1) * P + pl) * C + (size_t)cg * 8);
        __half* vdst = p.cache + ((((size_t)n * 2 + 1) * hwp + p0 + pl) * KM_L + u) * CF + coff + (size_t)cg * 8;
        *reinterpret_cast<uint4*>(vdst) = pf_v;
        *reinterpret_cast<uint4*>(vplane + km_off(pl * KM_L + u, cg, R)) = pf_v;
      }
      stamp(d_wv);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the K patch (generic proxy) precedes the async refill
      km_bar(gbar, GT);
      if (lane == 0) km_mbar_arrive(empty_bar(bK));                   // scores done: the K plane can be refilled now
      stamp(d_bar2);
      if (W > 1) {   // sum the partial scores of this pixel's warps
        acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
        for (int w2 = 0; w2 < W; ++w2) {
          const float4 v4 = *reinterpret_cast<const float4*>(s_part + ((size_t)(pl * W + w2) * 32 + lane) * 4);
          acc[0] += v4.x; acc[1] += v4.y; acc[2] += v4.z; acc[3] += v4.w;
        }
      }
      // lane (g, t): acc[0] = S[slot g][head 2t], acc[1] = S[g][2t+1], acc[2] = S[g+8][2t], acc[3] = S[g+8][2t+1]
      const float m_lo = s_mask[kb * KM_L + g], m_hi = s_mask[kb * KM_L + g + 8];
      float s0 = fmaf(acc[0], p.scale, m_lo), s1 = fmaf(acc[1], p.scale, m_lo);
      float s2 = fmaf(acc[2], p.scale, m_hi), s3 = fmaf(acc[3], p.scale, m_hi);
      float mxa = fmaxf(s0, s2), mxb = fmaxf(s1, s3);                        // heads 2t and 2t+1
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {                                    // reduce over g (the slots)
        mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, o));
        mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, o));
      }
      if constexpr (LB == 2) {   // the other half of the window: exchange the per-head max with the partner group
        if (sub == 0 && g == 0) {
          s_xm[pl * 8 + 2 * t] = mxa;
          s_xm[pl * 8 + 2 * t + 1] = mxb;
        }
        km_bar(pbar, 2 * GT);
        mxa = fmaxf(mxa, s_xm[(pl ^ 1) * 8 + 2 * t]);
        mxb = fmaxf(mxb, s_xm[(pl ^ 1) * 8 + 2 * t + 1]);
      }
      s0 = __expf(s0 - mxa); s2 = __expf(s2 - mxa);
      s1 = __expf(s1 - mxb); s3 = __expf(s3 - mxb);
      float suma = s0 + s2, sumb = s1 + s3;
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {
        suma += __shfl_xor_sync(0xffffffffu, suma, o);
        sumb += __shfl_xor_sync(0xffffffffu, sumb, o);
      }
      if constexpr (LB == 2) {   // ... and the per-head sum: P is the softmax over all 32 slots
        if (sub == 0 && g == 0) {
          s_xs[pl * 8 + 2 * t] = suma;
          s_xs[pl * 8 + 2 * t + 1] = sumb;
        }
        km_bar(pbar, 2 * GT);
        suma += s_xs[(pl ^ 1) * 8 + 2 * t];
        sumb += s_xs[(pl ^ 1) * 8 + 2 * t + 1];
      }
      const float inva = __frcp_rn(suma), invb = __frcp_rn(sumb);            // sums are in [1, 32]
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
      stamp(d_mid);

      // ---- O^T: rows = channels, columns = heads; every lane parks its two head columns of rows g and g + 8 in the
      //      pixel's (dead) K window, the gather below picks column head(c) ----
      {
        const uint32_t vb = km_smem_u32(vplane), pvb = km_smem_u32(pev) + (uint32_t)kb * pe_plane;
        auto pv_step = [&](uint32_t voff, uint32_t pvoff, uint32_t ooff) {
          uint32_t a[4], pe[4];
          // V~^T tile: matrices (slots 0-7, d 0-7) (slots 0-7, d 8-15) (slots 8-15, d 0-7) (slots 8-15, d 8-15), transposed
          km_ldsm_t(vb + voff, a);
          km_ldsm_t(pvb + pvoff, pe);
#pragma unroll
          for (int r = 0; r < 4; ++r) a[r] = km_hadd2(a[r], pe[r]);             // V + V_pe -> fp16
          float o[4];
          km_mma0(o, a, pb0, pb1);
          // The stores below overwrite bytes of the V window that OTHER lanes of this warp supplied to the ldmatrix above.
          // The data dependency ldmatrix -> HADD2 -> mma -> store already orders them in hardware; the __syncwarp makes the
          // ordering formal for the memory model (compute-sanitizer racecheck reported the pair as a WAR hazard without it).
          __syncwarp();
          // lane (g, t): o[0] = O^T[dt*16+g][2t], o[1] = [..][2t+1], o[2] = O^T[dt*16+8+g][2t], o[3] = [..][2t+1]
          *reinterpret_cast<uint32_t*>(vplane + ooff) = h2_as_u32(__floats2half2_rn(o[0], o[1]));
          *reinterpret_cast<uint32_t*>(vplane + ooff + 8 * 128) = h2_as_u32(__floats2half2_rn(o[2], o[3]));   // slot g + 8
        };
        if constexpr (SPEC) {
          constexpr int RT = PT * KM_L;
#pragma unroll
          for (int ii = 0; ii < NKW; ++ii) pv_step(km_frag_off<WT>(fv, ii, RT), km_frag_off<WT>(fpv, ii, KM_L), ot_off(ii));
        } else {
#pragma unroll 2
          for (int ii = 0; ii < nkw; ++ii) {
            const int dt = sub + W * ii;
            pv_step(km_off(pl * KM_L + rowv, 2 * dt + jv, R), km_off(rowv, 2 * dt + jv, KM_L), ot_off(ii));
          }
        }
      }
      if constexpr (LB == 2) km_bar(pbar, 2 * GT);   // both halves' partial O^T are parked
      else km_bar(gbar, GT);
      stamp(d_pv);
      // ---- gather column head(c) of O^T and store the pixel's output row (two channels per thread) ----
      {
        __half* orow = p.out + ((size_t)n * p.hw + (p0 + pl) / LB) * CF + coff;
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          if (g_lo[m] != 0xffffffffu) {
            uint32_t lo = *reinterpret_cast<const unsigned short*>(vplane + g_lo[m]);
            uint32_t hi = *reinterpret_cast<const unsigned short*>(vplane + g_hi[m]);
            if constexpr (LB == 2) {   // sum of the two 16-slot halves (fp32 add of the two fp16 partials, one rounding)
              const __half2 mine = u32_as_h2(lo | (hi << 16));
              const uint32_t lo2 = *reinterpret_cast<const unsigned short*>(vplane + g_lo[m] + partner);
              const uint32_t hi2 = *reinterpret_cast<const unsigned short*>(vplane + g_hi[m] + partner);
              const float2 fa = __half22float2(mine), fb = __half22float2(u32_as_h2(lo2 | (hi2 << 16)));
              *reinterpret_cast<uint32_t*>(orow + 2 * (cgx + GS * m)) = h2_as_u32(__floats2half2_rn(fa.x + fb.x, fa.y + fb.y));
            } else {
              *reinterpret_cast<uint32_t*>(orow + 2 * (cg + GT * m)) = lo | (hi << 16);
            }
          }
        }
      }
      stamp(d_gather);
    } else {   // ragged last tile of a row: nothing to compute, but stay inside the pipeline depth (empty counts)
      km_mbar_wait(full_bar(bV), (uint32_t)(sV / NB) & 1u);
      if (lane == 0) km_mbar_arrive(empty_bar(bK));
      if (i + 1 < my_tiles) load_kq(tile + 1);
      load_v(tile + 1, (i + 1) & 1);
    }
    // ---- release the V plane: window patch and O^T staging (generic proxy) precede the async refill ----
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) km_mbar_arrive(empty_bar(bV));
    stamp(d_store);
  }
  if (tl) {
    long long* o = dbg + (size_t)blockIdx.x * 16;
    o[0] = d_wait; o[1] = d_patch; o[2] = d_qk; o[3] = d_wv; o[4] = my_tiles;
    o[5] = d_bar2; o[6] = d_mid; o[7] = d_pv; o[8] = d_gather; o[9] = d_store;
    o[10] = d_loop0 - d_entry;            // prologue: barrier init, head table, first prefetch
    o[11] = clock64() - d_entry;          // whole CTA
    long long gt1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
    o[12] = d_gt0; o[13] = gt1;
  }
}

int g_km_sms = 0;
long long* g_km_dbg = nullptr;

template <int CT, int PT, int NB, int LB = 1>
int km_launch(const CUtensorMap& tm, const KvAttnParams& p, int tiles_per_row, int grid, size_t smem, cudaStream_t stream) {
  static size_t configured = 0;
  if (smem > configured) {
    L2D_CUDA(cudaFuncSetAttribute(kv_attn_mma_kernel<CT, PT, NB, LB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  launch_pdl_if(p.pdl != 0 && pdl_family(3), kv_attn_mma_kernel<CT, PT, NB, LB>, dim3(grid), dim3(KM_THREADS), smem, stream, tm, p, tiles_per_row,
                g_km_dbg);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

}  // namespace

void kv_attn_set_debug(long long* ptr) { g_km_dbg = ptr; }

// L2D_K1_SLICE=1 (developer A/B): at L == 16 split C = 1280 rows into two 640-channel slices of one pixel each (more,
// smaller tiles at the two coarse levels); default off
static bool k1_slice16() {
  static const bool on = [] {
    const char* e = getenv("L2D_K1_SLICE");
    return e && e[0] == '1';
  }();
  return on;
}

bool kv_attn_mma_supported(const KvAttnParams& p) {
  const int hd = p.heads > 0 ? p.C / p.heads : 0;
  // L == 16: C <= 640: ring of 4 x 40 KB planes + the row's two PE windows (<= 40 KB); C = 1280: 3 planes + 80 KB of PE windows.
  // L == 32 (two 16-slot blocks per pixel): C = 320 / 640, and C = 1280 as two 640-channel slices of whole heads.
  // Other widths / window lengths take the scalar kernel.
  const bool common = p.C % 64 == 0 && p.heads >= 1 && p.heads <= 8 && p.C % p.heads == 0 && hd % 8 == 0 && p.pe_ld % 8 == 0;
  if (p.L == KM_L) return common && (p.C <= 640 || p.C == 1280);
  if (p.L == 2 * KM_L)
    return common && p.hw >= 2 && (p.C == 320 || p.C == 640 || (p.C == 1280 && p.heads % 2 == 0 && 640 % hd == 0));
  return false;
}

int kv_attn_mma_launch(const KvAttnParams& p0, cudaStream_t stream) {
  KvAttnParams p = p0;
  const int hd = p.C / p.heads;
  const int LB = p.L / KM_L;                 // 16-slot blocks per pixel window
  const int hwp = p.hw * LB;                 // tiles are made of (half-)pixels of 16 slots
  const int c_full = p.C;
  // channel slices: a slice of whole heads is scheduled like a row of its own (own PE windows, own tiles)
  const bool sliced = p.C == 1280 && p.heads % 2 == 0 && 640 % hd == 0 && (LB == 2 || k1_slice16());
  if (sliced) {
    p.C = 640;
    p.heads /= 2;
  }
  p.c_slices = sliced ? 2 : 1;
  p.c_full = c_full;
  {
    static const int eager = [] {
      const char* e = getenv("L2D_K1_EAGER_PLANES");
      return e ? atoi(e) : 1;
    }();
    p.eager_planes = eager;
  }
  p.T = p.C / 8;
  p.hd8 = hd / 8;
  int P = 1280 / p.C;                        // 80 KB of K+V per tile: 4 / 2 / 1 (half-)pixels at C = 320 / 640 / 1280
  if (P >= 8) P = 8; else if (P >= 4) P = 4; else if (P >= 2) P = 2; else P = 1;   // 8 math warps, 8 / P per pixel
  if (sliced && LB == 1) P = 1;              // L == 16 slices: one pixel (40 KB of K+V) per tile
  while (P > hwp) P >>= 1;
  if (P < 1) P = 1;
  p.P = P;
  p.scale = 1.0f / sqrtf((float)hd);
  const int ncb = p.C / 64;
  const int NB = (p.C == 1280 || (LB == 2 && p.C == 640)) ? 3 : 4;
  const size_t plane_bytes = (size_t)ncb * P * KM_L * 128;
  const size_t smem = NB * plane_bytes + (size_t)2 * LB * ncb * KM_L * 128 + (size_t)P * p.C * sizeof(__half) +
                      (size_t)p.C * sizeof(__half) + (size_t)2 * P * p.C * sizeof(__half) + 8 * 32 * 16 + 2 * KM_L * 8 + 16 + 128 * 4 +
                      (2 * NB + 1) * 8 + 192 + 1024;
  if (smem > 227 * 1024) return fail(L2D_ERR_INVALID, "kv_attn(mma): tile does not fit in shared memory");
  CUtensorMap tm;
  const int64_t rows = (int64_t)p.n_rows * 2 * hwp * KM_L;
  int rc = get_tmap_2d(p.cache, rows, c_full, c_full, P * KM_L, &tm);
  if (rc != L2D_OK) return rc;
  if (g_km_sms == 0) {
    int dev = 0;
    L2D_CUDA(cudaGetDevice(&dev));
    L2D_CUDA(cudaDeviceGetAttribute(&g_km_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int tiles_per_row = ceil_div(hwp, P);
  const int nrv = p.n_rows * p.c_slices;
  const int total = tiles_per_row * nrv;
  int grid = total < g_km_sms ? total : g_km_sms;
  if (grid >= nrv) grid -= grid % nrv;                          // whole CTAs per (row, slice): no CTA stages two sets of PE windows
  if (LB == 2) {
    if (p.C == 320 && P == 4) return km_launch<320, 4, 4, 2>(tm, p, tiles_per_row, grid, smem, stream);
    if (p.C == 640 && P == 2) return km_launch<640, 2, 3, 2>(tm, p, tiles_per_row, grid, smem, stream);
    return fail(L2D_ERR_INVALID, "kv_attn(mma): unsupported 32-slot geometry");
  }
  if (p.C == 320 && P == 4) return km_launch<320, 4, 4>(tm, p, tiles_per_row, grid, smem, stream);
  if (p.C == 640 && P == 2) return km_launch<640, 2, 4>(tm, p, tiles_per_row, grid, smem, stream);
  if (p.C == 640 && P == 1 && sliced) return km_launch<640, 1, 4>(tm, p, tiles_per_row, grid, smem, stream);
  if (p.C == 1280 && P == 1) return km_launch<1280, 1, 3>(tm, p, tiles_per_row, grid, smem, stream);
  if (p.C > 640) return fail(L2D_ERR_INVALID, "kv_attn(mma): unsupported geometry");
  return km_launch<0, 0, 4>(tm, p, tiles_per_row, grid, smem, stream);
}

}  // namespace l2d
