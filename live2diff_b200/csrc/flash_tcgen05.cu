// K2/K3 on the 5th-generation tensor cores -- spatial self-attention (S x S) and text cross-attention (S x 77):
//     O = softmax(Q K^T / sqrt(hd)) V   per (batch, head), no mask.
// Reference semantics: diffusers 0.25.0 `Attention` + AttnProcessor2_0 (F.scaled_dot_product_attention) as used for
// attn1 / attn2 of BasicTransformerBlock (live2diff/animatediff/models/attention.py:173-194, 243, 251-253); heads = 8,
// hd = 40 / 80 at the two levels where S is large (64x64 and 32x32 latents: S = 4096 / 1024).
//
// One CTA = one or two 128-query tiles of one (batch, head); it walks the keys in tiles of 128:
//   warp 0     TMA producer: Q once, then K_j / V_j tiles ([128 rows x 64 columns] boxes, 128B swizzle) into two
//              independent rings (K is consumed one tile ahead of V)
//   warp 1     TMEM allocator + tcgen05.mma issuer.  The whole warp walks the issue loop and ONE ELECTED LANE executes the
//              MMAs and commits, so that descriptor arithmetic stays in uniform registers (inside an `if (lane == 0)` region
//              the compiler wraps every UTCHMMA in a per-lane waterfall loop, ~150 cycles per MMA against a floor of 32-64):
//                S_j = Q . K_j^T   (M = 128 queries, N = 128 keys, K = hd padded to 16s; both operands K-major in shared memory)
//                O  += P_j . V_j   (M = 128, N = hd rounded up to whole 64-channel swizzle atoms, K = 128 keys; A = P read
//                                  from TENSOR MEMORY, B = the V tile exactly as TMA delivered it = MN-major, 128B swizzle);
//                                  O stays in TMEM across ALL key tiles (accumulate flag)
//   warps 2-5  softmax + epilogue (a second quartet, warps 6-9, when the CTA carries two query tiles): thread = one query
//              row (tcgen05.ld 32x32b: TMEM lane = row), so row max / row sum need no shuffles.  A tile's 128 scores are pulled
//              into registers in one go (S is read from TMEM once and its accumulator goes back to the tensor core at the
//              START of the tile); the exp uses the row max of the EARLIER tiles as its reference and O / l are rescaled
//              only when a row outgrows the reference by more than 2^8 (lazy rescale); P is rounded to fp16 (like the fused
//              SDPA kernels the reference dispatches to) and written back to TMEM with tcgen05.st as fp16 pairs -- it never
//              touches shared memory.  The 1/l normalisation is applied once at the end.  (Measured and not used, see
//              profiles/README.md: a polynomial exp2 on the FMA pipe for one exponential in four; strict alternation of
//              the two quartets' exp phases through a named-barrier token, kept behind L2D_FLASH_PINGPONG=1.)
// hd = 40: the K extent of Q.K^T is padded to 48 by zeroing columns 40..47 of the Q tile in shared memory (the K tile's
// columns 40..47 then hold the next head's values, finite, times zero); O's columns 40..63 are never read.
// Keys beyond skv in the last tile (cross-attention: 77 keys) are masked to -inf before the softmax; their V rows are
// other rows of the same tensor or TMA zero fill, always finite.
#include <cuda.h>

#include <cstdio>
#include <type_traits>

#include "ops.cuh"

namespace l2d {

int get_tmap_2d(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out);   // gemm_tcgen05.cu

namespace {

__device__ __forceinline__ uint32_t ft_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ft_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ft_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ft_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ft_mbar_wait(uint32_t bar, uint32_t parity) {
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
    if ((++spins & 1023u) == 0) {   // never hang the GPU on a protocol bug: trap after ~2 s
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();
    }
  }
}
__device__ __forceinline__ void ft_tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void ft_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void ft_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ft_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ft_umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from TMEM (lane = row, one 32-bit column = two consecutive K elements), B from shared memory
__device__ __forceinline__ void ft_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of a converged warp (always the lowest: tcgen05.commit tracks the MMAs of the thread that executes it)
__device__ __forceinline__ bool ft_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void ft_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ft_tmem_ld32(uint32_t addr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr));
}
__device__ __forceinline__ void ft_tmem_ld16(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
}
__device__ __forceinline__ void ft_tmem_ld8(uint32_t addr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr));
}
__device__ __forceinline__ float ft_ex2(float x) {   // arguments <= 0: no overflow; tiny results flush to zero
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ft_tmem_st16(uint32_t addr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void ft_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptors (cute::UMMA::SmemDescriptor bit layout, version 1, SWIZZLE_128B):
//   K-major  (Q, K, P): rows of 128 B (64 fp16 along K), 8-row groups 1024 B apart (SBO); a K step of 16 = +32 B
//   MN-major (V as loaded: row = key = K index, 64 consecutive head channels = MN): canonical ((8,n),(8,k)):((1,LBO),(8,SBO))
//            in 16-byte units: 8-key groups 1024 B apart (SBO), 64-channel column blocks `lbo` bytes apart (LBO)
__device__ __forceinline__ uint64_t ft_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format F16 (0) @7/@10, a_major @15, b_major @16 (0 = K, 1 = MN),
// N >> 3 @17, M >> 4 @24
__host__ __device__ constexpr uint32_t ft_idesc(int n, int b_mn_major) {
  return (1u << 4) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

long long* g_ft_dbg = nullptr;   // l2d_flash_set_debug

struct FtParams {
  __half* o;
  int64_t ldo;
  int sq, skv, heads;
  float scale_log2;   // log2(e) / sqrt(hd)
  int pingpong;       // NQ == 2: the two query tiles take turns on the MUFU pipe (L2D_FLASH_PINGPONG=1; default 0: measured slower)
  long long* dbg;     // developer timeline of CTA (0,0,0), see l2d_flash_set_debug; nullptr = off
};

constexpr int FT_TILE = 128 * 128;   // bytes of one [128 rows x 64 fp16] swizzled block

// NQ = 128-query tiles per CTA.  The driver keeps kernels that use tcgen05 at ONE CTA per SM whatever their shared-memory
// and TMEM footprint (measured: cudaOccupancyMaxActiveBlocksPerMultiprocessor = 1 even at 48 KB / 256 columns), so the
// latency hiding that two co-resident CTAs would give is built into one: with NQ = 2 two query tiles share every K/V tile
// and run their own softmax warp quartets; while one tile's chain (S ready -> scores to registers -> exp -> P ready -> P.V)
// waits on the tensor core or a barrier, the other's exp pass keeps the MUFU pipe busy (one tile per CTA: 179 vs 112 us at
// the level-0 shape).  hd = 80 keeps NQ = 1: two query tiles would need 2 x 128 TMEM columns for O on top of S (256) and P.
template <int HD, int NQ>
struct FtCfg {
  static constexpr int KSTEPS = (HD + 15) / 16;     // k16 steps of Q.K^T
  static constexpr int NBLK = (HD + 63) / 64;       // 64-column blocks of a Q / K / V tile
  static constexpr int ON = KSTEPS * 16;            // columns of T = P.V that are read back (hd rounded up to 16)
  static constexpr int ON_MMA = NBLK * 64;          // MMA N of P.V: whole 64-channel swizzle atoms of the V tile
  static constexpr int STAGES = 2;                  // K / V ring depth
  static constexpr int SBUF = 2 / NQ;               // S accumulators per query tile (NQ * SBUF = 2 x 128 TMEM columns)
  static constexpr int TMEM_COLS = 512;             // S: 256 | O: NQ x ON_MMA <= 128 | P (fp16 pairs): NQ x 64
  static constexpr int THREADS = 64 + NQ * 128;     // TMA warp, MMA warp, 4 softmax warps per query tile
  static constexpr int TILE_BYTES = NBLK * FT_TILE;
  static constexpr int SMEM = TILE_BYTES * (NQ + 2 * STAGES) + 256 /* barriers */ + 1024 /* align */;
  static_assert(NQ * ON_MMA <= 128 && (NQ == 1 || NQ == 2), "TMEM layout");
};

template <int HD, int NQ>
__global__ void __launch_bounds__(FtCfg<HD, NQ>::THREADS, 1)
flash_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const FtParams p) {
  using Cfg = FtCfg<HD, NQ>;
  constexpr int ST = Cfg::STAGES, NBLK = Cfg::NBLK, KSTEPS = Cfg::KSTEPS, ON = Cfg::ON, SBUF = Cfg::SBUF;
  extern __shared__ __align__(1024) uint8_t ft_smem_raw[];
  uint8_t* smem = ft_smem_raw + ((1024u - (ft_smem_u32(ft_smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                  // [NQ] query tiles
  uint8_t* sK = sQ + NQ * Cfg::TILE_BYTES;
  uint8_t* sV = sK + ST * Cfg::TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ST * Cfg::TILE_BYTES);
  const uint32_t bar0 = ft_smem_u32(bars);
  // barrier map: q_full, q_ready | per query tile t: p_full, o_full, (unused), s_full[2], s_empty[2] | k/v full/empty[ST]
  const uint32_t q_full = bar0, q_ready = bar0 + 8;
  auto p_full = [&](int t) { return bar0 + 16 + 56u * t; };
  auto o_full = [&](int t) { return bar0 + 24 + 56u * t; };
  auto s_full = [&](int t, int a) { return bar0 + 40 + 56u * t + 8u * a; };
  auto s_empty = [&](int t, int a) { return bar0 + 56 + 56u * t + 8u * a; };
  const uint32_t kv0 = bar0 + 16 + 56u * NQ;
  auto k_full = [&](int s) { return kv0 + 8u * s; };
  auto k_empty = [&](int s) { return kv0 + 8u * (ST + s); };
  auto v_full = [&](int s) { return kv0 + 8u * (2 * ST + s); };
  auto v_empty = [&](int s) { return kv0 + 8u * (3 * ST + s); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 + 7 * NQ + 4 * ST);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, m0 = blockIdx.x * (128 * NQ);
  const int nk = (p.skv + 127) >> 7;
  const int col0 = h * HD;

  pdl_launch();
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_v) : "memory");
    ft_mbar_init(q_full, 1);
    ft_mbar_init(q_ready, 4 * NQ);
    for (int t = 0; t < NQ; ++t) {
      ft_mbar_init(p_full(t), 4);
      ft_mbar_init(o_full(t), 1);
      for (int a = 0; a < 2; ++a) {
        ft_mbar_init(s_full(t, a), 1);
        ft_mbar_init(s_empty(t, a), 4);
      }
    }
    for (int s = 0; s < ST; ++s) {
      ft_mbar_init(k_full(s), 1);
      ft_mbar_init(k_empty(s), 1);
      ft_mbar_init(v_full(s), 1);
      ft_mbar_init(v_empty(s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ft_smem_u32(tmem_ptr_smem)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  ft_tc_fence_before();
  __syncthreads();
  ft_tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // TMEM columns: S accumulator (tile t, buffer a) at (t * SBUF + a) * 128; T of tile t at 256 + t * ON_MMA
  auto tm_s = [&](int t, int a) { return (uint32_t)((t * SBUF + a) * 128); };
  auto tm_o = [&](int t) { return (uint32_t)(256 + t * Cfg::ON_MMA); };
  auto tm_p = [&](int t) { return (uint32_t)(384 + t * 64); };   // P_j as fp16 pairs: 128 keys = 64 columns

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      pdl_wait();   // q / k / v were written by the projection GEMMs right before this kernel
      ft_mbar_expect_tx(q_full, NQ * Cfg::TILE_BYTES);
      for (int t = 0; t < NQ; ++t)
        for (int cb = 0; cb < NBLK; ++cb)
          ft_tma_2d(ft_smem_u32(sQ + t * Cfg::TILE_BYTES + cb * FT_TILE), &tmap_q, q_full, col0 + cb * 64, b * p.sq + m0 + t * 128);
      auto load_k = [&](int j) {
        const int s = j % ST;
        ft_mbar_wait(k_empty(s), ((uint32_t)(j / ST) & 1u) ^ 1u);
        ft_mbar_expect_tx(k_full(s), Cfg::TILE_BYTES);
        for (int cb = 0; cb < NBLK; ++cb)
          ft_tma_2d(ft_smem_u32(sK + s * Cfg::TILE_BYTES + cb * FT_TILE), &tmap_k, k_full(s), col0 + cb * 64, b * p.skv + j * 128);
      };
      // K runs one tile ahead of V: S_{j+1} is issued while the softmax of tile j is still busy, V_j only feeds P_j . V_j
      load_k(0);
      for (int j = 0; j < nk; ++j) {
        if (j + 1 < nk) load_k(j + 1);
        const int sv = j % ST;
        ft_mbar_wait(v_empty(sv), ((uint32_t)(j / ST) & 1u) ^ 1u);
        ft_mbar_expect_tx(v_full(sv), Cfg::TILE_BYTES);
        for (int cb = 0; cb < NBLK; ++cb)
          ft_tma_2d(ft_smem_u32(sV + sv * Cfg::TILE_BYTES + cb * FT_TILE), &tmap_v, v_full(sv), col0 + cb * 64, b * p.skv + j * 128);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The WHOLE warp walks the loop (convergent control flow) and one elected lane executes the tcgen05 instructions, so the
    // descriptor arithmetic stays in uniform registers; inside an `if (lane == 0)` region the compiler wraps every UTCHMMA
    // in a per-lane waterfall loop, and the timeline showed this thread -- ~150 cycles per MMA against a tensor-core floor
    // of 32-64 -- pacing the whole kernel.
    {
      long long* mdbg = (p.dbg && lane == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? p.dbg + 8 * 16 * 8 : nullptr;
      constexpr uint32_t idesc_s = ft_idesc(128, 0);
      constexpr uint32_t idesc_o = ft_idesc(Cfg::ON_MMA, 1);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      auto issue_s = [&](int t, int j) {
        const int s = j % ST, a = j % SBUF;
        if (t == 0) ft_mbar_wait(k_full(s), (uint32_t)(j / ST) & 1u);
        ft_mbar_wait(s_empty(t, a), ((uint32_t)(j / SBUF) & 1u) ^ 1u);   // the softmax has drained this S accumulator (tile j-SBUF)
        ft_tc_fence_after();
        const uint32_t qa = ft_smem_u32(sQ + t * Cfg::TILE_BYTES), ka = ft_smem_u32(sK + s * Cfg::TILE_BYTES);
        if (ft_elect_one()) {
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k) {
            const uint32_t off = (uint32_t)(k >> 2) * FT_TILE + (uint32_t)(k & 3) * 32;
            ft_umma(tmem_u + tm_s(t, a), ft_desc(qa + off, 0), ft_desc(ka + off, 0), idesc_s, k != 0);
          }
          if (t == NQ - 1) ft_commit(k_empty(s));   // every query tile has read K_j
          ft_commit(s_full(t, a));
        }
        __syncwarp();
        if (mdbg && j < 16) mdbg[(t * 16 + j) * 2] = clock64();       // S_j issued
      };
      auto issue_pv = [&](int t, int j) {
        const int s = j % ST;
        if (t == 0) ft_mbar_wait(v_full(s), (uint32_t)(j / ST) & 1u);
        ft_mbar_wait(p_full(t), (uint32_t)j & 1u);
        ft_tc_fence_after();
        const uint32_t va = ft_smem_u32(sV + s * Cfg::TILE_BYTES);
        if (ft_elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {   // 128 keys = 8 k16 steps: P +8 TMEM columns per step; V: 16 key rows = 2 KB per step
            const uint64_t db = ft_desc(va + (uint32_t)kk * 2048, FT_TILE);
            ft_umma_ts(tmem_u + tm_o(t), tmem_u + tm_p(t) + (uint32_t)kk * 8, db, idesc_o, (j | kk) != 0);   // O accumulates across the key tiles
          }
          if (t == NQ - 1) ft_commit(v_empty(s));   // every query tile has read V_j
          ft_commit(o_full(t));
        }
        __syncwarp();
        if (mdbg && j < 16) mdbg[(t * 16 + j) * 2 + 1] = clock64();   // P_j . V_j issued
      };
      ft_mbar_wait(q_ready, 0);
      for (int t = 0; t < NQ; ++t) issue_s(t, 0);
      for (int j = 0; j < nk; ++j) {
        for (int t = 0; t < NQ; ++t) {
          // S_{j+1} first: its accumulator was handed back when the softmax pulled S_j into registers, P_j comes later
          if (j + 1 < nk) issue_s(t, j + 1);
          issue_pv(t, j);
        }
      }
    }
  } else {
    // ===================== softmax / correction / epilogue (4 warps per query tile) =====================
    const int t = (warp - 2) >> 2;             // query tile of this warp quartet
    const int q = warp & 3;                    // TMEM lane quarter this warp may touch
    const int r = q * 32 + lane;               // query row inside the tile
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* myQ = sQ + t * Cfg::TILE_BYTES;
    ft_mbar_wait(q_full, 0);
    if (HD % 16 != 0) {   // zero columns HD .. KSTEPS*16-1 of this row (hd = 40: chunk 5 of block 0)
      constexpr int chunk = HD / 8;
      *reinterpret_cast<uint4*>(myQ + (chunk >> 3) * FT_TILE + r * 128 + (((chunk & 7) ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
    }
    ft_fence_proxy_async();
    __syncwarp();
    if (lane == 0) ft_mbar_arrive(q_ready);

    // Online softmax with a LAZY reference (flash-attention 4's conditional rescale).  A tile's scores (128 fp32 per row) are
    // pulled from TMEM into registers in one go -- S is read exactly once and its accumulator goes back to the tensor core
    // at the START of the tile, so S_{j+1} = Q.K^T runs under the exp pass of tile j instead of after it (the ncu stall
    // profile of the version that kept S in TMEM until the end of the pass had the softmax warps waiting on s_full for a
    // quarter of all samples).  The exp uses the reference m_ref the row already has (the max of the earlier tiles) unless
    // some row of the warp exceeds its reference by more than 2^LAZY; only then are O (in TMEM, accumulated by the tensor
    // core across all key tiles) and l rescaled to the new max.  P <= 2^8 fits fp16 with the same relative precision; l and
    // O are fp32.
    float m_ref = -INFINITY, l_run = 0.f;
    const float c = p.scale_log2;
    constexpr float LAZY = 8.f;
    const uint32_t o_addr = t_lane + tm_o(t);
    // NQ == 2, optional (L2D_FLASH_PINGPONG=1): a token passed through two named barriers between warp w of tile 0 and warp
    // w + 4 of tile 1 (same scheduler) makes the two tiles' exp phases alternate strictly (flash-attention 3's ping-pong).
    // Measured 114 vs 111 us without it at the level-0 shape: the exp phase is bound by one warp's dependent instruction
    // stream, not by MUFU throughput, so two overlapping exp phases are faster than two serialised ones.
    const int bar_mine = 1 + t * 4 + (warp & 3), bar_other = 1 + (t ^ 1) * 4 + (warp & 3);
    const bool pp = NQ == 2 && p.pingpong;
    if (pp && t == 1) asm volatile("bar.arrive %0, 64;" ::"r"(bar_other) : "memory");   // tile 0 goes first

    long long* dbg = (p.dbg && lane == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? p.dbg + (warp - 2) * 16 * 8 : nullptr;
    for (int j = 0; j < nk; ++j) {
      const int a = j % SBUF;
      const int nvalid = min(128, p.skv - j * 128);
      long long* stamp = (dbg && j < 16) ? dbg + j * 8 : nullptr;
      if (stamp) stamp[0] = clock64();   // tile start
      ft_mbar_wait(s_full(t, a), (uint32_t)(j / SBUF) & 1u);
      ft_tc_fence_after();
      if (stamp) stamp[1] = clock64();   // S_j available
      const uint32_t s_addr = t_lane + tm_s(t, a);
      uint32_t v0[32], v1[32], v2[32], v3[32];
      ft_tmem_ld32(s_addr, v0);
      ft_tmem_ld32(s_addr + 32, v1);
      ft_tmem_ld32(s_addr + 64, v2);
      ft_tmem_ld32(s_addr + 96, v3);
      ft_tmem_ld_wait();
      if (stamp) stamp[2] = clock64();   // scores in registers
      // the scores are in registers: hand the accumulator back (tile j + SBUF may overwrite it)
      ft_tc_fence_before();
      __syncwarp();
      if (lane == 0) ft_mbar_arrive(s_empty(t, a));
      if (nvalid < 128) {   // last tile of a ragged key length (cross-attention: 77 keys): masked scores never win the max, exp -> 0
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          if (e >= nvalid) v0[e] = 0xff800000u;
          if (32 + e >= nvalid) v1[e] = 0xff800000u;
          if (64 + e >= nvalid) v2[e] = 0xff800000u;
          if (96 + e >= nvalid) v3[e] = 0xff800000u;
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        mx = fmaxf(mx, fmaxf(__uint_as_float(v0[e]), __uint_as_float(v0[e + 1])));
        mx = fmaxf(mx, fmaxf(__uint_as_float(v1[e]), __uint_as_float(v1[e + 1])));
        mx = fmaxf(mx, fmaxf(__uint_as_float(v2[e]), __uint_as_float(v2[e + 1])));
        mx = fmaxf(mx, fmaxf(__uint_as_float(v3[e]), __uint_as_float(v3[e + 1])));
      }
      if (j > 0) {   // P_{j-1} . V_{j-1} has completed: P may be overwritten and O is stable
        ft_mbar_wait(o_full(t), (uint32_t)(j - 1) & 1u);
        ft_tc_fence_after();
      }
      if (j == 0) {
        m_ref = mx;   // no history yet: the first tile's own max is the reference
      } else if (__any_sync(0xffffffffu, fmaf(mx, c, -m_ref * c) > LAZY)) {
        // some row of this warp outgrew its reference: move every row of the warp to its new max and rescale the history
        // (rows whose max did not move get the factor 2^0 = 1, exact)
        const float m_new = fmaxf(m_ref, mx);
        const float corr = ft_ex2((m_ref - m_new) * c);
#pragma unroll
        for (int cc = 0; cc < ON / 16; ++cc) {
          uint32_t tt[16];
          ft_tmem_ld16(o_addr + cc * 16, tt);
          ft_tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) tt[e] = __float_as_uint(__uint_as_float(tt[e]) * corr);
          ft_tmem_st16(o_addr + cc * 16, tt);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        ft_tc_fence_before();
        l_run *= corr;
        m_ref = m_new;
      }
      const float mneg = -m_ref * c;
      float rs = 0.f;
      const uint32_t p_addr = t_lane + tm_p(t);
      // p = 2^(s*c - m_ref*c), row sum in fp32, P -> fp16 pairs written to TMEM (tcgen05.st: lane = row, column = key pair),
      // where the tensor core reads it as the A operand of P.V -- P never touches shared memory
      auto emit = [&](const uint32_t (&v)[32], int cc) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float p0 = ft_ex2(fmaf(__uint_as_float(v[e]), c, mneg));
          const float p1 = ft_ex2(fmaf(__uint_as_float(v[e + 1]), c, mneg));
          rs += p0 + p1;
          pk[e >> 1] = h2_as_u32(__floats2half2_rn(p0, p1));
        }
        ft_tmem_st16(p_addr + cc * 16, pk);
      };
      if (stamp) stamp[3] = clock64();   // max / lazy check / P.V_{j-1} wait done
      if (pp) asm volatile("bar.sync %0, 64;" ::"r"(bar_mine) : "memory");     // my turn on the MUFU pipe
      if (stamp) stamp[4] = clock64();   // token acquired
      emit(v0, 0);
      emit(v1, 1);
      emit(v2, 2);
      emit(v3, 3);
      if (stamp) stamp[5] = clock64();   // exponentials + P stores issued
      if (pp) asm volatile("bar.arrive %0, 64;" ::"r"(bar_other) : "memory");  // the other tile's turn
      l_run += rs;
      // P_j written: hand it to the MMA warp
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      ft_tc_fence_before();
      __syncwarp();
      if (lane == 0) ft_mbar_arrive(p_full(t));
      if (stamp) stamp[6] = clock64();   // P_j handed over
    }
    if (pp && t == 0) asm volatile("bar.sync %0, 64;" ::"r"(bar_mine) : "memory");   // absorb tile 1's last hand-over
    // ---- last tile's P.V has landed: normalise, store ----
    ft_mbar_wait(o_full(t), (uint32_t)(nk - 1) & 1u);
    ft_tc_fence_after();
    const float inv = 1.f / l_run;
    const int row = m0 + t * 128 + r;
    __half* orow = p.o + ((size_t)b * p.sq + row) * p.ldo + col0;
    constexpr int OUT16 = (HD + 15) / 16;
#pragma unroll
    for (int cc = 0; cc < OUT16; ++cc) {
      uint32_t tt[16];
      ft_tmem_ld16(o_addr + cc * 16, tt);
      ft_tmem_ld_wait();
      if (row < p.sq) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (cc * 16 + i * 8 < HD) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(tt[i * 8 + e]) * inv;
            *reinterpret_cast<uint4*>(orow + cc * 16 + i * 8) = pack8(f);
          }
        }
      }
    }
  }
  ft_tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ft_tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}

template <int HD, int NQ>
int ft_launch(const __half* q, int64_t ldq, const __half* k, int64_t ldk, const __half* v, int64_t ldv, __half* o, int64_t ldo,
              int batch, int heads, int sq, int skv, cudaStream_t st) {
  using Cfg = FtCfg<HD, NQ>;
  static bool configured = false;
  if (!configured) {
    L2D_CUDA(cudaFuncSetAttribute(flash_tcgen05_kernel<HD, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    configured = true;
  }
  CUtensorMap tq, tk, tv;
  int rc = get_tmap_2d(q, (int64_t)batch * sq, (int64_t)heads * HD, ldq, 128, &tq);
  if (rc != L2D_OK) return rc;
  rc = get_tmap_2d(k, (int64_t)batch * skv, (int64_t)heads * HD, ldk, 128, &tk);
  if (rc != L2D_OK) return rc;
  rc = get_tmap_2d(v, (int64_t)batch * skv, (int64_t)heads * HD, ldv, 128, &tv);
  if (rc != L2D_OK) return rc;
  static const int pingpong = [] { const char* e = getenv("L2D_FLASH_PINGPONG"); return e ? atoi(e) : 0; }();
  FtParams p{o, ldo, sq, skv, heads, 1.4426950408889634f / sqrtf((float)HD), pingpong, g_ft_dbg};
  launch_pdl_if(pdl_family(2), flash_tcgen05_kernel<HD, NQ>, dim3(sq / (128 * NQ), heads, batch), dim3(Cfg::THREADS),
                (size_t)Cfg::SMEM, st, tq, tk, tv, p);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

}  // namespace

void set_flash_debug(long long* ptr) { g_ft_dbg = ptr; }

int flash_tcgen05_ctas_per_sm(int hd) {
  int n = 0;
  if (hd == 40) {
    cudaFuncSetAttribute(flash_tcgen05_kernel<40, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FtCfg<40, 2>::SMEM);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, flash_tcgen05_kernel<40, 2>, FtCfg<40, 2>::THREADS, FtCfg<40, 2>::SMEM);
    if (getenv("L2D_FLASH_OCC_DEBUG")) {   // developer print: occupancy as a function of the dynamic shared-memory request
      fprintf(stderr, "[flash occupancy] hd 40, NQ 2 (%d threads):", FtCfg<40, 2>::THREADS);
      for (int kb : {48, 96, 112, 160}) {
        int m = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m, flash_tcgen05_kernel<40, 2>, FtCfg<40, 2>::THREADS, (size_t)kb * 1024);
        fprintf(stderr, " %dKB:%d", kb, m);
      }
      fprintf(stderr, "\n");
    }
  } else if (hd == 80) {
    cudaFuncSetAttribute(flash_tcgen05_kernel<80, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FtCfg<80, 1>::SMEM);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, flash_tcgen05_kernel<80, 1>, FtCfg<80, 1>::THREADS, FtCfg<80, 1>::SMEM);
  }
  return n;
}

// hd 40 / 80, whole 128-query tiles, 16-byte aligned operands with row pitches that are multiples of 8 elements
bool attention_tcgen05_supported(const void* q, const void* k, const void* v, const void* o, int64_t ldq, int64_t ldk, int64_t ldv,
                                 int64_t ldo, int sq, int skv, int hd) {
  static const bool off = [] {
    const char* e = getenv("L2D_FLASH_LEGACY");
    return e && e[0] == '1';
  }();
  if (off) return false;
  auto al = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; };
  return (hd == 40 || hd == 80) && sq % 128 == 0 && skv >= 1 && al(q) && al(k) && al(v) && al(o) && ldq % 8 == 0 &&
         ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0;
}

int attention_tcgen05_launch(const __half* q, int64_t ldq, const __half* k, int64_t ldk, const __half* v, int64_t ldv, __half* o,
                             int64_t ldo, int batch, int heads, int sq, int skv, int hd, cudaStream_t st) {
  static const int force_nq1 = [] { const char* e = getenv("L2D_FLASH_NQ"); return e && atoi(e) == 1; }();   // developer A/B
  if (hd == 40 && sq % 256 == 0 && !force_nq1) return ft_launch<40, 2>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, sq, skv, st);
  if (hd == 40) return ft_launch<40, 1>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, sq, skv, st);
  if (hd == 80) return ft_launch<80, 1>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, sq, skv, st);
  return fail(L2D_ERR_INVALID, "attention(tcgen05): unsupported head_dim");
}

}  // namespace l2d

// developer hook (include/l2d_b200_debug.h): resident CTAs per SM of the tcgen05 attention kernel for head_dim 40 / 80
extern "C" int l2d_debug_flash_ctas_per_sm(int hd) { return l2d::flash_tcgen05_ctas_per_sm(hd); }
// developer hook: per-tile clock64 stamps of CTA (0,0,0) of the tcgen05 attention kernel (profiles/flash_timeline.py)
extern "C" void l2d_flash_set_debug(void* timeline) { l2d::set_flash_debug(static_cast<long long*>(timeline)); }
