// K7 (and, through im2col, K4) -- fp16 GEMM on the 5th-gen tensor cores:
//     out[M,N] = epilogue( A[M,K] . W[N,K]^T ),  fp32 accumulation in TMEM.
//
// Covers every dense contraction of the UNet step: nn.Linear (to_q/k/v, to_out, proj_in/out, GEGLU
// FF -- attention.py:173-205, motion_module.py:182-199,360; stream_motion_module.py:105-107,206),
// 1x1 convs (attention.py:62,87; resnet.py:227) and, fed by the im2col kernels, the 3x3 convs of
// resnet.py:57-65,92,141,194,214.
//
// Structure (one 128 x BN output tile per CTA, 320 threads):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor 2D loads of A[128x64] and W[BNx64] tiles
//              (128B swizzle) into a STAGES-deep shared-memory ring, mbarrier complete_tx
//   warp 1   : TMEM allocator + MMA issuer -- the converged warp walks the k-loop, one elected lane issues
//              tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) from shared-memory descriptors (kept in
//              uniform registers); tcgen05.commit releases ring slots and finally signals the accumulator-ready barrier
//   warps 2-9: epilogue -- tcgen05.ld 32x32b TMEM -> registers, fused bias / per-image bias (temb) /
//              SiLU / ReLU / GEGLU / residual, 256-bit global loads and stores
// Partial tiles rely on TMA out-of-bounds zero fill (loads) and predicated stores.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "ops.cuh"

namespace l2d {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 fp16 = 128 B = one swizzle-128B row

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
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
    // never hang the GPU: a protocol bug becomes a launch failure after ~2 s
    if ((++spins & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of a converged warp (the lowest: the same lane every time, as tcgen05.commit tracks the issuing thread's MMAs)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one request per 32-byte sector instead of two.  The epilogue's
// accesses are one row per lane (row pitch >= 640 B), so every 16-byte access was its own L2 request; measured +2.2 % frames/s
// in a same-box A/B (profiles/README.md, "GEMM family").  Hoisting all of a tile's bias / residual loads ahead of the
// accumulator wait (bias staged in shared memory) shortened the isolated epilogue by a third but made the frame 4 % slower
// (registers 126 -> 168, one more barrier per tile) and is not used.
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p)
               : "memory");
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
               "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

// K-major, 128B-swizzled operand tile descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major) | [32,46) SBO >> 4 = 1024 B
//   (8 rows x 128 B) | [46,48) version = 1 | [61,64) layout = SWIZZLE_128B (2)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format F16 (0) @7/@10, a/b K-major (0) @15/@16,
// N >> 3 @17, M >> 4 @24
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// Implicit-GEMM 3x3 convolution (pad 1, stride 1) over a channels-last activation [N,H,W,Cin]: the A tile of
// output pixels (bn images x bh rows x bw columns = 128) for tap (dy,dx) and channel block c0 is ONE 4-D TMA box
// at (c0, w0+dx-1, h0+dy-1, n0); out-of-range coordinates are zero-filled by the TMA unit = the conv's padding.
// No im2col matrix exists.  k-block kb <-> (tap = kb / cin_blocks, c0 = 64 * (kb % cin_blocks)), matching the
// [Cout, 9*Cin] (tap, cin) weight repack.
struct ConvGeom {
  int enabled;
  int cin_blocks;        // Cin / 64
  int H, W, N;
  int bw, bh, bn;        // tile extents, bw * bh * bn == 128
  int tiles_w, tiles_h;  // tiles per image row / column
};

struct GemmEpilogue {
  __half* out;
  int64_t ldo;
  const __half* bias;           // [N] or null
  const __half* rowgroup_bias;  // [M / rows_per_group, >=N] or null
  int64_t rg_ld;
  int rows_per_group;
  const __half* residual;       // [M, ldr] or null
  int64_t ldr;
  int act;
  // split-K (gridDim.z > 1): every split stores its fp32 partial tile to ws[z][M_pad][ws_ld]; a second
  // kernel (splitk_finish_kernel) sums the slices and applies the epilogue
  float* ws;
  int64_t ws_ld;
  int64_t ws_slice;   // elements per split slice = M_pad * ws_ld
  long long* dbg;     // optional per-CTA timeline (8 clock64 stamps per CTA), see l2d_gemm_set_debug
  // LayerNorm fusion (ops.cuh GemmFusion)
  float2* stats_out;        // producer: [M][stats_slots] (sum, sumsq) of the stored fp16 row values
  int stats_slots;
  const float2* ln_stats;   // consumer: row statistics of A
  int ln_slots;
  const float* ln_s;
  const float* ln_b;
  float ln_inv_c, ln_eps;
};

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Persistent kernel: grid = min(#tiles, #SMs); every CTA walks tiles t = blockIdx.x, +gridDim.x, ... (N-tile fastest,
// then M-tile, then K-split).  The smem operand ring runs continuously across tiles and the accumulator is
// double-buffered in TMEM, so tile i's epilogue overlaps tile i+1's TMA loads and MMAs.
// CL = true: split-K inside a thread-block cluster.  The `splits` CTAs of one output tile form a cluster (blockIdx.x =
// tile * splits + z); each runs its K range into TMEM, then scatters its fp32 partial tile over the cluster through
// distributed shared memory -- CTA z' receives, from every CTA, the column slice [z' * BN/splits, +BN/splits) -- and after
// one cluster barrier reduces its slice in CTA-rank order (deterministic) and applies the epilogue.  No fp32 workspace in
// HBM/L2 and no second kernel (the non-cluster path below keeps the slice buffer + splitk_finish_kernel).
// NACC: accumulator stages in TMEM (2 = tile i's epilogue overlaps tile i+1's MMAs).  ("Lean" instantiations -- half the
// operand ring, <= 256 TMEM columns, so that the next kernel's CTAs could become resident beside a running GEMM CTA and
// overlap their prologue under PDL -- were measured 8 % slower per frame and removed: profiles/README.md.)
template <int BN, int STAGES, bool CL = false, int NACC = 2>
__global__ void __launch_bounds__(320, 1)
gemm_f16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                        const GemmEpilogue epi, const ConvGeom cg, int M, int N, int K, int tiles_n, int tiles_m,
                        int splits) {
  using S = GemmSmem<BN>;
  constexpr uint32_t TMEM_COLS = NACC * BN <= 64 ? 64 : NACC * BN <= 128 ? 128 : NACC * BN <= 256 ? 256 : 512;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024 B alignment is required by the 128B swizzle atom (8 rows x 128 B)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * S::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full_bar = bars + 2 * STAGES;        // [2]
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;   // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_kb = (K + BK - 1) / BK;
  const int kb_per = (total_kb + splits - 1) / splits;
  const int total_tiles = tiles_n * tiles_m * splits;
  const int my_tiles = CL ? 1 : ((int)blockIdx.x < total_tiles) ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  float* cl_part = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES + 256);   // CL: [splits][128 rows][BN / splits] fp32

  struct Tile {
    int n0, m0, mt, z, kb_begin, num_kb, cn0, ch0, cw0;
  };
  auto decode = [&](int i) {
    const int t = CL ? (int)blockIdx.x / splits + ((int)blockIdx.x % splits) * (tiles_n * tiles_m)
                     : (int)blockIdx.x + i * (int)gridDim.x;
    Tile tl;
    tl.z = t / (tiles_n * tiles_m);
    const int rem = t - tl.z * (tiles_n * tiles_m);
    tl.mt = rem / tiles_n;
    tl.n0 = (rem - tl.mt * tiles_n) * BN;
    tl.m0 = tl.mt * BM;
    tl.kb_begin = tl.z * kb_per;
    tl.num_kb = max(0, min(total_kb, tl.kb_begin + kb_per) - tl.kb_begin);
    tl.cn0 = tl.ch0 = tl.cw0 = 0;
    if (cg.enabled) {   // output tile = images [cn0, cn0+bn) x rows [ch0, ch0+bh) x columns [cw0, cw0+bw)
      tl.cw0 = (tl.mt % cg.tiles_w) * cg.bw;
      tl.ch0 = ((tl.mt / cg.tiles_w) % cg.tiles_h) * cg.bh;
      tl.cn0 = (tl.mt / (cg.tiles_w * cg.tiles_h)) * cg.bn;
    }
    return tl;
  };

  long long* dbg = epi.dbg ? epi.dbg + (size_t)blockIdx.x * 8 : nullptr;   // timeline of this CTA's first tile
  pdl_launch();   // the setup below touches no global memory: it overlaps the previous kernel's tail (common.cuh)
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a2 = 0; a2 < 2; ++a2) {
      mbar_init(smem_u32(&tmem_full_bar[a2]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[a2]), 8);   // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (dbg && threadIdx.x == 0) {
    dbg[0] = dbg[1] = clock64();   // (setup precedes the PDL wait and is not part of the timeline)
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // PDL: the B operand is a weight matrix (constant for the whole step), so the weight tiles of the first ring slots
      // are requested while the previous kernel is still draining; the A tiles (activations written by that kernel)
      // follow after the wait.  One expect_tx per slot covers both operands.
      int pre = 0;
      if (my_tiles > 0) {
        const Tile t0 = decode(0);
        pre = t0.num_kb < STAGES ? t0.num_kb : STAGES;
        for (int kb = 0; kb < pre; ++kb) {
          const uint32_t fb = smem_u32(&full_bar[kb]);
          mbar_expect_tx(fb, S::STAGE_BYTES);
          tma_load_2d(smem_u32(smem_b + kb * S::B_BYTES), &tmap_b, fb, (t0.kb_begin + kb) * BK, t0.n0);
        }
      }
      pdl_wait();
      int it = 0;   // running k-block counter: the operand ring never drains between tiles
      for (int i = 0; i < my_tiles; ++i) {
        const Tile tl = decode(i);
        for (int kb = 0; kb < tl.num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          const bool prefetched = it < pre;               // slot `it` of tile 0: expect_tx + B already issued above
          if (!prefetched) mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[s]);
          if (!prefetched) mbar_expect_tx(fb, S::STAGE_BYTES);
          const int kk = tl.kb_begin + kb;
          if (cg.enabled) {
            const int tap = kk / cg.cin_blocks, cb = kk - tap * cg.cin_blocks;
            tma_load_4d(smem_u32(smem_a + s * S::A_BYTES), &tmap_a, fb, cb * BK, tl.cw0 + tap % 3 - 1, tl.ch0 + tap / 3 - 1,
                        tl.cn0);
          } else {
            tma_load_2d(smem_u32(smem_a + s * S::A_BYTES), &tmap_a, fb, kk * BK, tl.m0);
          }
          if (!prefetched) tma_load_2d(smem_u32(smem_b + s * S::B_BYTES), &tmap_b, fb, kk * BK, tl.n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The WHOLE warp walks the loop (convergent control flow) and one elected lane executes the tcgen05 instructions: the
    // operands of UTCHMMA live in uniform registers, and descriptor arithmetic done by all lanes of a converged warp is
    // provably uniform.  Inside an `if (lane == 0)` region it is not, and the compiler wraps every MMA in a per-lane
    // "waterfall" loop (ELECT / UTCHMMA / BRA.U.ANY) that cost ~40 cycles per instruction on top of the tensor-core floor.
    {
      constexpr uint32_t idesc = make_idesc(BN);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      int it = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const Tile tl = decode(i);
        const int as = i % NACC;
        // the epilogue must have drained this accumulator stage (tile i-NACC)
        mbar_wait(smem_u32(&tmem_empty_bar[as]), ((uint32_t)(i / NACC) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_u + (uint32_t)(as * BN);
        for (int kb = 0; kb < tl.num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          if (dbg && it == 0 && lane == 0) dbg[2] = clock64();        // first operand stage landed
          tc_fence_after();
          const uint64_t da = make_sw128_desc(smem_u32(smem_a + s * S::A_BYTES));
          const uint64_t db = make_sw128_desc(smem_u32(smem_b + s * S::B_BYTES));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the (>>4) address field
              umma_f16(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            }
            umma_commit(smem_u32(&empty_bar[s]));  // slot free once these MMAs have read it
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(smem_u32(&tmem_full_bar[as]));   // accumulator complete (fires immediately when num_kb == 0)
        __syncwarp();
        if (dbg && i == 0 && lane == 0) dbg[3] = clock64();        // last MMA of the first tile issued
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // 8 warps: warp w may touch TMEM lanes 32*(w%4)..+31 only, so two warps share each lane quarter and split the
    // tile's columns (h = 0: first half, h = 1: second half).
    pdl_wait();                         // bias rows / residual / output belong to the previous kernels until here
    const int q = warp & 3;             // TMEM lane quarter
    const int h = (warp - 2) >> 2;      // column half handled by this warp
    constexpr int NCH = BN / 16;        // 16-column chunks of the tile
    const bool geglu = epi.act == L2D_ACT_GEGLU;
    // chunk range of this warp (GEGLU: chunks of the value half; the matching gate chunk is cc + NCH/2)
    const int nch_eff = geglu ? NCH / 2 : NCH;
    const int cc_begin = h == 0 ? 0 : (nch_eff + 1) / 2;
    const int cc_end = h == 0 ? (nch_eff + 1) / 2 : nch_eff;
    for (int i = 0; i < my_tiles; ++i) {
      const Tile tl = decode(i);
      const int as = i % NACC;
      const int n0 = tl.n0;
      int row = tl.m0 + q * 32 + lane;
      bool row_ok = row < M;
      if (cg.enabled) {   // tile row -> (image, y, x) -> pixel index
        const int r = q * 32 + lane;
        const int iw = r % cg.bw, ih = (r / cg.bw) % cg.bh, in = r / (cg.bw * cg.bh);
        row_ok = tl.cn0 + in < cg.N;
        row = ((tl.cn0 + in) * cg.H + tl.ch0 + ih) * cg.W + tl.cw0 + iw;
      }
      const bool use_res = epi.residual != nullptr && splits == 1 && !geglu && row_ok;
      const __half* res_ptr = use_res ? epi.residual + (size_t)row * epi.ldr + n0 : nullptr;
      const __half* rg_ptr = (epi.rowgroup_bias && row_ok)
                                 ? epi.rowgroup_bias + (size_t)(row / epi.rows_per_group) * epi.rg_ld + n0 : nullptr;
      const __half* bias_ptr = epi.bias ? epi.bias + n0 : nullptr;
      // first chunk's residual is requested before the accumulator wait; each iteration requests the next one
      uint4 res0 = make_uint4(0, 0, 0, 0), res1 = res0;
      const bool v32 = (reinterpret_cast<uintptr_t>(epi.out) & 31) == 0 && (epi.ldo & 15) == 0 &&
                       (!epi.residual || ((reinterpret_cast<uintptr_t>(epi.residual) & 31) == 0 && (epi.ldr & 15) == 0));
      if (res_ptr && n0 + cc_begin * 16 < N) {
        if (v32 && n0 + cc_begin * 16 + 8 < N) {
          ldg256(res_ptr + cc_begin * 16, res0, res1);
        } else {
          res0 = *reinterpret_cast<const uint4*>(res_ptr + cc_begin * 16);
          if (n0 + cc_begin * 16 + 8 < N) res1 = *reinterpret_cast<const uint4*>(res_ptr + cc_begin * 16 + 8);
        }
      }
      // LayerNorm consumer: this row's mean / rstd from the producer's per-tile partial sums (fixed slot order)
      float ln_rstd = 1.f, ln_mr = 0.f;   // y = rstd * acc - (rstd * mean) * s[n] + b'[n]
      if (epi.ln_stats && row_ok) {
        const float2* sp = epi.ln_stats + (size_t)row * epi.ln_slots;
        float sm = 0.f, sq = 0.f;
        for (int t2 = 0; t2 < epi.ln_slots; ++t2) {
          const float2 v2 = __ldcg(sp + t2);
          sm += v2.x;
          sq += v2.y;
        }
        const float mean = sm * epi.ln_inv_c;
        ln_rstd = rsqrtf(fmaxf(sq * epi.ln_inv_c - mean * mean, 0.f) + epi.ln_eps);
        ln_mr = ln_rstd * mean;
      }
      float st_sum = 0.f, st_sq = 0.f;    // LayerNorm producer: sums over this warp's columns of the row
      mbar_wait(smem_u32(&tmem_full_bar[as]), (uint32_t)(i / NACC) & 1u);
      if (dbg && i == 0 && threadIdx.x == 64) dbg[4] = clock64();   // accumulator ready
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      if (CL) {
        // ---- cluster split-K, phase A: scatter this CTA's fp32 partial tile into the owners' shared memory ----
        const int W = BN / splits;                       // columns reduced by each CTA of the cluster (8 .. 64)
        const uint32_t my_slot = smem_u32(cl_part) + (uint32_t)((tl.z * 128 + q * 32 + lane) * W) * 4u;
#pragma unroll 1
        for (int cc = cc_begin; cc < cc_end; ++cc) {
          uint32_t r[16];
          if (tl.num_kb > 0) {
            tmem_ld16(taddr + cc * 16, r);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) r[e] = 0u;
          }
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const int col = cc * 16 + g4 * 4;
            const int owner = col / W, cin = col - owner * W;
            uint32_t raddr;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(my_slot + (uint32_t)cin * 4u), "r"(owner));
            asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(r[g4 * 4]), "r"(r[g4 * 4 + 1]),
                         "r"(r[g4 * 4 + 2]), "r"(r[g4 * 4 + 3])
                         : "memory");
          }
        }
      } else if (splits > 1) {
        // ---- split-K: store this split's fp32 partial tile (zeros when the split owns no k-blocks); slices are
        //      indexed by output row and summed by splitk_finish_kernel ----
        float* wrow = epi.ws + (size_t)tl.z * epi.ws_slice + (size_t)row * epi.ws_ld + n0;
#pragma unroll 1
        for (int cc = cc_begin; cc < cc_end; ++cc) {
          uint32_t r[16];
          if (tl.num_kb > 0) {
            tmem_ld16(taddr + cc * 16, r);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) r[e] = 0u;
          }
          if (row_ok) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              *reinterpret_cast<uint4*>(wrow + cc * 16 + q4 * 4) = make_uint4(r[q4 * 4], r[q4 * 4 + 1], r[q4 * 4 + 2], r[q4 * 4 + 3]);
          }
        }
      } else if (!geglu) {
        __half* out_ptr = epi.out + (size_t)row * epi.ldo + n0;
#pragma unroll 1
        for (int cc = cc_begin; cc < cc_end; ++cc) {
          const int c0 = cc * 16;                        // column offset inside the tile
          uint32_t r[16];
          tmem_ld16(taddr + c0, r);
          // request the next chunk's residual while the TMEM load is in flight
          uint4 nres0 = make_uint4(0, 0, 0, 0), nres1 = nres0;
          if (res_ptr && cc + 1 < cc_end && n0 + c0 + 16 < N) {
            if (v32 && n0 + c0 + 24 < N) {
              ldg256(res_ptr + c0 + 16, nres0, nres1);
            } else {
              nres0 = *reinterpret_cast<const uint4*>(res_ptr + c0 + 16);
              if (n0 + c0 + 24 < N) nres1 = *reinterpret_cast<const uint4*>(res_ptr + c0 + 24);
            }
          }
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]);
          const bool lo_ok = n0 + c0 < N, hi_ok = n0 + c0 + 8 < N;
          if (epi.ln_stats) {
            const float* sv = epi.ln_s + n0 + c0;
            const float* bv = epi.ln_b + n0 + c0;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              if (n0 + c0 + q4 * 4 < N) {
                const float4 s4 = __ldg(reinterpret_cast<const float4*>(sv + q4 * 4));
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bv + q4 * 4));
                v[q4 * 4 + 0] = fmaf(v[q4 * 4 + 0], ln_rstd, fmaf(-ln_mr, s4.x, b4.x));
                v[q4 * 4 + 1] = fmaf(v[q4 * 4 + 1], ln_rstd, fmaf(-ln_mr, s4.y, b4.y));
                v[q4 * 4 + 2] = fmaf(v[q4 * 4 + 2], ln_rstd, fmaf(-ln_mr, s4.z, b4.z));
                v[q4 * 4 + 3] = fmaf(v[q4 * 4 + 3], ln_rstd, fmaf(-ln_mr, s4.w, b4.w));
              }
            }
          }
          if (bias_ptr) {
            float b[8];
            if (lo_ok) {
              unpack8(ldg_cached(bias_ptr + c0), b);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] += b[e];
            }
            if (hi_ok) {
              unpack8(ldg_cached(bias_ptr + c0 + 8), b);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[8 + e] += b[e];
            }
          }
          if (rg_ptr) {
            float b[8];
            if (lo_ok) {
              unpack8(ldg_act(rg_ptr + c0), b);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] += b[e];
            }
            if (hi_ok) {
              unpack8(ldg_act(rg_ptr + c0 + 8), b);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[8 + e] += b[e];
            }
          }
          if (epi.act == L2D_ACT_SILU) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = silu_f(v[e]);
          } else if (epi.act == L2D_ACT_RELU) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
          }
          if (use_res) {
            float b[8];
            unpack8(res0, b);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] += b[e];
            unpack8(res1, b);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[8 + e] += b[e];
          }
          if (epi.act == L2D_ACT_RELU_POST) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
          }
          if (row_ok) {
            float lo[8], hi[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              lo[e] = v[e];
              hi[e] = v[8 + e];
            }
            const uint4 plo = pack8(lo), phi = pack8(hi);
            if (v32 && hi_ok) {
              stg256(out_ptr + c0, plo, phi);
            } else {
              if (lo_ok) *reinterpret_cast<uint4*>(out_ptr + c0) = plo;
              if (hi_ok) *reinterpret_cast<uint4*>(out_ptr + c0 + 8) = phi;
            }
            if (epi.stats_out) {   // statistics of the values as stored (fp16-rounded), like a LayerNorm reading them back
              float f[8];
              if (lo_ok) {
                unpack8(plo, f);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  st_sum += f[e];
                  st_sq = fmaf(f[e], f[e], st_sq);
                }
              }
              if (hi_ok) {
                unpack8(phi, f);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  st_sum += f[e];
                  st_sq = fmaf(f[e], f[e], st_sq);
                }
              }
            }
          }
          res0 = nres0;
          res1 = nres1;
        }
        if (epi.stats_out && row_ok)
          epi.stats_out[(size_t)row * epi.stats_slots + (n0 / BN) * 2 + h] = make_float2(st_sum, st_sq);
      } else {
        // GEGLU: tile columns [0,BN/2) = value half, [BN/2,BN) = matching gate half (weights interleaved by
        // l2d_geglu_interleave); output columns (n0/2) + ...
        constexpr int HB = BN / 2;
        const int n_out = N / 2;
        const int oc0 = n0 / 2;
        __half* out_ptr = epi.out + (size_t)row * epi.ldo + oc0;
#pragma unroll 1
        for (int cc = cc_begin; cc < cc_end; ++cc) {
          uint32_t rh[16], rgt[16];
          tmem_ld16(taddr + cc * 16, rh);
          tmem_ld16(taddr + HB + cc * 16, rgt);
          tmem_ld_wait();
          const int colh = n0 + cc * 16;       // column in the interleaved weight (value half)
          float v[16];
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            float bh[8], bg[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) bh[e] = bg[e] = 0.f;
            if (bias_ptr && colh + hlf * 8 < N) {
              unpack8(ldg_cached(bias_ptr + cc * 16 + hlf * 8), bh);
              unpack8(ldg_cached(bias_ptr + HB + cc * 16 + hlf * 8), bg);
            }
            float sh[8], sg[8];
            if (epi.ln_stats && colh + hlf * 8 < N) {   // LayerNorm consumer: (s, b') of the value and gate columns
              const float* svh = epi.ln_s + n0 + cc * 16 + hlf * 8;
              const float* bvh = epi.ln_b + n0 + cc * 16 + hlf * 8;
#pragma unroll
              for (int q4 = 0; q4 < 2; ++q4) {
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(svh + q4 * 4));
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(svh + HB + q4 * 4));
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(bvh + q4 * 4));
                const float4 d4 = __ldg(reinterpret_cast<const float4*>(bvh + HB + q4 * 4));
                sh[q4 * 4 + 0] = a4.x; sh[q4 * 4 + 1] = a4.y; sh[q4 * 4 + 2] = a4.z; sh[q4 * 4 + 3] = a4.w;
                sg[q4 * 4 + 0] = b4.x; sg[q4 * 4 + 1] = b4.y; sg[q4 * 4 + 2] = b4.z; sg[q4 * 4 + 3] = b4.w;
                bh[q4 * 4 + 0] = c4.x; bh[q4 * 4 + 1] = c4.y; bh[q4 * 4 + 2] = c4.z; bh[q4 * 4 + 3] = c4.w;
                bg[q4 * 4 + 0] = d4.x; bg[q4 * 4 + 1] = d4.y; bg[q4 * 4 + 2] = d4.z; bg[q4 * 4 + 3] = d4.w;
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) sh[e] = sg[e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              // the reference rounds the projection to fp16 before h * gelu(g) (GEGLU.forward on an fp16 Linear output)
              float ah, ag;
              if (epi.ln_stats) {
                ah = fmaf(__uint_as_float(rh[hlf * 8 + e]), ln_rstd, fmaf(-ln_mr, sh[e], bh[e]));
                ag = fmaf(__uint_as_float(rgt[hlf * 8 + e]), ln_rstd, fmaf(-ln_mr, sg[e], bg[e]));
              } else {
                ah = __uint_as_float(rh[hlf * 8 + e]) + bh[e];
                ag = __uint_as_float(rgt[hlf * 8 + e]) + bg[e];
              }
              const float hval = __half2float(__float2half_rn(ah));
              const float gval = __half2float(__float2half_rn(ag));
              v[hlf * 8 + e] = hval * gelu_erf_f(gval);
            }
          }
          if (row_ok && colh < N) {
            const int oc = oc0 + cc * 16;
            float lo[8], hi[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              lo[e] = v[e];
              hi[e] = v[8 + e];
            }
            if (v32 && oc + 8 < n_out && ((oc0 & 15) == 0)) {
              stg256(out_ptr + cc * 16, pack8(lo), pack8(hi));
            } else {
              if (oc < n_out) *reinterpret_cast<uint4*>(out_ptr + cc * 16) = pack8(lo);
              if (oc + 8 < n_out) *reinterpret_cast<uint4*>(out_ptr + cc * 16 + 8) = pack8(hi);
            }
          }
        }
      }
      // this warp has read everything it needs from accumulator stage `as`: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[as]));
      if (dbg && i == 0 && threadIdx.x == 64) dbg[5] = clock64();   // epilogue of the first tile done (warp 2)
    }
  }
  if constexpr (CL) {
    // every thread of every CTA of the cluster: partial tiles have landed in their owners' shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp >= 2) {
      // ---- phase B: this CTA reduces column slice z of the tile over the cluster, in CTA-rank order ----
      const Tile tl = decode(0);
      const int W = BN / splits;
      const int te = (int)threadIdx.x - 64;            // 0..255
      const int trow = te & 127, part = te >> 7;
      const int Wp = W >= 16 ? W / 2 : W;              // columns per thread group (multiple of 8)
      int row = tl.m0 + trow;
      bool row_ok = row < M;
      if (cg.enabled) {
        const int iw = trow % cg.bw, ih = (trow / cg.bw) % cg.bh, in = trow / (cg.bw * cg.bh);
        row_ok = tl.cn0 + in < cg.N;
        row = ((tl.cn0 + in) * cg.H + tl.ch0 + ih) * cg.W + tl.cw0 + iw;
      }
      if (row_ok && (W >= 16 || part == 0)) {
        const int cbase = part * Wp;                   // first column of this thread inside the slice
        for (int c8 = 0; c8 < Wp; c8 += 8) {
          const int col = tl.n0 + tl.z * W + cbase + c8;   // global output column
          if (col >= N) break;
          float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int z2 = 0; z2 < splits; ++z2) {
            const float* src = cl_part + (size_t)(z2 * 128 + trow) * W + cbase + c8;
            const float4 a = *reinterpret_cast<const float4*>(src), b2 = *reinterpret_cast<const float4*>(src + 4);
            v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b2.x; v[5] += b2.y; v[6] += b2.z; v[7] += b2.w;
          }
          if (epi.bias) {
            float b[8];
            unpack8(ldg_cached(epi.bias + col), b);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] += b[e];
          }
          if (epi.rowgroup_bias) {
            float b[8];
            unpack8(ldg_act(epi.rowgroup_bias + (size_t)(row / epi.rows_per_group) * epi.rg_ld + col), b);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] += b[e];
          }
          if (epi.act == L2D_ACT_SILU) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = silu_f(v[e]);
          } else if (epi.act == L2D_ACT_RELU) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
          }
          if (epi.residual) {
            float rr[8];
            unpack8(*reinterpret_cast<const uint4*>(epi.residual + (size_t)row * epi.ldr + col), rr);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] += rr[e];
          }
          if (epi.act == L2D_ACT_RELU_POST) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
          }
          *reinterpret_cast<uint4*>(epi.out + (size_t)row * epi.ldo + col) = pack8(v);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[6] = clock64();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// Second pass of a split-K GEMM: out[m, n0..n0+7] = epilogue(sum_z ws[z][m][n0..]) -- one thread per 8 columns
__global__ void __launch_bounds__(256) splitk_finish_kernel(const GemmEpilogue epi, int M, int N, int splits) {
  pdl_launch();
  pdl_wait();
  const int n8 = N >> 3;
  const size_t total = (size_t)M * n8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / n8), col = (int)(i % n8) * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* src = epi.ws + (size_t)row * epi.ws_ld + col;
    for (int z = 0; z < splits; ++z) {
      const float4 a = __ldcg(reinterpret_cast<const float4*>(src + (size_t)z * epi.ws_slice));
      const float4 b = __ldcg(reinterpret_cast<const float4*>(src + (size_t)z * epi.ws_slice + 4));
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if (epi.bias) {
      float b[8];
      unpack8(ldg_cached(epi.bias + col), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += b[e];
    }
    if (epi.rowgroup_bias) {
      float b[8];
      unpack8(ldg_act(epi.rowgroup_bias + (size_t)(row / epi.rows_per_group) * epi.rg_ld + col), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += b[e];
    }
    if (epi.act == L2D_ACT_SILU) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = silu_f(v[e]);
    } else if (epi.act == L2D_ACT_RELU) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
    }
    if (epi.residual) {
      float r[8];
      unpack8(*reinterpret_cast<const uint4*>(epi.residual + (size_t)row * epi.ldr + col), r);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += r[e];
    }
    if (epi.act == L2D_ACT_RELU_POST) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
    }
    *reinterpret_cast<uint4*>(epi.out + (size_t)row * epi.ldo + col) = pack8(v);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor-map cache + dispatch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  int64_t rows, cols, ld;
  int box_rows;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h ^= std::hash<int64_t>()(k.rows * 1315423911ll + k.cols) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h ^= std::hash<int64_t>()(k.ld * 31 + k.box_rows) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    return h;
  }
};

// 2D row-major fp16 [rows, cols] with row pitch ld elements; box = [box_rows, 64], 128B swizzle
static int get_tmap(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static std::mutex mu;
  TmapKey key{ptr, rows, cols, ld, box_rows};
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return L2D_OK;
  }
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(L2D_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(L2D_ERR_CUDA, "cuTensorMapEncodeTiled failed, CUresult " + std::to_string((int)r));
  if (cache.size() > 65536) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return L2D_OK;
}

// Cost model (SM cycles).  Measured on B200 (profiles/launches_r1.csv): a CTA's main loop is bound by the
// L2 -> shared-memory fill rate of its SM, ~42 B/clk, i.e. (128+BN)*128/42 cycles per 64-deep k-block, far above the
// tcgen05 issue time (BN/2 cycles).  So bigger N tiles are cheaper per flop, and small-M problems must be spread
// over the SMs by splitting K.
constexpr int kNumSms = 148;
constexpr int64_t kWsElems = 12 << 20;  // fp32 split-K workspace (48 MB)

struct GemmPlan {
  int bn, splits;
  bool cluster;   // split-K reduced inside a thread-block cluster (DSMEM) instead of slice buffer + finish kernel
};

// Off by default: measured neutral on B200 (profiles/README.md, "cluster split-K"): the shapes whose splits fit one wave of
// clusters are the ones where the finish kernel was already cheap.  L2D_SPLITK_CLUSTER=1 enables it.
static bool cluster_splitk_enabled() {
  static const bool on = [] {
    const char* e = getenv("L2D_SPLITK_CLUSTER");
    return e && e[0] == '1';
  }();
  return on;
}

static GemmPlan gemm_plan(int m, int n, int k, bool allow_split, int force_bn) {
  const int cands[4] = {256, 160, 128, 64};
  const int tiles_m = ceil_div(m, BM), num_kb = ceil_div(k, BK);
  GemmPlan best{64, 1, false};
  double best_cost = 1e30;
  const bool cl = cluster_splitk_enabled();
  for (int bn : cands) {
    if (force_bn > 0 && bn != force_bn) continue;
    const int tiles_n = ceil_div(n, bn), tiles = tiles_m * tiles_n;
    const double kb_cyc = (128.0 + bn) * 128.0 / 42.0;
    int max_s = 1;
    if (allow_split) max_s = std::max(1, std::min(32, num_kb / 4));
    for (int sp = 1; sp <= max_s; ++sp) {
      const int kb_per = ceil_div(num_kb, sp);
      if (sp > 1 && kb_per * (sp - 1) >= num_kb) continue;        // an empty last split: pointless
      for (int mode = 0; mode < 2; ++mode) {                      // 0: slice buffer + finish kernel, 1: cluster reduce
        if (mode == 1 && (sp == 1 || !cl)) continue;
        double cost;
        if (mode == 1) {
          // cluster split-K: the splits of a tile are one cluster of 2 / 4 / 8 CTAs, N tile 64 or 128 (partial tiles live in
          // shared memory next to a shortened operand ring); one wave only (the cluster kernel takes one tile per CTA);
          // clusters of 8 fill at most 16 per GPU (GPC granularity)
          if (!(sp == 2 || sp == 4 || sp == 8) || !(bn == 64 || bn == 128)) continue;
          const int slots = sp == 8 ? 128 : sp == 4 ? 144 : 148;
          if (tiles * sp > slots) continue;
          cost = kb_per * kb_cyc + 3000.0 + 1500.0 + 6.0 * bn;    // + DSMEM scatter, cluster barrier, slice reduce
        } else {
          if (sp > 1 && (int64_t)sp * tiles_m * BM * tiles_n * bn > kWsElems) continue;
          const int waves = ceil_div(tiles * sp, kNumSms);
          double cta = kb_per * kb_cyc + 3000.0;
          if (sp > 1) cta += 13.0 * bn;                             // fp32 store of the partial tile
          cost = waves * cta;
          if (sp > 1) cost += 6000.0 + (double)sp * m * n * 4.0 / (kNumSms * 40.0);   // finish kernel: launch + slice reads
        }
        if (cost < best_cost * 0.97 || (cost < best_cost && sp < best.splits)) {
          best_cost = cost;
          best = {bn, sp, mode == 1};
        }
      }
    }
  }
  return best;
}

int gemm_pick_tile_n(int m, int n, int k) { return gemm_plan(m, n, k, false, 0).bn; }
// row-statistics slots a producer GEMM of this shape writes per row: (N tiles) x (2 column halves)
int gemm_stats_slots(int m, int n, int k) { return 2 * ceil_div(n, gemm_plan(m, n, k, false, 0).bn); }

// exported for the K1 tensor-core kernel (kv_attn_mma.cu): 2-D map, box [box_rows, 64 columns], 128B swizzle
int get_tmap_2d(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out) {
  return get_tmap(ptr, rows, cols, ld, box_rows, out);
}

// channels-last activation [N,H,W,C] as a 4-D tensor map, box = [bn, bh, bw, 64 channels], 128B swizzle
static int get_tmap_conv(const void* ptr, int N, int H, int W, int C, int bw, int bh, int bn, CUtensorMap* out) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static std::mutex mu;
  TmapKey key{ptr, (int64_t)N * 1000003 + H, (int64_t)W * 1000003 + C, (int64_t)bw * 65536 + bh * 256 + bn, -4};
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return L2D_OK;
  }
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(L2D_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(L2D_ERR_CUDA, "cuTensorMapEncodeTiled(4D) failed, CUresult " + std::to_string((int)r));
  if (cache.size() > 65536) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return L2D_OK;
}

static int pow2_divisor(int x, int cap) {
  int p = 1;
  while (p * 2 <= cap && x % (p * 2) == 0) p *= 2;
  return p;
}

bool conv3x3_implicit_supported(int n_img, int h, int w, int cin) {
  if (cin % 64 != 0) return false;
  const int bw = pow2_divisor(w, 128), bh = pow2_divisor(h, 128 / bw);
  const int bn = 128 / (bw * bh);
  return bn <= 256 && bw * bh * bn == 128 && (bn == 1 || n_img >= 1);
}

static float* g_ws = nullptr;
static long long* g_dbg = nullptr;
// Set by the engine around its GEMM launches: the weight operand is engine-owned and constant, so the kernel may be
// launched with PDL and fetch weight tiles before the PDL wait.  Stand-alone l2d_gemm / l2d_conv3x3 calls (the caller may
// have just written `w`) launch fully serialised.
static thread_local bool t_weights_constant = false;
void gemm_weights_constant(bool on) { t_weights_constant = on; }

static int ensure_splitk_workspace() {
  if (g_ws) return L2D_OK;
  L2D_CUDA(cudaMalloc(&g_ws, kWsElems * sizeof(float)));
  return L2D_OK;
}

// cluster split-K launch: grid = tiles * splits, cluster = the `splits` CTAs of one tile
template <int BN, int STAGES>
static int launch_gemm_cluster(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& e, const ConvGeom& cg, int M,
                               int N, int K, int splits, int tiles_m, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * GemmSmem<BN>::STAGE_BYTES + 256 + (size_t)128 * BN * sizeof(float) + 1024;
  static bool configured = false;
  if (!configured) {
    L2D_CUDA(cudaFuncSetAttribute(gemm_f16_tcgen05_kernel<BN, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    configured = true;
  }
  const int tiles_n = ceil_div(N, BN);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles_n * tiles_m * splits);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = splits;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_family(0) && t_weights_constant && pdl_enabled()) ? 2 : 1;
  L2D_CUDA(cudaLaunchKernelEx(&cfg, gemm_f16_tcgen05_kernel<BN, STAGES, true>, ta, tb, e, cg, M, N, K, tiles_n, tiles_m, splits));
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

template <int BN, int STAGES, int NACC = 2>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& e, const ConvGeom& cg, int M,
                       int N, int K, int splits, int tiles_m, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * GemmSmem<BN>::STAGE_BYTES + (2 * STAGES + 6) * 8 + 1024;
  static bool configured = false;
  if (!configured) {
    L2D_CUDA(cudaFuncSetAttribute(gemm_f16_tcgen05_kernel<BN, STAGES, false, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    configured = true;
  }
  const int tiles_n = ceil_div(N, BN);
  const int total = tiles_n * tiles_m * splits;
  const int grid = total < kNumSms ? total : kNumSms;
  launch_pdl_if(pdl_family(0) && t_weights_constant, gemm_f16_tcgen05_kernel<BN, STAGES, false, NACC>, dim3(grid), dim3(320), smem, st,
                ta, tb, e, cg, M, N, K, tiles_n, tiles_m, splits);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

static int gemm_dispatch(const CUtensorMap& ta, const __half* w, int64_t ldw, const ConvGeom& cg, int tiles_m, int m_pad,
                         __half* out, int64_t ldo, int m, int n, int k, const __half* bias, const __half* rowgroup_bias,
                         int64_t rg_ld, int rows_per_group, const __half* residual, int64_t ldr, int act,
                         const GemmPlan& plan, cudaStream_t st, const GemmFusion* fx = nullptr) {
  const int bn = plan.bn;
  CUtensorMap tb;
  int rc = get_tmap(w, n, k, ldw, bn, &tb);
  if (rc != L2D_OK) return rc;
  GemmEpilogue e{out, ldo, bias, rowgroup_bias, rg_ld, rows_per_group > 0 ? rows_per_group : 1, residual, ldr, act,
                 nullptr, 0, 0, g_dbg};
  if (fx) {
    if (plan.splits != 1) return fail(L2D_ERR_INVALID, "gemm: LayerNorm fusion needs an unsplit plan");
    e.stats_out = fx->stats_out;
    e.stats_slots = 2 * ceil_div(n, bn);
    e.ln_stats = fx->ln_stats;
    e.ln_slots = fx->ln_slots;
    e.ln_s = fx->ln_s;
    e.ln_b = fx->ln_b;
    e.ln_inv_c = fx->ln_c > 0 ? 1.0f / (float)fx->ln_c : 0.f;
    e.ln_eps = fx->ln_eps;
    if (fx->ln_stats && (!fx->ln_s || !fx->ln_b || fx->ln_slots <= 0 || bias))
      return fail(L2D_ERR_INVALID, "gemm: LayerNorm consumer needs ln_s / ln_b / ln_slots and no separate bias");
  }
  if (plan.splits > 1 && plan.cluster) {
    if (bn == 64) return launch_gemm_cluster<64, 6>(ta, tb, e, cg, m, n, k, plan.splits, tiles_m, st);
    return launch_gemm_cluster<128, 4>(ta, tb, e, cg, m, n, k, plan.splits, tiles_m, st);
  }
  if (plan.splits > 1) {
    rc = ensure_splitk_workspace();
    if (rc != L2D_OK) return rc;
    e.ws = g_ws;
    e.ws_ld = (int64_t)ceil_div(n, bn) * bn;
    e.ws_slice = (int64_t)m_pad * e.ws_ld;
  }
  {
    switch (bn) {
      case 64: rc = launch_gemm<64, 6>(ta, tb, e, cg, m, n, k, plan.splits, tiles_m, st); break;
      case 128: rc = launch_gemm<128, 6>(ta, tb, e, cg, m, n, k, plan.splits, tiles_m, st); break;
      case 160: rc = launch_gemm<160, 5>(ta, tb, e, cg, m, n, k, plan.splits, tiles_m, st); break;
      case 256: rc = launch_gemm<256, 4>(ta, tb, e, cg, m, n, k, plan.splits, tiles_m, st); break;
      default: return fail(L2D_ERR_INVALID, "gemm: unsupported tile_n");
    }
  }
  if (rc != L2D_OK || plan.splits == 1) return rc;
  const size_t work = (size_t)m * (n / 8);
  const int blocks = (int)std::min<size_t>((work + 255) / 256, (size_t)kNumSms * 8);
  launch_pdl_if(pdl_family(1), splitk_finish_kernel, dim3(blocks), dim3(256), 0, st, e, m, n, plan.splits);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

int gemm_launch(const __half* a, int64_t lda, const __half* w, int64_t ldw, __half* out, int64_t ldo, int m, int n, int k,
                const __half* bias, const __half* rowgroup_bias, int64_t rg_ld, int rows_per_group,
                const __half* residual, int64_t ldr, int act, int force_bn, cudaStream_t st, const GemmFusion* fx) {
  const bool fused = fx && (fx->stats_out || fx->ln_stats);
  const GemmPlan plan = gemm_plan(m, n, k, act != L2D_ACT_GEGLU && !fused, force_bn);
  CUtensorMap ta;
  int rc = get_tmap(a, m, k, lda, BM, &ta);
  if (rc != L2D_OK) return rc;
  ConvGeom cg{};
  const int tiles_m = ceil_div(m, BM);
  return gemm_dispatch(ta, w, ldw, cg, tiles_m, tiles_m * BM, out, ldo, m, n, k, bias, rowgroup_bias, rg_ld, rows_per_group,
                       residual, ldr, act, plan, st, fused ? fx : nullptr);
}

// w'[n,k] = fp16(gamma[k] * w[n,k]) in place; ln_s[n] = sum_k w'[n,k]; ln_b[n] = bias[n] + sum_k beta[k] * w[n,k]  (one warp per row)
__global__ void __launch_bounds__(256) ln_fold_kernel(__half* __restrict__ w, const __half* __restrict__ bias,
                                                      const __half* __restrict__ gamma, const __half* __restrict__ beta,
                                                      float* __restrict__ ln_s, float* __restrict__ ln_b, int n, int k) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  __half* wr = w + (size_t)row * k;
  float s = 0.f, b = 0.f;
  for (int c = lane * 8; c < k; c += 256) {
    float wf[8], g[8], be[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(wr + c), wf);
    unpack8(*reinterpret_cast<const uint4*>(gamma + c), g);
    unpack8(*reinterpret_cast<const uint4*>(beta + c), be);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      o[e] = wf[e] * g[e];
      b = fmaf(be[e], wf[e], b);
    }
    const uint4 packed = pack8(o);
    *reinterpret_cast<uint4*>(wr + c) = packed;
    unpack8(packed, o);   // the sum runs over the weights the tensor cores will actually multiply by
#pragma unroll
    for (int e = 0; e < 8; ++e) s += o[e];
  }
  s = warp_sum(s);
  b = warp_sum(b);
  if (lane == 0) {
    ln_s[row] = s;
    ln_b[row] = b + (bias ? __half2float(bias[row]) : 0.f);
  }
}

int ln_fold_weights(__half* w, const __half* bias, const __half* gamma, const __half* beta, float* ln_s, float* ln_b, int n, int k,
                    cudaStream_t st) {
  if (k % 8 != 0) return fail(L2D_ERR_INVALID, "ln_fold_weights: K % 8 != 0");
  ln_fold_kernel<<<ceil_div(n, 8), 256, 0, st>>>(w, bias, gamma, beta, ln_s, ln_b, n, k);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

// out[N*H*W, cout] = epilogue( conv3x3(x[N,H,W,cin], w[cout, 9*cin]) ), pad 1, stride 1, x channels-last contiguous
int conv3x3_launch(const __half* x, int n_img, int h, int w_, int cin, const __half* w, __half* out, int64_t ldo, int cout,
                   const __half* bias, const __half* rowgroup_bias, int64_t rg_ld, int rows_per_group,
                   const __half* residual, int64_t ldr, int act, cudaStream_t st) {
  if (!conv3x3_implicit_supported(n_img, h, w_, cin)) return fail(L2D_ERR_INVALID, "conv3x3: shape not tileable");
  ConvGeom cg{};
  cg.enabled = 1;
  cg.cin_blocks = cin / BK;
  cg.H = h; cg.W = w_; cg.N = n_img;
  cg.bw = pow2_divisor(w_, 128);
  cg.bh = pow2_divisor(h, 128 / cg.bw);
  cg.bn = 128 / (cg.bw * cg.bh);
  cg.tiles_w = w_ / cg.bw;
  cg.tiles_h = h / cg.bh;
  const int tiles_m = cg.tiles_w * cg.tiles_h * ceil_div(n_img, cg.bn);
  const int m = n_img * h * w_, k = 9 * cin;
  // the plan's M is the padded tile count (tiles_m * 128 rows), which is what fills the SMs
  const GemmPlan plan = gemm_plan(tiles_m * BM, cout, k, true, 0);
  CUtensorMap ta;
  int rc = get_tmap_conv(x, n_img, h, w_, cin, cg.bw, cg.bh, cg.bn, &ta);
  if (rc != L2D_OK) return rc;
  // split-K slices are indexed by output pixel row, so a slice needs m rows (>= the largest pixel index + 1)
  return gemm_dispatch(ta, w, k, cg, tiles_m, m, out, ldo, m, cout, k, bias, rowgroup_bias, rg_ld, rows_per_group, residual,
                       ldr, act, plan, st);
}

}  // namespace l2d

using namespace l2d;

extern "C" int l2d_gemm_tile_n(int m, int n, int k) { return gemm_pick_tile_n(m, n, k); }

extern "C" void l2d_gemm_set_debug(void* timeline) {
  l2d::g_dbg = static_cast<long long*>(timeline);
}

extern "C" int l2d_conv3x3(const void* x, int n_img, int h, int w, int cin, const void* weight, void* out, int64_t ldo,
                           int cout, const void* bias, const void* rowgroup_bias, const void* residual, int64_t ldr,
                           int act, void* stream) {
  L2D_CHECK_ARG(x && weight && out, "null pointer");
  L2D_CHECK_ARG(n_img > 0 && h > 0 && w > 0 && cin > 0 && cout > 0, "empty problem");
  L2D_CHECK_ARG(cout % 8 == 0 && ldo % 8 == 0 && ldo >= cout, "Cout and ldo must be multiples of 8");
  L2D_CHECK_ARG(act == L2D_ACT_NONE || act == L2D_ACT_SILU || act == L2D_ACT_RELU || act == L2D_ACT_RELU_POST,
                "act must be none, SiLU, ReLU or ReLU-after-residual");
  L2D_CHECK_ARG(conv3x3_implicit_supported(n_img, h, w, cin), "need Cin % 64 == 0 and power-of-two tileable H, W");
  return conv3x3_launch((const __half*)x, n_img, h, w, cin, (const __half*)weight, (__half*)out, ldo, cout,
                        (const __half*)bias, (const __half*)rowgroup_bias, cout, h * w, (const __half*)residual, ldr, act,
                        (cudaStream_t)stream);
}

extern "C" int l2d_gemm(const void* a, int64_t lda, const void* w, void* out, int64_t ldo, int m, int n, int k,
                        const void* bias, const void* rowgroup_bias, int rows_per_group, const void* residual,
                        int64_t ldr, int act, void* stream) {
  L2D_CHECK_ARG(a && w && out, "null pointer");
  L2D_CHECK_ARG(m > 0 && n > 0 && k > 0, "empty problem");
  L2D_CHECK_ARG(k % 8 == 0 && n % 8 == 0 && lda % 8 == 0 && ldo % 8 == 0, "K, N, lda, ldo must be multiples of 8");
  L2D_CHECK_ARG(lda >= k, "lda < K");
  L2D_CHECK_ARG(act >= 0 && act <= 4, "bad act");
  L2D_CHECK_ARG(!residual || ldr % 8 == 0, "ldr must be a multiple of 8");
  L2D_CHECK_ARG(!rowgroup_bias || rows_per_group > 0, "rows_per_group must be > 0");
  L2D_CHECK_ARG(((uintptr_t)a % 16 == 0) && ((uintptr_t)w % 16 == 0) && ((uintptr_t)out % 16 == 0), "16-byte alignment");
  if (act == L2D_ACT_GEGLU) {
    const int bn = gemm_pick_tile_n(m, n, k);
    L2D_CHECK_ARG(n % bn == 0, "GEGLU: N must be a multiple of the N tile (see l2d_gemm_tile_n)");
  }
  return gemm_launch((const __half*)a, lda, (const __half*)w, k, (__half*)out, ldo, m, n, k, (const __half*)bias,
                     (const __half*)rowgroup_bias, n, rows_per_group, (const __half*)residual, ldr, act, 0,
                     (cudaStream_t)stream);
}
