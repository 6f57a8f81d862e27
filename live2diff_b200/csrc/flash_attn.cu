// K2/K3 -- spatial self-attention (S x S, S = hw) and text cross-attention (S x 77):
//     O = softmax(Q K^T / sqrt(hd)) V   per (batch, head), no mask.
//
// Reference semantics: diffusers 0.25.0 `Attention` + AttnProcessor2_0 (F.scaled_dot_product_attention)
// used for attn1/attn2 of BasicTransformerBlock (attention.py:173-194, 243, 251-253); heads = 8,
// hd = 40/80/160 (SURVEY A-10).
//
// Flash-style single pass (online softmax in fp32, P rounded to fp16 for the P.V product exactly like
// the fused SDPA kernels the reference dispatches to).  64 query rows per CTA (4 warps x 16 rows), 64-key
// tiles double-buffered with cp.async, operands fed to the tensor cores with ldmatrix.  This first
// version issues warp-level mma.sync (HMMA) tiles; the tcgen05/TMEM rewrite is listed in DESIGN.md.
#include "ops.cuh"

namespace l2d {

constexpr int FA_BM = 64;
constexpr int FA_BN = 64;

struct FlashParams {
  const __half* q;
  const __half* k;
  const __half* v;
  __half* o;
  int64_t ldq, ldk, ldv, ldo;
  int heads, sq, skv, hd;
  float scale_log2;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// load a [64 x hd] tile (rows row0.., zero-filled beyond `rows_total`) into smem with row pitch HDP halves
template <int HD>
__device__ __forceinline__ void load_tile(__half* smem, const __half* g, int64_t ld, int row0, int rows_total, int hd,
                                          int tid) {
  constexpr int HDP = HD + 8;
  const int chunks = hd >> 3;
  for (int i = tid; i < 64 * chunks; i += 128) {
    const int r = i / chunks, c = i - r * chunks;
    const int gr = row0 + r;
    const bool ok = gr < rows_total;
    const __half* src = g + (size_t)(ok ? gr : 0) * ld + c * 8;
    cp_async16((uint32_t)__cvta_generic_to_shared(smem + r * HDP + c * 8), src, ok ? 16 : 0);
  }
}

template <int HD>
__global__ void __launch_bounds__(128) flash_attn_kernel(const FlashParams p) {
  constexpr int HDP = HD + 8;       // padded pitch: conflict-free ldmatrix
  constexpr int KS = HD / 16;       // k-steps of Q.K^T
  constexpr int OT = HD / 8;        // n8 tiles of O
  extern __shared__ __align__(16) __half fa_smem[];
  __half* sQ = fa_smem;
  __half* sK = sQ + FA_BM * HDP;
  __half* sV = sK + 2 * FA_BN * HDP;

  pdl_launch();
  pdl_wait();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y, m0 = blockIdx.x * FA_BM;
  const __half* gq = p.q + (size_t)b * p.sq * p.ldq + (size_t)h * p.hd;
  const __half* gk = p.k + (size_t)b * p.skv * p.ldk + (size_t)h * p.hd;
  const __half* gv = p.v + (size_t)b * p.skv * p.ldv + (size_t)h * p.hd;

  // zero everything once: padding columns [hd, HD+8) must read as 0 for the k-dim of Q.K^T
  for (int i = tid; i < (FA_BM + 4 * FA_BN) * HDP / 8; i += 128) reinterpret_cast<uint4*>(fa_smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();

  const int nkv = (p.skv + FA_BN - 1) / FA_BN;
  load_tile<HD>(sQ, gq, p.ldq, m0, p.sq, p.hd, tid);
  load_tile<HD>(sK, gk, p.ldk, 0, p.skv, p.hd, tid);
  load_tile<HD>(sV, gv, p.ldv, 0, p.skv, p.hd, tid);
  cp_async_commit();

  float o_acc[OT][4];
#pragma unroll
  for (int i = 0; i < OT; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[KS][4];

  const int g = lane >> 2, t = lane & 3;
  for (int it = 0; it < nkv; ++it) {
    const int buf = it & 1;
    cp_async_wait<0>();
    __syncthreads();
    if (it + 1 < nkv) {
      load_tile<HD>(sK + (buf ^ 1) * FA_BN * HDP, gk, p.ldk, (it + 1) * FA_BN, p.skv, p.hd, tid);
      load_tile<HD>(sV + (buf ^ 1) * FA_BN * HDP, gv, p.ldv, (it + 1) * FA_BN, p.skv, p.hd, tid);
      cp_async_commit();
    }
    if (it == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = ks * 16 + (lane >> 4) * 8;
        ldsm_x4((uint32_t)__cvta_generic_to_shared(sQ + r * HDP + c), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    const __half* cK = sK + buf * FA_BN * HDP;
    const __half* cV = sV + buf * FA_BN * HDP;

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s_acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s_acc[i][0] = s_acc[i][1] = s_acc[i][2] = s_acc[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        const int mi = lane >> 3;
        const int r = np * 16 + (lane & 7) + (mi >> 1) * 8;
        const int c = ks * 16 + (mi & 1) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4((uint32_t)__cvta_generic_to_shared(cK + r * HDP + c), b0, b1, b2, b3);
        mma16816(s_acc[np * 2], qf[ks], b0, b1);
        mma16816(s_acc[np * 2 + 1], qf[ks], b2, b3);
      }
    }
    // ---- online softmax (rows g and g+8 of this warp's 16) ----
    // scores stay unscaled; the softmax scale is folded into the exp2 argument: p = 2^(s*c - m*c), one FFMA + one
    // MUFU per score.  Only the last key tile can hold out-of-range columns.
    const int kv0 = it * FA_BN;
    if (kv0 + FA_BN > p.skv) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = kv0 + nt * 8 + 2 * t + (e & 1);
          if (col >= p.skv) s_acc[nt][e] = -INFINITY;
        }
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s_acc[nt][0], s_acc[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s_acc[nt][2], s_acc[nt][3]));
    }
    float corr[2], mneg[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);            // running max of the raw scores
      corr[r] = exp2f((m_run[r] - m_new) * p.scale_log2);
      m_run[r] = m_new;
      mneg[r] = -m_new * p.scale_log2;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];  // P as A fragments for the 4 k16 steps over keys
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(fmaf(s_acc[nt][0], p.scale_log2, mneg[0]));
      const float p1 = exp2f(fmaf(s_acc[nt][1], p.scale_log2, mneg[0]));
      const float p2 = exp2f(fmaf(s_acc[nt][2], p.scale_log2, mneg[1]));
      const float p3 = exp2f(fmaf(s_acc[nt][3], p.scale_log2, mneg[1]));
      rs[0] += p0 + p1;                                       // fp32 row sum (as flash-attention does)
      rs[1] += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2 + 0] = h2_as_u32(__floats2half2_rn(p0, p1));
      pf[nt >> 1][(nt & 1) * 2 + 1] = h2_as_u32(__floats2half2_rn(p2, p3));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int i = 0; i < OT; ++i) {
      o_acc[i][0] *= corr[0];
      o_acc[i][1] *= corr[0];
      o_acc[i][2] *= corr[1];
      o_acc[i][3] *= corr[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < OT / 2; ++dp) {
        const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = dp * 16 + (lane >> 4) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t((uint32_t)__cvta_generic_to_shared(cV + r * HDP + c), b0, b1, b2, b3);
        mma16816(o_acc[dp * 2], pf[kk], b0, b1);
        mma16816(o_acc[dp * 2 + 1], pf[kk], b2, b3);
      }
    }
  }
  // ---- finalize ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int row0 = m0 + warp * 16 + g, row1 = row0 + 8;
  __half* go = p.o + (size_t)b * p.sq * p.ldo + (size_t)h * p.hd;
#pragma unroll
  for (int i = 0; i < OT; ++i) {
    const int d = i * 8 + 2 * t;
    if (d < p.hd) {
      if (row0 < p.sq)
        *reinterpret_cast<__half2*>(go + (size_t)row0 * p.ldo + d) = __floats2half2_rn(o_acc[i][0] * inv0, o_acc[i][1] * inv0);
      if (row1 < p.sq)
        *reinterpret_cast<__half2*>(go + (size_t)row1 * p.ldo + d) = __floats2half2_rn(o_acc[i][2] * inv1, o_acc[i][3] * inv1);
    }
  }
}

template <int HD>
static int launch_flash(const FlashParams& p, int batch, cudaStream_t st) {
  constexpr size_t smem = (size_t)(FA_BM + 4 * FA_BN) * (HD + 8) * sizeof(__half);
  static bool configured = false;
  if (!configured) {
    L2D_CUDA(cudaFuncSetAttribute(flash_attn_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid(ceil_div(p.sq, FA_BM), p.heads, batch);
  launch_pdl_if(pdl_family(2), flash_attn_kernel<HD>, grid, dim3(128), smem, st, p);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

int attention_launch(const __half* q, int64_t ldq, const __half* k, int64_t ldk, const __half* v, int64_t ldv, __half* o,
                     int64_t ldo, int batch, int heads, int sq, int skv, int hd, cudaStream_t st) {
  if (attention_tcgen05_supported(q, k, v, o, ldq, ldk, ldv, ldo, sq, skv, hd))
    return attention_tcgen05_launch(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, sq, skv, hd, st);
  FlashParams p{q, k, v, o, ldq, ldk, ldv, ldo, heads, sq, skv, hd, 1.4426950408889634f / sqrtf((float)hd)};
  const int hdp = ((hd + 15) / 16) * 16;
  switch (hdp) {
    case 16: return launch_flash<16>(p, batch, st);
    case 32: return launch_flash<32>(p, batch, st);
    case 48: return launch_flash<48>(p, batch, st);
    case 64: return launch_flash<64>(p, batch, st);
    case 80: return launch_flash<80>(p, batch, st);
    case 96: return launch_flash<96>(p, batch, st);
    case 128: return launch_flash<128>(p, batch, st);
    case 160: return launch_flash<160>(p, batch, st);
    default: return fail(L2D_ERR_INVALID, "attention: unsupported head_dim " + std::to_string(hd));
  }
}

}  // namespace l2d

using namespace l2d;

extern "C" int l2d_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                             int64_t ldo, int batch, int heads, int sq, int skv, int hd, void* stream) {
  L2D_CHECK_ARG(q && k && v && out, "null pointer");
  L2D_CHECK_ARG(batch > 0 && heads > 0 && sq > 0 && skv > 0, "empty problem");
  L2D_CHECK_ARG(hd % 8 == 0 && hd <= 160, "head_dim must be a multiple of 8 and <= 160");
  L2D_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0, "row strides must be multiples of 8");
  return attention_launch((const __half*)q, ldq, (const __half*)k, ldk, (const __half*)v, ldv, (__half*)out, ldo, batch,
                          heads, sq, skv, hd, (cudaStream_t)stream);
}
