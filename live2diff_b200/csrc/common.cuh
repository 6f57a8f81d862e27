// Shared device/host helpers for libl2d_b200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>

#include "../../include/l2d_b200.h"

namespace l2d {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
void count_launch(int n = 1);

#define L2D_CHECK_ARG(cond, msg)                                   \
  do {                                                             \
    if (!(cond)) return ::l2d::fail(L2D_ERR_INVALID, std::string(__func__) + ": " + (msg)); \
  } while (0)

#define L2D_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (call);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::l2d::fail(L2D_ERR_CUDA, std::string(__func__) + ": " #call ": " + cudaGetErrorString(_e)); \
  } while (0)

// kernel launch check (no sync): catches bad configs immediately
#define L2D_LAUNCH_CHECK()                                                                      \
  do {                                                                                          \
    ::l2d::count_launch();                                                                      \
    cudaError_t _e = cudaPeekAtLastError();                                                     \
    if (_e != cudaSuccess) {                                                                    \
      cudaGetLastError();                                                                       \
      return ::l2d::fail(L2D_ERR_CUDA, std::string(__func__) + ": launch: " + cudaGetErrorString(_e)); \
    }                                                                                           \
  } while (0)

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// A frame is ~850 dependent kernels of ~12 us each, so the launch -> CTA-scheduling -> prologue latency between two
// kernels is a double-digit share of the frame.  Every hot kernel is launched with the programmatic-stream-serialization
// attribute: it calls pdl_launch() first (lets the NEXT kernel's CTAs be scheduled as soon as all of this kernel's CTAs
// have started) and pdl_wait() before it touches global memory that the previous kernel may still be reading or writing
// (griddepcontrol.wait = the previous grid has completed and flushed).  Work that only touches memory no neighbouring
// kernel writes (barrier init, TMEM allocation, tensor-map prefetch, K1's PE-window / KV-plane loads) runs before the
// wait and overlaps the previous kernel's tail.  L2D_PDL=0 launches everything fully serialised (the waits are no-ops).
bool pdl_enabled();
bool pdl_family(int bit);   // L2D_PDL_MASK (developer): bit 0 gemm, 1 split-K finish, 2 flash, 3 K1, 4 GroupNorm, 5 LayerNorm
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_if(bool use_pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (use_pdl && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  return launch_pdl_if(true, kernel, grid, block, smem, st, std::forward<Args>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers -------------------------------------------------------------------------
struct __align__(16) half8 {
  __half2 v[4];
};

__device__ __forceinline__ uint4 ldg_stream(const void* p) {  // read-once data: keep it out of L1
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_cached(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
// Activations written by a neighbouring kernel: under PDL this kernel's lifetime overlaps its producer's, so the data
// is NOT read-only for the kernel's lifetime and the non-coherent path (ld.global.nc / __ldg) may return stale L1 lines.
// ld.global.cg reads at L2, the coherence point that griddepcontrol.wait orders against.  (Parameters -- weights, biases,
// PE tables -- are constant for the whole step and keep the .nc path.)
__device__ __forceinline__ uint4 ldg_act(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

__device__ __forceinline__ __half2 u32_as_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h2_as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 a = __half22float2(u32_as_h2(u.x)), b = __half22float2(u32_as_h2(u.y));
  float2 c = __half22float2(u32_as_h2(u.z)), d = __half22float2(u32_as_h2(u.w));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = h2_as_u32(__floats2half2_rn(f[0], f[1]));
  u.y = h2_as_u32(__floats2half2_rn(f[2], f[3]));
  u.z = h2_as_u32(__floats2half2_rn(f[4], f[5]));
  u.w = h2_as_u32(__floats2half2_rn(f[6], f[7]));
  return u;
}
// fp16 + fp16 -> fp16 (one rounding), the same value torch's half add produces
__device__ __forceinline__ uint4 hadd8(const uint4& a, const uint4& b) {
  uint4 r;
  r.x = h2_as_u32(__hadd2(u32_as_h2(a.x), u32_as_h2(b.x)));
  r.y = h2_as_u32(__hadd2(u32_as_h2(a.y), u32_as_h2(b.y)));
  r.z = h2_as_u32(__hadd2(u32_as_h2(a.z), u32_as_h2(b.z)));
  r.w = h2_as_u32(__hadd2(u32_as_h2(a.w), u32_as_h2(b.w)));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// single-instruction MUFU forms (1 ulp): an IEEE divide / __frcp_rn costs 8-10 instructions with a slow-path branch, which
// is what the GEGLU and SiLU epilogues are made of; the result is rounded to fp16 (2^-11) right after
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// x * sigmoid(x); x -> -inf: ex2 -> +inf, rcp -> 0, result -0; x -> +inf: ex2 -> 0, result x
__device__ __forceinline__ float silu_f(float x) { return x * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
// exact-erf GELU (diffusers GEGLU uses F.gelu default).  erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, two
// orders below an fp16 ulp of the result) with MUFU rcp/ex2: ~14 instructions instead of erff's ~40.
__device__ __forceinline__ float erf_as_f(float x) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = ex2_approx(-ax * ax * 1.4426950408889634f);
  return copysignf(fmaf(-p * t, e, 1.0f), x);
}
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erf_as_f(x * 0.70710678118654752f)); }

}  // namespace l2d
