// Issue-rate probe for the legacy tensor path on sm_100a: how many mma.sync.m16n8k16 (f16 in, f32 acc) per clock can an
// SM retire with W warps, each running `ILP` independent accumulator chains?  (Bounds kv_attn_mma.cu and flash_attn.cu.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/bin/mma_rate profiles/mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void probe(float* out, int iters, long long* cyc) {
  float acc[ILP][4];
  for (int i = 0; i < ILP; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  unsigned a0 = threadIdx.x * 0x3c003c00u, a1 = a0 ^ 0x1234u, a2 = a0 + 7, a3 = a1 + 9, b0 = 0x3c003c00u, b1 = 0x38003800u;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < ILP; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int warps) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  const int iters = 4096;
  probe<ILP><<<148, warps * 32>>>(out, iters, cyc);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<ILP><<<148, warps * 32>>>(out, iters, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double mmas = (double)iters * ILP * warps;   // per SM
  printf("warps/SM %2d ILP %d: %.1f cycles per mma per SM (%.1f per SMSP), %.0f TFLOP/s, dependent-chain latency ~%.1f cycles\n", warps,
         ILP, c / mmas, c / mmas * 4, 148.0 * mmas * 4096 / (ms * 1e-3) / 1e12, (double)c / iters / (ILP > 1 ? 1 : 1));
  cudaFree(out);
  cudaFree(cyc);
}
int main() {
  run<1>(1);
  run<1>(4);
  run<2>(4);
  run<4>(4);
  run<4>(8);
  run<8>(8);
  run<4>(16);
  return 0;
}
