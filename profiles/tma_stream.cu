// TMA streaming probe: what HBM read bandwidth can 148 persistent CTAs reach when each streams its own contiguous
// region of a [rows, C] fp16 matrix through a shared-memory ring, (a) as 64-column boxes (128 B pieces at a C*2-byte
// stride, the K1 layout) versus (b) through a [rows*C/64, 64] view whose boxes are fully contiguous in memory?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/bin/tma_stream profiles/tma_stream.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint32_t bar, uint32_t par) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(par) : "memory");
}
// stage = `boxes` boxes of [box_rows x 64] (box_rows*128 B each); step in rows of the map per stage = rows_per_stage
__global__ void __launch_bounds__(128, 1) stream(const __grid_constant__ CUtensorMap tm, int stages_total, int boxes, int box_rows,
                                                 int col_step, int rows_per_stage, int nstage, float* sink) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = sm_raw + ((1024u - (su32(sm_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[8], empty[8];
  const uint32_t stage_bytes = (uint32_t)boxes * box_rows * 128;
  const int per = stages_total / gridDim.x, s0 = blockIdx.x * per;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nstage; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&full[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 3;" ::"r"(su32(&empty[i])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0)
      for (int s = 0; s < per; ++s) {
        const int b = s % nstage, use = s / nstage;
        if (use > 0) mwait(su32(&empty[b]), (use - 1) & 1);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&full[b])), "r"(stage_bytes) : "memory");
        const int row = (s0 + s) * rows_per_stage;
        for (int x = 0; x < boxes; ++x) {
          const int c0 = col_step ? x * 64 : 0, c1 = col_step ? row : row + x * box_rows;
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(su32(sm + (size_t)b * stage_bytes + (size_t)x * box_rows * 128)), "l"(&tm), "r"(su32(&full[b])), "r"(c0),
                       "r"(c1) : "memory");
        }
      }
    return;
  }
  float acc = 0.f;
  for (int s = 0; s < per; ++s) {
    const int b = s % nstage;
    mwait(su32(&full[b]), (s / nstage) & 1);
    acc += reinterpret_cast<const float*>(sm + (size_t)b * stage_bytes)[threadIdx.x];   // touch the stage
    __syncwarp();
    if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(su32(&empty[b])) : "memory");
  }
  if (acc == 123.456f) sink[0] = acc;
}
int main() {
  PFN_enc enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  const int Cs[3] = {320, 640, 1280};
  for (int ci = 0; ci < 3; ++ci) {
    const int C = Cs[ci], ncb = C / 64;
    const size_t bytes = (size_t)3 << 30;                 // 3 GiB >> L2
    const int64_t rows = bytes / (C * 2);
    void* buf; float* sink;
    cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    const int R = 1280 / C * 16;                          // rows per K1 plane (40 KB)
    for (int mode = 0; mode < 2; ++mode) {
      CUtensorMap tm;
      cuuint32_t estr[2] = {1, 1};
      int boxes, box_rows, rows_per_stage, col_step;
      if (mode == 0) {   // strided 64-column boxes
        cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)rows}; cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)R};
        enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        boxes = ncb; box_rows = R; rows_per_stage = R; col_step = 1;
      } else {           // contiguous view [rows*ncb, 64], boxes of 160 rows (20 KB)
        cuuint64_t gdim[2] = {64, (cuuint64_t)rows * ncb}; cuuint64_t gstr[1] = {128};
        cuuint32_t box[2] = {64, 160};
        enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        boxes = 2; box_rows = 160; rows_per_stage = 320; col_step = 0;
      }
      for (int nstage = 2; nstage <= 4; nstage += 2) {
        const int stage_bytes = 40960;
        const int stages_total = (int)(bytes / stage_bytes) / 148 * 148;
        const size_t smem = (size_t)nstage * stage_bytes + 1024;
        cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        stream<<<148, 128, smem>>>(tm, stages_total, boxes, box_rows, col_step, rows_per_stage, nstage, sink);
        cudaEventRecord(e0);
        stream<<<148, 128, smem>>>(tm, stages_total, boxes, box_rows, col_step, rows_per_stage, nstage, sink);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("C=%4d %s ring %d x 40 KB: %.0f GB/s (%s)\n", C, mode ? "contiguous boxes" : "strided 64-col boxes", nstage,
               (double)stages_total * stage_bytes / (ms * 1e-3) / 1e9, cudaGetErrorString(err));
      }
    }
    cudaFree(buf); cudaFree(sink);
  }
  return 0;
}
