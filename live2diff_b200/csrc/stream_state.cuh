// Stream state that the reference keeps as Python-side tensors and advances with host control flow, as plain
// host+device functions (the same code runs in the device kernels of stream_state.cu and, for the CPU tests, on the host):
//   ring schedule   initialize_attn_bias_pe_and_update_idx / update_attn_bias
//                   (live2diff/pipeline_stream_animation_depth.py:403-414, 416-438)
//   re-noise RNG    torch.randn_like in predict_x0_batch (:596-598) -> counter-based Philox4x32-10 + Box-Muller, so the
//                   whole frame is one CUDA graph with no host-side generator state
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define L2D_HD __host__ __device__ __forceinline__
#else
#define L2D_HD inline
#endif

namespace l2d {

// :404-412 -- row 0 already sees the slot it is about to fill; rows >= 1 see only the sink slots; row 1 starts one slot
// further (the reference's unguarded `update_idx[1] = W0 + 1`).
L2D_HD void ring_init_row(int r, int32_t* valid, int64_t* pe_row, int64_t* update, int window, int warmup) {
  *valid = warmup + (r == 0 ? 1 : 0);
  for (int j = 0; j < window; ++j) pe_row[j] = j;
  *update = warmup + (r == 1 ? 1 : 0);
}

// :423-436 -- still filling: write the first masked slot, PE unchanged; full: rotate the rolling PEs right by one and
// overwrite the slot that now carries the largest PE (first index of the maximum, like Tensor.argmax on distinct values).
L2D_HD void ring_advance_row(int32_t* valid, int64_t* pe_row, int64_t* update, int window, int warmup) {
  if (*valid < window) {
    *update = *valid;
    *valid += 1;
    return;
  }
  const int64_t last = pe_row[window - 1];
  for (int j = window - 1; j > warmup; --j) pe_row[j] = pe_row[j - 1];
  pe_row[warmup] = last;
  int best = 0;
  for (int j = 1; j < window; ++j)
    if (pe_row[j] > pe_row[best]) best = j;
  *update = best;
}

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11)
L2D_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Standard normal for element `elem` of re-noise row `row` of frame `frame`: elements 2b and 2b+1 share one Philox block
// (counter = (b, row, frame_lo, frame_hi), key = seed) and take the cos / sin branch of one Box-Muller pair.
L2D_HD float stream_randn(uint64_t seed, uint64_t frame, uint32_t row, uint32_t elem) {
  uint32_t o[4];
  philox4x32_10(elem >> 1, row, (uint32_t)frame, (uint32_t)(frame >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), o);
  const float u1 = ((float)o[0] + 1.0f) * 2.3283064365386963e-10f;   // (0, 1]
  const float u2 = (float)o[1] * 2.3283064365386963e-10f;            // [0, 1]
  const float rad = sqrtf(-2.0f * logf(u1)), ang = 6.283185307179586f * u2;
  return (elem & 1u) ? rad * sinf(ang) : rad * cosf(ang);
}

}  // namespace l2d
