/*
 * l2d_b200_debug.h -- developer hooks of libl2d_b200.so used by the scripts under profiles/.  NOT part of the drop-in
 * ABI (include/l2d_b200.h): nothing a user of the reference needs, no stability promise.
 */
#ifndef L2D_B200_DEBUG_H
#define L2D_B200_DEBUG_H

#include "l2d_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* When non-NULL, every CTA of the tensor-core K1 kernel writes cycles spent waiting for TMA data, appending/staging,
 * computing, storing, and its tile count to timeline[cta*16 ..]; NULL disables. */
void l2d_kv_attn_set_debug(void* timeline);
/* When non-NULL, every GEMM CTA writes 8 clock64() stamps (start, setup done, first stage landed, last MMA issued,
 * accumulator ready, epilogue done, all warps joined, unused) to timeline[cta*8 ..]; NULL disables. */
void l2d_gemm_set_debug(void* timeline);
/* profiles/ablate_families.py: skip every launch of the families whose bit (1 << family) is set -- plus bit 6 =
 * LayerNorm only, bit 7 = GroupNorm only -- so that the drop in frame time measures that family's true cost on the
 * graph's critical path.  Results are garbage while a mask is set; 0 restores the real step. */
void l2d_unet_set_ablation(l2d_unet* u, int family_mask);
/* Drop the captured whole-frame graph of a device-resident stream (it still holds the launches an ablation mask
 * removed / restored); the next l2d_stream_frame captures again. */
void l2d_stream_invalidate_graph(l2d_stream* s);
/* Resident CTAs per SM of the tcgen05 spatial-attention kernel (head_dim 40 -> 2 expected, 80 -> 1). */
int l2d_debug_flash_ctas_per_sm(int hd);

#ifdef __cplusplus
}
#endif
#endif /* L2D_B200_DEBUG_H */
