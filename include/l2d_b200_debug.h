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
/* When non-NULL, CTA (0,0,0) of the tcgen05 attention kernel writes clock64() stamps of its first 16 key tiles:
 * timeline[(w*16 + j)*8 + k] for softmax warp w (0..7; 0-3 = query tile 0) -- k = 0 tile start, 1 S_j available, 2 scores in
 * registers, 3 max / lazy check / P.V wait done, 4 exp token acquired, 5 exponentials issued, 6 P_j handed over -- and
 * timeline[1024 + (t*16 + j)*2 + {0,1}] = the MMA thread's issue times of S_j and P_j.V_j for query tile t.  >= 1088 int64. */
void l2d_flash_set_debug(void* timeline);

#ifdef __cplusplus
}
#endif
#endif /* L2D_B200_DEBUG_H */
