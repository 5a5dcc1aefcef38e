/*
 * meshcnn_b200 -- diagnostics switches and profiling read-outs.  NOT part of the production C ABI (meshcnn_b200.h).
 *
 * Everything here is process-global mutable state, meant for the bottleneck-attribution tools under tools/ and for the A/B
 * tests: do not call from production code, do not flip while another host thread is launching kernels.
 */
#ifndef MESHCNN_B200_DEBUG_H
#define MESHCNN_B200_DEBUG_H

#include "meshcnn_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* diagnostics only: 1 = mcnn_meshconv_bwd_data_tc_signs launches one CTA per (edge tile, channel block) instead of the
 * persistent kernel (same arithmetic; the scatter order differs) */
int mcnn_debug_set_tc_bd_tile(int on);

/* diagnostics only: output-channel tiles of 128 per CTA in mcnn_meshconv_bwd_weight_tc (1 or 2; 0 = built-in choice: 2 when
 * Cout % 256 == 0).  mcnn_meshconv_bwd_weight_tc_ws() follows the setting. */
int mcnn_debug_set_tc_bw_ot(int ot);
/* diagnostics only: 0 = mcnn_meshconv_bwd_weight_tc keeps both operands in shared memory (the first-generation kernel);
 * 1 (default) = dout^T is held in tensor memory */
int mcnn_debug_set_tc_bw_tmem_a(int on);

/* diagnostics only: force the grid size of mcnn_meshconv_fwd_tc_sk (0 = one CTA per SM) */
int mcnn_debug_set_tc_sk_ctas(int n);

/* diagnostics only (library built with -DMCNN_TC_PROF): 64 clock64 stamps of one CTA of the last stream-K launch */
int mcnn_debug_sk_prof(long long* out64, int cta);

/* diagnostics only (-DMCNN_TC_PROF): %globaltimer ns at entry / exit of each of the first 256 CTAs of the last stream-K launch */
int mcnn_debug_sk_cta_times(unsigned long long* out512);

/* diagnostics only: bottleneck-attribution knobs of the tensor-core forward kernel (0 = normal operation) */
int mcnn_debug_set_tc(int flags);

/* diagnostics only: force the N tile (64 / 128 / 256) of the tensor-core forward kernel; 0 = built-in choice */
int mcnn_debug_set_tc_ntile(int ntile);

/* the CTA-pair (tcgen05 cta_group::2) forward kernel for Cout % 256 == 0: on / off */
int mcnn_debug_set_tc_pair(int on);

/* diagnostics only (-DMCNN_TC_PROF builds): copy the 16 cycle counters of CTA 0 of the forward kernel to host memory */
int mcnn_debug_tc_prof(long long* out16_host, int reset);

/* diagnostics only: 0 = always the three-kernel form (statistics, finalize, apply); 1 (default) = per-mesh statistics
 * (GroupNorm, InstanceNorm2d) use the fused one-launch form when the mesh's slab fits in shared memory */
int mcnn_debug_set_norm_fused(int on);

/* diagnostics only: which collapse kernel mcnn_pool_collapse launches for a mesh that fits on chip: 0 (default) = by the size
 * of the incoming level (sequential up to 3000 edges, rounds above), 1 = always the round kernel, 2 = always the sequential one */
int mcnn_debug_set_pool_mode(int mode);

/* diagnostics only: programmatic dependent launch of every kernel of the library (default off; MESHCNN_B200_PDL=1): 0 = plain
 * stream-ordered launches */
int mcnn_debug_set_pdl(int on);

#ifdef __cplusplus
}
#endif

#endif /* MESHCNN_B200_DEBUG_H */
