/*
 * rsb200_debug.h -- bring-up probes (hardware-behaviour experiments for tcgen05 descriptors and the MMA issue rate).
 * They live in their own library, librsb200_debug.so (csrc/rsb_debug.cu), and are NOT part of the product library
 * librsb200.so: nothing on the hot path, in the tools or in the tests calls them; scripts/gpu_probe_umma.py and
 * scripts/gpu_mma_rate.py do.
 */
#ifndef RSB200_DEBUG_H
#define RSB200_DEBUG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Hardware-behaviour probe for tcgen05 shared-memory descriptors (tests / bring-up only; see csrc/rsb_debug.cu).
 * mode 0: D[128][64] = A[row_offset + i][0:64] . B[n][0:64]^T with a K-major operand window starting `row_offset` rows
 *         into a TMA-written box of `a_rows` rows (descriptor base_offset as given).
 * mode 1: D[m][n] = sum_k A[k][m] * B[n][k] with A given as [64 k][128 m] (MN-major operand), descriptor LBO/SBO as given. */
int rsb_debug_umma(const void* a, int32_t a_total_rows, int32_t a_cols, const void* b, float* out, int32_t mode, int32_t a_rows,
                   int32_t a_blocks, int32_t row_offset, int32_t base_offset, int32_t lbo, int32_t sbo, int32_t k_step_bytes,
                   void* stream);

/* tcgen05.mma issue-rate probe (bring-up only): every CTA (pair != 0: every CTA pair) issues 4*iters MMAs of
 * M=128 (256 for a pair) x block_n x 16 on zeroed operands; out[cta] = cycles per MMA as seen by the issuing thread. */
int rsb_debug_mma_rate(float* out, int32_t grid, int32_t pair, int32_t block_n, int32_t iters, int32_t commit_each, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RSB200_DEBUG_H */
