/*
 * mnv_debug.h -- tuning / debug hook of the TUNING build of the kernel library
 * (minerva_b200/lib/libmnv_b200_tuning.so, compiled with -DMNV_TUNING).
 *
 * The product library libmnv_b200.so does NOT export this symbol: there every option below is a
 * compile-time constant, so the product has no process-global mutable state and the
 * SIMT checker kernel is not in its binary.  tools/ (opbench --mnv-opt, tune_opt, *_diag) and the
 * alternate-operand-path parity tests load the tuning build through minerva_b200._lib.load_tuning().
 *
 * Options are process-wide in the tuning build (it exists to A/B whole runs).  Keys:
 *   simt            1: route GEMM / conv through the one-thread-per-output SIMT checker kernel
 *   max_splits      >0: clamp split-K
 *   no_tma          1: gather B with threads even where TMA applies
 *   no_fwd_bwd      1: generic backward-data gather for stride 1 too
 *   no_ktab         1: table-free forward gather
 *   wait_hint       mbarrier.try_wait suspend hint (ns)
 *   no_wide no_deep no_tall tall_min_stages no_tail pf_dist   tile-shape / schedule selection
 *   no_tma_a        bit 0: no TMA-im2col fprop/dgrad, bit 1: no TMA MatMult A, bit 2: no TMA wgrad
 *   tma_tf32        0: FLOAT32-typed tensor maps (tensor core truncates) instead of TFLOAT32 (TMA rounds)
 *   force_tma_a no_klane no_s2d no_shift shift_dbg s2d_im2col   path selection for experiments
 *   sm_budget       SMs the persistent tensor-core kernel may occupy
 * Returns the previous value, -1 for an unknown key.
 */
#ifndef MNV_DEBUG_H_
#define MNV_DEBUG_H_
#ifdef __cplusplus
extern "C" {
#endif
int mnv_debug_set_option(const char* key, int value);
#ifdef __cplusplus
}
#endif
#endif /* MNV_DEBUG_H_ */
