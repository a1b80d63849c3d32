/*
 * wc_b200_debug.h — diagnostic entry points of libwc_b200.so.  NOT part of the drop-in boundary (include/wc_b200.h):
 * nothing on the product path calls them; they exist for the profiling drivers under scratch/ and for A/B tests.
 *
 * They are the library's only process-wide mutable state (one int of flags, default 0, and -- when bit 0 is set -- one
 * device buffer of globaltimer marks).  With the flags at their default the compute path is re-entrant and stateless as
 * include/wc_b200.h promises.
 */
#ifndef WC_B200_DEBUG_H_
#define WC_B200_DEBUG_H_
#ifdef __cplusplus
extern "C" {
#endif

/* bit 0: kernels record per-CTA globaltimer marks (allocates the mark buffer on first use); the other bits are
 * kernel-specific experiment switches documented next to their use in csrc/ (results may be wrong by construction:
 * they switch parts of a kernel off to time the rest).  Returns 0 or a cudaError_t.                                  */
int wc_debug_set(int flags);
/* copy the [160 blocks][8 slots] mark buffer to host memory (u64 nanoseconds); -1 if marks were never enabled        */
int wc_debug_marks(unsigned long long* host_out);
/* per-CTA / per-phase timestamps of the last Cholesky launch (cluster kernel: [8][32] u64; one-CTA kernel: [64] i64)  */
int wc_debug_chol_cluster_timing(unsigned long long* host_out);
int wc_debug_chol_timing(long long* host_out);
/* The program (stages, grid, jobs) the backward chain launch would run for a Cholesky layer with Cin = Cout = C, group size g,
 * K classes and E embedding filters, written to out[] WITHOUT touching the device (host logic only; used by the CPU tests):
 * out = {stages, grid, entries, stage_total[6], then per entry {type, stage, virtual CTAs, first part, chunks per part,
 * atomic + 2 * epilogue}}.  what: 1 unconditional kernel, 2 per-class kernels, 4 factorized branch, 8 adjoint buffer zeroed
 * by the caller (split class sums), 16 scratch matrices zeroed (K-split of stage 2).  Returns 0 or a WC_E_* code.          */
int wc_debug_bwd_chain_plan(int C, int g, int K, int E, int what, int* out, int out_len);

#ifdef __cplusplus
}
#endif
#endif
