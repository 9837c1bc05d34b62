/*
 * largesteps_b200_diag.h -- diagnostics and timing harnesses of libls_b200.so.  NOT part of the drop-in boundary
 * (include/largesteps_b200.h): nothing on the product path calls these; bench.py and the scripts under profiles/ do.
 */
#ifndef LARGESTEPS_B200_DIAG_H
#define LARGESTEPS_B200_DIAG_H

#include "largesteps_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* in-solver SpMM of the handle's own matrix copy on SoA planes, for profiling the dominant kernel:
 *   runs `launches` back-to-back launches of the solver's SpMM+dot kernel on its internal p/Ap planes.  */
int ls_pcg_bench_spmm(void *handle, int k, int launches, void *stream);
/* timing harness for the iteration kernels, launched back-to-back from C (a Python-level loop is launch-bound):
 *   `launches` launches rotating over `n_handles` handles (use enough handles that matrix+vectors exceed L2 for an
 *   HBM-cold number, one handle for the L2-resident number).  which: 0 SpMM+dot, 1 update, 2 p-update, 3 all three,
 *   4 SpMM without the dot-product epilogue (the plain SpMV of BASELINE's metric). */
int ls_pcg_bench(void **handles, int n_handles, int k, int which, int launches, void *stream);
/* with LS_PCG_PROFILE set in the environment the persistent kernel's CTA 0 accumulates SM-clock cycles per phase of
 * the last solve.  Fused solver: out8 = [phase A (SpMV + x/p/s update), all-reduce of p.s, phase B (r, z), all-reduce of
 * r.z / r.r (publishes z), true-residual restarts, 0, restarts, iterations]; round-1 kernel: [SpMM phase, all-reduce 1,
 * update phase, all-reduce 2, p-update phase, barrier 3, 0, iterations]                                                */
int ls_pcg_phase_cycles(void *handle, int64_t *out, int n /* 8, or 8 + 8*grid for the per-CTA table (.., smid, it) */, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LARGESTEPS_B200_DIAG_H */
