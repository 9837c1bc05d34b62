// ls_spmm.cu -- SpMM launch plumbing + the public y = A x entry point (AoS (V,k) torch layout).
// Replaces the torch sparse `M @ v` of parameterize.py:30 (to_differential) and scripts/main.py:192-195.
#include <stdlib.h>
#include "ls_spmm_kernel.cuh"
#include "ls_spmm_host.h"

namespace lsk {

// runtime tuning knobs (defaults chosen on B200; LS_SPMM_* environment variables exist for sweeps only)
void spmm_config(SpmmCfg *cfg) {
    static SpmmCfg s = {0, 0, 0, 0, 0};
    if (s.stages == 0) {
        auto geti = [](const char *name, int dflt, int lo, int hi) {
            const char *e = getenv(name);
            int v = e ? atoi(e) : dflt;
            return v < lo ? lo : (v > hi ? hi : v);
        };
        s.stages = geti("LS_SPMM_STAGES", 2, 2, SPMM_MAX_STAGES);
        s.cap = SPMM_NT * geti("LS_SPMM_CAPMUL", 8, 2, 24);
        s.unroll = 8;
        s.hint = geti("LS_SPMM_HINT", 1, 0, 2);   // in-solver matrix stream: L2 evict_first
        s.debug = geti("LS_SPMM_DEBUG", 0, 0, 3);  // diagnostics only
    }
    *cfg = s;
}

namespace {
template <int K, bool X4, bool YSOA, bool DOT, int U>
int prepare_t(const SpmmCfg &cfg, int *ctas_per_sm) {
    static thread_local int cached_dev = -1, cached_occ = 0, cached_st = 0, cached_cap = 0;
    LsDevInfo di;
    int rc = ls_dev_info(&di);
    if (rc) return rc;
    if (cached_dev != di.device || cached_st != cfg.stages || cached_cap != cfg.cap) {
        size_t smem = spmm_smem_bytes(cfg.stages, cfg.cap);
        if ((int)smem > di.max_smem_optin) {
            ls_set_error("SpMM stage configuration needs %zu bytes of shared memory, device allows %d", smem, di.max_smem_optin);
            return LS_ERR_UNSUPPORTED;
        }
        LS_CUDA_TRY(cudaFuncSetAttribute(spmm_tma_kernel<K, X4, YSOA, DOT, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        LS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmm_tma_kernel<K, X4, YSOA, DOT, U>, SPMM_THREADS, smem));
        if (occ < 1) occ = 1;
        cached_dev = di.device;
        cached_occ = occ;
        cached_st = cfg.stages;
        cached_cap = cfg.cap;
    }
    *ctas_per_sm = cached_occ;
    return LS_OK;
}

template <int K, bool X4, bool YSOA, bool DOT, int U>
int launch_t(const SpmmArgs &a, int grid, cudaStream_t stream) {
    size_t smem = spmm_smem_bytes(a.stages, a.cap);
    spmm_tma_kernel<K, X4, YSOA, DOT, U><<<grid, SPMM_THREADS, smem, stream>>>(a);
    LS_LAUNCH_CHECK();
    return LS_OK;
}
}  // namespace

#define LS_SPMM_DISPATCH(FN, ...)                                                                         \
    do {                                                                                                  \
        if (solver_layout) { /* x = solver p rows (1/2/4 floats), y = SoA planes, dot epilogue */          \
            switch (K) {                                                                                  \
                case 1: return FN<1, false, true, true, 8>(__VA_ARGS__);                                  \
                case 2: return FN<2, false, true, true, 8>(__VA_ARGS__);                                  \
                case 3: return FN<3, true, true, true, 8>(__VA_ARGS__);                                   \
                case 4: return FN<4, true, true, true, 8>(__VA_ARGS__);                                   \
            }                                                                                             \
        } else {             /* public: x, y = (V,k) row-major */                                          \
            switch (K) {                                                                                  \
                case 1: return FN<1, false, false, false, 8>(__VA_ARGS__);                                \
                case 2: return FN<2, false, false, false, 8>(__VA_ARGS__);                                \
                case 3: return FN<3, false, false, false, 8>(__VA_ARGS__);                                \
                case 4: return FN<4, false, false, false, 8>(__VA_ARGS__);                                \
            }                                                                                             \
        }                                                                                                 \
        ls_set_error("SpMM: k=%d out of range [1,4]", K);                                                 \
        return LS_ERR_BAD_ARG;                                                                            \
    } while (0)

int spmm_prepare(int K, bool solver_layout, const SpmmCfg &cfg, int *ctas_per_sm) {
    LS_SPMM_DISPATCH(prepare_t, cfg, ctas_per_sm);
}

int spmm_launch(int K, bool solver_layout, const SpmmCfg &cfg, const SpmmArgs &a, int grid, cudaStream_t stream) {
    LS_SPMM_DISPATCH(launch_t, a, grid, stream);
}

int spmm_grid_for(int64_t V, int sm_count, int occ) {
    int64_t g = (V + 63) / 64;
    int64_t cap = (int64_t)sm_count * occ;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

int spmm_plan(const int *rowptr, const int *part, int G, int cap, int4 *desc, int *desc_cnt, int *overflow,
              cudaStream_t stream) {
    spmm_plan_kernel<<<(G + 127) / 128, 128, 0, stream>>>(rowptr, part, G, cap, desc, desc_cnt, overflow);
    LS_LAUNCH_CHECK();
    return LS_OK;
}

}  // namespace lsk

extern "C" int ls_spmm_csr_f32(int64_t V, const int32_t *rowptr, const int32_t *col, const float *val, const float *x,
                               int64_t ldx, float *y, int64_t ldy, int k, void *stream_) {
    using namespace lsk;
    cudaStream_t stream = (cudaStream_t)stream_;
    LS_REQUIRE(V >= 0 && V < (int64_t)0x7ffffff0, "V out of range");
    LS_REQUIRE(k >= 1, "k must be >= 1");
    LS_REQUIRE(ldx >= k && ldy >= k, "leading dimension smaller than k");
    if (V == 0) return LS_OK;
    LS_REQUIRE(rowptr && col && val && x && y, "NULL pointer");
    LS_REQUIRE((((uintptr_t)rowptr | (uintptr_t)col | (uintptr_t)val) & 15) == 0,
               "rowptr/col/val must be 16-byte aligned (TMA bulk copies)");
    LsDevInfo di;
    int rc = ls_dev_info(&di);
    if (rc) return rc;
    SpmmCfg cfg;
    spmm_config(&cfg);
    for (int k0 = 0; k0 < k; k0 += 4) {
        const int kk = (k - k0) < 4 ? (k - k0) : 4;
        SpmmArgs a{};
        a.V = (int)V;
        a.stages = cfg.stages;
        a.cap = cfg.cap;
        a.hint = 0;
        a.rowptr = rowptr;
        a.col = col;
        a.val = val;
        a.x = x + k0;
        a.y = y + k0;
        a.ldx = ldx;
        a.ldy = ldy;
        int occ = 1;
        rc = spmm_prepare(kk, false, cfg, &occ);
        if (rc) return rc;
        rc = spmm_launch(kk, false, cfg, a, spmm_grid_for(V, di.sm_count, occ), stream);
        if (rc) return rc;
    }
    return LS_OK;
}
