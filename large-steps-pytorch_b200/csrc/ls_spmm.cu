// ls_spmm.cu -- public y = A x entry point (AoS (V,k) torch layout) on top of the TMA-staged SpMM kernel.
// Replaces the torch sparse `M @ v` of parameterize.py:30 (to_differential) and scripts/main.py:192-195.
#include <stdlib.h>
#include "ls_spmm_kernel.cuh"

namespace lsk {

// runtime tuning knobs (defaults chosen on B200; override with LS_SPMM_STAGES / LS_SPMM_CAPMUL for sweeps)
void spmm_config(int *stages, int *cap) {
    static int s_stages = 0, s_cap = 0;
    if (s_stages == 0) {
        const char *e1 = getenv("LS_SPMM_STAGES");
        const char *e2 = getenv("LS_SPMM_CAPMUL");
        int st = e1 ? atoi(e1) : 2;
        int cm = e2 ? atoi(e2) : 10;
        if (st < 2) st = 2;
        if (st > SPMM_MAX_STAGES) st = SPMM_MAX_STAGES;
        if (cm < 2) cm = 2;
        if (cm > 24) cm = 24;
        s_stages = st;
        s_cap = SPMM_NT * cm;
    }
    *stages = s_stages;
    *cap = s_cap;
}

template <int K, bool SOA, bool DOT>
int spmm_prepare(int stages, int cap, int *ctas_per_sm) {
    static thread_local int cached_dev = -1, cached_occ = 0, cached_st = 0, cached_cap = 0;
    LsDevInfo di;
    int rc = ls_dev_info(&di);
    if (rc) return rc;
    if (cached_dev != di.device || cached_st != stages || cached_cap != cap) {
        size_t smem = spmm_smem_bytes(stages, cap);
        if ((int)smem > di.max_smem_optin) {
            ls_set_error("SpMM stage configuration needs %zu bytes of shared memory, device allows %d", smem, di.max_smem_optin);
            return LS_ERR_UNSUPPORTED;
        }
        LS_CUDA_TRY(cudaFuncSetAttribute(spmm_tma_kernel<K, SOA, DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        LS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmm_tma_kernel<K, SOA, DOT>, SPMM_THREADS, smem));
        if (occ < 1) occ = 1;
        cached_dev = di.device;
        cached_occ = occ;
        cached_st = stages;
        cached_cap = cap;
    }
    *ctas_per_sm = cached_occ;
    return LS_OK;
}

template <int K, bool SOA, bool DOT>
int spmm_launch(const SpmmArgs &a, int grid, cudaStream_t stream) {
    size_t smem = spmm_smem_bytes(a.stages, a.cap);
    spmm_tma_kernel<K, SOA, DOT><<<grid, SPMM_THREADS, smem, stream>>>(a);
    LS_LAUNCH_CHECK();
    return LS_OK;
}

int spmm_grid_for(int64_t V, int sm_count, int occ) {
    int64_t g = (V + 63) / 64;
    int64_t cap = (int64_t)sm_count * occ;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// explicit instantiations used by ls_pcg.cu
template int spmm_prepare<1, true, true>(int, int, int *);
template int spmm_prepare<2, true, true>(int, int, int *);
template int spmm_prepare<3, true, true>(int, int, int *);
template int spmm_prepare<4, true, true>(int, int, int *);
template int spmm_launch<1, true, true>(const SpmmArgs &, int, cudaStream_t);
template int spmm_launch<2, true, true>(const SpmmArgs &, int, cudaStream_t);
template int spmm_launch<3, true, true>(const SpmmArgs &, int, cudaStream_t);
template int spmm_launch<4, true, true>(const SpmmArgs &, int, cudaStream_t);

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
    int stages, cap;
    spmm_config(&stages, &cap);
    for (int k0 = 0; k0 < k; k0 += 4) {
        int kk = (k - k0) < 4 ? (k - k0) : 4;
        SpmmArgs a{};
        a.V = (int)V;
        a.stages = stages;
        a.cap = cap;
        a.rowptr = rowptr;
        a.col = col;
        a.val = val;
        a.x = x + k0;
        a.y = y + k0;
        a.ldx = ldx;
        a.ldy = ldy;
        int occ = 1;
        switch (kk) {
            case 1: rc = spmm_prepare<1, false, false>(stages, cap, &occ); break;
            case 2: rc = spmm_prepare<2, false, false>(stages, cap, &occ); break;
            case 3: rc = spmm_prepare<3, false, false>(stages, cap, &occ); break;
            default: rc = spmm_prepare<4, false, false>(stages, cap, &occ); break;
        }
        if (rc) return rc;
        int grid = spmm_grid_for(V, di.sm_count, occ);
        switch (kk) {
            case 1: rc = spmm_launch<1, false, false>(a, grid, stream); break;
            case 2: rc = spmm_launch<2, false, false>(a, grid, stream); break;
            case 3: rc = spmm_launch<3, false, false>(a, grid, stream); break;
            default: rc = spmm_launch<4, false, false>(a, grid, stream); break;
        }
        if (rc) return rc;
    }
    return LS_OK;
}
