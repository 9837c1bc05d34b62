// ls_spmm_kernel.cuh -- the CSR SpMM kernel of the library (sm_100a): y = A x for K right-hand-side columns.
//
// Roofline: HBM.  Algorithmic bytes per launch = 8 nnz + 4 (V+1) + 8 K V  (SURVEY.md section 8 d).
//
// Design (rows have ~7 non-zeros on a triangle mesh, so neither warp-per-row nor naive thread-per-row coalesces):
//   * persistent CTAs, each owning a contiguous, nnz-balanced range of rows (`part`);
//   * a producer warp streams, per block of <= NT rows, the *contiguous* col/val/rowptr slices of that block into a
//     shared-memory ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx) -- 2/3 of all bytes of the
//     kernel never touch the LSU or the register file on their way in;
//   * NT consumer threads take one row each out of shared memory (odd stride -> conflict-free), gather x through
//     L1 (neighbouring rows share neighbours, so the gather is mostly L1/L2 hits), and write y coalesced;
//   * optional epilogue: the K dot products x_k . y_k (p.Ap of CG) reduced deterministically across the grid.
//
// Layouts: SOA = K planes of `ld` floats (plane k at base + k*ld; the solver's internal layout);
//          AOS = (V, K) row-major with leading dimension `ld` (the public torch layout).
#pragma once
#include "ls_common.cuh"

namespace lsk {

constexpr int SPMM_NT = 256;            // consumer threads = max rows per block
constexpr int SPMM_THREADS = SPMM_NT + 32;
constexpr int SPMM_MAX_STAGES = 4;
constexpr int SPMM_RP = SPMM_NT + 8;    // rowptr ints per stage
constexpr int SPMM_HDR_BYTES = 64 + 1088;   // barriers + reduction scratch

struct SpmmArgs {
    int V;
    int stages;             // 2..4
    int cap;                // col/val elements per stage, multiple of 4
    const int *rowptr;
    const int *col;
    const float *val;
    const float *x;
    float *y;
    long long ldx, ldy;
    const int *part;        // [gridDim.x + 1] row boundaries, or NULL for an even split
    const int *done;        // optional early-exit flag (device), NULL if unused
    double *partials;       // [gridDim.x * K]   (DOT only)
    unsigned int *ticket;   //                   (DOT only)
    double *dot_out;        // [K]               (DOT only)
};

inline size_t spmm_stage_bytes(int cap) { return 16 + (size_t)SPMM_RP * 4 + (size_t)cap * 8; }
inline size_t spmm_smem_bytes(int stages, int cap) { return SPMM_HDR_BYTES + (size_t)stages * spmm_stage_bytes(cap); }

template <int K, bool SOA>
__device__ __forceinline__ void load_x(const float *__restrict__ x, long long ld, int c, float (&v)[K]) {
    if (SOA) {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = __ldg(x + (size_t)k * ld + c);
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = __ldg(x + (size_t)c * ld + k);
    }
}

template <int K, bool SOA, bool DOT>
__global__ void __launch_bounds__(SPMM_THREADS) spmm_tma_kernel(const SpmmArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.done != nullptr && *reinterpret_cast<const volatile int *>(a.done) != 0) return;

    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + SPMM_MAX_STAGES;
    double *red = reinterpret_cast<double *>(smem_raw + 64);
    unsigned char *stage_base = smem_raw + SPMM_HDR_BYTES;
    const int stages = a.stages, cap = a.cap;
    const size_t stage_bytes = 16 + (size_t)SPMM_RP * 4 + (size_t)cap * 8;

    const int tid = threadIdx.x;
    const int G = gridDim.x, cta = blockIdx.x;
    int r_begin, r_end;
    if (a.part) {
        r_begin = a.part[cta];
        r_end = a.part[cta + 1];
    } else {
        r_begin = (int)((long long)a.V * cta / G);
        r_end = (int)((long long)a.V * (cta + 1) / G);
    }

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            ls_mbar_init(&full[s], 1);
            ls_mbar_init(&empty[s], SPMM_NT / 32);
        }
        ls_fence_mbar_init();
    }
    __syncthreads();

    if (tid >= SPMM_NT) {
        // ===================== producer warp (one elected lane) =====================
        if (tid == SPMM_NT && r_begin < r_end) {
            int r = r_begin;
            int s_nz = a.rowptr[r];
            int st = 0, use = 0;
            while (r < r_end) {
                unsigned char *S = stage_base + (size_t)st * stage_bytes;
                int *hdr = reinterpret_cast<int *>(S);
                int *s_rp = reinterpret_cast<int *>(S + 16);
                int *s_col = s_rp + SPMM_RP;
                float *s_val = reinterpret_cast<float *>(s_col + cap);
                int nr = min(SPMM_NT, r_end - r);
                int e_nz = a.rowptr[r + nr];
                int direct = 0;
                const int s_a = s_nz & ~3;
                if (e_nz - s_a > cap - 4) {
                    // block does not fit the stage: largest nr whose slice fits (binary search on rowptr)
                    int lo = 0, hi = nr;   // invariant: `lo` rows fit
                    while (lo < hi) {
                        int mid = (lo + hi + 1) >> 1;
                        if (a.rowptr[r + mid] - s_a <= cap - 4) lo = mid;
                        else hi = mid - 1;
                    }
                    if (lo == 0) {   // a single row longer than a stage: its consumer reads it straight from global
                        nr = 1;
                        direct = 1;
                    } else {
                        nr = lo;
                    }
                    e_nz = a.rowptr[r + nr];
                }
                if (use > 0) ls_mbar_wait(&empty[st], (use - 1) & 1);
                const int e_a = (e_nz + 3) & ~3;
                const int r_a = r & ~3;
                const int rp_n = ((r + nr + 1 + 3) & ~3) - r_a;
                hdr[0] = r;
                hdr[1] = nr;
                hdr[2] = s_a;
                hdr[3] = direct;
                const uint32_t nbytes_cv = direct ? 0u : (uint32_t)(e_a - s_a) * 4u;
                const uint32_t nbytes_rp = (uint32_t)rp_n * 4u;
                ls_mbar_expect_tx(&full[st], nbytes_rp + 2u * nbytes_cv);
                ls_bulk_g2s(s_rp, a.rowptr + r_a, nbytes_rp, &full[st]);
                if (nbytes_cv) {
                    ls_bulk_g2s(s_col, a.col + s_a, nbytes_cv, &full[st]);
                    ls_bulk_g2s(s_val, a.val + s_a, nbytes_cv, &full[st]);
                }
                r += nr;
                s_nz = e_nz;
                if (++st == stages) {
                    st = 0;
                    ++use;
                }
            }
        }
    } else {
        // ===================== consumers: one row per thread =====================
        double dacc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) dacc[k] = 0.0;
        int r = r_begin;
        int st = 0, use = 0;
        while (r < r_end) {
            const unsigned char *S = stage_base + (size_t)st * stage_bytes;
            const int *hdr = reinterpret_cast<const int *>(S);
            const int *s_rp = reinterpret_cast<const int *>(S + 16);
            const int *s_col = s_rp + SPMM_RP;
            const float *s_val = reinterpret_cast<const float *>(s_col + cap);
            ls_mbar_wait(&full[st], use & 1);
            const int r0 = hdr[0], nr = hdr[1], s_a = hdr[2], direct = hdr[3];
            if (tid < nr) {
                const int row = r0 + tid;
                const int off = (r0 & 3) + tid;
                const int j0 = s_rp[off], j1 = s_rp[off + 1];
                float acc[K];
#pragma unroll
                for (int k = 0; k < K; ++k) acc[k] = 0.f;
                if (!direct) {
                    const int *sc = s_col + (j0 - s_a);
                    const float *sv = s_val + (j0 - s_a);
                    const int len = j1 - j0;
                    int j = 0;
                    for (; j + 4 <= len; j += 4) {
                        int c[4];
                        float w[4];
                        float xv[4][K];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            c[u] = sc[j + u];
                            w[u] = sv[j + u];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) load_x<K, SOA>(a.x, a.ldx, c[u], xv[u]);
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int k = 0; k < K; ++k) acc[k] = fmaf(w[u], xv[u][k], acc[k]);
                    }
                    if (j < len) {   // 1..3 left: predicated (zero weight, own row as a safe address)
                        int c[3];
                        float w[3];
                        float xv[3][K];
#pragma unroll
                        for (int u = 0; u < 3; ++u) {
                            const bool ok = (j + u) < len;
                            c[u] = ok ? sc[j + u] : row;
                            w[u] = ok ? sv[j + u] : 0.f;
                        }
#pragma unroll
                        for (int u = 0; u < 3; ++u) load_x<K, SOA>(a.x, a.ldx, c[u], xv[u]);
#pragma unroll
                        for (int u = 0; u < 3; ++u)
#pragma unroll
                            for (int k = 0; k < K; ++k) acc[k] = fmaf(w[u], xv[u][k], acc[k]);
                    }
                } else {
                    for (int j = j0; j < j1; ++j) {
                        const int c = __ldg(a.col + j);
                        const float w = __ldg(a.val + j);
                        float xv[K];
                        load_x<K, SOA>(a.x, a.ldx, c, xv);
#pragma unroll
                        for (int k = 0; k < K; ++k) acc[k] = fmaf(w, xv[k], acc[k]);
                    }
                }
                if (SOA) {
#pragma unroll
                    for (int k = 0; k < K; ++k) a.y[(size_t)k * a.ldy + row] = acc[k];
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) a.y[(size_t)row * a.ldy + k] = acc[k];
                }
                if (DOT) {
                    float xr[K];
                    load_x<K, SOA>(a.x, a.ldx, row, xr);
#pragma unroll
                    for (int k = 0; k < K; ++k) dacc[k] += (double)xr[k] * (double)acc[k];
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) ls_mbar_arrive(&empty[st]);
            r = r0 + nr;
            if (++st == stages) {
                st = 0;
                ++use;
            }
        }
        if (DOT) {
            double tot[K];
            const bool last = ls_grid_reduce<K>(dacc, tot, a.partials, a.ticket, red, tid, SPMM_NT, 1, cta, G);
            if (last && tid == 0) {
#pragma unroll
                for (int k = 0; k < K; ++k) a.dot_out[k] = tot[k];
            }
        }
    }
}

}  // namespace lsk
