// ls_pcg.cu -- preconditioned conjugate gradients for M X = B, all K columns in one pass (sm_100a):
// handle / workspace management, solver configuration, the graph-mode solver and the C entry points.
//
// Replaces the reference's solve plug-ins (largesteps/solvers.py:26-39 CholeskySolver -> cholespy/CHOLMOD,
// solvers.py:41-126 ConjugateGradientSolver -> ~12 eager torch kernels + 1 host sync per iteration per axis).
//
// Three execution modes share the handle, the data layout and the per-column arithmetic:
//   * fused (ls_pcg_fused.cuh): the whole solve is ONE kernel with two grid synchronisations per iteration, in-kernel warm
//     start and true-residual guard, optional Chebyshev polynomial preconditioner -- the default for every k in 1..4;
//     cooperative grid (one CTA per SM), one CTA for tiny meshes, or (opt-in) one thread-block cluster;
//   * classic (ls_pcg_persistent.cuh): round 1's three-synchronisation persistent kernel (LS_PCG_ALGO=classic, A/B and fallback);
//   * graph (this file): one iteration = three kernels, a CUDA graph of CHUNK iterations replayed until a device-side
//     `done` flag is seen -- the fallback when no cooperative launch is possible:
//       K1  Ap = A p, pAp_k = p_k.Ap_k                     (SELL-32 or TMA-staged CSR SpMM + deterministic grid reduction)
//       K2  x += a p; r -= a Ap; rz' = r.(dinv r); rr = r.r (fused update + 2K dot products; last CTA does the
//           scalar state transition: beta, convergence per column, iteration count, done flag)
//       K3  p = dinv r + beta p
//
// Data layout in HBM (all inside the caller-provided workspace):
//   CSR copy  rowptr (V+1) int32, col (nnz) int32, val (nnz) fp32, padded so 16-byte TMA granules never leave it;
//             optionally re-ordered P A P^T (Morton order of the vertices, kept only if it gathers more coherently)
//   SELL-32   soff (V/32+1), ent (padded nnz) int2 {col, val}: the fast SpMM engine's copy
//   dinv      Vp fp32                      Jacobi 1/diag (0 in the padding)
//   x r Ap    K planes of Vp fp32 each     SoA: plane k holds column k; Vp = V rounded up to 32 (zero padded)
//   p         Vp rows of PW floats         PW = 1, 2, 4 for K = 1, 2, 3|4: a gather of p[col] is one load
//   ctrl      PcgCtrl                      device-resident scalars: the iteration never returns to the host for them
// Columns carry their own alpha/beta and freeze independently when ||r_k|| <= rtol ||b_k||, which is exactly the
// reference's "one CG per axis" (solvers.py:115-118) run in lock-step.  Dot products accumulate in fp64.
#include <new>
#include <string.h>
#include <stdlib.h>
#include "ls_spmm_host.h"
#include "ls_sell_kernel.cuh"
#include "ls_pcg_persistent.cuh"
#include "ls_pcg_fused.cuh"
#include "ls_fused_inst.h"

#ifndef LS_CLRES_DEFAULT
#define LS_CLRES_DEFAULT 0     // slices (x 32 vertices); 0 = off: measured slower than the cooperative grid, see clres_limit()
#endif

namespace {

constexpr int KMAX = 4;
constexpr int VEC_THREADS = 256;
constexpr int CHUNK = 8;   // CG iterations per graph launch

struct PcgCtrl {
    double rz[KMAX], pAp[KMAX], rr[KMAX], bb[KMAX];
    float beta[KMAX];
    float rtol2;
    int maxit;
    int it;
    int done;        // 0 running, 1 converged, 2 maxit reached, 3 breakdown
    int conv[KMAX];  // column frozen
    int k;
    int restart;     // warm start was worse than a cold start for some column: redo the initialisation from x = 0
};

struct PcgHandle {
    int64_t V, nnz, Vp;
    int k_max, precond;
    int device;
    int sm_count;
    // workspace carve-out (device)
    int *rowptr, *col;
    float *val, *dinv;
    float *x, *r, *p, *Ap;
    int *part;
    PcgCtrl *ctrl;
    double *part_spmm, *part_vec;
    unsigned int *tickets;   // [0] spmm, [1] vec
    float *info;
    int *flags;
    // launch geometry
    lsk::SpmmCfg cfg;
    int spmm_grid;
    int vec_grid;
    int4 *desc;
    int *desc_cnt;
    int *perm;       // new -> old row (NULL-equivalent when has_perm == 0)
    int *inv;        // old -> new
    int *scan;
    int has_perm;
    int planned;
    // SELL-32 engine (fast path)
    int *soff;
    int2 *ent;
    long long sell_cap;      // capacity of `ent` in entries
    long long sell_entries;  // padded entry count
    int nslices;
    int sell_on;
    int sell_grid;
    // pattern-only copy for matrices with one common off-diagonal value (ls_sell_kernel.cuh "PAT"; LS_PCG_PATTERN=0 switches it off)
    int *poff;
    int2 *pcol;
    float *diagp;
    unsigned int *patmm;     // [min, max] of the off-diagonal value bits
    long long pat_cap;       // capacity of `pcol` in pairs
    float offc;
    int pat_on;
    // persistent single-kernel solve (K = 3; warm starts enter it in resume mode)
    lsp::GridBar *gbar;
    double *part_persist;
    long long *dbg;
    unsigned long long *ring;
    int ring_slots;
    int persist_on, persist_grid, persist_res, persist_nsl_max, persist_threads;
    size_t persist_smem;
    // fused two-synchronisation solver (ls_pcg_fused.cuh): the default; one configuration for K = 3 (k = 1..3) and one for K = 4
    float *pv;               // owner copy of p, k_max planes
    float *z2, *cy, *cd;     // Chebyshev preconditioner: second published row buffer, iterate and direction planes
    int cheb_m;              // 0 / 1: Jacobi only; m >= 2: polynomial of degree m - 1 (precond = 2)
    float cheb_c0, cheb_c1[8], cheb_c2[8];
    float *gersh;            // [1] max_i sum_j |a_ij| / a_ii
    struct FusedCfg {
        int on, grid, res, nw, sync, cluster, nsl_max, pat, dp;
        size_t smem;
        const void *fn, *fn_prof;
    } fused[2];
    int max_smem_optin;
    int refine;              // max restarts from the true residual per solve
    float theta;
    int sell_tma;            // stand-alone SpMM: TMA-staged variant (0 = register-prefetch kernel)
    int sell_pf;             // ... halo (rows) of its bulk L2 prefetch of the gathered vector, 0 = off
    // graphs, one per K
    cudaGraphExec_t graph[KMAX + 1];
    cudaStream_t cap_stream;
    int *pinned_done;        // 2 ints, host pinned
    cudaEvent_t ev[2];
    size_t ws_bytes;
};

static bool want_pattern() {   // on unless LS_PCG_PATTERN=0 (A/B switch)
    const char *e = getenv("LS_PCG_PATTERN");
    return !(e && e[0] == '0');
}

struct Carve {
    size_t off = 0;
    size_t take(size_t bytes) {
        size_t o = off;
        off = ls_align_up(off + bytes, 256);
        return o;
    }
};

size_t carve_handle(PcgHandle *h, char *base, int64_t V, int64_t nnz, int k_max, int grid_cap) {
    Carve c;
    int64_t Vp = (V + 31) / 32 * 32;
    size_t o_rp = c.take((size_t)(V + 1 + 8) * 4);
    size_t o_col = c.take((size_t)(nnz + 8) * 4);
    size_t o_val = c.take((size_t)(nnz + 8) * 4);
    size_t o_dinv = c.take((size_t)Vp * 4);
    const size_t k_rows = k_max < 4 ? 4 : k_max;   // x and the owner's p may be stored as rows of 4 floats (fused solver, RES = 1)
    size_t o_x = c.take((size_t)Vp * 4 * k_rows);
    size_t o_r = c.take((size_t)Vp * 4 * k_max);
    size_t o_p = c.take((size_t)Vp * 4 * 4);            // p: rows of PW <= 4 floats
    size_t o_Ap = c.take((size_t)Vp * 4 * k_max);
    size_t o_pown = c.take((size_t)Vp * 4 * k_rows);
    size_t o_z2 = c.take((size_t)Vp * 4 * 4);
    size_t o_cy = c.take((size_t)Vp * 4 * k_max);
    size_t o_cd = c.take((size_t)Vp * 4 * k_max);
    size_t o_part = c.take((size_t)(grid_cap + 1) * 4);
    size_t o_desc = c.take((size_t)grid_cap * lsk::SPMM_BMAX * sizeof(int4));
    size_t o_dcnt = c.take((size_t)grid_cap * 4);
    const long long sell_cap = (long long)nnz + nnz / 2 + 32768;
    size_t o_soff = c.take((size_t)(Vp / 32 + 2) * 4);
    size_t o_ent = c.take((size_t)sell_cap * 8);
    size_t o_perm = c.take((size_t)(V + 8) * 4);
    size_t o_inv = c.take((size_t)(V + 8) * 4);
    size_t o_scan = c.take(ls_scan_scratch_elems(V + 1) * 4);
    size_t o_ctrl = c.take(sizeof(PcgCtrl));
    size_t o_ps = c.take((size_t)grid_cap * KMAX * 8);
    size_t o_pv = c.take((size_t)grid_cap * 3 * KMAX * 8);
    size_t o_gbar = c.take(64);
    size_t o_pp = c.take((size_t)2 * lsf::NVMAX * 256 * 8);
    size_t o_dbg = c.take((size_t)(8 + 8 * 256) * 8);
    constexpr int RING_SLOTS = 32768;                 // fast all-reduce slots (64 B each): 2 per iteration
    size_t o_ring = c.take((size_t)RING_SLOTS * 64);
    size_t o_tk = c.take(64);
    size_t o_info = c.take(64);
    size_t o_flags = c.take(64);
    // pattern-only copy (always carved: the workspace size must not depend on the environment)
    size_t o_poff = c.take((size_t)(Vp / 32 + 2) * 4);
    size_t o_pcol = c.take((size_t)sell_cap * 4);         // sell_cap / 2 pairs of 8 bytes
    size_t o_diagp = c.take((size_t)Vp * 4);
    size_t o_patmm = c.take(64);
    size_t o_gersh = c.take(64);
    if (h && base) {
        h->Vp = Vp;
        h->rowptr = (int *)(base + o_rp);
        h->col = (int *)(base + o_col);
        h->val = (float *)(base + o_val);
        h->dinv = (float *)(base + o_dinv);
        h->x = (float *)(base + o_x);
        h->r = (float *)(base + o_r);
        h->p = (float *)(base + o_p);
        h->Ap = (float *)(base + o_Ap);
        h->pv = (float *)(base + o_pown);
        h->z2 = (float *)(base + o_z2);
        h->cy = (float *)(base + o_cy);
        h->cd = (float *)(base + o_cd);
        h->part = (int *)(base + o_part);
        h->desc = (int4 *)(base + o_desc);
        h->desc_cnt = (int *)(base + o_dcnt);
        h->soff = (int *)(base + o_soff);
        h->ent = (int2 *)(base + o_ent);
        h->sell_cap = sell_cap;
        h->nslices = (int)(Vp / 32);
        h->perm = (int *)(base + o_perm);
        h->inv = (int *)(base + o_inv);
        h->scan = (int *)(base + o_scan);
        h->ctrl = (PcgCtrl *)(base + o_ctrl);
        h->part_spmm = (double *)(base + o_ps);
        h->part_vec = (double *)(base + o_pv);
        h->gbar = (lsp::GridBar *)(base + o_gbar);
        h->part_persist = (double *)(base + o_pp);
        h->dbg = (long long *)(base + o_dbg);
        h->ring = (unsigned long long *)(base + o_ring);
        h->ring_slots = RING_SLOTS;
        h->tickets = (unsigned int *)(base + o_tk);
        h->info = (float *)(base + o_info);
        h->flags = (int *)(base + o_flags);
        h->poff = (int *)(base + o_poff);
        h->pcol = (int2 *)(base + o_pcol);
        h->diagp = (float *)(base + o_diagp);
        h->patmm = (unsigned int *)(base + o_patmm);
        h->gersh = (float *)(base + o_gersh);
        h->pat_cap = sell_cap / 2;
    }
    return c.off;
}

constexpr int GRID_CAP = 148 * 8 * 2;   // upper bound on any persistent grid we launch (workspace sizing)

// ---- setup kernels --------------------------------------------------------------------------------
__global__ void k_pad_tail(int *rowptr, int *col, float *val, int64_t V, int64_t nnz) {
    int t = threadIdx.x;
    if (t < 8) {
        rowptr[V + 1 + t] = (int)nnz;
        col[nnz + t] = 0;
        val[nnz + t] = 0.f;
    }
}

__global__ void k_dinv(int64_t V, int64_t Vp, const int *__restrict__ rowptr, const int *__restrict__ col,
                       const float *__restrict__ val, int precond, float *__restrict__ dinv, int *__restrict__ flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Vp) return;
    if (i >= V) {
        dinv[i] = 0.f;
        return;
    }
    float d = 0.f;
    bool found = false;
    int s = rowptr[i], e = rowptr[i + 1];
    if (e < s) atomicOr(flags, 4);
    for (int j = s; j < e; ++j) {
        int c = col[j];
        if (c < 0 || c >= V) atomicOr(flags, 1);
        if (c == (int)i) {
            d += val[j];
            found = true;
        }
    }
    if (!found || !(d > 0.f)) atomicOr(flags, 2);
    dinv[i] = precond ? (1.0f / d) : 1.0f;
}

// Gershgorin bound of lambda_max(D^-1 A): max_i sum_j |a_ij| / a_ii   (positive floats order like their bit patterns)
__global__ void k_gershgorin(int64_t V, const int *__restrict__ rowptr, const int *__restrict__ col, const float *__restrict__ val,
                             float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float g = 0.f;
    if (i < V) {
        float d = 0.f, sabs = 0.f;
        for (int j = rowptr[i]; j < rowptr[i + 1]; ++j) {
            const float a = val[j];
            sabs += fabsf(a);
            if (col[j] == (int)i) d += a;
        }
        g = d > 0.f ? sabs / d : 0.f;
    }
    g = fmaxf(g, 0.f);
    unsigned int b = __float_as_uint(g);
    b = __reduce_max_sync(0xffffffffu, b);
    if ((threadIdx.x & 31) == 0 && b) atomicMax(reinterpret_cast<unsigned int *>(out), b);
}

// ---- permuted copy  A' = P A P^T  (perm[new] = old) ------------------------------------------------
__global__ void k_perm_inv_len(int64_t V, const int *__restrict__ perm, const int *__restrict__ rowptr,
                               int *__restrict__ inv, int *__restrict__ len, int *__restrict__ flags) {
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= V) return;
    const int o = perm[n];
    if (o < 0 || o >= V) {
        atomicOr(flags, 8);
        len[n] = 0;
        return;
    }
    inv[o] = (int)n;
    len[n] = rowptr[o + 1] - rowptr[o];
}
// one thread per new row: copy the old row with renumbered columns, then insertion-sort it by new column
__global__ void k_perm_rows(int64_t V, const int *__restrict__ perm, const int *__restrict__ inv,
                            const int *__restrict__ rowptr, const int *__restrict__ col, const float *__restrict__ val,
                            const int *__restrict__ rowptr_new, int *__restrict__ col_new, float *__restrict__ val_new,
                            int *__restrict__ flags) {
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= V) return;
    const int o = perm[n];
    if (o < 0 || o >= V) return;
    const int s = rowptr[o], e = rowptr[o + 1];
    const int d = rowptr_new[n];
    for (int j = s; j < e; ++j) {
        int c = col[j];
        if (c < 0 || c >= V) {
            atomicOr(flags, 1);
            c = o;
        }
        const int cn = inv[c];
        const float w = val[j];
        int a = d + (j - s) - 1;
        while (a >= d && col_new[a] > cn) {
            col_new[a + 1] = col_new[a];
            val_new[a + 1] = val_new[a];
            --a;
        }
        col_new[a + 1] = cn;
        val_new[a + 1] = w;
    }
}
__global__ void k_perm_check(int64_t V, const int *__restrict__ perm, const int *__restrict__ inv, int *__restrict__ flags) {
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= V) return;
    const int o = perm[n];
    if (o >= 0 && o < V && inv[o] != (int)n) atomicOr(flags, 8);   // not a permutation (duplicate target)
}

// Gather-locality score of a row order: number of (row, slot) pairs whose column is NOT within 8 entries of the
// same slot's column in the previous row (adjacent rows are adjacent lanes of a warp, 8 float4 rows of p = one 128-byte
// line).  Lower is better; used to decide whether the Morton re-ordering actually helps (a row-major grid is already
// perfectly coalesced, a scanner mesh or a shuffled numbering is not).
__global__ void k_locality_score(int64_t V, const int *__restrict__ rowptr, const int *__restrict__ col,
                                 unsigned long long *__restrict__ score) {
    unsigned int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (int64_t)gridDim.x * blockDim.x) {
        if ((i & 31) == 0) continue;   // first lane of a warp has no left neighbour
        const int s = rowptr[i], e = rowptr[i + 1], sp = rowptr[i - 1], ep = rowptr[i];
        const int n = min(e - s, ep - sp);
        for (int j = 0; j < n; ++j) {
            const int d = col[s + j] - col[sp + j];
            bad += (d < -8 || d > 8) ? 1u : 0u;
        }
        bad += (unsigned int)((e - s) - n);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(score, (unsigned long long)bad);
}

// nnz-balanced contiguous row partition: part[c] = first row r with weight(r) >= c * total / G,
// weight(r) = 2 * rowptr[r] + 5 * r   (~ bytes/4 streamed per non-zero and per row)
__global__ void k_partition(int64_t V, const int *__restrict__ rowptr, int G, int *__restrict__ part) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > G) return;
    if (c == G) {
        part[c] = (int)V;
        return;
    }
    long long total = 2LL * rowptr[V] + 5LL * V;
    long long target = total * c / G;
    int64_t lo = 0, hi = V;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        long long w = 2LL * rowptr[mid] + 5LL * mid;
        if (w < target) lo = mid + 1;
        else hi = mid;
    }
    part[c] = (int)lo;
}

// ---- solve kernels --------------------------------------------------------------------------------
struct VecArgs {
    int64_t V, Vp;
    float *x, *r, *p, *Ap;
    const float *dinv;
    PcgCtrl *ctrl;
    double *partials;
    unsigned int *ticket;
    const int *perm;   // new -> old row of the caller's (V,K) arrays, or NULL
    int bench;         // timing harness: ignore the done flag, skip the state transition
};

// cold start: x = 0, r = b, p = z = dinv r;  warm (stage 2): r = b - Ap (Ap = A x0 from K1), p = z
template <int K, bool WARM>
__global__ void __launch_bounds__(VEC_THREADS) k_init(VecArgs a, const float *__restrict__ b, float rtol, int maxit,
                                                      int only_if_restart) {
    __shared__ double red[3 * K * 32 + 3 * K + 1];
    if (only_if_restart && *reinterpret_cast<volatile int *>(&a.ctrl->restart) == 0) return;
    double acc[3 * K];   // [rz | bb | rr]
#pragma unroll
    for (int i = 0; i < 3 * K; ++i) acc[i] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.V; i += (int64_t)gridDim.x * blockDim.x) {
        const float di = a.dinv[i];
        const int64_t io = a.perm ? a.perm[i] : i;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float bv = b[io * K + k];
            float rv = bv;
            if (WARM) rv = bv - a.Ap[(size_t)k * a.Vp + i];
            else a.x[(size_t)k * a.Vp + i] = 0.f;
            const float z = di * rv;
            a.r[(size_t)k * a.Vp + i] = rv;
            a.p[(size_t)i * lsk::PRow<K>::PW + k] = z;
            acc[k] += (double)rv * (double)z;
            acc[K + k] += (double)bv * (double)bv;
            acc[2 * K + k] += (double)rv * (double)rv;
        }
        if (K == 3) a.p[(size_t)i * 4 + 3] = 0.f;
    }
    double tot[3 * K];
    const bool last = ls_grid_reduce<3 * K>(acc, tot, a.partials, a.ticket, red, threadIdx.x, VEC_THREADS, 1,
                                            blockIdx.x, gridDim.x);
    if (last && threadIdx.x == 0) {
        PcgCtrl *c = a.ctrl;
        int all = 1, worse = 0;
        const double rtol2 = (double)rtol * (double)rtol;
        for (int k = 0; k < K; ++k) {
            // a warm start whose residual exceeds ||b|| is worse than x = 0 and, in fp32, caps the attainable
            // accuracy at eps * kappa * ||x0|| / ||x||: fall back to the cold start (the reference CG has no such
            // guard, solvers.py:107-110, and loses accuracy when the gradient scale changes between steps)
            if (WARM && tot[2 * K + k] > tot[K + k]) worse = 1;
            c->rz[k] = tot[k];
            c->bb[k] = tot[K + k];
            c->rr[k] = tot[2 * K + k];
            c->pAp[k] = 1.0;
            c->beta[k] = 0.f;
            const int cv = tot[2 * K + k] <= rtol2 * tot[K + k];   // b_k == 0, or the warm start is already good enough
            c->conv[k] = cv;
            all &= cv;
        }
        for (int k = K; k < KMAX; ++k) {
            c->conv[k] = 1;
            c->rr[k] = 0.0;
            c->bb[k] = 0.0;
        }
        c->rtol2 = rtol * rtol;
        c->maxit = maxit;
        c->it = 0;
        c->k = K;
        c->restart = worse;
        c->done = (all && !worse) ? 1 : 0;
    }
}

// warm start stage 1: x = x0 (AoS -> SoA), p = x0 (SpMM input), done = 0 so that K1 runs
template <int K>
__global__ void __launch_bounds__(VEC_THREADS) k_warm_load(VecArgs a, const float *__restrict__ x0) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.V; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t io = a.perm ? a.perm[i] : i;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float v = x0[io * K + k];
            a.x[(size_t)k * a.Vp + i] = v;
            a.p[(size_t)i * lsk::PRow<K>::PW + k] = v;
        }
        if (K == 3) a.p[(size_t)i * 4 + 3] = 0.f;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) a.ctrl->done = 0;
}

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

// p is stored as rows of PW floats (PW = 1, 2, 4 for K = 1, 2, 3|4) so that the SpMM gathers one row with one load.
// These helpers move the 4 rows 4*i4 .. 4*i4+3 between that layout and per-column float4 registers.
template <int K>
__device__ __forceinline__ void load_p_rows(const float *p, int64_t i4, float4 (&pv)[K]) {
    constexpr int PW = lsk::PRow<K>::PW;
    if (PW == 1) {
        pv[0] = ld4(p + 4 * i4);
    } else if (PW == 2) {
        const float4 a = ld4(p + 8 * i4), b = ld4(p + 8 * i4 + 4);     // rows (0,1) and (2,3)
        pv[0] = make_float4(a.x, a.z, b.x, b.z);
        if (K > 1) pv[K > 1 ? 1 : 0] = make_float4(a.y, a.w, b.y, b.w);
    } else {
        const float4 r0 = ld4(p + 16 * i4), r1 = ld4(p + 16 * i4 + 4), r2 = ld4(p + 16 * i4 + 8), r3 = ld4(p + 16 * i4 + 12);
        pv[0] = make_float4(r0.x, r1.x, r2.x, r3.x);
        if (K > 1) pv[K > 1 ? 1 : 0] = make_float4(r0.y, r1.y, r2.y, r3.y);
        if (K > 2) pv[K > 2 ? 2 : 0] = make_float4(r0.z, r1.z, r2.z, r3.z);
        if (K > 3) pv[K > 3 ? 3 : 0] = make_float4(r0.w, r1.w, r2.w, r3.w);
    }
}
template <int K>
__device__ __forceinline__ void store_p_rows(float *p, int64_t i4, const float4 (&pv)[K]) {
    constexpr int PW = lsk::PRow<K>::PW;
    if (PW == 1) {
        st4(p + 4 * i4, pv[0]);
    } else if (PW == 2) {
        const float4 &c0 = pv[0], &c1 = pv[K > 1 ? 1 : 0];
        st4(p + 8 * i4, make_float4(c0.x, c1.x, c0.y, c1.y));
        st4(p + 8 * i4 + 4, make_float4(c0.z, c1.z, c0.w, c1.w));
    } else {
        const float4 &c0 = pv[0], &c1 = pv[K > 1 ? 1 : 0], &c2 = pv[K > 2 ? 2 : 0];
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 &c3 = (K > 3) ? pv[K > 3 ? 3 : 0] : z;
        st4(p + 16 * i4, make_float4(c0.x, c1.x, c2.x, c3.x));
        st4(p + 16 * i4 + 4, make_float4(c0.y, c1.y, c2.y, c3.y));
        st4(p + 16 * i4 + 8, make_float4(c0.z, c1.z, c2.z, c3.z));
        st4(p + 16 * i4 + 12, make_float4(c0.w, c1.w, c2.w, c3.w));
    }
}

// scalar state transition run by the last CTA of K2: beta, per-column convergence, iteration count, done flag
template <int K>
__device__ __forceinline__ void pcg_transition(PcgCtrl *c, const double (&tot)[2 * K]) {
    int all = 1, bad = 0;
    for (int k = 0; k < K; ++k) {
        if (c->conv[k]) continue;
        const double pAp = c->pAp[k];
        if (!(pAp > 0.0) || !(tot[k] == tot[k])) bad = 1;   // not SPD, or NaN crept in
        const double rz_old = c->rz[k];
        c->beta[k] = (rz_old > 0.0) ? (float)(tot[k] / rz_old) : 0.f;
        c->rz[k] = tot[k];
        c->rr[k] = tot[K + k];
        const int cv = tot[K + k] <= (double)c->rtol2 * c->bb[k];
        c->conv[k] = cv;
        if (cv) c->beta[k] = 0.f;
        all &= cv;
    }
    const int it = c->it + 1;
    c->it = it;
    if (bad) c->done = 3;
    else if (all) c->done = 1;
    else if (it >= c->maxit) c->done = 2;
}

// K3: p = dinv r + beta p   (same loads-first structure as K2)
template <int K>
__global__ void __launch_bounds__(VEC_THREADS, 4) k_pupdate(VecArgs a) {
    PcgCtrl *c = a.ctrl;
    const int64_t n4 = a.Vp >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float4 d, pv[K], rv[K];
    auto load = [&](int64_t j) {
        d = ld4(a.dinv + 4 * j);
        load_p_rows<K>(a.p, j, pv);
#pragma unroll
        for (int k = 0; k < K; ++k) rv[k] = ld4(a.r + (size_t)k * a.Vp + 4 * j);
    };
    if (i < n4) load(i);
    if (!a.bench && *reinterpret_cast<volatile int *>(&c->done) != 0) return;
    float beta[K];
#pragma unroll
    for (int k = 0; k < K; ++k) beta[k] = c->beta[k];
    for (bool first = true; i < n4; i += stride, first = false) {
        if (!first) load(i);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float be = beta[k];
            pv[k].x = fmaf(be, pv[k].x, d.x * rv[k].x);
            pv[k].y = fmaf(be, pv[k].y, d.y * rv[k].y);
            pv[k].z = fmaf(be, pv[k].z, d.z * rv[k].z);
            pv[k].w = fmaf(be, pv[k].w, d.w * rv[k].w);
        }
        store_p_rows<K>(a.p, i, pv);
    }
}

// K2: x += alpha p, r -= alpha Ap, rz' = r.(dinv r), rr = r.r ; last CTA: scalar state transition.
// One float4 of rows per thread, one column at a time (4-5 float4 loads in flight, ~80 registers); the vector loads of
// the first column are issued BEFORE the dependent scalar chain (done flag -> pAp/rz -> fp64 divide) so it hides under them.
template <int K>
__global__ void __launch_bounds__(VEC_THREADS, 3) k_update_cs(VecArgs a) {
    __shared__ double red[2 * K * 32 + 2 * K + 1];
    PcgCtrl *c = a.ctrl;
    const int64_t n4 = a.Vp >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f), x0, r0, q0, pv[K];
    if (i0 < n4) {   // issued before the dependent scalar chain below
        d = ld4(a.dinv + 4 * i0);
        load_p_rows<K>(a.p, i0, pv);
        x0 = ld4(a.x + 4 * i0);
        r0 = ld4(a.r + 4 * i0);
        q0 = ld4(a.Ap + 4 * i0);
    }
    if (!a.bench && *reinterpret_cast<volatile int *>(&c->done) != 0) return;
    float alpha[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double pAp = c->pAp[k];
        alpha[k] = (c->conv[k] || !(pAp > 0.0)) ? 0.f : (float)(c->rz[k] / pAp);
    }
    double acc[2 * K];
#pragma unroll
    for (int q = 0; q < 2 * K; ++q) acc[q] = 0.0;
    for (int64_t i = i0; i < n4; i += stride) {
        if (i != i0) {
            d = ld4(a.dinv + 4 * i);
            load_p_rows<K>(a.p, i, pv);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const size_t o = (size_t)k * a.Vp + 4 * i;
            float4 xv, rv, qv;
            if (k == 0 && i == i0) {
                xv = x0; rv = r0; qv = q0;
            } else {
                xv = ld4(a.x + o); rv = ld4(a.r + o); qv = ld4(a.Ap + o);
            }
            const float al = alpha[k];
            const float4 pk = pv[k];
            xv.x = fmaf(al, pk.x, xv.x); xv.y = fmaf(al, pk.y, xv.y); xv.z = fmaf(al, pk.z, xv.z); xv.w = fmaf(al, pk.w, xv.w);
            rv.x = fmaf(-al, qv.x, rv.x); rv.y = fmaf(-al, qv.y, rv.y); rv.z = fmaf(-al, qv.z, rv.z); rv.w = fmaf(-al, qv.w, rv.w);
            st4(a.x + o, xv);
            st4(a.r + o, rv);
            const float r2x = rv.x * rv.x, r2y = rv.y * rv.y, r2z = rv.z * rv.z, r2w = rv.w * rv.w;
            acc[k] += (double)(d.x * r2x) + (double)(d.y * r2y) + (double)(d.z * r2z) + (double)(d.w * r2w);
            acc[K + k] += (double)r2x + (double)r2y + (double)r2z + (double)r2w;
        }
    }
    double tot[2 * K];
    const bool last = ls_grid_reduce<2 * K>(acc, tot, a.partials, a.ticket, red, threadIdx.x, VEC_THREADS, 1,
                                            blockIdx.x, gridDim.x);
    if (last && threadIdx.x == 0 && !a.bench) pcg_transition<K>(c, tot);
}

// x (SoA) -> out (AoS), info
template <int K>
__global__ void __launch_bounds__(VEC_THREADS) k_final(VecArgs a, float *__restrict__ out, float *__restrict__ info) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.V; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t io = a.perm ? a.perm[i] : i;
#pragma unroll
        for (int k = 0; k < K; ++k) out[io * K + k] = a.x[(size_t)k * a.Vp + i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const PcgCtrl *c = a.ctrl;
        float tmp[8];
        tmp[0] = (float)c->it;
        tmp[1] = (float)c->done;
        for (int k = 0; k < KMAX; ++k) tmp[2 + k] = (k < K && c->bb[k] > 0.0) ? (float)sqrt(c->rr[k] / c->bb[k]) : 0.f;
        tmp[6] = tmp[7] = 0.f;
        for (int j = 0; j < 8; ++j)
            if (info) info[j] = tmp[j];
    }
}

VecArgs vec_args(PcgHandle *h, int which_ticket) {
    VecArgs a;
    a.V = h->V;
    a.Vp = h->Vp;
    a.x = h->x;
    a.r = h->r;
    a.p = h->p;
    a.Ap = h->Ap;
    a.dinv = h->dinv;
    a.ctrl = h->ctrl;
    a.partials = h->part_vec;
    a.ticket = h->tickets + which_ticket;
    a.perm = h->has_perm ? h->perm : nullptr;
    a.bench = 0;
    return a;
}

lsk::SpmmArgs spmm_args(PcgHandle *h, int K, bool with_done) {
    lsk::SpmmArgs s{};
    s.V = (int)h->V;
    s.stages = h->cfg.stages;
    s.cap = h->cfg.cap;
    s.hint = h->cfg.hint;
    s.debug = h->cfg.debug;
    s.desc = h->planned ? h->desc : nullptr;
    s.desc_cnt = h->desc_cnt;
    s.rowptr = h->rowptr;
    s.col = h->col;
    s.val = h->val;
    s.x = h->p;
    s.y = h->Ap;
    s.ldx = (K == 1) ? 1 : (K == 2 ? 2 : 4);   // p rows
    s.ldy = h->Vp;
    s.part = h->part;
    s.done = with_done ? &h->ctrl->done : nullptr;
    s.partials = h->part_spmm;
    s.ticket = h->tickets + 0;
    s.dot_out = h->ctrl->pAp;
    return s;
}

// TMA-staged SELL SpMM (ls_sell_kernel.cuh): per-warp shared-memory rings fed by cp.async.bulk, launched with programmatic
// stream serialisation so that its matrix prefetch overlaps the tail of the previous kernel in the stream.
template <int K, bool DOT, int NW, int DEPTH, int MINB>
int launch_sell_tma_t(PcgHandle *h, const lsk::SellArgs &a, cudaStream_t s) {
    static bool prepared = false;
    const size_t smem = lsk::sell_tma_smem_bytes(NW, DEPTH);
    if (!prepared) {
        LS_CUDA_TRY(cudaFuncSetAttribute(lsk::spmm_sell_tma_kernel<K, DOT, NW, DEPTH, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prepared = true;
    }
    cudaLaunchConfig_t lc = {};
    int g = h->nslices < h->sm_count * MINB ? h->nslices : h->sm_count * MINB;
    lc.gridDim = dim3(g < 1 ? 1 : g);
    lc.blockDim = dim3(NW * 32);
    lc.dynamicSmemBytes = smem;
    lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = (h->sell_tma >= 10) ? 0 : 1;   // LS_SELL_TMA >= 10: same kernels without PDL (A/B)
    LS_CUDA_TRY(cudaLaunchKernelEx(&lc, lsk::spmm_sell_tma_kernel<K, DOT, NW, DEPTH, MINB>, a));
    g_ls_launches.fetch_add(1, std::memory_order_relaxed);
    return LS_OK;
}
template <int K, bool DOT = true>
int launch_sell_tma(PcgHandle *h, const lsk::SellArgs &a, cudaStream_t s) {
    if constexpr (K == 3) {
        switch (h->sell_tma % 10) {
            case 2: return launch_sell_tma_t<K, DOT, 24, 4, 1>(h, a, s);
            case 4: return launch_sell_tma_t<K, DOT, 16, 6, 1>(h, a, s);
            case 5: return launch_sell_tma_t<K, DOT, 16, 3, 2>(h, a, s);   // two CTAs per SM: the next launch's prefetch overlaps this one's tail
            case 6: return launch_sell_tma_t<K, DOT, 24, 2, 2>(h, a, s);
            case 7: return launch_sell_tma_t<K, DOT, 16, 2, 2>(h, a, s);   // 2 x 66 KB of rings: ~95 KB of L1 left for the gathers
            default: break;
        }
    }
    if (h->sell_tma % 10 == 1) return launch_sell_tma_t<K, DOT, 32, 3, 1>(h, a, s);
    return launch_sell_tma_t<K, DOT, 32, 2, 1>(h, a, s);
}

template <int K>
int launch_spmm(PcgHandle *h, bool with_done, cudaStream_t s) {
    if (h->sell_on) {
        lsk::SellArgs a{};
        a.V = (int)h->V;
        a.nslices = h->nslices;
        a.soff = h->soff;
        a.ent = h->ent;
        a.p = h->p;
        a.y = h->Ap;
        a.ldy = h->Vp;
        a.done = with_done ? &h->ctrl->done : nullptr;
        a.partials = h->part_spmm;
        a.ticket = h->tickets + 0;
        a.dot_out = h->ctrl->pAp;
        a.pf_halo = h->sell_pf;
        if (h->sell_tma) return launch_sell_tma<K>(h, a, s);
        lsk::spmm_sell_kernel<K, true><<<h->sell_grid, lsk::SELL_THREADS, 0, s>>>(a);
        LS_LAUNCH_CHECK();
        return LS_OK;
    }
    return lsk::spmm_launch(K, true, h->cfg, spmm_args(h, K, with_done), h->spmm_grid, s);
}

template <int K>
int launch_iteration(PcgHandle *h, cudaStream_t s) {
    int rc = launch_spmm<K>(h, true, s);
    if (rc) return rc;
    k_update_cs<K><<<h->vec_grid, VEC_THREADS, 0, s>>>(vec_args(h, 1));
    LS_LAUNCH_CHECK();
    k_pupdate<K><<<h->vec_grid, VEC_THREADS, 0, s>>>(vec_args(h, 1));
    LS_LAUNCH_CHECK();
    return LS_OK;
}

// host-side resources only the graph-mode solver needs: created on its first use (they cost ~0.3 ms at handle creation)
int graph_host_resources(PcgHandle *h) {
    if (h->cap_stream) return LS_OK;
    LS_CUDA_TRY(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    LS_CUDA_TRY(cudaEventCreateWithFlags(&h->ev[0], cudaEventDisableTiming));
    LS_CUDA_TRY(cudaEventCreateWithFlags(&h->ev[1], cudaEventDisableTiming));
    LS_CUDA_TRY(cudaMallocHost((void **)&h->pinned_done, 64));
    return LS_OK;
}

template <int K>
int build_graph(PcgHandle *h) {
    if (h->graph[K]) return LS_OK;
    {
        const int rc0 = graph_host_resources(h);
        if (rc0) return rc0;
    }
    cudaGraph_t g = nullptr;
    LS_CUDA_TRY(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
    int rc = LS_OK;
    for (int i = 0; i < CHUNK && rc == LS_OK; ++i) rc = launch_iteration<K>(h, h->cap_stream);
    cudaError_t e = cudaStreamEndCapture(h->cap_stream, &g);
    if (rc) {
        if (g) cudaGraphDestroy(g);
        return rc;
    }
    LS_CUDA_TRY(e);
    // launches recorded during capture were counted once; replays are counted in solve
    e = cudaGraphInstantiate(&h->graph[K], g, 0);
    cudaGraphDestroy(g);
    LS_CUDA_TRY(e);
    return LS_OK;
}

int finish_info(PcgHandle *h, float rtol, int maxit, float *info_src, float *info_host, cudaStream_t stream) {
    if (!info_host) return LS_OK;
    LS_CUDA_TRY(cudaMemcpyAsync(info_host, info_src, 8 * sizeof(float), cudaMemcpyDefault, stream));
    LS_CUDA_TRY(cudaStreamSynchronize(stream));
    const int st = (int)info_host[1];
    if (st == 3) {
        ls_set_error("CG breakdown after %d iterations (matrix not SPD or NaN in the right-hand side)", (int)info_host[0]);
        return LS_ERR_BREAKDOWN;
    }
    if (st == 2) {
        ls_set_error("PCG did not reach rtol=%g within maxit=%d (relres %g %g %g %g)", (double)rtol, maxit,
                     (double)info_host[2], (double)info_host[3], (double)info_host[4], (double)info_host[5]);
        return LS_ERR_NOT_CONVERGED;
    }
    return LS_OK;
}

int solve_persistent(PcgHandle *h, const float *b, float *x, float rtol, int maxit, float *info_dev, float *info_host,
                     cudaStream_t stream, bool resume) {
    lsp::PersistArgs a{};
    if (resume) {   // warm start: k_warm_load / SpMM / k_init<WARM> left x, r, p and the scalars in global memory
        a.resume_rz = h->ctrl->rz;
        a.resume_rr = h->ctrl->rr;
        a.resume_bb = h->ctrl->bb;
        a.resume_conv = h->ctrl->conv;
        a.resume_done = &h->ctrl->done;
    }
    a.V = (int)h->V;
    a.Vp = h->Vp;
    a.nslices = h->nslices;
    a.nsl_max = h->persist_nsl_max;
    a.soff = h->soff;
    a.ent = h->ent;
    a.dinv = h->dinv;
    a.x = h->x;
    a.r = h->r;
    a.Ap = h->Ap;
    a.p = h->p;
    a.b = b;
    a.out = x;
    a.perm = h->has_perm ? h->perm : nullptr;
    a.rtol = rtol;
    a.maxit = maxit;
    a.bar = h->gbar;
    a.partials = h->part_persist;
    a.info = info_dev ? info_dev : h->info;
    a.dbg = getenv("LS_PCG_PROFILE") ? h->dbg : nullptr;
    a.poff = h->poff;
    a.pcol = h->pcol;
    a.diagp = h->diagp;
    a.offc = h->offc;
    LS_CUDA_TRY(cudaMemsetAsync(h->gbar, 0, sizeof(lsp::GridBar), stream));   // barrier counter restarts at 0
    {
        // fast all-reduce slots this solve can touch: 2 per iteration (beyond the ring the kernel uses the slow path)
        long long need = 2LL * maxit + 8;
        if (need > h->ring_slots) need = h->ring_slots;
        const char *e = getenv("LS_PCG_FASTRED");   // "0": fenced all-reduce only; "N": at most N fast slots (tests the hand-over)
        a.ring = h->ring;
        a.ring_slots = (e && e[0] == '0') ? 0 : (h->persist_grid <= 255 ? (int)need : 0);
        if (e && atoi(e) > 0 && atoi(e) < a.ring_slots) a.ring_slots = atoi(e);
        if (a.ring_slots > 0) LS_CUDA_TRY(cudaMemsetAsync(h->ring, 0, (size_t)a.ring_slots * 64, stream));
    }
    void *params[] = {(void *)&a};
    const void *fn = h->persist_res ? (const void *)lsp::pcg_persistent_kernel<3, 1, false, lsp::PWARPS> : (const void *)lsp::pcg_persistent_kernel<3, 0, false, lsp::PWARPS>;
    if (h->persist_threads == lsp::PT_SMALL) fn = (const void *)lsp::pcg_persistent_kernel<3, 1, false, lsp::PT_SMALL / 32>;
    else if (a.dbg) {   // profiling build of the same kernel (LS_PCG_PROFILE): per-phase cycle counters in CTA 0
        fn = h->persist_res ? (const void *)lsp::pcg_persistent_kernel<3, 1, true, lsp::PWARPS> : (const void *)lsp::pcg_persistent_kernel<3, 0, true, lsp::PWARPS>;
        LS_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
    }
    if (h->pat_on) {   // same kernel, pattern-only phase A
        if (h->persist_threads == lsp::PT_SMALL) fn = (const void *)lsp::pcg_persistent_kernel<3, 1, false, lsp::PT_SMALL / 32, true>;
        else if (a.dbg) {
            fn = h->persist_res ? (const void *)lsp::pcg_persistent_kernel<3, 1, true, lsp::PWARPS, true> : (const void *)lsp::pcg_persistent_kernel<3, 0, true, lsp::PWARPS, true>;
            LS_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
        } else
            fn = h->persist_res ? (const void *)lsp::pcg_persistent_kernel<3, 1, false, lsp::PWARPS, true> : (const void *)lsp::pcg_persistent_kernel<3, 0, false, lsp::PWARPS, true>;
    }
    cudaError_t ce = cudaLaunchCooperativeKernel(fn, dim3(h->persist_grid), dim3(h->persist_threads), params, h->persist_smem, stream);
    if (ce != cudaSuccess) {
        // e.g. the device is partitioned (MPS / MIG limits) and cannot co-schedule the grid: not fatal, the graph-mode
        // solver computes the same thing; remember the failure so later solves go there directly
        cudaGetLastError();
        h->persist_on = 0;
        ls_set_error("cooperative launch of the persistent solver failed (%s); using the graph-mode solver", cudaGetErrorString(ce));
        return -1;
    }
    g_ls_launches.fetch_add(1, std::memory_order_relaxed);
    return finish_info(h, rtol, maxit, a.info, info_host, stream);
}

// ---- fused two-synchronisation solver (ls_pcg_fused.cuh) ------------------------------------------------------------
// instantiation table: (K, RES, NW, PAT, SYNC, PROF) -> kernel, or NULL when that combination is not built
// instantiation table: (K, RES, NW, PAT, SYNC, PROF, CHEB) -> kernel, or NULL when that combination is not built.  The
// instantiations live in three translation units (ls_fused_a/b/c.cu) so that they compile in parallel.
const void *fused_fn(int K, int res, int nw, int pat, int sync, int prof, int cheb = 0) {
    if (cheb) return (K == 3 && !prof) ? ls_fused_fn_cheb(res, nw, pat, sync) : nullptr;
    if (K == 3 && !prof) return ls_fused_fn_jacobi(res, nw, pat, sync);
    return ls_fused_fn_misc(K, res, nw, pat, sync, prof);
}

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return (e && e[0]) ? atoi(e) : dflt;
}

// LS_PCG_CLRES=N (opt-in): meshes of (one CTA's worth) < slices <= N run as ONE cluster of 16 CTAs with every vector -- the
// published rows included -- in (distributed) shared memory: RES = 4 of ls_pcg_fused.cuh; inside the iteration nothing but matrix
// entries comes from global memory.  Built because a 2.5 K-vertex solve "should not touch global memory"  -- and
// measured slower than the cooperative grid with the Chebyshev steps (profiles/r02_cluster_resident.jsonl): icosphere 0.278 vs
// 0.233 ms, 10 K vertices 0.47 vs 0.24 ms.  Per iteration (CTA 0, 2562 vertices): gather pass 2.3 k cycles (14 remote 8-byte
// loads per row against ~20 B/clk of DSMEM bandwidth per SM), the two exchanges 2.3 k + 3.4 k (fp64 shuffle trees, 16 remote
// stores, barrier.cluster with release/acquire, fixed-order re-sum, fp64 division) -- a cluster synchronisation that carries a
// deterministic reduction is ~2 k cycles, not the 0.4 k of a bare barrier.cluster, and 16 SMs are 16 SMs.
constexpr int CLRES_CS = 16;
static int clres_limit() { return env_int("LS_PCG_CLRES", LS_CLRES_DEFAULT); }
static bool clres_regime(int nslices) {
    if (env_int("LS_PCG_CLUSTER", -1) == 0) return false;
    return nslices > env_int("LS_PCG_ONECTA", lsp::PWARPS) && nslices <= clres_limit();
}

// Choose grid / cluster, residency and CTA shape for one K.  Small meshes (the CTA-resident rows of <= 16 SMs hold them)
// run as ONE thread-block cluster; everything else as a cooperative grid with one CTA per SM.
int configure_fused(PcgHandle *h, const LsDevInfo &di, int K, PcgHandle::FusedCfg *c) {
    memset(c, 0, sizeof(*c));
    if (!h->sell_on) return LS_OK;
    const char *algo = getenv("LS_PCG_ALGO");
    if (algo && (algo[0] == 'c' || algo[0] == 'C')) return LS_OK;          // A/B: round-1 three-synchronisation kernel
    const char *mode = getenv("LS_PCG_MODE");
    if (mode && (mode[0] == 'g' || mode[0] == 'G')) return LS_OK;
    const int cheb = (K == 3 && h->cheb_m > 1) ? 1 : 0;
    const int pat = (K == 3 && h->pat_on) ? 1 : 0;
    const int W = lsp::PWARPS;
    auto cap_slices = [&](int res, int dp, int sync) {   // slices per CTA that fit in shared memory at this residency level
        if (res == 0) return 1 << 30;
        int n = 0;
        while (lsf::fused_smem_bytes(K, res, n + 1, dp, cheb, sync) <= (size_t)di.max_smem_optin) ++n;
        return n;
    };
    const int cap3 = cap_slices(3, pat, 1), cap2c = cap_slices(2, 0, 1), cap2 = cap_slices(2, 0, 0), cap1 = cap_slices(1, 0, 0);
    const int cap4 = cheb ? 0 : (cap_slices(4, 0, 1) < 63 ? cap_slices(4, 0, 1) : 63);   // (63: the owner of a row is found by a 16-bit multiply)
    const int want_cluster = env_int("LS_PCG_CLUSTER", -1);   // -1 auto, 0 never, N force cluster size N
    const int force_res = env_int("LS_PCG_RES", -1);
    // ---- one CTA (everything, including the gathered vector, in shared memory) or, on request, one cluster
    // A cluster of 16 was measured slower than the cooperative grid for mid-size meshes (bunny x2: 1.81 vs 0.72 ms): 16 SMs
    // give 16 x ~28 B/clk of L2 bandwidth and cluster.sync flushes L1 each time, so it is opt-in (LS_PCG_CLUSTER=N).
    int cs = 0;
    if (want_cluster != 0) {
        // one CTA only while every warp has at most one slice: beyond that the single SM is instruction-issue bound (81 slices: 8.4 k
        // cycles for phase A alone) and the cooperative grid wins despite its ~2 x 3.5 k cycles of synchronisation per iteration
        if (h->nslices <= env_int("LS_PCG_ONECTA", W)) cs = 1;
        else if (!cheb && clres_regime(h->nslices)) cs = CLRES_CS;
        if (want_cluster > 0) cs = want_cluster;
        if (cs > 0 && (h->nslices + cs - 1) / cs > cap2c) cs = 0;
    }
    if (cs > 0) {
        const int nsl_max = (h->nslices + cs - 1) / cs;
        int res = (cs == 1 && K == 3 && h->cheb_m <= 1 && nsl_max <= cap3 && !(force_res >= 0 && force_res < 3)) ? 3 : 2;
        int nwc = W;
        if (cs > 1 && !cheb && nsl_max <= cap4 && !(force_res >= 0 && force_res < 4)) {
            res = 4;
            if (K == 3 && nsl_max <= lsp::PT_SMALL / 32 && !(getenv("LS_PCG_SMALLCTA") && getenv("LS_PCG_SMALLCTA")[0] == '0')) nwc = lsp::PT_SMALL / 32;
        }
        if (res == 4 && !fused_fn(K, res, nwc, pat, 1, 0, cheb)) { res = 2; nwc = W; }
        const int dp = (pat && nsl_max <= cap_slices(res, 1, 1)) ? 1 : 0;
        const void *fn = fused_fn(K, res, nwc, pat, 1, 0, cheb);
        const size_t smem = lsf::fused_smem_bytes(K, res, nsl_max, dp, cheb, 1);
        bool ok = fn != nullptr;
        if (ok && cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin) != cudaSuccess) ok = false;
        if (ok && cs > 8 && cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) ok = false;
        if (ok && cs > 1) {
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3(cs);
            lc.blockDim = dim3(nwc * 32);
            lc.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            lc.attrs = at;
            lc.numAttrs = 1;
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, fn, &lc) != cudaSuccess || ncl < 1) ok = false;
        }
        if (ok) {
            c->on = 1; c->grid = cs; c->res = res; c->nw = nwc; c->sync = 1; c->cluster = cs; c->nsl_max = nsl_max; c->pat = pat; c->dp = dp;
            c->smem = smem; c->fn = fn; c->fn_prof = cheb ? nullptr : fused_fn(K, res, nwc, pat, 1, 1);
            if (c->fn_prof) {
                cudaFuncSetAttribute(c->fn_prof, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin);
                if (cs > 8) cudaFuncSetAttribute(c->fn_prof, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            }
            cudaGetLastError();
            return LS_OK;
        }
        cudaGetLastError();
    }
    // ---- cooperative grid, one CTA per SM
    int coop = 0;
    if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, di.device) != cudaSuccess || !coop) {
        cudaGetLastError();
        return LS_OK;
    }
    int g = di.sm_count < h->nslices ? di.sm_count : h->nslices;
    if (g > 255) g = 255;
    if (g < 1) g = 1;
    const int nsl_max = (h->nslices + g - 1) / g;
    int res = nsl_max <= cap2 ? 2 : (nsl_max <= cap1 ? 1 : 0);
    if (force_res >= 0 && force_res < res) res = force_res;
    int nw = W;
    const char *et = getenv("LS_PCG_SMALLCTA");
    if (K == 3 && res == 2 && nsl_max <= 16 && !(et && et[0] == '0')) nw = lsp::PT_SMALL / 32;
    const void *fn = fused_fn(K, res, nw, pat, 0, 0, cheb);
    if (!fn) return LS_OK;
    int dp = (pat && res >= 1 && nsl_max <= cap_slices(res, 1, 0)) ? 1 : 0;
    if (dp && res == 1) {
        // L1 is what the shared-memory carve-out leaves of 256 KB, and the gathers re-use the published rows from it: if the kernel
        // fits a smaller carve-out without the pattern diagonal in shared memory, leave the diagonal in global memory
        // (V = 1e6: 196 instead of 228 KB, i.e. 60 instead of 28 KB of L1: 1.761 vs 1.775 ms, profiles/r02_l1_carveout_ab.jsonl)
        auto carve_kb = [](size_t bytes) {
            static const int steps[] = {8, 16, 32, 64, 100, 132, 164, 196, 228};
            for (int st : steps)
                if (bytes + 1024 <= (size_t)st * 1024) return st;
            return 228;
        };
        if (carve_kb(lsf::fused_smem_bytes(K, res, nsl_max, 0, cheb, 0)) < carve_kb(lsf::fused_smem_bytes(K, res, nsl_max, 1, cheb, 0))) dp = 0;
    }
    const int dp_env = env_int("LS_PCG_DP", -1);   // A/B: 0 the diagonal stays in global memory, 1 in shared memory whenever it fits
    if (dp_env == 0) dp = 0;
    if (dp_env == 1) dp = (pat && res >= 1 && nsl_max <= cap_slices(res, 1, 0)) ? 1 : 0;
    const size_t smem = lsf::fused_smem_bytes(K, res, nsl_max, dp, cheb, 0);
    // the attribute is per function and device, shared by every handle: always the device maximum, never a per-handle size
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin) != cudaSuccess) {
        cudaGetLastError();
        return LS_OK;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, nw * 32, smem) != cudaSuccess || occ < 1 || occ * di.sm_count < g) {
        cudaGetLastError();
        return LS_OK;
    }
    c->on = 1; c->grid = g; c->res = res; c->nw = nw; c->sync = 0; c->cluster = 0; c->nsl_max = nsl_max; c->pat = pat; c->dp = dp;
    c->smem = smem; c->fn = fn; c->fn_prof = cheb ? nullptr : fused_fn(K, res, nw, pat, 0, 1);
    if (c->fn_prof) cudaFuncSetAttribute(c->fn_prof, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin);
    cudaGetLastError();
    return LS_OK;
}

int solve_fused(PcgHandle *h, const float *b, float *x, const float *x0, int k, float rtol, int maxit, float *info_dev,
                float *info_host, cudaStream_t stream) {
    PcgHandle::FusedCfg &c = h->fused[k == 4 ? 1 : 0];
    lsf::FusedArgs a{};
    a.V = (int)h->V;
    a.Vp = h->Vp;
    a.nslices = h->nslices;
    a.nsl_max = c.nsl_max;
    a.kb = k;
    a.soff = h->soff;
    a.ent = h->ent;
    a.poff = h->poff;
    a.pcol = h->pcol;
    a.diagp = h->diagp;
    a.offc = h->offc;
    a.dinv = h->dinv;
    a.x = h->x;
    a.pv = h->pv;
    a.r = h->r;
    a.s = h->Ap;
    a.z = h->p;
    a.z2 = h->z2;
    a.cy = h->cy;
    a.cd = h->cd;
    a.dp_smem = c.dp;
    a.cheb_m = (k == 4) ? 0 : h->cheb_m;     // (the K = 4 instantiations carry the Jacobi preconditioner only)
    a.cheb_c0 = h->cheb_c0;
    for (int j = 0; j < 8; ++j) {
        a.cheb_c1[j] = h->cheb_c1[j];
        a.cheb_c2[j] = h->cheb_c2[j];
    }
    a.b = b;
    a.out = x;
    a.x0 = x0;
    a.perm = h->has_perm ? h->perm : nullptr;
    a.rtol = rtol;
    a.maxit = maxit;
    a.refine = h->refine;
    a.theta = h->theta;
    a.bar = h->gbar;
    a.partials = h->part_persist;
    a.info = info_dev ? info_dev : h->info;
    const bool prof = getenv("LS_PCG_PROFILE") != nullptr && c.fn_prof != nullptr;
    a.dbg = prof ? h->dbg : nullptr;
    const void *fn = prof ? c.fn_prof : c.fn;
    void *params[] = {(void *)&a};
    cudaError_t ce;
    if (c.sync == 1) {
        if (c.cluster > 1) {
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3(c.grid);
            lc.blockDim = dim3(c.nw * 32);
            lc.dynamicSmemBytes = c.smem;
            lc.stream = stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = c.cluster;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            lc.attrs = at;
            lc.numAttrs = 1;
            ce = cudaLaunchKernelExC(&lc, fn, params);
        } else {
            ce = cudaLaunchKernel(fn, dim3(1), dim3(c.nw * 32), params, c.smem, stream);
        }
    } else {
        LS_CUDA_TRY(cudaMemsetAsync(h->gbar, 0, sizeof(lsp::GridBar), stream));
        long long need = 2LL * maxit + 64;
        if (need > h->ring_slots) need = h->ring_slots;
        const char *e = getenv("LS_PCG_FASTRED");
        a.ring = h->ring;
        a.ring_slots = (e && e[0] == '0') ? 0 : (int)need;
        if (e && atoi(e) > 0 && atoi(e) < a.ring_slots) a.ring_slots = atoi(e);
        if (a.ring_slots > 0) LS_CUDA_TRY(cudaMemsetAsync(h->ring, 0, (size_t)a.ring_slots * 64, stream));
        ce = cudaLaunchCooperativeKernel(fn, dim3(c.grid), dim3(c.nw * 32), params, c.smem, stream);
    }
    if (ce != cudaSuccess) {
        cudaGetLastError();
        c.on = 0;   // e.g. a partitioned device that cannot co-schedule the grid: the older paths compute the same thing
        ls_set_error("launch of the fused solver failed (%s); falling back", cudaGetErrorString(ce));
        return -1;
    }
    g_ls_launches.fetch_add(1, std::memory_order_relaxed);
    return finish_info(h, rtol, maxit, a.info, info_host, stream);
}

template <int K>
int solve_k(PcgHandle *h, const float *b, float *x, const float *x0, float rtol, int maxit, float *info_dev,
            float *info_host, cudaStream_t stream) {
    if (h->fused[K == 4 ? 1 : 0].on) {
        const int frc = solve_fused(h, b, x, x0, K, rtol, maxit, info_dev, info_host, stream);
        if (frc != -1) return frc;
    }
    if (K == 3 && h->persist_on && x0 == nullptr) {
        const int prc = solve_persistent(h, b, x, rtol, maxit, info_dev, info_host, stream, false);
        if (prc != -1) return prc;     // -1: cooperative launch refused, fall through to the graph-mode solver
    }
    int occ;
    int rc = lsk::spmm_prepare(K, true, h->cfg, &occ);
    if (rc) return rc;
    rc = build_graph<K>(h);
    if (rc) return rc;
    VecArgs va = vec_args(h, 1);
    if (x0) {
        k_warm_load<K><<<h->vec_grid, VEC_THREADS, 0, stream>>>(va, x0);
        LS_LAUNCH_CHECK();
        rc = launch_spmm<K>(h, false, stream);
        if (rc) return rc;
        k_init<K, true><<<h->vec_grid, VEC_THREADS, 0, stream>>>(va, b, rtol, maxit, 0);
        LS_LAUNCH_CHECK();
        k_init<K, false><<<h->vec_grid, VEC_THREADS, 0, stream>>>(va, b, rtol, maxit, 1);   // runs only if `restart`
        LS_LAUNCH_CHECK();
        if (K == 3 && h->persist_on) {   // iterate in the persistent kernel from the state the three kernels above left
            const int prc = solve_persistent(h, b, x, rtol, maxit, info_dev, info_host, stream, true);
            if (prc != -1) return prc;
        }
    } else {
        k_init<K, false><<<h->vec_grid, VEC_THREADS, 0, stream>>>(va, b, rtol, maxit, 0);
        LS_LAUNCH_CHECK();
    }
    // iterate: replay the CHUNK-iteration graph; the device `done` flag of chunk c-1 is checked while chunk c runs
    // (kernels of a chunk enqueued after convergence see `done` and return immediately).
    int launched = 0, nq = 0;
    bool finished = false;
    while (!finished) {
        LS_CUDA_TRY(cudaGraphLaunch(h->graph[K], stream));
        g_ls_launches.fetch_add(3 * CHUNK, std::memory_order_relaxed);
        LS_CUDA_TRY(cudaMemcpyAsync(&h->pinned_done[nq & 1], &h->ctrl->done, sizeof(int), cudaMemcpyDeviceToHost, stream));
        LS_CUDA_TRY(cudaEventRecord(h->ev[nq & 1], stream));
        ++nq;
        launched += CHUNK;
        if (nq >= 2) {
            LS_CUDA_TRY(cudaEventSynchronize(h->ev[(nq - 2) & 1]));
            if (h->pinned_done[(nq - 2) & 1] != 0) finished = true;
        }
        if (!finished && launched >= maxit) {   // every iteration maxit allows is enqueued: drain
            LS_CUDA_TRY(cudaEventSynchronize(h->ev[(nq - 1) & 1]));
            finished = true;
        }
    }
    k_final<K><<<h->vec_grid, VEC_THREADS, 0, stream>>>(va, x, info_dev ? info_dev : h->info);
    LS_LAUNCH_CHECK();
    if (info_host) {
        LS_CUDA_TRY(cudaMemcpyAsync(info_host, info_dev ? info_dev : h->info, 8 * sizeof(float), cudaMemcpyDefault, stream));
        LS_CUDA_TRY(cudaStreamSynchronize(stream));
        const int st = (int)info_host[1];
        if (st == 3) {
            ls_set_error("CG breakdown after %d iterations (matrix not SPD or NaN in the right-hand side)", (int)info_host[0]);
            return LS_ERR_BREAKDOWN;
        }
        if (st == 2) {
            ls_set_error("PCG did not reach rtol=%g within maxit=%d (relres %g %g %g %g)", (double)rtol, maxit,
                         (double)info_host[2], (double)info_host[3], (double)info_host[4], (double)info_host[5]);
            return LS_ERR_NOT_CONVERGED;
        }
    }
    return LS_OK;
}

}  // namespace

extern "C" int ls_pcg_workspace_bytes(int64_t V, int64_t nnz, int k_max, size_t *bytes_out) {
    LS_REQUIRE(bytes_out != nullptr, "bytes_out is NULL");
    LS_REQUIRE(V > 0 && nnz > 0 && V < (int64_t)0x7ffffff0 && nnz < (int64_t)0x7ffffff0, "size out of range");
    LS_REQUIRE(k_max >= 1 && k_max <= KMAX, "k_max must be in [1,4]");
    *bytes_out = carve_handle(nullptr, nullptr, V, nnz, k_max, GRID_CAP);
    return LS_OK;
}

extern "C" int ls_pcg_create(void **handle_out, int64_t V, int64_t nnz, const int32_t *rowptr, const int32_t *col,
                             const float *val, const int32_t *perm_new2old, int precond, int k_max, void *workspace,
                             size_t workspace_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LS_REQUIRE(handle_out != nullptr, "handle_out is NULL");
    *handle_out = nullptr;
    LS_REQUIRE(V > 0 && nnz > 0 && V < (int64_t)0x7ffffff0 && nnz < (int64_t)0x7ffffff0, "size out of range");
    LS_REQUIRE(k_max >= 1 && k_max <= KMAX, "k_max must be in [1,4]");
    LS_REQUIRE(precond >= 0 && precond <= 3, "precond must be 0 (none), 1 (Jacobi), 2 (Chebyshev polynomial over Jacobi) or 3 (auto)");
    LS_REQUIRE(rowptr && col && val, "NULL CSR pointer");
    LS_REQUIRE(workspace != nullptr && ((uintptr_t)workspace & 255) == 0, "workspace NULL or not 256-byte aligned");
    LsDevInfo di;
    int rc = ls_dev_info(&di);
    if (rc) return rc;
    size_t need = carve_handle(nullptr, nullptr, V, nnz, k_max, GRID_CAP);
    if (workspace_bytes < need) {
        ls_set_error("PCG workspace too small: %zu < %zu", workspace_bytes, need);
        return LS_ERR_WORKSPACE;
    }
    PcgHandle *h = new (std::nothrow) PcgHandle();
    LS_REQUIRE(h != nullptr, "out of host memory");
    memset(h, 0, sizeof(*h));
    h->V = V;
    h->nnz = nnz;
    h->k_max = k_max;
    h->precond = precond;
    h->device = di.device;
    h->sm_count = di.sm_count;
    h->max_smem_optin = di.max_smem_optin;
    h->refine = env_int("LS_PCG_REFINE", 1);
    h->sell_tma = env_int("LS_SELL_TMA", 3);   // 32 warps x 2 slots of 2 KB: measured best (profiles/r02_sell_tma_variants.jsonl)
    h->sell_pf = env_int("LS_SELL_PF", 1024);
    h->theta = 3.0f;
    h->ws_bytes = need;
    carve_handle(h, (char *)workspace, V, nnz, k_max, GRID_CAP);

    auto fail = [&](int code) {
        ls_pcg_destroy(h);
        return code;
    };
#define TRY_OR_FAIL(expr)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ls_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return fail(LS_ERR_CUDA);                                                          \
        }                                                                                      \
    } while (0)

    // zero what has to start at zero (vector planes incl. their padding rows, scalars, counters): NOT the matrix copies,
    // which are written in full below -- at V = 1e6 that is ~100 MB of memset instead of ~500 MB
    TRY_OR_FAIL(cudaMemsetAsync(h->dinv, 0, (size_t)((char *)h->soff - (char *)h->dinv), stream));
    TRY_OR_FAIL(cudaMemsetAsync(h->perm, 0, (size_t)((char *)h->poff - (char *)h->perm), stream));
    h->has_perm = perm_new2old ? 1 : 0;
    if (perm_new2old && !getenv("LS_FORCE_REORDER")) {
        // keep the caller's numbering when it already gathers at least as coherently as the Morton order would
        const unsigned gb = (unsigned)((V + 255) / 256);
        unsigned long long *sc = reinterpret_cast<unsigned long long *>(h->part_vec);   // scratch, zeroed above
        k_locality_score<<<gb > 2048 ? 2048 : gb, 256, 0, stream>>>(V, rowptr, col, sc);
        g_ls_launches.fetch_add(1);
        TRY_OR_FAIL(cudaGetLastError());
        // a numbering whose neighbouring rows already gather from neighbouring columns (a grid, a remesher's output) is kept
        // without ever building the permuted copy: fewer than 1 in 8 (row, slot) pairs break the coalescing
        unsigned long long hs0 = 0;
        TRY_OR_FAIL(cudaMemcpyAsync(&hs0, sc, sizeof(hs0), cudaMemcpyDeviceToHost, stream));
        TRY_OR_FAIL(cudaStreamSynchronize(stream));
        if (hs0 * 8ull <= (unsigned long long)nnz) {
            perm_new2old = nullptr;
            h->has_perm = 0;
            TRY_OR_FAIL(cudaMemsetAsync(sc, 0, 16, stream));
        }
        // otherwise the score of the permuted order needs the permuted CSR: build it, score it, then decide
    }
    if (perm_new2old) {
        // internal copy in the caller's locality order: A' = P A P^T, rows re-sorted by new column
        const unsigned gb = (unsigned)((V + 255) / 256);
        TRY_OR_FAIL(cudaMemcpyAsync(h->perm, perm_new2old, (size_t)V * 4, cudaMemcpyDeviceToDevice, stream));
        TRY_OR_FAIL(cudaMemsetAsync(h->inv, 0xff, (size_t)V * 4, stream));
        k_perm_inv_len<<<gb, 256, 0, stream>>>(V, h->perm, rowptr, h->inv, h->rowptr, h->flags);
        g_ls_launches.fetch_add(1);
        TRY_OR_FAIL(cudaGetLastError());
        k_perm_check<<<gb, 256, 0, stream>>>(V, h->perm, h->inv, h->flags);
        g_ls_launches.fetch_add(1);
        TRY_OR_FAIL(cudaGetLastError());
        rc = ls_exclusive_scan_i32(h->rowptr, h->rowptr, V, h->scan, stream);
        if (rc) return fail(rc);
        k_perm_rows<<<gb, 256, 0, stream>>>(V, h->perm, h->inv, rowptr, col, val, h->rowptr, h->col, h->val, h->flags);
        g_ls_launches.fetch_add(1);
        TRY_OR_FAIL(cudaGetLastError());
        if (!getenv("LS_FORCE_REORDER")) {
            unsigned long long *sc = reinterpret_cast<unsigned long long *>(h->part_vec);
            k_locality_score<<<gb > 2048 ? 2048 : gb, 256, 0, stream>>>(V, h->rowptr, h->col, sc + 1);
            g_ls_launches.fetch_add(1);
            TRY_OR_FAIL(cudaGetLastError());
            unsigned long long hs[2] = {0, 0};
            TRY_OR_FAIL(cudaMemcpyAsync(hs, sc, sizeof(hs), cudaMemcpyDeviceToHost, stream));
            TRY_OR_FAIL(cudaStreamSynchronize(stream));
            TRY_OR_FAIL(cudaMemsetAsync(sc, 0, sizeof(hs), stream));
            if (hs[0] <= hs[1]) {   // native order is at least as good: drop the permutation
                h->has_perm = 0;
                TRY_OR_FAIL(cudaMemcpyAsync(h->rowptr, rowptr, (size_t)(V + 1) * 4, cudaMemcpyDeviceToDevice, stream));
                TRY_OR_FAIL(cudaMemcpyAsync(h->col, col, (size_t)nnz * 4, cudaMemcpyDeviceToDevice, stream));
                TRY_OR_FAIL(cudaMemcpyAsync(h->val, val, (size_t)nnz * 4, cudaMemcpyDeviceToDevice, stream));
            }
        }
    } else {
        TRY_OR_FAIL(cudaMemcpyAsync(h->rowptr, rowptr, (size_t)(V + 1) * 4, cudaMemcpyDeviceToDevice, stream));
        TRY_OR_FAIL(cudaMemcpyAsync(h->col, col, (size_t)nnz * 4, cudaMemcpyDeviceToDevice, stream));
        TRY_OR_FAIL(cudaMemcpyAsync(h->val, val, (size_t)nnz * 4, cudaMemcpyDeviceToDevice, stream));
    }
    k_pad_tail<<<1, 32, 0, stream>>>(h->rowptr, h->col, h->val, V, nnz);
    g_ls_launches.fetch_add(1);
    TRY_OR_FAIL(cudaGetLastError());
    k_dinv<<<(unsigned)((h->Vp + 255) / 256), 256, 0, stream>>>(V, h->Vp, h->rowptr, h->col, h->val, precond, h->dinv, h->flags);
    g_ls_launches.fetch_add(1);
    TRY_OR_FAIL(cudaGetLastError());

    // launch geometry
    lsk::spmm_config(&h->cfg);

    int occ = 1;
    rc = lsk::spmm_prepare(3, true, h->cfg, &occ);
    if (rc) return fail(rc);
    h->spmm_grid = lsk::spmm_grid_for(V, di.sm_count, occ);
    if (h->spmm_grid > GRID_CAP) h->spmm_grid = GRID_CAP;
    int64_t vg = (h->Vp / 4 + VEC_THREADS - 1) / VEC_THREADS;   // one float4 per thread per column
    if (vg > GRID_CAP) vg = GRID_CAP;                           // beyond that the kernels grid-stride
    if (vg < 1) vg = 1;
    h->vec_grid = (int)vg;
    k_partition<<<(h->spmm_grid + 1 + 127) / 128, 128, 0, stream>>>(V, h->rowptr, h->spmm_grid, h->part);
    g_ls_launches.fetch_add(1);
    TRY_OR_FAIL(cudaGetLastError());

    // SELL-32 copy of the (re-ordered) CSR: the fast in-solver SpMM engine
    {
        const unsigned wb = (unsigned)(((int64_t)h->nslices * 32 + 255) / 256);
        lsk::sell_width_kernel<<<wb, 256, 0, stream>>>((int)V, h->nslices, h->rowptr, h->soff);
        g_ls_launches.fetch_add(1);
        TRY_OR_FAIL(cudaGetLastError());
        rc = ls_exclusive_scan_i32(h->soff, h->soff, h->nslices, h->scan, stream);
        if (rc) return fail(rc);
        lsk::sell_fill_kernel<<<wb, 256, 0, stream>>>((int)V, h->nslices, h->rowptr, h->col, h->val, h->soff, h->ent,
                                                       h->sell_cap);
        g_ls_launches.fetch_add(1);
        TRY_OR_FAIL(cudaGetLastError());
    }

    // block plan: every CTA's block boundaries, so the producer warp never chases rowptr at run time
    rc = lsk::spmm_plan(h->rowptr, h->part, h->spmm_grid, h->cfg.cap, h->desc, h->desc_cnt, h->flags + 1, stream);
    if (rc) return fail(rc);

    // pattern-only copy (opt-in): are all off-diagonal values bitwise equal?
    const bool want_pat = want_pattern();
    unsigned int hmm[2] = {0xffffffffu, 0u};
    h->pat_on = 0;
    if (want_pat) {
        TRY_OR_FAIL(cudaMemsetAsync(h->patmm, 0xff, 4, stream));
        TRY_OR_FAIL(cudaMemsetAsync(h->patmm + 1, 0, 4, stream));
        lsk::pat_detect_kernel<<<(unsigned)((V + 255) / 256), 256, 0, stream>>>((int)V, h->rowptr, h->col, h->val, h->patmm);
        g_ls_launches.fetch_add(1);
        TRY_OR_FAIL(cudaGetLastError());
        TRY_OR_FAIL(cudaMemcpyAsync(hmm, h->patmm, sizeof(hmm), cudaMemcpyDeviceToHost, stream));
    }

    float hgersh = 0.f;
    if (precond == 3) {
        // auto: the polynomial pays where the iteration is synchronisation-bound and its vectors fit in shared memory -- the
        // cooperative grid at residency level 2 (measured: 4K..250K vertices 13-26 % faster, V = 1e6 13 % slower, one CTA slower)
        const int g = di.sm_count < h->nslices ? di.sm_count : h->nslices;
        const int nsl_max = (h->nslices + g - 1) / g;
        const bool fits = lsf::fused_smem_bytes(3, 2, nsl_max, 1, 1, 0) <= (size_t)di.max_smem_optin;
        precond = (h->nslices > env_int("LS_PCG_ONECTA", lsp::PWARPS) && fits) ? 2 : 1;
        // ... and not where one cluster holds everything in shared memory: a synchronisation costs a tenth there, plain CG's
        // fewer SpMVs win
        if (clres_regime(h->nslices)) precond = 1;
        h->precond = precond;
    }
    if (precond == 2) {
        TRY_OR_FAIL(cudaMemsetAsync(h->gersh, 0, 64, stream));
        k_gershgorin<<<(unsigned)((V + 255) / 256), 256, 0, stream>>>(V, h->rowptr, h->col, h->val, h->gersh);
        g_ls_launches.fetch_add(1);
        TRY_OR_FAIL(cudaGetLastError());
        TRY_OR_FAIL(cudaMemcpyAsync(&hgersh, h->gersh, sizeof(float), cudaMemcpyDeviceToHost, stream));
    }
    int hflags2[2] = {0, 0};
    int sell_total = 0;
    TRY_OR_FAIL(cudaMemcpyAsync(hflags2, h->flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    TRY_OR_FAIL(cudaMemcpyAsync(&sell_total, h->soff + h->nslices, sizeof(int), cudaMemcpyDeviceToHost, stream));
    TRY_OR_FAIL(cudaStreamSynchronize(stream));
    const int hflags = hflags2[0];
    h->planned = (hflags2[1] == 0) ? 1 : 0;
    h->sell_entries = sell_total;
    {
        // engine choice: SELL unless padding blew past the buffer (very long rows) or LS_SPMM_ENGINE=csr
        const char *e = getenv("LS_SPMM_ENGINE");
        const bool want_csr = e && (e[0] == 'c' || e[0] == 'C');
        h->sell_on = (!want_csr && sell_total > 0 && (long long)sell_total <= h->sell_cap) ? 1 : 0;
        int socc = 0;
        TRY_OR_FAIL(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&socc, lsk::spmm_sell_kernel<3, true>, lsk::SELL_THREADS, 0));
        if (socc < 1) socc = 1;
        int64_t sg = ((int64_t)h->nslices + lsk::SELL_WARPS - 1) / lsk::SELL_WARPS;   // >= one slice per warp
        if (sg > (int64_t)di.sm_count * socc) sg = (int64_t)di.sm_count * socc;
        if (sg > GRID_CAP) sg = GRID_CAP;
        if (sg < 1) sg = 1;
        h->sell_grid = (int)sg;
    }
    if (want_pat && h->sell_on && hmm[0] == hmm[1]) {
        // every off-diagonal entry carries the same value: build the 4-byte-per-entry copy (no further host round trip:
        // its padded size is bounded by the general SELL copy's, which fits)
        memcpy(&h->offc, &hmm[0], 4);
        const unsigned wb = (unsigned)(((int64_t)h->nslices * 32 + 255) / 256);
        lsk::pat_width_kernel<<<wb, 256, 0, stream>>>((int)V, h->nslices, h->rowptr, h->col, h->poff);
        g_ls_launches.fetch_add(1);
        TRY_OR_FAIL(cudaGetLastError());
        rc = ls_exclusive_scan_i32(h->poff, h->poff, h->nslices, h->scan, stream);
        if (rc) return fail(rc);
        lsk::pat_fill_kernel<<<wb, 256, 0, stream>>>((int)V, h->nslices, h->rowptr, h->col, h->val, h->poff, h->pcol, h->pat_cap,
                                                      h->offc, h->diagp);
        g_ls_launches.fetch_add(1);
        TRY_OR_FAIL(cudaGetLastError());
        h->pat_on = 1;
    }
    {
        // persistent single-kernel solve: one 768-thread CTA per SM (256 for mid-size meshes), cooperative launch; r / Ap / dinv in shared memory
        // when the CTA's rows fit (RES = 1), in global memory otherwise (RES = 0)
        const char *e = getenv("LS_PCG_MODE");
        const bool want_graph = e && (e[0] == 'g' || e[0] == 'G');
        h->persist_on = 0;
        int coop = 0;
        TRY_OR_FAIL(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, di.device));
        if (!want_graph && h->sell_on && coop) {
            int g = di.sm_count < h->nslices ? di.sm_count : h->nslices;
            if (h->nslices <= 4 * lsp::PWARPS) g = 1;   // tiny mesh (<= 3K rows): one CTA, grid barriers become __syncthreads
            if (g > 255) g = 255;
            if (g < 1) g = 1;
            const int nsl_max = (h->nslices + g - 1) / g;
            const char *er = getenv("LS_PCG_RES");
            int res = 1;
            size_t smem = lsp::persist_smem_bytes(3, 1, nsl_max);
            if ((er && er[0] == '0') || (int)smem > di.max_smem_optin) {
                res = 0;
                smem = lsp::persist_smem_bytes(3, 0, nsl_max);
            }
            cudaError_t ce = res ? cudaFuncSetAttribute(lsp::pcg_persistent_kernel<3, 1, false, lsp::PWARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin)
                                 : cudaFuncSetAttribute(lsp::pcg_persistent_kernel<3, 0, false, lsp::PWARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin);
            int occ = 0;
            if (ce == cudaSuccess)
                ce = res ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lsp::pcg_persistent_kernel<3, 1, false, lsp::PWARPS>, lsp::PT, smem)
                         : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lsp::pcg_persistent_kernel<3, 0, false, lsp::PWARPS>, lsp::PT, smem);
            h->persist_threads = lsp::PT;
            const char *et = getenv("LS_PCG_SMALLCTA");
            if (ce == cudaSuccess && occ >= 1 && res == 1 && g > 1 && nsl_max <= 16 && !(et && et[0] == '0')) {
                // a CTA owns only a handful of slices: 8 warps are enough and make every CTA-level barrier cheaper
                if (cudaFuncSetAttribute(lsp::pcg_persistent_kernel<3, 1, false, lsp::PT_SMALL / 32>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin) == cudaSuccess)
                    h->persist_threads = lsp::PT_SMALL;
                else
                    cudaGetLastError();
            }
            if (ce == cudaSuccess && occ >= 1 && h->pat_on) {
                // the pattern-only instantiations need the same opt-in shared memory size; if that fails, stay general
                cudaError_t cp = cudaSuccess;
                if (h->persist_threads == lsp::PT_SMALL)
                    cp = cudaFuncSetAttribute(lsp::pcg_persistent_kernel<3, 1, false, lsp::PT_SMALL / 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin);
                else if (res)
                    cp = cudaFuncSetAttribute(lsp::pcg_persistent_kernel<3, 1, false, lsp::PWARPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin);
                else
                    cp = cudaFuncSetAttribute(lsp::pcg_persistent_kernel<3, 0, false, lsp::PWARPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin);
                if (cp != cudaSuccess) {
                    cudaGetLastError();
                    h->pat_on = 0;
                }
            }
            if (ce == cudaSuccess && occ >= 1) {
                h->persist_on = 1;
                h->persist_grid = g;
                h->persist_res = res;
                h->persist_nsl_max = nsl_max;
                h->persist_smem = smem;
            } else {
                cudaGetLastError();   // not fatal: the graph path stays available
            }
        }
    }
    h->cheb_m = 0;
    if (precond == 2 && hgersh > 0.f) {
        // Chebyshev semi-iteration for D^-1 A on [b/30, b], b = 1.02 x the Gershgorin bound: theta, delta, sigma = theta/delta,
        // rho_0 = 1/sigma;  d_0 = g/theta;  rho_j = 1/(2 sigma - rho_{j-1});  d_j = rho_j rho_{j-1} d_{j-1} + 2 rho_j/delta (g - B y_j)
        int m = env_int("LS_PCG_CHEB_M", 4);
        if (m < 2) m = 2;
        if (m > 8) m = 8;
        const double b = 1.02 * (double)hgersh, a = b / 30.0;
        const double th = 0.5 * (b + a), de = 0.5 * (b - a), sg = th / de;
        double rho = 1.0 / sg;
        h->cheb_c0 = (float)(1.0 / th);
        for (int j = 1; j < m; ++j) {
            const double rn = 1.0 / (2.0 * sg - rho);
            h->cheb_c1[j - 1] = (float)(rn * rho);
            h->cheb_c2[j - 1] = (float)(2.0 * rn / de);
            rho = rn;
        }
        h->cheb_m = m;
    }
    rc = configure_fused(h, di, 3, &h->fused[0]);
    if (rc) return fail(rc);
    if (k_max >= 4) {
        rc = configure_fused(h, di, 4, &h->fused[1]);
        if (rc) return fail(rc);
    }
    if (hflags & 8) {
        ls_set_error("perm_new2old is not a permutation of [0, V)");
        return fail(LS_ERR_BAD_ARG);
    }
    if (hflags & (1 | 4)) {
        ls_set_error("CSR is malformed (column index out of range or decreasing rowptr)");
        return fail(LS_ERR_INDEX_RANGE);
    }
    if (hflags & 2) {
        ls_set_error("matrix has a missing or non-positive diagonal entry: not SPD");
        return fail(LS_ERR_BREAKDOWN);
    }
#undef TRY_OR_FAIL
    *handle_out = h;
    return LS_OK;
}

extern "C" int ls_pcg_solve(void *handle, const float *b, float *x, const float *x0, int k, float rtol, int maxit,
                            float *info_dev, float *info_host, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PcgHandle *h = (PcgHandle *)handle;
    LS_REQUIRE(h != nullptr, "handle is NULL");
    LS_REQUIRE(b != nullptr && x != nullptr, "b or x is NULL");
    LS_REQUIRE(k >= 1 && k <= h->k_max, "k out of range for this handle");
    LS_REQUIRE(rtol > 0.f && maxit > 0, "rtol and maxit must be positive");
    int dev = -1;
    LS_CUDA_TRY(cudaGetDevice(&dev));
    LS_REQUIRE(dev == h->device, "handle was created on a different device");
    switch (k) {
        case 1: return solve_k<1>(h, b, x, x0, rtol, maxit, info_dev, info_host, stream);
        case 2: return solve_k<2>(h, b, x, x0, rtol, maxit, info_dev, info_host, stream);
        case 3: return solve_k<3>(h, b, x, x0, rtol, maxit, info_dev, info_host, stream);
        default: return solve_k<4>(h, b, x, x0, rtol, maxit, info_dev, info_host, stream);
    }
}

extern "C" int ls_pcg_set_refinement(void *handle, int max_restarts, float theta) {
    PcgHandle *h = (PcgHandle *)handle;
    LS_REQUIRE(h != nullptr, "handle is NULL");
    LS_REQUIRE(max_restarts >= 0 && max_restarts <= 8, "max_restarts must be in [0, 8]");
    LS_REQUIRE(theta >= 1.0f && theta < 1e6f, "theta must be >= 1");
    h->refine = max_restarts;
    h->theta = theta;
    return LS_OK;
}

extern "C" int ls_pcg_destroy(void *handle) {
    PcgHandle *h = (PcgHandle *)handle;
    if (!h) return LS_OK;
    for (int k = 0; k <= KMAX; ++k)
        if (h->graph[k]) cudaGraphExecDestroy(h->graph[k]);
    if (h->ev[0]) cudaEventDestroy(h->ev[0]);
    if (h->ev[1]) cudaEventDestroy(h->ev[1]);
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    if (h->pinned_done) cudaFreeHost(h->pinned_done);
    delete h;
    return LS_OK;
}

extern "C" int ls_pcg_bench_spmm(void *handle, int k, int launches, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PcgHandle *h = (PcgHandle *)handle;
    LS_REQUIRE(h != nullptr, "handle is NULL");
    LS_REQUIRE(k >= 1 && k <= h->k_max, "k out of range for this handle");
    int occ;
    int rc = lsk::spmm_prepare(k, true, h->cfg, &occ);
    if (rc) return rc;
    for (int i = 0; i < launches; ++i) {
        switch (k) {
            case 1: rc = launch_spmm<1>(h, false, stream); break;
            case 2: rc = launch_spmm<2>(h, false, stream); break;
            case 3: rc = launch_spmm<3>(h, false, stream); break;
            default: rc = launch_spmm<4>(h, false, stream); break;
        }
        if (rc) return rc;
    }
    return LS_OK;
}

namespace {
template <int K>
int bench_one(PcgHandle *h, int which, cudaStream_t stream) {
    int rc = LS_OK;
    VecArgs va = vec_args(h, 1);
    va.bench = 1;
    if (which == 0 || which == 3) rc = launch_spmm<K>(h, false, stream);
    if (which == 4) {   // pure y = A p, no dot-product epilogue (the SpMV of BASELINE's metric)
        if constexpr (K == 3) {
            if (h->sell_on && h->sell_tma) {
                lsk::SellArgs a{};
                a.V = (int)h->V;
                a.nslices = h->nslices;
                a.soff = h->soff;
                a.ent = h->ent;
                a.p = h->p;
                a.y = h->Ap;
                a.ldy = h->Vp;
                a.pf_halo = h->sell_pf;
                rc = launch_sell_tma<3, false>(h, a, stream);
            } else {
                rc = launch_spmm<K>(h, false, stream);
            }
        } else {
            rc = launch_spmm<K>(h, false, stream);
        }
    }
    if (rc) return rc;
    if (which == 1 || which == 3) {
        k_update_cs<K><<<h->vec_grid, VEC_THREADS, 0, stream>>>(va);
        LS_LAUNCH_CHECK();
    }
    if (which == 2 || which == 3) {
        k_pupdate<K><<<h->vec_grid, VEC_THREADS, 0, stream>>>(va);
        LS_LAUNCH_CHECK();
    }
    return LS_OK;
}
}  // namespace

extern "C" int ls_pcg_bench(void **handles, int n_handles, int k, int which, int launches, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LS_REQUIRE(handles != nullptr && n_handles >= 1, "no handles");
    LS_REQUIRE(which >= 0 && which <= 4, "which: 0 SpMM+dot, 1 update, 2 p-update, 3 one full iteration, 4 SpMM without the dot epilogue");
    for (int i = 0; i < n_handles; ++i) {
        PcgHandle *h = (PcgHandle *)handles[i];
        LS_REQUIRE(h != nullptr, "NULL handle");
        LS_REQUIRE(k >= 1 && k <= h->k_max, "k out of range for this handle");
        int occ;
        int rc = lsk::spmm_prepare(k, true, h->cfg, &occ);
        if (rc) return rc;
    }
    for (int i = 0; i < launches; ++i) {
        PcgHandle *h = (PcgHandle *)handles[i % n_handles];
        int rc;
        switch (k) {
            case 1: rc = bench_one<1>(h, which, stream); break;
            case 2: rc = bench_one<2>(h, which, stream); break;
            case 3: rc = bench_one<3>(h, which, stream); break;
            default: rc = bench_one<4>(h, which, stream); break;
        }
        if (rc) return rc;
    }
    return LS_OK;
}

extern "C" int ls_pcg_phase_cycles(void *handle, int64_t *out, int n, void *stream_) {
    PcgHandle *h = (PcgHandle *)handle;
    LS_REQUIRE(h != nullptr && out != nullptr, "NULL pointer");
    LS_REQUIRE(n >= 8 && n <= 8 + 8 * 256, "n out of range");
    LS_CUDA_TRY(cudaMemcpyAsync(out, h->dbg, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
    LS_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream_));
    return LS_OK;
}

extern "C" int ls_pcg_describe(void *handle, int64_t *out8) {
    PcgHandle *h = (PcgHandle *)handle;
    LS_REQUIRE(h != nullptr && out8 != nullptr, "NULL pointer");
    const PcgHandle::FusedCfg &fc = h->fused[0];
    if (fc.on) {   // fused two-synchronisation solver: [engine, padded entries, grid, cluster size, mode 10 + RES, preconditioner, threads, re-ordered]
        out8[0] = fc.pat ? 2 : 1;
        out8[1] = h->sell_entries;
        out8[2] = fc.grid;
        out8[3] = fc.cluster;
        out8[4] = 10 + fc.res;
        out8[5] = (h->cheb_m > 1) ? 2 : h->precond;    // preconditioner actually in use (auto resolved)
        out8[6] = fc.nw * 32;
        out8[7] = h->has_perm;
        return LS_OK;
    }
    out8[0] = (h->pat_on && h->persist_on) ? 2 : h->sell_on;   // 2 = pattern-only SELL-32 in the persistent kernel, 1 = SELL-32 engine, 0 = TMA-staged CSR engine
    out8[1] = h->sell_entries;            // padded entries of the SELL copy
    out8[2] = h->sell_on ? h->sell_grid : h->spmm_grid;
    out8[3] = h->vec_grid;
    out8[4] = h->persist_on ? (h->persist_res ? 2 : 1) : 0;   // 0 graph of 3 kernels, 1 persistent (global r/Ap), 2 persistent (smem r/Ap)
    out8[5] = h->persist_on ? h->persist_grid : 0;
    out8[6] = h->planned;
    out8[7] = h->has_perm;
    return LS_OK;
}

extern "C" int64_t ls_pcg_spmm_bytes(void *handle, int k) {
    PcgHandle *h = (PcgHandle *)handle;
    if (!h) return 0;
    return 8 * h->nnz + 4 * (h->V + 1) + 8 * (int64_t)k * h->V;
}
