/* roundtrip.c -- the large-steps hot path driven through the C ABI alone (no Python, no torch): what a native host of
 * the reference's path would call.  Mirrors the reference's own usage (README "Parameterization", scripts/main.py):
 *
 *     M = compute_matrix(verts, faces, lambda_)          ls_assemble_count / ls_assemble_fill
 *     u = to_differential(M, verts)                      ls_spmm_csr_f32
 *     v = from_differential(M, u, 'Cholesky')            ls_pcg_create / ls_pcg_solve
 *     AdamUniform.step()                                 ls_adam_uniform_step
 *
 * and checks  v == verts  (||v - verts|| / ||verts|| <= 1e-5) on the host.
 *
 * build:  make -C examples/c_host            (gcc + the CUDA runtime for cudaMalloc/cudaMemcpy only)
 * run:    examples/c_host/roundtrip [n]      n x n plane, default 512 (262 144 vertices)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <cuda_runtime_api.h>
#include "largesteps_b200.h"

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));            \
            return 2;                                                                              \
        }                                                                                          \
    } while (0)
#define LS(call)                                                                                   \
    do {                                                                                           \
        int s_ = (call);                                                                           \
        if (s_ != LS_OK) {                                                                         \
            fprintf(stderr, "%s:%d: %s (%s)\n", __FILE__, __LINE__, ls_status_string(s_), ls_last_error()); \
            return 3;                                                                              \
        }                                                                                          \
    } while (0)

static size_t up256(size_t b) { return (b + 255) & ~(size_t)255; }

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 512;
    if (n < 2 || n > 4096) {
        fprintf(stderr, "n must be in [2, 4096]\n");
        return 1;
    }
    const int64_t V = (int64_t)n * n, F = 2 * (int64_t)(n - 1) * (n - 1);
    const float lambda = 19.0f;

    /* a height field over the unit square, two triangles per cell */
    float *verts = (float *)malloc(sizeof(float) * 3 * V);
    int32_t *faces = (int32_t *)malloc(sizeof(int32_t) * 3 * F);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const float x = (float)j / (n - 1), y = (float)i / (n - 1);
            float *p = verts + 3 * ((int64_t)i * n + j);
            p[0] = x;
            p[1] = y;
            p[2] = 0.1f * sinf(6.f * x) * cosf(5.f * y);
        }
    int64_t f = 0;
    for (int i = 0; i + 1 < n; ++i)
        for (int j = 0; j + 1 < n; ++j) {
            const int32_t a = i * n + j, b = a + 1, c = a + n, d = c + 1;
            faces[3 * f + 0] = a; faces[3 * f + 1] = b; faces[3 * f + 2] = d; ++f;
            faces[3 * f + 0] = a; faces[3 * f + 1] = d; faces[3 * f + 2] = c; ++f;
        }

    float *d_verts, *d_u, *d_v, *d_info;
    int32_t *d_faces;
    CU(cudaMalloc((void **)&d_verts, sizeof(float) * 3 * V));
    CU(cudaMalloc((void **)&d_faces, sizeof(int32_t) * 3 * F));
    CU(cudaMalloc((void **)&d_u, sizeof(float) * 3 * V));
    CU(cudaMalloc((void **)&d_v, sizeof(float) * 3 * V));
    CU(cudaMalloc((void **)&d_info, sizeof(float) * 8));
    CU(cudaMemcpy(d_verts, verts, sizeof(float) * 3 * V, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_faces, faces, sizeof(int32_t) * 3 * F, cudaMemcpyHostToDevice));

    /* M = I + lambda L (uniform Laplacian): count, then fill the CSR the solver streams */
    size_t asm_bytes = 0;
    LS(ls_assemble_workspace_bytes(F, V, &asm_bytes));
    void *d_asm;
    CU(cudaMalloc(&d_asm, up256(asm_bytes)));
    int64_t nnz = 0;
    LS(ls_assemble_count(d_faces, 4, F, V, d_asm, asm_bytes, &nnz, NULL));
    int32_t *d_rowptr, *d_col;
    float *d_val;
    CU(cudaMalloc((void **)&d_rowptr, sizeof(int32_t) * (V + 1)));
    CU(cudaMalloc((void **)&d_col, sizeof(int32_t) * nnz));
    CU(cudaMalloc((void **)&d_val, sizeof(float) * nnz));
    LS(ls_assemble_fill(d_faces, 4, d_verts, F, V, /*cotan*/ 0, /*diag_shift*/ 1.0f, /*scale*/ lambda, d_asm, asm_bytes, nnz,
                        NULL, NULL, NULL, d_rowptr, d_col, d_val, NULL));

    /* u = M verts */
    LS(ls_spmm_csr_f32(V, d_rowptr, d_col, d_val, d_verts, 3, d_u, 3, 3, NULL));

    /* v = M^-1 u */
    size_t ws_bytes = 0;
    LS(ls_pcg_workspace_bytes(V, nnz, 3, &ws_bytes));
    void *d_ws, *solver = NULL;
    CU(cudaMalloc(&d_ws, up256(ws_bytes)));
    LS(ls_pcg_create(&solver, V, nnz, d_rowptr, d_col, d_val, NULL, /*Jacobi*/ 1, 3, d_ws, ws_bytes, NULL));
    float info[8];
    LS(ls_pcg_solve(solver, d_u, d_v, NULL, 3, 1e-7f, 10000, d_info, info, NULL));

    float *v = (float *)malloc(sizeof(float) * 3 * V);
    CU(cudaMemcpy(v, d_v, sizeof(float) * 3 * V, cudaMemcpyDeviceToHost));
    double num = 0.0, den = 0.0;
    for (int64_t i = 0; i < 3 * V; ++i) {
        const double d = (double)v[i] - (double)verts[i];
        num += d * d;
        den += (double)verts[i] * (double)verts[i];
    }
    const double err = sqrt(num / den);
    int64_t desc[8];
    LS(ls_pcg_describe(solver, desc));

    /* one AdamUniform step on u with v standing in for the gradient (t = 1, so c = 1 - beta^1):
     *   g1 = b1 g1 + (1-b1) g;  g2 = b2 g2 + (1-b2) g^2;  u -= lr (g1/c1) / (1e-8 + sqrt(max(g2)/c2))   (optimize.py:35-41) */
    float *d_g1, *d_g2;
    void *d_scratch;
    CU(cudaMalloc((void **)&d_g1, sizeof(float) * 3 * V));
    CU(cudaMalloc((void **)&d_g2, sizeof(float) * 3 * V));
    CU(cudaMalloc(&d_scratch, 256));
    CU(cudaMemset(d_g1, 0, sizeof(float) * 3 * V));
    CU(cudaMemset(d_g2, 0, sizeof(float) * 3 * V));
    const double b1 = 0.9, b2 = 0.999;
    LS(ls_adam_uniform_step(d_u, d_v, d_g1, d_g2, 3 * V, 0.01f, (float)b1, (float)b2, (float)(1.0 - b1), (float)(1.0 - b2),
                            (float)(1.0 - b1), (float)(1.0 - b2), d_scratch, NULL));
    CU(cudaDeviceSynchronize());

    printf("V=%lld nnz=%lld iterations=%d status=%d relres=%.2e roundtrip_rel_l2=%.2e mode=%lld kernels_launched=%llu\n",
           (long long)V, (long long)nnz, (int)info[0], (int)info[1], info[2], err, (long long)desc[4],
           (unsigned long long)ls_launch_count());
    LS(ls_pcg_destroy(solver));
    if (!(err <= 1e-5) || nnz != 7 * V - 8 * (int64_t)n + 2) {
        fprintf(stderr, "FAILED\n");
        return 4;
    }
    printf("c-abi roundtrip ok\n");
    return 0;
}
