// ls_capi.cu -- library plumbing: version, error strings, device info, prefix scan.
#include <stdarg.h>
#include <string.h>
#include "ls_common.cuh"

std::atomic<uint64_t> g_ls_launches{0};
static thread_local char g_ls_err[512] = "";

void ls_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_ls_err, sizeof(g_ls_err), fmt, ap);
    va_end(ap);
}

extern "C" int ls_version(void) { return 10000 * 0 + 100 * 1 + 0; }
extern "C" const char *ls_last_error(void) { return g_ls_err; }
extern "C" uint64_t ls_launch_count(void) { return g_ls_launches.load(); }
extern "C" const char *ls_status_string(int s) {
    switch (s) {
        case LS_OK: return "ok";
        case LS_ERR_BAD_ARG: return "bad argument";
        case LS_ERR_CUDA: return "CUDA error";
        case LS_ERR_BREAKDOWN: return "CG breakdown (matrix not SPD or NaN input)";
        case LS_ERR_NOT_CONVERGED: return "not converged within maxit";
        case LS_ERR_UNSUPPORTED: return "unsupported configuration";
        case LS_ERR_INDEX_RANGE: return "index out of range / unsorted COO";
        case LS_ERR_WORKSPACE: return "workspace too small";
        default: return "unknown status";
    }
}

int ls_dev_info(LsDevInfo *out) {
    static thread_local LsDevInfo cache = {-1, 0, 0, 0};
    int dev = 0;
    LS_CUDA_TRY(cudaGetDevice(&dev));
    if (cache.device != dev) {
        LsDevInfo d;
        d.device = dev;
        LS_CUDA_TRY(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
        LS_CUDA_TRY(cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        LS_CUDA_TRY(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
        if (d.cc_major < 10) {
            ls_set_error("libls_b200 requires an sm_100a (Blackwell) device, found compute capability %d.x", d.cc_major);
            return LS_ERR_UNSUPPORTED;
        }
        cache = d;
    }
    *out = cache;
    return LS_OK;
}

// ---- exclusive scan (int32) -----------------------------------------------------------------------
// Three-pass: per-tile sums -> recursive scan of the sums -> per-tile scan with carry-in.  Tiles of 4096.
namespace {
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const int *__restrict__ in, int *__restrict__ sums,
                                                              int64_t n) {
    __shared__ int wsum[SCAN_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + (int64_t)i * SCAN_THREADS + threadIdx.x;
        if (idx < n) s += in[idx];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = threadIdx.x < SCAN_THREADS / 32 ? wsum[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) sums[blockIdx.x] = v;
    }
}

// each thread owns SCAN_ITEMS consecutive elements (blocked arrangement) so the scan is exact and simple
__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles(const int *__restrict__ in, int *__restrict__ out,
                                                          const int *__restrict__ carry, int64_t n, int write_total) {
    __shared__ int wsum[SCAN_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = warp_incl_scan(s, lane);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < SCAN_THREADS / 32 ? wsum[lane] : 0;
        int wi = warp_incl_scan(w, lane);
        if (lane < SCAN_THREADS / 32) wsum[lane] = wi - w;
    }
    __syncthreads();
    int run = (carry ? carry[blockIdx.x] : 0) + wsum[warp] + incl - s;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
    // total goes to out[n] (written by the thread that owns element n-1)
    if (write_total && n > 0 && base <= n - 1 && n - 1 < base + SCAN_ITEMS) out[n] = run;
}
__global__ void scan_zero_total(int *out) { out[0] = 0; }
}  // namespace

size_t ls_scan_scratch_elems(int64_t n) {
    size_t total = 0;
    int64_t m = n;
    while (m > 1) {
        m = (m + SCAN_TILE - 1) / SCAN_TILE;
        total += (size_t)m * 2 + 8;   // sums + scanned sums (+1 total) per level
        if (m == 1) break;
    }
    return total + 16;
}

static int scan_rec(const int *in, int *out, int64_t n, int *scratch, cudaStream_t stream, int write_total) {
    if (n <= 0) {
        if (write_total) {
            scan_zero_total<<<1, 1, 0, stream>>>(out);
            LS_LAUNCH_CHECK();
        }
        return LS_OK;
    }
    int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles == 1) {
        scan_tiles<<<1, SCAN_THREADS, 0, stream>>>(in, out, nullptr, n, write_total);
        LS_LAUNCH_CHECK();
        return LS_OK;
    }
    int *sums = scratch;
    int *carry = scratch + tiles;
    int *rest = scratch + 2 * tiles + 8;
    scan_tile_sums<<<(unsigned)tiles, SCAN_THREADS, 0, stream>>>(in, sums, n);
    LS_LAUNCH_CHECK();
    int rc = scan_rec(sums, carry, tiles, rest, stream, 0);
    if (rc != LS_OK) return rc;
    scan_tiles<<<(unsigned)tiles, SCAN_THREADS, 0, stream>>>(in, out, carry, n, write_total);
    LS_LAUNCH_CHECK();
    return LS_OK;
}

int ls_exclusive_scan_i32(const int *in, int *out, int64_t n, int *scratch, cudaStream_t stream) {
    return scan_rec(in, out, n, scratch, stream, 1);
}
