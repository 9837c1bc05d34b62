// ls_fused_c.cu -- profiling instantiations (per-phase cycle counters) and the 4-column instantiations
#include "ls_pcg_fused.cuh"
#include "ls_fused_inst.h"

namespace {
template <int K, int RES, int NW, bool PAT, int SYNC, bool PROF, bool CHEB = false>
const void *ffn() { return (const void *)lsf::pcg_fused_kernel<K, RES, NW, PAT, SYNC, PROF, CHEB>; }
constexpr int W = lsp::PWARPS, WS = lsp::PT_SMALL / 32;
}  // namespace

const void *ls_fused_fn_misc(int K, int res, int nw, int pat, int sync, int prof) {
    if (K == 3 && prof && nw == W) {
        if (sync == 0 && res == 1) return pat ? ffn<3, 1, W, true, 0, true>() : ffn<3, 1, W, false, 0, true>();
        if (sync == 0 && res == 2) return pat ? ffn<3, 2, W, true, 0, true>() : ffn<3, 2, W, false, 0, true>();
        if (sync == 1 && res == 2) return pat ? ffn<3, 2, W, true, 1, true>() : ffn<3, 2, W, false, 1, true>();
        if (sync == 1 && res == 3) return pat ? ffn<3, 3, W, true, 1, true>() : ffn<3, 3, W, false, 1, true>();
    }
    if (K == 4 && !prof && !pat && nw == W) {
        if (sync == 0 && res == 0) return ffn<4, 0, W, false, 0, false>();
        if (sync == 0 && res == 1) return ffn<4, 1, W, false, 0, false>();
        if (sync == 0 && res == 2) return ffn<4, 2, W, false, 0, false>();
        if (sync == 1 && res == 2) return ffn<4, 2, W, false, 1, false>();
    }
    return nullptr;
}
