// decode_dual.cu — fused SwiGLU up-projection on the decode family (ggq_mm_swiglu, T <= 16):
//     C[T, O] = silu(X . dequant(Wg)^T) * (X . dequant(Wu)^T)
// One launch streams both packed matrices once; a tile is 8 gate rows over the same 8 up rows, so both projections of
// an output land in the same lane's accumulator fragment and the store applies silu(gate) * up (decode_impl.cuh, DUAL).
// No reference counterpart (the fusion that follows the gate/up matmuls in a SwiGLU FFN).
#include "decode_impl.cuh"

namespace ggq {
namespace dec {

// The flat (tile, chunk) walk only (AT == 1): 12 warps, else 8 warps, else cluster split-K; one CTA per SM.
template <int FMT>
static bool make_plan_dual(const MmArgs& a, int T, Plan& pl) {
    if (make_plan_cfg<FMT>(a, T, 12, 1, false, pl) && pl.p.stages >= 2) return true;
    if (make_plan_cfg<FMT>(a, T, 8, 1, false, pl) && pl.p.stages >= 2) return true;
    int best_s = 0, best_nw = 0, best_waste = 1 << 30;
    for (int S = 2; S <= 8; ++S)
        for (int nw = 12; nw >= 8; nw -= 4) {
            Plan t;
            if (!make_plan_cfg<FMT>(a, T, nw, 1, false, t, S) || t.p.stages < 2) continue;
            const int waste = S * t.p.cps * 1024 / t.p.nc;
            if (waste < best_waste) {
                best_waste = waste;
                best_s = S;
                best_nw = nw;
            }
        }
    return best_s != 0 && make_plan_cfg<FMT>(a, T, best_nw, 1, false, pl, best_s);
}

template <int FMT>
static int launch_fmt_dual(const MmArgs& a, bool plan_only) {
    Plan pl;
    const int T = static_cast<int>(a.T);
    if constexpr (FMT == 1) {
        if (make_plan_wide(a, T, pl)) {   // Q4_K single token on large layers: the wide-chunk geometry (decode_impl.cuh)
            if (plan_only) return 0;
            return launch_kernel<1, 1, 1, 8, 1, true, true, true>(pl, a.stream, a.W2);
        }
    }
    if (!make_plan_dual<FMT>(a, T, pl) || pl.at != 1) return GGQ_E_FAMILY;
    if (plan_only) return 0;
    if (T == 1 && pl.p.n_slices == 1)
        return pl.nw == 12 ? launch_kernel<FMT, 1, 1, 12, 1, true, true>(pl, a.stream, a.W2)
                           : launch_kernel<FMT, 1, 1, 8, 1, true, true>(pl, a.stream, a.W2);
    if (pl.nt == 1)
        return pl.nw == 12 ? launch_kernel<FMT, 1, 1, 12, 1, false, true>(pl, a.stream, a.W2)
                           : launch_kernel<FMT, 1, 1, 8, 1, false, true>(pl, a.stream, a.W2);
    return pl.nw == 12 ? launch_kernel<FMT, 2, 1, 12, 1, false, true>(pl, a.stream, a.W2)
                       : launch_kernel<FMT, 2, 1, 8, 1, false, true>(pl, a.stream, a.W2);
}

}  // namespace dec

// a.W = gate, a.W2 = up, one output.  GGQ_E_FAMILY when the shape has no single-pass plan (the caller composes).
int launch_decode_dual(int fmt, const MmArgs& a, bool plan_only) {
    if (!a.W2 || a.sync || a.n_out != 1 || a.T < 1 || a.T > 16 || a.O < 8) return GGQ_E_FAMILY;
    if (reinterpret_cast<uintptr_t>(a.W2) & 15) return GGQ_E_FAMILY;
    if (!decode_supports(fmt, a)) return GGQ_E_FAMILY;
    switch (fmt) {
        case GGQ_Q8_0: return dec::launch_fmt_dual<0>(a, plan_only);
        case GGQ_Q4_K: return dec::launch_fmt_dual<1>(a, plan_only);
        case GGQ_Q6_K: return dec::launch_fmt_dual<2>(a, plan_only);
    }
    return GGQ_E_FORMAT;
}

}  // namespace ggq
