// decode.cu — the decode family's mm entry points (kernel and planner: decode_impl.cuh).
#include "decode_impl.cuh"

namespace ggq {

void decode_set_trace(void* buf) {
    dec::g_trace = static_cast<unsigned long long*>(buf);
    dec::g_trace_n = 0;
}

bool decode_supports(int fmt, const MmArgs& a) {
    if (a.T < 1 || a.O < 1 || a.K < fmt_qk(fmt)) return false;
    if ((reinterpret_cast<uintptr_t>(a.W) & 15) || (reinterpret_cast<uintptr_t>(a.X) & 15) || (a.ldx & 7)) return false;
    const int64_t nb = a.K / fmt_qk(fmt);
    if ((nb * fmt_blk(fmt)) & 15) return false;  // rows must be whole 16-byte vectors (Q8_0 / Q6_K: nb % 8 == 0)
    return true;
}

// Host-only view of the decode planner (tests / DESIGN.md): out = {KW, AT, NT, slices, chunks per
// slice, stages, grid, batches, smem bytes}.  Returns 0, or GGQ_E_FAMILY when no plan fits.
int decode_plan(int fmt, const MmArgs& a, int* out, bool* wide) {
    dec::Plan pl;
    if (wide) *wide = false;
    const int T = static_cast<int>(a.T > 16 ? 16 : a.T);
    bool ok = false;
    switch (fmt) {
        case GGQ_Q8_0: ok = dec::make_plan<0>(a, T, pl); break;
        case GGQ_Q4_K:
            ok = dec::make_plan_wide(a, T, pl);
            if (ok && wide) *wide = true;
            if (!ok) ok = dec::make_plan<1>(a, T, pl);
            break;
        case GGQ_Q6_K: ok = dec::make_plan<2>(a, T, pl); break;
    }
    if (!ok) return GGQ_E_FAMILY;
    const int v[9] = {pl.p.KW, pl.at, pl.nt, pl.p.n_slices, pl.p.cps, pl.p.stages, pl.grid * 100 + pl.occ,
                      pl.p.num_batches, static_cast<int>(pl.smem)};
    for (int i = 0; i < 9; ++i) out[i] = v[i];
    return 0;
}

int launch_decode(int fmt, const MmArgs& a) {
    switch (fmt) {
        case GGQ_Q8_0: return dec::launch_fmt<0>(a);
        case GGQ_Q4_K: return dec::launch_fmt<1>(a);
        case GGQ_Q6_K: return dec::launch_fmt<2>(a);
    }
    return GGQ_E_FORMAT;
}

}  // namespace ggq
