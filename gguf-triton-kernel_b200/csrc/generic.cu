// generic.cu — the any-shape family and the standalone dequantize op.
//
// generic GEMM: one warp per output row o, a chunk of TT tokens per pass.  Every weight is expanded
// with the bit-exact scalar recipe of formats.cuh (so this family reproduces X.float() @ dequant(W).T
// up to fp32 summation order), byte loads only — no alignment or size assumption beyond K % QK == 0.
// It is the path for shapes the fast families reject: rows that are not 4-byte multiples (Q8_0 /
// Q6_K with an odd block count, e.g. the reference test grid's K = 32 / 256 cases,
// test/test_mmq_q8_0.py:20, test_mmq_q6_k.py:20) and 16 < T < 64.
#include "common.cuh"
#include "formats.cuh"
#include "../../include/ggq.h"

namespace ggq {

template <class F, int TT>
__global__ void __launch_bounds__(128) generic_mm_kernel(const uint8_t* __restrict__ W, const __half* __restrict__ X,
                                                         int64_t ldx, OutPtrs outs, int64_t ldc, int64_t O,
                                                         int64_t T, int64_t K) {
    const int lane = threadIdx.x & 31;
    const int64_t o = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
    const int64_t t0 = static_cast<int64_t>(blockIdx.y) * TT;
    if (o >= O) return;
    const int64_t nb = K / F::QK;
    const uint8_t* row = W + o * nb * F::BLK;
    float acc[TT];
#pragma unroll
    for (int i = 0; i < TT; ++i) acc[i] = 0.f;
    for (int64_t b = 0; b < nb; ++b) {
        const uint8_t* blk = row + b * F::BLK;
#pragma unroll
        for (int i = 0; i < F::QK / 32; ++i) {
            const int e = lane + 32 * i;
            const float w = __half2float(dequant_elem(F{}, blk, e));
            const int64_t k = b * F::QK + e;
#pragma unroll
            for (int tt = 0; tt < TT; ++tt) {
                if (t0 + tt < T) acc[tt] = fmaf(w, __half2float(X[(t0 + tt) * ldx + k]), acc[tt]);
            }
        }
    }
#pragma unroll
    for (int tt = 0; tt < TT; ++tt) {
        float v = acc[tt];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        acc[tt] = v;
    }
    if (lane == 0) {
#pragma unroll
        for (int tt = 0; tt < TT; ++tt) {
            if (t0 + tt < T) {
                const __half h = __float2half_rn(acc[tt]);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (i < outs.n) outs.p[i][(t0 + tt) * ldc + o] = h;
            }
        }
    }
}

template <class F>
static int launch_generic_t(const MmArgs& a) {
    constexpr int TT = 8;
    dim3 grid(static_cast<unsigned>((a.O + 3) / 4), static_cast<unsigned>((a.T + TT - 1) / TT));
    if (grid.y > 65535) return GGQ_E_SHAPE;
    generic_mm_kernel<F, TT><<<grid, 128, 0, a.stream>>>(a.W, static_cast<const __half*>(a.X), a.ldx, make_outs(a),
                                                         a.ldc, a.O, a.T, a.K);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

int launch_generic(int fmt, const MmArgs& a) {
    switch (fmt) {
        case GGQ_Q8_0: return launch_generic_t<Q8_0>(a);
        case GGQ_Q4_K: return launch_generic_t<Q4_K>(a);
        case GGQ_Q6_K: return launch_generic_t<Q6_K>(a);
    }
    return GGQ_E_FORMAT;
}

// ---- standalone dequantize: one thread per element pair, coalesced fp16x2 stores -------------
template <class F>
__global__ void __launch_bounds__(256) dequant_kernel(const uint8_t* __restrict__ W, __half2* __restrict__ out,
                                                      int64_t n_pairs) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < n_pairs; p += stride) {
        const int64_t e0 = 2 * p;
        const uint8_t* blk = W + (e0 / F::QK) * F::BLK;
        const int e = static_cast<int>(e0 % F::QK);
        out[p] = __halves2half2(dequant_elem(F{}, blk, e), dequant_elem(F{}, blk, e + 1));
    }
}

template <class F>
static int launch_dequant_t(const uint8_t* W, void* out, int64_t O, int64_t K, cudaStream_t s) {
    const int64_t n_pairs = O * K / 2;
    if (n_pairs == 0) return 0;
    int64_t blocks = (n_pairs + 255) / 256;
    const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    dequant_kernel<F><<<static_cast<unsigned>(blocks), 256, 0, s>>>(W, static_cast<__half2*>(out), n_pairs);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

int launch_dequant(int fmt, const uint8_t* W, void* out, int64_t O, int64_t K, cudaStream_t s) {
    int rc = 0;
    if (launch_dequant64(fmt, W, out, O, K, s, &rc)) return rc;  // same code path as the prefill GEMM's B operand
    switch (fmt) {
        case GGQ_Q8_0: return launch_dequant_t<Q8_0>(W, out, O, K, s);
        case GGQ_Q4_K: return launch_dequant_t<Q4_K>(W, out, O, K, s);
        case GGQ_Q6_K: return launch_dequant_t<Q6_K>(W, out, O, K, s);
    }
    return GGQ_E_FORMAT;
}

}  // namespace ggq
