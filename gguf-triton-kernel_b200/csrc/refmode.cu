// refmode.cu — "reference arithmetic" mode: Q8_1-quantized activations, integer block dots, fp16 accumulation,
// in exactly the operation order of the reference's CPU implementations, so the result equals
// kernels/cpu_impls/mmq_*_q8_1_cpu bit for bit.  One thread per output element walks the
// blocks sequentially; this is a parity tool, not a fast path.
//   Q8_0  kernels/cpu_impls/mmq_q8_0_q8_1_cpu.py:37-54   r = fp16(fp16(d_w*d_x) * dot);            C = fp16(C + r)
//   Q4_K  kernels/cpu_impls/mmq_q4_k_q8_1_cpu.py:94-117  r = ((d*sc)*d_x)*dot - (dmin*m)*s_x (fp32); C = fp16(C + fp16(r))
//   Q6_K  kernels/cpu_impls/mmq_q6_k_q8_1_cpu.py:117-150 r = d_x*((d*sc1)*dot1 + (d*sc2)*dot2) (fp32); C = fp16(C + fp16(r))
// Every fp32 operation is an explicitly rounded intrinsic (__fmul_rn / __fadd_rn / __fsub_rn): no FMA contraction.
#include "../../include/ggq.h"
#include "common.cuh"
#include "formats.cuh"

namespace ggq {

__device__ __forceinline__ float hf(const uint8_t* p) { return __half2float(load_half_bytes(p)); }
__device__ __forceinline__ float acc16(float c, float r32) {  // C (fp16 value held as float) += r
    const float r16 = __half2float(__float2half_rn(r32));     // torch rounds the Python scalar to the tensor dtype first
    return __half2float(__float2half_rn(__fadd_rn(c, r16)));
}

template <int FMT>
__global__ void __launch_bounds__(128) refmode_kernel(const uint8_t* __restrict__ W, const uint8_t* __restrict__ XQ,
                                                      __half* __restrict__ C, int64_t O, int64_t T, int64_t K) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= O * T) return;
    const int64_t o = idx % O, t = idx / O;
    const int64_t nb32 = K / 32;
    const uint8_t* xrow = XQ + t * nb32 * 36;
    float c = 0.f;
    if (FMT == GGQ_Q8_0) {
        const uint8_t* wrow = W + o * nb32 * 34;
        for (int64_t b = 0; b < nb32; ++b) {
            const uint8_t* wb = wrow + b * 34;
            const uint8_t* xb = xrow + b * 36;
            int dot = 0;
            for (int i = 0; i < 32; ++i) dot += static_cast<int>(static_cast<int8_t>(wb[2 + i])) * static_cast<int>(static_cast<int8_t>(xb[4 + i]));
            const float sc = __half2float(__float2half_rn(__fmul_rn(hf(wb), hf(xb))));  // fp16 * fp16 -> fp16
            c = acc16(c, __fmul_rn(sc, static_cast<float>(dot)));
        }
    } else if (FMT == GGQ_Q4_K) {
        const uint8_t* wrow = W + o * (K / 256) * 144;
        for (int64_t sb = 0; sb < K / 256; ++sb) {
            const uint8_t* wb = wrow + sb * 144;
            const float d = hf(wb), dmin = hf(wb + 2);
            for (int j = 0; j < 8; ++j) {
                int sc, m;
                q4k_scale_min(wb + 4, j, sc, m);
                const uint8_t* xb = xrow + (sb * 8 + j) * 36;
                int dot = 0;
                for (int i = 0; i < 32; ++i) {
                    const int byte = wb[16 + (j >> 1) * 32 + i];
                    const int q = (j & 1) ? (byte >> 4) : (byte & 15);
                    dot += q * static_cast<int>(static_cast<int8_t>(xb[4 + i]));
                }
                const float t1 = __fmul_rn(__fmul_rn(__fmul_rn(d, static_cast<float>(sc)), hf(xb)), static_cast<float>(dot));
                const float t2 = __fmul_rn(__fmul_rn(dmin, static_cast<float>(m)), hf(xb + 2));
                c = acc16(c, __fsub_rn(t1, t2));
            }
        }
    } else {
        const uint8_t* wrow = W + o * (K / 256) * 210;
        for (int64_t sb = 0; sb < K / 256; ++sb) {
            const uint8_t* wb = wrow + sb * 210;
            const float d = hf(wb + 208);
            for (int j = 0; j < 8; ++j) {
                const uint8_t* xb = xrow + (sb * 8 + j) * 36;
                int dot1 = 0, dot2 = 0;
                for (int i = 0; i < 16; ++i) {
                    dot1 += q6k_quant(wb, 32 * j + i) * static_cast<int>(static_cast<int8_t>(xb[4 + i]));
                    dot2 += q6k_quant(wb, 32 * j + 16 + i) * static_cast<int>(static_cast<int8_t>(xb[20 + i]));
                }
                const float s1 = __fmul_rn(d, static_cast<float>(static_cast<int8_t>(wb[192 + 2 * j])));
                const float s2 = __fmul_rn(d, static_cast<float>(static_cast<int8_t>(wb[192 + 2 * j + 1])));
                const float inner = __fadd_rn(__fmul_rn(s1, static_cast<float>(dot1)), __fmul_rn(s2, static_cast<float>(dot2)));
                c = acc16(c, __fmul_rn(hf(xb), inner));
            }
        }
    }
    C[t * O + o] = __float2half_rn(c);
}

}  // namespace ggq

using namespace ggq;

extern "C" int ggq_mm_ref_q8_1(int fmt, const void* W, const void* XQ, void* C, int64_t O, int64_t T, int64_t K,
                               void* stream) {
    if (fmt < GGQ_Q8_0 || fmt > GGQ_Q6_K) return GGQ_E_FORMAT;
    if (O < 0 || T < 0 || K < 0 || K % fmt_qk(fmt) != 0) return GGQ_E_SHAPE;
    if (O == 0 || T == 0) return 0;
    if (!W || !XQ || !C) return GGQ_E_POINTER;
    const int64_t n = O * T;
    const unsigned grid = static_cast<unsigned>((n + 127) / 128);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint8_t* w = static_cast<const uint8_t*>(W);
    const uint8_t* x = static_cast<const uint8_t*>(XQ);
    __half* c = static_cast<__half*>(C);
    switch (fmt) {
        case GGQ_Q8_0: refmode_kernel<GGQ_Q8_0><<<grid, 128, 0, s>>>(w, x, c, O, T, K); break;
        case GGQ_Q4_K: refmode_kernel<GGQ_Q4_K><<<grid, 128, 0, s>>>(w, x, c, O, T, K); break;
        default: refmode_kernel<GGQ_Q6_K><<<grid, 128, 0, s>>>(w, x, c, O, T, K); break;
    }
    count_launch();
    return static_cast<int>(cudaGetLastError());
}
