// pack.cu — GPU packers for the formats the reference packs in pure Python, byte-identical to them:
//   Q8_0 weights      utils/quantize/q8_0.py:4-49   (all-fp16 arithmetic: d = max|x| / 127, q = rint(x / d), d = 1 if max == 0)
//   Q8_1 activations  utils/quantize/q8_1.py:18-70  (d = max|x| / 127 or 0, q = rint(x / (d or 1)), s = d * fp16(sum q))
// plus an fp32 Q6_K dequantizer (the reference's dequantize_q6_k returns fp32, utils/quantize/q6_k.py:157).
// One warp per 32-element block.  fp16 division is done as an IEEE fp32 division rounded to fp16, which is the
// correctly rounded fp16 quotient (24 >= 2*11 + 2 bits), i.e. exactly what torch's CPU half kernels produce.
#include "../../include/ggq.h"
#include "common.cuh"
#include <algorithm>

#include "formats.cuh"

namespace ggq {

template <bool Q8_1>
__global__ void __launch_bounds__(256) quantize_q8_kernel(const __half* __restrict__ x, uint8_t* __restrict__ out,
                                                          int64_t nblocks) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    constexpr int BLK = Q8_1 ? 36 : 34;
    for (int64_t b = warp; b < nblocks; b += nwarps) {
        const __half v = x[b * 32 + lane];
        const float vf = __half2float(v);
        float amax = fabsf(vf);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, s));
        __half d = Q8_1 ? __float2half_rn(0.f) : __float2half_rn(1.f);
        if (amax != 0.f) d = __float2half_rn(amax / 127.f);
        const float dsafe = (Q8_1 && __half2float(d) == 0.f) ? 1.f : __half2float(d);
        const __half qh = hrint(__float2half_rn(vf / dsafe));  // fp16 quotient, then round-half-even
        int q = __half2int_rn(qh);
        q = max(-127, min(127, q));
        uint8_t* blk = out + b * BLK;
        blk[(Q8_1 ? 4 : 2) + lane] = static_cast<uint8_t>(static_cast<int8_t>(q));
        if (Q8_1) {
            int sum = q;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
            if (lane == 0) {
                // fp16(sum) (torch: int32 -> float16), then the fp16 product d * fp16(sum): exact in fp32, rounded once
                const float sum16 = __half2float(__float2half_rn(static_cast<float>(sum)));
                const __half sh = __float2half_rn(__half2float(d) * sum16);
                // 16-bit stores of the raw bit patterns (blocks are 2-byte aligned).  NB: writing `bits & 0xff` byte by
                // byte made nvcc 12.9 emit F2I.U8.F16 (a VALUE conversion of the half) for the low byte.
                reinterpret_cast<unsigned short*>(blk)[0] = __half_as_ushort(d);
                reinterpret_cast<unsigned short*>(blk)[1] = __half_as_ushort(sh);
            }
        } else if (lane == 0) {
            reinterpret_cast<unsigned short*>(blk)[0] = __half_as_ushort(d);
        }
    }
}

__global__ void __launch_bounds__(256) dequant_q6k_f32_kernel(const uint8_t* __restrict__ W, float* __restrict__ out,
                                                              int64_t n) {
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < n;
         e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint8_t* blk = W + (e / 256) * 210;
        const int k = static_cast<int>(e % 256);
        const float d = __half2float(load_half_bytes(blk + 208));
        const float ds = d * static_cast<float>(static_cast<int8_t>(blk[192 + (k >> 4)]));
        out[e] = ds * static_cast<float>(q6k_quant(blk, k));  // exact (q6_k.py:126-135)
    }
}

}  // namespace ggq

using namespace ggq;

extern "C" {

int ggq_quantize_q8_0_f16(const void* x, void* out, int64_t n, void* stream) {
    if (n < 0 || n % 32 != 0) return GGQ_E_SHAPE;
    if (n == 0) return 0;
    if (!x || !out) return GGQ_E_POINTER;
    const int64_t nb = n / 32;
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>((nb + 7) / 8, static_cast<int64_t>(num_sms()) * 16));
    quantize_q8_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(x),
                                                                                  static_cast<uint8_t*>(out), nb);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

int ggq_quantize_q8_1_f16(const void* x, void* out, int64_t n, void* stream) {
    if (n < 0 || n % 32 != 0) return GGQ_E_SHAPE;
    if (n == 0) return 0;
    if (!x || !out) return GGQ_E_POINTER;
    const int64_t nb = n / 32;
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>((nb + 7) / 8, static_cast<int64_t>(num_sms()) * 16));
    quantize_q8_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(x),
                                                                                 static_cast<uint8_t*>(out), nb);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

int ggq_dequant_q6_k_f32(const void* W, void* out, int64_t O, int64_t K, void* stream) {
    if (O < 0 || K < 0 || K % 256 != 0) return GGQ_E_SHAPE;
    if (O == 0 || K == 0) return 0;
    if (!W || !out) return GGQ_E_POINTER;
    const int64_t n = O * K;
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(num_sms()) * 16));
    dequant_q6k_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint8_t*>(W),
                                                                               static_cast<float*>(out), n);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
