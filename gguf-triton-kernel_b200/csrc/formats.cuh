// formats.cuh — GGUF block layouts and the bit-exact dequantization recipes.
//
// Layouts (reference: kernels/mmq_q4_k.py:1-16, kernels/mmq_q6_k.py:1-14, utils/quantize/q6_k.py:15-19,
// utils/quantize/q4_k_ref.c:76-89, q6_k_ref.c:62-68):
//   Q8_0   34 B: [0:2] fp16 d            [2:34] int8 qs[32]                         w = d*q
//   Q4_K  144 B: [0:2] fp16 d  [2:4] fp16 dmin  [4:16] 6-bit sc[8]/m[8]  [16:144] qs  w = d*sc*q - dmin*m
//   Q6_K  210 B: [0:128] ql  [128:192] qh  [192:208] int8 sc[16]  [208:210] fp16 d  w = d*sc*(q-32)
//
// Bit-exact recipes (checked against the reference dequantizers, SURVEY §8c / tests):
//   Q8_0: RN16(d*q)                         one fp16 multiply (utils/quantize/q8_0.py:94)
//   Q4_K: RN16(RN32((d*sc)*q - dmin*m))     fp32; d*sc, *q and dmin*m are exact, so one FMA gives the
//                                           same bits as numpy's mul-then-sub (utils/quantize/q4_k.py:137-143,156)
//   Q6_K: RN16((d*sc)*(q-32))               product exact in fp32 (utils/quantize/q6_k.py:126-135)
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace ggq {

struct Q8_0 {
    static constexpr int ID = 0;
    static constexpr int QK = 32;    // weights per block
    static constexpr int BLK = 34;   // bytes per block
};
struct Q4_K {
    static constexpr int ID = 1;
    static constexpr int QK = 256;
    static constexpr int BLK = 144;
};
struct Q6_K {
    static constexpr int ID = 2;
    static constexpr int QK = 256;
    static constexpr int BLK = 210;
};

__host__ __device__ inline int fmt_qk(int fmt) { return fmt == 0 ? 32 : 256; }
__host__ __device__ inline int fmt_blk(int fmt) { return fmt == 0 ? 34 : (fmt == 1 ? 144 : 210); }

__device__ __forceinline__ __half half_from_bits(uint32_t lo16) {
    return __ushort_as_half(static_cast<unsigned short>(lo16 & 0xffffu));
}
__device__ __forceinline__ __half load_half_bytes(const uint8_t* p) {  // 1-byte aligned
    return half_from_bits(static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8));
}

// 6-bit scale / min of sub-block j (0..7) from the 12 scale bytes (q4_k_ref.c:174-186).
__device__ __forceinline__ void q4k_scale_min(const uint8_t* s, int j, int& sc, int& m) {
    if (j < 4) {
        sc = s[j] & 63;
        m = s[j + 4] & 63;
    } else {
        sc = (s[j + 4] & 0x0F) | ((s[j - 4] >> 6) << 4);
        m = (s[j + 4] >> 4) | ((s[j] >> 6) << 4);
    }
}

// ---- scalar, alignment-free element accessors (dequant op, generic GEMM) --------------------
// `blk` points at the block holding element e (0 <= e < QK); returns the fp16-rounded weight.

__device__ __forceinline__ __half dequant_elem(Q8_0, const uint8_t* blk, int e) {
    const __half d = load_half_bytes(blk);
    const int q = static_cast<int8_t>(blk[2 + e]);
    return __hmul(d, __int2half_rn(q));
}

__device__ __forceinline__ __half dequant_elem(Q4_K, const uint8_t* blk, int e) {
    const float d = __half2float(load_half_bytes(blk));
    const float dmin = __half2float(load_half_bytes(blk + 2));
    const int j = e >> 5;  // sub-block
    int sc, m;
    q4k_scale_min(blk + 4, j, sc, m);
    // chunk c = e/64: byte 16 + 32c + l; low nibble -> element 64c + l, high nibble -> 64c + 32 + l
    const int byte = blk[16 + ((e >> 6) << 5) + (e & 31)];
    const int q = (e & 32) ? (byte >> 4) : (byte & 0x0F);
    const float ds = d * static_cast<float>(sc);      // exact (11 x 6 bits)
    const float dm = dmin * static_cast<float>(m);    // exact
    return __float2half_rn(fmaf(ds, static_cast<float>(q), -dm));
}

__device__ __forceinline__ int q6k_quant(const uint8_t* blk, int e) {  // returns q - 32
    const int h = e >> 7;          // half of the super-block
    const int r = e & 127;
    const int g = r >> 5;          // 0..3: which (ql nibble, qh bit pair)
    const int l = r & 31;
    const int qlb = blk[64 * h + ((g & 1) << 5) + l];
    const int lo = (g & 2) ? (qlb >> 4) : (qlb & 0x0F);
    const int hi = (blk[128 + 32 * h + l] >> (2 * g)) & 3;
    return (lo | (hi << 4)) - 32;
}

__device__ __forceinline__ __half dequant_elem(Q6_K, const uint8_t* blk, int e) {
    const float d = __half2float(load_half_bytes(blk + 208));
    const int sc = static_cast<int8_t>(blk[192 + (e >> 4)]);
    const float ds = d * static_cast<float>(sc);  // exact
    return __float2half_rn(ds * static_cast<float>(q6k_quant(blk, e)));  // product exact, one rounding
}

}  // namespace ggq
