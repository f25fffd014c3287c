// kquant_pack.cuh — Q4_K / Q6_K packers, byte-identical to the reference's compiled packers
// (utils/quantize/q4_k_ref.c:188-368 make_qkx2_quants + quantize_row_q4_K_ref, q6_k_ref.c:153-340 make_qx_quants +
// quantize_row_q6_K_ref).
//
// Byte identity means the same fp32 operations in the same order with the same roundings: every sum below runs over
// the elements in index order, products are grouped as the reference's C expressions group them, nothing may be
// contracted into an FMA (the .cu including this header is compiled with -fmad=false; the reference is plain
// `gcc -O2` x86-64, which has no FMA either), division and square root are the IEEE-rounded ones.
// What is new is the decomposition: the scale search of one sub-block (32 or 16 weights) is independent of every other
// sub-block, so ONE THREAD owns one sub-block with its weights in registers (quants packed 8 / 4 per 32-bit word), the
// 8 (Q4_K) or 16 (Q6_K) threads of a super-block agree on the super-block scale with warp shuffles and each writes its
// own bytes.  The same code compiles for the host (tests/host/kquant_host.cpp), where the "shuffles" are plain arrays;
// that build is what tests/test_kquant_pack_cpu.py compares with the reference library on the CPU.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#include <cuda_fp16.h>
#define KQ_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define KQ_HD inline
#endif

namespace ggq {
namespace kq {

KQ_HD int f2i_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    int i;
    memcpy(&i, &f, 4);
    return i;
#endif
}
KQ_HD float i2f_bits(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// round to nearest (ties to even) through the 1.5 * 2^23 trick, as q4_k_ref.c:165-172 does it (|v| <= 4194303)
KQ_HD int nearest(float v) {
    const float t = v + 12582912.f;
    return (f2i_bits(t) & 0x007fffff) - 0x00400000;
}
KQ_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// fp32 -> fp16 bits, round to nearest even (finite inputs; q4_k_ref.c:137-163 is the same rounding, bit-twiddled)
KQ_HD uint16_t f2h_bits(float f) {
#ifdef __CUDA_ARCH__
    return __half_as_ushort(__float2half_rn(f));
#else
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const int32_t exp = static_cast<int32_t>((x >> 23) & 0xff) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff) return static_cast<uint16_t>(sign | (man ? 0x7e00u : 0x7c00u));
    if (exp >= 31) return static_cast<uint16_t>(sign | 0x7c00u);
    if (exp <= 0) {
        if (exp < -10) return static_cast<uint16_t>(sign);
        man |= 0x800000u;
        const int shift = 14 - exp;
        uint32_t h = man >> shift;
        const uint32_t rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1))) ++h;
        return static_cast<uint16_t>(sign | h);
    }
    uint32_t h = (static_cast<uint32_t>(exp) << 10) | (man >> 13);
    const uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;
    return static_cast<uint16_t>(sign | h);
#endif
}
KQ_HD float h2f_bits(uint16_t h) {
#ifdef __CUDA_ARCH__
    return __half2float(__ushort_as_half(h));
#else
    const uint32_t sign = (h >> 15) & 1, exp = (h >> 10) & 0x1f, man = h & 0x3ff;
    float v;
    if (exp == 0) v = std::ldexp(static_cast<float>(man), -24);
    else if (exp == 31) v = man ? NAN : INFINITY;
    else v = std::ldexp(static_cast<float>(man | 0x400), static_cast<int>(exp) - 25);
    return sign ? -v : v;
#endif
}

// =====================================================================================================================
// Q4_K: sub-block of 32 weights -> (scale, min, 32 four-bit quants)
// =====================================================================================================================
struct Sub4 {
    float scale, minv;   // w ~ scale * q - minv, minv >= 0
    uint32_t q[4];       // nibble i of the sub-block in bits 4*(i%8) of word i/8
};
KQ_HD void put4(uint32_t (&q)[4], int i, int v) { q[i >> 3] = (q[i >> 3] & ~(0xFu << (4 * (i & 7)))) | (static_cast<uint32_t>(v) << (4 * (i & 7))); }
KQ_HD int get4(const uint32_t (&q)[4], int i) { return static_cast<int>((q[i >> 3] >> (4 * (i & 7))) & 0xFu); }

// weighted least-squares search over 21 candidate grids (q4_k_ref.c:188-279 with n = 32, nmax = 15, rmin = -1,
// rdelta = 0.1, nstep = 20, squared error; the weights are av_x + |x|, q4_k_ref.c:302-307)
KQ_HD Sub4 search_q4(const float (&x)[32]) {
    Sub4 r;
    float sum_x2 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) sum_x2 += x[i] * x[i];
    const float av = sqrtf(sum_x2 / 32);
    float lo = x[0], hi = x[0];
    float sum_w = av + fabsf(x[0]);
    float sum_x = sum_w * x[0];
#pragma unroll
    for (int i = 1; i < 32; ++i) {
        if (x[i] < lo) lo = x[i];
        if (x[i] > hi) hi = x[i];
        const float w = av + fabsf(x[i]);
        sum_w += w;
        sum_x += w * x[i];
    }
    if (lo > 0) lo = 0;
    r.q[0] = r.q[1] = r.q[2] = r.q[3] = 0u;
    if (hi == lo) {
        r.scale = 0.f;
        r.minv = -lo;
        return r;
    }
    float iscale = 15 / (hi - lo);
    float scale = 1 / iscale;
    float best = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int l = clampi(nearest(iscale * (x[i] - lo)), 0, 15);
        put4(r.q, i, l);
        float diff = scale * l + lo - x[i];
        diff = diff * diff;
        const float w = av + fabsf(x[i]);
        best += w * diff;
    }
    for (int is = 0; is <= 20; ++is) {
        iscale = (-1.f + 0.1f * is + 15) / (hi - lo);
        float sum_l = 0.f, sum_l2 = 0.f, sum_xl = 0.f;
        uint32_t aux[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int l = clampi(nearest(iscale * (x[i] - lo)), 0, 15);
            aux[i >> 3] |= static_cast<uint32_t>(l) << (4 * (i & 7));
            const float w = av + fabsf(x[i]);
            sum_l += w * l;
            sum_l2 += w * l * l;
            sum_xl += w * l * x[i];
        }
        const float D = sum_w * sum_l2 - sum_l * sum_l;
        if (D > 0) {
            float this_scale = (sum_w * sum_xl - sum_x * sum_l) / D;
            float this_min = (sum_l2 * sum_x - sum_l * sum_xl) / D;
            if (this_min > 0) {
                this_min = 0;
                this_scale = sum_xl / sum_l2;
            }
            float cur = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float diff = this_scale * get4(aux, i) + this_min - x[i];
                diff = diff * diff;
                const float w = av + fabsf(x[i]);
                cur += w * diff;
            }
            if (cur < best) {
                r.q[0] = aux[0];
                r.q[1] = aux[1];
                r.q[2] = aux[2];
                r.q[3] = aux[3];
                best = cur;
                scale = this_scale;
                lo = this_min;
            }
        }
    }
    r.scale = scale;
    r.minv = -lo;
    return r;
}

// Super-block step for one sub-block, given the super-block maxima (q4_k_ref.c:321-358): 6-bit scale / min codes and
// the final quants against the fp16-rounded d, dmin.
struct Code4 {
    uint8_t ls, lm;
};
KQ_HD Code4 code_q4(const Sub4& s, float max_scale, float max_min) {
    const float inv_scale = max_scale > 0 ? 63.f / max_scale : 0.f;
    const float inv_min = max_min > 0 ? 63.f / max_min : 0.f;
    uint8_t ls = static_cast<uint8_t>(nearest(inv_scale * s.scale));
    uint8_t lm = static_cast<uint8_t>(nearest(inv_min * s.minv));
    if (ls > 63) ls = 63;
    if (lm > 63) lm = 63;
    return Code4{ls, lm};
}
KQ_HD void requant_q4(const float (&x)[32], Sub4& s, Code4 c, uint16_t d_bits, uint16_t dmin_bits) {
    const float d = h2f_bits(d_bits) * c.ls;
    if (!d) return;  // keeps the quants of the search
    const float dm = h2f_bits(dmin_bits) * c.lm;
#pragma unroll
    for (int i = 0; i < 32; ++i) put4(s.q, i, clampi(nearest((x[i] + dm) / d), 0, 15));
}
// the 12 scale bytes of a block_q4_K from the eight (ls, lm) pairs (q4_k_ref.c:327-338)
KQ_HD void scale_bytes_q4(const uint8_t (&ls)[8], const uint8_t (&lm)[8], uint8_t (&out)[12]) {
    for (int j = 0; j < 12; ++j) out[j] = 0;
    for (int j = 0; j < 8; ++j) {
        if (j < 4) {
            out[j] = ls[j];
            out[j + 4] = lm[j];
        } else {
            out[j + 4] = static_cast<uint8_t>((ls[j] & 0xF) | ((lm[j] & 0xF) << 4));
            out[j - 4] |= static_cast<uint8_t>((ls[j] >> 4) << 6);
            out[j] |= static_cast<uint8_t>((lm[j] >> 4) << 6);
        }
    }
}

// =====================================================================================================================
// Q6_K: sub-block of 16 weights -> (scale, 16 six-bit quants stored as q + 32)
// =====================================================================================================================
struct Sub6 {
    float scale;
    uint32_t q[4];  // byte i of the sub-block in word i/4
};
KQ_HD void put8(uint32_t (&q)[4], int i, int v) { q[i >> 2] = (q[i >> 2] & ~(0xFFu << (8 * (i & 3)))) | (static_cast<uint32_t>(v) << (8 * (i & 3))); }
KQ_HD int get8(const uint32_t (&q)[4], int i) { return static_cast<int>((q[i >> 2] >> (8 * (i & 3))) & 0xFFu); }

constexpr float GROUP_EPS = 1e-15f;

// q6_k_ref.c:153-249 with n = 16, nmax = 32, rmse_type = 1 (weights x^2), no external weights
KQ_HD Sub6 search_q6(const float (&x)[16]) {
    Sub6 r;
    r.q[0] = r.q[1] = r.q[2] = r.q[3] = 0u;
    float top = 0.f, atop = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float ax = fabsf(x[i]);
        if (ax > atop) {
            atop = ax;
            top = x[i];
        }
    }
    if (atop < GROUP_EPS) {
        r.scale = 0.f;
        return r;
    }
    float iscale = -32 / top;
    float sumlx = 0.f, suml2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int l = clampi(nearest(iscale * x[i]), -32, 31);
        put8(r.q, i, l + 32);
        const float w = x[i] * x[i];
        sumlx += w * x[i] * l;
        suml2 += w * l * l;
    }
    float scale = suml2 ? sumlx / suml2 : 0.0f;
    float best = scale * sumlx;
    for (int is = -9; is <= 9; ++is) {
        if (is == 0) continue;
        iscale = -(32 + 0.1f * is) / top;
        sumlx = suml2 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int l = clampi(nearest(iscale * x[i]), -32, 31);
            const float w = x[i] * x[i];
            sumlx += w * x[i] * l;
            suml2 += w * l * l;
        }
        if (suml2 > 0 && sumlx * sumlx > best * suml2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) put8(r.q, i, 32 + clampi(nearest(iscale * x[i]), -32, 31));
            scale = sumlx / suml2;
            best = scale * sumlx;
        }
    }
    r.scale = scale;
    return r;
}
// super-block step (q6_k_ref.c:280-301): int8 scale code against iscale = -128 / max_scale, final quants against fp16 d
KQ_HD int8_t code_q6(float scale, float iscale) {
    const int v = nearest(iscale * scale);
    return static_cast<int8_t>(v < 127 ? v : 127);
}
KQ_HD void requant_q6(const float (&x)[16], Sub6& s, int8_t code, uint16_t d_bits) {
    const float d = h2f_bits(d_bits) * code;
    if (!d) return;
#pragma unroll
    for (int i = 0; i < 16; ++i) put8(s.q, i, clampi(nearest(x[i] / d), -32, 31) + 32);
}

}  // namespace kq
}  // namespace ggq
