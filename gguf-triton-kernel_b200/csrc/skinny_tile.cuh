// skinny_tile.cuh — decode-grade dequantization of one row x 64 consecutive weights for the skinny family.
//
// Same interface and output order as prefill_tile.cuh's dequant64 (out[c] = weights 8c .. 8c+7 as fp16, K order), but
// built from packed half2 arithmetic so that it fits the instruction budget of an HBM-bound kernel (Q4_K at 6 TB/s
// leaves ~3.4 lane-operations per weight in total; the bit-exact fp32 recipes cost 2.75 - 3.5):
//
//   Q8_0  RN16(d * q)                          = the reference value (prefill_tile.cuh's recipe is already half2)
//   Q4_K  RN16(d * (sc*q) - RN16(dmin*m))      1.75 op / weight.  sc*q is an exact fp16 integer (<= 945): one HFMA2 turns
//                                              the magic-biased nibble (1024 + q, or 64 + q for high nibbles) into sc*q,
//                                              a second one applies d and the min.  Differs from the reference
//                                              RN16(RN32(d*sc*q - dmin*m)) only by the fp16 rounding of dmin*m.
//   Q6_K  RN16(RN16(d*sc) * (q6 - 32))         2.4 op / weight
// These values feed tcgen05.mma directly and are never stored, so they are held to the matmul tolerance (Tier 1,
// measured ~4e-4 rel. Frobenius), not to the bit-exactness of the standalone dequantize ops.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "prefill_tile.cuh"

namespace ggq {
namespace skn {

__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }

// ---- Q4_K: kb = 0..3, sub-blocks 2kb (low nibbles of qs[32kb ..]) -> out[0..3], 2kb+1 (high nibbles) -> out[4..7] ----
__device__ __forceinline__ void dq64_q4_k(const uint8_t* p, int kb, uint4 (&out)[8]) {
    const uint4 h = *reinterpret_cast<const uint4*>(p);
    const __half2 d2 = u2h(__byte_perm(h.x, 0u, 0x1010));
    const float dmin = pre::hbits2f(h.x >> 16);
    const uint4* qs = reinterpret_cast<const uint4*>(p + 16 + 32 * kb);
    const uint4 qa = qs[0], qb = qs[1];
    const uint32_t w[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        uint32_t sc, m;
        if (kb < 2) {  // sub-blocks 0..3: plain 6-bit fields (q4_k_ref.c:176-178)
            const int s = 16 * kb + 8 * hf;
            sc = (h.y >> s) & 63u;
            m = (h.z >> s) & 63u;
        } else {       // sub-blocks 4..7 (:180-183)
            const int s = 16 * (kb - 2) + 8 * hf;
            sc = ((h.w >> s) & 0xFu) | (((h.y >> (s + 6)) & 3u) << 4);
            m = ((h.w >> (s + 4)) & 0xFu) | (((h.z >> (s + 6)) & 3u) << 4);
        }
        const float scf = static_cast<float>(sc);
        const __half2 sc2 = __float2half2_rn(scf);                                   // exact (<= 63)
        const __half2 nb2 = __float2half2_rn(scf * (hf ? -64.f : -1024.f));          // exact (multiple of 64 / 1024)
        const __half2 nc2 = __float2half2_rn(-(dmin * static_cast<float>(m)));       // RN16(-(dmin * m))
        const uint32_t mask = hf ? 0xF0F0F0F0u : 0x0F0F0F0Fu;
        const uint32_t magic = hf ? 0x54545454u : 0x64646464u;                       // halves 64 + q / 1024 + q
        uint4* o = out + 4 * hf;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t t0 = w[2 * c] & mask, t1 = w[2 * c + 1] & mask;
            const __half2 a0 = __hfma2(u2h(__byte_perm(t0, magic, 0x4140)), sc2, nb2);   // sc * q, exact
            const __half2 a1 = __hfma2(u2h(__byte_perm(t0, magic, 0x4342)), sc2, nb2);
            const __half2 a2 = __hfma2(u2h(__byte_perm(t1, magic, 0x4140)), sc2, nb2);
            const __half2 a3 = __hfma2(u2h(__byte_perm(t1, magic, 0x4342)), sc2, nb2);
            o[c].x = h2u(__hfma2(a0, d2, nc2));
            o[c].y = h2u(__hfma2(a1, d2, nc2));
            o[c].z = h2u(__hfma2(a2, d2, nc2));
            o[c].w = h2u(__hfma2(a3, d2, nc2));
        }
    }
}

// ---- Q6_K: kb = 0..3 (half h = kb >> 1, groups 2gp, 2gp+1 with gp = kb & 1), see prefill_tile.cuh dequant_q6_k_core ----
// lw[8 gi + i] = the 32 ql bytes of group 2gp + gi, hw = the 32 qh bytes of half h, scw = 4 int8 scales, d = block scale
__device__ __forceinline__ void dq64_q6_k_core(const uint32_t (&lw)[16], const uint32_t (&hw)[8], uint32_t scw, float d, int gp,
                                               uint4 (&out)[8]) {
    __half2 a2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        a2[i] = __float2half2_rn(d * static_cast<float>(static_cast<int>(static_cast<int8_t>((scw >> (8 * i)) & 0xffu))));
    const __half2 bias = __float2half2_rn(1056.f);  // (1024 + q) - 1056 = q - 32, exact
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {  // group g = 2gp + gi
        const uint32_t rot = static_cast<uint32_t>(4 * gp + 2 * gi - 4) & 31u;
        uint4* o = out + 4 * gi;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {  // 8 weights of sub-block 2*gi + (c4 >> 1) of the four
            const __half2 s2 = a2[2 * gi + (c4 >> 1)];
            uint32_t r[4];
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const uint32_t lo = (lw[8 * gi + 2 * c4 + v] >> (4 * gp)) & 0x0F0F0F0Fu;
                const uint32_t hr = __funnelshift_r(hw[2 * c4 + v], hw[2 * c4 + v], rot);
                const uint32_t q = (hr & 0x30303030u) | lo;  // four 6-bit quants
                r[2 * v] = h2u(__hmul2(__hsub2(u2h(__byte_perm(q, 0x64646464u, 0x4140)), bias), s2));
                r[2 * v + 1] = h2u(__hmul2(__hsub2(u2h(__byte_perm(q, 0x64646464u, 0x4342)), bias), s2));
            }
            o[c4] = make_uint4(r[0], r[1], r[2], r[3]);
        }
    }
}
// block at a 2-byte aligned address (GGQ_SKINNY_EXACT builds and tests of the arithmetic)
template <bool ODD>
__device__ __forceinline__ void dq64_q6_k_al(const uint8_t* b, int kb, uint4 (&out)[8]) {
    const int h = kb >> 1, gp = kb & 1;
    const float d = pre::hbits2f(*reinterpret_cast<const uint16_t*>(b + 208));
    const uint32_t scw = pre::ld32_any(b + 192 + 8 * h + 4 * gp);  // scales of sub-blocks 8h + 4gp + 0..3
    uint32_t hw[8], lw[16], t[8];
    pre::ld8w<ODD>(b + 128 + 32 * h, hw);
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
        pre::ld8w<ODD>(b + 64 * h + 32 * gi, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) lw[8 * gi + i] = t[i];
    }
    dq64_q6_k_core(lw, hw, scw, d, gp, out);
}
// p: 16-byte aligned superset of the block column in this row; off: byte offset of the block inside it
template <int FMT>
__device__ __forceinline__ void dq64(const uint8_t* p, int off, int kb, uint4 (&out)[8]) {
#ifdef GGQ_SKINNY_EXACT
    pre::dequant64(pre::Unit<FMT>{}, p, off, kb, out);
#else
    if constexpr (FMT == 0) {
        pre::dequant_q8_0<2>(p, off, kb, out);
    } else if constexpr (FMT == 1) {
        dq64_q4_k(p, kb, out);
    } else {
        // (dq64_q6_k_sm, the conflict-free loader that sped the prefill kernel up by 18 %, measured slower here: unrolled, its
        // alignment variants thrash the instruction cache (173 us at T=16 on the lm_head), rolled it loses the overlap of
        // one chunk's loads with the previous chunk's arithmetic (117 us); this loader: 112 us)
        const uint8_t* b = p + off;
        if (off & 2) dq64_q6_k_al<true>(b, kb, out);
        else dq64_q6_k_al<false>(b, kb, out);
    }
#endif
}

}  // namespace skn
}  // namespace ggq
