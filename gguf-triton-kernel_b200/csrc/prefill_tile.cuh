// prefill_tile.cuh — bit-exact vectorized dequantization of one row x 64 consecutive weights.
//
// Used by BOTH the prefill GEMM (source = shared-memory staging units filled by TMA) and the
// standalone dequantize op ggq_dequant_*_f16 (source = global memory), so the Tier-0 test
// "dequantized weights are bit-exact" checks exactly the values that are fed to tcgen05.mma.
//
// A "unit row" is the 16-byte aligned superset of the blocks of one weight row that cover UNIT_K
// consecutive weights; `p` points at the unit row, `off` is the byte offset of the first block inside
// it (0 for Q4_K; 0/8 for Q8_0; 0,2,..,14 for Q6_K — identical for every row of a block column because
// rows are whole 16-byte vectors).  out[c] holds weights 8c..8c+7 of the 64 as fp16.
//
// Recipes (reference: utils/quantize/q8_0.py:94, q4_k.py:137-143,156, q6_k.py:126-135):
//   Q8_0  RN16(d * q)                      one fp16 multiply
//   Q4_K  RN16(fma(d*sc, q, -(dmin*m)))    fp32, products exact
//   Q6_K  RN16((d*sc) * (q6 - 32))         fp32 product exact; written as a multiply (not an FMA against
//                                          -32*d*sc) so that zeros keep the reference's sign: (-s)*0 = -0
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace ggq {
namespace pre {

template <int FMT> struct Unit;
template <> struct Unit<0> {  // Q8_0: 4 blocks = 128 weights = 136 B; block column starts at 136c (0 or 8 mod 16)
    static constexpr int QK = 32, BLK = 34, UNIT_K = 128, UNIT_BYTES = 136, BOX_BYTES = 144;
};
template <> struct Unit<1> {  // Q4_K: 1 super-block = 256 weights = 144 B, always 16-byte aligned
    static constexpr int QK = 256, BLK = 144, UNIT_K = 256, UNIT_BYTES = 144, BOX_BYTES = 144;
};
template <> struct Unit<2> {  // Q6_K: 1 super-block = 256 weights = 210 B at 210c (even offsets 0..14 mod 16); the box is
                              // 240 B = an odd number of 16-byte vectors per row (dequant_q6_k_sm)
    static constexpr int QK = 256, BLK = 210, UNIT_K = 256, UNIT_BYTES = 210, BOX_BYTES = 240;
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);  // cvt.rn.f16x2.f32
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float hbits2f(uint32_t bits) {
    return __half2float(__ushort_as_half(static_cast<unsigned short>(bits)));
}
__device__ __forceinline__ float byte2f(uint32_t w, int i) { return static_cast<float>((w >> (8 * i)) & 0xffu); }

// ---- packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2 process two floats per instruction) ----------------
struct f2 { float x, y; };
__device__ __forceinline__ unsigned long long f2_bits(f2 v) {
    return (static_cast<unsigned long long>(__float_as_uint(v.y)) << 32) | __float_as_uint(v.x);
}
__device__ __forceinline__ f2 bits_f2(unsigned long long b) {
    return f2{__uint_as_float(static_cast<uint32_t>(b)), __uint_as_float(static_cast<uint32_t>(b >> 32))};
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return bits_f2(r);
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return bits_f2(r);
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
    return bits_f2(r);
}
// bytes i, i+1 of w as the floats 2^23 + byte (one PRMT each, no integer->float conversion)
__device__ __forceinline__ f2 magic2(uint32_t w, int i) {
    return f2{__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650 + i)), __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7651 + i))};
}

// 32-bit load at an address that is only 2-byte aligned (alignment is uniform across the warp)
__device__ __forceinline__ uint32_t ld32_any(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~uintptr_t{3});
    const uint32_t sh = static_cast<uint32_t>(a & 3) * 8;
    if (sh == 0) return q[0];
    return __funnelshift_r(q[0], q[1], sh);
}

// All three are templated on HALF: 2 = all 64 weights into out[0..7]; 0 / 1 = only the first / second 32
// weights into out[0..3] (the 2-CTA prefill kernel splits a row's k-block over two threads).

// ---- Q8_0 -------------------------------------------------------------------------------------
// The arithmetic on the 17 words of two consecutive blocks (68 bytes from a 4-byte aligned start): w[0] = {d_A, q0 q1},
// w[8] = {q30 q31 of A, d_B}, w[9..16] = the quants of B.
template <int HALF>
__device__ __forceinline__ void dequant_q8_0_core(const uint32_t (&w)[17], uint4* out) {
    const __half2 bias = __float2half2_rn(1152.f);
    auto cvt = [&](uint32_t q4, const __half2 d2, uint32_t& lo, uint32_t& hi) {
        const uint32_t u = q4 ^ 0x80808080u;                        // q + 128
        uint32_t a = __byte_perm(u, 0x64646464u, 0x5140);           // halves 1024 + (q + 128)
        uint32_t b = __byte_perm(u, 0x64646464u, 0x7362);
        __half2 ha = __hsub2(*reinterpret_cast<__half2*>(&a), bias);  // exact q
        __half2 hb = __hsub2(*reinterpret_cast<__half2*>(&b), bias);
        ha = __hmul2(ha, d2);                                       // RN16(d * q)
        hb = __hmul2(hb, d2);
        lo = *reinterpret_cast<uint32_t*>(&ha);
        hi = *reinterpret_cast<uint32_t*>(&hb);
    };
    if (HALF != 1) {  // block A: quants start 2 bytes into word 0
        const __half2 d2 = __half2half2(__ushort_as_half(static_cast<unsigned short>(w[0] & 0xffffu)));
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t v0 = __funnelshift_r(w[2 * c], w[2 * c + 1], 16);
            const uint32_t v1 = __funnelshift_r(w[2 * c + 1], w[2 * c + 2], 16);
            cvt(v0, d2, out[c].x, out[c].y);
            cvt(v1, d2, out[c].z, out[c].w);
        }
    }
    if (HALF != 0) {  // block B: quants are word aligned (34 + 2 = 36)
        uint4* o = out + (HALF == 2 ? 4 : 0);
        const __half2 d2 = __half2half2(__ushort_as_half(static_cast<unsigned short>(w[8] >> 16)));
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            cvt(w[9 + 2 * c], d2, o[c].x, o[c].y);
            cvt(w[10 + 2 * c], d2, o[c].z, o[c].w);
        }
    }
}
// loader 1: 32-bit loads from the 4-byte aligned start (global memory; the skinny kernel)
// kb: which 64-weight half of the 128-weight unit (blocks 2kb, 2kb+1)
template <int HALF>
__device__ __forceinline__ void dequant_q8_0(const uint8_t* p, int off, int kb, uint4* out) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p + off + 68 * kb);  // 4-byte aligned; 17 words
    uint32_t w[17];
#pragma unroll
    for (int i = 0; i < 17; ++i) w[i] = (HALF == 0 && i > 8) || (HALF == 1 && i < 8) ? 0u : q[i];
    dequant_q8_0_core<HALF>(w, out);
}
// loader 2 (shared memory, thread = row, 144-byte rows = 9 vectors): 128-bit loads of aligned vectors, conflict free,
// and a warp-uniform word shift (the 68-byte pair starts 0, 4, 8 or 12 bytes into a vector).  With loader 1 the 17
// 32-bit loads of 32 rows hit 8 banks: 4-way conflicts, 61 % of the Q8_0 prefill kernel's shared-memory wavefronts.
template <int WS>
__device__ __forceinline__ void q8_0_gather(const uint4* v, uint32_t (&w)[17]) {
    uint32_t a[20];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const uint4 t = v[i];
        a[4 * i] = t.x; a[4 * i + 1] = t.y; a[4 * i + 2] = t.z; a[4 * i + 3] = t.w;
    }
#pragma unroll
    for (int i = 0; i < 17; ++i) w[i] = a[WS + i];
}
template <int HALF>
__device__ __forceinline__ void dequant_q8_0_sm(const uint8_t* row16, int off, int kb, uint4* out) {
    const int o = off + 68 * kb;   // 0, 8, 68 or 76
    const uint4* v = reinterpret_cast<const uint4*>(row16 + (o & ~15));
    uint32_t w[17];
    switch ((o & 15) >> 2) {
        case 0: q8_0_gather<0>(v, w); break;
        case 1: q8_0_gather<1>(v, w); break;
        case 2: q8_0_gather<2>(v, w); break;
        default: q8_0_gather<3>(v, w); break;
    }
    dequant_q8_0_core<HALF>(w, out);
}

// ---- Q4_K -------------------------------------------------------------------------------------
// kb: 0..3, the 64-weight chunk of the super-block (sub-blocks 2kb = low nibbles, 2kb+1 = high nibbles)
template <int HALF>
__device__ __forceinline__ void dequant_q4_k(const uint8_t* p, int kb, uint4* out) {
    const uint4 h = *reinterpret_cast<const uint4*>(p);
    const float d = hbits2f(h.x & 0xffffu), dmin = hbits2f(h.x >> 16);
    const uint4* qs = reinterpret_cast<const uint4*>(p + 16 + 32 * kb);
    const uint4 qa = qs[0], qb = qs[1];
    const uint32_t w[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        if (HALF != 2 && HALF != hf) continue;
        uint32_t sc, m;
        if (kb < 2) {  // sub-blocks 0..3: plain 6-bit fields (q4_k_ref.c:176-178)
            const int s = 16 * kb + 8 * hf;
            sc = (h.y >> s) & 63u;
            m = (h.z >> s) & 63u;
        } else {       // sub-blocks 4..7: low 4 bits in bytes 8..11, high 2 bits in the top of bytes 0..7 (:180-183)
            const int s = 16 * (kb - 2) + 8 * hf;
            sc = ((h.w >> s) & 0xFu) | (((h.y >> (s + 6)) & 3u) << 4);
            m = ((h.w >> (s + 4)) & 0xFu) | (((h.z >> (s + 6)) & 3u) << 4);
        }
        const float ds = d * static_cast<float>(sc);      // exact
        const float dm = -(dmin * static_cast<float>(m));  // exact
        const f2 ds2{ds, ds}, dm2{dm, dm}, m23{-8388608.f, -8388608.f};
        uint4* o = out + ((HALF == 2 && hf == 1) ? 4 : 0);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t q0 = (hf ? (w[2 * c] >> 4) : w[2 * c]) & 0x0F0F0F0Fu;
            const uint32_t q1 = (hf ? (w[2 * c + 1] >> 4) : w[2 * c + 1]) & 0x0F0F0F0Fu;
            // q as float = (2^23 + q) - 2^23 (exact), then one FMA: RN32(ds*q - dm)
            const f2 a0 = fma2(ds2, add2(magic2(q0, 0), m23), dm2), a1 = fma2(ds2, add2(magic2(q0, 2), m23), dm2);
            const f2 a2 = fma2(ds2, add2(magic2(q1, 0), m23), dm2), a3 = fma2(ds2, add2(magic2(q1, 2), m23), dm2);
            o[c].x = pack2(a0.x, a0.y);
            o[c].y = pack2(a1.x, a1.y);
            o[c].z = pack2(a2.x, a2.y);
            o[c].w = pack2(a3.x, a3.y);
        }
    }
}

// ---- Q6_K -------------------------------------------------------------------------------------
// kb: 0..3; half h = kb >> 1, nibble/bit-pair selector gp = kb & 1 (groups g = 2gp, 2gp + 1 of 32 weights):
//   weight 128h + 32g + l = (ql[64h + 32(g&1) + l] nibble gp) | ((qh[32h + l] >> 2g) & 3) << 4   (q6_k_ref.c:320-336)
// 8 consecutive 32-bit words starting at a 2-byte aligned address (alignment uniform across the warp)
template <bool ODD> __device__ __forceinline__ void ld8w(const uint8_t* p, uint32_t w[8]) {
    if (!ODD) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = q[i];
    } else {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p - 2);
        uint32_t r[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) r[i] = q[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = __funnelshift_r(r[i], r[i + 1], 16);
    }
}

// The arithmetic, shared by every loader: lw[8 gi + i] = the 32 ql bytes of group g = 2gp + gi, hw = the 32 qh bytes of
// half h, scw = the int8 scales of sub-blocks 8h + 4gp + 0..3, d = the block scale.
template <int HALF>
__device__ __forceinline__ void dequant_q6_k_core(const uint32_t (&lw)[16], const uint32_t (&hw)[8], uint32_t scw, float d, int gp,
                                                  uint4* out) {
    float ds[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        ds[i] = d * static_cast<float>(static_cast<int>(static_cast<int8_t>((scw >> (8 * i)) & 0xffu)));  // exact
    const f2 m32{-8388640.f, -8388640.f};  // -(2^23 + 32): (2^23 + q) + this == q - 32 exactly
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {  // group g = 2gp + gi
        if (HALF != 2 && HALF != gi) continue;
        // the 2 qh bits of group g sit at bits 2g, 2g+1 of every byte: rotate them to bits 4, 5 (no bits of a
        // neighbouring byte can reach positions 4, 5: the rotation is by -4, -2, 0 or +2)
        const uint32_t rot = static_cast<uint32_t>(4 * gp + 2 * gi - 4) & 31u;
        uint4* o = out + ((HALF == 2 && gi == 1) ? 4 : 0);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {  // 8 weights: l = 8*c4 .. 8*c4+7, sub-block 2*gi + (c4 >> 1) of the four
            const float s = ds[2 * gi + (c4 >> 1)];
            const f2 s2{s, s};
            uint32_t r[4];
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const uint32_t lo = (lw[8 * gi + 2 * c4 + v] >> (4 * gp)) & 0x0F0F0F0Fu;
                const uint32_t hr = __funnelshift_r(hw[2 * c4 + v], hw[2 * c4 + v], rot);
                const uint32_t q = (hr & 0x30303030u) | lo;  // four 6-bit quants
                const f2 a = mul2(s2, add2(magic2(q, 0), m32)), c = mul2(s2, add2(magic2(q, 2), m32));
                r[2 * v] = pack2(a.x, a.y);
                r[2 * v + 1] = pack2(c.x, c.y);
            }
            o[c4] = make_uint4(r[0], r[1], r[2], r[3]);
        }
    }
}

// loader 1: block at a 2-byte aligned address (global memory / the 16-byte aligned supersets of the decode family)
template <int HALF, bool ODD>
__device__ __forceinline__ void dequant_q6_k_al(const uint8_t* b, int kb, uint4* out) {
    const int h = kb >> 1, gp = kb & 1;
    const float d = hbits2f(*reinterpret_cast<const uint16_t*>(b + 208));
    const uint32_t scw = ld32_any(b + 192 + 8 * h + 4 * gp);  // scales of sub-blocks 8h + 4gp + 0..3
    uint32_t hw[8], lw[16];
    ld8w<ODD>(b + 128 + 32 * h, hw);
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
        if (HALF != 2 && HALF != gi) continue;
        uint32_t t[8];
        ld8w<ODD>(b + 64 * h + 32 * gi, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) lw[8 * gi + i] = t[i];
    }
    dequant_q6_k_core<HALF>(lw, hw, scw, d, gp, out);
}

// loader 2 (shared memory, thread = row): `row16` = 16-byte aligned start of the row's staged bytes, the block begins `off`
// bytes in (even, 0..14, the same for every row of a block column, so the switch below is warp-uniform).  TMA can only
// start a box at a 16-byte aligned global address (an element-granular start faults: tools/probes/tma_probe.cu), so the
// block cannot be staged aligned.  With loader 1 the 32-bit loads of the 32 rows, a multiple of 16 bytes apart, fall into
// 4 or 8 banks: 4- to 8-way conflicts made the shared-memory pipe the bound of the Q6_K prefill and skinny kernels.
// Here the row pitch is an ODD multiple of 16 bytes (Q6K_ROW_PITCH) and every load is a 128-bit load of an aligned
// vector — conflict free with thread = row — and the words are shifted into place in registers.
constexpr int Q6K_ROW_PITCH = 240;
template <int WS, bool ODD>   // off = 4 WS + (ODD ? 2 : 0)
__device__ __forceinline__ void q6k_gather(const uint4* ql, const uint4* qh, uint32_t (&lw)[16], uint32_t (&hw)[8],
                                           const uint4* sc, int sel, uint32_t& scw, uint32_t& dbits) {
    {   // scales (block bytes 192..207) and d (208, 209): row bytes [192 + off, 210 + off) lie inside the two vectors sc[0..1]
        const uint4 s0 = sc[0], s1 = sc[1];
        const uint32_t t[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        uint32_t u[5];   // the block's words 48..52 (bytes 192..211)
#pragma unroll
        for (int i = 0; i < 5; ++i) u[i] = ODD ? __funnelshift_r(t[WS + i], t[WS + i + 1 < 8 ? WS + i + 1 : 7], 16) : t[WS + i];
        scw = sel == 0 ? u[0] : sel == 1 ? u[1] : sel == 2 ? u[2] : u[3];   // sel = 2h + gp
        dbits = u[4] & 0xffffu;
    }
    constexpr int NL = (WS == 0 && !ODD) ? 4 : 5, NH = (WS == 0 && !ODD) ? 2 : 3;
    uint32_t a[20], c[12];
#pragma unroll
    for (int i = 0; i < NL; ++i) {
        const uint4 v = ql[i];
        a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < NH; ++i) {
        const uint4 v = qh[i];
        c[4 * i] = v.x; c[4 * i + 1] = v.y; c[4 * i + 2] = v.z; c[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) lw[i] = ODD ? __funnelshift_r(a[WS + i], a[WS + i + 1], 16) : a[WS + i];
#pragma unroll
    for (int i = 0; i < 8; ++i) hw[i] = ODD ? __funnelshift_r(c[WS + i], c[WS + i + 1], 16) : c[WS + i];
}
template <int HALF>
__device__ __forceinline__ void dequant_q6_k_sm(const uint8_t* row16, int off, int kb, uint4* out) {
    const int h = kb >> 1, gp = kb & 1;
    const uint4* ql = reinterpret_cast<const uint4*>(row16 + 64 * h);
    const uint4* qh = reinterpret_cast<const uint4*>(row16 + 128 + 32 * h);
    const uint4* sc = reinterpret_cast<const uint4*>(row16 + 192);
    const int sel = 2 * h + gp;
    uint32_t lw[16], hw[8], scw, dbits;
    switch (off >> 1) {
        case 0: q6k_gather<0, false>(ql, qh, lw, hw, sc, sel, scw, dbits); break;
        case 1: q6k_gather<0, true>(ql, qh, lw, hw, sc, sel, scw, dbits); break;
        case 2: q6k_gather<1, false>(ql, qh, lw, hw, sc, sel, scw, dbits); break;
        case 3: q6k_gather<1, true>(ql, qh, lw, hw, sc, sel, scw, dbits); break;
        case 4: q6k_gather<2, false>(ql, qh, lw, hw, sc, sel, scw, dbits); break;
        case 5: q6k_gather<2, true>(ql, qh, lw, hw, sc, sel, scw, dbits); break;
        case 6: q6k_gather<3, false>(ql, qh, lw, hw, sc, sel, scw, dbits); break;
        default: q6k_gather<3, true>(ql, qh, lw, hw, sc, sel, scw, dbits); break;
    }
    dequant_q6_k_core<HALF>(lw, hw, scw, hbits2f(dbits), gp, out);
}

template <int HALF>
__device__ __forceinline__ void dequant_q6_k(const uint8_t* p, int off, int kb, uint4* out) {
    const uint8_t* b = p + off;  // 2-byte aligned; (off & 2) is uniform across the block column
    if (off & 2) dequant_q6_k_al<HALF, true>(b, kb, out);
    else dequant_q6_k_al<HALF, false>(b, kb, out);
}

template <int HALF> __device__ __forceinline__ void dequant_part(Unit<0>, const uint8_t* p, int off, int kb, uint4* out) {
    dequant_q8_0<HALF>(p, off, kb, out);
}
template <int HALF> __device__ __forceinline__ void dequant_part(Unit<1>, const uint8_t* p, int, int kb, uint4* out) {
    dequant_q4_k<HALF>(p, kb, out);
}
template <int HALF> __device__ __forceinline__ void dequant_part(Unit<2>, const uint8_t* p, int off, int kb, uint4* out) {
    dequant_q6_k<HALF>(p, off, kb, out);
}
template <class U> __device__ __forceinline__ void dequant64(U u, const uint8_t* p, int off, int kb, uint4 out[8]) {
    dequant_part<2>(u, p, off, kb, out);
}

}  // namespace pre
}  // namespace ggq
