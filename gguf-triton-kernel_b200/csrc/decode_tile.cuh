// decode_tile.cuh — lane-level code of the decode family (T <= 16 per pass).
//
// One warp owns a 16-row weight tile.  Packed row chunks sit in shared memory exactly as they lie in
// HBM (TMA bulk copies, no repacking); each lane unpacks the 2x4 weights it needs for one
// mma.sync.m16n8k16 (f16 x f16 -> f32) straight into the A fragment, the B fragment is read from the
// raw fp16 activation rows (also TMA-staged), and block / sub-block scales are applied to the fp32
// MMA result ("post-scaling"), so dequantized weights never exist outside registers:
//
//   Q8_0  A = (q + 128) * 2^-24,   D starts at -128 * 2^-24 * sum32(x);   acc += (d * 2^24) * D
//   Q4_K  A = q * 2^-24 (low nibbles) / q * 2^-20 (high nibbles, in place);
//         acc += (d*sc_j * 2^24|2^20) * D_j - (dmin*m_j) * sum32_j(x)
//   Q6_K  A = q6 * 2^-24,           D starts at -32 * 2^-24 * sum16(x);    acc += (d*sc_j * 2^24) * D_j
//
// Integer -> fp16 costs one LOP3/PRMT per weight PAIR: an integer < 1024 placed in the low mantissa
// bits of a half whose exponent field is zero IS the subnormal value n * 2^-24, which the tensor core
// multiplies exactly; the power of two is folded into the fp32 scale.  (The classic 0x6400 "1024 + n"
// magic is avoided on purpose: its bias has to be cancelled in fp32 afterwards, which costs ~10 bits of
// the accumulator — 8e-6 vs 2e-7 relative error in the host emulator — while the subnormal form keeps
// the addends at the magnitude of the signal.)  Products are exact and accumulate in fp32.  Because the K order
// inside an MMA is free as long as A and B agree, lane t of a quad always takes the 4 weights that share one 32-bit
// word and the 4 matching consecutive activations.
//
// Two tile codes per format: the general one (T = 2..16: the 8 MMA columns are 8 tokens, NT = 1 or 2 n-tiles) and the
// single-token GEMV one (`*_gv`: the 8 columns are 8 SUB-BLOCKS of the one token, see the notes above
// compute_q4_k_gv_impl), selected by the `GV` template flag of Tile<>.
//
// This header is also compiled for the host (tests/host/emu_decode.cpp) where a 32-thread warp
// emulator supplies mma/syncwarp, so the bit manipulation is verified against the oracle without a GPU.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#include <cuda_fp16.h>
#define GGQ_DEV __device__ __forceinline__
namespace ggq {
namespace dec {
GGQ_DEV uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
GGQ_DEV uint32_t funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
GGQ_DEV float h2f(uint32_t bits) { return __half2float(__ushort_as_half(static_cast<unsigned short>(bits))); }
GGQ_DEV uint32_t f2u(float f) { return __float_as_uint(f); }
GGQ_DEV float u2f(uint32_t u) { return __uint_as_float(u); }
GGQ_DEV float shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
GGQ_DEV void mma16816_bf16(float d[4], const uint32_t a[4], const uint32_t b[2], const float c[4]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};\n"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}
// D = A(16x16, row) * B(16x8, col) + C, fp16 inputs, fp32 accumulate (HMMA.16816.F32 on sm_100a)
GGQ_DEV void mma16816(float d[4], const uint32_t a[4], const uint32_t b[2], const float c[4]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};\n"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}
}  // namespace dec
}  // namespace ggq
#else
#include "../../tests/host/cuda_shim.h"
#define GGQ_DEV inline
#endif

namespace ggq {
namespace dec {

// ---- per-format staging geometry -------------------------------------------------------------
// CHUNK_BLOCKS blocks of one row form the unit a pipeline stage holds for each of the 16 rows; a stage is
// CHUNK_BLOCKS / PREP_BLOCKS sub-tiles of [16 rows x SLOT bytes], each filled by one 2-D TMA box.
// SLOT is the shared-memory pitch between the 16 rows of a stage; it is >= the largest copy
// (chunk bytes + 16-byte alignment slack) and chosen so the per-lane loads below are bank-conflict free.
template <int FMT> struct Geo;
template <> struct Geo<0> {  // Q8_0: 16 blocks = 512 weights = 544 B in one TMA box of 560 B (16 B of the next chunk ride along so
                             // that slot/4 % 32 == 12 and the lanes' 32-bit loads are conflict free; measured faster than
                             // two 272 B boxes: TMA cost is per box row)
#ifndef GGQ_Q8_0_CHUNK
#define GGQ_Q8_0_CHUNK 16
#endif
    // (GGQ_Q8_0_CHUNK = 8: 272 B chunks in 304 B slots, the same bank pattern — slot/4 % 32 == 12)
    static constexpr int QK = 32, BLK = 34, CHUNK_BLOCKS = GGQ_Q8_0_CHUNK, CHUNK_ELEMS = 32 * CHUNK_BLOCKS,
                         CHUNK_BYTES = 34 * CHUNK_BLOCKS, SLOT = CHUNK_BLOCKS == 16 ? 560 : 304;
    static constexpr int PREP_BLOCKS = CHUNK_BLOCKS;      // blocks handled per prep/compute sub-step (= one TMA box) of a stage
    static constexpr int GROUP = 32;            // activations per pre-summed group (one block)
    static constexpr int SCRATCH_PER_BLOCK = 0;  // bytes of prepared scales per (row, block)
    static constexpr float TBL_MUL = -128.f / 16777216.f;   // cancels the +128 of (q ^ 0x80)
};
template <> struct Geo<1> {  // Q4_K: 2 blocks = 512 weights = 288 B = one TMA box per stage (288 B pitch: rows land 8 banks apart, so
                             // the lanes' 64-bit loads are conflict free).  Small stages let 12-16 warps per SM keep >= 2 stages
                             // each; 4-block stages with 8 warps measured slower once the per-chunk control code was trimmed.
#ifndef GGQ_Q4K_CHUNK
#define GGQ_Q4K_CHUNK 2
#endif
    static constexpr int QK = 256, BLK = 144, CHUNK_BLOCKS = GGQ_Q4K_CHUNK, CHUNK_ELEMS = 256 * CHUNK_BLOCKS,
                         CHUNK_BYTES = 144 * CHUNK_BLOCKS, SLOT = 288;
    static constexpr int PREP_BLOCKS = 2;
    static constexpr int GROUP = 32;
    static constexpr int SCRATCH_PER_BLOCK = 80;  // 64 B payload + 16 B pad: rows 20 banks apart
    static constexpr float TBL_MUL = 1.f;
};
template <> struct Geo<2> {  // Q6_K: 2 blocks = 512 weights = 420 B; the 16-byte aligned superset is always 432 B
    static constexpr int QK = 256, BLK = 210, CHUNK_BLOCKS = 2, CHUNK_ELEMS = 512, CHUNK_BYTES = 420, SLOT = 432;
    static constexpr int PREP_BLOCKS = 2;
    static constexpr int GROUP = 16;
    static constexpr int SCRATCH_PER_BLOCK = 80;  // 64 B payload + 16 B pad: rows 20 banks apart
    static constexpr float TBL_MUL = -32.f / 16777216.f;    // the -32 of (q6 - 32)
};

struct Lane {
    int lane, g, t;  // g = lane >> 2 (row / token group), t = lane & 3 (k group)
};

// What one warp needs to consume one stage.
struct StageArgs {
    const uint8_t* rows;      // stage base: 16 row slots of Geo::SLOT bytes
    int data_off;             // byte offset of this sub-step's first block inside every slot
    int nblk;                 // blocks in this sub-step (<= PREP_BLOCKS; even for Q8_0 / Q6_K)
    const uint8_t* xrow[2];   // per n-tile: this lane's activation row (token), at slice-relative k = 0
    bool xv[2];               // per n-tile: does this lane's token exist?  Loads of missing tokens are predicated
                              // off (their MMA columns are never stored), which also keeps the shared-memory
                              // wavefront count proportional to T instead of 8
    int k0;                   // slice-relative element index of the sub-step's first weight
    const float* tbl;         // [k / GROUP][8 * NT] pre-summed activations times Geo::TBL_MUL
    uint8_t* scratch;         // this warp's prepared-scale area (SCRATCH_PER_BLOCK * 16 * PREP_BLOCKS)
};

GGQ_DEV uint32_t ld32(const uint8_t* p) { return *reinterpret_cast<const uint32_t*>(p); }
GGQ_DEV uint32_t ld32p(const uint8_t* p, bool v) { return v ? *reinterpret_cast<const uint32_t*>(p) : 0u; }
GGQ_DEV uint2 ld64p(const uint8_t* p, bool v) { return v ? *reinterpret_cast<const uint2*>(p) : uint2{0u, 0u}; }
GGQ_DEV uint4 ld128p(const uint8_t* p, bool v) { return v ? *reinterpret_cast<const uint4*>(p) : uint4{0u, 0u, 0u, 0u}; }
// Predicated shared-memory load that KEEPS the destination of inactive lanes (no zeroing): the lanes whose token does
// not exist feed MMA columns that are never stored, so whatever they hold is fine.
GGQ_DEV void ld128k(uint4& d, const uint8_t* p, bool v) {
#ifdef __CUDACC__
    asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %5, 0;\n @q ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n}\n"
                 : "+r"(d.x), "+r"(d.y), "+r"(d.z), "+r"(d.w)
                 : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))), "r"(static_cast<int>(v)));
#else
    if (v) d = *reinterpret_cast<const uint4*>(p);
#endif
}
GGQ_DEV void ld64k(uint2& d, const uint8_t* p, bool v) {
#ifdef __CUDACC__
    asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %3, 0;\n @q ld.shared.v2.b32 {%0, %1}, [%2];\n}\n"
                 : "+r"(d.x), "+r"(d.y)
                 : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))), "r"(static_cast<int>(v)));
#else
    if (v) d = *reinterpret_cast<const uint2*>(p);
#endif
}
GGQ_DEV void ld32k(uint32_t& d, const uint8_t* p, bool v) {
#ifdef __CUDACC__
    asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n @q ld.shared.b32 %0, [%1];\n}\n"
                 : "+r"(d)
                 : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))), "r"(static_cast<int>(v)));
#else
    if (v) d = *reinterpret_cast<const uint32_t*>(p);
#endif
}
GGQ_DEV uint2 ld64(const uint8_t* p) { return *reinterpret_cast<const uint2*>(p); }
GGQ_DEV uint4 ld128(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }
GGQ_DEV float4 ld128f(const uint8_t* p) { return *reinterpret_cast<const float4*>(p); }
GGQ_DEV float2 ld64f(const float* p) { return *reinterpret_cast<const float2*>(p); }
GGQ_DEV uint32_t ld16(const uint8_t* p) { return *reinterpret_cast<const uint16_t*>(p); }
GGQ_DEV float ldf(const uint8_t* p) { return *reinterpret_cast<const float*>(p); }

// acc[0,1] belong to row g, acc[2,3] to row g+8 (mma C fragment order)
template <int NT> struct Acc {
    float v[NT][4];
};

// =============================================================================================
// Q8_0
// =============================================================================================
// Two blocks = 68 B = 17 words:  [d0 q0 q1][q2..q5]...[q26..q29][q30 q31 d1][Q0..Q3]...[Q28..Q31]
// even block: lane t takes word 1+t (q[2+4t..5+4t]) and word 5+t (q[18+4t..21+4t]); for t == 3 the
// second word is rebuilt as [q30 q31 q0 q1] from words 8 and 0.  odd block: words 9+t and 13+t.
GGQ_DEV void q8_to_h2(uint32_t w, uint32_t& lo, uint32_t& hi) {
    const uint32_t u = w ^ 0x80808080u;       // q + 128 in every byte
    lo = prmt(u, 0u, 0x5140);                 // halves (0x00|u0, 0x00|u1) = (q + 128) * 2^-24
    hi = prmt(u, 0u, 0x7362);
}

template <int NT, bool FULL>
GGQ_DEV void compute_q8_0(const Lane& L, const StageArgs& s, Acc<NT>& acc) {
    using G = Geo<0>;
    const uint8_t* r0 = s.rows + L.g * G::SLOT + s.data_off;
    const uint8_t* r1 = r0 + 8 * G::SLOT;
    const uint32_t wrap_sel = (L.t == 3) ? 0x7610u : 0x3210u;
    const int t4 = 4 * L.t;
    // activation fragments: lanes whose token does not exist keep whatever these hold (their columns are never stored)
    uint32_t xk[NT][2][2];
    uint2 xo[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            xk[nt][j][0] = xk[nt][j][1] = 0u;
            xo[nt][j] = uint2{0u, 0u};
        }
#pragma unroll
    for (int p = 0; p < G::PREP_BLOCKS / 2; ++p) {
        if (!FULL && 2 * p >= s.nblk) break;
        const uint8_t* a = r0 + 68 * p;
        const uint8_t* b = r1 + 68 * p;
        const uint32_t a0w = ld32(a), a8w = ld32(a + 32), b0w = ld32(b), b8w = ld32(b + 32);
        uint32_t ae1 = ld32(a + 4 + t4), ae2 = ld32(a + 20 + t4), ao1 = ld32(a + 36 + t4), ao2 = ld32(a + 52 + t4);
        uint32_t be1 = ld32(b + 4 + t4), be2 = ld32(b + 20 + t4), bo1 = ld32(b + 36 + t4), bo2 = ld32(b + 52 + t4);
        ae2 = prmt(ae2, a0w, wrap_sel);
        be2 = prmt(be2, b0w, wrap_sel);
        const float da_e = h2f(a0w & 0xffffu) * 16777216.f, da_o = h2f(a8w >> 16) * 16777216.f;
        const float db_e = h2f(b0w & 0xffffu) * 16777216.f, db_o = h2f(b8w >> 16) * 16777216.f;
        const int kb = s.k0 + 64 * p;  // first weight of the even block (slice relative)
        uint32_t fa[4], fb[4];         // A fragments of the two k16 steps of a block
        // ---- even block: activations start at k = 2 (mod 4) -> two 32-bit loads per fragment
        q8_to_h2(ae1, fa[0], fa[2]);
        q8_to_h2(be1, fa[1], fa[3]);
        q8_to_h2(ae2, fb[0], fb[2]);
        q8_to_h2(be2, fb[1], fb[3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const uint8_t* x = s.xrow[nt] + 2 * kb;
            const float2 c = ld64f(s.tbl + (kb >> 5) * (8 * NT) + 8 * nt + 2 * L.t);
            float d[4] = {c.x, c.y, c.x, c.y};
            uint32_t* bf = xk[nt][0];
            uint32_t* bg = xk[nt][1];
            ld32k(bf[0], x + 2 * (2 + t4), s.xv[nt]);
            ld32k(bf[1], x + 2 * (4 + t4), s.xv[nt]);
            mma16816(d, fa, bf, d);
            ld32k(bg[0], x + 2 * (18 + t4), s.xv[nt]);
            ld32k(bg[1], x + 2 * ((20 + t4) & 31), s.xv[nt]);
            mma16816(d, fb, bg, d);
            acc.v[nt][0] = fmaf(da_e, d[0], acc.v[nt][0]);
            acc.v[nt][1] = fmaf(da_e, d[1], acc.v[nt][1]);
            acc.v[nt][2] = fmaf(db_e, d[2], acc.v[nt][2]);
            acc.v[nt][3] = fmaf(db_e, d[3], acc.v[nt][3]);
        }
        // ---- odd block: words are aligned with k = 0 (mod 4) -> one 64-bit load per fragment
        q8_to_h2(ao1, fa[0], fa[2]);
        q8_to_h2(bo1, fa[1], fa[3]);
        q8_to_h2(ao2, fb[0], fb[2]);
        q8_to_h2(bo2, fb[1], fb[3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const uint8_t* x = s.xrow[nt] + 2 * (kb + 32);
            const float2 c = ld64f(s.tbl + ((kb >> 5) + 1) * (8 * NT) + 8 * nt + 2 * L.t);
            float d[4] = {c.x, c.y, c.x, c.y};
            uint2& v = xo[nt][0];
            uint2& u = xo[nt][1];
            ld64k(v, x + 2 * t4, s.xv[nt]);
            const uint32_t bf[2] = {v.x, v.y};
            mma16816(d, fa, bf, d);
            ld64k(u, x + 2 * (16 + t4), s.xv[nt]);
            const uint32_t bg[2] = {u.x, u.y};
            mma16816(d, fb, bg, d);
            acc.v[nt][0] = fmaf(da_o, d[0], acc.v[nt][0]);
            acc.v[nt][1] = fmaf(da_o, d[1], acc.v[nt][1]);
            acc.v[nt][2] = fmaf(db_o, d[2], acc.v[nt][2]);
            acc.v[nt][3] = fmaf(db_o, d[3], acc.v[nt][3]);
        }
    }
}


// ---- Q8_0, one token (GEMV): the 8 MMA columns carry 8 consecutive blocks (see the Q4_K GEMV notes below) ----
// Lane (g, t) loads the activations of block g of the 8-block group in the pattern of that block's parity (even
// blocks: k = 2 + 4t.., odd blocks: k = 4t..) once per group; column j of block j's MMAs is the dot product, held by
// lane t == j / 2, which also is the only lane that converts that block's scale.
template <bool FULL>
GGQ_DEV void compute_q8_0_gv(const Lane& L, const StageArgs& s, Acc<1>& acc) {
    using G = Geo<0>;
    const uint8_t* r0 = s.rows + L.g * G::SLOT + s.data_off;
    const uint8_t* r1 = r0 + 8 * G::SLOT;
    const uint32_t wrap_sel = (L.t == 3) ? 0x7610u : 0x3210u;
    const int t4 = 4 * L.t;
    const bool even = !(L.g & 1);
    const int o1 = even ? 2 + t4 : t4, o3 = o1 + 16, o4 = (o3 + 2) & 31;
    float a0 = acc.v[0][0], a1 = acc.v[0][1], a2 = acc.v[0][2], a3 = acc.v[0][3];
#pragma unroll
    for (int u = 0; u < G::PREP_BLOCKS / 8; ++u) {
        if (!FULL && 8 * u >= s.nblk) break;
        const int kb8 = s.k0 + 256 * u;
        const uint8_t* x = s.xrow[0] + 2 * (kb8 + 32 * L.g);
        const uint32_t b1[2] = {ld32(x + 2 * o1), ld32(x + 2 * (o1 + 2))};
        const uint32_t b2[2] = {ld32(x + 2 * o3), ld32(x + 2 * o4)};
        const float2 tb = ld64f(s.tbl + (kb8 >> 5) + 2 * L.t);
        const float c[4] = {tb.x, tb.y, tb.x, tb.y};
        // this lane's scales: the pair of blocks (2t, 2t + 1) of the group, rows g and g + 8
        const uint8_t* da = r0 + 68 * (4 * u + L.t);
        const uint8_t* db = r1 + 68 * (4 * u + L.t);
        const float sa_e = h2f(ld16(da)) * 16777216.f, sa_o = h2f(ld16(da + 34)) * 16777216.f;
        const float sb_e = h2f(ld16(db)) * 16777216.f, sb_o = h2f(ld16(db + 34)) * 16777216.f;
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
            const uint8_t* a = r0 + 68 * (4 * u + pp);
            const uint8_t* b = r1 + 68 * (4 * u + pp);
            const uint32_t a0w = ld32(a), b0w = ld32(b);
            uint32_t ae1 = ld32(a + 4 + t4), ae2 = ld32(a + 20 + t4), ao1 = ld32(a + 36 + t4), ao2 = ld32(a + 52 + t4);
            uint32_t be1 = ld32(b + 4 + t4), be2 = ld32(b + 20 + t4), bo1 = ld32(b + 36 + t4), bo2 = ld32(b + 52 + t4);
            ae2 = prmt(ae2, a0w, wrap_sel);
            be2 = prmt(be2, b0w, wrap_sel);
            uint32_t fa[4], fb[4];
            float d[4];
            q8_to_h2(ae1, fa[0], fa[2]);
            q8_to_h2(be1, fa[1], fa[3]);
            q8_to_h2(ae2, fb[0], fb[2]);
            q8_to_h2(be2, fb[1], fb[3]);
            mma16816(d, fa, b1, c);
            mma16816(d, fb, b2, d);
            if (L.t == pp) {
                a0 = fmaf(sa_e, d[0], a0);
                a2 = fmaf(sb_e, d[2], a2);
            }
            q8_to_h2(ao1, fa[0], fa[2]);
            q8_to_h2(bo1, fa[1], fa[3]);
            q8_to_h2(ao2, fb[0], fb[2]);
            q8_to_h2(bo2, fb[1], fb[3]);
            mma16816(d, fa, b1, c);
            mma16816(d, fb, b2, d);
            if (L.t == pp) {
                a1 = fmaf(sa_o, d[1], a1);
                a3 = fmaf(sb_o, d[3], a3);
            }
        }
    }
    acc.v[0][0] = a0;
    acc.v[0][1] = a1;
    acc.v[0][2] = a2;
    acc.v[0][3] = a3;
}

// =============================================================================================
// Q4_K
// =============================================================================================
// Q4_K specifics of this family:
//  * activations are stored PERMUTED inside every group of four (x0 x2 x1 x3), so the nibble masks
//    (w & 0x000F000F -> weights l, l+2;  (w >> 8) & 0x000F000F -> l+1, l+3) need no byte shuffle;
//  * the -dmin*m_j*sum_j(x) term of all 8 sub-blocks of a block is ONE extra bf16 MMA per block:
//    A = the 6-bit mins m_j (exact in bf16), B = the 32-activation sums split into bf16 hi + lo
//    (k slots 0..7 = hi_j, 8..15 = lo_j; ~2^-17 relative), then acc -= dmin * D;
//  * prep decodes each 16-byte block header once per (row, block) into an 80-byte scratch entry:
//      float s[8]  = d*sc_j * 2^24 (even j: low nibbles = q*2^-24) / * 2^20 (odd j: high nibbles = q*2^-20)
//      u32  mp[4]  = bf16x2(m_2t, m_2t+1)        float dmin
GGQ_DEV uint32_t bf16_bits_rn(float f) {  // float -> bf16 bit pattern, round to nearest even (finite inputs)
    uint32_t u = f2u(f);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return u >> 16;
}

template <bool FULL>
GGQ_DEV void prep_q4_k(const Lane& L, const StageArgs& s) {
    using G = Geo<1>;
    for (int p = L.lane; p < 16 * (FULL ? G::PREP_BLOCKS : s.nblk); p += 32) {
        const int row = p & 15, blk = p >> 4;
        const uint4 h = ld128(s.rows + row * G::SLOT + s.data_off + blk * G::BLK);
        const float d = h2f(h.x & 0xffffu), dmin = h2f(h.x >> 16);
        const uint32_t u0 = h.y, u1 = h.z, u2 = h.w;
        const uint32_t sc_lo = u0 & 0x3f3f3f3fu, m_lo = u1 & 0x3f3f3f3fu;  // sub-blocks 0..3 (q4_k_ref.c:176-178)
        const uint32_t sc_hi = (u2 & 0x0f0f0f0fu) | ((u0 >> 2) & 0x30303030u);         // 4..7 (:180-183)
        const uint32_t m_hi = ((u2 >> 4) & 0x0f0f0f0fu) | ((u1 >> 2) & 0x30303030u);
        // byte -> fp32 without the integer pipe: PRMT builds the bit pattern of 2^23 + n, one FMA removes the 2^23
        // and applies the scale: (2^23 + n) * c - 2^23 * c == n * c exactly (n * c has <= 17 significant bits)
        const float d24 = d * 16777216.f, d20 = d * 1048576.f;
        const float n24 = d24 * -8388608.f, n20 = d20 * -8388608.f;
        uint8_t* e = s.scratch + (blk * 16 + row) * G::SCRATCH_PER_BLOCK;
        float4 a, b;
        a.x = fmaf(u2f(prmt(sc_lo, 0x4B000000u, 0x7650)), d24, n24);
        a.y = fmaf(u2f(prmt(sc_lo, 0x4B000000u, 0x7651)), d20, n20);
        a.z = fmaf(u2f(prmt(sc_lo, 0x4B000000u, 0x7652)), d24, n24);
        a.w = fmaf(u2f(prmt(sc_lo, 0x4B000000u, 0x7653)), d20, n20);
        b.x = fmaf(u2f(prmt(sc_hi, 0x4B000000u, 0x7650)), d24, n24);
        b.y = fmaf(u2f(prmt(sc_hi, 0x4B000000u, 0x7651)), d20, n20);
        b.z = fmaf(u2f(prmt(sc_hi, 0x4B000000u, 0x7652)), d24, n24);
        b.w = fmaf(u2f(prmt(sc_hi, 0x4B000000u, 0x7653)), d20, n20);
        // mins as bf16 pairs: float(n) = (2^23 + n) - 2^23; integers < 64 are exact in bf16 = the upper halves
        const float f0 = u2f(prmt(m_lo, 0x4B000000u, 0x7650)) - 8388608.f, f1 = u2f(prmt(m_lo, 0x4B000000u, 0x7651)) - 8388608.f;
        const float f2 = u2f(prmt(m_lo, 0x4B000000u, 0x7652)) - 8388608.f, f3 = u2f(prmt(m_lo, 0x4B000000u, 0x7653)) - 8388608.f;
        const float f4 = u2f(prmt(m_hi, 0x4B000000u, 0x7650)) - 8388608.f, f5 = u2f(prmt(m_hi, 0x4B000000u, 0x7651)) - 8388608.f;
        const float f6 = u2f(prmt(m_hi, 0x4B000000u, 0x7652)) - 8388608.f, f7 = u2f(prmt(m_hi, 0x4B000000u, 0x7653)) - 8388608.f;
        uint4 m;
        m.x = prmt(f2u(f0), f2u(f1), 0x7632);
        m.y = prmt(f2u(f2), f2u(f3), 0x7632);
        m.z = prmt(f2u(f4), f2u(f5), 0x7632);
        m.w = prmt(f2u(f6), f2u(f7), 0x7632);
        *reinterpret_cast<float4*>(e) = a;
        *reinterpret_cast<float4*>(e + 16) = b;
        *reinterpret_cast<uint4*>(e + 32) = m;
        *reinterpret_cast<float*>(e + 48) = dmin;
    }
}

// Per (block, token) the staging pass turns the raw activations into what compute_q4_k reads:
// the permuted fp16 row (in place) and the bf16 hi/lo split of the eight 32-activation sums, laid out
// as the B fragment of the min-term MMA: xb[(blk * TPAD + token) * 4 + t] = {bf16x2(hi_2t, hi_2t+1),
// bf16x2(lo_2t, lo_2t+1)}.  One call handles one entry = 64 activations (sub-blocks 2t, 2t+1).
GGQ_DEV void q4k_stage_pair(uint8_t* x, uint2* xb_entry, bool token_valid) {
    // one B-fragment entry: sub-blocks 2t and 2t+1 (64 activations at `x`)
    if (!token_valid) {
        *xb_entry = uint2{0u, 0u};
        return;
    }
    uint32_t hi[2], lo[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        float sum = 0.f;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            uint4* ptr = reinterpret_cast<uint4*>(x + 64 * j + 16 * v);
            const uint4 q = *ptr;
            sum += h2f(q.x & 0xffffu);
            sum += h2f(q.x >> 16);
            sum += h2f(q.y & 0xffffu);
            sum += h2f(q.y >> 16);
            sum += h2f(q.z & 0xffffu);
            sum += h2f(q.z >> 16);
            sum += h2f(q.w & 0xffffu);
            sum += h2f(q.w >> 16);
            uint4 o;  // (x0 x1 x2 x3) -> (x0 x2 x1 x3), per group of four
            o.x = prmt(q.x, q.y, 0x5410);
            o.y = prmt(q.x, q.y, 0x7632);
            o.z = prmt(q.z, q.w, 0x5410);
            o.w = prmt(q.z, q.w, 0x7632);
            *ptr = o;
        }
        const uint32_t h = f2u(sum) & 0xffff0000u;          // bf16 hi part (truncated), exact remainder below
        hi[j] = h >> 16;
        lo[j] = bf16_bits_rn(sum - u2f(h));
    }
    *xb_entry = uint2{hi[0] | (hi[1] << 16), lo[0] | (lo[1] << 16)};
}

// FULL = the sub-step holds all PREP_BLOCKS blocks: no per-block bounds checks, so the whole sub-step is one basic
// block and ptxas can hoist the next chunk's shared-memory loads over the current chunk's math.
template <int NT, bool FULL>
GGQ_DEV void compute_q4_k_impl(const Lane& L, const StageArgs& s, Acc<NT>& acc) {
    using G = Geo<1>;
    const uint8_t* r0 = s.rows + L.g * G::SLOT + s.data_off;
    const uint8_t* r1 = r0 + 8 * G::SLOT;
    const float zero[4] = {0.f, 0.f, 0.f, 0.f};
    const uint2* xbt = reinterpret_cast<const uint2*>(s.tbl);
    uint4 xek[2][NT], xok[2][NT];
    uint2 xbk[NT];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            xek[j][nt] = xok[j][nt] = uint4{0u, 0u, 0u, 0u};
            xbk[nt] = uint2{0u, 0u};
        }
#pragma unroll
    for (int i = 0; i < G::PREP_BLOCKS; ++i) {
        if (!FULL && i >= s.nblk) break;
        const uint8_t* q0 = r0 + i * G::BLK + 16 + 8 * L.t;
        const uint8_t* q1 = r1 + i * G::BLK + 16 + 8 * L.t;
        const uint8_t* e0 = s.scratch + (i * 16 + L.g) * G::SCRATCH_PER_BLOCK;
        const uint8_t* e1 = e0 + 8 * G::SCRATCH_PER_BLOCK;
        const float4 sa0 = ld128f(e0), sa1 = ld128f(e0 + 16), sb0 = ld128f(e1), sb1 = ld128f(e1 + 16);
        const float sa[8] = {sa0.x, sa0.y, sa0.z, sa0.w, sa1.x, sa1.y, sa1.z, sa1.w};
        const float sb[8] = {sb0.x, sb0.y, sb0.z, sb0.w, sb1.x, sb1.y, sb1.z, sb1.w};
        {   // ---- min term of the whole block: one bf16 MMA per n-tile
            const uint32_t mpa = ld32(e0 + 32 + 4 * L.t), mpb = ld32(e1 + 32 + 4 * L.t);
            const float dmin_a = ldf(e0 + 48), dmin_b = ldf(e1 + 48);
            const uint32_t ma[4] = {mpa, mpb, mpa, mpb};
            const int blk = (s.k0 >> 8) + i;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                uint2& xb = xbk[nt];
                ld64k(xb, reinterpret_cast<const uint8_t*>(xbt + (blk * (8 * NT) + 8 * nt + L.g) * 4 + L.t), s.xv[nt]);
                const uint32_t bfr[2] = {xb.x, xb.y};
                float dm[4];
                mma16816_bf16(dm, ma, bfr, zero);
                acc.v[nt][0] = fmaf(-dmin_a, dm[0], acc.v[nt][0]);
                acc.v[nt][1] = fmaf(-dmin_a, dm[1], acc.v[nt][1]);
                acc.v[nt][2] = fmaf(-dmin_b, dm[2], acc.v[nt][2]);
                acc.v[nt][3] = fmaf(-dmin_b, dm[3], acc.v[nt][3]);
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint2 wa = ld64(q0 + 32 * c), wb = ld64(q1 + 32 * c);
            uint32_t e1f[4], e2f[4], o1f[4], o2f[4];  // A fragments: even/odd sub-block, first/second k16 step
            e1f[0] = wa.x & 0x000F000Fu;  e1f[2] = (wa.x >> 8) & 0x000F000Fu;   // q * 2^-24 (fp16 subnormals)
            o1f[0] = wa.x & 0x00F000F0u;  o1f[2] = (wa.x >> 8) & 0x00F000F0u;   // q * 2^-20
            e1f[1] = wb.x & 0x000F000Fu;  e1f[3] = (wb.x >> 8) & 0x000F000Fu;
            o1f[1] = wb.x & 0x00F000F0u;  o1f[3] = (wb.x >> 8) & 0x00F000F0u;
            e2f[0] = wa.y & 0x000F000Fu;  e2f[2] = (wa.y >> 8) & 0x000F000Fu;
            o2f[0] = wa.y & 0x00F000F0u;  o2f[2] = (wa.y >> 8) & 0x00F000F0u;
            e2f[1] = wb.y & 0x000F000Fu;  e2f[3] = (wb.y >> 8) & 0x000F000Fu;
            o2f[1] = wb.y & 0x00F000F0u;  o2f[3] = (wb.y >> 8) & 0x00F000F0u;
            const int kb = s.k0 + 256 * i + 64 * c;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                uint4& xe = xek[c & 1][nt];
                uint4& xo = xok[c & 1][nt];
                ld128k(xe, s.xrow[nt] + 2 * (kb + 8 * L.t), s.xv[nt]);   // permuted: (x0 x2 | x1 x3 | x4 x6 | x5 x7)
                ld128k(xo, s.xrow[nt] + 2 * (kb + 32 + 8 * L.t), s.xv[nt]);
                float de[4], dd[4];
                uint32_t bf[2] = {xe.x, xe.y};
                mma16816(de, e1f, bf, zero);
                bf[0] = xe.z;
                bf[1] = xe.w;
                mma16816(de, e2f, bf, de);
                bf[0] = xo.x;
                bf[1] = xo.y;
                mma16816(dd, o1f, bf, zero);
                bf[0] = xo.z;
                bf[1] = xo.w;
                mma16816(dd, o2f, bf, dd);
                float a0 = acc.v[nt][0], a1 = acc.v[nt][1], a2 = acc.v[nt][2], a3 = acc.v[nt][3];
                a0 = fmaf(sa[2 * c], de[0], a0);
                a1 = fmaf(sa[2 * c], de[1], a1);
                a2 = fmaf(sb[2 * c], de[2], a2);
                a3 = fmaf(sb[2 * c], de[3], a3);
                a0 = fmaf(sa[2 * c + 1], dd[0], a0);
                a1 = fmaf(sa[2 * c + 1], dd[1], a1);
                a2 = fmaf(sb[2 * c + 1], dd[2], a2);
                a3 = fmaf(sb[2 * c + 1], dd[3], a3);
                acc.v[nt][0] = a0;
                acc.v[nt][1] = a1;
                acc.v[nt][2] = a2;
                acc.v[nt][3] = a3;
            }
        }
    }
}



// ---- Q4_K, one token (GEMV) ----------------------------------------------------------------------
// With a single token only column 0 of the m16n8 MMA would carry data.  Instead the 8 columns carry the 8 sub-blocks
// of a block: lane (g, t) loads the activations of sub-block g (its t-th quarter) ONCE per block and uses them as the
// B fragment of all 16 MMAs of the block, so D_j[row, n] = <weights of sub-block j, activations of sub-block n>.
// Only the column n == j is meaningful; it lives in lane t == j / 2, which applies the sub-block scale with a
// predicated FMA.  Per block and lane: 1 activation load (instead of 8), 2 scale loads (instead of 8), 20 FMAs
// (instead of 36) and no min-term MMA: -dmin * m_j * sum_j(x) is one more FMA per (row, sub-block).  The four
// accumulators of a lane are partial sums over (row g / g + 8) x (even / odd sub-blocks); gemv_finalize() folds
// them over the quad.
// prep: scratch entry = 64 B per (row, block): four 16-byte pieces {s_2t, s_2t+1, -dmin*m_2t, -dmin*m_2t+1},
// piece t stored at position t ^ ((row >> 1) & 3) so that both the 128-bit stores of prep (lane = row) and the
// 128-bit loads of compute (lane = (g, t)) are bank-conflict free.
template <bool FULL>
GGQ_DEV void prep_q4_k_gv(const Lane& L, const StageArgs& s) {
    using G = Geo<1>;
    for (int p = L.lane; p < 16 * (FULL ? G::PREP_BLOCKS : s.nblk); p += 32) {
        const int row = p & 15, blk = p >> 4;
        const uint4 h = ld128(s.rows + row * G::SLOT + s.data_off + blk * G::BLK);
        const float d = h2f(h.x & 0xffffu), dmin = h2f(h.x >> 16);
        const uint32_t u0 = h.y, u1 = h.z, u2 = h.w;
        const uint32_t sc_lo = u0 & 0x3f3f3f3fu, m_lo = u1 & 0x3f3f3f3fu;
        const uint32_t sc_hi = (u2 & 0x0f0f0f0fu) | ((u0 >> 2) & 0x30303030u);
        const uint32_t m_hi = ((u2 >> 4) & 0x0f0f0f0fu) | ((u1 >> 2) & 0x30303030u);
        // (2^23 + n) * c - 2^23 * c == n * c exactly (one rounding, n * c has <= 17 significant bits)
        const float d24 = d * 16777216.f, d20 = d * 1048576.f;
        const float n24 = d24 * -8388608.f, n20 = d20 * -8388608.f;
        const float dm = -dmin, ndm = dmin * 8388608.f;
        float4 c0, c1, c2, c3;
        c0.x = fmaf(u2f(prmt(sc_lo, 0x4B000000u, 0x7650)), d24, n24);
        c0.y = fmaf(u2f(prmt(sc_lo, 0x4B000000u, 0x7651)), d20, n20);
        c1.x = fmaf(u2f(prmt(sc_lo, 0x4B000000u, 0x7652)), d24, n24);
        c1.y = fmaf(u2f(prmt(sc_lo, 0x4B000000u, 0x7653)), d20, n20);
        c2.x = fmaf(u2f(prmt(sc_hi, 0x4B000000u, 0x7650)), d24, n24);
        c2.y = fmaf(u2f(prmt(sc_hi, 0x4B000000u, 0x7651)), d20, n20);
        c3.x = fmaf(u2f(prmt(sc_hi, 0x4B000000u, 0x7652)), d24, n24);
        c3.y = fmaf(u2f(prmt(sc_hi, 0x4B000000u, 0x7653)), d20, n20);
        c0.z = fmaf(u2f(prmt(m_lo, 0x4B000000u, 0x7650)), dm, ndm);
        c0.w = fmaf(u2f(prmt(m_lo, 0x4B000000u, 0x7651)), dm, ndm);
        c1.z = fmaf(u2f(prmt(m_lo, 0x4B000000u, 0x7652)), dm, ndm);
        c1.w = fmaf(u2f(prmt(m_lo, 0x4B000000u, 0x7653)), dm, ndm);
        c2.z = fmaf(u2f(prmt(m_hi, 0x4B000000u, 0x7650)), dm, ndm);
        c2.w = fmaf(u2f(prmt(m_hi, 0x4B000000u, 0x7651)), dm, ndm);
        c3.z = fmaf(u2f(prmt(m_hi, 0x4B000000u, 0x7652)), dm, ndm);
        c3.w = fmaf(u2f(prmt(m_hi, 0x4B000000u, 0x7653)), dm, ndm);
        uint8_t* e = s.scratch + (blk * 16 + row) * 64;
        const int sw = ((row >> 1) & 3) << 4;
        *reinterpret_cast<float4*>(e + sw) = c0;
        *reinterpret_cast<float4*>(e + (sw ^ 16)) = c1;
        *reinterpret_cast<float4*>(e + (sw ^ 32)) = c2;
        *reinterpret_cast<float4*>(e + (sw ^ 48)) = c3;
    }
}

// staging for the GEMV kernel: permute 64 activations in place (as q4k_stage_pair) and write the two fp32 sums
GGQ_DEV void q4k_stage_pair_gv(uint8_t* x, float* sums2) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        float sum = 0.f;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            uint4* ptr = reinterpret_cast<uint4*>(x + 64 * j + 16 * v);
            const uint4 q = *ptr;
            sum += h2f(q.x & 0xffffu);
            sum += h2f(q.x >> 16);
            sum += h2f(q.y & 0xffffu);
            sum += h2f(q.y >> 16);
            sum += h2f(q.z & 0xffffu);
            sum += h2f(q.z >> 16);
            sum += h2f(q.w & 0xffffu);
            sum += h2f(q.w >> 16);
            uint4 o;
            o.x = prmt(q.x, q.y, 0x5410);
            o.y = prmt(q.x, q.y, 0x7632);
            o.z = prmt(q.z, q.w, 0x5410);
            o.w = prmt(q.z, q.w, 0x7632);
            *ptr = o;
        }
        sums2[j] = sum;
    }
}

template <bool FULL>
GGQ_DEV void compute_q4_k_gv_impl(const Lane& L, const StageArgs& s, Acc<1>& acc) {
    using G = Geo<1>;
    const uint8_t* r0 = s.rows + L.g * G::SLOT + s.data_off;
    const uint8_t* r1 = r0 + 8 * G::SLOT;
    const float zero[4] = {0.f, 0.f, 0.f, 0.f};
    const int piece = (L.t ^ ((L.g >> 1) & 3)) << 4;
    float a0 = acc.v[0][0], a1 = acc.v[0][1], a2 = acc.v[0][2], a3 = acc.v[0][3];
#pragma unroll
    for (int i = 0; i < G::PREP_BLOCKS; ++i) {
        if (!FULL && i >= s.nblk) break;
        const uint8_t* q0 = r0 + i * G::BLK + 16 + 8 * L.t;
        const uint8_t* q1 = r1 + i * G::BLK + 16 + 8 * L.t;
        const uint8_t* e0 = s.scratch + (i * 16 + L.g) * 64 + piece;
        const float4 ca = ld128f(e0), cb = ld128f(e0 + 8 * 64);
        const int kb = s.k0 + 256 * i;
        const float2 sm = ld64f(s.tbl + (kb >> 5) + 2 * L.t);                 // sums of sub-blocks 2t, 2t+1
        const uint4 xr = ld128(s.xrow[0] + 2 * (kb + 32 * L.g + 8 * L.t));  // sub-block g, quarter t (permuted)
        const uint32_t b1[2] = {xr.x, xr.y}, b2[2] = {xr.z, xr.w};
        a0 = fmaf(ca.z, sm.x, a0);
        a1 = fmaf(ca.w, sm.y, a1);
        a2 = fmaf(cb.z, sm.x, a2);
        a3 = fmaf(cb.w, sm.y, a3);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint2 wa = ld64(q0 + 32 * c), wb = ld64(q1 + 32 * c);
            uint32_t e1f[4], e2f[4], o1f[4], o2f[4];
            float de[4], dd[4];
#ifndef GGQ_Q4K_NO_BYTEFRAG
            // odd sub-block through whole BYTES (16 * hi + lo) * 2^-24: one PRMT replaces a shift and a mask per word; the
            // low-nibble cross term it drags in is exactly de[odd column], which this lane already holds
            const uint32_t ta = prmt(wa.x, 0u, 0x4341), tb = prmt(wb.x, 0u, 0x4341);
            const uint32_t ua = prmt(wa.y, 0u, 0x4341), ub = prmt(wb.y, 0u, 0x4341);
            e1f[0] = wa.x & 0x000F000Fu;  e1f[2] = ta & 0x000F000Fu;
            e1f[1] = wb.x & 0x000F000Fu;  e1f[3] = tb & 0x000F000Fu;
            o1f[0] = wa.x & 0x00FF00FFu;  o1f[2] = ta;
            o1f[1] = wb.x & 0x00FF00FFu;  o1f[3] = tb;
            e2f[0] = wa.y & 0x000F000Fu;  e2f[2] = ua & 0x000F000Fu;
            e2f[1] = wb.y & 0x000F000Fu;  e2f[3] = ub & 0x000F000Fu;
            o2f[0] = wa.y & 0x00FF00FFu;  o2f[2] = ua;
            o2f[1] = wb.y & 0x00FF00FFu;  o2f[3] = ub;
            mma16816(de, e1f, b1, zero);
            mma16816(de, e2f, b2, de);
            mma16816(dd, o1f, b1, zero);
            mma16816(dd, o2f, b2, dd);
            if (L.t == c) {  // columns 2c (sub-block 2c) and 2c + 1 live in this lane
                a0 = fmaf(ca.x, de[0], a0);
                a2 = fmaf(cb.x, de[2], a2);
                a1 = fmaf(ca.y, dd[1] - de[1], a1);
                a3 = fmaf(cb.y, dd[3] - de[3], a3);
            }
#else
            e1f[0] = wa.x & 0x000F000Fu;  e1f[2] = (wa.x >> 8) & 0x000F000Fu;
            o1f[0] = wa.x & 0x00F000F0u;  o1f[2] = (wa.x >> 8) & 0x00F000F0u;
            e1f[1] = wb.x & 0x000F000Fu;  e1f[3] = (wb.x >> 8) & 0x000F000Fu;
            o1f[1] = wb.x & 0x00F000F0u;  o1f[3] = (wb.x >> 8) & 0x00F000F0u;
            e2f[0] = wa.y & 0x000F000Fu;  e2f[2] = (wa.y >> 8) & 0x000F000Fu;
            o2f[0] = wa.y & 0x00F000F0u;  o2f[2] = (wa.y >> 8) & 0x00F000F0u;
            e2f[1] = wb.y & 0x000F000Fu;  e2f[3] = (wb.y >> 8) & 0x000F000Fu;
            o2f[1] = wb.y & 0x00F000F0u;  o2f[3] = (wb.y >> 8) & 0x00F000F0u;
            mma16816(de, e1f, b1, zero);
            mma16816(de, e2f, b2, de);
            mma16816(dd, o1f, b1, zero);
            mma16816(dd, o2f, b2, dd);
            if (L.t == c) {  // columns 2c (sub-block 2c) and 2c + 1 live in this lane
                a0 = fmaf(ca.x, de[0], a0);
                a2 = fmaf(cb.x, de[2], a2);
                a1 = fmaf(ca.y, dd[1], a1);
                a3 = fmaf(cb.y, dd[3], a3);
            }
#endif
        }
    }
    acc.v[0][0] = a0;
    acc.v[0][1] = a1;
    acc.v[0][2] = a2;
    acc.v[0][3] = a3;
}


// GEMV accumulators -> the standard C-fragment layout (column 0 = the token): v[0] = row g, v[2] = row g + 8
GGQ_DEV void gemv_finalize(Acc<1>& acc) {
    float u = acc.v[0][0] + acc.v[0][1], v = acc.v[0][2] + acc.v[0][3];
    u += shfl_xor(u, 1);
    v += shfl_xor(v, 1);
    u += shfl_xor(u, 2);
    v += shfl_xor(v, 2);
    acc.v[0][0] = u;
    acc.v[0][1] = 0.f;
    acc.v[0][2] = v;
    acc.v[0][3] = 0.f;
}

// =============================================================================================
// Q6_K
// =============================================================================================
// prep: scratch[blk * 16 + row][(h*2 + lh)*4 + grp] = d * sc[8h + 2grp + lh] * 2^24  (fp32, exact)
template <bool FULL>
GGQ_DEV void prep_q6_k(const Lane& L, const StageArgs& s) {
    using G = Geo<2>;
    for (int p = L.lane; p < 16 * (FULL ? G::PREP_BLOCKS : s.nblk); p += 32) {
        const int row = p & 15, blk = p >> 4;
        const uint8_t* b = s.rows + row * G::SLOT + s.data_off + blk * G::BLK;  // 2-byte aligned
        const float d = h2f(ld16(b + 208)) * 16777216.f;
        float sc[16];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {  // scales 2jj, 2jj+1
            const uint32_t two = ld16(b + 192 + 2 * jj);
            sc[2 * jj] = d * static_cast<float>(static_cast<int>(static_cast<int8_t>(two & 0xff)));
            sc[2 * jj + 1] = d * static_cast<float>(static_cast<int>(static_cast<int8_t>(two >> 8)));
        }
        float4* out = reinterpret_cast<float4*>(s.scratch + (blk * 16 + row) * G::SCRATCH_PER_BLOCK);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int lh = 0; lh < 2; ++lh) {  // sub-block j = 8h + 2grp + lh  ->  out[h*2 + lh] = (grp 0..3)
                float4 v;
                v.x = sc[8 * h + 0 + lh];
                v.y = sc[8 * h + 2 + lh];
                v.z = sc[8 * h + 4 + lh];
                v.w = sc[8 * h + 6 + lh];
                out[h * 2 + lh] = v;
            }
    }
}

// 32-bit load at a 2-byte aligned address (odd blocks of a Q6_K row start at 2 mod 4).
template <bool ODD> GGQ_DEV uint32_t ldw(const uint8_t* p) {
    if (!ODD) return ld32(p);
    const uint8_t* a = p - 2;
    return funnelshift_r(ld32(a), ld32(a + 4), 16);
}

// ql word A (elements l..l+3 of groups 0 and 2), ql word B (groups 1 and 3), qh word -> four words of
// 6-bit quants, one per group (q6_k_ref.c:320-336)
GGQ_DEV void q6_bytes(uint32_t qla, uint32_t qlb, uint32_t qh, uint32_t g[4]) {
    g[0] = (qla & 0x0F0F0F0Fu) | ((qh << 4) & 0x30303030u);
    g[1] = (qlb & 0x0F0F0F0Fu) | ((qh << 2) & 0x30303030u);
    g[2] = ((qla >> 4) & 0x0F0F0F0Fu) | (qh & 0x30303030u);
    g[3] = ((qlb >> 4) & 0x0F0F0F0Fu) | ((qh >> 2) & 0x30303030u);
}

template <int NT, bool ODD>
GGQ_DEV void compute_q6_k_block(const Lane& L, const StageArgs& s, int i, const uint8_t* r0, const uint8_t* r1,
                                Acc<NT>& acc) {
    using G = Geo<2>;
    const uint8_t* b0 = r0 + i * G::BLK;
    const uint8_t* b1 = r1 + i * G::BLK;
    const uint8_t* sc0 = s.scratch + (i * 16 + L.g) * G::SCRATCH_PER_BLOCK;
    const uint8_t* sc1 = sc0 + 8 * G::SCRATCH_PER_BLOCK;
    uint2 xk6[NT][2];  // activation fragments; lanes without a token keep whatever these hold
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) xk6[nt][0] = xk6[nt][1] = uint2{0u, 0u};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int lh = 0; lh < 2; ++lh) {
            const int l = 16 * lh + 4 * L.t;
            uint32_t ga[4], gb[4];
            q6_bytes(ldw<ODD>(b0 + 64 * h + l), ldw<ODD>(b0 + 64 * h + 32 + l), ldw<ODD>(b0 + 128 + 32 * h + l), ga);
            q6_bytes(ldw<ODD>(b1 + 64 * h + l), ldw<ODD>(b1 + 64 * h + 32 + l), ldw<ODD>(b1 + 128 + 32 * h + l), gb);
            const float4 sa = ld128f(sc0 + 16 * (h * 2 + lh)), sb = ld128f(sc1 + 16 * (h * 2 + lh));
            const float sca[4] = {sa.x, sa.y, sa.z, sa.w}, scb[4] = {sb.x, sb.y, sb.z, sb.w};
#pragma unroll
            for (int grp = 0; grp < 4; ++grp) {
                uint32_t fa[4];
                fa[0] = prmt(ga[grp], 0u, 0x5140);  // q6 * 2^-24
                fa[2] = prmt(ga[grp], 0u, 0x7362);
                fa[1] = prmt(gb[grp], 0u, 0x5140);
                fa[3] = prmt(gb[grp], 0u, 0x7362);
                const int kk = s.k0 + 256 * i + 128 * h + 32 * grp + l;  // this lane's first activation
                const int j16 = (s.k0 + 256 * i) / 16 + 8 * h + 2 * grp + lh;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    uint2& xv = xk6[nt][grp & 1];
                    ld64k(xv, s.xrow[nt] + 2 * kk, s.xv[nt]);
                    const float2 c = ld64f(s.tbl + j16 * (8 * NT) + 8 * nt + 2 * L.t);
                    float d[4] = {c.x, c.y, c.x, c.y};
                    const uint32_t bf[2] = {xv.x, xv.y};
                    mma16816(d, fa, bf, d);
                    acc.v[nt][0] = fmaf(sca[grp], d[0], acc.v[nt][0]);
                    acc.v[nt][1] = fmaf(sca[grp], d[1], acc.v[nt][1]);
                    acc.v[nt][2] = fmaf(scb[grp], d[2], acc.v[nt][2]);
                    acc.v[nt][3] = fmaf(scb[grp], d[3], acc.v[nt][3]);
                }
            }
        }
    }
}

template <int NT>
GGQ_DEV void compute_q6_k(const Lane& L, const StageArgs& s, Acc<NT>& acc) {
    using G = Geo<2>;
    const uint8_t* r0 = s.rows + L.g * G::SLOT + s.data_off;
    const uint8_t* r1 = r0 + 8 * G::SLOT;
#pragma unroll
    for (int i = 0; i < G::PREP_BLOCKS; i += 2) {  // chunk starts at an even block: even blocks 4-byte aligned
        // (rows are whole 16-byte vectors => K/256 is a multiple of 8 => every sub-step is full: no bounds check)
        compute_q6_k_block<NT, false>(L, s, i, r0, r1, acc);
        compute_q6_k_block<NT, true>(L, s, i + 1, r0, r1, acc);
    }
}


// ---- Q6_K, one token (GEMV): the 8 MMA columns carry the 8 sub-blocks (16 weights each) of half a block ----------
// prep: scratch entry = 64 B per (row, block); piece t = d * 2^24 * {sc[2t], sc[2t+1], sc[8+2t], sc[8+2t+1]} (the four
// sub-blocks whose column lives in lane t), stored at position t ^ ((row >> 1) & 3) (conflict-free stores and loads).
template <bool FULL>
GGQ_DEV void prep_q6_k_gv(const Lane& L, const StageArgs& s) {
    using G = Geo<2>;
    for (int p = L.lane; p < 16 * (FULL ? G::PREP_BLOCKS : s.nblk); p += 32) {
        const int row = p & 15, blk = p >> 4;
        const uint8_t* b = s.rows + row * G::SLOT + s.data_off + blk * G::BLK;  // 2-byte aligned
        const float d = h2f(ld16(b + 208)) * 16777216.f;
        // int8 -> fp32 without the integer pipe: byte ^ 0x80 dropped into the mantissa of 2^23; (2^23 + u) - (2^23 + 128)
        // is exact, and so is the product with d (<= 19 significant bits)
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = (ld16(b + 192 + 4 * k) | (ld16(b + 194 + 4 * k) << 16)) ^ 0x80808080u;
        float sc[16];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            sc[4 * k + 0] = (u2f(prmt(w[k], 0x4B000000u, 0x7650)) - 8388736.f) * d;
            sc[4 * k + 1] = (u2f(prmt(w[k], 0x4B000000u, 0x7651)) - 8388736.f) * d;
            sc[4 * k + 2] = (u2f(prmt(w[k], 0x4B000000u, 0x7652)) - 8388736.f) * d;
            sc[4 * k + 3] = (u2f(prmt(w[k], 0x4B000000u, 0x7653)) - 8388736.f) * d;
        }
        uint8_t* e = s.scratch + (blk * 16 + row) * 64;
        const int sw = ((row >> 1) & 3) << 4;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float4 v;
            v.x = sc[2 * t];
            v.y = sc[2 * t + 1];
            v.z = sc[8 + 2 * t];
            v.w = sc[8 + 2 * t + 1];
            *reinterpret_cast<float4*>(e + (sw ^ (16 * t))) = v;
        }
    }
}

template <bool ODD>
GGQ_DEV void compute_q6_k_gv_block(const Lane& L, const StageArgs& s, int i, const uint8_t* r0, const uint8_t* r1,
                                   float& a0, float& a1, float& a2, float& a3) {
    using G = Geo<2>;
    const uint8_t* b0 = r0 + i * G::BLK;
    const uint8_t* b1 = r1 + i * G::BLK;
    const int piece = (L.t ^ ((L.g >> 1) & 3)) << 4;
    const uint8_t* e0 = s.scratch + (i * 16 + L.g) * 64 + piece;
    const float4 sa = ld128f(e0), sb = ld128f(e0 + 8 * 64);
    const float sca[4] = {sa.x, sa.y, sa.z, sa.w}, scb[4] = {sb.x, sb.y, sb.z, sb.w};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int kbh = s.k0 + 256 * i + 128 * h;
        const uint2 xr = ld64(s.xrow[0] + 2 * (kbh + 16 * L.g + 4 * L.t));  // sub-block g of this half, elements 4t..4t+3
        const uint32_t bf[2] = {xr.x, xr.y};
        const float2 tb = ld64f(s.tbl + (kbh >> 4) + 2 * L.t);              // -32 * 2^-24 * sum(x) of sub-blocks 2t, 2t+1
        const float c[4] = {tb.x, tb.y, tb.x, tb.y};
#pragma unroll
        for (int lh = 0; lh < 2; ++lh) {
            const int l = 16 * lh + 4 * L.t;
            uint32_t ga[4], gb[4];
            q6_bytes(ldw<ODD>(b0 + 64 * h + l), ldw<ODD>(b0 + 64 * h + 32 + l), ldw<ODD>(b0 + 128 + 32 * h + l), ga);
            q6_bytes(ldw<ODD>(b1 + 64 * h + l), ldw<ODD>(b1 + 64 * h + 32 + l), ldw<ODD>(b1 + 128 + 32 * h + l), gb);
#pragma unroll
            for (int grp = 0; grp < 4; ++grp) {  // sub-block 8h + 2grp + lh = column 2grp + lh: lane t == grp
                uint32_t fa[4];
                fa[0] = prmt(ga[grp], 0u, 0x5140);  // q6 * 2^-24
                fa[2] = prmt(ga[grp], 0u, 0x7362);
                fa[1] = prmt(gb[grp], 0u, 0x5140);
                fa[3] = prmt(gb[grp], 0u, 0x7362);
                float d[4];
                mma16816(d, fa, bf, c);
                if (L.t == grp) {
                    if (lh == 0) {
                        a0 = fmaf(sca[2 * h], d[0], a0);
                        a2 = fmaf(scb[2 * h], d[2], a2);
                    } else {
                        a1 = fmaf(sca[2 * h + 1], d[1], a1);
                        a3 = fmaf(scb[2 * h + 1], d[3], a3);
                    }
                }
            }
        }
    }
}

GGQ_DEV void compute_q6_k_gv(const Lane& L, const StageArgs& s, Acc<1>& acc) {
    using G = Geo<2>;
    const uint8_t* r0 = s.rows + L.g * G::SLOT + s.data_off;
    const uint8_t* r1 = r0 + 8 * G::SLOT;
    float a0 = acc.v[0][0], a1 = acc.v[0][1], a2 = acc.v[0][2], a3 = acc.v[0][3];
#pragma unroll
    for (int i = 0; i < G::PREP_BLOCKS; i += 2) {
        compute_q6_k_gv_block<false>(L, s, i, r0, r1, a0, a1, a2, a3);
        compute_q6_k_gv_block<true>(L, s, i + 1, r0, r1, a0, a1, a2, a3);
    }
    acc.v[0][0] = a0;
    acc.v[0][1] = a1;
    acc.v[0][2] = a2;
    acc.v[0][3] = a3;
}

// ---- activation staging pass (after the TMA copies of the K-slice have landed) ------------------
// Builds the per-slice table every thread block needs next to the raw fp16 rows:
//   Q8_0 / Q6_K: tbl[j][col] = TBL_MUL * sum of the GROUP activations of group j of token col (fp32)
//   Q4_K:        the bf16 hi/lo B fragments of the min-term MMA, and permutes the rows in place
// `ne` activations per token row (multiple of 256 for Q4_K), rows `x_stride` bytes apart.
template <int FMT, int NT, bool GV = false>
GGQ_DEV void stage_activations(uint8_t* xs, uint32_t x_stride, float* tbl, int ne, int T, int tid, int nthreads) {
    using G = Geo<FMT>;
    constexpr int TPAD = 8 * NT;
#ifdef __CUDACC__
    if (FMT == 1) {
        // Q4_K on the device: one 16-byte vector (8 activations) per thread and step -- permute it in place, add its 8
        // values, fold the 4 vectors of a 32-activation sub-block with two shuffles.  (The host emulator keeps the
        // one-thread-per-64-activations form below; the sums differ in the last bits only.)  `ne` is a multiple of 256,
        // so a warp is always entirely inside or entirely outside the T * ne / 8 vectors.
        const int vpr = ne / 8, total = T * vpr, lane = tid & 31;
        uint2* xb = reinterpret_cast<uint2*>(tbl);
        for (int idx = tid; idx < total; idx += nthreads) {
            const int tok = idx / vpr, v = idx - tok * vpr;
            uint4* ptr = reinterpret_cast<uint4*>(xs + tok * x_stride) + v;
            const uint4 q = *ptr;
            float sum = h2f(q.x & 0xffffu);
            sum += h2f(q.x >> 16);
            sum += h2f(q.y & 0xffffu);
            sum += h2f(q.y >> 16);
            sum += h2f(q.z & 0xffffu);
            sum += h2f(q.z >> 16);
            sum += h2f(q.w & 0xffffu);
            sum += h2f(q.w >> 16);
            uint4 o;  // (x0 x1 x2 x3) -> (x0 x2 x1 x3), per group of four
            o.x = prmt(q.x, q.y, 0x5410);
            o.y = prmt(q.x, q.y, 0x7632);
            o.z = prmt(q.z, q.w, 0x5410);
            o.w = prmt(q.z, q.w, 0x7632);
            *ptr = o;
            sum += shfl_xor(sum, 1);
            sum += shfl_xor(sum, 2);               // sum of sub-block v / 4
            const float other = shfl_xor(sum, 4);  // ... and of its neighbour in the 64-activation pair
            if (GV) {
                if ((lane & 3) == 0) tbl[v >> 2] = sum;
            } else if ((lane & 7) == 0) {
                const int e = v >> 3, b = e >> 2, t = e & 3;
                const uint32_t h0 = f2u(sum) & 0xffff0000u, h1 = f2u(other) & 0xffff0000u;  // bf16 hi (truncated), lo = rest
                const uint32_t l0 = bf16_bits_rn(sum - u2f(h0)), l1 = bf16_bits_rn(other - u2f(h1));
                xb[(b * TPAD + tok) * 4 + t] = uint2{(h0 >> 16) | h1, l0 | (l1 << 16)};
            }
        }
        if (!GV) {  // token columns that do not exist read as zero
            const int nent = ne / 64;
            for (int idx = tid; idx < nent * (TPAD - T); idx += nthreads) {
                const int e = idx % nent, col = T + idx / nent;
                xb[((e >> 2) * TPAD + col) * 4 + (e & 3)] = uint2{0u, 0u};
            }
        }
        return;
    }
#endif
    if (FMT == 1 && GV) {  // one token: tbl[sub-block] = fp32 sum of its 32 activations
        for (int e = tid; e < ne / 64; e += nthreads) q4k_stage_pair_gv(xs + e * 128, tbl + 2 * e);
    } else if (GV) {       // one token: tbl[group] = TBL_MUL * sum of its GROUP activations
        for (int j = tid; j < ne / G::GROUP; j += nthreads) {
            const uint8_t* src = xs + j * G::GROUP * 2;
            float sum = 0.f;
#pragma unroll
            for (int v = 0; v < G::GROUP / 8; ++v) {
                const uint4 q = ld128(src + 16 * v);
                sum += h2f(q.x & 0xffffu);
                sum += h2f(q.x >> 16);
                sum += h2f(q.y & 0xffffu);
                sum += h2f(q.y >> 16);
                sum += h2f(q.z & 0xffffu);
                sum += h2f(q.z >> 16);
                sum += h2f(q.w & 0xffffu);
                sum += h2f(q.w >> 16);
            }
            tbl[j] = sum * G::TBL_MUL;
        }
    } else if (FMT == 1) {
        uint2* xb = reinterpret_cast<uint2*>(tbl);
        const int nent = ne / 64;  // entries per token
        for (int idx = tid; idx < nent * TPAD; idx += nthreads) {
            const int e = idx % nent, col = idx / nent;  // consecutive threads -> consecutive 128-byte pieces of a row
            const int b = e >> 2, t = e & 3;
            q4k_stage_pair(xs + (col < T ? col : 0) * x_stride + e * 128, xb + (b * TPAD + col) * 4 + t, col < T);
        }
    } else {
        const int ngrp = ne / G::GROUP;
        for (int idx = tid; idx < ngrp * TPAD; idx += nthreads) {
            const int j = idx / TPAD, col = idx % TPAD;
            float sum = 0.f;
            if (col < T) {
                const uint8_t* src = xs + col * x_stride + j * G::GROUP * 2;
#pragma unroll
                for (int v = 0; v < G::GROUP / 8; ++v) {
                    const uint4 q = ld128(src + 16 * v);
                    sum += h2f(q.x & 0xffffu);
                    sum += h2f(q.x >> 16);
                    sum += h2f(q.y & 0xffffu);
                    sum += h2f(q.y >> 16);
                    sum += h2f(q.z & 0xffffu);
                    sum += h2f(q.z >> 16);
                    sum += h2f(q.w & 0xffffu);
                    sum += h2f(q.w >> 16);
                }
            }
            tbl[idx] = sum * G::TBL_MUL;
        }
    }
}

// ---- format dispatch ---------------------------------------------------------------------------
// FULL = the sub-step holds all PREP_BLOCKS blocks (the common case: no bounds checks, prep is a single pass)
template <int FMT, int NT, bool GV = false> struct Tile;
template <> struct Tile<0, 1, true> {
    template <bool FULL> static GGQ_DEV void prep(const Lane&, const StageArgs&) {}
    template <bool FULL> static GGQ_DEV void compute(const Lane& L, const StageArgs& s, Acc<1>& a) { compute_q8_0_gv<FULL>(L, s, a); }
};
template <> struct Tile<2, 1, true> {
    template <bool FULL> static GGQ_DEV void prep(const Lane& L, const StageArgs& s) { prep_q6_k_gv<FULL>(L, s); }
    template <bool FULL> static GGQ_DEV void compute(const Lane& L, const StageArgs& s, Acc<1>& a) { compute_q6_k_gv(L, s, a); }
};
template <> struct Tile<1, 1, true> {
    template <bool FULL> static GGQ_DEV void prep(const Lane& L, const StageArgs& s) { prep_q4_k_gv<FULL>(L, s); }
    template <bool FULL> static GGQ_DEV void compute(const Lane& L, const StageArgs& s, Acc<1>& a) { compute_q4_k_gv_impl<FULL>(L, s, a); }
};
template <int NT> struct Tile<0, NT, false> {
    template <bool FULL> static GGQ_DEV void prep(const Lane&, const StageArgs&) {}
    template <bool FULL> static GGQ_DEV void compute(const Lane& L, const StageArgs& s, Acc<NT>& a) { compute_q8_0<NT, FULL>(L, s, a); }
};
template <int NT> struct Tile<1, NT, false> {
    template <bool FULL> static GGQ_DEV void prep(const Lane& L, const StageArgs& s) { prep_q4_k<FULL>(L, s); }
    template <bool FULL> static GGQ_DEV void compute(const Lane& L, const StageArgs& s, Acc<NT>& a) { compute_q4_k_impl<NT, FULL>(L, s, a); }
};
template <int NT> struct Tile<2, NT, false> {
    template <bool FULL> static GGQ_DEV void prep(const Lane& L, const StageArgs& s) { prep_q6_k<FULL>(L, s); }
    template <bool FULL> static GGQ_DEV void compute(const Lane& L, const StageArgs& s, Acc<NT>& a) { compute_q6_k<NT>(L, s, a); }
};

}  // namespace dec
}  // namespace ggq
