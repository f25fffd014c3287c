// refmode.cu — "reference arithmetic" mode: Q8_1-quantized activations, integer block dots, fp16 accumulation,
// in exactly the operation order of the reference's CPU implementations, so the result equals
// kernels/cpu_impls/mmq_*_q8_1_cpu bit for bit.  One thread per output element walks the blocks
// sequentially — the fp16 accumulation chain of an output IS sequential in the reference — and consecutive threads own
// consecutive weight rows of the same token.  Four kernels: `refmode_mma_q4k_kernel` (Q4_K, T >= 3: the block dots of
// 16 rows x 8 tokens by one integer tensor-core MMA, the chains in the accumulator-fragment layout),
// `refmode_tile_q4k_kernel` (Q4_K: rows staged through a
// cp.async-filled shared-memory tile, HBM-friendly), `refmode_fast_kernel` (rows that are whole 32-bit words: 128-bit /
// 32-bit vector loads; both with the integer block dots on DP4A and the activation block read as a broadcast) and the
// byte-wise `refmode_kernel` for every other shape.  The integer dots are exact in any order; the floating-point
// operations are the same explicitly rounded intrinsics in the same order in both.
//   Q8_0  kernels/cpu_impls/mmq_q8_0_q8_1_cpu.py:37-54   r = fp16(fp16(d_w*d_x) * dot);            C = fp16(C + r)
//   Q4_K  kernels/cpu_impls/mmq_q4_k_q8_1_cpu.py:94-117  r = ((d*sc)*d_x)*dot - (dmin*m)*s_x (fp32); C = fp16(C + fp16(r))
//   Q6_K  kernels/cpu_impls/mmq_q6_k_q8_1_cpu.py:117-150 r = d_x*((d*sc1)*dot1 + (d*sc2)*dot2) (fp32); C = fp16(C + fp16(r))
// Every fp32 operation is an explicitly rounded intrinsic (__fmul_rn / __fadd_rn / __fsub_rn): no FMA contraction.
#include <cstdlib>

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

// ---- vectorised form -----------------------------------------------------------------------------------------------
// The accumulator of the vectorised kernels is kept in fp16 and advanced with ONE half-precision add: fp16(fp32(c + r16))
// — what torch's CPU kernel and the byte-wise kernel above compute — equals the correctly rounded fp16 sum, because a
// second rounding from a format of P >= 2p + 2 significand bits (24 >= 2 * 11 + 2) is innocuous for +, -, *, / (Figueroa).
// __hadd_rn / __hadd2_rn are never contracted into an FMA.
__device__ __forceinline__ __half acc16h(__half c, float r32) { return __hadd_rn(c, __float2half_rn(r32)); }
__device__ __forceinline__ float hlo(uint32_t w) { return __half2float(__ushort_as_half(static_cast<unsigned short>(w & 0xffffu))); }
__device__ __forceinline__ float hhi(uint32_t w) { return __half2float(__ushort_as_half(static_cast<unsigned short>(w >> 16))); }
__device__ __forceinline__ int dp4(uint32_t a, uint32_t b, int c) { return __dp4a(static_cast<int>(a), static_cast<int>(b), c); }
__device__ __forceinline__ int dp4_us(uint32_t a, uint32_t b, int c) {   // unsigned bytes of a x signed bytes of b
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// n + 1 words from the 4-byte aligned address at or 2 bytes below p, shifted so that out[i] = the word at p + 4 i
template <int N>
__device__ __forceinline__ void ld_words(const uint8_t* p, uint32_t (&out)[N]) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~uintptr_t{3});
    if ((a & 2) == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = __ldg(q + i);
    } else {
        uint32_t prev = __ldg(q);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint32_t next = __ldg(q + i + 1);
            out[i] = __funnelshift_r(prev, next, 16);
            prev = next;
        }
    }
}

// Requirements (checked by the launcher): W 16-byte aligned, XQ 4-byte aligned, rows whole 32-bit words (Q8_0 / Q6_K: an
// even number of blocks per row).  A thread owns ONE weight row and TT consecutive tokens: the row's blocks are loaded and
// unpacked once per block, the activation words are warp-uniform loads (every lane of a warp reads the same address).
template <int FMT, int TT>
__global__ void __launch_bounds__(128) refmode_fast_kernel(const uint8_t* __restrict__ W, const uint8_t* __restrict__ XQ,
                                                           __half* __restrict__ C, int64_t O, int64_t T, int64_t K) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t groups = (T + TT - 1) / TT;
    if (idx >= O * groups) return;
    const int64_t o = idx % O, t0 = (idx / O) * TT;
    const int nt = static_cast<int>(T - t0 < TT ? T - t0 : TT);   // warp-uniform unless the warp straddles two token groups
    const int64_t nb32 = K / 32;
    const int64_t xstride = nb32 * 9;                                               // words per token
    const uint32_t* xrow = reinterpret_cast<const uint32_t*>(XQ) + t0 * xstride;    // 9 words per Q8_1 block: {d, s}, qs[32]
    __half c[TT];
#pragma unroll
    for (int tt = 0; tt < TT; ++tt) c[tt] = __ushort_as_half(0);
    if (FMT == GGQ_Q8_0) {
        const uint32_t* wrow = reinterpret_cast<const uint32_t*>(W + o * nb32 * 34);
        for (int64_t b = 0; b < nb32; b += 2) {   // two blocks = 68 bytes = 17 words
            uint32_t w[17], q[2][8];
#pragma unroll
            for (int i = 0; i < 17; ++i) w[i] = __ldg(wrow + (b >> 1) * 17 + i);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                q[0][i] = __funnelshift_r(w[i], w[i + 1], 16);
                q[1][i] = w[9 + i];
            }
            const float dw[2] = {hlo(w[0]), hhi(w[8])};
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int tt = 0; tt < TT; ++tt) {
                    if (tt >= nt) break;
                    const uint32_t* xb = xrow + tt * xstride + (b + half) * 9;
                    const uint32_t x0 = __ldg(xb);
                    int dot = 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) dot = dp4(q[half][i], __ldg(xb + 1 + i), dot);
                    const float sc = __half2float(__float2half_rn(__fmul_rn(dw[half], hlo(x0))));  // fp16 * fp16 -> fp16
                    c[tt] = acc16h(c[tt], __fmul_rn(sc, static_cast<float>(dot)));
                }
            }
        }
    } else if (FMT == GGQ_Q4_K) {
        const uint4* wrow = reinterpret_cast<const uint4*>(W + o * (K / 256) * 144);   // 9 vectors per super-block
        for (int64_t sb = 0; sb < K / 256; ++sb) {
            const uint4 h = __ldg(wrow + sb * 9);
            const float d = hlo(h.x), dmin = hhi(h.x);
#pragma unroll
            for (int pr = 0; pr < 4; ++pr) {   // sub-blocks 2 pr (low nibbles) and 2 pr + 1 (high nibbles) share 32 bytes
                const uint4 qa = __ldg(wrow + sb * 9 + 1 + 2 * pr), qb = __ldg(wrow + sb * 9 + 2 + 2 * pr);
                const uint32_t w[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int j = 2 * pr + hf;
                    int sc, m;   // q4_k_ref.c:174-186, from the header words h.y = s[0..3], h.z = s[4..7], h.w = s[8..11]
                    if (j < 4) {
                        sc = (h.y >> (8 * j)) & 63;
                        m = (h.z >> (8 * j)) & 63;
                    } else {
                        const int sh = 8 * (j - 4);
                        sc = ((h.w >> sh) & 0x0F) | ((((h.y >> sh) & 0xFF) >> 6) << 4);
                        m = (((h.w >> sh) & 0xFF) >> 4) | ((((h.z >> sh) & 0xFF) >> 6) << 4);
                    }
                    uint32_t q[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) q[i] = (hf ? (w[i] >> 4) : w[i]) & 0x0F0F0F0Fu;
                    const float dsc = __fmul_rn(d, static_cast<float>(sc)), dm = __fmul_rn(dmin, static_cast<float>(m));
#pragma unroll
                    for (int tt = 0; tt < TT; ++tt) {
                        if (tt >= nt) break;
                        const uint32_t* xb = xrow + tt * xstride + (sb * 8 + j) * 9;
                        const uint32_t x0 = __ldg(xb);
                        int dot = 0;
#pragma unroll
                        for (int i = 0; i < 8; ++i) dot = dp4(q[i], __ldg(xb + 1 + i), dot);
                        const float t1 = __fmul_rn(__fmul_rn(dsc, hlo(x0)), static_cast<float>(dot));
                        const float t2 = __fmul_rn(dm, hhi(x0));
                        c[tt] = acc16h(c[tt], __fsub_rn(t1, t2));
                    }
                }
            }
        }
    } else {
        const uint8_t* wrow = W + o * (K / 256) * 210;
        for (int64_t sb = 0; sb < K / 256; ++sb) {
            const uint8_t* wb = wrow + sb * 210;   // 2-byte aligned (odd blocks of a row start 2 bytes into a word)
            uint32_t scw[5];                       // bytes 192..211: 16 int8 scales, d
            ld_words<5>(wb + 192, scw);
            const float d = hlo(scw[4]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t ql[16], qh[8];
                ld_words<16>(wb + 64 * h, ql);
                ld_words<8>(wb + 128 + 32 * h, qh);
#pragma unroll
                for (int g = 0; g < 4; ++g) {      // j = 4 h + g: weights 128 h + 32 g + l, l = 0..31 (q6_k_ref.c:320-336)
                    const int j = 4 * h + g;
                    uint32_t q[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t lo = (ql[8 * (g & 1) + i] >> (4 * (g >> 1))) & 0x0F0F0F0Fu;
                        const uint32_t hi = ((qh[i] >> (2 * g)) & 0x03030303u) << 4;
                        q[i] = __vsub4(lo | hi, 0x20202020u);   // q6 - 32 in every byte, as int8
                    }
                    const uint32_t s2 = (scw[j >> 1] >> (16 * (j & 1))) & 0xffffu;   // scales 2 j, 2 j + 1
                    const float s1 = __fmul_rn(d, static_cast<float>(static_cast<int>(static_cast<int8_t>(s2 & 0xffu))));
                    const float s2f = __fmul_rn(d, static_cast<float>(static_cast<int>(static_cast<int8_t>(s2 >> 8))));
#pragma unroll
                    for (int tt = 0; tt < TT; ++tt) {
                        if (tt >= nt) break;
                        const uint32_t* xb = xrow + tt * xstride + (sb * 8 + j) * 9;
                        const uint32_t x0 = __ldg(xb);
                        int dot1 = 0, dot2 = 0;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            dot1 = dp4(q[i], __ldg(xb + 1 + i), dot1);
                            dot2 = dp4(q[4 + i], __ldg(xb + 5 + i), dot2);
                        }
                        const float inner = __fadd_rn(__fmul_rn(s1, static_cast<float>(dot1)), __fmul_rn(s2f, static_cast<float>(dot2)));
                        c[tt] = acc16h(c[tt], __fmul_rn(hlo(x0), inner));
                    }
                }
            }
        }
    }
#pragma unroll
    for (int tt = 0; tt < TT; ++tt)
        if (tt < nt) C[(t0 + tt) * O + o] = c[tt];
}

// ---- Q4_K, tiled: HBM-rate form of the same arithmetic ---------------------------------------------------------------
// One thread per weight row still (the accumulation chain of an output is sequential), but the packed rows reach the
// threads through shared memory: the CTA's 128 threads copy a tile of 128 rows x 1 super-block (144 B per row) with
// coalesced 16-byte cp.async into a double-buffered stage (odd row pitch in 16-byte vectors: the threads' 128-bit reads of
// their own row are conflict free; small stages keep five CTAs = 20 warps per SM for the single-token case, whose per-row
// chain of dependent operations needs the warps), and the Q8_1 activations of the CTA's token tile are staged once.
// Loads of the next stage overlap the arithmetic of the current one, so the kernel streams the weights once at HBM rate
// instead of waiting on every thread's own scattered 16-byte loads.
// Row pitch of a stage (template parameter RT_PITCH): 144 B = the rows packed densely, 9 vectors = an odd pitch, so the
// threads' 128-bit reads of their own rows are conflict free (vector slot (9 r + c) mod 8 = (r + c) mod 8) and so are the
// 32-bit fragment loads of the tensor-core kernel (bank 36 g + tg = 4 g + tg mod 32); 36 KB for both stages, five CTAs
// per SM.  176 B (11 vectors, four CTAs per SM) was the first version and stays selectable (GGQ_REFMODE_PITCH=176).
constexpr int RT_ROWS = 128, RT_SB = 1;
__device__ __forceinline__ void cp_async16_s(uint32_t dst_shared, const void* src) {   // destination as a shared-space address
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_shared), "l"(src) : "memory");
}
// The copy plan of a stage (source pointer and shared-memory address of each 16-byte piece a thread copies) is kept as
// LOOP-CARRIED state — sources advance by one super-block, destinations hop between the two buffers by +/- one buffer
// size — because ptxas otherwise re-derives a loop-invariant plan inside the stage loop (division by 9, two multiplies
// and 64-bit address assembly per piece: ~110 of the ~450 instructions per stage and thread) instead of holding it.
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int TT, int RT_PITCH>
__global__ void __launch_bounds__(RT_ROWS) refmode_tile_q4k_kernel(const uint8_t* __restrict__ W, const uint8_t* __restrict__ XQ,
                                                                   __half* __restrict__ C, int64_t O, int64_t T, int64_t K) {
    extern __shared__ __align__(16) uint8_t rsm[];
    const int tid = threadIdx.x;
    const int64_t row0 = static_cast<int64_t>(blockIdx.x) * RT_ROWS;
    const int64_t t0 = static_cast<int64_t>(blockIdx.y) * TT;
    const int nt = static_cast<int>(T - t0 < TT ? T - t0 : TT);
    const int nsb = static_cast<int>(K / 256);
    const int64_t rowB = static_cast<int64_t>(nsb) * 144;
    const int xwords = nsb * 8 * 9;                                // words per token in XQ
    const int xsw = nsb * 8 * 10;                                  // words per token in shared memory (d, s as two fp32)
    uint32_t* xs = reinterpret_cast<uint32_t*>(rsm);               // [TT][xsw]
    uint8_t* ws = rsm + static_cast<size_t>(TT) * xsw * 4;         // [2][RT_ROWS][RT_PITCH] (xsw * 4 is a multiple of 16)
    const int nstage = (nsb + RT_SB - 1) / RT_SB;

    // the 16-byte pieces this thread copies per stage (128 rows x 9 pieces, piece p = tid + 128 j: consecutive threads copy
    // consecutive pieces of a row) are the same in every stage: source row / destination offset are computed once
    static_assert(RT_SB == 1, "one super-block (9 vectors) per row and stage");
    const uint8_t* src[9];   // advanced by one super-block per issued stage
    uint32_t dsto[9];        // shared-space address of the piece in the buffer of the next stage
    const uint32_t ws_s = static_cast<uint32_t>(__cvta_generic_to_shared(ws));
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const int p = tid + RT_ROWS * j, r = p / 9, c = p - 9 * r;
        const int64_t row = min(row0 + r, O - 1);                  // rows past the end re-read the last row (never stored)
        src[j] = W + row * rowB + c * 16;
        dsto[j] = ws_s + static_cast<uint32_t>(r * RT_PITCH + c * 16);
    }
    int bstep = RT_ROWS * RT_PITCH;   // to the other buffer
    auto issue = [&](int) {   // the next stage (super-block st) -> buffer st & 1; called once per stage, in order
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            cp_async16_s(dsto[j], src[j]);
            src[j] += 144;
            dsto[j] += bstep;
        }
        bstep = -bstep;
        cp_async_commit();
    };
    issue(0);
    const int nblk = nsb * 8;
    {   // the token tile's activations, once (XQ rows are 36 * K/32 bytes: 4-byte aligned), regrouped per token as
        // (d, s) of all blocks converted to fp32 (exact), then the blocks' 8 quant words at 32-byte aligned offsets: a
        // thread reads a block with two 128-bit and one 64-bit broadcast load instead of nine 32-bit loads + 2 converts
        const uint32_t* xg = reinterpret_cast<const uint32_t*>(XQ) + t0 * xwords;
        for (int bi = tid; bi < nt * nblk; bi += RT_ROWS) {            // one Q8_1 block (9 words) per thread and turn
            const int tt = TT == 1 ? 0 : bi / nblk, b = bi - tt * nblk;
            const uint32_t* g = xg + static_cast<int64_t>(tt) * xwords + 9 * b;
            uint32_t v[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) v[i] = __ldg(g + i);
            uint32_t* xt = xs + tt * xsw;
            *reinterpret_cast<float2*>(xt + 2 * b) = make_float2(hlo(v[0]), hhi(v[0]));
            *reinterpret_cast<uint4*>(xt + 2 * nblk + 8 * b) = make_uint4(v[1], v[2], v[3], v[4]);
            *reinterpret_cast<uint4*>(xt + 2 * nblk + 8 * b + 4) = make_uint4(v[5], v[6], v[7], v[8]);
        }
    }
    __half c[TT];
#pragma unroll
    for (int tt = 0; tt < TT; ++tt) c[tt] = __ushort_as_half(0);
    for (int st = 0; st < nstage; ++st) {
        if (st + 1 < nstage) {
            issue(st + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();   // stage st (and, the first time, the activations) are visible to every thread
        const uint8_t* mine = ws + static_cast<size_t>(st & 1) * RT_ROWS * RT_PITCH + tid * RT_PITCH;
        const int nhere = min(RT_SB, nsb - st * RT_SB);
        for (int u = 0; u < nhere; ++u) {
            const int sb = st * RT_SB + u;
            const uint4* wb = reinterpret_cast<const uint4*>(mine + u * 144);
            const uint4 h = wb[0];
            const float d = hlo(h.x), dmin = hhi(h.x);
#pragma unroll
            for (int pr = 0; pr < 4; ++pr) {
                const uint4 qa = wb[1 + 2 * pr], qb = wb[2 + 2 * pr];
                const uint32_t w[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int j = 2 * pr + hf;
                    int sc, m;
                    if (j < 4) {
                        sc = (h.y >> (8 * j)) & 63;
                        m = (h.z >> (8 * j)) & 63;
                    } else {
                        const int sh = 8 * (j - 4);
                        sc = ((h.w >> sh) & 0x0F) | ((((h.y >> sh) & 0xFF) >> 6) << 4);
                        m = (((h.w >> sh) & 0xFF) >> 4) | ((((h.z >> sh) & 0xFF) >> 6) << 4);
                    }
                    // low nibbles: masked in place; high nibbles: left in place as unsigned bytes 16 q — their dot with
                    // the signed activation bytes is 16 x the sub-block's dot, exactly (one mask per word, no shift)
                    uint32_t q[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) q[i] = w[i] & (hf ? 0xF0F0F0F0u : 0x0F0F0F0Fu);
                    const float dsc = __fmul_rn(d, static_cast<float>(sc)), dm = __fmul_rn(dmin, static_cast<float>(m));
#pragma unroll
                    for (int tt = 0; tt < TT; ++tt) {
                        if (tt >= nt) break;
                        const uint32_t* xt = xs + tt * xsw;                         // the same addresses in every thread: broadcast
                        const float2 xds = *reinterpret_cast<const float2*>(xt + 2 * (sb * 8 + j));
                        const uint4 xa = *reinterpret_cast<const uint4*>(xt + 2 * nblk + 8 * (sb * 8 + j));
                        const uint4 xb = *reinterpret_cast<const uint4*>(xt + 2 * nblk + 8 * (sb * 8 + j) + 4);
                        const uint32_t xq[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
                        int dot = 0;
#pragma unroll
                        for (int i = 0; i < 8; ++i) dot = hf ? dp4_us(q[i], xq[i], dot) : dp4(q[i], xq[i], dot);
                        if (hf) dot >>= 4;
                        const float t1 = __fmul_rn(__fmul_rn(dsc, xds.x), static_cast<float>(dot));
                        const float t2 = __fmul_rn(dm, xds.y);
                        c[tt] = acc16h(c[tt], __fsub_rn(t1, t2));
                    }
                }
            }
        }
        __syncthreads();   // everyone is done with buffer st & 1 before stage st + 2 is copied into it
    }
    const int64_t row = row0 + tid;
    if (row < O) {
#pragma unroll
        for (int tt = 0; tt < TT; ++tt)
            if (tt < nt) C[(t0 + tt) * O + row] = c[tt];
    }
}

// ---- Q4_K, two or more tokens: the integer block dots on the tensor cores ---------------------------------------------
// For T >= 2 the thread-per-row form above spends its time in the per-token work of every sub-block (nine activation
// loads and eight DP4A for each 32-element dot).  Here a warp owns 32 weight rows and a tile of 8 tokens, and one
// mma.sync.m16n8k32 (s8 x s8 -> s32; nibbles 0..15 and Q8_1 quants are exact int8 operands, a 32-element sum of
// products is exact in int32) delivers the 16 x 8 dots of a sub-block: a thread then holds the dots of rows g, g + 8 and
// tokens 2 tg, 2 tg + 1 (g = lane / 4, tg = lane % 4) and runs those four outputs' floating-point chains — the same
// explicitly rounded operations in the same order, the two tokens of a row as one fp16x2 add.  The stage pipeline is the
// tile kernel's (128 rows x 1 super-block by cp.async, double buffered); the 8 tokens' Q8_1 blocks of the super-block
// travel with the stage (288 B per token, token pitch 400 B: B-fragment loads of the 8 tokens hit distinct banks), so the
// shared-memory footprint does not depend on K and four CTAs stay resident.  d*sc and dmin*m of a (row, sub-block) are
// computed once by the lane that owns the row (lane l <-> row 32 warp + l) and handed to the lanes that need them by
// warp shuffles.  Requirements (checked by the launcher): W and XQ 16-byte aligned.
constexpr int RM_TOK = 8, RM_XPITCH = 100;   // words per token and stage: 72 + 28, = 4 (mod 32)
__device__ __forceinline__ void mma_s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(0));
}

template <int RT_PITCH>
__global__ void __launch_bounds__(RT_ROWS) refmode_mma_q4k_kernel(const uint8_t* __restrict__ W, const uint8_t* __restrict__ XQ,
                                                                  __half* __restrict__ C, int64_t O, int64_t T, int64_t K) {
    extern __shared__ __align__(16) uint8_t rsm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
    const int64_t row0 = static_cast<int64_t>(blockIdx.x) * RT_ROWS;
    const int64_t t0 = static_cast<int64_t>(blockIdx.y) * RM_TOK;
    const int nt = static_cast<int>(T - t0 < RM_TOK ? T - t0 : RM_TOK);
    const int nsb = static_cast<int>(K / 256);
    const int64_t rowB = static_cast<int64_t>(nsb) * 144;
    const int64_t xrowB = static_cast<int64_t>(nsb) * 288;        // bytes of one token's Q8_1 row
    uint8_t* ws = rsm;                                                                   // [2][RT_ROWS][RT_PITCH]
    uint8_t* xs = rsm + 2 * RT_ROWS * RT_PITCH;                                          // [2][RM_TOK][RM_XPITCH words]

    // copy plan of a stage, the same in every stage: 9 weight pieces per thread (as in the tile kernel) and the
    // 8 tokens x 18 activation pieces over the first 144 piece slots (tokens past the end re-read the last token)
    const uint8_t* src[9];   // advanced by one super-block per issued stage
    uint32_t dsto[9];        // shared-space address of the piece in the buffer of the next stage
    const uint32_t ws_s = static_cast<uint32_t>(__cvta_generic_to_shared(ws));
    const uint32_t xs_s = static_cast<uint32_t>(__cvta_generic_to_shared(xs));
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const int p = tid + RT_ROWS * j, r = p / 9, c = p - 9 * r;
        const int64_t row = min(row0 + r, O - 1);
        src[j] = W + row * rowB + c * 16;
        dsto[j] = ws_s + static_cast<uint32_t>(r * RT_PITCH + c * 16);
    }
    const uint8_t* xsrc[2];
    uint32_t xdst[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int p = min(tid + RT_ROWS * j, RM_TOK * 18 - 1), tk = p / 18, c = p - 18 * tk;
        xsrc[j] = XQ + (t0 + min(tk, nt - 1)) * xrowB + c * 16;
        xdst[j] = xs_s + static_cast<uint32_t>(tk * (RM_XPITCH * 4) + c * 16);
    }
    int bstep = RT_ROWS * RT_PITCH, xstep = RM_TOK * RM_XPITCH * 4;   // to the other buffer
    auto issue = [&](int) {   // the next stage (super-block st) -> buffer st & 1; called once per stage, in order
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            cp_async16_s(dsto[j], src[j]);
            src[j] += 144;
            dsto[j] += bstep;
        }
        cp_async16_s(xdst[0], xsrc[0]);
        if (tid < RM_TOK * 18 - RT_ROWS) cp_async16_s(xdst[1], xsrc[1]);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            xsrc[j] += 288;
            xdst[j] += xstep;
        }
        bstep = -bstep;
        xstep = -xstep;
        cp_async_commit();
    };
    issue(0);
    __half2 acc[2][2];   // [16-row tile of the warp][row g / row g + 8] = tokens (2 tg, 2 tg + 1)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i >> 1][i & 1] = __half2half2(__ushort_as_half(0));
    for (int st = 0; st < nsb; ++st) {
        if (st + 1 < nsb) {
            issue(st + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();   // stage st is visible to every thread
        const uint8_t* wst = ws + static_cast<size_t>(st & 1) * RT_ROWS * RT_PITCH;
        const uint32_t* xst = reinterpret_cast<const uint32_t*>(xs + static_cast<size_t>(st & 1) * RM_TOK * RM_XPITCH * 4);
        // this lane's own row (row tid of the tile): d * sc and dmin * m of the 8 sub-blocks
        float dsc[8], dm[8];
        {
            const uint4 h = *reinterpret_cast<const uint4*>(wst + tid * RT_PITCH);
            const float d = hlo(h.x), dmin = hhi(h.x);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int sc, m;   // q4_k_ref.c:174-186, from the header words h.y = s[0..3], h.z = s[4..7], h.w = s[8..11]
                if (j < 4) {
                    sc = (h.y >> (8 * j)) & 63;
                    m = (h.z >> (8 * j)) & 63;
                } else {
                    const int sh = 8 * (j - 4);
                    sc = ((h.w >> sh) & 0x0F) | ((((h.y >> sh) & 0xFF) >> 6) << 4);
                    m = (((h.w >> sh) & 0xFF) >> 4) | ((((h.z >> sh) & 0xFF) >> 6) << 4);
                }
                dsc[j] = __fmul_rn(d, static_cast<float>(sc));
                dm[j] = __fmul_rn(dmin, static_cast<float>(m));
            }
        }
#pragma unroll
        for (int pr = 0; pr < 4; ++pr) {   // sub-blocks 2 pr (low nibbles) and 2 pr + 1 (high nibbles) share 32 bytes
            uint32_t wq[2][4];             // [16-row tile][row g: words tg, 4 + tg; row g + 8: words tg, 4 + tg]
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const uint32_t* wr = reinterpret_cast<const uint32_t*>(wst + (32 * warp + 16 * mt + g) * RT_PITCH) + 4 + 8 * pr;
                wq[mt][0] = wr[tg];
                wq[mt][1] = wr[4 + tg];
                wq[mt][2] = wr[8 * (RT_PITCH / 4) + tg];
                wq[mt][3] = wr[8 * (RT_PITCH / 4) + 4 + tg];
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int j = 2 * pr + hf;
                const uint32_t* xb = xst + g * RM_XPITCH + j * 9;          // B fragment: token g, bytes 4 tg.. and 16 + 4 tg..
                const uint32_t b0 = xb[1 + tg], b1 = xb[5 + tg];
                const uint32_t xw0 = xst[(2 * tg) * RM_XPITCH + j * 9], xw1 = xst[(2 * tg + 1) * RM_XPITCH + j * 9];
                const float dx0 = hlo(xw0), sx0 = hhi(xw0), dx1 = hlo(xw1), sx1 = hhi(xw1);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    int dot[4];            // (row g, token 2 tg), (g, 2 tg + 1), (g + 8, 2 tg), (g + 8, 2 tg + 1)
                    mma_s8(dot, (wq[mt][0] >> (4 * hf)) & 0x0F0F0F0Fu, (wq[mt][2] >> (4 * hf)) & 0x0F0F0F0Fu,
                           (wq[mt][1] >> (4 * hf)) & 0x0F0F0F0Fu, (wq[mt][3] >> (4 * hf)) & 0x0F0F0F0Fu, b0, b1);
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const int owner = 16 * mt + 8 * rr + g;            // the lane that holds this row's scales
                        const float ds = __shfl_sync(0xffffffffu, dsc[j], owner), dmn = __shfl_sync(0xffffffffu, dm[j], owner);
                        const float r0 = __fsub_rn(__fmul_rn(__fmul_rn(ds, dx0), static_cast<float>(dot[2 * rr])), __fmul_rn(dmn, sx0));
                        const float r1 = __fsub_rn(__fmul_rn(__fmul_rn(ds, dx1), static_cast<float>(dot[2 * rr + 1])), __fmul_rn(dmn, sx1));
                        acc[mt][rr] = __hadd2_rn(acc[mt][rr], __floats2half2_rn(r0, r1));
                    }
                }
            }
        }
        __syncthreads();   // everyone is done with buffer st & 1 before stage st + 2 is copied into it
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int64_t row = row0 + 32 * warp + 16 * mt + 8 * rr + g;
            if (row >= O) continue;
            if (2 * tg < nt) C[(t0 + 2 * tg) * O + row] = __low2half(acc[mt][rr]);
            if (2 * tg + 1 < nt) C[(t0 + 2 * tg + 1) * O + row] = __high2half(acc[mt][rr]);
        }
}

static int refmode_pitch() {   // dev: GGQ_REFMODE_PITCH=176 selects the first version's padded row pitch
    static const int pitch = getenv("GGQ_REFMODE_PITCH") && atoi(getenv("GGQ_REFMODE_PITCH")) == 176 ? 176 : 144;
    return pitch;
}

static bool launch_mma_q4k(const uint8_t* w, const uint8_t* x, __half* c, int64_t O, int64_t T, int64_t K, cudaStream_t s) {
    const int pitch = refmode_pitch();
    const int smem = 2 * RT_ROWS * pitch + 2 * RM_TOK * RM_XPITCH * 4;   // 43264 B: five CTAs per SM (176: 51456 B, four)
    const int64_t gx = (O + RT_ROWS - 1) / RT_ROWS, gy = (T + RM_TOK - 1) / RM_TOK;
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || gx > 0x7fffffff || gy > 65535) return false;
    auto go = [&](auto kern) {
        static bool configured[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !configured[dev]) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            configured[dev] = true;
        }
        kern<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(gy)), RT_ROWS, smem, s>>>(w, x, c, O, T, K);
    };
    if (pitch == 144) go(refmode_mma_q4k_kernel<144>);
    else go(refmode_mma_q4k_kernel<176>);
    return true;
}

// returns false when the shape does not fit (the activations of a token tile must fit next to the two stages)
static bool launch_tile_q4k(const uint8_t* w, const uint8_t* x, __half* c, int64_t O, int64_t T, int64_t K, cudaStream_t s) {
    // the largest token tile (8, 4, 1) whose activations fit next to the two stages with two CTAs per SM (113 KB each;
    // one token at K = 4096: 50 KB, four CTAs per SM);
    // a layer with more tokens than the tile streams its weights once per token tile
    constexpr size_t LIMIT = 113 * 1024;
    const int pitch = refmode_pitch();
    auto smem_of = [&](int t) { return static_cast<size_t>(t) * (K / 32) * 40 + 2 * static_cast<size_t>(RT_ROWS) * pitch; };
    int tt = T == 1 ? 1 : T <= 4 ? 4 : 8;
    while (tt > 1 && smem_of(tt) > LIMIT) tt = tt == 8 ? 4 : 1;
    const size_t smem = smem_of(tt);
    if (smem > LIMIT || (O + RT_ROWS - 1) / RT_ROWS > 0x7fffffff || (T + tt - 1) / tt > 65535) return false;
    const dim3 grid(static_cast<unsigned>((O + RT_ROWS - 1) / RT_ROWS), static_cast<unsigned>((T + tt - 1) / tt));
    auto go = [&](auto kern) {
        static bool configured[3][64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        const int slot = tt == 1 ? 0 : tt == 4 ? 1 : 2;
        if (dev >= 0 && dev < 64 && !configured[slot][dev]) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(LIMIT));
            configured[slot][dev] = true;
        }
        kern<<<grid, RT_ROWS, smem, s>>>(w, x, c, O, T, K);
    };
    if (pitch == 144) {
        if (tt == 1) go(refmode_tile_q4k_kernel<1, 144>);
        else if (tt == 4) go(refmode_tile_q4k_kernel<4, 144>);
        else go(refmode_tile_q4k_kernel<8, 144>);
    } else {
        if (tt == 1) go(refmode_tile_q4k_kernel<1, 176>);
        else if (tt == 4) go(refmode_tile_q4k_kernel<4, 176>);
        else go(refmode_tile_q4k_kernel<8, 176>);
    }
    return true;
}

template <int FMT>
static void launch_fast(const uint8_t* w, const uint8_t* x, __half* c, int64_t O, int64_t T, int64_t K, cudaStream_t s) {
    if (FMT == GGQ_Q4_K) {
        static const bool no_tile = getenv("GGQ_REFMODE_NOTILE") != nullptr;   // dev: the untiled vector kernel
        static const int mma_min_t = getenv("GGQ_REFMODE_MMA_MIN_T") ? atoi(getenv("GGQ_REFMODE_MMA_MIN_T")) : 3;   // T = 2 measured equal (96 vs 103 us on the lm_head); dev: A/B
        if (!no_tile && T >= mma_min_t && launch_mma_q4k(w, x, c, O, T, K, s)) return;
        if (!no_tile && launch_tile_q4k(w, x, c, O, T, K, s)) return;
    }
    auto grid = [&](int tt) { return static_cast<unsigned>((O * ((T + tt - 1) / tt) + 127) / 128); };
    if (T == 1) refmode_fast_kernel<FMT, 1><<<grid(1), 128, 0, s>>>(w, x, c, O, T, K);
    else if (T <= 4) refmode_fast_kernel<FMT, 4><<<grid(4), 128, 0, s>>>(w, x, c, O, T, K);
    else refmode_fast_kernel<FMT, 8><<<grid(8), 128, 0, s>>>(w, x, c, O, T, K);
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
    if ((n + 127) / 128 > 0x7fffffff) return GGQ_E_SHAPE;
    const unsigned grid = static_cast<unsigned>((n + 127) / 128);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint8_t* w = static_cast<const uint8_t*>(W);
    const uint8_t* x = static_cast<const uint8_t*>(XQ);
    __half* c = static_cast<__half*>(C);
    static const bool slow = getenv("GGQ_REFMODE_BYTEWISE") != nullptr;   // dev: always the byte-wise kernel
    const int64_t nb = K / fmt_qk(fmt);
    const bool vec_ok = !slow && (reinterpret_cast<uintptr_t>(W) & 15) == 0 && (reinterpret_cast<uintptr_t>(XQ) & 3) == 0 &&
                        (fmt == GGQ_Q4_K || nb % 2 == 0);
    if (vec_ok) {
        switch (fmt) {
            case GGQ_Q8_0: launch_fast<GGQ_Q8_0>(w, x, c, O, T, K, s); break;
            case GGQ_Q4_K: launch_fast<GGQ_Q4_K>(w, x, c, O, T, K, s); break;
            default: launch_fast<GGQ_Q6_K>(w, x, c, O, T, K, s); break;
        }
        count_launch();
        return static_cast<int>(cudaGetLastError());
    }
    switch (fmt) {
        case GGQ_Q8_0: refmode_kernel<GGQ_Q8_0><<<grid, 128, 0, s>>>(w, x, c, O, T, K); break;
        case GGQ_Q4_K: refmode_kernel<GGQ_Q4_K><<<grid, 128, 0, s>>>(w, x, c, O, T, K); break;
        default: refmode_kernel<GGQ_Q6_K><<<grid, 128, 0, s>>>(w, x, c, O, T, K); break;
    }
    count_launch();
    return static_cast<int>(cudaGetLastError());
}
