// packk.cu — GPU packers for Q4_K and Q6_K, byte-identical to the reference's compiled packers (see kquant_pack.cuh
// for the arithmetic contract; this file must be compiled with -fmad=false).
// One thread per sub-block (32 weights for Q4_K, 16 for Q6_K) runs the scale search with its weights in registers;
// the 8 / 16 threads of a super-block sit in one warp, agree on the super-block maxima with shuffles and write their
// own slice of the block.  The search is ~10^4 dependent fp32 operations per thread, so the kernels are compute
// (latency) bound; the 128256 x 4096 lm_head packs in tens of milliseconds instead of the minute the single-threaded
// reference takes.
#include "../../include/ggq.h"
#include "common.cuh"
#include "kquant_pack.cuh"

namespace ggq {
namespace {

using namespace kq;

__device__ __forceinline__ uint32_t spread4(uint32_t nib16) {  // nibbles 0..3 of the low 16 bits -> bytes 0..3
    return (nib16 & 0xFu) | ((nib16 & 0xF0u) << 4) | ((nib16 & 0xF00u) << 8) | ((nib16 & 0xF000u) << 12);
}

// ---- Q4_K: 8 threads per super-block, 4 super-blocks per warp -----------------------------------------------------------
__global__ void __launch_bounds__(128) pack_q4_k_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int64_t nsub) {
    const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // sub-block index
    const bool live = gid < nsub;            // nsub is a multiple of 8: a group of 8 lanes is live or dead together
    const int lane = threadIdx.x & 31, j = lane & 7;
    float xs[32];
    if (live) {
        const float4* src = reinterpret_cast<const float4*>(x + gid * 32);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const float4 q = src[v];
            xs[4 * v] = q.x;
            xs[4 * v + 1] = q.y;
            xs[4 * v + 2] = q.z;
            xs[4 * v + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) xs[i] = 0.f;
    }
    Sub4 s = search_q4(xs);
    // super-block maxima (q4_k_ref.c:309-318: running maxima from 0, strict >)
    float max_scale = s.scale > 0.f ? s.scale : 0.f, max_min = s.minv > 0.f ? s.minv : 0.f;
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) {
        const float a = __shfl_xor_sync(0xffffffffu, max_scale, m), b = __shfl_xor_sync(0xffffffffu, max_min, m);
        max_scale = a > max_scale ? a : max_scale;
        max_min = b > max_min ? b : max_min;
    }
    const uint16_t d_bits = f2h_bits(max_scale / 63.f), dmin_bits = f2h_bits(max_min / 63.f);
    const Code4 c = code_q4(s, max_scale, max_min);
    requant_q4(xs, s, c, d_bits, dmin_bits);
    // gather the eight (ls, lm) codes in every lane of the group; lane j == 0 writes the 16-byte header
    uint8_t ls[8], lm[8];
    const uint32_t mine = c.ls | (static_cast<uint32_t>(c.lm) << 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t v = __shfl_sync(0xffffffffu, mine, (lane & ~7) + k);
        ls[k] = static_cast<uint8_t>(v & 0xff);
        lm[k] = static_cast<uint8_t>(v >> 8);
    }
    // the partner's quants: byte l of pair p is nibble l of sub-block 2p (low) and of sub-block 2p + 1 (high)
    uint32_t pq[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) pq[k] = __shfl_xor_sync(0xffffffffu, s.q[k], 1);
    if (!live) return;
    uint8_t* blk = out + (gid >> 3) * 144;
    if (j == 0) {
        uint8_t sb[12];
        scale_bytes_q4(ls, lm, sb);
        uint4 h;
        h.x = d_bits | (static_cast<uint32_t>(dmin_bits) << 16);
        h.y = sb[0] | (sb[1] << 8) | (sb[2] << 16) | (static_cast<uint32_t>(sb[3]) << 24);
        h.z = sb[4] | (sb[5] << 8) | (sb[6] << 16) | (static_cast<uint32_t>(sb[7]) << 24);
        h.w = sb[8] | (sb[9] << 8) | (sb[10] << 16) | (static_cast<uint32_t>(sb[11]) << 24);
        *reinterpret_cast<uint4*>(blk) = h;
    }
    if ((j & 1) == 0) {
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o[2 * k] = spread4(s.q[k] & 0xFFFFu) | (spread4(pq[k] & 0xFFFFu) << 4);
            o[2 * k + 1] = spread4(s.q[k] >> 16) | (spread4(pq[k] >> 16) << 4);
        }
        uint4* dst = reinterpret_cast<uint4*>(blk + 16 + 32 * (j >> 1));
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

// ---- Q6_K: 16 threads per super-block, 2 super-blocks per warp -------------------------------------------------------
__global__ void __launch_bounds__(128) pack_q6_k_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int64_t nsub) {
    const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool live = gid < nsub;            // nsub is a multiple of 16
    const int lane = threadIdx.x & 31, j = lane & 15;
    float xs[16];
    if (live) {
        const float4* src = reinterpret_cast<const float4*>(x + gid * 16);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const float4 q = src[v];
            xs[4 * v] = q.x;
            xs[4 * v + 1] = q.y;
            xs[4 * v + 2] = q.z;
            xs[4 * v + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) xs[i] = 0.f;
    }
    Sub6 s = search_q6(xs);
    // the FIRST sub-block with the largest |scale| gives the signed super-block scale (q6_k_ref.c:262-270)
    float max_scale = 0.f, max_abs = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float sc = __shfl_sync(0xffffffffu, s.scale, (lane & ~15) + k);
        const float a = fabsf(sc);
        if (a > max_abs) {
            max_abs = a;
            max_scale = sc;
        }
    }
    const bool zero = max_abs < GROUP_EPS;
    int8_t code = 0;
    uint16_t d_bits = 0;
    if (!zero) {
        const float iscale = -128.f / max_scale;
        d_bits = f2h_bits(1 / iscale);
        code = code_q6(s.scale, iscale);
        requant_q6(xs, s, code, d_bits);
    } else {
        s.q[0] = s.q[1] = s.q[2] = s.q[3] = 0u;   // an all-zero block is 210 zero bytes
    }
    // sub-block j = 8h + 2g + lp (h: half, g: 32-element group, lp: low / high 16 of the group).  The lane with g == 0
    // collects the quants of g = 1, 2, 3 (lanes j + 2, j + 4, j + 6) and writes 16 bytes of ql[0:32], ql[32:64], qh.
    uint32_t q1[4], q2[4], q3[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        q1[k] = __shfl_down_sync(0xffffffffu, s.q[k], 2);
        q2[k] = __shfl_down_sync(0xffffffffu, s.q[k], 4);
        q3[k] = __shfl_down_sync(0xffffffffu, s.q[k], 6);
    }
    if (!live) return;
    uint8_t* blk = out + (gid >> 4) * 210;     // 2-byte aligned
    blk[192 + j] = static_cast<uint8_t>(code);
    if (j == 0) *reinterpret_cast<uint16_t*>(blk + 208) = d_bits;
    if ((j & 6) == 0) {   // g == 0: j = 8h + lp
        const int h = j >> 3, lp = j & 1;
        uint16_t* qla = reinterpret_cast<uint16_t*>(blk + 64 * h + 16 * lp);
        uint16_t* qlb = reinterpret_cast<uint16_t*>(blk + 64 * h + 32 + 16 * lp);
        uint16_t* qh = reinterpret_cast<uint16_t*>(blk + 128 + 32 * h + 16 * lp);
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // four quants (bytes) per word
            const uint32_t L0 = s.q[k], L1 = q1[k], L2 = q2[k], L3 = q3[k];
            const uint32_t a = (L0 & 0x0F0F0F0Fu) | ((L2 & 0x0F0F0F0Fu) << 4);
            const uint32_t b = (L1 & 0x0F0F0F0Fu) | ((L3 & 0x0F0F0F0Fu) << 4);
            const uint32_t hh = ((L0 >> 4) & 0x03030303u) | (((L1 >> 4) & 0x03030303u) << 2) | (((L2 >> 4) & 0x03030303u) << 4) |
                                (((L3 >> 4) & 0x03030303u) << 6);
            qla[2 * k] = static_cast<uint16_t>(a);
            qla[2 * k + 1] = static_cast<uint16_t>(a >> 16);
            qlb[2 * k] = static_cast<uint16_t>(b);
            qlb[2 * k + 1] = static_cast<uint16_t>(b >> 16);
            qh[2 * k] = static_cast<uint16_t>(hh);
            qh[2 * k + 1] = static_cast<uint16_t>(hh >> 16);
        }
    }
}

template <typename K>
int launch_pack(K kern, const void* x, void* out, int64_t n, int sub, void* stream) {
    if (n < 0 || n % 256 != 0) return GGQ_E_SHAPE;   // the reference raises ValueError (utils/quantize/q4_k.py:72-73)
    if (n == 0) return 0;
    if (!x || !out) return GGQ_E_POINTER;
    if (reinterpret_cast<uintptr_t>(x) & 15) return GGQ_E_POINTER;   // float4 loads
    const int64_t nsub = n / sub;
    const int64_t blocks = (nsub + 127) / 128;
    if (blocks > 0x7fffffff) return GGQ_E_SHAPE;
    kern<<<static_cast<unsigned>(blocks), 128, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const float*>(x),
                                                                                    static_cast<uint8_t*>(out), nsub);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

}  // namespace
}  // namespace ggq

extern "C" int ggq_quantize_q4_k_f32(const void* x, void* out, int64_t n, void* stream) {
    return ggq::launch_pack(ggq::pack_q4_k_kernel, x, out, n, 32, stream);
}
extern "C" int ggq_quantize_q6_k_f32(const void* x, void* out, int64_t n, void* stream) {
    return ggq::launch_pack(ggq::pack_q6_k_kernel, x, out, n, 16, stream);
}
