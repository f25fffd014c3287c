// cuda_shim.h — host stand-ins for the few CUDA built-ins used by csrc/decode_tile.cuh, plus a
// 32-thread warp emulator for mma.sync.m16n8k16.  Test infrastructure only (tests/host/).
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };

namespace ggq {
namespace dec {

inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    const uint64_t v = (static_cast<uint64_t>(b) << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t n = (sel >> (4 * i)) & 0xF;
        uint32_t byte = static_cast<uint32_t>((v >> (8 * (n & 7))) & 0xFF);
        if (n & 8) byte = (byte & 0x80) ? 0xFF : 0x00;  // sign replicate mode (unused here)
        r |= byte << (8 * i);
    }
    return r;
}
inline uint32_t funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    const uint64_t v = (static_cast<uint64_t>(hi) << 32) | lo;
    return static_cast<uint32_t>(v >> (sh & 31));
}
inline float h2f(uint32_t bits) {
    const uint32_t h = bits & 0xffffu, sign = (h >> 15) & 1, exp = (h >> 10) & 0x1f, man = h & 0x3ff;
    float v;
    if (exp == 0) v = std::ldexp(static_cast<float>(man), -24);
    else if (exp == 31) v = man ? NAN : INFINITY;
    else v = std::ldexp(static_cast<float>(man | 0x400), static_cast<int>(exp) - 25);
    return sign ? -v : v;
}

inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

struct WarpEmu {
    std::barrier<> bar{32};
    uint32_t a[32][4];
    uint32_t b[32][2];
    float x[32];
};
inline thread_local WarpEmu* tls_warp = nullptr;
inline thread_local int tls_lane = 0;

inline float frag_half(uint32_t reg, int hi) { return h2f(hi ? (reg >> 16) : (reg & 0xffff)); }

// PTX ISA fragment layouts of mma.m16n8k16 (.f16): see the comments in decode_tile.cuh
inline void mma16816(float d[4], const uint32_t a[4], const uint32_t b[2], const float c[4]) {
    WarpEmu& w = *tls_warp;
    const int lane = tls_lane, g = lane >> 2, t = lane & 3;
    std::memcpy(w.a[lane], a, 16);
    std::memcpy(w.b[lane], b, 8);
    const float cin[4] = {c[0], c[1], c[2], c[3]};
    w.bar.arrive_and_wait();
    for (int i = 0; i < 4; ++i) {
        const int row = g + ((i & 2) ? 8 : 0), col = 2 * t + (i & 1);
        float acc = cin[i];
        for (int k = 0; k < 16; ++k) {
            const int ta = (k & 7) >> 1;
            const float av = frag_half(w.a[(row & 7) * 4 + ta][(row >= 8 ? 1 : 0) + (k >= 8 ? 2 : 0)], k & 1);
            const float bv = frag_half(w.b[col * 4 + ta][k >= 8 ? 1 : 0], k & 1);
            acc += av * bv;  // products of fp16 values are exact in fp32
        }
        d[i] = acc;
    }
    w.bar.arrive_and_wait();
}
inline float frag_bf16(uint32_t reg, int hi) { return u2f((hi ? (reg >> 16) : (reg & 0xffff)) << 16); }
inline void mma16816_bf16(float d[4], const uint32_t a[4], const uint32_t b[2], const float c[4]) {
    WarpEmu& w = *tls_warp;
    const int lane = tls_lane, g = lane >> 2, t = lane & 3;
    std::memcpy(w.a[lane], a, 16);
    std::memcpy(w.b[lane], b, 8);
    const float cin[4] = {c[0], c[1], c[2], c[3]};
    w.bar.arrive_and_wait();
    for (int i = 0; i < 4; ++i) {
        const int row = g + ((i & 2) ? 8 : 0), col = 2 * t + (i & 1);
        float acc = cin[i];
        for (int k = 0; k < 16; ++k) {
            const int ta = (k & 7) >> 1;
            acc += frag_bf16(w.a[(row & 7) * 4 + ta][(row >= 8 ? 1 : 0) + (k >= 8 ? 2 : 0)], k & 1) *
                   frag_bf16(w.b[col * 4 + ta][k >= 8 ? 1 : 0], k & 1);
        }
        d[i] = acc;
    }
    w.bar.arrive_and_wait();
}
inline void syncwarp() { tls_warp->bar.arrive_and_wait(); }
inline float shfl_xor(float v, int m) {
    WarpEmu& w = *tls_warp;
    w.x[tls_lane] = v;
    w.bar.arrive_and_wait();
    const float r = w.x[tls_lane ^ m];
    w.bar.arrive_and_wait();
    return r;
}

}  // namespace dec
}  // namespace ggq
