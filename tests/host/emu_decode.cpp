// emu_decode.cpp — runs the decode family's lane-level code (csrc/decode_tile.cuh) on the CPU under a
// 32-thread warp emulator and checks it against a scalar restatement of the block formats.
// Usage: emu_decode <fmt 0|1|2> <T 1..16> <K> <seed> [gemv 0|1]      exit 0 = pass.   Test infrastructure only.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "../../gguf-triton-kernel_b200/csrc/decode_tile.cuh"

using namespace ggq::dec;

static uint16_t f2h(float f) {  // round-to-nearest-even float -> half (finite inputs)
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    int32_t exp = static_cast<int32_t>((x >> 23) & 0xff) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
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
}

// ---- scalar format restatement, exact in double ----------------------
static double weight(int fmt, const uint8_t* row, int k) {
    if (fmt == 0) {
        const uint8_t* b = row + (k / 32) * 34;
        return static_cast<double>(h2f(b[0] | (b[1] << 8))) * static_cast<int8_t>(b[2 + k % 32]);
    }
    if (fmt == 1) {
        const uint8_t* b = row + (k / 256) * 144;
        const int e = k % 256, j = e / 32;
        const double d = h2f(b[0] | (b[1] << 8)), dmin = h2f(b[2] | (b[3] << 8));
        const uint8_t* s = b + 4;
        int sc, m;
        if (j < 4) { sc = s[j] & 63; m = s[j + 4] & 63; }
        else { sc = (s[j + 4] & 0xF) | ((s[j - 4] >> 6) << 4); m = (s[j + 4] >> 4) | ((s[j] >> 6) << 4); }
        const int byte = b[16 + (e / 64) * 32 + e % 32];
        const int q = (e & 32) ? (byte >> 4) : (byte & 15);
        return d * sc * q - dmin * m;
    }
    const uint8_t* b = row + (k / 256) * 210;
    const int e = k % 256, h = e / 128, r = e % 128, g = r / 32, l = r % 32;
    const int qlb = b[64 * h + (g & 1) * 32 + l];
    const int lo = (g & 2) ? (qlb >> 4) : (qlb & 15);
    const int hi = (b[128 + 32 * h + l] >> (2 * g)) & 3;
    const double d = h2f(b[208] | (b[209] << 8));
    return d * static_cast<int8_t>(b[192 + e / 16]) * ((lo | (hi << 4)) - 32);
}

template <int FMT, int NT, bool GV = false>
static int run(int T, int K, unsigned seed) {
    using G = Geo<FMT>;
    const int nb = K / G::QK, rowB = nb * G::BLK;
    std::mt19937 rng(seed);
    std::uniform_int_distribution<int> byte(0, 255);
    std::uniform_real_distribution<float> uni(0.25f, 1.0f);
    std::normal_distribution<float> nrm(0.f, 1.f);
    // 16 packed rows, random payload, sane fp16 scales
    std::vector<uint8_t> W(16 * static_cast<size_t>(rowB) + 64);
    for (auto& v : W) v = static_cast<uint8_t>(byte(rng));
    for (int r = 0; r < 16; ++r)
        for (int b = 0; b < nb; ++b) {
            uint8_t* blk = &W[static_cast<size_t>(r) * rowB + b * G::BLK];
            auto put = [&](int off, float v) { const uint16_t h = f2h(v); blk[off] = h & 0xff; blk[off + 1] = h >> 8; };
            const float sgn = (byte(rng) & 1) ? -1.f : 1.f;
            if (FMT == 0) put(0, sgn * uni(rng) * 0.02f);
            if (FMT == 1) { put(0, uni(rng) * 0.002f); put(2, uni(rng) * 0.002f); }
            if (FMT == 2) put(208, sgn * uni(rng) * 0.0005f);
        }
    // activations: T rows of fp16, row pitch K + 8 halves (any 16-byte multiple works)
    const int xpitch = (K + 8) * 2;
    alignas(16) static uint8_t Xbuf[16 * (8192 + 8) * 2 + 64];
    if (K > 8192) return 2;
    struct XV { uint8_t* p; uint8_t* data() { return p; } uint8_t& operator[](size_t i) { return p[i]; } } X{Xbuf};
    std::memset(Xbuf, 0, sizeof(Xbuf));
    std::vector<double> xd(static_cast<size_t>(T) * K);
    for (int t = 0; t < T; ++t)
        for (int k = 0; k < K; ++k) {
            const uint16_t h = f2h(nrm(rng));
            std::memcpy(&X[static_cast<size_t>(t) * xpitch + 2 * k], &h, 2);
            xd[static_cast<size_t>(t) * K + k] = h2f(h);
        }
    // per-slice activation table (and, for Q4_K, the in-place permutation) exactly as the kernel builds it
    const int tpad = 8 * NT, ngrp = K / G::GROUP;
    std::vector<float> tbl(static_cast<size_t>(ngrp) * tpad + 64, 0.f);
    stage_activations<FMT, NT, GV>(X.data(), static_cast<uint32_t>(xpitch), tbl.data(), K, T, 0, 1);
    // emulate the staging of every chunk and run the warp
    alignas(16) static uint8_t stage[16 * 1024];
    alignas(16) static uint8_t scratch[16 * 8 * 64 + 64];
    WarpEmu warp;
    std::vector<Acc<NT>> accs(32);
    for (auto& a : accs) std::memset(&a, 0, sizeof(a));
    const int nchunks = (nb + G::CHUNK_BLOCKS - 1) / G::CHUNK_BLOCKS;
    auto lane_fn = [&](int lane) {
        tls_warp = &warp;
        tls_lane = lane;
        Lane L{lane, lane >> 2, lane & 3};
        for (int c = 0; c < nchunks; ++c) {
            const int b0 = c * G::CHUNK_BLOCKS;
            const int nblk = std::min(G::CHUNK_BLOCKS, nb - b0);
            const int goff = b0 * G::BLK;       // byte offset inside a (16-byte aligned) row
            constexpr int SUBTILES = G::CHUNK_BLOCKS / G::PREP_BLOCKS;
            if (lane < 16) {                    // what the TMA boxes of a stage do (zero fill past the row end)
                for (int sub = 0; sub < SUBTILES; ++sub) {
                    const int src = (goff + sub * G::PREP_BLOCKS * G::BLK) & ~15;
                    uint8_t* dst = stage + sub * 16 * G::SLOT + lane * G::SLOT;
                    for (int i = 0; i < G::SLOT; ++i) dst[i] = (src + i < rowB) ? W[static_cast<size_t>(lane) * rowB + src + i] : 0;
                }
            }
            syncwarp();
            StageArgs s{};
            for (int nt = 0; nt < NT; ++nt) {
                s.xrow[nt] = &X[static_cast<size_t>(std::min(8 * nt + L.g, T - 1)) * xpitch];
                s.xv[nt] = 8 * nt + L.g < T;
            }
            s.tbl = tbl.data();
            s.scratch = scratch;
            for (int b = 0; b < nblk; b += G::PREP_BLOCKS) {  // same sub-stepping as decode.cu
                s.rows = stage + (b / G::PREP_BLOCKS) * 16 * G::SLOT;
                s.data_off = goff & 15;
                s.nblk = std::min(G::PREP_BLOCKS, nblk - b);
                s.k0 = (b0 + b) * G::QK;
                if (s.nblk == G::PREP_BLOCKS) {  // the kernel's fast path
                    Tile<FMT, NT, GV>::template prep<true>(L, s);
                    syncwarp();
                    Tile<FMT, NT, GV>::template compute<true>(L, s, accs[lane]);
                } else {
                    Tile<FMT, NT, GV>::template prep<false>(L, s);
                    syncwarp();
                    Tile<FMT, NT, GV>::template compute<false>(L, s, accs[lane]);
                }
                syncwarp();
            }
        }
        if constexpr (GV) gemv_finalize(accs[lane]);
    };
    std::vector<std::thread> th;
    for (int l = 0; l < 32; ++l) th.emplace_back(lane_fn, l);
    for (auto& t : th) t.join();
    // compare with the scalar restatement
    double num = 0, den = 0, worst = 0;
    for (int r = 0; r < 16; ++r)
        for (int t = 0; t < T; ++t) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += weight(FMT, &W[static_cast<size_t>(r) * rowB], k) * xd[static_cast<size_t>(t) * K + k];
            // C fragment: row r -> g = r & 7, regs 0/1 (r < 8) or 2/3; col t -> lane t/2 within quad of n-tile t/8
            const int nt = t / 8, col = t % 8, lane = (r & 7) * 4 + col / 2, reg = (r >= 8 ? 2 : 0) + (col & 1);
            const double got = accs[lane].v[nt][reg];
            num += (got - ref) * (got - ref);
            den += ref * ref;
            worst = std::max(worst, std::fabs(got - ref));
        }
    const double rel = std::sqrt(num / den);
    std::printf("fmt=%d NT=%d gemv=%d T=%d K=%d rel_fro=%.3e max_abs=%.3e\n", FMT, NT, int(GV), T, K, rel, worst);
    return rel < 2e-6 ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc < 5) return 2;
    const int fmt = std::atoi(argv[1]), T = std::atoi(argv[2]), K = std::atoi(argv[3]);
    const unsigned seed = static_cast<unsigned>(std::atoi(argv[4]));
    const bool two = T > 8;
    if (argc > 5 && std::atoi(argv[5]) == 1) {  // single-token GEMV tile code
        if (T != 1) return 2;
        switch (fmt) {
            case 0: return run<0, 1, true>(T, K, seed);
            case 1: return run<1, 1, true>(T, K, seed);
            case 2: return run<2, 1, true>(T, K, seed);
        }
        return 2;
    }
    switch (fmt) {
        case 0: return two ? run<0, 2>(T, K, seed) : run<0, 1>(T, K, seed);
        case 1: return two ? run<1, 2>(T, K, seed) : run<1, 1>(T, K, seed);
        case 2: return two ? run<2, 2>(T, K, seed) : run<2, 1>(T, K, seed);
    }
    return 2;
}
