// kquant_host.cpp — host build of csrc/kquant_pack.cuh (the sub-block searches and super-block steps the GPU packers
// run one thread per sub-block) with the cross-thread steps written as plain loops.  Test infrastructure: compiled by
// tests/test_kquant_pack_cpu.py with `g++ -O2 -ffp-contract=off` and compared byte for byte with the reference's
// compiled packers (oracle/_ref).
#include "../../gguf-triton-kernel_b200/csrc/kquant_pack.cuh"

using namespace ggq::kq;

extern "C" void host_quantize_q4_k(const float* x, uint8_t* out, long long n) {
    for (long long b = 0; b < n / 256; ++b, x += 256, out += 144) {
        Sub4 s[8];
        float xs[8][32];
        float max_scale = 0.f, max_min = 0.f;
        for (int j = 0; j < 8; ++j) {
            for (int i = 0; i < 32; ++i) xs[j][i] = x[32 * j + i];
            s[j] = search_q4(xs[j]);
            if (s[j].scale > max_scale) max_scale = s[j].scale;
            if (s[j].minv > max_min) max_min = s[j].minv;
        }
        const uint16_t d_bits = f2h_bits(max_scale / 63.f), dmin_bits = f2h_bits(max_min / 63.f);
        uint8_t ls[8], lm[8];
        for (int j = 0; j < 8; ++j) {
            const Code4 c = code_q4(s[j], max_scale, max_min);
            ls[j] = c.ls;
            lm[j] = c.lm;
            requant_q4(xs[j], s[j], c, d_bits, dmin_bits);
        }
        uint8_t sb[12];
        scale_bytes_q4(ls, lm, sb);
        out[0] = d_bits & 0xff;
        out[1] = d_bits >> 8;
        out[2] = dmin_bits & 0xff;
        out[3] = dmin_bits >> 8;
        for (int j = 0; j < 12; ++j) out[4 + j] = sb[j];
        for (int p = 0; p < 4; ++p)
            for (int l = 0; l < 32; ++l) out[16 + 32 * p + l] = static_cast<uint8_t>(get4(s[2 * p].q, l) | (get4(s[2 * p + 1].q, l) << 4));
    }
}

extern "C" void host_quantize_q6_k(const float* x, uint8_t* out, long long n) {
    for (long long b = 0; b < n / 256; ++b, x += 256, out += 210) {
        Sub6 s[16];
        float xs[16][16];
        float max_scale = 0.f, max_abs = 0.f;
        for (int j = 0; j < 16; ++j) {
            for (int i = 0; i < 16; ++i) xs[j][i] = x[16 * j + i];
            s[j] = search_q6(xs[j]);
            const float a = fabsf(s[j].scale);
            if (a > max_abs) {
                max_abs = a;
                max_scale = s[j].scale;
            }
        }
        if (max_abs < GROUP_EPS) {
            for (int i = 0; i < 210; ++i) out[i] = 0;
            continue;
        }
        const float iscale = -128.f / max_scale;
        const uint16_t d_bits = f2h_bits(1 / iscale);
        for (int j = 0; j < 16; ++j) {
            const int8_t c = code_q6(s[j].scale, iscale);
            out[192 + j] = static_cast<uint8_t>(c);
            requant_q6(xs[j], s[j], c, d_bits);
        }
        out[208] = d_bits & 0xff;
        out[209] = d_bits >> 8;
        for (int h = 0; h < 2; ++h)
            for (int l = 0; l < 32; ++l) {
                const int L0 = get8(s[8 * h + 0 + l / 16].q, l % 16), L1 = get8(s[8 * h + 2 + l / 16].q, l % 16);
                const int L2 = get8(s[8 * h + 4 + l / 16].q, l % 16), L3 = get8(s[8 * h + 6 + l / 16].q, l % 16);
                out[64 * h + l] = static_cast<uint8_t>((L0 & 0xF) | ((L2 & 0xF) << 4));
                out[64 * h + 32 + l] = static_cast<uint8_t>((L1 & 0xF) | ((L3 & 0xF) << 4));
                out[128 + 32 * h + l] = static_cast<uint8_t>((L0 >> 4) | ((L1 >> 4) << 2) | ((L2 >> 4) << 4) | ((L3 >> 4) << 6));
            }
    }
}
