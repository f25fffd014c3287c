// api.cu — the extern "C" boundary of libggq.so (include/ggq.h): validation + family dispatch.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>

#include "../../include/ggq.h"
#include "common.cuh"
#include "formats.cuh"

namespace ggq {

static std::atomic<int64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cache[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

static int validate(int fmt, const void* W, const void* X, void* const* C_out, int n_out, int64_t ldx, int64_t ldc,
                    int64_t O, int64_t T, int64_t K) {
    if (fmt < GGQ_Q8_0 || fmt > GGQ_Q6_K) return GGQ_E_FORMAT;
    if (O < 0 || T < 0 || K < 0 || K % fmt_qk(fmt) != 0) return GGQ_E_SHAPE;
    if (n_out < 1 || n_out > 8 || ldx < K || ldc < O) return GGQ_E_SHAPE;
    if (O > (int64_t{1} << 31) || T > (int64_t{1} << 24) || K > (int64_t{1} << 24)) return GGQ_E_SHAPE;
    if (O == 0 || T == 0) return 0;
    if (!C_out) return GGQ_E_POINTER;
    for (int i = 0; i < n_out; ++i)
        if (!C_out[i]) return GGQ_E_POINTER;
    if (K > 0 && (!W || !X)) return GGQ_E_POINTER;
    return 0;
}

static int select_family(int fmt, const MmArgs& a) {
    if (a.O < 16) return GGQ_FAMILY_GENERIC;  // less than one 16-row MMA tile: one warp per row is the better fit
    // skinny (tcgen05, weights as the TMEM A operand): every 17 <= T <= 127, and the T <= 16 shapes where it measured
    // faster than the mma.sync decode kernel on B200 (profiles/README.md): its cost per weight does not grow with T, the
    // decode kernel's does (one more n-tile from T = 9, shared-memory re-reads of the activations from T = 5).  Small
    // layers stay with the decode kernel (lower fixed cost per launch).
    static const bool no_skinny = [] { const char* e = getenv("GGQ_SKINNY"); return e && e[0] == '0'; }();
    if (a.T >= 2 && a.T <= 127 && !no_skinny && skinny_supports(fmt, a)) {
        static const int thr[3] = {6, 9, 9};    // Q8_0, Q4_K, Q6_K: smallest T that goes to the skinny kernel
        const bool big = a.O * a.K >= (int64_t{48} << 20);
        if (a.T > 16 || (big && a.T >= thr[fmt]) || !decode_supports(fmt, a)) return GGQ_FAMILY_SKINNY;
    }
    if (a.T <= 16 && decode_supports(fmt, a)) return GGQ_FAMILY_DECODE;
    if (a.T >= 64 && prefill_supports(fmt, a)) return GGQ_FAMILY_PREFILL;
    if (a.T > 16 && decode_supports(fmt, a)) return GGQ_FAMILY_DECODE;  // looped over 16-token groups
    return GGQ_FAMILY_GENERIC;
}

// K == 0: the product is an all-zero [T, O]; written by a tiny kernel so the call stays asynchronous.
__global__ void zero_out_kernel(OutPtrs outs, int64_t ldc, int64_t O, int64_t T) {
    const int64_t n = O * T;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t t = i / O, o = i % O;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < outs.n) outs.p[j][t * ldc + o] = __float2half_rn(0.f);
    }
}

static int mm(int fmt, const void* W, const void* X, int64_t ldx, void* const* C_out, int n_out, int64_t ldc, int64_t O,
              int64_t T, int64_t K, int family, void* stream) {
    const int v = validate(fmt, W, X, C_out, n_out, ldx, ldc, O, T, K);
    if (v != 0) return v;
    if (O == 0 || T == 0) return 0;
    MmArgs a{};
    a.W = static_cast<const uint8_t*>(W);
    a.X = X;
    for (int i = 0; i < n_out; ++i) a.C[i] = C_out[i];
    a.n_out = n_out;
    a.ldx = ldx;
    a.ldc = ldc;
    a.O = O;
    a.T = T;
    a.K = K;
    a.stream = static_cast<cudaStream_t>(stream);
    if (K == 0) {
        zero_out_kernel<<<64, 256, 0, a.stream>>>(make_outs(a), ldc, O, T);
        count_launch();
        return static_cast<int>(cudaGetLastError());
    }
    if (family == GGQ_FAMILY_AUTO) family = select_family(fmt, a);
    switch (family) {
        case GGQ_FAMILY_GENERIC: return launch_generic(fmt, a);
        case GGQ_FAMILY_DECODE: return decode_supports(fmt, a) ? launch_decode(fmt, a) : GGQ_E_FAMILY;
        case GGQ_FAMILY_PREFILL: return prefill_supports(fmt, a) ? launch_prefill(fmt, a) : GGQ_E_FAMILY;
        case GGQ_FAMILY_SKINNY: return skinny_supports(fmt, a) ? launch_skinny(fmt, a) : GGQ_E_FAMILY;
    }
    return GGQ_E_FAMILY;
}

// C = silu(G) * C elementwise over [T, O] fp16 (the composed form of ggq_mm_swiglu: G = gate projection, C = up projection)
__global__ void silu_mul_kernel(const __half* __restrict__ G, __half* C, int64_t n) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float g = __half2float(G[i]), u = __half2float(C[i]);
        C[i] = __float2half_rn(g / (1.f + __expf(-g)) * u);
    }
}

static MmArgs placeholder_args(int64_t O, int64_t T, int64_t K) {
    MmArgs a{};
    a.W = reinterpret_cast<const uint8_t*>(uintptr_t{256});  // alignment-neutral placeholders (host-side planning only)
    a.X = reinterpret_cast<const void*>(uintptr_t{256});
    a.n_out = 1;
    a.ldx = K;
    a.ldc = O;
    a.O = O;
    a.T = T;
    a.K = K;
    return a;
}

static int dequant(int fmt, const void* W, void* out, int64_t O, int64_t K, void* stream) {
    if (O < 0 || K < 0 || K % fmt_qk(fmt) != 0) return GGQ_E_SHAPE;
    if (K > (int64_t{1} << 30)) return GGQ_E_SHAPE;   // the kernels index a row with 32-bit arithmetic
    if (O == 0 || K == 0) return 0;
    if (!W || !out) return GGQ_E_POINTER;
    return launch_dequant(fmt, static_cast<const uint8_t*>(W), out, O, K, static_cast<cudaStream_t>(stream));
}

}  // namespace ggq

using namespace ggq;

extern "C" {

int ggq_mm_q8_0_f16(const void* W, const void* X, void* C, int64_t O, int64_t T, int64_t K, void* stream) {
    void* outs[1] = {C};
    return mm(GGQ_Q8_0, W, X, K, outs, 1, O, O, T, K, GGQ_FAMILY_AUTO, stream);
}
int ggq_mm_q4_k_f16(const void* W, const void* X, void* C, int64_t O, int64_t T, int64_t K, void* stream) {
    void* outs[1] = {C};
    return mm(GGQ_Q4_K, W, X, K, outs, 1, O, O, T, K, GGQ_FAMILY_AUTO, stream);
}
int ggq_mm_q6_k_f16(const void* W, const void* X, void* C, int64_t O, int64_t T, int64_t K, void* stream) {
    void* outs[1] = {C};
    return mm(GGQ_Q6_K, W, X, K, outs, 1, O, O, T, K, GGQ_FAMILY_AUTO, stream);
}

int ggq_mm_ex(int fmt, const void* W, const void* X, int64_t ldx, void* const* C_out, int n_out, int64_t ldc, int64_t O,
              int64_t T, int64_t K, int family, void* stream) {
    if (family < GGQ_FAMILY_AUTO || family > GGQ_FAMILY_SKINNY) return GGQ_E_FAMILY;
    return mm(fmt, W, X, ldx, C_out, n_out, ldc, O, T, K, family, stream);
}

int64_t ggq_mm_swiglu_workspace(int fmt, int64_t O, int64_t T, int64_t K) {
    if (fmt < GGQ_Q8_0 || fmt > GGQ_Q6_K) return GGQ_E_FORMAT;
    if (O < 0 || T < 0 || K < 0 || K % fmt_qk(fmt) != 0) return GGQ_E_SHAPE;
    if (O == 0 || T == 0 || K == 0) return 0;
    MmArgs a = placeholder_args(O, T, K);
    a.W2 = a.W;
    return launch_decode_dual(fmt, a, true) == 0 ? 0 : T * O * 2;
}

int ggq_mm_swiglu(int fmt, const void* Wg, const void* Wu, const void* X, void* C, int64_t O, int64_t T, int64_t K,
                  void* workspace, int64_t workspace_bytes, void* stream) {
    void* outs[1] = {C};
    int v = validate(fmt, Wg, X, outs, 1, K, O, O, T, K);
    if (v != 0) return v;
    if (O == 0 || T == 0) return 0;
    if (K > 0 && !Wu) return GGQ_E_POINTER;
    if (K == 0) return mm(fmt, Wg, X, K, outs, 1, O, O, T, K, GGQ_FAMILY_AUTO, stream);  // silu(0) * 0 = 0
    MmArgs a = placeholder_args(O, T, K);
    a.W = static_cast<const uint8_t*>(Wg);
    a.W2 = static_cast<const uint8_t*>(Wu);
    a.X = X;
    a.C[0] = C;
    a.stream = static_cast<cudaStream_t>(stream);
    static const bool no_fused = getenv("GGQ_SWIGLU_COMPOSED") != nullptr;   // dev: always take the composed form
    if (!no_fused) {
        v = launch_decode_dual(fmt, a, false);
        if (v != GGQ_E_FAMILY) return v;
    }
    // composed: gate projection into the workspace, up projection into C, one elementwise pass
    if (!workspace) return GGQ_E_POINTER;
    if (workspace_bytes < T * O * 2) return GGQ_E_SHAPE;
    void* ws[1] = {workspace};
    v = mm(fmt, Wg, X, K, ws, 1, O, O, T, K, GGQ_FAMILY_AUTO, stream);
    if (v != 0) return v;
    v = mm(fmt, Wu, X, K, outs, 1, O, O, T, K, GGQ_FAMILY_AUTO, stream);
    if (v != 0) return v;
    const int64_t n = T * O;
    const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(num_sms()) * 8));
    silu_mul_kernel<<<blocks, 256, 0, a.stream>>>(static_cast<const __half*>(workspace), static_cast<__half*>(C), n);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

int ggq_mm_sync(int fmt, const void* W, const void* X, int64_t ldx, void* C, int64_t ldc, int64_t O, int64_t T, int64_t K,
                const ggq_peer_sync* sync, int* ctas_out, void* stream) {
    if (!sync || !ctas_out || sync->world < 2 || sync->world > 8 || sync->rank < 0 || sync->rank >= sync->world ||
        sync->x_owner < 0 || sync->x_owner >= sync->world || !sync->counter || !sync->c_land || !C)
        return GGQ_E_POINTER;
    const bool owner = sync->rank == sync->x_owner;
    void* outs[1] = {static_cast<__half*>(C) + sync->rank * O};   // this rank's own columns
    // the non-owners never dereference X (their activations arrive in x_land): validate() only needs a non-null pointer
    const int v = validate(fmt, W, owner ? X : W, outs, 1, ldx, ldc, O, T, K);
    if (v != 0) return v;
    if (O < 16 || T < 1 || T > 8 || K < fmt_qk(fmt)) return GGQ_E_FAMILY;
    if (O % 4 != 0 || K % 4 != 0 || T * K > 65536 || ldc < sync->world * O || ldc % 4 != 0) return GGQ_E_SHAPE;
    if ((reinterpret_cast<uintptr_t>(C) & 7) != 0) return GGQ_E_POINTER;
    if (sync->c_land_half < sync->world * T * O * 4 || sync->c_land_half % 16 != 0) return GGQ_E_SHAPE;
    if (sync->x_land_half < T * K * 4 || sync->x_land_half % 16 != 0) return GGQ_E_SHAPE;
    for (int i = 0; i < sync->world; ++i) {
        if (i == sync->rank) continue;
        if (!sync->c_land_peer[i]) return GGQ_E_POINTER;
        if (owner && !sync->x_land_peer[i]) return GGQ_E_POINTER;
    }
    if (!owner && !sync->x_land) return GGQ_E_POINTER;
    MmArgs a{};
    a.W = static_cast<const uint8_t*>(W);
    a.X = owner ? X : W;
    a.C[0] = outs[0];
    a.n_out = 1;
    a.ldx = ldx;
    a.ldc = ldc;
    a.O = O;
    a.T = T;
    a.K = K;
    a.stream = static_cast<cudaStream_t>(stream);
    PeerSync ps{};
    ps.rank = sync->rank;
    ps.world = sync->world;
    ps.x_owner = sync->x_owner;
    ps.epoch = sync->epoch;
    ps.counter = sync->counter;
    ps.epoch_dev = sync->epoch_dev;
    ps.C_full = static_cast<__half*>(C);
    if (sync->epoch_dev) {  // replayable mode: kernel-maintained epoch, odd epochs use the alternate buffers
        if ((owner && !sync->X_alt) || !sync->C_alt || (reinterpret_cast<uintptr_t>(sync->C_alt) & 7) != 0) return GGQ_E_POINTER;
        ps.X_alt = static_cast<const uint8_t*>(owner ? sync->X_alt : W);
        ps.C_alt = static_cast<__half*>(sync->C_alt);
    }
    ps.x_land = static_cast<uint4*>(sync->x_land);
    ps.c_land = static_cast<uint4*>(sync->c_land);
    for (int i = 0; i < 8; ++i) {
        const bool peer = i < sync->world && i != sync->rank;
        ps.x_land_peer[i] = (peer && owner) ? static_cast<uint4*>(sync->x_land_peer[i]) : nullptr;
        ps.c_land_peer[i] = peer ? static_cast<uint4*>(sync->c_land_peer[i]) : nullptr;
    }
    ps.x_half_lines = static_cast<uint32_t>(sync->x_land_half / 16);
    ps.c_half_lines = static_cast<uint32_t>(sync->c_land_half / 16);
    ps.status = sync->status;
    ps.timeout_ns = sync->timeout_ns ? sync->timeout_ns : 2000000000ull;
    a.sync = &ps;
    a.ctas_out = ctas_out;
    if (!decode_supports(fmt, a)) return GGQ_E_FAMILY;
    return launch_decode(fmt, a);
}

int ggq_dequant_q8_0_f16(const void* W, void* out, int64_t O, int64_t K, void* stream) {
    return dequant(GGQ_Q8_0, W, out, O, K, stream);
}
int ggq_dequant_q4_k_f16(const void* W, void* out, int64_t O, int64_t K, void* stream) {
    return dequant(GGQ_Q4_K, W, out, O, K, stream);
}
int ggq_dequant_q6_k_f16(const void* W, void* out, int64_t O, int64_t K, void* stream) {
    return dequant(GGQ_Q6_K, W, out, O, K, stream);
}

int64_t ggq_packed_nbytes(int fmt, int64_t O, int64_t K) {
    if (fmt < GGQ_Q8_0 || fmt > GGQ_Q6_K) return GGQ_E_FORMAT;
    if (O < 0 || K < 0 || K % fmt_qk(fmt) != 0) return GGQ_E_SHAPE;
    return O * (K / fmt_qk(fmt)) * fmt_blk(fmt);
}

int ggq_select_family(int fmt, int64_t O, int64_t T, int64_t K) {
    if (fmt < GGQ_Q8_0 || fmt > GGQ_Q6_K) return GGQ_E_FORMAT;
    if (O < 0 || T < 0 || K < 0 || K % fmt_qk(fmt) != 0) return GGQ_E_SHAPE;
    MmArgs a{};
    a.W = reinterpret_cast<const uint8_t*>(uintptr_t{256});  // alignment-neutral placeholder
    a.X = reinterpret_cast<const void*>(uintptr_t{256});
    a.n_out = 1;
    a.ldx = K;
    a.ldc = O;
    a.O = O;
    a.T = T;
    a.K = K;
    return select_family(fmt, a);
}

int ggq_decode_plan(int fmt, int64_t O, int64_t T, int64_t K, int* out9) {
    if (fmt < GGQ_Q8_0 || fmt > GGQ_Q6_K) return GGQ_E_FORMAT;
    if (O < 1 || T < 1 || K < fmt_qk(fmt) || K % fmt_qk(fmt) != 0 || !out9) return GGQ_E_SHAPE;
    MmArgs a{};
    a.W = reinterpret_cast<const uint8_t*>(uintptr_t{256});
    a.X = reinterpret_cast<const void*>(uintptr_t{256});
    a.n_out = 1;
    a.ldx = K;
    a.ldc = O;
    a.O = O;
    a.T = T;
    a.K = K;
    if (!decode_supports(fmt, a)) return GGQ_E_FAMILY;
    return decode_plan(fmt, a, out9);
}

int ggq_describe(int fmt, int64_t O, int64_t T, int64_t K, char* out, int cap) {
    if (fmt < GGQ_Q8_0 || fmt > GGQ_Q6_K) return GGQ_E_FORMAT;
    if (O < 1 || T < 1 || K < fmt_qk(fmt) || K % fmt_qk(fmt) != 0) return GGQ_E_SHAPE;
    if (!out || cap < 1) return GGQ_E_POINTER;
    MmArgs a{};
    a.W = reinterpret_cast<const uint8_t*>(uintptr_t{256});
    a.X = reinterpret_cast<const void*>(uintptr_t{256});
    a.n_out = 1;
    a.ldx = K;
    a.ldc = O;
    a.O = O;
    a.T = T;
    a.K = K;
    static const char* const names[3] = {"Q8_0", "Q4_K", "Q6_K"};
    switch (select_family(fmt, a)) {
        case GGQ_FAMILY_GENERIC:
            snprintf(out, cap, "ggq::gen::generic_kernel<%s> (one warp per output row)", names[fmt]);
            return 0;
        case GGQ_FAMILY_SKINNY: skinny_describe(fmt, a, out, cap); return 0;
        case GGQ_FAMILY_PREFILL:
            snprintf(out, cap, "ggq::pre::two::prefill2_kernel<%s> (tcgen05.mma cta_group::2, 512 x 256 tiles)", names[fmt]);
            return 0;
        case GGQ_FAMILY_DECODE: {
            int v[9];
            bool wide = false;
            if (decode_plan(fmt, a, v, &wide) != 0) return GGQ_E_FAMILY;
            const int tt = static_cast<int>(T > 16 ? 16 : T);
            const bool gv = tt == 1 && v[1] == 1 && v[3] == 1;
            snprintf(out, cap, "ggq::dec::decode_kernel<%s,NT=%d,AT=%d,%s%s> grid=%d occ=%d stages=%d k-slices=%d%s", names[fmt], v[2],
                     v[1], gv ? "GV=1 (single-token GEMV tile code)" : "GV=0", wide ? ",WIDE (4-block stages)" : "", v[6] / 100,
                     v[6] % 100, v[5], v[3], T > 16 ? " (16-token passes)" : "");
            return 0;
        }
    }
    return GGQ_E_FAMILY;
}

void ggq_dev_set_trace(void* buf) { decode_set_trace(buf); }

int64_t ggq_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* ggq_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case GGQ_E_SHAPE: return "ggq: invalid shape (K must be a multiple of the block size; sizes/strides in range)";
        case GGQ_E_POINTER: return "ggq: null data pointer";
        case GGQ_E_FAMILY: return "ggq: requested kernel family does not support this shape";
        case GGQ_E_FORMAT: return "ggq: unknown quant format";
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "ggq: unknown error";
}

int ggq_version(void) { return GGQ_VERSION; }

}  // extern "C"
