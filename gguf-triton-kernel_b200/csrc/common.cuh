// common.cuh — host-side helpers shared by the translation units of libggq.so.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ggq {

// In-kernel N-split exchange (see ggq_peer_sync in include/ggq.h); world == 0 means disabled.
struct PeerSync {            // device-side view of ggq_peer_sync (include/ggq.h)
    int rank, world, x_owner;
    uint32_t epoch;
    uint32_t* counter;
    uint32_t* epoch_dev;      // replayable mode: epoch of a call = *epoch_dev + 1 (kernel-maintained)
    const uint8_t* X_alt;     // odd epochs read these activations ...
    __half* C_alt;            // ... and store here (column 0 of the full row)
    __half* C_full;           // even epochs: column 0 of the full [T, ldc] row
    uint4* x_land;            // LL lines of the activations, two parity halves
    uint4* x_land_peer[8];
    uint4* c_land;            // LL lines of the peers' output slices, two parity halves
    uint4* c_land_peer[8];
    uint32_t x_half_lines, c_half_lines;   // 16-byte lines per parity half
    uint32_t* status;         // first GGQ_SYNC_* code of a wait that gave up (or null)
    unsigned long long timeout_ns;
};

struct MmArgs {
    const uint8_t* W;     // packed weight rows
    const void* X;        // fp16 [T, ldx]
    void* C[8];           // fp16 outputs, each [T, ldc]
    int n_out;
    int64_t ldx, ldc;
    int64_t O, T, K;
    cudaStream_t stream;
    const PeerSync* sync;  // decode family only; nullptr = plain call
    int* ctas_out;
    const uint8_t* W2;     // fused SwiGLU (launch_decode_dual): the up matrix, W is the gate matrix; else nullptr
};

// Output pointers passed to kernels by value.  n > 1 = the same tile is also stored to peer-mapped
// buffers of other ranks (fused N-split all-gather).
struct OutPtrs {
    __half* p[8];
    int n;
};
inline OutPtrs make_outs(const MmArgs& a) {
    OutPtrs o;
    for (int i = 0; i < 8; ++i) o.p[i] = static_cast<__half*>(a.C[i < a.n_out ? i : 0]);
    o.n = a.n_out;
    return o;
}

void count_launch(int n = 1);
int num_sms();  // SM count of the current device (cached per device)

// families (one launcher per translation unit); return cudaError_t / GGQ_E_*
int launch_generic(int fmt, const MmArgs& a);
int launch_decode(int fmt, const MmArgs& a);
int launch_decode_dual(int fmt, const MmArgs& a, bool plan_only);  // decode_dual.cu; GGQ_E_FAMILY = no fused plan
int launch_prefill(int fmt, const MmArgs& a);
int launch_skinny(int fmt, const MmArgs& a);
bool decode_supports(int fmt, const MmArgs& a);
bool prefill_supports(int fmt, const MmArgs& a);
bool skinny_supports(int fmt, const MmArgs& a);
int decode_plan(int fmt, const MmArgs& a, int* out9, bool* wide = nullptr);   // wide: the Q4_K wide-chunk GEMV geometry
void decode_set_trace(void* buf);
int skinny_describe(int fmt, const MmArgs& a, char* out, int cap);

int launch_dequant(int fmt, const uint8_t* W, void* out, int64_t O, int64_t K, cudaStream_t s);
// vectorized dequantize built on the prefill family's dequant64(); returns 0 if the shape is not eligible
int launch_dequant64(int fmt, const uint8_t* W, void* out, int64_t O, int64_t K, cudaStream_t s, int* rc);

}  // namespace ggq
