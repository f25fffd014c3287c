// prefill.cu — tensor-bound family: C[T, O] = X[T, K] . dequant(W)[O, K]^T for large T on tcgen05.
//
// A CTA PAIR (cta_group::2, cluster of 2) computes a 512-token x 256-out-feature tile: two M256 N256 fp32
// accumulators fill all 512 TMEM columns of both SMs, so every dequantized weight is reused by 512 tokens.
// Warp roles per CTA (10 warps):
//
//   warp 0      TMA producer   X tiles [2 x 128 tokens, 64 k] (2-D tensor map, SWIZZLE_128B, .cta_group::2) and
//                              packed weight "units" [128 rows, one block column] (2-D tensor map over the raw
//                              GGUF bytes viewed as int32 — rows are copied verbatim, no repacking)
//   warp 1      MMA issuer     leader CTA only, one elected thread: tcgen05.mma.cta_group::2.kind::f16
//                              (M256 N256 K16), A and B from shared memory, D in TMEM; tcgen05.commit releases
//                              the stage in both CTAs and finally publishes the accumulators
//   warps 2..9  dequant        two ping-pong groups of 128 threads = the CTA's 128 weight rows: packed unit row ->
//                              64 bit-exact fp16 weights (prefill_tile.cuh) -> K-major SWIZZLE_128B B tile
//                              (hand-swizzled 16-byte stores + fence.proxy.async), then the same warps run the
//                              epilogue: tcgen05.ld 32x32b.x32 -> fp16 -> global
//
// Roofline: 2*T*O*K FLOP on the fp16 tensor pipe; packed weights are read T/512 times (L2-resident
// across the token tiles that run concurrently), X is read O/256 times.
#include <algorithm>
#include <cstdlib>

#include "../../include/ggq.h"
#include "common.cuh"
#include "formats.cuh"
#include "prefill_tile.cuh"
#include "ptx.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace ggq {
namespace pre {

constexpr int BK = 64;           // k per pipeline stage: one 128-byte swizzle atom of fp16
constexpr int UNIT_ROWS = 128;   // rows of a packed staging unit (one TMA box)
constexpr int DQ_WARPS = 8;

struct Params {
    OutPtrs outs;
    int64_t ldc, O, T;
    int K;
};

// =================================================================================================
// 2-CTA variant (cta_group::2): a CTA PAIR computes 512 tokens x 256 out-features.  Each UMMA is
// M256 N256 K16 across the pair: every CTA contributes the A rows of its own 128 tokens and the B rows of
// its own 128 out-features, so a CTA dequantizes only HALF of the weight tile (128 rows) per stage while
// the tensor pipes of both SMs stay as busy as in the 1-CTA kernel — the dequant cost per FLOP halves
// and the packed staging ring shrinks enough to hold two full block columns (no TMA bubble per column).
//   * TMA loads of X use .cta_group::2 and complete on the LEADER's x_full barrier (both CTAs' bytes);
//   * dequant warps of both CTAs arrive on the leader's b_full barrier (remote mbarrier arrive);
//   * only the leader issues tcgen05.mma.cta_group::2; tcgen05.commit multicasts to both CTAs' barriers.
// =================================================================================================
namespace two {

constexpr int STAGES2 = 3;
constexpr int A2_BYTES = 2 * 128 * BK * 2;   // 32 KB: this CTA's 128 tokens of both 256-token MMA sets
constexpr int B2_BYTES = 128 * BK * 2;       // 16 KB: this CTA's 128 out-feature rows
constexpr int STAGE2_BYTES = A2_BYTES + B2_BYTES;
constexpr int NUM_THREADS2 = 320;

template <int FMT> struct Ring2 {
    static constexpr int DEPTH = FMT == 2 ? 2 : 3;  // units = whole block columns of this CTA's 128 rows
    static constexpr int UNIT_BYTES = 128 * Unit<FMT>::BOX_BYTES;
    static constexpr int KB_PER_UNIT = Unit<FMT>::UNIT_K / BK;
};
template <int FMT> constexpr int smem_bytes2() {
    return 1024 + STAGES2 * STAGE2_BYTES + Ring2<FMT>::DEPTH * Ring2<FMT>::UNIT_BYTES + 256;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {  // acquire at cluster scope
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d_2cta(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(bar_cluster_addr)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {  // arrives on `bar` in BOTH CTAs of the pair
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}
// M = 256 (pair), N = 256
constexpr uint32_t IDESC2 = (1u << 4) | (static_cast<uint32_t>(256 >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);

template <int FMT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS2, 1)
prefill2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const Params p,
                const int pairs_t) {
    using U = Unit<FMT>;
    using R = Ring2<FMT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t{1023});
    uint8_t* stages = smem;                                   // [STAGES2][A(2 x 128 tokens) | B(128 rows)]
    uint8_t* units = smem + STAGES2 * STAGE2_BYTES;           // [DEPTH][128 rows][BOX_BYTES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(units + R::DEPTH * R::UNIT_BYTES);
    uint64_t* x_full = bars;                  // [STAGES2]  leader's is the one that counts (tx of both CTAs)
    uint64_t* b_full = bars + STAGES2;        // [STAGES2]  leader: its 8 dequant warps + 1 relayed arrival of the peer;
                                              //            peer: its 8 dequant warps (watched by the peer's relay thread)
    uint64_t* free_ = bars + 2 * STAGES2;     // [STAGES2]  per CTA, multicast commit
    uint64_t* w_full = bars + 3 * STAGES2;    // [DEPTH]    per CTA
    uint64_t* w_empty = w_full + R::DEPTH;    // [DEPTH]    per CTA
    uint64_t* acc_full = w_empty + R::DEPTH;  // [1]        per CTA, multicast commit
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int64_t tp = pair % pairs_t, op = pair / pairs_t;     // token tiles vary fastest
    const int64_t tok_base = tp * 512 + rank * 128;             // + 256 * set
    const int64_t o0 = op * 256;                                // pair's out-feature tile
    const int64_t orow0 = o0 + rank * 128;                      // this CTA's B rows
    const int num_kb = p.K / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES2; ++s) {
            mbar_init(&x_full[s], 1);
            mbar_init(&b_full[s], DQ_WARPS / 2 + (rank == 0 ? 1 : 0));
            mbar_init(&free_[s], 1);
        }
        for (int u = 0; u < R::DEPTH; ++u) {
            mbar_init(&w_full[u], 1);
            mbar_init(&w_empty[u], DQ_WARPS);
        }
        mbar_init(acc_full, 1);
        fence_mbar_init();
        prefetch_tmap(&map_x);
        prefetch_tmap(&map_w);
    }
    if (warp == 1) tmem_alloc2(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();   // barriers of both CTAs are initialised before anyone signals them remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                if (kb % R::KB_PER_UNIT == 0) {
                    const int col = kb / R::KB_PER_UNIT;
                    const int slot = col % R::DEPTH;
                    const uint32_t use = static_cast<uint32_t>(col / R::DEPTH);
                    if (use > 0) mbar_wait(&w_empty[slot], (use - 1) & 1u);
                    mbar_arrive_expect_tx(&w_full[slot], R::UNIT_BYTES);
                    tma_load_2d(units + slot * R::UNIT_BYTES, &map_w, ((col * U::UNIT_BYTES) & ~15) / 4,
                                static_cast<int>(orow0), &w_full[slot]);
                }
                const int s = kb % STAGES2;
                const uint32_t use = static_cast<uint32_t>(kb / STAGES2);
                if (use > 0) mbar_wait(&free_[s], (use - 1) & 1u);
                const uint32_t xbar = map_to_cta(smem_u32(&x_full[s]), 0);  // the leader's barrier
                if (rank == 0) mbar_arrive_expect_tx(&x_full[s], 2 * A2_BYTES);  // bytes of both CTAs
                uint8_t* a = stages + s * STAGE2_BYTES;
                tma_load_2d_2cta(a, &map_x, kb * BK, static_cast<int>(tok_base), xbar);
                tma_load_2d_2cta(a + A2_BYTES / 2, &map_x, kb * BK, static_cast<int>(tok_base) + 256, xbar);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        if (rank == 0 && lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES2;
                const uint32_t ph = static_cast<uint32_t>(kb / STAGES2) & 1u;
                mbar_wait(&x_full[s], ph);
                mbar_wait_cluster(&b_full[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(stages + s * STAGE2_BYTES);
                const uint32_t b_addr = a_addr + A2_BYTES;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t ad = smem_desc_sw128(a_addr + t * (A2_BYTES / 2) + k * 32);
                        const uint64_t bd = smem_desc_sw128(b_addr + k * 32);
                        umma_f16_2cta(tmem_base + t * 256, ad, bd, IDESC2, (kb | k) != 0 ? 1u : 0u);
                    }
                }
                umma_commit_2cta(&free_[s]);
            }
            umma_commit_2cta(acc_full);
        } else if (rank == 1 && lane == 0) {
            // ---- relay (peer CTA): one cluster-scope release per stage instead of one per dequant warp — a
            // cluster-scope mbarrier arrive costs a membar, which stalled the dequant warps (ncu: stall_membar)
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES2;
                mbar_wait(&b_full[s], static_cast<uint32_t>(kb / STAGES2) & 1u);
                mbar_arrive_cluster(map_to_cta(smem_u32(&b_full[s]), 0));
            }
        }
    } else {
        // ================= dequant warps: two ping-pong groups of 4 warps =================
        // Group g (128 threads = the CTA's 128 weight rows) dequantizes the stages kb = g, g+2, ...: every thread
        // has two MMA periods per stage, which hides the fixed latency of a stage hand-over (barrier waits,
        // proxy fence, relay to the leader) behind the other group's stage.
        const int dq = threadIdx.x - 64;        // 0..255
        const int dwarp = warp - 2;
        const int urow = dq & 127;              // B row inside this CTA's half tile
        const int grp = dq >> 7;
        const uint32_t sw = static_cast<uint32_t>(urow & 7);
        for (int kb = grp; kb < num_kb; kb += 2) {
            const int col = kb / R::KB_PER_UNIT, kin = kb % R::KB_PER_UNIT;
            const int slot = col % R::DEPTH;
            mbar_wait(&w_full[slot], static_cast<uint32_t>(col / R::DEPTH) & 1u);
            const int s = kb % STAGES2;
            const uint32_t use = static_cast<uint32_t>(kb / STAGES2);
            if (use > 0) mbar_wait(&free_[s], (use - 1) & 1u);
            uint4 v[8];
            const int off = (col * U::UNIT_BYTES) & 15;
            if constexpr (FMT == 2)   // conflict-free 128-bit loads + register realignment (same arithmetic as dequant64)
                dequant_q6_k_sm<2>(units + slot * R::UNIT_BYTES + urow * U::BOX_BYTES, off, kin, v);
            else if constexpr (FMT == 0)
                dequant_q8_0_sm<2>(units + slot * R::UNIT_BYTES + urow * U::BOX_BYTES, off, kin, v);
            else
                dequant64(U{}, units + slot * R::UNIT_BYTES + urow * U::BOX_BYTES, off, kin, v);
            uint8_t* brow = stages + s * STAGE2_BYTES + A2_BYTES + (urow >> 3) * 1024 + (urow & 7) * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(brow + ((static_cast<uint32_t>(j) ^ sw) << 4)) = v[j];
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&b_full[s]);  // own CTA's barrier (the peer's is relayed to the leader by its warp 1)
                // last stage of this block column that this group handles -> the unit may be overwritten
                if (kin + 2 >= R::KB_PER_UNIT || kb + 2 >= num_kb) mbar_arrive(&w_empty[slot]);
            }
        }
        // ---- epilogue: this CTA's 128 TMEM lanes = its 128 tokens of each set, all 256 out-features ----
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int chalf = dwarp >> 2;
        const bool vec_ok = (p.ldc % 8 == 0);
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
            const int64_t tok = tok_base + t * 256 + q * 32 + lane;
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                const int col0 = chalf * 128 + cc * 32;
                uint32_t r[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(t * 256 + col0), r);
                tmem_ld_wait();
                if (tok < p.T) {
                    uint32_t h[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) h[i] = pack2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
                    const int64_t oc = o0 + col0;
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        if (o >= p.outs.n) break;
                        __half* dst = p.outs.p[o] + tok * p.ldc + oc;
                        if (vec_ok && oc + 32 <= p.O && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                reinterpret_cast<uint4*>(dst)[i] = make_uint4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (oc + i < p.O)
                                    dst[i] = __ushort_as_half(static_cast<unsigned short>((h[i >> 1] >> (16 * (i & 1))) & 0xffffu));
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    cluster_sync_all();   // nobody signals a barrier of, or reads the shared memory of, a CTA that has exited
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, 512);
    }
}

}  // namespace two

// ---- host: tensor maps ------------------------------------------------------------------------
template <int FMT>
static int launch_t(const MmArgs& a) {
    using U = Unit<FMT>;
    const int64_t rowB = a.K / U::QK * U::BLK;
    alignas(64) CUtensorMap map_x, map_w;
    // X: fp16 [T, K] (row stride ldx), box 64 k x 128 tokens, 128-byte swizzle; rows >= T are zero-filled
    if (!make_map_2d(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, a.X, static_cast<uint64_t>(a.K), static_cast<uint64_t>(a.T),
                     static_cast<uint64_t>(a.ldx) * 2, BK, 128, CU_TENSOR_MAP_SWIZZLE_128B))
        return static_cast<int>(cudaErrorInvalidValue);
    // W: the packed rows viewed as int32 [O, rowB/4], box BOX_BYTES/4 x 128 rows; rows >= O and bytes past the end of a row
    // are zero-filled
    if (!make_map_2d(&map_w, CU_TENSOR_MAP_DATA_TYPE_INT32, a.W, static_cast<uint64_t>(rowB / 4), static_cast<uint64_t>(a.O),
                     static_cast<uint64_t>(rowB), U::BOX_BYTES / 4, UNIT_ROWS, CU_TENSOR_MAP_SWIZZLE_NONE))
        return static_cast<int>(cudaErrorInvalidValue);
    Params p;
    p.outs = make_outs(a);
    p.ldc = a.ldc;
    p.O = a.O;
    p.T = a.T;
    p.K = static_cast<int>(a.K);
    auto kern2 = two::prefill2_kernel<FMT>;
    static int configured2[64] = {0};
    int dev2 = 0;
    cudaGetDevice(&dev2);
    if (dev2 >= 0 && dev2 < 64 && !configured2[dev2]) {
        cudaError_t e = cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, two::smem_bytes2<FMT>());
        if (e != cudaSuccess) return static_cast<int>(e);
        configured2[dev2] = 1;
    }
    // token tiles vary fastest, so the pairs resident at the same time read the same packed weight tiles (L2 hits)
    const int pairs_t = static_cast<int>((a.T + 511) / 512);
    const int64_t pairs_o = (a.O + 255) / 256;
    const int64_t ctas = 2 * pairs_t * pairs_o;
    if (ctas > 0x7fffffff) return GGQ_E_SHAPE;
    kern2<<<static_cast<unsigned>(ctas), two::NUM_THREADS2, two::smem_bytes2<FMT>(), a.stream>>>(map_x, map_w, p, pairs_t);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

// ---- standalone dequantize through the same dequant64 functions (global-memory source) -----------
template <int FMT>
__global__ void __launch_bounds__(256) dequant64_kernel(const uint8_t* __restrict__ W, __half* __restrict__ out, int64_t O,
                                                        int K) {
    using U = Unit<FMT>;
    const int64_t rowB = static_cast<int64_t>(K / U::QK) * U::BLK;
    const int nkb = K / BK;
    const int64_t total = O * nkb;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t row = i / nkb;
        const int kb = static_cast<int>(i % nkb);
        const int col = kb / (U::UNIT_K / BK), kin = kb % (U::UNIT_K / BK);
        const int byte0 = col * U::UNIT_BYTES;
        uint4 v[8];
        dequant64(U{}, W + row * rowB + (byte0 & ~15), byte0 & 15, kin, v);
        uint4* dst = reinterpret_cast<uint4*>(out + row * K + static_cast<int64_t>(kb) * BK);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = v[j];
    }
}

}  // namespace pre

// rows must be whole 16-byte vectors and K a multiple of the staging unit (256; 128 for Q8_0)
static bool aligned_rows(int fmt, const uint8_t* W, int64_t K) {
    const int unit_k = fmt == GGQ_Q8_0 ? 128 : 256;
    if (K < unit_k || K % unit_k != 0) return false;
    if (reinterpret_cast<uintptr_t>(W) & 15) return false;
    return ((K / fmt_qk(fmt)) * fmt_blk(fmt)) % 16 == 0;
}

bool prefill_supports(int fmt, const MmArgs& a) {
    if (a.T < 1 || a.O < 1 || !aligned_rows(fmt, a.W, a.K)) return false;
    if ((reinterpret_cast<uintptr_t>(a.X) & 15) || (a.ldx & 7)) return false;
    if (a.T > (int64_t{1} << 30) || a.O > (int64_t{1} << 30)) return false;
    return true;
}

int launch_prefill(int fmt, const MmArgs& a) {
    switch (fmt) {
        case GGQ_Q8_0: return pre::launch_t<0>(a);
        case GGQ_Q4_K: return pre::launch_t<1>(a);
        case GGQ_Q6_K: return pre::launch_t<2>(a);
    }
    return GGQ_E_FORMAT;
}

// returns 1 if handled (out written asynchronously), 0 if the shape needs the scalar dequant kernel
int launch_dequant64(int fmt, const uint8_t* W, void* out, int64_t O, int64_t K, cudaStream_t s, int* rc) {
    if (!aligned_rows(fmt, W, K) || (reinterpret_cast<uintptr_t>(out) & 15)) return 0;
    const int64_t total = O * (K / pre::BK);
    const int64_t blocks = std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(num_sms()) * 8);
    __half* o = static_cast<__half*>(out);
    switch (fmt) {
        case GGQ_Q8_0: pre::dequant64_kernel<0><<<static_cast<unsigned>(blocks), 256, 0, s>>>(W, o, O, static_cast<int>(K)); break;
        case GGQ_Q4_K: pre::dequant64_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, s>>>(W, o, O, static_cast<int>(K)); break;
        case GGQ_Q6_K: pre::dequant64_kernel<2><<<static_cast<unsigned>(blocks), 256, 0, s>>>(W, o, O, static_cast<int>(K)); break;
        default: *rc = GGQ_E_FORMAT; return 1;
    }
    count_launch();
    *rc = static_cast<int>(cudaGetLastError());
    return 1;
}

}  // namespace ggq
