// decode_impl.cuh — kernel, planner and launcher templates of the decode family, included by decode.cu (the mm entry
// points) and decode_dual.cu (the fused SwiGLU up-projection), which instantiate disjoint sets of kernels.
//
// HBM-bound skinny GEMM (T <= 16 tokens per pass): C[T, O] = X[T, K] . dequant(W)[O, K]^T
//
// Persistent kernel (12 warps per SM, or two 8-warp CTAs), no dedicated producer: every warp runs its OWN ring of 2-D
// TMA boxes (cp.async.bulk.tensor + mbarrier complete_tx) over the packed bytes of 16-row tiles, so there is no
// cross-warp synchronisation on the weight stream at all:
//
//     for each of my (tile, k-chunk) items:   wait(full[stage]) -> prep scales -> 16 x {8,16} MMA tile
//                                             -> re-arm the stage with the item STAGES ahead
//
// A stage holds one chunk (KGeo<FMT, WIDE>::CHUNK_BLOCKS blocks: the format's tile geometry, doubled for the Q4_K GEMV) of
// each of the tile's 16 rows, copied verbatim (16-byte aligned supersets where a chunk starts mid-vector, Q6_K).  Activations are staged once as raw fp16 rows by bulk
// copy, plus a table of per-sub-block activation sums that cancels the integer->fp16 bias (decode_tile.cuh).
//
// Work decomposition (static, chosen on the host, `ggq_decode_plan`):
//   * single K-slice (the activations of all tokens fit next to the rings): CTA c owns a contiguous range of tiles and
//     its tiles x chunks items are cut into NW equal ranges, one per warp; a tile cut by a range boundary is finished
//     by the warp holding its head, the others park their partial sums in shared memory and raise a flag;
//   * otherwise clusters of S = 2..8 CTAs split K: CTA r stages K-slice r of the activations, partial sums travel
//     rank S-1 -> ... -> 0 through per-warp DSMEM mailboxes (st.async + mbarrier);
//   * last resort: K-slices staged one after the other with AT = 4 live tiles per warp and KW warps per tile.
// T == 1 uses the GEMV tile code (`GV`): the 8 MMA columns carry 8 sub-blocks instead of 8 tokens.
// `DUAL`: fused SwiGLU up-projection, 8 gate + 8 up rows per tile (decode_dual.cu).
// Launches use programmatic stream serialization: the prologue and the weight prefetch of a launch overlap the drain
// of the previous kernel; `griddepcontrol.wait` sits in front of the first access to the activations.
// With a `ggq_peer_sync` the kernel also does the N-split exchange (activations and output tiles as flag-in-data lines
// into peer-mapped landing buffers, bounded waits).
// HBM traffic = packed weight bytes once (+ <= 3 % activations/outputs); roofline: HBM bandwidth.
#pragma once
#include <cstring>
#include <algorithm>
#include <cstdlib>

#include "../../include/ggq.h"
#include "common.cuh"
#include "decode_tile.cuh"
#include "formats.cuh"
#include "ptx.cuh"
#include "tma.cuh"

namespace ggq {
namespace dec {

constexpr int MAX_NW = 16;                // most warps per CTA of any configuration
constexpr int MAX_STAGES = 6;
constexpr int SMEM_LIMIT = 227 * 1024;    // opt-in dynamic shared memory per CTA on sm_100
constexpr int SMEM_LIMIT_2 = 113 * 1024;  // per CTA when two CTAs share an SM (228 KB - 2 x 1 KB reserved)

// Kernel geometry = the format's tile geometry (decode_tile.cuh) with the CHUNK — what one ring stage holds of a row —
// optionally doubled.  WIDE (Q4_K single-token GEMV): 4 blocks = 576 B per row and stage, filled by two back-to-back
// 288 B boxes: the per-chunk control code (barrier wait, refill, cursor arithmetic) runs half as often — the rings alone
// already stream at the copy peak, what the GEMV loses is instruction time (measured: lm_head 51.8 -> 49.4 us, 27.5 M ->
// 24.5 M warp instructions; with T = 8 the same geometry loses 30 %, gpurun_out/r2_ab.log).  The lane-level tile code
// only knows PREP_BLOCKS / SLOT and is the same for both.
template <int FMT, bool WIDE> struct KGeo : Geo<FMT> {};
template <> struct KGeo<1, true> : Geo<1> {
    static constexpr int CHUNK_BLOCKS = 4, CHUNK_ELEMS = 1024, CHUNK_BYTES = 576;
};

struct Params {
    const uint8_t* W;
    const uint8_t* X;
    OutPtrs outs;
    int64_t ldx_bytes, ldc, O, rowB;
    int T, K, nb;
    int num_tiles;     // ceil(O / 16)
    int KW;            // warps per tile
    int nc;            // chunks per row
    int cps;           // chunks per K-slice
    int n_slices;
    int num_batches;
    int stages;
    int l2_prefetch_bytes; // fused exchange: bytes of the CTA's weight range prefetched into L2 before the first wait
    unsigned long long* trace;  // dev (ggq_dev_set_trace): per-CTA globaltimer stamps of this launch, [grid][8], or null
    int dbg_skip_compute;  // GGQ_DECODE_NOCOMPUTE=1: stream the weights through the TMA rings but skip the math (roofline probe)
    PeerSync sync;      // world == 0: no cross-GPU synchronisation
    uint32_t x_stride;  // bytes between token rows in shared memory
    uint32_t off_bars, off_x, off_tbl, off_ring, off_scr, off_red, off_mbox;
};

// ---- flag-in-data ("LL") lines of the fused N-split exchange: 16 bytes = {data0, epoch, data1, epoch} ---------------
__device__ __forceinline__ void ll_store(uint4* dst, uint32_t d0, uint32_t d1, uint32_t flag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(d0), "r"(flag), "r"(d1), "r"(flag) : "memory");
}
__device__ __forceinline__ uint4 ll_load(const uint4* src) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src) : "memory");
    return v;
}
// Polls a line of this rank's landing buffer until both flag words carry `flag`.  Bounded: gives up after `timeout_ns`
// of the global timer (or as soon as another thread has given up) and records `code` in *status.  `dead` short-cuts
// every later wait of a thread that has given up once.
static __device__ __noinline__ uint2 ll_wait_slow(const uint4* src, uint32_t flag, const PeerSync& sy, uint32_t code, bool& dead) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 32; ++i) {
            const uint4 v = ll_load(src);
            if (v.y == flag && v.w == flag) return make_uint2(v.x, v.z);
        }
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const bool other = sy.status != nullptr && *reinterpret_cast<volatile const uint32_t*>(sy.status) != 0u;
        if (other || t - t0 > sy.timeout_ns) {
            if (sy.status != nullptr) atomicCAS(sy.status, 0u, code);
            dead = true;
            return make_uint2(0u, 0u);
        }
    }
}
__device__ __forceinline__ uint2 ll_wait(const uint4* src, uint32_t flag, const PeerSync& sy, uint32_t code, bool& dead) {
    if (dead) return make_uint2(0u, 0u);
    const uint4 v = ll_load(src);
    if (v.y == flag && v.w == flag) return make_uint2(v.x, v.z);
    return ll_wait_slow(src, flag, sy, code, dead);
}

// NW warps per CTA, MINB CTAs per SM (register budget), NT 8-token n-tiles, AT live tiles per warp
// GV: single-token (GEMV) tile code, see decode_tile.cuh
// DUAL: fused SwiGLU up-projection (ggq_mm_swiglu).  A tile is 8 rows of the gate matrix (map_w, MMA rows 0..7) over
// the SAME 8 rows of the up matrix (map_w2, MMA rows 8..15): the m16n8 accumulator fragment of a lane then holds
// gate[row g] in v[0..1] and up[row g] in v[2..3], and the store writes silu(gate) * up — no extra pass, no extra traffic.
template <int FMT, int NT, int AT, int NW, int MINB, bool GV = false, bool DUAL = false, bool WIDE = false>
__global__ void __launch_bounds__(NW * 32, MINB)
decode_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_w2, const Params p) {
    static_assert(!GV || (NT == 1 && AT == 1), "the GEMV tile code is single-token, one live tile");
    static_assert(!DUAL || AT == 1, "the fused SwiGLU store is part of the flat (tile, chunk) walk");
    constexpr int TILE_ROWS = DUAL ? 8 : 16;   // output rows per tile
    using G = KGeo<FMT, WIDE>;
    constexpr int SUBTILES = G::CHUNK_BLOCKS / G::PREP_BLOCKS;  // TMA boxes per stage
    constexpr int STAGE_BYTES = SUBTILES * 16 * G::SLOT;
    constexpr int SCR_BYTES = 16 * G::PREP_BLOCKS * G::SCRATCH_PER_BLOCK;
    extern __shared__ __align__(128) uint8_t smem[];

    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const Lane L{lane, lane >> 2, lane & 3};
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bars);  // [0]: activations, [1 + w*stages + s]: ring
    uint8_t* xs = smem + p.off_x;
    float* tbl = reinterpret_cast<float*>(smem + p.off_tbl);
    const int STG = p.stages;
    uint8_t* ring = smem + p.off_ring + static_cast<size_t>(w) * STG * STAGE_BYTES;
    uint8_t* scr = smem + p.off_scr + static_cast<size_t>(w) * SCR_BYTES;
    float* red = reinterpret_cast<float*>(smem + p.off_red);
    float* mbox = reinterpret_cast<float*>(smem + p.off_mbox);
    uint64_t* my_full = bars + 1 + w * STG;

    // Programmatic dependent launch: the next kernel in the stream may start its own prologue (barrier init, weight
    // prefetch) while this one runs; everything that depends on earlier kernels sits behind pdl_wait() in stage_x.
    // dev: phase stamps (0 entry, 1 barriers + first boxes issued, 2 pdl_wait passed, 3 activations staged, 4 warp 0's
    // items done, 5 CTA done)
    auto stamp = [&](int ev) {
        if (p.trace != nullptr && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.trace[blockIdx.x * 8 + ev] = t;
        }
    };
    stamp(0);
    pdl_launch_dependents();
    // cross-GPU exchange: the epoch of this launch (host-supplied, or kernel-maintained in replayable mode, where odd
    // epochs use the alternate activation / output buffers); settled in stage_x once the previous kernel is complete
    uint32_t epoch = p.sync.epoch;
    bool alt = false;
    uint32_t* epoch_word = reinterpret_cast<uint32_t*>(bars + 1 + MAX_NW * MAX_STAGES + 2 * MAX_NW) + MAX_NW;
    __half* c_full = nullptr;   // fused exchange: column 0 of this epoch's full [T, ldc] result (set in stage_x)
    bool ll_dead = false;       // a cross-GPU wait of this thread has given up
    auto out_ptr = [&](int o) -> __half* {
        return (c_full != nullptr && o == 0) ? c_full + static_cast<int64_t>(p.sync.rank) * p.O : p.outs.p[o];
    };
    // every warp initialises its own barriers (ring stages, cluster mailbox pair) and starts streaming right away; the
    // barrier of the activations (bars[0], warp 0) is first used after the __syncthreads at the top of stage_x
    if (lane == 0) {
        if (w == 0) {
            prefetch_tmap(&map_w);
            if constexpr (DUAL) prefetch_tmap(&map_w2);
            mbar_init(&bars[0], 1);
        }
        for (int s = 0; s < STG; ++s) mbar_init(my_full + s, 1);
        mbar_init(&bars[1 + MAX_NW * MAX_STAGES + w], 1);            // cluster mailbox: filled
        mbar_init(&bars[1 + MAX_NW * MAX_STAGES + MAX_NW + w], 1);   // cluster mailbox: free
        fence_mbar_init();
    }
    __syncwarp();

    // fused exchange: send the 16 rows x T tokens of a finished tile (just stored to this rank's C by this warp) to
    // every peer's landing buffer as LL lines
    auto ll_send_tile = [&](int tile) {
        __syncwarp();
        const int row0 = tile * 16;
        const int nl = min(16, static_cast<int>(p.O) - row0) >> 2;   // lines per token
        const int per_lines = static_cast<int>(p.O >> 2);
        const size_t par_off = static_cast<size_t>(epoch & 1u) * p.sync.c_half_lines;
        const __half* own = c_full + static_cast<int64_t>(p.sync.rank) * p.O;
        const int n = p.T * nl;
        for (int j = lane; j < n * (p.sync.world - 1); j += 32) {
            const int pr = j / n, rem = j - pr * n;
            const int t = rem / nl, q = rem - t * nl;
            const int peer = pr + (pr >= p.sync.rank ? 1 : 0);
            const uint2 v = __ldcg(reinterpret_cast<const uint2*>(own + static_cast<int64_t>(t) * p.ldc + row0 + 4 * q));
            ll_store(p.sync.c_land_peer[peer] + par_off + (static_cast<size_t>(p.sync.rank * p.T + t) * per_lines + tile * 4 + q),
                     v.x, v.y, epoch);
        }
    };

    const int KW = p.KW, WT = NW / KW, tg = w / KW, sub = w % KW;
    // (K-sliced fallback) live tile `a` of batch `batch` is round batch*AT + a; a round spreads WT tiles over every CTA
    // ---- epilogue of one batch: (reduce over the KW warps of a tile,) round to fp16, store ------------
    auto epilogue = [&](Acc<NT>* acc, int batch) {
        if constexpr (GV) gemv_finalize(acc[0]);
        if (KW > 1) {
#pragma unroll
            for (int a = 0; a < AT; ++a) {
                __syncthreads();  // previous readers of `red` are done
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) red[((w * NT + nt) * 4 + i) * 32 + lane] = acc[a].v[nt][i];
                __syncthreads();
                if (sub == 0) {
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float v = acc[a].v[nt][i];
                            for (int k = 1; k < KW; ++k) v += red[(((w + k) * NT + nt) * 4 + i) * 32 + lane];
                            acc[a].v[nt][i] = v;
                        }
                }
            }
        }
        if (sub != 0) return;
#pragma unroll
        for (int a = 0; a < AT; ++a) {
            const int tile = ((batch * AT + a) * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x)) * WT + tg;
            if (tile >= p.num_tiles) continue;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (GV && (i & 1)) continue;  // single token: column 0 only
                    const int64_t row = static_cast<int64_t>(tile) * 16 + L.g + ((i & 2) ? 8 : 0);
                    const int col = 8 * nt + 2 * L.t + (i & 1);
                    if (row < p.O && col < p.T) {
                        const __half h = __float2half_rn(acc[a].v[nt][i]);
                        const int64_t at = col * p.ldc + row;
                        out_ptr(0)[at] = h;
                        for (int o = 1; o < p.outs.n; ++o) out_ptr(o)[at] = h;
                    }
                }
            if (c_full != nullptr) ll_send_tile(tile);
        }
    };

    const bool skip_math = p.dbg_skip_compute != 0;
    const int tok0 = min(L.g, p.T - 1), tok1 = min(8 + L.g, p.T - 1);
    uint32_t ring_phase = 0, x_phase = 0;
    int cstage = 0;

    // ---- stage one K-slice of the activations + its block-sum table (all threads) -------------------
    auto stage_x = [&](int slice, bool first) {
        const int e0 = slice * p.cps * G::CHUNK_ELEMS;
        const int ne = min(p.cps * G::CHUNK_ELEMS, p.K - e0);
        const bool xsync = p.sync.world > 1;
        __syncthreads();  // every warp is done with the previous slice's x / tbl
        if (tid == 0 && first) {
            pdl_wait();  // earlier kernels in the stream (the producers of X, earlier users of C) are complete
            stamp(2);
            if (p.sync.epoch_dev != nullptr)  // written back by the last CTA of the previous launch at its very end
                epoch = *reinterpret_cast<volatile const uint32_t*>(p.sync.epoch_dev) + 1u;
            *epoch_word = epoch;
        }
        if (xsync && first) {
            __syncthreads();
            epoch = *epoch_word;
            alt = p.sync.epoch_dev != nullptr && (epoch & 1u) != 0u;
            c_full = alt ? p.sync.C_alt : p.sync.C_full;
        }
        const bool owner = !xsync || p.sync.rank == p.sync.x_owner;
        if (xsync && owner && first) {
            // owner rank: CTA b < world pushes the activations to peer b as LL lines (data and "it is there" in the same
            // 16-byte store: one NVLink hop, no fence), so the pushes to all peers run in parallel.  (The grid has >= world
            // CTAs whenever there are >= world tiles; else CTA 0 serves every peer.)
            const int nserve = min(static_cast<int>(gridDim.x), p.sync.world);
            if (static_cast<int>(blockIdx.x) < nserve) {
                const uint8_t* const Xsrc = alt ? p.sync.X_alt : p.X;
                const int lpr = p.K >> 2, total = p.T * lpr;
                const bool all = nserve < p.sync.world;
                for (int r = 0; r < p.sync.world; ++r) {
                    if (all ? blockIdx.x != 0 : r != static_cast<int>(blockIdx.x)) continue;
                    uint4* dst = p.sync.x_land_peer[r];
                    if (dst == nullptr) continue;
                    dst += static_cast<size_t>(epoch & 1u) * p.sync.x_half_lines;
                    for (int i = tid; i < total; i += NW * 32) {
                        const int t = i / lpr, c = i - t * lpr;
                        const uint2 v = *reinterpret_cast<const uint2*>(Xsrc + t * p.ldx_bytes + static_cast<int64_t>(c) * 8);
                        ll_store(dst + i, v.x, v.y, epoch);
                    }
                }
            }
        }
        if (owner) {
            if (tid == 0) {
                const uint8_t* const Xg = alt ? p.sync.X_alt : p.X;
                mbar_arrive_expect_tx(&bars[0], static_cast<uint32_t>(p.T * ne * 2));
                for (int t = 0; t < p.T; ++t)
                    bulk_g2s(xs + t * p.x_stride, Xg + t * p.ldx_bytes + 2 * static_cast<int64_t>(e0),
                             static_cast<uint32_t>(ne * 2), &bars[0]);
            }
            mbar_wait(&bars[0], x_phase);
            x_phase ^= 1;
        } else {
            // peers: the activations arrive as LL lines in the local landing buffer; every CTA polls the lines of this
            // K-slice and unpacks them straight into shared memory
            const uint4* src = p.sync.x_land + static_cast<size_t>(epoch & 1u) * p.sync.x_half_lines;
            const int lpr = p.K >> 2, l0 = e0 >> 2, nl = ne >> 2;
            for (int i = tid; i < p.T * nl; i += NW * 32) {
                const int t = i / nl, c = i - t * nl;
                const uint2 v = ll_wait(src + t * lpr + l0 + c, epoch, p.sync, GGQ_SYNC_TIMEOUT_X, ll_dead);
                *reinterpret_cast<uint2*>(xs + t * p.x_stride + c * 8) = v;
            }
            __syncthreads();
        }
        stage_activations<FMT, NT, GV>(xs, p.x_stride, tbl, ne, p.T, tid, NW * 32);
        __syncthreads();
    };

    // ---- consume the chunk sitting in ring stage `cstage` (chunk `ci` of K-slice `slice`) -----------
    // ring stages are used round-robin, so all stages of one lap share the mbarrier parity `ring_phase`
    StageArgs sa;
    sa.xrow[0] = xs + tok0 * p.x_stride;
    sa.xrow[1] = xs + tok1 * p.x_stride;
    sa.xv[0] = L.g < p.T;
    sa.xv[1] = 8 + L.g < p.T;
    sa.tbl = tbl;
    sa.scratch = scr;
    auto consume = [&](int slice, int ci, Acc<NT>& acc) {
        mbar_wait(my_full + cstage, ring_phase);
        if (skip_math) return;
        const int b0 = (slice * p.cps + ci) * G::CHUNK_BLOCKS;
        StageArgs s = sa;
        s.data_off = (b0 * G::BLK) & 15;
        if (b0 + G::CHUNK_BLOCKS <= p.nb) {  // whole chunk (warp-uniform): no bounds checks anywhere below
#pragma unroll
            for (int u = 0; u < SUBTILES; ++u) {
                s.rows = ring + cstage * STAGE_BYTES + u * 16 * G::SLOT;
                s.data_off = ((b0 + u * G::PREP_BLOCKS) * G::BLK) & 15;
                s.nblk = G::PREP_BLOCKS;
                s.k0 = ci * G::CHUNK_ELEMS + u * G::PREP_BLOCKS * G::QK;
                Tile<FMT, NT, GV>::template prep<true>(L, s);
                __syncwarp();
                Tile<FMT, NT, GV>::template compute<true>(L, s, acc);
                __syncwarp();  // all lanes are done reading the stage and the scratch
            }
        } else {
            const int nblk = p.nb - b0;
#pragma unroll
            for (int u = 0; u < SUBTILES; ++u) {
                const int b = u * G::PREP_BLOCKS;
                if (b < nblk) {
                    s.rows = ring + cstage * STAGE_BYTES + u * 16 * G::SLOT;
                    s.data_off = ((b0 + b) * G::BLK) & 15;
                    s.nblk = min(G::PREP_BLOCKS, nblk - b);
                    s.k0 = ci * G::CHUNK_ELEMS + b * G::QK;
                    Tile<FMT, NT, GV>::template prep<false>(L, s);
                    __syncwarp();
                    Tile<FMT, NT, GV>::template compute<false>(L, s, acc);
                    __syncwarp();
                }
            }
        }
    };
    auto next_stage = [&]() {
        if (++cstage == STG) {
            cstage = 0;
            ring_phase ^= 1u;
        }
    };
    // one elected lane: the 2-D TMA boxes (16 rows x SLOT bytes of the raw packed rows, each starting at the 16-byte
    // aligned superset of its blocks; rows >= O and bytes past the row end are zero-filled) of chunk `chunk` of the
    // tile whose first row is `row0`, into ring stage `stage`
    auto issue_boxes = [&](int row0, int chunk, int stage) {
        if (lane == 0) {
            const int goff = chunk * (G::CHUNK_BLOCKS * G::BLK);
            uint64_t* bar = my_full + stage;
            mbar_arrive_expect_tx(bar, STAGE_BYTES);
#pragma unroll
            for (int u = 0; u < SUBTILES; ++u) {
                uint8_t* dst = ring + stage * STAGE_BYTES + u * 16 * G::SLOT;
                const int x = ((goff + u * G::PREP_BLOCKS * G::BLK) & ~15) >> 2;
                tma_load_2d(dst, &map_w, x, row0, bar);   // DUAL: the maps have 8-row boxes
                if constexpr (DUAL) tma_load_2d(dst + 8 * G::SLOT, &map_w2, x, row0, bar);
            }
        }
    };

    if constexpr (AT == 1) {
        // ======== single K-slice (the plan has AT == 1 exactly then): flat (tile, chunk) items ==============
        // The CTA owns the contiguous tiles [tile_lo, tile_hi); their tiles x chunks items are cut into NW equal
        // contiguous ranges, one per warp, so every warp streams the same number of bytes whatever O and K are.
        // A tile cut by a range boundary is finished by the warp that holds its head: the warps holding the rest
        // (always the FIRST thing in their range) park their partial sums in `red` and raise a flag.
        // Cluster split-K (S = p.n_slices > 1, launched as clusters of S CTAs): when the activations of all tokens do
        // not fit in one CTA's shared memory, CTA r of a cluster stages only K-slice r of them and walks the SAME
        // (tile, chunk) items over its slice of every row.  The partial sums of a warp's part of a tile travel down
        // the cluster, rank S-1 -> ... -> rank 0, through one distributed-shared-memory mailbox per warp (the warps
        // with the same index run the same item sequence in every rank); rank 0 then finishes the tile as usual.
        const int S = p.n_slices;
        const int crank = S > 1 ? static_cast<int>(cl_ctarank()) : 0;
        const int ncl = static_cast<int>(gridDim.x) / S, cl = static_cast<int>(blockIdx.x) / S;
        const int tile_lo = static_cast<int>(static_cast<int64_t>(cl) * p.num_tiles / ncl);
        const int tile_hi = static_cast<int>(static_cast<int64_t>(cl + 1) * p.num_tiles / ncl);
        const int nsc = p.cps;                                    // logical chunks per tile, the same in every rank
        const int nsc_mine = min(p.cps, p.nc - crank * p.cps);    // chunks that exist in my slice (the last may be short)
        const int items = (tile_hi - tile_lo) * nsc;
        auto range_begin = [&](int k) { return static_cast<int>(static_cast<int64_t>(k) * items / NW); };
        const int ibeg = range_begin(w), iend = range_begin(w + 1);
        uint64_t* mb_full = bars + 1 + MAX_NW * MAX_STAGES + w;            // my mailbox has been filled (by rank + 1)
        uint64_t* mb_empty = bars + 1 + MAX_NW * MAX_STAGES + MAX_NW + w;  // rank - 1 has read my last message
        volatile uint32_t* flags = reinterpret_cast<volatile uint32_t*>(bars + 1 + MAX_NW * MAX_STAGES + 2 * MAX_NW);  // [NW] intra-CTA hand-off
        float* my_slot = red + static_cast<size_t>(w) * (NT * 4 * 32);
        float4* my_mbox = reinterpret_cast<float4*>(mbox) + static_cast<size_t>(w) * (NT * 32);
        uint32_t sent = 0, full_phase = 0, empty_phase = 0;

        int pi = ibeg, ptile = tile_lo + ibeg / nsc, pci = ibeg % nsc;  // producer cursor, STG real items ahead
        auto p_advance = [&]() {
            ++pi;
            if (++pci == nsc) {
                pci = 0;
                ++ptile;
            }
        };
        auto produce = [&](int stage) {
            issue_boxes(ptile * TILE_ROWS, crank * p.cps + pci, stage);
            p_advance();
            while (pi < iend && pci >= nsc_mine) p_advance();  // chunks past the end of a short last slice do not exist
        };
        while (pi < iend && pci >= nsc_mine) p_advance();
        for (int s = 0; s < STG && pi < iend; ++s) produce(s);
        if (tid < NW) flags[tid] = 0u;
        if (p.l2_prefetch_bytes > 0) {
            // fused exchange: this launch is about to wait — for the previous step's exchange to finish (pdl_wait) and,
            // on the peers, for the activations to arrive over NVLink.  Pull the head of the CTA's weight range
            // (contiguous whole rows) into L2 meanwhile, so that HBM streams during those latencies.
            const uint8_t* base = p.W + static_cast<int64_t>(tile_lo) * 16 * p.rowB;
            const int64_t row_end = min(static_cast<int64_t>(tile_hi) * 16, p.O);
            const int64_t total = min((row_end - static_cast<int64_t>(tile_lo) * 16) * p.rowB, static_cast<int64_t>(p.l2_prefetch_bytes));
            for (int64_t off = static_cast<int64_t>(tid) * 4096; off < total; off += static_cast<int64_t>(NW) * 32 * 4096)
                prefetch_l2_bulk(base + off, static_cast<uint32_t>(min(static_cast<int64_t>(4096), total - off) & ~int64_t{15}));
        }
        if (S > 1) cl_sync();     // every CTA of the cluster runs and has initialised its barriers before anyone signals them
        stamp(1);
        stage_x(crank, true);     // (its barriers also publish the cleared flags inside the CTA)
        stamp(3);

        auto store_tile = [&](int tile, const Acc<NT>& acc) {
            if constexpr (DUAL) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        if (GV && i == 1) continue;  // single token: column 0 only
                        const int64_t row = static_cast<int64_t>(tile) * 8 + L.g;
                        const int col = 8 * nt + 2 * L.t + i;
                        if (row < p.O && col < p.T) {
                            // both projections rounded to fp16 first: the result equals silu(mmq(gate)) * mmq(up) of two calls
                            const float gte = __half2float(__float2half_rn(acc.v[nt][i]));
                            const float up = __half2float(__float2half_rn(acc.v[nt][i + 2]));
                            p.outs.p[0][col * p.ldc + row] = __float2half_rn(gte / (1.f + __expf(-gte)) * up);
                        }
                    }
            } else {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (GV && (i & 1)) continue;  // single token: column 0 only
                        const int64_t row = static_cast<int64_t>(tile) * 16 + L.g + ((i & 2) ? 8 : 0);
                        const int col = 8 * nt + 2 * L.t + (i & 1);
                        if (row < p.O && col < p.T) {
                            const __half h = __float2half_rn(acc.v[nt][i]);
                            const int64_t at = col * p.ldc + row;
                            out_ptr(0)[at] = h;
                            for (int o = 1; o < p.outs.n; ++o) out_ptr(o)[at] = h;
                        }
                    }
                if (c_full != nullptr) ll_send_tile(tile);
            }
        };

        int tile = tile_lo + ibeg / nsc, ci = ibeg % nsc;
        bool head = ci == 0;  // does this warp hold the first chunk of the tile it is working on?
        Acc<NT> acc;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc.v[nt][i] = 0.f;
        for (int i = ibeg; i < iend;) {
            if (ci < nsc_mine) {
                consume(crank, ci, acc);
                if (pi < iend) produce(cstage);
                next_stage();
            }
            ++i;
            ++ci;
            if (ci == nsc || i == iend) {  // my part of `tile` is done
                if constexpr (GV) gemv_finalize(acc);
                bool mine = true;          // does this CTA finish / hand over the part inside the CTA?
                if (S > 1) {
                    if (crank < S - 1) {   // add what the ranks above me accumulated for this part
                        if (lane == 0) mbar_arrive_expect_tx(mb_full, NT * 4 * 32 * 4);
                        mbar_wait(mb_full, full_phase);
                        full_phase ^= 1u;
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) {
                            const float4 v = my_mbox[nt * 32 + lane];
                            acc.v[nt][0] += v.x;
                            acc.v[nt][1] += v.y;
                            acc.v[nt][2] += v.z;
                            acc.v[nt][3] += v.w;
                        }
                        __syncwarp();
                        if (lane == 0) cl_mbar_arrive(cl_map(smem_u32(mb_empty), crank + 1));
                    }
                    if (crank > 0) {       // pass it down
                        if (sent > 0) {    // the previous message has been read
                            cl_mbar_wait(mb_empty, empty_phase);
                            empty_phase ^= 1u;
                        }
                        const uint32_t dst = cl_map(smem_u32(my_mbox), crank - 1);
                        const uint32_t dbar = cl_map(smem_u32(mb_full), crank - 1);
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt)
                            cl_st_async_f32x4(dst + (nt * 32 + lane) * 16, acc.v[nt][0], acc.v[nt][1], acc.v[nt][2], acc.v[nt][3], dbar);
                        ++sent;
                        mine = false;
                    }
                }
                if (!mine) {
                } else if (!head) {
                    // the tile began in an earlier warp: hand my partial sums to it
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int r = 0; r < 4; ++r) my_slot[(nt * 4 + r) * 32 + lane] = acc.v[nt][r];
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) flags[w] = 1u;
                } else {
                    if (ci != nsc) {
                        // the rest of the tile is in the following warps (each parks it before doing anything else)
                        const int tile_end = (tile - tile_lo + 1) * nsc;
                        for (int k = w + 1; k < NW; ++k) {
                            const int bk = range_begin(k);
                            if (bk >= tile_end) break;
                            if (range_begin(k + 1) == bk) continue;  // empty range
                            while (flags[k] == 0u) {
                            }
                            __threadfence_block();
                            const float* slot = red + static_cast<size_t>(k) * (NT * 4 * 32);
#pragma unroll
                            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                                for (int r = 0; r < 4; ++r) acc.v[nt][r] += slot[(nt * 4 + r) * 32 + lane];
                        }
                    }
                    store_tile(tile, acc);
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc.v[nt][r] = 0.f;
                head = true;
                if (ci == nsc) {
                    ci = 0;
                    ++tile;
                }
            }
        }
        if (S > 1) cl_sync();  // nobody writes into the shared memory of a CTA that has exited
    } else {
        // ======== K-sliced walk: AT live tiles per warp share each staged activation slice ========
        auto tile_of = [&](int batch, int a) -> int64_t {
            return (static_cast<int64_t>(batch * AT + a) * gridDim.x + blockIdx.x) * WT + tg;
        };
        auto slice_chunks = [&](int slice) { return min(p.cps, p.nc - slice * p.cps); };
        // producer cursor: walks exactly the item sequence the consumer loops below walk
        struct Cur {
            int batch, slice, a, ci;
            bool done;
        };
        auto normalize = [&](Cur& c) {
            while (true) {
                if (c.batch >= p.num_batches) { c.done = true; return; }
                if (c.slice >= p.n_slices) { c.slice = 0; c.a = 0; c.ci = sub; ++c.batch; continue; }
                if (c.a >= AT) { c.a = 0; c.ci = sub; ++c.slice; continue; }
                if (tile_of(c.batch, c.a) >= p.num_tiles || c.ci >= slice_chunks(c.slice)) { ++c.a; c.ci = sub; continue; }
                return;
            }
        };
        Cur pc{0, 0, 0, sub, false};
        normalize(pc);
        auto produce = [&](int stage) {
            issue_boxes(static_cast<int>(tile_of(pc.batch, pc.a) * 16), pc.slice * p.cps + pc.ci, stage);
            pc.ci += KW;
            normalize(pc);
        };
        for (int s = 0; s < STG && !pc.done; ++s) produce(s);

        for (int batch = 0; batch < p.num_batches; ++batch) {
            Acc<NT> acc[AT];
#pragma unroll
            for (int a = 0; a < AT; ++a)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[a].v[nt][i] = 0.f;
            for (int slice = 0; slice < p.n_slices; ++slice) {
                const int nsc = slice_chunks(slice);
                stage_x(slice, batch == 0 && slice == 0);
#pragma unroll
                for (int a = 0; a < AT; ++a) {
                    if (tile_of(batch, a) >= p.num_tiles) continue;  // warp-uniform
                    for (int ci = sub; ci < nsc; ci += KW) {
                        consume(slice, ci, acc[a]);
                        if (!pc.done) produce(cstage);
                        next_stage();
                    }
                }
            }
            epilogue(acc, batch);
        }
    }

    stamp(4);
    if (p.trace != nullptr && p.sync.world <= 1) {
        __syncthreads();
        stamp(5);
    }
    if (p.sync.world > 1) {
        // ---- fused N-split exchange, receive side: the peers' slices arrive as LL lines in this rank's landing buffer;
        // every CTA polls an equal share of them and writes plain fp16 into C.  When the kernel completes, the full
        // C[T, world * O] is present in this rank's buffer (nothing else to wait for: no flags, no fence).
        __syncthreads();
        const int per_lines = static_cast<int>(p.O >> 2), tl = p.T * per_lines;
        const int64_t L = static_cast<int64_t>(p.sync.world - 1) * tl;
        const int64_t lb = static_cast<int64_t>(blockIdx.x) * L / gridDim.x, le = static_cast<int64_t>(blockIdx.x + 1) * L / gridDim.x;
        const uint4* land = p.sync.c_land + static_cast<size_t>(epoch & 1u) * p.sync.c_half_lines;
        for (int64_t i = lb + tid; i < le; i += NW * 32) {
            const int sp = static_cast<int>(i / tl), rem = static_cast<int>(i - static_cast<int64_t>(sp) * tl);
            const int t = rem / per_lines, q = rem - t * per_lines;
            const int src = sp + (sp >= p.sync.rank ? 1 : 0);
            const uint2 v = ll_wait(land + (static_cast<size_t>(src * p.T + t) * per_lines + q), epoch, p.sync, GGQ_SYNC_TIMEOUT_PEER, ll_dead);
            *reinterpret_cast<uint2*>(c_full + static_cast<int64_t>(t) * p.ldc + static_cast<int64_t>(src) * p.O + 4 * q) = v;
        }
        __syncthreads();
        stamp(5);
        if (tid == 0) {
            const uint32_t arrived = atomicAdd(p.sync.counter, 1u) + 1u;
            if (arrived == gridDim.x) {   // last CTA of this rank: leave the state ready for the next launch
                *p.sync.counter = 0u;
                if (p.sync.epoch_dev != nullptr) *p.sync.epoch_dev = epoch;
            }
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------
static unsigned long long* g_trace = nullptr;   // dev: [16 launches][320 CTAs][8 stamps] device buffer (ggq_dev_set_trace)
static unsigned g_trace_n = 0;

struct Plan {
    Params p;
    int nt, at, grid, nw, occ;
    size_t smem;
};

// One configuration attempt: NW warps per CTA, OCC CTAs per SM.
template <int FMT, bool WIDE = false>
static bool make_plan_cfg(const MmArgs& a, int T, int NW, int OCC, bool allow_slicing, Plan& pl, int S = 1) {
    using G = KGeo<FMT, WIDE>;
    Params& p = pl.p;
    p = Params{};
    p.W = a.W;
    p.X = static_cast<const uint8_t*>(a.X);
    p.outs = make_outs(a);
    p.ldx_bytes = a.ldx * 2;
    p.ldc = a.ldc;
    p.O = a.O;
    p.T = T;
    p.K = static_cast<int>(a.K);
    p.nb = static_cast<int>(a.K / G::QK);
    p.rowB = static_cast<int64_t>(p.nb) * G::BLK;
    p.num_tiles = static_cast<int>(a.W2 ? (a.O + 7) / 8 : (a.O + 15) / 16);   // fused SwiGLU: 8 gate + 8 up rows per tile
    p.nc = (p.nb + G::CHUNK_BLOCKS - 1) / G::CHUNK_BLOCKS;
    pl.nt = T > 8 ? 2 : 1;
    const int tpad = 8 * pl.nt;
    const int sms = num_sms() * OCC;  // CTA slots
    const int SMEM_LIMIT = OCC == 2 ? SMEM_LIMIT_2 : dec::SMEM_LIMIT;
    pl.nw = NW;
    pl.occ = OCC;
    if (a.sync) p.sync = *a.sync;
    p.l2_prefetch_bytes = 0;
    {
        static const int skip = [] { const char* e = getenv("GGQ_DECODE_NOCOMPUTE"); return (e && e[0] == '1') ? 1 : 0; }();
        p.dbg_skip_compute = skip;
    }

    int kw = 1;
    while (kw * 2 <= NW && NW % (kw * 2) == 0 && static_cast<int64_t>(p.num_tiles) * kw < static_cast<int64_t>(sms) * NW &&
           kw * 2 <= p.nc)
        kw *= 2;
    p.KW = kw;
    const int wt = NW / kw;

    constexpr int STAGE_BYTES = (G::CHUNK_BLOCKS / G::PREP_BLOCKS) * 16 * G::SLOT;
    constexpr int SCR_BYTES = 16 * G::PREP_BLOCKS * G::SCRATCH_PER_BLOCK;
    const uint32_t xpad = (FMT == 1) ? 64u : 32u;  // x row pitch = 64 (128-bit loads) / 32 (64-bit loads) mod 128
    auto layout = [&](int cps, int stages, int at, bool commit) -> size_t {
        const size_t elems = static_cast<size_t>(cps) * G::CHUNK_ELEMS;
        const uint32_t xstride = static_cast<uint32_t>(elems * 2 + xpad);
        size_t off = 0;
        auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 127) & ~size_t{127}; return o; };
        const size_t o_bars = take(8 * (1 + MAX_NW * MAX_STAGES + 2 * MAX_NW) + 4 * MAX_NW + 16);  // mbarriers + a flag word per warp + epoch
        const size_t o_x = take(static_cast<size_t>(T) * xstride);
        const size_t o_tbl = take(elems / G::GROUP * tpad * 4);
        const size_t o_ring = take(static_cast<size_t>(NW) * stages * STAGE_BYTES);
        const size_t o_scr = take(static_cast<size_t>(NW) * SCR_BYTES);
        const size_t o_red = take(static_cast<size_t>(NW) * pl.nt * 4 * 32 * 4);  // one partial-sum slot per warp
        const size_t o_mbox = take(S > 1 ? static_cast<size_t>(NW) * pl.nt * 4 * 32 * 4 : 0);  // cluster mailboxes
        (void)at;
        if (commit) {
            p.x_stride = xstride;
            p.off_bars = static_cast<uint32_t>(o_bars);
            p.off_x = static_cast<uint32_t>(o_x);
            p.off_tbl = static_cast<uint32_t>(o_tbl);
            p.off_ring = static_cast<uint32_t>(o_ring);
            p.off_scr = static_cast<uint32_t>(o_scr);
            p.off_red = static_cast<uint32_t>(o_red);
            p.off_mbox = static_cast<uint32_t>(o_mbox);
        }
        return off;
    };

    // whole K in one slice if it fits next to a >= 2-stage ring; otherwise the largest even slicing
    int cps = p.nc, at = 1, stages = 2;
    if (S > 1) {  // cluster split-K: CTA r of a cluster of S owns K-slice r (cps chunks; the last slice may be shorter)
        cps = (p.nc + S - 1) / S;
        if (cps * (S - 1) >= p.nc || OCC != 1 || a.sync) return false;
        if (layout(cps, 2, 1, false) > static_cast<size_t>(SMEM_LIMIT)) return false;
    } else if (layout(cps, 2, 1, false) > static_cast<size_t>(SMEM_LIMIT)) {
        if (!allow_slicing) return false;
        at = 4;
        int slices = 2;
        for (;; ++slices) {
            cps = (p.nc + slices - 1) / slices;
            if (layout(cps, 2, at, false) <= static_cast<size_t>(SMEM_LIMIT)) break;
            if (cps == 1) return false;
        }
    }
    while (stages < MAX_STAGES && layout(cps, stages + 1, at, false) <= static_cast<size_t>(SMEM_LIMIT)) ++stages;
    p.cps = cps;
    p.n_slices = (p.nc + cps - 1) / cps;
    p.stages = stages;
    pl.at = at;
    pl.smem = layout(cps, stages, at, true);
    if (at == 1) {  // flat (tile, chunk) walk: KW is not used, every CTA / cluster owns >= 1 tile
        if (a.sync && a.sync->world > 1 && S == 1) {
            // L2 prefetch budget: at most ~56 MB per GPU in flight ahead of the TMA rings (L2 is 126 MB), 4 KB granules
            static const int mb = [] { const char* e = getenv("GGQ_SYNC_L2_PREFETCH_MB"); return e ? atoi(e) : 56; }();
            const int64_t per_cta = (static_cast<int64_t>(mb) << 20) / std::max(1, std::min(sms, p.num_tiles));
            p.l2_prefetch_bytes = static_cast<int>(std::min<int64_t>(per_cta, int64_t{1} << 22)) & ~4095;
        }
        p.KW = 1;
        pl.grid = S * std::max(1, std::min(sms / S, p.num_tiles));
        p.num_batches = 1;
    } else {
        pl.grid = std::max(1, std::min(sms, (p.num_tiles + wt - 1) / wt));
        const int rounds = (p.num_tiles + pl.grid * wt - 1) / (pl.grid * wt);
        p.num_batches = (rounds + at - 1) / at;
    }
    if (a.ctas_out) *a.ctas_out = pl.grid;
    return true;
}

// Preference order (measured on B200, profiles/README.md): 12 warps with >= 3 ring stages each; else two 8-warp
// CTAs per SM (16 resident warps, 2 stages) when the problem state fits in half an SM's shared memory; else one
// CTA per SM with the full 227 KB (activations of many tokens / K-slicing).
template <int FMT>
static bool make_plan(const MmArgs& a, int T, Plan& pl) {
    static const char* const force = getenv("GGQ_PLAN_FORCE");  // dev: "nw,occ[,cluster]", e.g. "8,2" | "12,1" | "8,1,2"
    if (const char* f = force) {
        const int nw = atoi(f), occ = (strchr(f, ',') ? atoi(strchr(f, ',') + 1) : 1);
        const char* c2 = strchr(f, ',') ? strchr(strchr(f, ',') + 1, ',') : nullptr;
        const int S = c2 ? atoi(c2 + 1) : 1;
        return make_plan_cfg<FMT>(a, T, nw, occ, nw == 8 && occ == 1 && S == 1, pl, S);
    }
    // small problems (a handful of (tile, chunk) items per warp): the weights are on chip before the activations are and
    // the launch is bound by how many warps unpack them — 16 resident warps (two 8-warp CTAs) beat 12 (measured,
    // Q4_K 14336 x 4096 T=1: 11.4 vs 12.2 us; 4096 x 4096: 6.0 vs 6.4 us); large ones stream best with 12 warps x 3 stages
    {
        using G = Geo<FMT>;
        const int64_t nc = (a.K / G::QK + G::CHUNK_BLOCKS - 1) / G::CHUNK_BLOCKS;
        const int64_t items_per_warp = ((a.O + 15) / 16) * nc / (static_cast<int64_t>(num_sms()) * 12);
        if (items_per_warp < 8 && !a.sync && make_plan_cfg<FMT>(a, T, 8, 2, false, pl)) return true;
    }
    if (make_plan_cfg<FMT>(a, T, 12, 1, false, pl) && pl.p.stages >= 3) return true;
    if (make_plan_cfg<FMT>(a, T, 8, 2, false, pl)) return true;
    if (make_plan_cfg<FMT>(a, T, 12, 1, false, pl) && pl.p.stages >= 2) return true;
    if (make_plan_cfg<FMT>(a, T, 8, 1, false, pl) && pl.p.stages >= 2) return true;
    // the activations of all tokens do not fit in one CTA: clusters of 2 / 4 / 8 CTAs split K (and the activations)
    static const bool no_cluster = getenv("GGQ_NO_CLUSTER") != nullptr;
    if (!no_cluster) {
        // smallest padding of K first (S * cps chunks are walked for nc real ones), then the smaller cluster (fewer
        // hops, and small clusters tile the GPCs without leaving SMs idle), then 12 warps before 8
        int best_s = 0, best_nw = 0, best_waste = 1 << 30;
        for (int S = 2; S <= 8; ++S)
            for (int nw = 12; nw >= 8; nw -= 4) {
                Plan t;
                if (!make_plan_cfg<FMT>(a, T, nw, 1, false, t, S) || t.p.stages < 2) continue;
                const int waste = S * t.p.cps * 1024 / t.p.nc;
                if (waste < best_waste) {
                    best_waste = waste;
                    best_s = S;
                    best_nw = nw;
                }
            }
        if (best_s) return make_plan_cfg<FMT>(a, T, best_nw, 1, false, pl, best_s);
    }
    return make_plan_cfg<FMT>(a, T, 8, 1, true, pl);
}

// Q4_K single-token GEMV: the wide-chunk geometry (KGeo<1, true>), 8 warps x 2 stages of 9.2 KB, whenever it fits (10
// warps measured slower: 53.5 vs 49.5 us on the lm_head).  It won on every size measured, 4096 x 4096 (6.0 -> 5.6 us) to
// 128256 x 4096 (51.8 -> 49.4 us), gpurun_out/r2_wide.log / r2_wide2.log; shapes whose activations do not fit next to
// the wide rings (K = 28672) fall back to the fine-grained geometry.
static bool make_plan_wide(const MmArgs& s, int T, Plan& pl) {
    static const bool no_gemv = getenv("GGQ_NO_GEMV") != nullptr;
    static const bool off = [] { const char* e = getenv("GGQ_WIDE"); return e && e[0] == '0'; }();   // dev: GGQ_WIDE=0
    if (T != 1 || no_gemv || off) return false;
    return make_plan_cfg<1, true>(s, 1, 8, 1, false, pl) && pl.p.stages >= 2 && pl.at == 1 && pl.p.n_slices == 1;
}

template <int FMT, int NT, int AT, int NW, int MINB, bool GV = false, bool DUAL = false, bool WIDE = false>
static int launch_kernel(const Plan& pl, cudaStream_t stream, const uint8_t* W2 = nullptr) {
    auto kern = decode_kernel<FMT, NT, AT, NW, MINB, GV, DUAL, WIDE>;
    static int configured_dev_mask[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured_dev_mask[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             MINB == 2 ? SMEM_LIMIT_2 : SMEM_LIMIT);
        if (e != cudaSuccess) return static_cast<int>(e);
        configured_dev_mask[dev] = 1;
    }
    // the packed rows viewed as int32 [O, rowB / 4]; box = 16 rows (fused SwiGLU: 8 rows of each matrix) x SLOT bytes.
    // Encoding a map costs ~1 us of host time, as much as the rest of the call: tma.cuh keeps the last few per thread.
    auto map_of = [&](const uint8_t* w) {
        return cached_map_2d(CU_TENSOR_MAP_DATA_TYPE_INT32, w, static_cast<uint64_t>(pl.p.rowB / 4), static_cast<uint64_t>(pl.p.O),
                             static_cast<uint64_t>(pl.p.rowB), Geo<FMT>::SLOT / 4, DUAL ? 8 : 16, CU_TENSOR_MAP_SWIZZLE_NONE, dev);
    };
    const CUtensorMap* m1 = map_of(pl.p.W);
    const CUtensorMap* m2 = DUAL ? map_of(W2) : m1;
    if (!m1 || !m2) return static_cast<int>(cudaErrorInvalidValue);
    const CUtensorMap& map_w = *m1;
    const CUtensorMap& map_w2 = *m2;
    static const bool no_pdl = getenv("GGQ_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pl.grid);
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    const int S = (AT == 1) ? pl.p.n_slices : 1;
    if (S > 1) {  // cluster split-K: as many clusters as can be co-resident (the kernel is persistent)
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = S;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
        cfg.attrs = attr;
        cfg.numAttrs = na;
        int max_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters < 1) {
            cudaGetLastError();
            return GGQ_E_FAMILY;
        }
        cfg.gridDim = dim3(S * std::max(1, std::min(max_clusters, pl.grid / S)));
    }
    if (!no_pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    Params prm = pl.p;
    if (g_trace != nullptr) prm.trace = g_trace + static_cast<size_t>(g_trace_n++ % 16) * (320 * 8);
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, map_w, map_w2, prm);
    count_launch();
    return static_cast<int>(e != cudaSuccess ? e : cudaGetLastError());
}

template <int FMT>
static int launch_fmt(const MmArgs& a) {
    if (a.sync && a.T > 8) return GGQ_E_FAMILY;  // the fused exchange is a single-pass, one-n-tile feature
    for (int64_t t0 = 0; t0 < a.T; t0 += 16) {  // T > 16: 16-token passes (weights re-read per pass)
        MmArgs s = a;
        s.X = static_cast<const __half*>(a.X) + t0 * a.ldx;
        for (int i = 0; i < a.n_out; ++i) s.C[i] = static_cast<__half*>(a.C[i]) + t0 * a.ldc;
        const int T = static_cast<int>(std::min<int64_t>(16, a.T - t0));
        Plan pl;
        static const bool no_gemv = getenv("GGQ_NO_GEMV") != nullptr;
        if constexpr (FMT == 1) {
            if (make_plan_wide(s, T, pl)) {
                const int rcw = launch_kernel<1, 1, 1, 8, 1, true, false, true>(pl, a.stream);
                if (rcw != 0) return rcw;
                continue;
            }
        }
        if (!make_plan<FMT>(s, T, pl)) return GGQ_E_FAMILY;
        int rc;
        if (T == 1 && pl.at == 1 && pl.p.n_slices == 1 && !no_gemv) {  // single token: GEMV tile code
            rc = pl.occ == 2   ? launch_kernel<FMT, 1, 1, 8, 2, true>(pl, a.stream)
                 : pl.nw == 12 ? launch_kernel<FMT, 1, 1, 12, 1, true>(pl, a.stream)
                               : launch_kernel<FMT, 1, 1, 8, 1, true>(pl, a.stream);
        } else if (pl.occ == 2) {  // AT == 1 by construction
            rc = pl.nt == 1 ? launch_kernel<FMT, 1, 1, 8, 2>(pl, a.stream) : launch_kernel<FMT, 2, 1, 8, 2>(pl, a.stream);
        } else if (pl.nw == 12) {
            rc = pl.nt == 1 ? launch_kernel<FMT, 1, 1, 12, 1>(pl, a.stream) : launch_kernel<FMT, 2, 1, 12, 1>(pl, a.stream);
        } else if (pl.nt == 1) {
            rc = pl.at == 1 ? launch_kernel<FMT, 1, 1, 8, 1>(pl, a.stream) : launch_kernel<FMT, 1, 4, 8, 1>(pl, a.stream);
        } else {
            rc = pl.at == 1 ? launch_kernel<FMT, 2, 1, 8, 1>(pl, a.stream) : launch_kernel<FMT, 2, 4, 8, 1>(pl, a.stream);
        }
        if (rc != 0) return rc;
    }
    return 0;
}

}  // namespace dec
}  // namespace ggq
