// skinny.cu — HBM-bound skinny GEMM for 2 <= T <= 128 tokens on tcgen05:  C[T, O] = X[T, K] . dequant(W)[O, K]^T
//
// The weights are the M side of the MMA and never touch shared memory after decoding: a 128-row tile of packed blocks
// is TMA-staged (verbatim GGUF bytes, one 2-D box of WU "units" of each row, a unit = UNIT_K consecutive weights), every
// dequant thread owns ONE weight row = ONE tensor-memory lane, decodes the row's unit in registers (skinny_tile.cuh) and
// writes it with tcgen05.st into one of NBUF A-operand buffers in TMEM.  The activations are the N side: [N tokens x
// UNIT_K] fp16 tiles, TMA-staged with the 128-byte swizzle, read by the tensor core straight from shared memory
// (B operand).  tcgen05.mma (M128, N = 16/32/64/128, K16; A from TMEM, B from smem) accumulates a row tile over K in
// TMEM.  Shared-memory traffic per weight is the packed bytes (once in, once out) plus 2N/128 bytes of activations —
// the mma.sync decode kernel re-reads T*2/16 bytes of activations per weight and is shared-memory-bound from T = 8 on.
// The cost per weight is independent of T up to the MMA time (N / 2 cycles per 128 x 16 weights).
//
// Warp roles (4G + 6 warps, one CTA per SM, persistent; a "step" = one unit of each of the 128 rows):
//   warps 0..4G-1     dequant   G ping-pong groups of 128 threads; group g handles the steps g, g + G, ... (each in its
//                               own A buffer), so a group has G step periods for one step and the hand-over latencies
//                               (barrier waits, tcgen05.st completion) of one group hide behind the others
//   warps 4G..4G+3    epilogue  tcgen05.ld 32x32b.x16 -> fp16 -> C (or fp32 partial sums -> workspace, see below)
//   warp  4G+4        producer  TMA: packed boxes [128 rows x ROW_BYTES] into a ring of `depth` slots, X tiles (one 3-D
//                               box per step) into a ring of `xst` stages
//   warp  4G+5        MMA       per step UNIT_K/16 x tcgen05.mma; tcgen05.commit frees the A buffer / X stage and
//                               publishes the accumulator of a finished tile part
// The single-thread roles sit in the HIGHEST warp ids (the warp scheduler favours them: in warps 0 / 1 they were starved
// by the dequant warps) and run warp-uniform loops with one elected lane issuing, so their operands stay in uniform
// registers (inside `if (lane == 0)` every tcgen05.mma cost an ELECT + 4 x R2UR + branch sequence).
// mbarriers: w_full/w_empty (ring), x_full/x_free, a_full/a_free (A buffers), acc_full/acc_free (2 accumulators).
//
// Work split: items = (row tile, box) pairs in tile-major order, cut into gridDim.x equal contiguous ranges, so every
// CTA streams the same number of bytes whatever O and K are.  A tile cut by a range boundary is finished by the CTA that
// holds its box 0: the CTAs holding the rest (always the FIRST thing in their range) store fp32 partial sums into
// their workspace slot and raise a flag; the finisher adds them in CTA order (deterministic).  All CTAs are co-resident
// (grid <= SM count), so the finisher's wait cannot deadlock.
// Launches use programmatic stream serialization: weights are prefetched before griddepcontrol.wait, activations are
// read and global memory is written only after it.
// HBM traffic = packed weight bytes once (+ activations from L2 once per row tile); roofline: HBM bandwidth.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "../../include/ggq.h"
#include "common.cuh"
#include "formats.cuh"
#include "prefill_tile.cuh"
#include "ptx.cuh"
#include "skinny_tile.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace ggq {
namespace skn {

using pre::Unit;

constexpr int TM = 128;            // weight rows per tile = UMMA M = TMEM lanes
constexpr int EP_WARPS = 4;
constexpr int MAX_DEPTH = 12, MAX_XST = 12, MAX_NBUF = 4;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int WS_POOL = 4;         // workspaces used round-robin by consecutive launches
constexpr int MAX_N = 128;

struct Params {
    OutPtrs outs;
    int64_t ldc, O, items;
    int T, B;               // tokens; TMA boxes per row (WU steps each)
    int depth, xst;         // ring depths
    int probe;              // dev (GGQ_SKINNY_PROBE): 1 = no dequant / tcgen05.st, 2 = also no MMA, 3 = also no X loads
    float* ws;              // [grid][N][128] fp32 partial sums of tile parts that do not hold box 0
    uint32_t* flags;        // [grid][4] one word per (CTA, lane quarter): 1 = the partial sums are in the workspace
    long long* prof;        // dev (GGQ_SKINNY_PROF=1): clock64 stamps of CTA 0, [64 steps][16 events]
};
#ifdef GGQ_SKINNY_PROF_BUILD   // dev builds only: the stamps cost ~1.5 % of the dequant loop's issue slots
#define SKN_STAMP(j, ev) \
    do { if (p.prof != nullptr && blockIdx.x == 0 && (j) < 64) p.prof[(j) * 16 + (ev)] = clock64(); } while (0)
#else
#define SKN_STAMP(j, ev) do { } while (0)
#endif

constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }

// FMT: quant format; N: MMA N (tokens, padded); WU: units (prefill_tile.cuh Unit<FMT>) per TMA box row; G: dequant groups;
// MAXT: row tiles per pass (they share every activation stage)
template <int FMT, int N, int WU, int G, int MAXT> struct Cfg {
    using U = Unit<FMT>;
    static constexpr int UNIT_K = U::UNIT_K;
    static constexpr int A_COLS = UNIT_K / 2;            // TMEM columns of one A buffer (2 fp16 per column)
    static constexpr int KSTEPS = UNIT_K / 16;           // MMAs per step
    static constexpr int XATOMS = UNIT_K / 64;           // 128-byte swizzle atoms per step
    static constexpr int RAW = WU * U::UNIT_BYTES;       // packed bytes of one box row
    static constexpr int MAX_OFF = 16 - cgcd(RAW, 16);   // the box starts at the 16-byte aligned superset
    // + 16 for 288-byte Q4_K rows: a 304-byte pitch keeps the threads' 128-bit loads (thread = row) conflict free
    static constexpr int ROW_BYTES = ((RAW + MAX_OFF + 15) & ~15) + ((FMT == 1 && WU == 2) ? 16 : 0);
    static constexpr int W_SLOT = TM * ROW_BYTES;        // bytes of a packed staging slot
    static constexpr int X_ATOM = N * 128;               // bytes of one [N tokens x 64 k] atom
    static constexpr int X_STAGE = WU * XATOMS * X_ATOM; // activations of one box column (WU steps), all N tokens
    static constexpr int CHUNKS = UNIT_K / 64;           // 64-weight dequant chunks per step and row
    // TMEM columns: two sets of MAXT accumulators (a pass computes into one while the epilogue drains the other),
    // then the A buffers
    static constexpr int ACC_SET = MAXT * N;
    static constexpr int A_COL0 = 2 * ACC_SET;
    static constexpr int NBUF = (512 - A_COL0) / A_COLS < MAX_NBUF ? (512 - A_COL0) / A_COLS : MAX_NBUF;
    static constexpr int DQ_WARPS = 4 * G;
    static constexpr int EP_WARP0 = DQ_WARPS, P_WARP = DQ_WARPS + EP_WARPS, MMA_WARP = P_WARP + 1;
    static constexpr int NUM_THREADS = (MMA_WARP + 1) * 32;
    static_assert(NBUF >= G, "every dequant group needs its own A buffer");
    static_assert(A_COL0 + NBUF * A_COLS <= 512, "tensor memory");
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ bool elect_one() {  // one lane of the (converged) warp
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// A ring position without divisions: index and the parity of the number of completed laps.
struct RingPos {
    uint32_t idx, lap;
    __device__ __forceinline__ void advance(uint32_t size) {
        if (++idx == size) {
            idx = 0;
            lap ^= 1u;
        }
    }
};

// The cells (row tile, box column) of one CTA, walked in the order every role uses: the CTA's tiles (a contiguous range
// of the tile-major item order: the first tile may start at box b_first, the last may end before box b_last) are taken
// in passes of MAXT tiles; inside a pass the box columns are the OUTER loop, so the MAXT tiles of a pass share each
// activation stage (one X stage per (pass, column) instead of one per (tile, column)).  All members are warp-uniform;
// everything is incremental (no divisions: the single-warp roles are latency-bound on exactly this code).
template <int MAXT> struct Cells {
    uint32_t ntiles, b_first, b_last, Bb;   // region: tiles, first box of tile 0, end box (exclusive) of the last tile
    uint32_t pt0, nt, kb, ti, t_lo, t_hi;   // pass base / size, column, tile inside the pass, tiles present at this column
    uint32_t kmax;                          // last column of this pass that has cells
    uint32_t pass_par;                      // pass index & 1
    bool has_first, has_last;               // the pass holds the region's first / last tile
    bool done;
    __device__ __forceinline__ void start_pass() {
        nt = min(static_cast<uint32_t>(MAXT), ntiles - pt0);
        has_first = pt0 == 0;
        has_last = pt0 + nt == ntiles;
        kmax = (has_last && nt == 1) ? b_last - 1 : Bb - 1;
    }
    __device__ __forceinline__ void seek() {  // first column at or after kb (possibly in a later pass) that has cells
        for (;;) {
            t_lo = (has_first && kb < b_first) ? 1u : 0u;
            t_hi = (has_last && kb >= b_last) ? nt - 1 : nt;
            if (t_lo < t_hi) {
                ti = t_lo;
                return;
            }
            if (!next_column()) return;
        }
    }
    __device__ __forceinline__ bool next_column() {  // false: no more columns
        if (++kb == Bb) {
            kb = 0;
            pt0 += MAXT;
            pass_par ^= 1u;
            if (pt0 >= ntiles) {
                done = true;
                return false;
            }
            start_pass();
        }
        return true;
    }
    __device__ __forceinline__ void init(uint32_t ntiles_, uint32_t b_first_, uint32_t b_last_, uint32_t Bb_) {
        ntiles = ntiles_; b_first = b_first_; b_last = b_last_; Bb = Bb_;
        pt0 = 0; kb = 0; ti = 0; t_lo = 0; t_hi = 0; pass_par = 0; nt = 0; kmax = 0; has_first = has_last = false;
        done = ntiles == 0;
        if (!done) {
            start_pass();
            seek();
        }
    }
    // next cell; returns true when it is the first cell of a new stage (column)
    __device__ __forceinline__ bool next() {
        if (++ti < t_hi) return false;
        next_stage();
        return true;
    }
    __device__ __forceinline__ void next_stage() {  // first cell of the next (pass, column) that has cells
        if (next_column()) seek();
    }
    __device__ __forceinline__ uint32_t tile_rel() const { return pt0 + ti; }
    __device__ __forceinline__ bool first_of_stage() const { return ti == t_lo; }
    __device__ __forceinline__ bool last_of_stage() const { return ti + 1 == t_hi; }
    __device__ __forceinline__ bool last_of_pass() const { return ti + 1 == t_hi && kb == kmax; }
    __device__ __forceinline__ bool first_of_tile() const { return kb == ((has_first && ti == 0) ? b_first : 0u); }
};

template <int FMT, int N, int WU, int G, int MAXT>
__global__ void __launch_bounds__((Cfg<FMT, N, WU, G, MAXT>::NUM_THREADS), 1)
skinny_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const Params p) {
    using C = Cfg<FMT, N, WU, G, MAXT>;
    using U = Unit<FMT>;
    extern __shared__ uint8_t smem_raw[];
    // align through the shared-window offset, not through a generic-pointer cast: the loads below stay LDS
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* xring = smem;                                        // [xst][X_STAGE], 1024-byte aligned atoms
    uint8_t* wring = smem + p.xst * C::X_STAGE;                   // [depth][128 rows][ROW_BYTES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(wring + p.depth * C::W_SLOT);
    uint64_t* w_full = bars;                     // [MAX_DEPTH] TMA tx
    uint64_t* w_empty = w_full + MAX_DEPTH;      // [MAX_DEPTH] every dequant warp
    uint64_t* x_full = w_empty + MAX_DEPTH;      // [MAX_XST]   TMA tx
    uint64_t* x_free = x_full + MAX_XST;         // [MAX_XST]   tcgen05.commit
    uint64_t* a_full = x_free + MAX_XST;         // [MAX_NBUF]  the 4 warps of a group
    uint64_t* a_free = a_full + MAX_NBUF;        // [MAX_NBUF]  tcgen05.commit
    uint64_t* acc_full = a_free + MAX_NBUF;      // [2]         tcgen05.commit
    uint64_t* acc_free = acc_full + 2;           // [2]         4 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ibeg = static_cast<int64_t>(blockIdx.x) * p.items / gridDim.x;
    const int64_t iend = static_cast<int64_t>(blockIdx.x + 1) * p.items / gridDim.x;
    const uint32_t Bb = static_cast<uint32_t>(p.B);
    // region of this CTA: tiles tile0 .. tile0 + ntiles - 1; tile 0 from box b_first, the last tile up to box b_last
    const uint32_t tile0 = static_cast<uint32_t>(ibeg / Bb), b_first = static_cast<uint32_t>(ibeg % Bb);
    const uint32_t tile_l = iend > ibeg ? static_cast<uint32_t>((iend - 1) / Bb) : tile0;
    const uint32_t ntiles = iend > ibeg ? tile_l - tile0 + 1 : 0u;
    const uint32_t b_last = iend > ibeg ? static_cast<uint32_t>((iend - 1) % Bb) + 1 : 0u;

    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_DEPTH; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&w_empty[s], C::DQ_WARPS);
        }
        for (int s = 0; s < MAX_XST; ++s) {
            mbar_init(&x_full[s], 1);
            mbar_init(&x_free[s], 1);
        }
        for (int s = 0; s < MAX_NBUF; ++s) {
            mbar_init(&a_full[s], 4);
            mbar_init(&a_free[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_free[s], EP_WARPS);
        }
        fence_mbar_init();
        prefetch_tmap(&map_w);
        prefetch_tmap(&map_x);
    }
    if (warp == C::MMA_WARP) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == C::P_WARP) {
        // ================= producer: W ring and X ring (warp-uniform loop, one elected lane issues) =================
        Cells<MAXT> cw, cx;   // W cursor (cells), X cursor (stages)
        cw.init(ntiles, b_first, b_last, Bb);
        cx.init(ntiles, b_first, b_last, Bb);
        RingPos wr{0, 0}, xr{0, 0};
        uint32_t wcount = 0, xcount = 0;
        auto issue_w = [&]() {
            if (elect_one()) {
                mbar_arrive_expect_tx(&w_full[wr.idx], C::W_SLOT);
                tma_load_2d(wring + wr.idx * C::W_SLOT, &map_w, static_cast<int>(((cw.kb * C::RAW) & ~15u) >> 2),
                            static_cast<int>((tile0 + cw.tile_rel()) * TM), &w_full[wr.idx]);
            }
            __syncwarp();
            if (lane == 0) SKN_STAMP(wcount, 0);
            ++wcount;
            wr.advance(p.depth);
            cw.next();
        };
        auto issue_x = [&]() {
            if (elect_one()) {
                mbar_arrive_expect_tx(&x_full[xr.idx], C::X_STAGE);
                tma_load_3d(xring + xr.idx * C::X_STAGE, &map_x, 0, 0, static_cast<int>(cx.kb * WU * C::XATOMS), &x_full[xr.idx]);
            }
            __syncwarp();
            ++xcount;
            xr.advance(p.xst);
            cx.next_stage();
        };
        // weights first (nothing in this library writes them): fill the ring before the dependency wait
        while (!cw.done && wcount < static_cast<uint32_t>(p.depth)) issue_w();
        pdl_wait();  // the activations may be produced by the previous kernel of the stream
        if (p.probe >= 3) cx.done = true;  // dev: no activation loads at all
        while (!cx.done && xcount < static_cast<uint32_t>(p.xst)) issue_x();
        while (!cw.done || !cx.done) {
            // test_wait, not try_wait: a try_wait on the ring that is not ready may suspend this warp for microseconds
            // while the other ring starves.  (lap ^ 1 = parity of the previous use of the slot)
            bool any = false;
            if (!cx.done && mbar_test_wait(&x_free[xr.idx], xr.lap ^ 1u)) { issue_x(); any = true; }
            if (!cw.done && mbar_test_wait(&w_empty[wr.idx], wr.lap ^ 1u)) { issue_w(); any = true; }
            if (!any) __nanosleep(64);   // both rings full: leave the issue slots to the dequant warps
        }
    } else if (warp == C::MMA_WARP) {
        // ================= MMA issuer =================
        constexpr uint32_t IDESC = umma_idesc_f16(TM, N);
        Cells<MAXT> c;
        c.init(ntiles, b_first, b_last, Bb);
        RingPos ar{0, 0}, xr{0, 0};
        uint32_t step = 0, passes = 0;
        bool new_stage = true;
        while (!c.done) {
            const uint32_t accs = c.pass_par;
            if (new_stage && p.probe < 3) mbar_wait(&x_full[xr.idx], xr.lap);
            const uint32_t d_tmem = tmem_base + accs * C::ACC_SET + c.ti * N;
            const bool tile_start = c.first_of_tile(), stage_end = c.last_of_stage(), pass_end = c.last_of_pass();
            const uint32_t x_cell = smem_u32(xring + xr.idx * C::X_STAGE);
#pragma unroll
            for (int s = 0; s < WU; ++s, ++step) {
                if (lane == 0) SKN_STAMP(step, 1);
                mbar_wait(&a_full[ar.idx], ar.lap);
                tc_fence_after();
                if (lane == 0) SKN_STAMP(step, 2);
                const uint32_t a_tmem = tmem_base + C::A_COL0 + ar.idx * C::A_COLS;
                // K-major SWIZZLE_128B descriptor of the step's first atom; the k-steps below only add to its address field
                const uint64_t bd0 = smem_desc_sw128(x_cell + s * (C::XATOMS * C::X_ATOM));
                if (elect_one()) {
                    if (p.probe < 2) {
#pragma unroll
                        for (int kk = 0; kk < C::KSTEPS; ++kk) {
                            const uint64_t bd = bd0 + static_cast<uint64_t>(((kk >> 2) * C::X_ATOM + (kk & 3) * 32) >> 4);
                            umma_f16_ts(d_tmem, a_tmem + kk * 8, bd, IDESC, (tile_start && s == 0 && kk == 0) ? 0u : 1u);
                        }
                    }
                    umma_commit(&a_free[ar.idx]);
                    if (s == WU - 1) {
                        if (stage_end) umma_commit(&x_free[xr.idx]);
                        if (pass_end) umma_commit(&acc_full[accs]);
                    }
                }
                __syncwarp();
                if (lane == 0) SKN_STAMP(step, 3);
                ar.advance(C::NBUF);
            }
            if (stage_end) xr.advance(p.xst);
            new_stage = c.next();
            if (pass_end) {
                ++passes;
                // the set the next pass computes into was last used by pass `passes - 2`: wait until the epilogue drained it
                if (!c.done && passes >= 2) mbar_wait(&acc_free[c.pass_par], ((passes >> 1) - 1) & 1u);
            }
        }
    } else if (warp < C::DQ_WARPS) {
        // ================= dequant groups: thread = weight row = TMEM lane =================
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int grp = warp >> 2;                     // ping-pong group
        const int row = q * 32 + lane;                 // row inside the tile
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        const bool stamp = (warp == 0 && lane == 0);
        Cells<MAXT> c;
        c.init(ntiles, b_first, b_last, Bb);
        RingPos wr{0, 0}, ar{0, 0};
        uint32_t step = 0, sg = 0;    // step counter (profiling only), step % G
        for (; !c.done; c.next()) {
            // EVERY dequant warp observes EVERY phase of the ring barriers (waits for the box, arrives on w_empty), also
            // for boxes whose steps all belong to other groups: a parity wait that skips a phase can return early (the
            // phase before the skipped one has the parity it is looking for) or never (two phases later it has again).
            mbar_wait(&w_full[wr.idx], wr.lap);
#pragma unroll
            for (uint32_t s = 0; s < WU; ++s, ++step) {
                const bool mine = sg == static_cast<uint32_t>(grp);
                if (++sg == G) sg = 0;
                const uint32_t ab = ar.idx, alap = ar.lap;
                ar.advance(C::NBUF);
                if (!mine) continue;
                if (stamp) SKN_STAMP(step, 5);
                if (step >= C::NBUF) {
                    mbar_wait_backoff(&a_free[ab], alap ^ 1u, 32);  // the MMAs of step - NBUF have read this buffer
                    tc_fence_after();
                }
                if (stamp) SKN_STAMP(step, 6);
                if (p.probe == 0) {
                    const uint8_t* brow = wring + wr.idx * C::W_SLOT + row * C::ROW_BYTES;
                    // Q4_K blocks are 16-byte aligned: point at the block; Q8_0 / Q6_K: offset inside the aligned superset
                    const uint8_t* up = FMT == 1 ? brow + s * U::UNIT_BYTES : brow;
                    const int off = FMT == 1 ? 0 : static_cast<int>(((c.kb * C::RAW) & 15u) + s * U::UNIT_BYTES);
                    const uint32_t a_tmem = tmem_base + lane_base + C::A_COL0 + ab * C::A_COLS;
#pragma unroll
                    for (int kb = 0; kb < C::CHUNKS; ++kb) {
                        uint4 v[8];
                        dq64<FMT>(up, off, kb, v);
                        tmem_st32(a_tmem + kb * 32, v);
                    }
                    if (stamp) SKN_STAMP(step, 7);
                    tmem_st_wait();
                }
                if (stamp) SKN_STAMP(step, 8);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[ab]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&w_empty[wr.idx]);
            wr.advance(p.depth);
        }
    } else {
        // ================= epilogue warps =================
        const int q = warp & 3;
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        pdl_wait();  // C (and the workspace) may still be in use by earlier kernels of the stream
        for (uint32_t pt0 = 0, pass = 0; pt0 < ntiles; pt0 += MAXT, ++pass) {
            const uint32_t nt = min(static_cast<uint32_t>(MAXT), ntiles - pt0), accs = pass & 1u;
            mbar_wait_backoff(&acc_full[accs], (pass >> 1) & 1u, 256);  // a whole pass away: do not spin on issue slots
            tc_fence_after();
            for (uint32_t ti = 0; ti < nt; ++ti) {
                const uint32_t trel = pt0 + ti, tile = tile0 + trel;
                const bool head = (trel != 0 || b_first == 0);                       // this CTA holds box 0 of the tile
                const bool whole = head && (trel != ntiles - 1 || b_last == Bb);     // ... and all the others
                const int64_t tile_end = static_cast<int64_t>(tile + 1) * Bb;
                const int64_t grow = static_cast<int64_t>(tile) * TM + q * 32 + lane;
                // the CTAs after this one that hold the rest of the tile (each stored it first thing)
                int nfol = 0;
                if (head && !whole) {
                    for (int c = static_cast<int>(blockIdx.x) + 1; c < static_cast<int>(gridDim.x); ++c) {
                        const int64_t cb = static_cast<int64_t>(c) * p.items / gridDim.x;
                        if (cb >= tile_end) break;
                        ++nfol;
                        if (cb == static_cast<int64_t>(c + 1) * p.items / gridDim.x) continue;  // empty range
                        if (lane == 0)
                            while (ld_acquire_gpu(p.flags + c * 4 + q) == 0u) {
                            }
                    }
                    __syncwarp();
                }
#pragma unroll 1
                for (int cg = 0; cg < N / 16; ++cg) {
                    if (cg * 16 >= p.T) break;
                    uint32_t r[16];
                    tmem_ld16(tmem_base + lane_base + accs * C::ACC_SET + ti * N + cg * 16, r);
                    tmem_ld_wait();
                    if (!head) {
                        float* dst = p.ws + (static_cast<size_t>(blockIdx.x) * N + cg * 16) * TM + q * 32 + lane;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (cg * 16 + i < p.T) __stcg(dst + i * TM, __uint_as_float(r[i]));
                        continue;
                    }
                    for (int f = 1; f <= nfol; ++f) {
                        const int c = static_cast<int>(blockIdx.x) + f;
                        if (static_cast<int64_t>(c) * p.items / gridDim.x == static_cast<int64_t>(c + 1) * p.items / gridDim.x) continue;
                        const float* src = p.ws + (static_cast<size_t>(c) * N + cg * 16) * TM + q * 32 + lane;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (cg * 16 + i < p.T) r[i] = __float_as_uint(__uint_as_float(r[i]) + __ldcg(src + i * TM));
                    }
                    if (grow < p.O) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int t = cg * 16 + i;
                            if (t < p.T) {
                                const __half h = __float2half_rn(__uint_as_float(r[i]));
                                const int64_t at = static_cast<int64_t>(t) * p.ldc + grow;
                                p.outs.p[0][at] = h;
                                if (p.outs.n > 1) {   // uniform branch: the common single-output launch issues one store
#pragma unroll
                                    for (int o = 1; o < 8; ++o)
                                        if (o < p.outs.n) p.outs.p[o][at] = h;
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    if (!head) {
                        __threadfence();
                        st_release_gpu(p.flags + blockIdx.x * 4 + q, 1u);
                    } else {
                        for (int f = 1; f <= nfol; ++f) p.flags[(blockIdx.x + f) * 4 + q] = 0u;  // consumed: ready for the next launch
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[accs]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == C::MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---- host ---------------------------------------------------------------------------------------
struct DevWs {
    float* data[WS_POOL];
    uint32_t* flags[WS_POOL];
    int sms;
    bool ready;
};
static DevWs g_ws[64];
static std::mutex g_ws_mu;
static std::atomic<unsigned> g_ws_rr{0};

// Workspace of the tile parts that cross CTA ranges; allocated on the first call per device (not capturable: make
// one call outside stream capture first — any warm-up does).
static int get_ws(int dev, float** data, uint32_t** flags) {
    if (dev < 0 || dev >= 64) return static_cast<int>(cudaErrorInvalidDevice);
    std::lock_guard<std::mutex> lock(g_ws_mu);
    DevWs& w = g_ws[dev];
    if (!w.ready) {
        w.sms = num_sms();
        const size_t per = static_cast<size_t>(w.sms) * MAX_N * TM * sizeof(float);
        const size_t fl = static_cast<size_t>(w.sms) * 4 * sizeof(uint32_t);
        uint8_t* base = nullptr;
        cudaError_t e = cudaMalloc(&base, WS_POOL * (per + fl));
        if (e != cudaSuccess) return static_cast<int>(e);
        e = cudaMemset(base, 0, WS_POOL * (per + fl));
        if (e != cudaSuccess) return static_cast<int>(e);
        for (int i = 0; i < WS_POOL; ++i) {
            w.data[i] = reinterpret_cast<float*>(base + i * per);
            w.flags[i] = reinterpret_cast<uint32_t*>(base + WS_POOL * per + i * fl);
        }
        w.ready = true;
    }
    const unsigned k = g_ws_rr.fetch_add(1, std::memory_order_relaxed) % WS_POOL;
    *data = w.data[k];
    *flags = w.flags[k];
    return 0;
}

template <int FMT, int N, int WU, int G, int MAXT>
static int launch_n(const MmArgs& a) {
    using C = Cfg<FMT, N, WU, G, MAXT>;
    using U = Unit<FMT>;
    int dev = 0;
    cudaGetDevice(&dev);
    const int64_t rowB = a.K / U::QK * U::BLK;
    Params p;
    p.outs = make_outs(a);
    p.ldc = a.ldc;
    p.O = a.O;
    p.T = static_cast<int>(a.T);
    const int units = static_cast<int>(a.K / C::UNIT_K);
    if (units % WU != 0) return GGQ_E_FAMILY;
    p.B = units / WU;
    const int64_t tiles = (a.O + TM - 1) / TM;
    p.items = tiles * p.B;
    static const int probe = [] { const char* e = getenv("GGQ_SKINNY_PROBE"); return e ? atoi(e) : 0; }();
    p.probe = probe;
    static const int xst_env = [] { const char* e = getenv("GGQ_SKINNY_XST"); return e ? atoi(e) : 0; }();  // dev
    p.xst = std::min(MAX_XST, std::max(2, (48 * 1024) / C::X_STAGE));
    if (xst_env) p.xst = std::min(MAX_XST, std::max(2, xst_env));
    const int budget = SMEM_LIMIT - 1024 - p.xst * C::X_STAGE - 512;
    p.depth = std::min(MAX_DEPTH, budget / C::W_SLOT);
    if (p.depth < 2) return GGQ_E_FAMILY;
    const size_t smem = 1024 + static_cast<size_t>(p.xst) * C::X_STAGE + static_cast<size_t>(p.depth) * C::W_SLOT + 512;
    const int grid = static_cast<int>(std::min<int64_t>(num_sms(), p.items));
    int rc = get_ws(dev, &p.ws, &p.flags);
    if (rc != 0) return rc;
    static const bool prof_on = [] { const char* e = getenv("GGQ_SKINNY_PROF"); return e && e[0] == '1'; }();
    p.prof = nullptr;
    if (prof_on) {
        cudaMalloc(&p.prof, 64 * 16 * sizeof(long long));
        cudaMemset(p.prof, 0, 64 * 16 * sizeof(long long));
    }

    auto kern = skinny_kernel<FMT, N, WU, G, MAXT>;
    static int configured[64] = {0};
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) return static_cast<int>(e);
        configured[dev] = 1;
    }
    // X: fp16 [T, K] as (64 k, T tokens, K/64 groups); box (64, N, XATOMS): rows >= T are zero-filled
    const CUtensorMap* map_x = cached_map_x3d(a.X, static_cast<uint64_t>(a.K), static_cast<uint64_t>(a.T),
                                              static_cast<uint64_t>(a.ldx), N, WU * C::XATOMS, dev);
    // W: the packed rows viewed as int32 [O, rowB/4], box ROW_BYTES/4 x 128 rows; rows >= O and bytes past the row end
    // are zero-filled
    const CUtensorMap* map_w = cached_map_2d(CU_TENSOR_MAP_DATA_TYPE_INT32, a.W, static_cast<uint64_t>(rowB / 4),
                                             static_cast<uint64_t>(a.O), static_cast<uint64_t>(rowB), C::ROW_BYTES / 4, TM,
                                             CU_TENSOR_MAP_SWIZZLE_NONE, dev);
    if (!map_x || !map_w) return static_cast<int>(cudaErrorInvalidValue);

    static const bool no_pdl = getenv("GGQ_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(C::NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = a.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = no_pdl ? 0 : 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, *map_x, *map_w, p);
    count_launch();
    if (prof_on) {  // dev: dump CTA 0's time stamps (cycles relative to the first W issue)
        static long long host[64 * 16];
        cudaStreamSynchronize(a.stream);
        cudaMemcpy(host, p.prof, sizeof(host), cudaMemcpyDeviceToHost);
        cudaFree(p.prof);
        fprintf(stderr, "skinny prof fmt=%d N=%d WU=%d G=%d MAXT=%d NBUF=%d depth=%d xst=%d items/cta=%lld\n", FMT, N, WU, G,
                MAXT, C::NBUF, p.depth, p.xst, static_cast<long long>(p.items / grid));
        fprintf(stderr, "step  W_issued | mma: x_ok a_ok issued - | dq(warp 0): w_ok a_free_ok st_issued st_done\n");
        for (int j = 0; j < 64; ++j) {
            fprintf(stderr, "%3d", j);
            for (int ev = 0; ev < 9; ++ev) fprintf(stderr, " %8lld", host[j * 16 + ev] ? host[j * 16 + ev] - host[0] : -1LL);
            fprintf(stderr, "\n");
        }
    }
    return static_cast<int>(e != cudaSuccess ? e : cudaGetLastError());
}

// Tokens -> MMA N; the smaller N, the more row tiles share an activation stage (MAXT * N accumulator columns per set).
// Two units per TMA box (longer box rows, half as many boxes) whenever the row has an even number of units.
template <int FMT>
static int launch_fmt(const MmArgs& a) {
    static const int wu_env = [] { const char* e = getenv("GGQ_SKINNY_WU"); return e ? atoi(e) : 0; }();  // dev: 1 = one unit per box
    const int units = static_cast<int>(a.K / Unit<FMT>::UNIT_K);
    const bool two = units % 2 == 0 && wu_env != 1;
    if (a.T <= 16) return two ? launch_n<FMT, 16, 2, 3, 4>(a) : launch_n<FMT, 16, 1, 3, 4>(a);
    if (a.T <= 32) return two ? launch_n<FMT, 32, 2, 3, 2>(a) : launch_n<FMT, 32, 1, 3, 2>(a);
    if (a.T <= 64) return launch_n<FMT, 64, 1, 3, 1>(a);
    return launch_n<FMT, 128, 1, 2, 1>(a);
}

}  // namespace skn

bool skinny_supports(int fmt, const MmArgs& a) {
    if (a.T < 1 || a.T > skn::MAX_N || a.O < 1) return false;
    const int unit_k = fmt == GGQ_Q8_0 ? 128 : 256;
    if (a.K < unit_k || a.K % unit_k != 0) return false;
    if ((reinterpret_cast<uintptr_t>(a.W) & 15) || (reinterpret_cast<uintptr_t>(a.X) & 15) || (a.ldx & 7)) return false;
    if (((a.K / fmt_qk(fmt)) * fmt_blk(fmt)) % 16 != 0) return false;  // rows must be whole 16-byte vectors
    if (a.O > (int64_t{1} << 30)) return false;
    return a.sync == nullptr;
}

// "skinny_kernel<FMT, N, WU, G, MAXT>" of the launch launch_skinny() would make (bench.py / ggq_describe)
int skinny_describe(int fmt, const MmArgs& a, char* out, int cap) {
    const int unit_k = fmt == GGQ_Q8_0 ? 128 : 256;
    const int units = static_cast<int>(a.K / unit_k);
    const bool two = units % 2 == 0;
    int N, WU, G, MAXT;
    if (a.T <= 16) { N = 16; WU = two ? 2 : 1; G = 3; MAXT = 4; }
    else if (a.T <= 32) { N = 32; WU = two ? 2 : 1; G = 3; MAXT = 2; }
    else if (a.T <= 64) { N = 64; WU = 1; G = 3; MAXT = 1; }
    else { N = 128; WU = 1; G = 2; MAXT = 1; }
    static const char* const names[3] = {"Q8_0", "Q4_K", "Q6_K"};
    return snprintf(out, cap, "ggq::skn::skinny_kernel<%s,N=%d,WU=%d,G=%d,MAXT=%d> (tcgen05.mma, weights in TMEM)", names[fmt],
                    N, WU, G, MAXT);
}

int launch_skinny(int fmt, const MmArgs& a) {
    switch (fmt) {
        case GGQ_Q8_0: return skn::launch_fmt<0>(a);
        case GGQ_Q4_K: return skn::launch_fmt<1>(a);
        case GGQ_Q6_K: return skn::launch_fmt<2>(a);
    }
    return GGQ_E_FORMAT;
}

}  // namespace ggq
