/*
 * ggq.h — C ABI of the B200-native GGUF mmq path (libggq.so).
 *
 * This is the drop-in boundary for ONE hot path of PowerfulGhost/gguf-triton-kernel: multiplying
 * GGUF block-quantized weights (Q8_0 / Q4_K / Q6_K) by fp16 activations.  Every entry point below
 * names the reference interface it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - All data pointers are DEVICE pointers on the device that is current on the calling thread.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls enqueue work
 *     on that stream and return; they never synchronize, never allocate device memory and never
 *     throw.  (The reference launches its Triton kernels the same way: asynchronously on the
 *     current stream, kernels/mmq_q8_0.py:133-147.)
 *   - Return value: 0 on success; a positive value is a cudaError_t from the launch; a negative
 *     value is one of GGQ_E_* (argument errors, detected before anything is enqueued).
 *   - Naming: O = out-features (reference arg `M`, rows of the packed weight), T = tokens (reference
 *     arg `N`, rows of the activation), K = in-features.  C[T, O] = X[T, K] . dequant(W)[O, K]^T.
 *   - Packed weight W: row-major rows of K/QK blocks, exactly the byte stream the reference packers
 *     emit (utils/quantize/q8_0.py:41-47; q4_k_ref.c:76-89 `block_q4_K`; q6_k_ref.c:62-68
 *     `block_q6_K`).  No repacking, no alignment requirement beyond what cudaMalloc/torch give the
 *     base pointer; shapes whose rows are not 16-byte multiples take a slower general kernel.
 *   - Arithmetic: fp16 activations (NOT re-quantized to Q8_1), fp32 accumulation, fp16 output.
 */
#ifndef GGQ_H_
#define GGQ_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GGQ_VERSION 108 /* 0.1.8 */

/* argument errors (negative so they cannot collide with cudaError_t) */
#define GGQ_E_SHAPE     (-1) /* K not a multiple of the block size (the reference's only assert:   */
                             /* kernels/mmq_q8_0.py:124, mmq_q4_k.py:263, mmq_q6_k.py:211), or     */
                             /* O/T/K/ldc negative or out of range                                  */
#define GGQ_E_POINTER   (-2) /* NULL data pointer with a non-empty problem                          */
#define GGQ_E_FAMILY    (-3) /* requested kernel family cannot run this shape                       */
#define GGQ_E_FORMAT    (-4) /* unknown quant format id                                             */

/* quant formats */
#define GGQ_Q8_0 0 /* 32 weights / 34 B:  fp16 d, int8 qs[32]                                       */
#define GGQ_Q4_K 1 /* 256 weights / 144 B: fp16 d, fp16 dmin, 12 B 6-bit scales+mins, qs[128]       */
#define GGQ_Q6_K 2 /* 256 weights / 210 B: ql[128], qh[64], int8 scales[16], fp16 d                 */

/* kernel families (ggq_mm_ex `family`) */
#define GGQ_FAMILY_AUTO    0 /* decode / skinny for T<=16, skinny for 17..127, prefill from 128 on   */
#define GGQ_FAMILY_GENERIC 1 /* any shape, any alignment; one warp per output row                    */
#define GGQ_FAMILY_DECODE  2 /* HBM-bound skinny GEMM, T<=16: TMA bulk-staged packed rows,           */
                             /* register unpack, mma.sync m16n8k16 f16->f32                           */
#define GGQ_FAMILY_PREFILL 3 /* tensor-bound GEMM: packed tiles -> smem dequant -> tcgen05.mma/TMEM   */
#define GGQ_FAMILY_SKINNY  4 /* HBM-bound skinny GEMM, 2<=T<=128: packed tiles -> register dequant ->  */
                             /* tcgen05.st (weights = TMEM A operand) -> tcgen05.mma, N = tokens      */

/*
 * mmq entry points.  Replace the bodies of
 *     kernels/mmq_q8_0.py:102  mmq_q8_0(A, B, M, N, K)   (+ Triton kernel :13-93)
 *     kernels/mmq_q4_k.py:240  mmq_q4_k(A, B, M, N, K)   (+ Triton kernels :30-229)
 *     kernels/mmq_q6_k.py:197  mmq_q6_k(A, B, M, N, K)   (+ Triton kernels :28-186)
 * with  W = A.data_ptr(), X = B.data_ptr(), O = M, T = N.  C is the fp16 [T, O] contiguous output the
 * reference allocates at mmq_q8_0.py:128 (element [t, o] at O*t + o, :91).
 */
int ggq_mm_q8_0_f16(const void* W, const void* X, void* C, int64_t O, int64_t T, int64_t K, void* stream);
int ggq_mm_q4_k_f16(const void* W, const void* X, void* C, int64_t O, int64_t T, int64_t K, void* stream);
int ggq_mm_q6_k_f16(const void* W, const void* X, void* C, int64_t O, int64_t T, int64_t K, void* stream);

/*
 * Extended form used by the tests (family pinning), the N-split multi-GPU driver (ldc / peer
 * outputs) and the benchmarks.
 *   fmt     GGQ_Q8_0 | GGQ_Q4_K | GGQ_Q6_K
 *   ldx     row stride of X in elements (>= K)
 *   ldc     row stride of every output in elements (>= O); element [t, o] is written at ldc*t + o
 *   n_out   number of output buffers (1..8); the same tile is stored to every C_out[i].  With
 *           n_out > 1 the extra pointers are peer-mapped buffers of the other ranks (NVLink), which
 *           fuses the N-split all-gather into the epilogue.
 *   family  GGQ_FAMILY_*
 */
int ggq_mm_ex(int fmt, const void* W, const void* X, int64_t ldx, void* const* C_out, int n_out, int64_t ldc,
              int64_t O, int64_t T, int64_t K, int family, void* stream);

/*
 * Fused SwiGLU up-projection, the step that follows the gate/up matmuls of a Llama FFN (the reference has
 * no counterpart: its callers run mmq twice and apply silu(gate) * up in torch):
 *     C[T, O] = silu(G) * U,   G = fp16(X . dequant(Wg)^T),  U = fp16(X . dequant(Wu)^T),  silu and product in fp32
 * i.e. exactly what `F.silu(mmq(Wg, X).float()) * mmq(Wu, X).float()` rounded to fp16 gives.  Wg / Wu: packed [O, K] of the
 * same format.  T <= 16 on decode-eligible shapes: ONE kernel streams both matrices once and applies the activation in
 * the accumulator registers (no intermediate [T, 2*O] in HBM).  Other shapes: gate GEMM -> workspace, up GEMM -> C,
 * one elementwise pass; that form needs `workspace_bytes` >= ggq_mm_swiglu_workspace(...) (0 when the fused kernel
 * runs; `workspace` may then be NULL).  The query assumes 16-byte aligned Wg / Wu / X (what cudaMalloc and torch give);
 * with a less aligned pointer the call takes the composed form and returns GGQ_E_POINTER if no workspace was passed.
 */
int64_t ggq_mm_swiglu_workspace(int fmt, int64_t O, int64_t T, int64_t K);
int ggq_mm_swiglu(int fmt, const void* Wg, const void* Wu, const void* X, void* C, int64_t O, int64_t T, int64_t K,
                  void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Decode-family GEMV/skinny GEMM with the N-split exchange fused INTO the kernel (T <= 8, T*K <= 65536):
 * ONE kernel per rank and step, no NCCL call, no barrier kernel, no system-scope fence.  Both transfers use a
 * flag-in-data ("LL") line format: a 16-byte line {data0, epoch, data1, epoch} carries 4 fp16 values and is written
 * with one 16-byte store over NVLink; the receiver polls the line itself, so data and "it is there" arrive in the
 * same NVLink hop.
 *   - activations: the owner rank (x_owner) reads X from its own memory and pushes it as lines into every peer's
 *     x_land buffer; the peers' CTAs poll their local x_land and unpack straight into shared memory;
 *   - outputs: every finished 16-row tile is stored to this rank's C and sent as lines into every peer's c_land; at
 *     the end every CTA polls its share of the lines addressed to this rank and writes them into C as plain fp16:
 *     when the kernel completes, the full C[T, world * O] is present in this rank's buffer.
 * W holds this rank's O rows.  `C` points at column 0 of this rank's [T, ldc] result (ldc >= world * O); the rank's
 * own columns start at rank * O.  O % 4 == 0 and K % 4 == 0.
 * Landing buffers live in peer-mapped (symmetric) memory, zero-initialised once: x_land = 2 halves (epoch parity) of
 * `x_land_half` bytes each, >= T*K*4; c_land = 2 halves of `c_land_half` bytes each, >= world*T*O*4.
 * `epoch` (>= 1) must increase by 1 per call on all ranks; `counter` is a zero-initialised uint32 in local device
 * memory.  Every rank of the group must make the matching call: a rank that waits in vain gives up after
 * `timeout_ns` (see `status`).
 */
typedef struct ggq_peer_sync {
    int32_t rank, world;
    int32_t x_owner;       /* rank whose X is the step's activations */
    uint32_t epoch;        /* this call's epoch (ignored in replayable mode) */
    uint32_t* counter;     /* local: CTAs that finished this call (the kernel resets it) */
    /* Replayable mode (epoch_dev != NULL): nothing in the call changes from step to step, so the launch can be
     * captured in a CUDA graph and replayed.  The epoch of a call is *epoch_dev + 1, read by the kernel; the last CTA
     * writes it back.  Calls with an EVEN epoch use (X, C), calls with an ODD epoch use (X_alt, C_alt): the output
     * double buffering that keeps a fast rank from overwriting results a slow rank's consumer is still reading. */
    uint32_t* epoch_dev;   /* device word, zero-initialised, local to this rank */
    const void* X_alt;     /* owner rank only */
    void* C_alt;
    void* x_land;          /* this rank's landing buffer for the activations (unused on the owner) */
    void* x_land_peer[8];  /* owner rank: the peers' x_land (own entry ignored) */
    int64_t x_land_half;   /* bytes of one parity half of x_land */
    void* c_land;          /* this rank's landing buffer for the peers' output slices */
    void* c_land_peer[8];  /* the peers' c_land (own entry ignored) */
    int64_t c_land_half;   /* bytes of one parity half of c_land */
    /* Bounded waits: every poll of a line that has not arrived gives up after `timeout_ns` nanoseconds of the
     * device's global timer (0 = the default of 2 s).  A wait that gives up stores a GGQ_SYNC_* code (first one wins)
     * into *status — a zero-initialised uint32 in local device memory, may be NULL — and the kernel runs to completion
     * with whatever arrived, so a missing or late rank costs a bounded time and is reported instead of hanging the
     * GPU.  The results of a step whose *status is non-zero are undefined. */
    uint32_t* status;
    uint64_t timeout_ns;
} ggq_peer_sync;
#define GGQ_SYNC_OK 0u
#define GGQ_SYNC_TIMEOUT_X 1u     /* the activations did not arrive */
#define GGQ_SYNC_TIMEOUT_PEER 2u  /* a peer's output lines did not arrive */

/* Returns 0 and writes the number of CTAs launched to *ctas_out. */
int ggq_mm_sync(int fmt, const void* W, const void* X, int64_t ldx, void* C, int64_t ldc, int64_t O, int64_t T, int64_t K,
                const ggq_peer_sync* sync, int* ctas_out, void* stream);

/*
 * The whole step with HOST activations and a HOST result: what a caller of kernels/mmq_q4_k.py:240 does around the call
 * (`.cuda()` of the activations, `.cpu()` of the product) behind one C call.  A pipe owns `depth` device slots (X and C
 * staging buffers, allocated once by ggq_host_pipe_create on the current device — the ONLY functions of this library
 * that allocate) and three streams: H2D of step i+1 and D2H of step i-1 overlap the kernel of step i.
 *   W_dev    packed weights on the pipe's device (resident state of the layer, as in the reference)
 *   X_host   fp16 [T, K] in host memory (page-locked memory makes the copy asynchronous); must stay valid and unchanged
 *            until the step has been copied in — at the latest after ggq_host_pipe_sync
 *   C_host   fp16 [T, O] in host memory, complete after ggq_host_pipe_sync (or once `depth` later calls have returned
 *            and been synchronised by the caller's own means: see ggq_host_pipe_stream)
 * T*K*2 <= max_x_bytes and T*O*2 <= max_c_bytes, else GGQ_E_SHAPE.  Calls on one pipe must come from one thread at a time.
 */
#define GGQ_HOST_PIPE_MAX_DEPTH 8
typedef struct ggq_host_pipe ggq_host_pipe;
int ggq_host_pipe_create(ggq_host_pipe** out, int64_t max_x_bytes, int64_t max_c_bytes, int depth);
int ggq_mm_host(ggq_host_pipe* pipe, int fmt, const void* W_dev, const void* X_host, void* C_host, int64_t O, int64_t T,
                int64_t K);
int ggq_host_pipe_sync(ggq_host_pipe* pipe);
/* cudaStream_t of the pipe: which = 0 copy-in, 1 kernels, 2 copy-out (e.g. to record the caller's own events) */
void* ggq_host_pipe_stream(ggq_host_pipe* pipe, int which);
void ggq_host_pipe_destroy(ggq_host_pipe* pipe);

/*
 * N-split exchange of a prefill-sized result by the copy engines: the same [rows x width_bytes] column block (row pitch
 * pitch_bytes, identical on both sides) is copied from `src` to each of the n_dst (<= 8) peer-mapped destinations with
 * one 2-D asynchronous copy per peer on `stream`.  multigpu/nsplit.py: every rank computes its [T, O/N] slice into its
 * own [T, O] buffer and pushes the slice into the same columns of every peer's buffer — large DMA transfers over NVLink
 * instead of 64-byte peer stores from the GEMM epilogue (8 x B200, T=4096, O=28672: 2.0 ms -> see DESIGN.md §5).
 */
int ggq_push_columns(const void* src, void* const* dst, int n_dst, int64_t pitch_bytes, int64_t width_bytes, int64_t rows,
                     void* stream);

/*
 * Dequantize packed rows to fp16 [O, K] with the SAME device functions the prefill GEMM uses.
 * Bit-exact targets: utils/quantize/q8_0.py:52-100 dequantize_q8_0, q4_k.py:146-158 dequantize_q4_k,
 * q6_k.py:138-159 dequantize_q6_k (fp32 there; `.half()` of it here).
 */
int ggq_dequant_q8_0_f16(const void* W, void* out, int64_t O, int64_t K, void* stream);
int ggq_dequant_q4_k_f16(const void* W, void* out, int64_t O, int64_t K, void* stream);
int ggq_dequant_q6_k_f16(const void* W, void* out, int64_t O, int64_t K, void* stream);

/*
 * "Next" rows of the path, same conventions:
 *   ggq_quantize_q8_0_f16   fp16 x[n] -> Q8_0 blocks (n/32 * 34 B), byte-identical to utils/quantize/q8_0.py:4-49
 *   ggq_quantize_q8_1_f16   fp16 x[n] -> Q8_1 blocks (n/32 * 36 B), byte-identical to utils/quantize/q8_1.py:18-70
 *   ggq_dequant_q6_k_f32    fp32 [O, K], bit-identical to utils/quantize/q6_k.py:138-159 (which returns fp32)
 * n must be a multiple of 32 (the reference raises ValueError, q8_0.py:14-15).
 */
int ggq_quantize_q8_0_f16(const void* x, void* out, int64_t n, void* stream);
/*
 * K-quant packers: fp32 x[n] (16-byte aligned, n a multiple of 256) -> Q4_K blocks (n/256 * 144 B) / Q6_K blocks
 * (n/256 * 210 B), byte-identical to the reference's compiled packers utils/quantize/q4_k_ref.c:281-368
 * quantize_row_q4_K_ref and q6_k_ref.c:251-340 quantize_row_q6_K_ref (same fp32 operation sequence, no FMA).
 */
int ggq_quantize_q4_k_f32(const void* x, void* out, int64_t n, void* stream);
int ggq_quantize_q6_k_f32(const void* x, void* out, int64_t n, void* stream);
int ggq_quantize_q8_1_f16(const void* x, void* out, int64_t n, void* stream);
int ggq_dequant_q6_k_f32(const void* W, void* out, int64_t O, int64_t K, void* stream);

/*
 * Reference-arithmetic mode: XQ = activations packed as Q8_1 (T rows of K/32 36-byte blocks, e.g. by
 * ggq_quantize_q8_1_f16); integer block dots and fp16 accumulation in the exact operation order of
 * kernels/cpu_impls/mmq_{q8_0,q4_k,q6_k}_q8_1_cpu.py, so C[T, O] equals their result bit for bit.  One thread per weight
 * row (the fp16 accumulation chain of an output is sequential in the reference), vector loads and DP4A block dots when
 * the rows are whole 32-bit words (Q4_K always — through a cp.async-filled shared-memory tile; Q8_0 / Q6_K: an even
 * number of blocks per row), byte-wise otherwise.  Q4_K with T >= 3 and XQ 16-byte aligned: the block dots of 16 rows x
 * 8 tokens by one integer tensor-core MMA, the chains in the accumulator-fragment layout (same bits).  Kernel selection
 * depends only on shape and pointer alignment (W 16-byte aligned and XQ 4-byte aligned for the vector kernels).
 * For callers that want the reference's own numbers: its int8 activation noise (~5e-3) is outside the tolerance the
 * fp16-activation entry points above are held to.
 */
int ggq_mm_ref_q8_1(int fmt, const void* W, const void* XQ, void* C, int64_t O, int64_t T, int64_t K, void* stream);

/* Bytes of a packed [O, K] weight (O * K/QK * block bytes), or GGQ_E_* (<0). */
int64_t ggq_packed_nbytes(int fmt, int64_t O, int64_t K);

/* Which family GGQ_FAMILY_AUTO resolves to for this problem (GGQ_FAMILY_*), or GGQ_E_* (<0). */
int ggq_select_family(int fmt, int64_t O, int64_t T, int64_t K);

/* Host-only: the decode family's static work decomposition for a problem (no GPU needed).
 * out9 = {warps per tile, live tiles per warp, 8-token n-tiles, K-slices, chunks per slice, pipeline
 * stages per warp, CTAs, batches, dynamic shared memory bytes}.  0, or GGQ_E_* (<0). */
int ggq_decode_plan(int fmt, int64_t O, int64_t T, int64_t K, int* out9);

/* Host-only: one-line description (kernel template + plan) of the launch GGQ_FAMILY_AUTO makes for this problem,
 * written to out[cap] (NUL-terminated).  0, or GGQ_E_* (<0).  bench.py reports it as `roofline.kernel`. */
int ggq_describe(int fmt, int64_t O, int64_t T, int64_t K, char* out, int cap);

/* Development aid: `buf` = device buffer of 16 x 320 x 8 uint64 (or NULL to stop).  While set, every decode-family launch
 * writes per-CTA %globaltimer stamps of its phases (entry, first boxes issued, dependency wait passed, activations
 * staged, warp 0 done, CTA done) into slot (launch index % 16).  tools/trace_decode.py prints the timeline. */
void ggq_dev_set_trace(void* buf);

/* Kernels launched by this library since load (all families; for bench.py's `gpu_launches`). */
int64_t ggq_launch_count(void);

/* Static description of an error code returned by any function above. */
const char* ggq_error_string(int code);

int ggq_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GGQ_H_ */
