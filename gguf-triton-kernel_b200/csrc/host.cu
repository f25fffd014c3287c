// host.cu — the mmq call with HOST activations and a HOST result (include/ggq.h, `ggq_host_pipe`).
//
// The reference's callers hold the packed weights on the device and hand every step a fresh activation tensor
// (kernels/mmq_q4_k.py:240-289 takes device tensors; the caller's `.cuda()` / `.cpu()` are the copies).  This is
// that whole step behind one C call: H2D of X, the mm kernel, D2H of C — on three streams with `depth` rotating
// device slots, so the copy-in of step i+1 and the copy-out of step i-1 overlap the kernel of step i and the kernels
// stay back to back on their stream (programmatic dependent launch keeps working: only event waits sit between
// them).  All of it is enqueued asynchronously; `ggq_host_pipe_sync` waits for everything submitted so far.
#include <new>

#include "../../include/ggq.h"
#include "common.cuh"
#include "formats.cuh"

struct ggq_host_pipe {
    int depth = 0;
    int device = 0;
    int64_t x_cap = 0, c_cap = 0;    // bytes per slot
    uint64_t calls = 0;
    cudaStream_t s_in = nullptr, s_mm = nullptr, s_out = nullptr;
    uint8_t* x_dev[GGQ_HOST_PIPE_MAX_DEPTH] = {};
    uint8_t* c_dev[GGQ_HOST_PIPE_MAX_DEPTH] = {};
    cudaEvent_t x_ready[GGQ_HOST_PIPE_MAX_DEPTH] = {};   // H2D of the slot's X is complete
    cudaEvent_t mm_done[GGQ_HOST_PIPE_MAX_DEPTH] = {};   // the kernel has read X and written C of the slot
    cudaEvent_t c_out[GGQ_HOST_PIPE_MAX_DEPTH] = {};     // D2H of the slot's C is complete
};

extern "C" {

void ggq_host_pipe_destroy(ggq_host_pipe* p) {
    if (!p) return;
    for (int i = 0; i < p->depth; ++i) {
        if (p->x_dev[i]) cudaFree(p->x_dev[i]);
        if (p->c_dev[i]) cudaFree(p->c_dev[i]);
        if (p->x_ready[i]) cudaEventDestroy(p->x_ready[i]);
        if (p->mm_done[i]) cudaEventDestroy(p->mm_done[i]);
        if (p->c_out[i]) cudaEventDestroy(p->c_out[i]);
    }
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_mm) cudaStreamDestroy(p->s_mm);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    delete p;
}

int ggq_host_pipe_create(ggq_host_pipe** out, int64_t max_x_bytes, int64_t max_c_bytes, int depth) {
    if (!out) return GGQ_E_POINTER;
    *out = nullptr;
    if (depth < 1 || depth > GGQ_HOST_PIPE_MAX_DEPTH || max_x_bytes < 1 || max_c_bytes < 1) return GGQ_E_SHAPE;
    ggq_host_pipe* p = new (std::nothrow) ggq_host_pipe();
    if (!p) return static_cast<int>(cudaErrorMemoryAllocation);
    p->depth = depth;
    p->x_cap = (max_x_bytes + 255) & ~int64_t{255};
    p->c_cap = (max_c_bytes + 255) & ~int64_t{255};
    cudaError_t e = cudaGetDevice(&p->device);
    auto ok = [&](cudaError_t r) {
        if (e == cudaSuccess) e = r;
        return e == cudaSuccess;
    };
    ok(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&p->s_mm, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < depth && e == cudaSuccess; ++i) {
        ok(cudaMalloc(reinterpret_cast<void**>(&p->x_dev[i]), static_cast<size_t>(p->x_cap)));
        ok(cudaMalloc(reinterpret_cast<void**>(&p->c_dev[i]), static_cast<size_t>(p->c_cap)));
        ok(cudaEventCreateWithFlags(&p->x_ready[i], cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&p->mm_done[i], cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&p->c_out[i], cudaEventDisableTiming));
    }
    if (e != cudaSuccess) {
        ggq_host_pipe_destroy(p);
        return static_cast<int>(e);
    }
    *out = p;
    return 0;
}

int ggq_mm_host(ggq_host_pipe* p, int fmt, const void* W_dev, const void* X_host, void* C_host, int64_t O, int64_t T,
                int64_t K) {
    if (!p) return GGQ_E_POINTER;
    if (fmt < GGQ_Q8_0 || fmt > GGQ_Q6_K) return GGQ_E_FORMAT;
    if (O < 0 || T < 0 || K < 0 || K % ggq::fmt_qk(fmt) != 0) return GGQ_E_SHAPE;
    if (O == 0 || T == 0) return 0;
    if (!C_host || (K > 0 && (!W_dev || !X_host))) return GGQ_E_POINTER;
    const int64_t xb = T * K * 2, cb = T * O * 2;
    if (xb > p->x_cap || cb > p->c_cap) return GGQ_E_SHAPE;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev != p->device) return GGQ_E_POINTER;   // the pipe's buffers live on p->device
    const int s = static_cast<int>(p->calls % static_cast<uint64_t>(p->depth));
    const bool reused = p->calls >= static_cast<uint64_t>(p->depth);
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) {
        if (e == cudaSuccess) e = r;
    };
    // copy-in: the slot's X buffer is free once the kernel `depth` calls ago has run
    if (reused) ok(cudaStreamWaitEvent(p->s_in, p->mm_done[s], 0));
    if (xb > 0) ok(cudaMemcpyAsync(p->x_dev[s], X_host, static_cast<size_t>(xb), cudaMemcpyHostToDevice, p->s_in));
    ok(cudaEventRecord(p->x_ready[s], p->s_in));
    // kernel: needs X, and the slot's C buffer copied out by the D2H `depth` calls ago
    ok(cudaStreamWaitEvent(p->s_mm, p->x_ready[s], 0));
    if (reused) ok(cudaStreamWaitEvent(p->s_mm, p->c_out[s], 0));
    if (e != cudaSuccess) return static_cast<int>(e);
    void* outs[1] = {p->c_dev[s]};
    const int rc = ggq_mm_ex(fmt, W_dev, p->x_dev[s], K, outs, 1, O, O, T, K, GGQ_FAMILY_AUTO, p->s_mm);
    if (rc != 0) return rc;
    ok(cudaEventRecord(p->mm_done[s], p->s_mm));
    // copy-out
    ok(cudaStreamWaitEvent(p->s_out, p->mm_done[s], 0));
    ok(cudaMemcpyAsync(C_host, p->c_dev[s], static_cast<size_t>(cb), cudaMemcpyDeviceToHost, p->s_out));
    ok(cudaEventRecord(p->c_out[s], p->s_out));
    ++p->calls;
    return static_cast<int>(e);
}

int ggq_host_pipe_sync(ggq_host_pipe* p) {
    if (!p) return GGQ_E_POINTER;
    // every call ends with its D2H on s_out, and s_out's work is ordered after the call's H2D and kernel
    cudaError_t e = cudaStreamSynchronize(p->s_out);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->s_mm);
    return static_cast<int>(e);
}

int ggq_push_columns(const void* src, void* const* dst, int n_dst, int64_t pitch_bytes, int64_t width_bytes, int64_t rows,
                     void* stream) {
    if (n_dst < 0 || n_dst > 8 || pitch_bytes < width_bytes || width_bytes < 0 || rows < 0) return GGQ_E_SHAPE;
    if (n_dst == 0 || width_bytes == 0 || rows == 0) return 0;
    if (!src || !dst) return GGQ_E_POINTER;
    for (int i = 0; i < n_dst; ++i)
        if (!dst[i]) return GGQ_E_POINTER;
    for (int i = 0; i < n_dst; ++i) {
        const cudaError_t e = cudaMemcpy2DAsync(dst[i], static_cast<size_t>(pitch_bytes), src, static_cast<size_t>(pitch_bytes),
                                                static_cast<size_t>(width_bytes), static_cast<size_t>(rows), cudaMemcpyDeviceToDevice,
                                                static_cast<cudaStream_t>(stream));
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    return 0;
}

void* ggq_host_pipe_stream(ggq_host_pipe* p, int which) {
    if (!p) return nullptr;
    return which == 0 ? p->s_in : which == 1 ? p->s_mm : which == 2 ? p->s_out : nullptr;
}

}  // extern "C"
