// torch_binding.cpp — the PyTorch extension over the C ABI (include/ggq.h): `_ggq_torch.*.so`.
//
// The reference's entry points take and return torch tensors (kernels/mmq_q8_0.py:102, mmq_q4_k.py:240,
// mmq_q6_k.py:197); this shim does their operand checks, allocates the fp16 [N, M] result on A's device and
// calls libggq.so on the current stream.  No arithmetic lives here and there is no fallback of any kind: a non-zero
// return code of the library raises.  (kernels/_ext.py binds the same C ABI through ctypes for everything that is
// not per-step; this extension exists because a ctypes call costs ~9 us of host time per step and this one ~2.)
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "../../include/ggq.h"

namespace {

const int QK[3] = {32, 256, 256};
const int BLK[3] = {34, 144, 210};

void check_rc(int rc, const char* what) {
    TORCH_CHECK(rc == 0, what, " failed (", rc, "): ", ggq_error_string(rc));
}

// C[N, M] (fp16) = B[N, K] @ dequant(A)[M, K]^T
at::Tensor mm(int64_t fmt, const at::Tensor& A, const at::Tensor& B, int64_t M, int64_t N, int64_t K, int64_t family,
              const c10::optional<at::Tensor>& out) {
    TORCH_CHECK(fmt >= 0 && fmt <= 2, "unknown quant format ", fmt);
    TORCH_CHECK_TYPE(A.scalar_type() == at::kChar || A.scalar_type() == at::kByte, "A must be int8 packed blocks, got ",
                     A.scalar_type());
    TORCH_CHECK_TYPE(B.scalar_type() == at::kHalf, "B must be float16, got ", B.scalar_type());
    TORCH_CHECK_VALUE(A.is_cuda() && B.is_cuda() && A.device() == B.device(),
                      "A and B must live on the same CUDA device (no CPU path)");
    TORCH_CHECK_VALUE(A.is_contiguous() && B.is_contiguous(), "A and B must be contiguous");
    TORCH_CHECK_VALUE(M >= 0 && N >= 0 && K >= 0, "negative size");
    const int64_t want = M * (K / QK[fmt]) * BLK[fmt];
    TORCH_CHECK_VALUE(A.numel() == want, "A has ", A.numel(), " bytes, expected ", want, " for M=", M, ", K=", K);
    TORCH_CHECK_VALUE(B.numel() == N * K, "B has ", B.numel(), " elements, expected N*K=", N * K);
    c10::cuda::CUDAGuard guard(A.device());
    at::Tensor C;
    if (out.has_value()) {
        C = *out;
        TORCH_CHECK_VALUE(C.scalar_type() == at::kHalf && C.dim() == 2 && C.size(0) == N && C.size(1) == M &&
                              C.is_contiguous() && C.device() == A.device(),
                          "out must be a contiguous float16 [N, M] tensor on A's device");
    } else {
        C = at::empty({N, M}, B.options());
    }
    void* outs[1] = {C.data_ptr()};
    cudaStream_t stream = at::cuda::getCurrentCUDAStream(A.device().index()).stream();
    check_rc(ggq_mm_ex(static_cast<int>(fmt), A.data_ptr(), B.data_ptr(), K, outs, 1, M, M, N, K, static_cast<int>(family),
                       stream),
             "ggq_mm");
    return C;
}

// fused SwiGLU up-projection: C[N, M] = silu(B @ dequant(Ag)^T) * (B @ dequant(Au)^T)
at::Tensor mm_swiglu(int64_t fmt, const at::Tensor& Ag, const at::Tensor& Au, const at::Tensor& B, int64_t M, int64_t N,
                     int64_t K) {
    TORCH_CHECK(fmt >= 0 && fmt <= 2, "unknown quant format ", fmt);
    for (const at::Tensor* A : {&Ag, &Au}) {
        TORCH_CHECK_TYPE(A->scalar_type() == at::kChar || A->scalar_type() == at::kByte, "packed weights must be int8");
        TORCH_CHECK_VALUE(A->is_cuda() && A->device() == B.device() && A->is_contiguous(),
                          "packed weights must be contiguous and on B's CUDA device");
        TORCH_CHECK_VALUE(A->numel() == M * (K / QK[fmt]) * BLK[fmt], "packed size does not match M, K");
    }
    TORCH_CHECK_TYPE(B.scalar_type() == at::kHalf, "B must be float16");
    TORCH_CHECK_VALUE(B.is_cuda() && B.is_contiguous() && B.numel() == N * K, "B must be a contiguous CUDA [N, K] tensor");
    c10::cuda::CUDAGuard guard(B.device());
    at::Tensor C = at::empty({N, M}, B.options());
    const int64_t ws_bytes = ggq_mm_swiglu_workspace(static_cast<int>(fmt), M, N, K);
    TORCH_CHECK(ws_bytes >= 0, "ggq_mm_swiglu_workspace failed (", ws_bytes, "): ", ggq_error_string(static_cast<int>(ws_bytes)));
    at::Tensor ws;   // the composed form's gate projection (shapes the fused kernel does not take)
    if (ws_bytes > 0) ws = at::empty({ws_bytes}, B.options().dtype(at::kByte));
    cudaStream_t stream = at::cuda::getCurrentCUDAStream(B.device().index()).stream();
    check_rc(ggq_mm_swiglu(static_cast<int>(fmt), Ag.data_ptr(), Au.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K,
                           ws_bytes > 0 ? ws.data_ptr() : nullptr, ws_bytes, stream),
             "ggq_mm_swiglu");
    return C;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "PyTorch extension over libggq.so (C ABI: include/ggq.h)";
    m.def("mm", &mm, py::arg("fmt"), py::arg("A"), py::arg("B"), py::arg("M"), py::arg("N"), py::arg("K"),
          py::arg("family") = 0, py::arg("out") = py::none());
    m.def("mm_swiglu", &mm_swiglu, py::arg("fmt"), py::arg("Ag"), py::arg("Au"), py::arg("B"), py::arg("M"), py::arg("N"),
          py::arg("K"));
    m.def("version", []() { return ggq_version(); });
}
