// tma.cuh — 2-D tensor maps (host) and the tensor-tile TMA load (device), shared by both fast families.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "ptx.cuh"

namespace ggq {

// cp.async.bulk.tensor.2d global -> shared (UTMALDG), completion on an mbarrier; OOB elements are zero-filled
// and still counted in the transaction bytes.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

inline PFN_cuTensorMapEncodeTiled tensor_map_encoder() {  // resolved through the runtime: no -lcuda
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(f);
    });
    return fn;
}

// row-major [outer, inner] tensor with `row_stride_bytes` between rows; box = [box_outer, box_inner]
inline bool make_map_2d(CUtensorMap* m, CUtensorMapDataType dt, const void* base, uint64_t inner, uint64_t outer,
                        uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle sw) {
    auto enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t dims[2] = {inner, outer};
    const cuuint64_t strides[1] = {row_stride_bytes};
    const cuuint32_t box[2] = {box_inner, box_outer};
    const cuuint32_t estr[2] = {1, 1};
    return enc(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// fp16 [T, K] activations (row stride ldx elements) as the 3-D tensor (64 k, T tokens, K/64 k-groups): one box
// (64, box_tokens, box_groups) lands in shared memory as box_groups consecutive [box_tokens x 128 B] SWIZZLE_128B atoms,
// the K-major UMMA operand layout.  Tokens >= T are zero-filled.
inline bool make_map_x3d(CUtensorMap* m, const void* base, uint64_t K, uint64_t T, uint64_t ldx, uint32_t box_tokens,
                         uint32_t box_groups) {
    auto enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t dims[3] = {64, T, K / 64};
    const cuuint64_t strides[2] = {ldx * 2, 128};
    const cuuint32_t box[3] = {64, box_tokens, box_groups};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Encoding a map costs ~1 us of host time, as much as the rest of a call: the last few maps of this thread are kept,
// keyed by everything that goes into the encoding.  The returned map is copied into the kernel parameters at launch.
inline const CUtensorMap* cached_map_2d(CUtensorMapDataType dt, const void* base, uint64_t inner, uint64_t outer,
                                        uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer,
                                        CUtensorMapSwizzle sw, int dev) {
    struct Slot {
        const void* base;
        uint64_t inner, outer, stride;
        uint32_t bi, bo;
        int dt, sw, dev;
        bool valid;
        alignas(64) CUtensorMap map;
    };
    constexpr int SLOTS = 16;
    static thread_local Slot cache[SLOTS] = {};
    static thread_local unsigned next = 0;
    for (Slot& c : cache)
        if (c.valid && c.base == base && c.inner == inner && c.outer == outer && c.stride == row_stride_bytes &&
            c.bi == box_inner && c.bo == box_outer && c.dt == static_cast<int>(dt) && c.sw == static_cast<int>(sw) &&
            c.dev == dev)
            return &c.map;
    Slot& c = cache[next++ % SLOTS];
    c.valid = false;
    if (!make_map_2d(&c.map, dt, base, inner, outer, row_stride_bytes, box_inner, box_outer, sw)) return nullptr;
    c.base = base; c.inner = inner; c.outer = outer; c.stride = row_stride_bytes;
    c.bi = box_inner; c.bo = box_outer; c.dt = static_cast<int>(dt); c.sw = static_cast<int>(sw); c.dev = dev;
    c.valid = true;
    return &c.map;
}

// the same cache policy for the 3-D activation maps of make_map_x3d
inline const CUtensorMap* cached_map_x3d(const void* base, uint64_t K, uint64_t T, uint64_t ldx, uint32_t box_tokens,
                                         uint32_t box_groups, int dev) {
    struct Slot {
        const void* base;
        uint64_t K, T, ldx;
        uint32_t bt, bg;
        int dev;
        bool valid;
        alignas(64) CUtensorMap map;
    };
    constexpr int SLOTS = 8;
    static thread_local Slot cache[SLOTS] = {};
    static thread_local unsigned next = 0;
    for (Slot& c : cache)
        if (c.valid && c.base == base && c.K == K && c.T == T && c.ldx == ldx && c.bt == box_tokens && c.bg == box_groups &&
            c.dev == dev)
            return &c.map;
    Slot& c = cache[next++ % SLOTS];
    c.valid = false;
    if (!make_map_x3d(&c.map, base, K, T, ldx, box_tokens, box_groups)) return nullptr;
    c.base = base; c.K = K; c.T = T; c.ldx = ldx; c.bt = box_tokens; c.bg = box_groups; c.dev = dev;
    c.valid = true;
    return &c.map;
}

}  // namespace ggq
