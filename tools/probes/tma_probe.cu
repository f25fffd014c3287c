// dev probe: which element-granular TMA box starts are legal?  nvcc -arch=sm_100a tma_probe.cu -o tma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../gguf-triton-kernel_b200/csrc/tma.cuh"
using namespace ggq;

__global__ void k(const __grid_constant__ CUtensorMap map, int c0, int rows, int box_bytes, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&bar, rows * box_bytes);
        tma_load_2d(sm, &map, c0, 0, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < rows * box_bytes; i += blockDim.x) out[i] = sm[i];
}

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    int idx = -1;
    const int rowB = 210 * 8, rows = 8;
    std::vector<uint8_t> h(rowB * rows);
    for (size_t i = 0; i < h.size(); ++i) h[i] = static_cast<uint8_t>(i * 7 + (i >> 8));
    uint8_t *d, *o;
    cudaMalloc(&d, h.size());
    cudaMalloc(&o, 65536);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    struct Case { CUtensorMapDataType dt; int eb; int box_elems; int c0; const char* name; };
    const Case cases[] = {
        {CU_TENSOR_MAP_DATA_TYPE_INT32, 4, 60, 0, "int32 box240 c0=0"},
        {CU_TENSOR_MAP_DATA_TYPE_INT32, 4, 60, 52, "int32 box240 c0=52 (208 B, 16-aligned)"},
        {CU_TENSOR_MAP_DATA_TYPE_INT32, 4, 60, 53, "int32 box240 c0=53 (212 B, 4-aligned)"},
        {CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 120, 0, "u16 box240 c0=0"},
        {CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 120, 104, "u16 box240 c0=104 (208 B)"},
        {CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 120, 105, "u16 box240 c0=105 (210 B, 2-aligned)"},
        {CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 128, 105, "u16 box256 c0=105"},
        {CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 240, 210, "u8 box240 c0=210"},
        {CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 240, 211, "u8 box240 c0=211 (odd)"},
    };
    for (const Case& c : cases) {
        if (++idx != only && only >= 0) continue;
        alignas(64) CUtensorMap m;
        if (!make_map_2d(&m, c.dt, d, rowB / c.eb, rows, rowB, c.box_elems, rows, CU_TENSOR_MAP_SWIZZLE_NONE)) {
            printf("%-45s encode FAILED\n", c.name);
            continue;
        }
        const int bb = c.box_elems * c.eb;
        cudaMemset(o, 0xEE, 65536);
        k<<<1, 128, rows * bb>>>(m, c.c0, rows, bb, o);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("%-45s RUN ERROR %s\n", c.name, cudaGetErrorString(e));
            return 1;  // context is dead
        }
        std::vector<uint8_t> r(rows * bb);
        cudaMemcpy(r.data(), o, r.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int row = 0; row < rows; ++row)
            for (int j = 0; j < bb; ++j) {
                const int g = c.c0 * c.eb + j;
                const uint8_t want = g < rowB ? h[row * rowB + g] : 0;
                if (r[row * bb + j] != want) ++bad;
            }
        printf("%-45s ok, mismatches=%d\n", c.name, bad);
    }
    return 0;
}
