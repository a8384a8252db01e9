// Standalone probe of the tensor-map copy used by sdv_gemm_tma_kernel: one 128 x 16 box of doubles, 128-byte swizzle.
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tools/probes/tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, const int c0, const int r0, double* out) {
    extern __shared__ unsigned char raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + 16384);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(16384) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(s32(base)), "l"(&map), "r"(s32(bar)), "r"(c0), "r"(r0) : "memory");
    }
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(s32(bar)), "r"(0) : "memory");
    } while (!done);
    for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) {
        const int r = i / 16, c = i % 16;
        out[i] = *reinterpret_cast<const double*>(base + r * 128 + (((c >> 1) ^ (r & 7)) << 4) + (c & 1) * 8);
    }
}

typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int rows = 300, cols = 50, ld = 52;
    std::vector<double> h((size_t)rows * ld);
    for (int r = 0; r < rows; ++r) for (int c = 0; c < ld; ++c) h[(size_t)r * ld + c] = r * 1000.0 + c;
    double *d, *o;
    cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, 128 * 16 * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(e), (int)q, p);
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    const cuuint32_t box[2] = {16, 128};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = ((enc_fn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)cr);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 1024 + 64);
    for (int t = 0; t < 2; ++t) {
        const int c0 = t ? 40 : 8, r0 = t ? 200 : 64;
        probe<<<1, 256, 16384 + 1024 + 64>>>(m, c0, r0, o);
        e = cudaDeviceSynchronize();
        printf("kernel (c0=%d, r0=%d): %s\n", c0, r0, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<double> g(128 * 16);
        cudaMemcpy(g.data(), o, g.size() * 8, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r = 0; r < 128; ++r) for (int c = 0; c < 16; ++c) {
            const double want = (r0 + r < rows && c0 + c < cols) ? (r0 + r) * 1000.0 + c0 + c : 0.0;
            if (g[r * 16 + c] != want) { if (bad < 5) printf("  mismatch r=%d c=%d got %.1f want %.1f\n", r, c, g[r * 16 + c], want); ++bad; }
        }
        printf("  mismatches: %d\n", bad);
    }
    return 0;
}
