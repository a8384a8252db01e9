// Standalone run of sdv_gemm_tma_kernel on a synthetic banded product, checked against the host.
// nvcc -gencode arch=compute_100a,code=sm_100a -Ipttools_b200/csrc -Iinclude [-DGT_DBG=n] -o gemm_tma_probe tools/probes/gemm_tma_probe.cu
#include "../../pttools_b200/csrc/ptt_sdv_gemm.cu"
#include <vector>
namespace ptt { char* error_buffer() { static char b[512]; return b; } }

int main() {
    const int n_z = 300, n_p = 200, n_q = 548, kw = 384, n_tiles = 3, n_tn = 2;
    const long long a2_ld = 3648;
    std::vector<double> hom((size_t)n_tiles * 128 * kw), ha2((size_t)n_p * a2_ld, 0.0);
    for (size_t i = 0; i < hom.size(); ++i) hom[i] = 1.0 + (double)(i % 97) * 0.01;
    for (int p = 0; p < n_p; ++p) for (int k = 0; k < n_q; ++k) ha2[(size_t)p * a2_ld + k] = 0.5 + (double)((p * 31 + k * 7) % 89) * 0.02;
    std::vector<int> heb = {61, 189, 318};
    double *om, *a2, *out; int* eb;
    cudaMalloc(&om, hom.size() * 8); cudaMalloc(&a2, ha2.size() * 8); cudaMalloc(&out, (size_t)n_p * n_z * 8); cudaMalloc(&eb, 12);
    cudaMemcpy(om, hom.data(), hom.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(a2, ha2.data(), ha2.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(eb, heb.data(), 12, cudaMemcpyHostToDevice);
    cudaMemset(out, 0, (size_t)n_p * n_z * 8);
    CUtensorMap m1, m2;
    printf("maps: %d %d\n", (int)ptt::gt_make_map(&m1, om, kw, n_tiles * 128, kw), (int)ptt::gt_make_map(&m2, a2, n_q, n_p, a2_ld));
    cudaFuncSetAttribute(ptt::sdv_gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ptt::GT_SMEM);
    ptt::sdv_gemm_tma_kernel<<<n_tiles * n_tn, ptt::GT_THREADS, ptt::GT_SMEM>>>(m1, m2, n_z, n_p, kw, n_tn, 0, eb, 2.0, out, n_z);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<double> g((size_t)n_p * n_z);
    cudaMemcpy(g.data(), out, g.size() * 8, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int p = 0; p < n_p; p += 7) for (int i = 0; i < n_z; i += 5) {
        double acc = 0;
        const int t = i / 128;
        for (int k = 0; k < kw; ++k) { const int e2 = heb[t] + k; if (e2 < n_q) acc += hom[(size_t)i * kw + k] * ha2[(size_t)p * a2_ld + e2]; }
        const double d = fabs(g[(size_t)p * n_z + i] / (2.0 * acc) - 1);
        if (d > worst) worst = d;
    }
    printf("max rel dev vs host: %.3e\n", worst);
    return 0;
}
