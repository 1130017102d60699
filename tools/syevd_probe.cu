// syevd_probe.cu -- how long does the per-step eigensolve take on this GPU, and is there a faster cuSOLVER
// entry point than the legacy Dsyevd?  (Bench tooling only; the library calls cusolverDnDsyevd.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/syevd_probe.cu -lcusolver -lcublas -o tools/syevd_probe
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

int main(int argc, char **argv)
{
    cusolverDnHandle_t h; cusolverDnCreate(&h);
    cusolverDnParams_t prm; cusolverDnCreateParams(&prm);
    printf("{");
    bool first = true;
    for (int n : {576, 1000, 2000, 4000}) {
        std::vector<double> A((size_t)n * n);
        std::mt19937_64 rng(n);
        std::normal_distribution<double> nd;
        for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) { double v = nd(rng); A[(size_t)i * n + j] = v; A[(size_t)j * n + i] = v; }
        double *dA, *dA0, *dW; int *dInfo;
        CK(cudaMalloc(&dA, sizeof(double) * n * n)); CK(cudaMalloc(&dA0, sizeof(double) * n * n));
        CK(cudaMalloc(&dW, sizeof(double) * n)); CK(cudaMalloc(&dInfo, sizeof(int)));
        CK(cudaMemcpy(dA0, A.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        // legacy API
        int lwork = 0;
        cusolverDnDsyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, dA, n, dW, &lwork);
        double *dWork; CK(cudaMalloc(&dWork, sizeof(double) * lwork));
        float best = 1e30f;
        for (int r = 0; r < 3; ++r) {
            CK(cudaMemcpy(dA, dA0, sizeof(double) * n * n, cudaMemcpyDeviceToDevice));
            cudaEventRecord(e0);
            cusolverDnDsyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, dA, n, dW, dWork, lwork, dInfo);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms;
        }
        printf("%s\"dsyevd_N%d_ms\": %.3f", first ? "" : ", ", n, best); first = false;
        // 64-bit generic API
        size_t wsD = 0, wsH = 0;
        cusolverDnXsyevd_bufferSize(h, prm, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, dA, n,
                                    CUDA_R_64F, dW, CUDA_R_64F, &wsD, &wsH);
        void *dW2 = nullptr; CK(cudaMalloc(&dW2, wsD ? wsD : 1));
        std::vector<char> hW(wsH ? wsH : 1);
        best = 1e30f;
        for (int r = 0; r < 3; ++r) {
            CK(cudaMemcpy(dA, dA0, sizeof(double) * n * n, cudaMemcpyDeviceToDevice));
            cudaEventRecord(e0);
            cusolverDnXsyevd(h, prm, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, dA, n, CUDA_R_64F, dW,
                             CUDA_R_64F, dW2, wsD, hW.data(), wsH, dInfo);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms;
        }
        printf(", \"xsyevd_N%d_ms\": %.3f", n, best);
        // H2D / D2H of the matrices involved at this size (pinned)
        double *hp; CK(cudaMallocHost(&hp, sizeof(double) * n * n));
        cudaEventRecord(e0); CK(cudaMemcpyAsync(dA, hp, sizeof(double) * n * n, cudaMemcpyHostToDevice)); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf(", \"h2d_N%d_ms\": %.3f", n, ms);
        cudaFreeHost(hp);
        cudaFree(dA); cudaFree(dA0); cudaFree(dW); cudaFree(dInfo); cudaFree(dWork); cudaFree(dW2);
    }
    printf("}\n");
    return 0;
}
