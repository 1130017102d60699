// fp64_peaks.cu -- measures the FP64 ceilings of the B200 this runs on (MEASURED_PEAKS.json has
// no FP64 entry): DMMA.8x8x4 issue rate, DFMA issue rate, and cuBLAS DGEMM throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/fp64_peaks.cu -lcublas -o tools/fp64_peaks
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void dmma_kernel(double *out, int iters)
{
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dfma_kernel(double *out, int iters)
{
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA with a mix of DFMA (1 DFMA per `ratio` DMMA) to see whether they share the pipe
template <int ILP>
__global__ void mix_kernel(double *out, int iters)
{
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c[ILP][2], d[4] = {0, 1, 2, 3};
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
            if (i % 4 == 0) d[(i / 4) % 4] = fma(d[(i / 4) % 4], a, b);
        }
    }
    double s = d[0] + d[1] + d[2] + d[3];
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f, int reps = 5)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
    }
    return best;
}

int main(int argc, char **argv)
{
    const char *path = argc > 1 ? argv[1] : "gpurun_out/fp64_peaks.json";
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    FILE *fp = fopen(path, "w");
    if (!fp) fp = stdout;
    fprintf(fp, "{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
    const int iters = 20000;
    // DMMA: warps per SM sweep
    for (int warps : {4, 8, 16, 32}) {
        const int blocks = sms, threads = warps * 32;
        float ms = time_ms([&] { dmma_kernel<16><<<blocks, threads>>>(out, iters); });
        double flops = 2.0 * 256 * 16 * (double)iters * warps * blocks;
        printf("DMMA ilp16 warps/SM=%2d : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
        fprintf(fp, ", \"dmma_tflops_w%d\": %.3f", warps, flops / ms * 1e-9);
    }
    {
        const int warps = 8;
        float ms = time_ms([&] { dmma_kernel<4><<<sms, warps * 32>>>(out, iters); });
        double flops = 2.0 * 256 * 4 * (double)iters * warps * sms;
        printf("DMMA ilp4  warps/SM= 8 : %8.3f ms  %7.2f TFLOP/s\n", ms, flops / ms * 1e-9);
        fprintf(fp, ", \"dmma_tflops_ilp4_w8\": %.3f", flops / ms * 1e-9);
        ms = time_ms([&] { dmma_kernel<32><<<sms, warps * 32>>>(out, iters); });
        flops = 2.0 * 256 * 32 * (double)iters * warps * sms;
        printf("DMMA ilp32 warps/SM= 8 : %8.3f ms  %7.2f TFLOP/s\n", ms, flops / ms * 1e-9);
        fprintf(fp, ", \"dmma_tflops_ilp32_w8\": %.3f", flops / ms * 1e-9);
    }
    for (int warps : {8, 16, 32}) {
        float ms = time_ms([&] { dfma_kernel<16><<<sms, warps * 32>>>(out, iters); });
        double flops = 2.0 * 32 * 16 * (double)iters * warps * sms;
        printf("DFMA ilp16 warps/SM=%2d : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
        fprintf(fp, ", \"dfma_tflops_w%d\": %.3f", warps, flops / ms * 1e-9);
    }
    {
        const int warps = 8;
        float ms = time_ms([&] { mix_kernel<16><<<sms, warps * 32>>>(out, iters); });
        double flops = 2.0 * 256 * 16 * (double)iters * warps * sms;
        printf("DMMA+DFMA(1:4) warps=8 : %8.3f ms  %7.2f TFLOP/s (DMMA flops only)\n", ms, flops / ms * 1e-9);
        fprintf(fp, ", \"dmma_tflops_with_dfma_1in4_w8\": %.3f", flops / ms * 1e-9);
    }
    // cuBLAS DGEMM
    {
        cublasHandle_t h; cublasCreate(&h);
        for (int n : {4096, 8192}) {
            double *A, *B, *C;
            CK(cudaMalloc(&A, sizeof(double) * n * n)); CK(cudaMalloc(&B, sizeof(double) * n * n));
            CK(cudaMalloc(&C, sizeof(double) * n * n));
            CK(cudaMemset(A, 0, sizeof(double) * n * n)); CK(cudaMemset(B, 0, sizeof(double) * n * n));
            const double one = 1.0, zero = 0.0;
            float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); }, 5);
            double tf = 2.0 * n * (double)n * n / ms * 1e-9;
            printf("cuBLAS DGEMM %d^3       : %8.3f ms  %7.2f TFLOP/s (burst, best of 5)\n", n, ms, tf);
            fprintf(fp, ", \"cublas_dgemm_%d_tflops\": %.3f", n, tf);
            if (n == 8192) {
                // sustained: back to back for ~3 s
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                int reps = std::max(3, (int)(3000.0f / ms));
                cudaEventRecord(e0);
                for (int r = 0; r < reps; ++r) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float tot; cudaEventElapsedTime(&tot, e0, e1);
                double tfs = 2.0 * n * (double)n * n * reps / tot * 1e-9;
                printf("cuBLAS DGEMM %d^3       : sustained %7.2f TFLOP/s over %d reps\n", n, tfs, reps);
                fprintf(fp, ", \"cublas_dgemm_8192_tflops_sustained\": %.3f", tfs);
            }
            cudaFree(A); cudaFree(B); cudaFree(C);
        }
        cublasDestroy(h);
    }
    // sustained DMMA (3 s)
    {
        const int warps = 8;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float one = time_ms([&] { dmma_kernel<16><<<sms, warps * 32>>>(out, iters); }, 2);
        int reps = std::max(3, (int)(3000.0f / one));
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) dmma_kernel<16><<<sms, warps * 32>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float tot; cudaEventElapsedTime(&tot, e0, e1);
        double tf = 2.0 * 256 * 16 * (double)iters * warps * sms * reps / tot * 1e-9;
        printf("DMMA sustained (3 s)    : %7.2f TFLOP/s\n", tf);
        fprintf(fp, ", \"dmma_tflops_sustained\": %.3f", tf);
    }
    fprintf(fp, "}\n");
    if (fp != stdout) fclose(fp);
    return 0;
}
