// kernels_probe.cu -- in-run measurement of the FP64 tensor-pipe ceiling (the roofline denominator).
//
// MEASURED_PEAKS.json carries HBM and bf16 only; the FP64 ceiling of this path is the issue rate of
// DMMA.8x8x4 (mma.sync.m8n8k4.f64), the instruction jac_dmma_kernel is built on.  bench.py calls
// qcdmet_measure_fp64_peak() in the same process, right after the timed region, so the denominator of
// roofline.frac is measured on the same GPU, at the same clocks, in the driver's own run.
#include "common.cuh"
#include <algorithm>

namespace qcdmet {

namespace {
constexpr int PROBE_ILP = 16;      // independent accumulator pairs per thread
constexpr int PROBE_WARPS = 8;     // warps per SM: two per sub-partition, as the Jacobian kernel runs

__global__ void __launch_bounds__(PROBE_WARPS * 32) dmma_probe_kernel(double *out, int iters)
{
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c[PROBE_ILP][2];
#pragma unroll
    for (int i = 0; i < PROBE_ILP; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < PROBE_ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < PROBE_ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

}  // namespace qcdmet

using namespace qcdmet;

/* Sustained DMMA.8x8x4 rate of `device` (< 0: current) in TFLOP/s, measured for about `seconds` (0.05 .. 5)
 * with one 8-warp CTA per SM and 16 independent accumulators per thread (the configuration of
 * tools/fp64_peaks.cu, whose committed result is profiles/fp64_peaks.json). */
extern "C" int qcdmet_measure_fp64_peak(int device, double seconds, double *tflops)
{
    QC_CHECK(tflops != nullptr, QCDMET_E_ARG, "tflops is NULL");
    *tflops = 0.0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device available");
        return QCDMET_E_CUDA;
    }
    int prev = -1;
    QC_CUDA(cudaGetDevice(&prev));
    if (device < 0) device = prev;
    QC_CHECK(device < ndev, QCDMET_E_ARG, "device %d out of range", device);
    QC_CUDA(cudaSetDevice(device));
    int rc = QCDMET_OK;
    double *out = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    do {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { rc = QCDMET_E_CUDA; break; }
        const int sms = prop.multiProcessorCount, threads = PROBE_WARPS * 32;
        if (cudaMalloc((void **)&out, sizeof(double) * sms * threads) != cudaSuccess) { rc = QCDMET_E_ALLOC; break; }
        if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { rc = QCDMET_E_CUDA; break; }
        const int iters = 20000;
        const double flopsPerLaunch = 2.0 * 256 * PROBE_ILP * (double)iters * PROBE_WARPS * sms;
        // calibrate: one launch
        dmma_probe_kernel<<<sms, threads>>>(out, iters);
        if (cudaDeviceSynchronize() != cudaSuccess) { rc = QCDMET_E_CUDA; break; }
        cudaEventRecord(e0);
        dmma_probe_kernel<<<sms, threads>>>(out, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { rc = QCDMET_E_CUDA; break; }
        float one = 0.f;
        cudaEventElapsedTime(&one, e0, e1);
        seconds = std::min(std::max(seconds, 0.05), 5.0);
        const int reps = std::max(3, (int)(seconds * 1e3 / std::max(one, 1e-3f)));
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) dmma_probe_kernel<<<sms, threads>>>(out, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { rc = QCDMET_E_CUDA; break; }
        float tot = 0.f;
        cudaEventElapsedTime(&tot, e0, e1);
        *tflops = flopsPerLaunch * reps / (tot * 1e-3) * 1e-12;
    } while (0);
    if (rc != QCDMET_OK) { cudaGetLastError(); set_error("FP64 peak probe failed on device %d", device); }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(out);
    cudaSetDevice(prev);
    return rc;
}
