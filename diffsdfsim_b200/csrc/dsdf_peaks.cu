// Measured arithmetic peaks of the device the library runs on: dependent-FMA-chain microbenchmarks for the FP64 and
// FP32 pipes (SURVEY.md s8d asks the builder to measure these; the stepping kernels are FP64-latency / issue bound, so
// their roofline is quoted against the DFMA rate, not against HBM).  Eight independent chains per thread hide the FMA
// latency; the result is written back so the compiler cannot drop the loop.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dsdf_b200.h"

namespace dsdf {

template <class T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T a, T b) {
    T x0 = (T)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
        x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

template <class T> static int run_peak(int iters, void* scratch, double* tflops, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fma_peak_kernel<T><<<blocks, threads, 0, st>>>((T*)scratch, iters / 8 + 1, (T)0.999, (T)0.001);      // warm-up
    cudaEventRecord(e0, st);
    fma_peak_kernel<T><<<blocks, threads, 0, st>>>((T*)scratch, iters, (T)0.999, (T)0.001);
    cudaEventRecord(e1, st);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e != cudaSuccess) return (int)e;
    *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    return 0;
}

}  // namespace dsdf

// scratch: device buffer of at least 148 * 8 * 256 * 8 bytes (2.4 MB).  Results in TFLOP/s (host pointers).
extern "C" int dsdf_fma_peaks(int iters, void* scratch, double* fp64_tflops, double* fp32_tflops, void* stream) {
    if (iters <= 0 || !scratch) return -1;
    int rc = 0;
    if (fp64_tflops) rc = dsdf::run_peak<double>(iters, scratch, fp64_tflops, (cudaStream_t)stream);
    if (!rc && fp32_tflops) rc = dsdf::run_peak<float>(iters * 4, scratch, fp32_tflops, (cudaStream_t)stream);
    return rc;
}
