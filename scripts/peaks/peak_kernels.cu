// Peak-rate microbenchmarks for the roofline denominators MEASURED_PEAKS.json does not carry
// ("re-measure TF32 / FP64-tensor / FP32-FFMA peaks with the same method").
// Measurement infrastructure only -- not linked into libriemann_b200.so.
//   fp64 DFMA   : 8 independent FMA chains per thread, register resident
//   fp64 DMMA   : mma.sync.m8n8k4.f64, 8 independent accumulators per warp
//   fp32 FFMA   : 8 independent FMA chains per thread
// Each entry launches `blocks` x 256 threads for `iters` loop trips and returns the flop count.
#include <cuda_runtime.h>
#include <stdint.h>

template <typename T>
__global__ void __launch_bounds__(256) fma_kernel(T* out, int iters, T a, T b) {
    T x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll 8
        for (int u = 0; u < 8; ++u) {
            x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
            x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
        }
    }
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int q = 0; q < 8; ++q) { c[q][0] = threadIdx.x; c[q][1] = q; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[q][0]), "+d"(c[q][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += c[q][0] + c[q][1];
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s;
}

extern "C" double peak_launch(int which, int blocks, int iters, void* d_out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const double threads = (double)blocks * 256.0;
    if (which == 0) {          // fp64 DFMA: 64 FMA per trip per thread
        fma_kernel<double><<<blocks, 256, 0, st>>>((double*)d_out, iters, 0.999999, 1e-9);
        return threads * iters * 64.0 * 2.0;
    } else if (which == 1) {   // fp32 FFMA
        fma_kernel<float><<<blocks, 256, 0, st>>>((float*)d_out, iters, 0.999999f, 1e-9f);
        return threads * iters * 64.0 * 2.0;
    } else {                   // DMMA m8n8k4: 32 mma per trip per warp, 512 flop each
        dmma_kernel<<<blocks, 256, 0, st>>>((double*)d_out, iters, 0.999999, 1e-9);
        return (threads / 32.0) * iters * 32.0 * 512.0;
    }
}
