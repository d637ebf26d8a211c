// microbenchmark: dispatch rate of small tcgen05.mma (M = 128, N = 64 / 128) with A in tensor memory or shared memory,
// for the issue forms the fused logistic sweep could use.  One CTA; prints cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I riemann_b200/csrc scripts/peaks/mma_rate.cu -o scripts/peaks/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_gemm.cuh"
using namespace tc;

template <bool ACC>
__device__ __forceinline__ void mma_ts_w(bool bf16, uint32_t d, uint32_t a, uint32_t blo, uint32_t bhi, uint32_t idesc) {
    if (bf16)
        asm volatile("{\n.reg .pred pe, pa;\n.reg .b64 bd;\nelect.sync _|pe, 0xffffffff;\nsetp.ne.b32 pa, %5, 0;\nmov.b64 bd, {%2, %3};\n"
                     "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, pa;\n}\n" ::"r"(d), "r"(a), "r"(blo), "r"(bhi), "r"(idesc), "n"(ACC ? 1 : 0) : "memory");
    else
        asm volatile("{\n.reg .pred pe, pa;\n.reg .b64 bd;\nelect.sync _|pe, 0xffffffff;\nsetp.ne.b32 pa, %5, 0;\nmov.b64 bd, {%2, %3};\n"
                     "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, pa;\n}\n" ::"r"(d), "r"(a), "r"(blo), "r"(bhi), "r"(idesc), "n"(ACC ? 1 : 0) : "memory");
}
__device__ __forceinline__ void mma_ss_w(uint32_t d, uint32_t alo, uint32_t blo, uint32_t bhi, uint32_t idesc) {
    asm volatile("{\n.reg .pred pe, pa;\n.reg .b64 ad, bd;\nelect.sync _|pe, 0xffffffff;\nsetp.ne.b32 pa, 1, 0;\nmov.b64 ad, {%1, %3};\nmov.b64 bd, {%2, %3};\n"
                 "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %4, pa;\n}\n" ::"r"(d), "r"(alo), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
// single-thread forms (caller is one elected lane)
__device__ __forceinline__ void mma_ts_1(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n.reg .pred pa;\nsetp.ne.b32 pa, 1, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, pa;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}
// four MMAs per asm block, ONE elect; descriptor advanced inside the block
__device__ __forceinline__ void mma_ts_w4(uint32_t d, uint32_t a, uint32_t blo, uint32_t bhi, uint32_t idesc) {
    asm volatile("{\n.reg .pred pe, pa;\n.reg .b64 b0, b1, b2, b3;\n.reg .b32 l1, l2, l3, a1, a2, a3;\nelect.sync _|pe, 0xffffffff;\nsetp.ne.b32 pa, 1, 0;\n"
                 "add.u32 l1, %2, 2;\nadd.u32 l2, %2, 4;\nadd.u32 l3, %2, 6;\nadd.u32 a1, %1, 8;\nadd.u32 a2, %1, 16;\nadd.u32 a3, %1, 24;\n"
                 "mov.b64 b0, {%2, %3};\nmov.b64 b1, {l1, %3};\nmov.b64 b2, {l2, %3};\nmov.b64 b3, {l3, %3};\n"
                 "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], b0, %4, pa;\n"
                 "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [a1], b1, %4, pa;\n"
                 "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [a2], b2, %4, pa;\n"
                 "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [a3], b3, %4, pa;\n}\n" ::"r"(d), "r"(a), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}

__global__ void __launch_bounds__(640, 1) rate_kernel(long long* out, int reps, int load, float* sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int done;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); done = 0; }
    if (warp == 1) tmem_alloc(&slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp >= 4) {
        // background load of the pointwise kind: packed FMAs and MUFU (load = 1), plus tensor-memory loads (load = 2)
        float2 acc = make_float2(threadIdx.x * 1e-3f, 0.5f), c = make_float2(1.0001f, 0.9999f);
        float m = 0.3f + lane * 1e-3f;
        if (load > 0) {
            while (!done) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    acc = __ffma2_rn(acc, c, make_float2(1e-7f, -1e-7f));
                    acc = __ffma2_rn(acc, c, make_float2(-1e-7f, 1e-7f));
                    acc = __ffma2_rn(acc, c, make_float2(1e-7f, -1e-7f));
                    float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(m)); m = 0.5f * y + 0.1f;
                }
                if (load == 2) { float v[16]; tmem_ld_32x16(((uint32_t)((warp & 3) * 32) << 16) + 448, v); m += v[3] * 1e-30f; }
            }
        }
        if (acc.x + acc.y + m == 12345.0f) sink[threadIdx.x] = acc.x;
        __syncthreads();
        return;
    }
    if (warp != 0) { __syncthreads(); if (warp == 1) tmem_dealloc(0, 512); return; }
    const uint64_t dk0 = umma_desc_kmajor<128>(smem_u32(smem));
    const uint32_t blo = (uint32_t)dk0, bhi = (uint32_t)(dk0 >> 32);
    uint32_t phase = 0;
    auto finish = [&](int v, long long t0, int n) {
        asm volatile("{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        mbar_wait(&bar, phase); phase ^= 1;
        const long long t2 = clock64();
        if (lane == 0) { out[3 * v] = t1 - t0; out[3 * v + 1] = t2 - t0; out[3 * v + 2] = n; }
        __syncwarp();
    };
    const uint32_t i64 = umma_idesc_tf32(128, 64), i128 = umma_idesc_tf32(128, 128), i256 = umma_idesc_tf32(128, 256);
    const uint32_t b64 = umma_idesc_bf16(128, 64);
    // V0: TS tf32 N=64, per-MMA elect (the sweep's form), 16 MMAs per rep with runtime-looped operands
    { long long t0 = clock64(); for (int r = 0; r < reps; ++r) for (int k = 0; k < 16; ++k) mma_ts_w<true>(false, 384, (uint32_t)(k * 8), blo + (uint32_t)((k >> 2) * 512 + (k & 3) * 2), bhi, i64); finish(0, t0, reps * 16); }
    // V1: same, fully unrolled
    { long long t0 = clock64(); for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 16; ++k) mma_ts_w<true>(false, 384, (uint32_t)(k * 8), blo + (uint32_t)((k >> 2) * 512 + (k & 3) * 2), bhi, i64); } finish(1, t0, reps * 16); }
    // V2: TS bf16 N=64 unrolled
    { long long t0 = clock64(); for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 16; ++k) mma_ts_w<true>(true, 384, (uint32_t)(k * 8), blo + (uint32_t)((k >> 2) * 512 + (k & 3) * 2), bhi, b64); } finish(2, t0, reps * 16); }
    // V3: SS tf32 N=64 unrolled
    { long long t0 = clock64(); for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 16; ++k) mma_ss_w(384, blo + 2048 + (uint32_t)((k & 3) * 2), blo + (uint32_t)((k >> 2) * 512 + (k & 3) * 2), bhi, i64); } finish(3, t0, reps * 16); }
    // V4: TS tf32 N=128 unrolled
    { long long t0 = clock64(); for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 16; ++k) mma_ts_w<true>(false, 256, (uint32_t)(k * 8), blo + (uint32_t)((k >> 2) * 512 + (k & 3) * 2), bhi, i128); } finish(4, t0, reps * 16); }
    // V5: TS tf32 N=256 unrolled
    { long long t0 = clock64(); for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 16; ++k) mma_ts_w<true>(false, 256, (uint32_t)(k * 8), blo + (uint32_t)((k >> 2) * 512 + (k & 3) * 2), bhi, i256); } finish(5, t0, reps * 16); }
    // V6: TS tf32 N=64, four MMAs per asm block with one elect
    { long long t0 = clock64(); for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int b = 0; b < 4; ++b) mma_ts_w4(384, (uint32_t)(b * 32), blo + (uint32_t)(b * 512), bhi, i64); } finish(6, t0, reps * 16); }
    // V7: TS tf32 N=64, one thread issues (divergent branch), unrolled
    { long long t0 = clock64(); if (lane == 0) { for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 16; ++k) mma_ts_1(384, (uint32_t)(k * 8), dk0 + (uint64_t)((k >> 2) * 512 + (k & 3) * 2), i64); } } __syncwarp(); finish(7, t0, reps * 16); }
    // V8: TS tf32 N=64 unrolled, alternating two accumulators (is the RMW of one accumulator the limit?)
    { long long t0 = clock64(); for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 16; ++k) mma_ts_w<true>(false, 384 + (k & 1) * 64, (uint32_t)(k * 8), blo + (uint32_t)((k >> 2) * 512 + (k & 3) * 2), bhi, i64); } finish(8, t0, reps * 16); }
    // V9: SS tf32 N=256 unrolled (the dense GEMM's MMA) for reference
    { long long t0 = clock64(); for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 16; ++k) mma_ss_w(256, blo + 2048 + (uint32_t)((k & 3) * 2), blo + (uint32_t)((k & 3) * 2), bhi, i256); } finish(9, t0, reps * 16); }
    // V10: the fused sweep's GEMM1 exactly: 13 TF32 MMAs (the first one overwrites), then 7 x 2 bf16 MMAs with A from two
    // other TMEM regions and B from two other shared-memory regions; D alternates between two buffers per repetition
    { long long t0 = clock64(); for (int r = 0; r < reps; ++r) {
        const uint32_t tz = 384 + (uint32_t)(r & 1) * 64;
        mma_ts_w<false>(false, tz, 0, blo, bhi, i64);
#pragma unroll
        for (int k = 1; k < 13; ++k) mma_ts_w<true>(false, tz, (uint32_t)(k * 8), blo + (uint32_t)((k >> 2) * 512 + (k & 3) * 2), bhi, i64);
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const uint32_t off = (uint32_t)((k >> 2) * 512 + (k & 3) * 2);
            mma_ts_w<true>(true, tz, 128 + (uint32_t)(k * 8), blo + 2048 + off, bhi, b64);
            mma_ts_w<true>(true, tz, 192 + (uint32_t)(k * 8), blo + 3072 + off, bhi, b64);
        } } finish(10, t0, reps * 27); }
    // V11: the same with every MMA accumulating (no overwrite)
    { long long t0 = clock64(); for (int r = 0; r < reps; ++r) {
        const uint32_t tz = 384 + (uint32_t)(r & 1) * 64;
#pragma unroll
        for (int k = 0; k < 13; ++k) mma_ts_w<true>(false, tz, (uint32_t)(k * 8), blo + (uint32_t)((k >> 2) * 512 + (k & 3) * 2), bhi, i64);
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const uint32_t off = (uint32_t)((k >> 2) * 512 + (k & 3) * 2);
            mma_ts_w<true>(true, tz, 128 + (uint32_t)(k * 8), blo + 2048 + off, bhi, b64);
            mma_ts_w<true>(true, tz, 192 + (uint32_t)(k * 8), blo + 3072 + off, bhi, b64);
        } } finish(11, t0, reps * 27); }
    done = 1;
    __syncthreads();
}

int main() {
    long long* d; cudaMalloc(&d, 36 * 8); cudaMemset(d, 0, 36 * 8);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    const char* names[12] = {"TS tf32 N=64 looped", "TS tf32 N=64 unrolled", "TS bf16 N=64 unrolled", "SS tf32 N=64 unrolled", "TS tf32 N=128 unrolled",
                             "TS tf32 N=256 unrolled", "TS tf32 N=64 4-per-asm", "TS tf32 N=64 one thread", "TS tf32 N=64 two accumulators", "SS tf32 N=256 unrolled", "sweep GEMM1: 13 tf32 + 14 bf16", "same, all accumulating"};
    float* sink; cudaMalloc(&sink, 4096);
    for (int load = 0; load < 2; ++load) {
        for (int pass = 0; pass < 2; ++pass) {
            rate_kernel<<<1, 640, 70000>>>(d, 64, load, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        long long h[36]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("== background load %d (0 none, 1 FFMA2 + MUFU on 16 warps, 2 the same + tcgen05.ld)\n", load);
        for (int v = 0; v < 12; ++v) printf("%-32s issue %7.1f cycles/MMA   complete %7.1f cycles/MMA\n", names[v], (double)h[3 * v] / h[3 * v + 2], (double)h[3 * v + 1] / h[3 * v + 2]);
    }
    return 0;
}
