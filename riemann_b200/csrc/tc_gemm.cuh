// riemann_b200 -- tcgen05 / TMA / TMEM building blocks for the fp32-accurate (3xTF32) tensor path.
//
//   C[m][n] = sum_k A[m][k] B[n][k],   A = Ah + Al,  B = Bh + Bl  (fp32 values split so that the
//   "hi" parts are exactly representable in TF32: 10 explicit mantissa bits)
//   C ~= Ah Bh^T + Ah Bl^T + Al Bh^T   accumulated in fp32 in tensor memory (the dropped
//   Al Bl^T term is 2^-22 relative), i.e. fp32-level accuracy from three TF32 MMAs.
//
// Operands are K-major fp32 tiles [rows][32] (one 128-byte swizzle row per tile row) written to
// shared memory by TMA (cp.async.bulk.tensor, SWIZZLE_128B) and read by tcgen05.mma.kind::tf32
// through shared-memory matrix descriptors; accumulators live in TMEM and come back through
// tcgen05.ld.  One elected thread issues TMA, one elected thread issues MMA, four warps run the
// epilogue.  Written directly in PTX (no CUTLASS templates); encodings follow the PTX ISA
// tcgen05 matrix-descriptor / instruction-descriptor tables.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

constexpr int TM = 128;          // rows of A per CTA  (UMMA M)
constexpr int TN = 256;          // rows of B per CTA  (UMMA N)
constexpr int TK = 32;           // fp32 per k-block = 128 bytes = one swizzle row (single-pass kernels)
#ifndef RMN_TC_TK3
#define RMN_TC_TK3 16
#endif
// k-block of the 3-pass kernels: 16 fp32 = 64-byte rows (SWIZZLE_64B), so a stage (Ah Al Bh Bl) is 48 KB and
// FOUR stages fit -- with 96-KB stages only two did, and the TMA latency of a stage (~2,500 cycles) was longer
// than the 1,600 cycles of tensor work it fed
constexpr int TK3 = RMN_TC_TK3;
constexpr int UK = 8;            // K per tcgen05.mma.kind::tf32
constexpr int STAGES = 2;
constexpr int A_BYTES = TM * TK * 4;                 // 16 KB
constexpr int B_BYTES = TN * TK * 4;                 // 32 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // Ah Al Bh Bl = 96 KB
constexpr int EPI_STAGE_FLOATS = 32 * 20;      // per epilogue warp: 32 rows x 16 columns (+4 pad) for the coalescing transpose
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 8 * EPI_STAGE_FLOATS * 4;
constexpr int THREADS = 384;     // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 epilogue (2 per TMEM lane quarter)
constexpr int EPI_WARPS = 8;
constexpr int ACC_STAGES = 2;    // accumulator double buffering: epilogue(i) overlaps mainloop(i+1)
constexpr int TMEM_COLS = ACC_STAGES * TN;   // 512 = all of TMEM

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// shared-memory matrix descriptor: K-major, rows of ROWB = 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B) bytes,
// 8-row groups 8 * ROWB bytes apart
template <int ROWB>
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr) {
    static_assert(ROWB == 128 || ROWB == 64, "row bytes = swizzle span");
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address, 16-byte units   [0,14)
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major) [16,30)
    d |= (uint64_t)((8 * ROWB) >> 4) << 32;             // stride byte offset = 8 rows * ROWB            [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)                   [46,48)
    d |= (uint64_t)(ROWB == 128 ? 2 : 4) << 61;         // layout type: SWIZZLE_128B = 2, SWIZZLE_64B = 4 [61,64)
    return d;
}
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t saddr) { return umma_desc_kmajor<128>(saddr); }
// instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) /*c = f32*/ | (2u << 7) /*a = tf32*/ | (2u << 10) /*b = tf32*/ |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f16 with bf16 operands, fp32 accumulate (K = 16 per instruction): the proposal-only Fisher-metric GEMM
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) /*c = f32*/ | (1u << 7) /*a = bf16*/ | (1u << 10) /*b = bf16*/ |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane base + t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// split an fp32 value into a TF32-exact "hi" (low 13 mantissa bits cleared) and the fp32 remainder
__host__ __device__ inline void split_tf32(float x, float& hi, float& lo) {
#ifdef __CUDA_ARCH__
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
#else
    union { float f; uint32_t u; } c; c.f = x; c.u &= 0xFFFFE000u; hi = c.f;
#endif
    lo = x - hi;
}

struct GemmMaps {
    CUtensorMap ah, al, bh, bl;
};

// host: 2-D bf16 row-major [rows][cols] tensor map, box = [box_rows][64] (one 128-byte swizzle row), OOB -> 0
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows);
// host: 2-D fp32 row-major [rows][cols] tensor map, box = [box_rows][tk], tk = 32 (SWIZZLE_128B, single-pass
// kernels) or TK3 (3-pass kernels; SWIZZLE_64B when 16), OOB -> 0
int make_tmap_2d(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_rows, int tk = TK);

}  // namespace tc
