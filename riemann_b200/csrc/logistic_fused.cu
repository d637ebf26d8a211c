// riemann_b200 -- fused tcgen05 likelihood sweep of the logistic-regression model:
// ONE kernel per sweep, "flash-attention shaped".  Z and R never touch HBM.
//
//   log L(theta_c) = sum_i y_i z_ci - softplus(z_ci),   z_ci = theta_c . x_i        (model, absent in the reference;
//   grad_c         = sum_i (y_i - sigmoid(z_ci)) x_i                                  Model protocol model.py:27-55,
//   (mMALA)  w_ci  = p_ci (1 - p_ci)                                                  driven by hamiltonian.py:76-91)
//
// A CTA owns 128 chains (one TMEM lane each) and a contiguous range of data rows, streamed as tiles of 64 rows:
//
//   TMA thread    ONE ring of SA slots; a slot holds the 64-row tile three times over the same rows:
//                   Xh   fp32 (x rounded to nearest TF32)        [64][dp32]   SWIZZLE_128B boxes of 32 columns
//                   Xlb  fp16 ((x - Xh) 2^10)                    [64][dp32]   SWIZZLE_128B boxes of 64 columns
//                   Xhb  fp16 (x 2^-10)                          [64][dp32]   SWIZZLE_128B boxes of 64 columns
//                 8 bytes per element, all of it read by BOTH products: the slot is released by GEMM2.
//   MMA warp      GEMM1  Z[128 x 64] = Theta_h Xh^T  (kind::tf32)  +  Theta_hb Xlb^T + Theta_lb Xhb^T  (kind::f16, fp16):
//                        the two correction terms are 2^-11 of the product, so fp16 operands (2^-12 relative) leave
//                        z fp32-accurate at HALF the instruction count of a TF32 correction (K = 16 per MMA).
//                        A = Theta from TENSOR MEMORY (written once per CTA; fp16 parts packed two per column)
//                 GEMM2  G[128 x dp32] += R[128 x 64] Xhb[64 x dp32]   (kind::f16: the gradient only shapes the
//                        proposal); A = R as fp16 pairs from tensor memory, written in place over Z by the pointwise
//                        warps; B = the SAME Xhb boxes read MN-major -- a 16-bit operand may be MN-major in the plain
//                        128-byte swizzle, so the tile TMA wrote K-major for GEMM1 is, read with the other descriptor,
//                        the transposed operand of GEMM2.  (Round 2's first version kept the gradient in TF32, whose
//                        only MN-major layout is the 32-byte-atom swizzle: a second shared-memory copy of the tile in
//                        a second ring, 96 KB of L2 -> SM traffic per tile instead of 64 -- and that traffic, not the
//                        tensor pipe, was what bounded the sweep once the correction MMAs were halved.)
//   4 x 4 warps   pointwise stage, all four warpgroups on every tile (16 of its 64 rows each): tcgen05.ld the logits, fp32
//                 sigmoid / softplus (one MUFU.EX2, one MUFU.RCP and a degree-9 polynomial for log1p per element, two
//                 elements per instruction with the packed fp32 FMA of sm_100),
//                 log-likelihood partial sums in fp64, R = y - p rounded to fp16 -> tcgen05.st back into the first 8 of
//                 the warpgroup's 16 TMEM columns (and W = p(1-p) to HBM for the mMALA metric GEMM)
//
//   TMEM columns: Theta_h [0, dp32) | Theta_hb, Theta_lb (dp32 / 2 each) | G (dp32) | Z/R buffer 0, 1 (64 each)
//
// GEMM1 and GEMM2 are issued by two warps (1 and 3): the accumulators differ, so the order in which the tensor pipe takes the
// two instruction streams cannot change the result.  GEMM1 of tile t+1 runs while the pointwise warps process tile t;
// GEMM1 of tile t+2 reuses the Z buffer GEMM2 of tile t reads R from and waits for that GEMM2's commit (z_free).
//
// Measured dead end (gpurun r2k): releasing the ring box by box (eight one-box slots, TMA two tiles ahead) was SLOWER, 2.35 ms
// per 1,024-chain sweep against 1.77 -- four commits and four barrier waits per tile cost the MMA thread more than the
// exposed TMA latency it hid.
//
// Accuracy.  Xh / Theta_h are rounded to nearest TF32 and the remainders to (scaled) fp16, so the pair carries 22 bits and
// the dropped terms are ~2^-24 of a product; the per-row terms are fp32 (|error| ~1e-7 each), summed in fp64.  R and x
// enter the gradient product as fp16 (2^-12 relative).  Measured budget: tests/test_gpu_logistic.py (riemann_b200/budgets.py).
// The result is a deterministic function of theta (fixed tile order, no atomics).
#include <algorithm>
#include "common.cuh"
#include "tc_gemm.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "logistic_fused.cuh"
#include <stdlib.h>

namespace lgf {
using namespace tc;

namespace {

constexpr int NT = 64;                     // data rows per tile = UMMA N of GEMM1 = K extent of GEMM2
constexpr int CB = 128;                    // chains per CTA = UMMA M
constexpr int BOX_BYTES = NT * 128;        // one TMA box: 64 rows x 32 fp32
constexpr int PWG = 4;                     // pointwise warpgroups: each takes NT / PWG = 16 of a tile's 64 rows
constexpr int PCOLS = NT / PWG;
constexpr int THREADS = 128 + 128 * PWG;   // warp 0 TMA, 1 MMA (GEMM1), 2 TMEM alloc, 3 MMA (GEMM2), then PWG x 4 pointwise warps
constexpr int TMEM_COLS_ALLOC = 512;

template <int DP32> struct Cfg {
    static constexpr int NBOX = DP32 / 32;                         // fp32 boxes (32 columns) of the Xh part
    static constexpr int NB16 = DP32 / 64;                         // fp16 boxes (64 columns) of the Xlb / Xhb parts
    static constexpr int OFF_LB = NBOX * BOX_BYTES;                // slot: Xh | Xlb | Xhb
    static constexpr int OFF_HB = OFF_LB + NB16 * BOX_BYTES;
    static constexpr int A_BYTES = OFF_HB + NB16 * BOX_BYTES;      // 64 KB (dp32 = 128) / 32 KB (64): 8 bytes per element
    static constexpr int SA = (DP32 == 128) ? 3 : 6;               // ring slots
    static constexpr int TXA = A_BYTES;
    static constexpr int OFF_BAR = SA * A_BYTES;
    static constexpr int SMEM = OFF_BAR + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int COL_TH = 0, COL_TB = DP32, COL_LB = DP32 + DP32 / 2, COL_G = 2 * DP32, COL_Z = 3 * DP32;
    static_assert(3 * DP32 + 2 * NT <= 512, "TMEM columns");
    static_assert(SMEM <= 232448, "shared memory");
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// MMA issue for a warp in UNIFORM control flow: every lane executes the statement, `elect.sync` picks the one that issues.
// The B descriptor comes as two 32-bit words (only the start-address field in the low word changes from MMA to MMA, so
// advancing it is ONE integer add) and the accumulate flag is a literal.  With warp-uniform operands the compiler keeps
// them in uniform registers and an MMA costs two or three instructions instead of a divergent issue sequence.
template <bool ACC>
__device__ __forceinline__ void umma_tf32_ts_w(uint32_t tmem_d, uint32_t tmem_a, uint32_t bdesc_lo, uint32_t bdesc_hi, uint32_t idesc) {
    asm volatile(
        "{\n"
        ".reg .pred pe, pa;\n"
        ".reg .b64 bd;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "setp.ne.b32 pa, %5, 0;\n"
        "mov.b64 bd, {%2, %3};\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, pa;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "n"(ACC ? 1 : 0) : "memory");
}
// the same issue form for kind::f16 (16-bit operands, K = 16 per instruction): A is 16 packed pairs per lane = 8 TMEM columns
template <bool ACC>
__device__ __forceinline__ void umma_f16_ts_w(uint32_t tmem_d, uint32_t tmem_a, uint32_t bdesc_lo, uint32_t bdesc_hi, uint32_t idesc) {
    asm volatile(
        "{\n"
        ".reg .pred pe, pa;\n"
        ".reg .b64 bd;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "setp.ne.b32 pa, %5, 0;\n"
        "mov.b64 bd, {%2, %3};\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, pa;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "n"(ACC ? 1 : 0) : "memory");
}
// MMA issue with the issuing lane chosen ONCE (elect_leader) instead of an elect.sync per instruction
__device__ __forceinline__ uint32_t elect_leader() {
    uint32_t r;
    asm volatile("{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\nselp.u32 %0, 1, 0, pe;\n}\n" : "=r"(r));
    return r;
}
template <bool ACC, bool BF16>
__device__ __forceinline__ void umma_ts_l(uint32_t leader, uint32_t tmem_d, uint32_t tmem_a, uint32_t bdesc_lo, uint32_t bdesc_hi, uint32_t idesc) {
    if (BF16)
        asm volatile("{\n.reg .pred pe, pa;\n.reg .b64 bd;\nsetp.ne.b32 pe, %6, 0;\nsetp.ne.b32 pa, %5, 0;\nmov.b64 bd, {%2, %3};\n"
                     "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, pa;\n}\n"
                     ::"r"(tmem_d), "r"(tmem_a), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "n"(ACC ? 1 : 0), "r"(leader) : "memory");
    else
        asm volatile("{\n.reg .pred pe, pa;\n.reg .b64 bd;\nsetp.ne.b32 pe, %6, 0;\nsetp.ne.b32 pa, %5, 0;\nmov.b64 bd, {%2, %3};\n"
                     "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, pa;\n}\n"
                     ::"r"(tmem_d), "r"(tmem_a), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "n"(ACC ? 1 : 0), "r"(leader) : "memory");
}
__device__ __forceinline__ void umma_commit_l(uint32_t leader, uint64_t* bar) {
    asm volatile("{\n.reg .pred pe;\nsetp.ne.b32 pe, %1, 0;\n"
                 "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(bar)), "r"(leader) : "memory");
}
// all THREADS of the CTA, from either of the two places the roles end at (a named barrier counts arrivals, not places)
__device__ __forceinline__ void cta_sync_all() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n"
        ".reg .pred pe;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
        "}\n" ::"r"(smem_u32(bar)) : "memory");
}

// shared-memory matrix descriptor, MN-major, 16-bit elements, plain 128-byte swizzle (layout type 2): 64 elements of the
// MN index are contiguous (one 128-byte row), MN blocks of 64 are `lbo` bytes apart; the K index walks the 128-byte rows,
// groups of 8 rows (one swizzle pattern) are `sbo` bytes apart.  This is the box TMA writes for a [rows = K][64 columns =
// MN] 16-bit tile with CU_TENSOR_MAP_SWIZZLE_128B -- the very tile that is a K-major operand when rows are taken as MN.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_b16(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ float rn_tf32(float x) {          // round to nearest TF32 (10 explicit mantissa bits)
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
// two fp16 values in one 32-bit word, `even` in the low half: consecutive along the contraction of a kind::f16 MMA,
// both in a shared-memory row (little endian) and in a tensor-memory column of an A operand
__device__ __forceinline__ uint32_t pack_f16(float even, float odd) {
    const __half2 h2 = __floats2half2_rn(even, odd);                     // .x (low half) = even
    return *reinterpret_cast<const uint32_t*>(&h2);
}
// The 16-bit operands are fp16 with power-of-two scales that cancel in every product: the remainders x - Xh, theta - Th
// (<= 2^-12 of the value) are stored times 2^10, the values they multiply times 2^-10.  Both factors then sit in fp16's
// normal range for |x|, |theta| between ~0.06 and ~6e4 (smaller ones lose relative, not absolute, accuracy), and the
// correction terms keep 11 bits: their error is 2^-24 of a product, what TF32 remainders gave (bf16 operands: 2^-21).
constexpr float CS_UP = 1024.0f, CS_DN = 1.0f / 1024.0f;
constexpr uint32_t umma_idesc_f16(int M, int N) {                        // kind::f16, fp16 A and B, fp32 accumulate
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// log1p(e) for two elements at once, e in [0, 1]: degree-9 Chebyshev interpolant in t = 2e - 1, Horner form with the
// packed fp32 FMA of sm_100 (FFMA2: one instruction, two lanes); |error| < 7e-8
__device__ __forceinline__ float2 f2(float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 log1p_unit2(float2 e) {
    const float2 t = __ffma2_rn(f2(2.0f), e, f2(-1.0f));
    float2 p = f2(7.152816828e-06f);
    p = __ffma2_rn(p, t, f2(-2.401530172e-05f));
    p = __ffma2_rn(p, t, f2(6.393497408e-05f));
    p = __ffma2_rn(p, t, f2(-2.240642703e-04f));
    p = __ffma2_rn(p, t, f2(8.235513423e-04f));
    p = __ffma2_rn(p, t, f2(-3.088083925e-03f));
    p = __ffma2_rn(p, t, f2(1.234561512e-02f));
    p = __ffma2_rn(p, t, f2(-5.555534548e-02f));
    p = __ffma2_rn(p, t, f2(3.333333346e-01f));
    p = __ffma2_rn(p, t, f2(4.054651039e-01f));
    return p;
}

template <int DP32, bool HASW>
__global__ void __launch_bounds__(THREADS, 1)
lg_fused_sweep_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xlb,
                      const __grid_constant__ CUtensorMap map_xhb, SweepArgs a) {
    using C = Cfg<DP32>;
    constexpr int SA = C::SA;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* fullA = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* emptyA = fullA + SA;
    uint64_t* z_full = emptyA + SA;              // [2] GEMM1 of a tile complete
    uint64_t* r_full = z_full + 2;               // [2] pointwise stage wrote R
    uint64_t* z_free = r_full + 2;               // [2] GEMM2 of a tile complete: its R (= the Z buffer) may be overwritten
    uint64_t* th_ready = z_free + 2;             // Theta is in TMEM
    uint64_t* g_full = th_ready + 1;             // last GEMM2 complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cb = blockIdx.x % a.nblk, rs = blockIdx.x / a.nblk;
    const int64_t t_begin = (int64_t)rs * a.tps;
    const int64_t t_end = min(a.tiles_total, t_begin + a.tps);
    const int ntile = (int)max((int64_t)0, t_end - t_begin);

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_xh); tma_prefetch_desc(&map_xlb); tma_prefetch_desc(&map_xhb); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < SA; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&z_full[b], 1); mbar_init(&r_full[b], 4 * PWG); mbar_init(&z_free[b], 1); }
        mbar_init(th_ready, 4);
        mbar_init(g_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS_ALLOC);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Role dispatch.  Every warp but the GEMM1 issuer works inside this branch and RETURNS from it; the GEMM1 warp's loop is
    // the kernel's tail.  Code that no other thread can rejoin is convergent for the compiler, so it keeps the descriptors
    // and tensor-memory addresses of the MMAs in uniform registers and advances them on the uniform datapath: 3-4
    // instructions per MMA instead of the ~12 (R2UR, elect, vote, ...) it emits for an MMA inside a divergent role branch,
    // which, sharing a scheduler with four pointwise warps, held GEMM1 to ~70 cycles per MMA where the tensor pipe needs 32
    // (scripts/peaks/mma_rate.cu).
    if (warp != 1) {
    if (warp == 0 && lane == 0) {
        // ===== TMA producer: Xh | Xlb | Xhb of a tile into one slot (released by GEMM2) =====
        for (int t = 0; t < ntile; ++t) {
            const int s = t % SA;
            mbar_wait(&emptyA[s], ((t / SA) & 1) ^ 1);
            uint8_t* st = smem + s * C::A_BYTES;
            const int row0 = (int)((t_begin + t) * NT);
            mbar_expect_tx(&fullA[s], C::TXA);
#pragma unroll
            for (int b = 0; b < C::NBOX; ++b) tma_load_2d(st + b * BOX_BYTES, &map_xh, &fullA[s], b * 32, row0);
#pragma unroll
            for (int b = 0; b < C::NB16; ++b) {
                tma_load_2d(st + C::OFF_LB + b * BOX_BYTES, &map_xlb, &fullA[s], b * 64, row0);
                tma_load_2d(st + C::OFF_HB + b * BOX_BYTES, &map_xhb, &fullA[s], b * 64, row0);
            }
        }
    } else if (warp == 3) {
        // ===== second MMA issuer: GEMM2.  Its own warp on another scheduler -- an MMA costs its issuing warp ~12
        // instructions (descriptor arithmetic, R2UR, elect / vote), and with four pointwise warps on the same scheduler
        // the GEMM1 warp needs ~70 cycles per MMA where the tensor pipe needs 32 (scripts/peaks/mma_rate.cu); GEMM2 on that
        // warp too put 300 more cycles per tile on the critical chain.  Different accumulators (Z / G), so the order
        // in which the tensor pipe takes the two streams does not change a bit; the one hazard -- GEMM1 of tile t + 2
        // overwrites the Z buffer GEMM2 of tile t reads R from -- is covered by z_free.
        const uint32_t idesc2 = umma_idesc_f16(CB, DP32) | (1u << 16);           // B is MN-major
        constexpr uint32_t t_g = C::COL_G;
        const uint64_t dm0 = umma_desc_mnmajor_b16(0, BOX_BYTES, 1024);
        const uint32_t dm_hi = (uint32_t)(dm0 >> 32), dm_lo0 = (uint32_t)dm0;
        const uint32_t smem0 = smem_u32(smem);
        const bool dbg = a.dbg && blockIdx.x == 0 && lane == 0;
        auto stamp = [&](int t, int k) { if (dbg && t < 256) a.dbg[t * 8 + k] = clock64(); };
        for (int t = 0; t < ntile; ++t) {
            const int s = t % SA;
            mbar_wait(&r_full[t & 1], (t >> 1) & 1);
            tc_fence_after();
            stamp(t, 2);
            // R of warpgroup w (data rows 16 w .. 16 w + 15 of the tile) sits as 8 packed columns at the start of
            // the warpgroup's 16 logit columns; one K = 16 MMA per warpgroup, 16 rows = 2,048 bytes of the Xhb boxes
            const uint32_t tr = C::COL_Z + (uint32_t)((t & 1) * NT);
            const uint32_t lo_b = dm_lo0 + (((smem0 + (uint32_t)(s * C::A_BYTES + C::OFF_HB)) & 0x3FFFF) >> 4);
            if (t == 0) umma_f16_ts_w<false>(t_g, tr, lo_b, dm_hi, idesc2);
            else umma_f16_ts_w<true>(t_g, tr, lo_b, dm_hi, idesc2);
#pragma unroll
            for (int ks = 1; ks < NT / 16; ++ks)
                umma_f16_ts_w<true>(t_g, tr + ks * PCOLS, lo_b + ks * (2048 >> 4), dm_hi, idesc2);
            umma_commit_elect(&emptyA[s]);           // the slot is free once GEMM2 has read it
            umma_commit_elect(&z_free[t & 1]);
            stamp(t, 3);
        }
        if (ntile > 0) umma_commit_elect(g_full);
    } else if (warp >= 4) {
        // ===== pointwise warpgroups: thread = chain (TMEM lane 32 q + lane) =====
        const int q = warp & 3, wg = (warp - 4) >> 2;
        const int64_t c = (int64_t)cb * CB + q * 32 + lane;
        const bool okc = c < a.K;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        if (wg == 0 && ntile > 0) {
            // Theta of this chain -> TMEM (hi and lo parts)
            const int slot = (a.fixed_slot >= 0) ? a.fixed_slot : (okc ? (a.cur[c] ^ 1) : 0);
            const double* th = a.Th + ((int64_t)slot * a.K + (okc ? c : 0)) * a.dp;
            for (int j0 = 0; j0 < DP32; j0 += 32) {
                uint32_t hi[32], hb[16], lb[16];
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    float h[2], x[2], l[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int j = j0 + e + i;
                        const double v = (okc && j < a.d) ? th[j] : 0.0;
                        h[i] = rn_tf32((float)v);
                        x[i] = (float)v;
                        l[i] = (float)(v - (double)h[i]);
                        hi[e + i] = __float_as_uint(h[i]);
                    }
                    hb[e >> 1] = pack_f16(x[0] * CS_DN, x[1] * CS_DN);
                    lb[e >> 1] = pack_f16(l[0] * CS_UP, l[1] * CS_UP);
                }
                tmem_st_32x32(lane_base + C::COL_TH + j0, hi);
                tmem_st_32x16(lane_base + C::COL_TB + (j0 >> 1), hb);
                tmem_st_32x16(lane_base + C::COL_LB + (j0 >> 1), lb);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(th_ready);
        }
        // Every tile is split between the PWG warpgroups (16 of its 64 data rows each).  The per-element arithmetic is a
        // long dependent chain (EX2, ten FMAs, RCP); with one warpgroup per tile the tile's pointwise LATENCY, not its
        // instruction count, set the pace (the tensor pipe idled waiting for R).  Four warps per scheduler, each with a
        // quarter of the tile, bring R(t) in while GEMM1(t+1) is still running.
        double ll = 0.0;
        float* wrow = HASW ? a.W + (okc ? c : 0) * a.ldw : nullptr;
        const int col0 = wg * PCOLS;
        // The tile's label sign masks (16 words per warpgroup, the same for every chain) come straight from global
        // memory into registers, one tile AHEAD, so the stage starts the moment GEMM1 completes.
        const uint4* ysg = reinterpret_cast<const uint4*>(a.ys + t_begin * NT + col0);
        uint4 ymn[PCOLS / 4];
#pragma unroll
        for (int i = 0; i < PCOLS / 4; ++i) ymn[i] = (ntile > 0) ? __ldg(ysg + i) : make_uint4(0, 0, 0, 0);
        for (int t = 0; t < ntile; ++t) {
            uint32_t ymv[PCOLS];
#pragma unroll
            for (int i = 0; i < PCOLS / 4; ++i) { ymv[4 * i] = ymn[i].x; ymv[4 * i + 1] = ymn[i].y; ymv[4 * i + 2] = ymn[i].z; ymv[4 * i + 3] = ymn[i].w; }
            if (t + 1 < ntile) {
#pragma unroll
                for (int i = 0; i < PCOLS / 4; ++i) ymn[i] = __ldg(ysg + (size_t)(t + 1) * (NT / 4) + i);
            }
            const int64_t row0 = (t_begin + t) * NT + col0;
            mbar_wait(&z_full[t & 1], (t >> 1) & 1);
            tc_fence_after();
            const bool dbgp = a.dbg && blockIdx.x == 0 && warp == 4 && lane == 0 && t < 256;
            if (dbgp) a.dbg[t * 8 + 4] = clock64();
            const uint32_t tz = lane_base + C::COL_Z + (uint32_t)((t & 1) * NT + col0);
            float v[PCOLS];
            tmem_ld_32x16(tz, v);
            if (dbgp) a.dbg[t * 8 + 5] = clock64();
            uint32_t rr[PCOLS / 2];
            float wv[HASW ? PCOLS : 2];
            float2 part = f2(0.0f);
#pragma unroll
            for (int e = 0; e < PCOLS; e += 2) {
                const uint2 ym = make_uint2(ymv[e], ymv[e + 1]);                // 0x80000000 where y = 1
                const float s0 = __uint_as_float(__float_as_uint(v[e]) ^ ym.x);    // s = (1 - 2y) z
                const float s1 = __uint_as_float(__float_as_uint(v[e + 1]) ^ ym.y);
                const float2 ex = make_float2(ex2_approx(-1.4426950408889634f * fabsf(v[e])),     // e = exp(-|z|)
                                              ex2_approx(-1.4426950408889634f * fabsf(v[e + 1])));
                // softplus(s) = max(s, 0) + log1p(e) = softplus(z) - y z;  the log-likelihood term is its negative
                const float2 sp = __ffma2_rn(log1p_unit2(ex), f2(1.0f), make_float2(fmaxf(s0, 0.0f), fmaxf(s1, 0.0f)));
                part = __ffma2_rn(sp, f2(-1.0f), part);
                const float2 q = __ffma2_rn(ex, f2(1.0f), f2(1.0f));
                const float2 inv = make_float2(rcp_approx(q.x), rcp_approx(q.y));
                const float2 ei = __ffma2_rn(ex, inv, f2(0.0f));
                const float g0 = (s0 >= 0.0f) ? inv.x : ei.x;                   // sigmoid(s);  y - p = (2y - 1) sigmoid(s)
                const float g1 = (s1 >= 0.0f) ? inv.y : ei.y;
                rr[e >> 1] = pack_f16(__uint_as_float(__float_as_uint(g0) | (~ym.x & 0x80000000u)),        // nearest fp16
                                      __uint_as_float(__float_as_uint(g1) | (~ym.y & 0x80000000u)));
                if (HASW) { const float2 w2 = __ffma2_rn(ei, inv, f2(0.0f)); wv[e] = w2.x; wv[e + 1] = w2.y; }
                if ((e & 6) == 6) { ll += (double)(part.x + part.y); part = f2(0.0f); }
            }
            if (dbgp) a.dbg[t * 8 + 6] = clock64();
            tmem_st_32x8(tz, rr);
            if (HASW && okc && row0 < a.ldw) {                       // ldw is a multiple of 32, row0 of 16
                if (a.w_bf16) {
                    __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(a.W) + (okc ? c : 0) * a.ldw + row0;
                    uint32_t pk[PCOLS / 2];
#pragma unroll
                    for (int e = 0; e < PCOLS; e += 2) {
                        const __nv_bfloat162 b2 = __floats2bfloat162_rn(wv[e], wv[e + 1]);
                        pk[e / 2] = *reinterpret_cast<const uint32_t*>(&b2);
                    }
#pragma unroll
                    for (int e = 0; e < PCOLS / 2; e += 4)
                        *reinterpret_cast<uint4*>(wb + 2 * e) = make_uint4(pk[e], pk[e + 1], pk[e + 2], pk[e + 3]);
                } else {
#pragma unroll
                    for (int e = 0; e < PCOLS; e += 4)
                        *reinterpret_cast<float4*>(wrow + row0 + e) = make_float4(wv[e], wv[e + 1], wv[e + 2], wv[e + 3]);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&r_full[t & 1]);
            if (dbgp) a.dbg[t * 8 + 7] = clock64();
        }
        // rows beyond N (the last tile's padding) are zero rows of X: z = 0 exactly, and each contributed
        // -softplus(0) in the arithmetic above -- take those terms out again
        if (t_end == a.tiles_total && ntile > 0) {
            const int64_t lo = (a.tiles_total - 1) * NT + col0;
            const int64_t npad = lo + PCOLS - max(a.N, lo);
            if (npad > 0) {
                const float2 sp0 = log1p_unit2(f2(ex2_approx(-0.0f)));
                ll += (double)npad * (double)sp0.x;
            }
        }
        if (okc) a.llp[((int64_t)rs * PWG + wg) * a.K + c] = ll;
        if (wg == 0) {
            // gradient partial of this (chain block, row split)
            float* gout = a.gp + ((int64_t)rs * a.K + (okc ? c : 0)) * DP32;
            if (ntile > 0) {
                mbar_wait(g_full, 0);
                tc_fence_after();
#pragma unroll
                for (int j0 = 0; j0 < DP32; j0 += 32) {
                    float v[32];
                    tmem_ld_32x32(lane_base + C::COL_G + j0, v);
                    if (okc) {                                        // GEMM2 contracted R with x 2^-10
#pragma unroll
                        for (int e = 0; e < 32; e += 4)
                            *reinterpret_cast<float4*>(gout + j0 + e) = make_float4(v[e] * CS_UP, v[e + 1] * CS_UP, v[e + 2] * CS_UP, v[e + 3] * CS_UP);
                    }
                }
            } else if (okc) {
                for (int j = 0; j < DP32; ++j) gout[j] = 0.0f;
            }
        }
    }
    tc_fence_before();
    cta_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS_ALLOC);
    }
    return;
    }
    {
        // ===== GEMM1 issuer (warp 1): the kernel's convergent tail, issuing lane elected once (see the role dispatch) =====
        // Every operand is thread-invariant (the CTA owns all 512 TMEM columns, so its TMEM base is 0 and the addresses are
        // literals; descriptors differ by an add on their low word) and lives in uniform registers.
        const uint32_t idesc1 = umma_idesc_tf32(CB, NT), idesc1b = umma_idesc_f16(CB, NT);
        constexpr uint32_t t_th = C::COL_TH, t_tb = C::COL_TB, t_lb = C::COL_LB;
        const int d8 = (a.d + 7) / 8, d16 = (a.d + 15) / 16;                    // K steps of GEMM1 that hold data
        const uint64_t dk0 = umma_desc_kmajor<128>(0);                           // descriptor with a zero start address
        const uint32_t dk_hi = (uint32_t)(dk0 >> 32), dk_lo0 = (uint32_t)dk0;
        const uint32_t smem0 = smem_u32(smem);
        const bool dbg = a.dbg && blockIdx.x == 0 && lane == 0;
        auto stamp = [&](int t, int k) { if (dbg && t < 256) a.dbg[t * 8 + k] = clock64(); };
        const uint32_t leader = elect_leader();
        if (ntile > 0) {
            if (tmem_base != 0) __trap();            // 512 columns allocated: the base cannot be anything else
            mbar_wait(th_ready, 0);
            tc_fence_after();
        }
        for (int t = 0; t < ntile; ++t) {
            const int s = t % SA;
            mbar_wait(&fullA[s], (t / SA) & 1);
            if (t >= 2) mbar_wait(&z_free[t & 1], ((t >> 1) - 1) & 1);          // GEMM2 of tile t - 2 has read its R
            tc_fence_after();
            stamp(t, 0);
            const uint32_t tz = C::COL_Z + (uint32_t)((t & 1) * NT);
            const uint32_t lo_h = dk_lo0 + (((smem0 + (uint32_t)(s * C::A_BYTES)) & 0x3FFFF) >> 4);
            const uint32_t lo_lb = lo_h + (C::OFF_LB >> 4), lo_hb = lo_h + (C::OFF_HB >> 4);
            // Theta_h Xh^T: TF32, K = 8 = 32 bytes of a row per MMA, four per 32-column box
            umma_ts_l<false, false>(leader, tz, t_th, lo_h, dk_hi, idesc1);
            for (int k = 1; k < d8; ++k)
                umma_ts_l<true, false>(leader, tz, t_th + (uint32_t)(k * 8), lo_h + (uint32_t)((k >> 2) * (BOX_BYTES >> 4) + (k & 3) * 2), dk_hi, idesc1);
            // Theta_hb Xlb^T + Theta_lb Xhb^T: fp16, K = 16 = 32 bytes of a row (8 packed TMEM columns) per MMA
            for (int k = 0; k < d16; ++k) {
                const uint32_t off = (uint32_t)((k >> 2) * (BOX_BYTES >> 4) + (k & 3) * 2);
                umma_ts_l<true, true>(leader, tz, t_tb + (uint32_t)(k * 8), lo_lb + off, dk_hi, idesc1b);
                umma_ts_l<true, true>(leader, tz, t_lb + (uint32_t)(k * 8), lo_hb + off, dk_hi, idesc1b);
            }
            umma_commit_l(leader, &z_full[t & 1]);
            stamp(t, 1);
        }
    }
    tc_fence_before();
    cta_sync_all();
}

// X[N][d] (fp64) -> Xh [N][ldx] fp32 (x rounded to nearest TF32), Xlb = fp16((x - Xh) 2^10), Xhb = fp16(x 2^-10) [N][ldx] (row pitch
// ldx = d rounded up to 8), and the label sign masks
__global__ void __launch_bounds__(256)
lgf_prep_kernel(int64_t N, int d, int ldx, int64_t nys, const double* __restrict__ X, const double* __restrict__ y,
                float* __restrict__ Xh, __half* __restrict__ Xlb, __half* __restrict__ Xhb,
                uint32_t* __restrict__ ys) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx < N * ldx) {
        const int64_t i = idx / ldx;
        const int k = (int)(idx % ldx);
        const double x = (k < d) ? X[i * d + k] : 0.0;
        const float hi = rn_tf32((float)x);
        Xh[idx] = hi;
        Xlb[idx] = __float2half_rn((float)(x - (double)hi) * CS_UP);
        Xhb[idx] = __float2half_rn((float)x * CS_DN);
    }
    if (idx < nys) ys[idx] = (idx < N && y[idx] != 0.0) ? 0x80000000u : 0u;
}

// sum the partials over the row splits into slot 0 of llpart / gpart (fixed order)
__global__ void __launch_bounds__(128)
lgf_reduce_kernel(int64_t K, int ns, int dp32, int dp, const double* __restrict__ llp, const float* __restrict__ gp,
                  double* __restrict__ llpart, double* __restrict__ gpart) {
    const int lane = threadIdx.x & 31;
    const int64_t c = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (c >= K) return;
    double ll = 0.0;
    for (int q = lane; q < PWG * ns; q += 32) ll += llp[(int64_t)q * K + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ll += __shfl_xor_sync(0xffffffffu, ll, o);
    if (lane == 0) llpart[c] = ll;
    for (int j = lane; j < dp; j += 32) {
        double g = 0.0;
        if (j < dp32)
            for (int s = 0; s < ns; ++s) g += (double)gp[((int64_t)s * K + c) * dp32 + j];
        gpart[c * dp + j] = g;
    }
}

}  // namespace

bool supported(int d) { return d >= 1 && d <= 128; }
int dp32_of(int d) { return d <= 64 ? 64 : 128; }

void make_geometry(Geometry* g, int64_t N, int d, int64_t K) {
    g->dp32 = dp32_of(d);
    g->ldx = (d + 7) / 8 * 8;                  // rows start on 32-byte sectors
    g->tiles_total = (N + NT - 1) / NT;
    g->nys = g->tiles_total * NT;
    g->nblk = (int)((K + CB - 1) / CB);
    // Row splits.  One CTA per SM (its shared memory sees to that), so the sweep takes waves x (tiles per CTA + a fixed
    // start / drain cost per CTA: TMEM allocation, Theta into tensor memory, first TMA round trip, G out -- measured
    // ~12 us = 8 tiles, gpurun r2bd).  Pick the split count with the least modelled time; at least 8 tiles per CTA.
    const int sms = 148, ovh = 8;
    const int64_t max_ns = std::max<int64_t>(1, g->tiles_total / 8);
    int ns = 1;
    int64_t best = -1;
    for (int v = 1; v <= (int)std::min<int64_t>(max_ns, 4 * sms); ++v) {
        const int64_t tps = (g->tiles_total + v - 1) / v;
        const int64_t ctas = (int64_t)g->nblk * ((g->tiles_total + tps - 1) / tps);
        const int64_t cost = ((ctas + sms - 1) / sms) * (tps + ovh);
        if (best < 0 || cost < best) { best = cost; ns = v; }
    }
    if (const char* e = getenv("RMN_LGF_NS")) { const int v = atoi(e); if (v >= 1 && v <= max_ns) ns = v; }   // A/B aid
    g->tps = (int)((g->tiles_total + ns - 1) / ns);
    g->ns = (int)((g->tiles_total + g->tps - 1) / g->tps);
}

int prep_x(int64_t N, int d, const Geometry& g, const double* X, const double* y, float* Xh, float* Xl, uint32_t* ys,
           cudaStream_t st) {
    const int64_t n = std::max<int64_t>(N * g.ldx, g.nys);
    // the second buffer (N ldx fp32 words) holds the two fp16 copies one after the other
    __half* Xlb = reinterpret_cast<__half*>(Xl);
    lgf_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(N, d, g.ldx, g.nys, X, y, Xh, Xlb, Xlb + N * g.ldx, ys);
    RMN_KERNEL_CHECK();
    return RMN_OK;
}

int make_maps(Maps* m, const Geometry& g, int64_t N, const float* Xh, const float* Xl) {
    // the maps are ldx columns wide: the boxes of the last column block reach past it and are zero-filled there
    const __half* Xlb = reinterpret_cast<const __half*>(Xl);      // 2-byte elements: the bf16 map type only names the size
    if (int rc = make_tmap_2d(&m->xh, Xh, (uint64_t)N, (uint64_t)g.ldx, (uint64_t)g.ldx, NT)) return rc;
    if (int rc = make_tmap_2d_bf16(&m->xlb, Xlb, (uint64_t)N, (uint64_t)g.ldx, (uint64_t)g.ldx, NT)) return rc;
    return make_tmap_2d_bf16(&m->xhb, Xlb + N * g.ldx, (uint64_t)N, (uint64_t)g.ldx, (uint64_t)g.ldx, NT);
}

int sweep(const Maps& m, const Geometry& g, SweepArgs a, cudaStream_t st) {
    a.nblk = g.nblk; a.tps = g.tps; a.tiles_total = g.tiles_total;
    static bool attr[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr[dev & 63]) {
        RMN_CUDA(cudaFuncSetAttribute(lg_fused_sweep_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::SMEM));
        RMN_CUDA(cudaFuncSetAttribute(lg_fused_sweep_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::SMEM));
        RMN_CUDA(cudaFuncSetAttribute(lg_fused_sweep_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM));
        RMN_CUDA(cudaFuncSetAttribute(lg_fused_sweep_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM));
        attr[dev & 63] = true;
    }
    const unsigned grid = (unsigned)(g.nblk * g.ns);
    if (g.dp32 == 64) {
        if (a.W) lg_fused_sweep_kernel<64, true><<<grid, THREADS, Cfg<64>::SMEM, st>>>(m.xh, m.xlb, m.xhb, a);
        else lg_fused_sweep_kernel<64, false><<<grid, THREADS, Cfg<64>::SMEM, st>>>(m.xh, m.xlb, m.xhb, a);
    } else {
        if (a.W) lg_fused_sweep_kernel<128, true><<<grid, THREADS, Cfg<128>::SMEM, st>>>(m.xh, m.xlb, m.xhb, a);
        else lg_fused_sweep_kernel<128, false><<<grid, THREADS, Cfg<128>::SMEM, st>>>(m.xh, m.xlb, m.xhb, a);
    }
    RMN_KERNEL_CHECK();
    return RMN_OK;
}

int reduce(const Geometry& g, int64_t K, int dp, const double* llp, const float* gp, double* llpart, double* gpart,
           cudaStream_t st) {
    lgf_reduce_kernel<<<(unsigned)((K * 32 + 127) / 128), 128, 0, st>>>(K, g.ns, g.dp32, dp, llp, gp, llpart, gpart);
    RMN_KERNEL_CHECK();
    return RMN_OK;
}

}  // namespace lgf
