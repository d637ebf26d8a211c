// riemann_b200 -- tcgen05 3xTF32 GEMM kernel (see tc_gemm.cuh): C[m][n] (fp32) = A B^T, store-only epilogue.
// The dense-Gaussian sampler multiplies the INCREMENT delta = theta' - theta by P with it (forming P y' directly would
// cancel 30 - 29.97 in fp32 for the 0.1 I + 0.9 11^T target; the increment has no common mode), the logistic sampler
// its Fisher-metric product.  This is the fp32-accurate tensor-core counterpart of gemm_abt_kernel (fp64 DMMA) for
// SURVEY.md row D4 / BASELINE config 3.  (Round 1 also carried an epilogue that fused the MH row reductions; its
// eight epilogue warps stalled on L2 latency and it measured slower than a separate row pass -- removed.)
#include "common.cuh"
#include "tc_gemm.cuh"

namespace tc {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

int make_tmap_2d(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_rows, int tk) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) { rmn_set_error("cuTensorMapEncodeTiled not available from the driver"); return RMN_ERR_CUDA; }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld_elems * sizeof(float)};
    if (tk != 32 && tk != 16) { rmn_set_error("make_tmap_2d: k-block must be 16 or 32"); return RMN_ERR_PARAM; }
    cuuint32_t box[2] = {(cuuint32_t)tk, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     tk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { rmn_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return RMN_ERR_CUDA; }
    return RMN_OK;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) { rmn_set_error("cuTensorMapEncodeTiled not available from the driver"); return RMN_ERR_CUDA; }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld_elems * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { rmn_set_error("cuTensorMapEncodeTiled (bf16) failed (%d)", (int)r); return RMN_ERR_CUDA; }
    return RMN_OK;
}

// Persistent kernel: gridDim.x CTAs (one per SM) walk the tile list  tile = blockIdx.x + i * gridDim.x,
// tile -> (m_tile, n_tile) with n fastest, so concurrently running CTAs share the same rows of A in L2.
// Three independent pipelines (Blackwell canonical form):
//   TMA warp   --full/empty[STAGES]-->   MMA thread   --tmem_full/tmem_empty[ACC_STAGES]-->   8 epilogue warps
// The accumulator is double buffered in TMEM (2 x 256 columns), so the epilogue of tile i runs while the
// tensor cores work on tile i+1.
// PASSES = 3: fp32-accurate split product (Ah Bh + Ah Bl + Al Bh), k-blocks of TK3 = 16 (SWIZZLE_64B): 4 stages of 48 KB.
// PASSES = 1: plain TF32 product of the "hi" maps only (the Fisher-metric GEMM of the logistic sampler,
//             where the product only shapes a proposal), 4 stages of 48 KB so the TMA latency stays hidden
//             behind one third of the tensor work per stage.
// Launch bounds: the kernel runs ONE persistent CTA per SM (its shared memory sees to that); the "3" only caps the
// registers at 56 per thread (no spills) so that memory-bound kernels of another stream -- the dense sampler's row pass of
// the other half-batch -- find 44K free registers on the SM while the tensor cores work (RMN_TC_MINB to change).
#ifndef RMN_TC_MINB
#define RMN_TC_MINB 3
#endif
// BF16 (PASSES = 1 only): the operands are bf16 [rows][64] k-blocks and the instruction is kind::f16 -- twice the
//             contraction depth per 128-byte row and per MMA, for products that only shape a proposal.
template <int PASSES, bool BF16 = false>
__global__ void __launch_bounds__(THREADS, RMN_TC_MINB)
tf32x3_gemm_kernel(const __grid_constant__ GemmMaps maps, int64_t M, int N, int Kdim, float* __restrict__ C,
                   int ldc, int ksplit, int kb_per, int64_t split_stride, int m_fastest, int tn, int bbox,
                   int64_t wide_tiles, long long* dbg) {
    // dbg: optional {first CTA start, last CTA end} in globaltimer ns (set_debug_stamp; the dense sampler's timeline aid)
    if (dbg && threadIdx.x == 0) atomicMin(reinterpret_cast<unsigned long long*>(dbg), (unsigned long long)rmn_globaltimer());
    // Tile list: the first `wide_tiles` tiles are tn columns wide; every further one is a HALF tile (tn / 2 columns, two
    // per remaining full tile).  512 tiles of 128 x 256 on 148 SMs are 3.46 waves = 4 rounds of the persistent loop with
    // the last one 46 % full; with the last 68 tiles cut in two, 136 CTAs work half a round: 3.5 rounds.  bbox = rows of
    // B per TMA box (the B maps' box height): a tile issues (width / bbox) loads per operand.
    // tn: columns of C per tile, TN = 256 or 128 (the B tensor maps must have been built with box_rows = tn).  The
    // narrow tile is for outputs with fewer than #SM tiles of 128 x 256 (2,048 chains x 1,024 columns = 64 of them):
    // it doubles the tile count instead of leaving half of the SMs idle.  Shared-memory regions keep their 256-row size.
    // m_fastest: tile order.  0 = n fastest (CTAs running together share rows of A in L2), 1 = m fastest
    // (they share rows of B: the logits GEMM, whose A -- the chains' Theta -- is tiny and whose B -- the data
    // matrix -- should cross HBM once).
    // ksplit > 1: the k-blocks are divided into ksplit contiguous ranges; a tile is
    // (m_tile, n_tile, split) and split s stores its partial product at C + s * split_stride (summed by
    // the caller) -- this is how a product with few output tiles but a long contraction (the logistic
    // gradient R X: K x d output, N data rows deep) still fills all SMs.
    static_assert(!BF16 || PASSES == 1, "the bf16 operands have no split form");
    constexpr int TK = BF16 ? 64 : ((PASSES == 1) ? tc::TK : tc::TK3);  // elements per k-block
    constexpr int UK = BF16 ? 16 : tc::UK;                              // contraction depth of one MMA
    constexpr int ROWB = TK * (BF16 ? 2 : 4);                           // bytes per tile row = swizzle span
    constexpr int A_BYTES = TM * ROWB, B_BYTES = TN * ROWB;
    constexpr int STAGE_BYTES = (PASSES == 1) ? (A_BYTES + B_BYTES) : 2 * (A_BYTES + B_BYTES);
    constexpr int STAGES = (tc::STAGES * tc::STAGE_BYTES) / STAGE_BYTES;   // 4 x 48 KB (or 2 x 96 KB with TK3 = 32)
    constexpr int OFF_BH = (PASSES == 1) ? A_BYTES : 2 * A_BYTES;
    static_assert(STAGES * STAGE_BYTES <= tc::STAGES * tc::STAGE_BYTES, "stage ring must fit SMEM_BYTES");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint64_t* tmem_empty = tmem_full + ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);
    float* epi_stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);   // [EPI_WARPS][32][20]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB_all = Kdim / TK;
    // kb_per = k-blocks per split (the last one may be short, none is empty: launch_common)
    const int n_tiles = (N + tn - 1) / tn;
    const int64_t m_tiles = (M + TM - 1) / TM;
    const int64_t mn_tiles = m_tiles * n_tiles;
    const int64_t num_tiles = wide_tiles + 2 * (mn_tiles * ksplit - wide_tiles);
    // tile index -> (output tile, column offset inside it, width)
    auto decode = [&](int64_t tile, int64_t& wt, int& noff, int& w) {
        if (tile < wide_tiles) { wt = tile; noff = 0; w = tn; }
        else { const int64_t j = tile - wide_tiles; wt = wide_tiles + (j >> 1); w = tn >> 1; noff = (int)(j & 1) * w; }
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.ah); tma_prefetch_desc(&maps.bh);
        if (PASSES == 3) { tma_prefetch_desc(&maps.al); tma_prefetch_desc(&maps.bl); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer: runs ahead across tiles, bounded by the stage ring =====
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int64_t wt; int noff, w;
            decode(tile, wt, noff, w);
            const int64_t mn = wt % mn_tiles;
            const int kb0 = (int)(wt / mn_tiles) * kb_per;
            const int kb1 = min(KB_all, kb0 + kb_per);
            const int m0 = (int)(m_fastest ? mn % m_tiles : mn / n_tiles) * TM;
            const int n0 = (int)(m_fastest ? mn / m_tiles : mn % n_tiles) * tn + noff;
            const uint32_t tx_bytes = (uint32_t)((PASSES == 1 ? 1 : 2) * (A_BYTES + w * ROWB));
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int s = it % STAGES;
                mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                uint8_t* st = smem + s * STAGE_BYTES;
                mbar_expect_tx(&full[s], tx_bytes);
                tma_load_2d(st, &maps.ah, &full[s], kb * TK, m0);
                if (PASSES == 3) tma_load_2d(st + A_BYTES, &maps.al, &full[s], kb * TK, m0);
                for (int b = 0; b < w; b += bbox) {               // consecutive boxes continue the same swizzled layout
                    tma_load_2d(st + OFF_BH + b * ROWB, &maps.bh, &full[s], kb * TK, n0 + b);
                    if (PASSES == 3) tma_load_2d(st + 2 * A_BYTES + B_BYTES + b * ROWB, &maps.bl, &full[s], kb * TK, n0 + b);
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer (one thread) =====
        // UMMA N = the valid width of B rounded up to 16 (a d-wide gradient tile does not pay for 256 columns)
        uint32_t it = 0, ti = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
            const int a = ti % ACC_STAGES;
            mbar_wait(&tmem_empty[a], ((ti / ACC_STAGES) & 1) ^ 1);      // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(a * TN);
            int64_t wt; int noff, w;
            decode(tile, wt, noff, w);
            const int n_eff = (N >= tn) ? w : ((N + 15) / 16 * 16);
            const uint32_t idesc = BF16 ? umma_idesc_bf16(TM, n_eff) : umma_idesc_tf32(TM, n_eff);
            const int kb0 = (int)(wt / mn_tiles) * kb_per;
            const int kb1 = min(KB_all, kb0 + kb_per);
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int s = it % STAGES;
                mbar_wait(&full[s], (it / STAGES) & 1);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                const uint64_t dah = umma_desc_kmajor<ROWB>(sa);
                const uint64_t dal = umma_desc_kmajor<ROWB>(sa + A_BYTES);
                const uint64_t dbh = umma_desc_kmajor<ROWB>(sa + OFF_BH);
                const uint64_t dbl = umma_desc_kmajor<ROWB>(sa + 2 * A_BYTES + B_BYTES);
#pragma unroll
                for (int k = 0; k < TK / UK; ++k) {
                    const uint64_t adv = (uint64_t)((k * 32) >> 4);       // +32 bytes per K step, 16-byte units
                    if (BF16) umma_bf16(tacc, dah + adv, dbh + adv, idesc, ((kb - kb0) | k) != 0);
                    else umma_tf32(tacc, dah + adv, dbh + adv, idesc, ((kb - kb0) | k) != 0);
                    if (PASSES == 3) {
                        umma_tf32(tacc, dah + adv, dbl + adv, idesc, 1);
                        umma_tf32(tacc, dal + adv, dbh + adv, idesc, 1);
                    }
                }
                umma_commit(&empty[s]);                 // frees the stage when these MMAs retire
            }
            umma_commit(&tmem_full[a]);                 // accumulator complete
        }
    } else if (warp >= 4) {
        // ===== epilogue: warp w reads TMEM lanes 32*(w%4).. (one row per thread), columns half (w-4)/4 =====
        const int q = warp & 3, half = (warp - 4) >> 2;
        uint32_t ti = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
            const int a = ti % ACC_STAGES;
            int64_t wt; int noff, w;
            decode(tile, wt, noff, w);
            const int64_t mn = wt % mn_tiles;
            const int64_t m0 = (m_fastest ? mn % m_tiles : mn / n_tiles) * TM;
            const int n0 = (int)(m_fastest ? mn / m_tiles : mn % n_tiles) * tn + noff;
            float* Cs = C + (wt / mn_tiles) * split_stride;
            mbar_wait(&tmem_full[a], (ti / ACC_STAGES) & 1);
            tc_fence_after();
            // Epilogue data mapping: tcgen05.ld hands each thread one TMEM lane (= output row) x 16 columns;
            // every 32 x 16 chunk is transposed through shared memory so that a warp instruction touches
            // 8 rows x 64 contiguous bytes of C and of the epilogue's input arrays instead of 32 rows x 16 B.
            // Lane L then works on rows it*8 + L/4 (it = 0..3), columns 4 (L%4) .. +3 of the chunk.
            float* stg = epi_stage + (warp - 4) * EPI_STAGE_FLOATS;
            const int rsub = lane >> 2, cg = (lane & 3) * 4;
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * TN + half * (w / 2));
            for (int c0 = 0; c0 < w / 2; c0 += 16) {
                const int n = n0 + half * (w / 2) + c0;
                if (n >= N) continue;                                   // warp-uniform
                bool okr[4];
#pragma unroll
                for (int it = 0; it < 4; ++it) okr[it] = (m0 + q * 32 + it * 8 + rsub < M) && (n + cg < N);
                float v[16];
                tmem_ld_32x16(trow + (uint32_t)c0, v);
                float4* w4 = reinterpret_cast<float4*>(stg + lane * 20);
#pragma unroll
                for (int i = 0; i < 4; ++i) w4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int rr = it * 8 + rsub;
                    const int64_t mr = m0 + q * 32 + rr;
                    if (!okr[it]) continue;
                    const float4 val = *reinterpret_cast<const float4*>(stg + rr * 20 + cg);
                    *reinterpret_cast<float4*>(Cs + mr * ldc + n + cg) = val;
                }
                __syncwarp();
            }
            // this warp is done with accumulator a: hand it back to the MMA thread
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[a]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
    if (dbg && threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(dbg + 1), (unsigned long long)rmn_globaltimer());
}

static thread_local long long* g_dbg_stamp = nullptr;
// the NEXT launch of this thread records {start, end} globaltimer stamps at p[0], p[1] (p[0] preset to LLONG_MAX, p[1] to 0)
void set_debug_stamp(long long* p) { g_dbg_stamp = p; }

// function attributes (dynamic shared memory limit) are per device; set them outside any stream capture
int prepare_kernels() {
    static bool attr_done[64] = {false};
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    bool& attr = attr_done[dev_id & 63];
    if (!attr) {
        RMN_CUDA(cudaFuncSetAttribute(tf32x3_gemm_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        RMN_CUDA(cudaFuncSetAttribute(tf32x3_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        RMN_CUDA(cudaFuncSetAttribute(tf32x3_gemm_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr = true;
    }
    return RMN_OK;
}

static int launch_common(const GemmMaps& maps, int64_t M, int N, int Kdim, float* C, int ldc,
                         cudaStream_t st, int passes = 3, int ksplit = 1, int64_t split_stride = 0, int m_fastest = 0,
                         int tn = TN, bool mixed = false, bool bf16 = false) {
    if (tn != TN && tn != 128) { rmn_set_error("tf32x3 gemm: tile width must be 256 or 128"); return RMN_ERR_PARAM; }
    const int tk = bf16 ? 64 : ((passes == 1) ? TK : TK3);              // k-block of the kernel variant (elements)
    if (bf16 && passes != 1) { rmn_set_error("tf32x3 gemm: bf16 operands are single pass"); return RMN_ERR_PARAM; }
    if (Kdim % tk != 0 || ldc % 4 != 0) { rmn_set_error("tf32x3 gemm: K must be a multiple of the k-block (%d), ld of 4", tk); return RMN_ERR_PARAM; }
    if (ksplit < 1 || ksplit > Kdim / tk) { rmn_set_error("tf32x3 gemm: bad ksplit"); return RMN_ERR_PARAM; }
    // every split must own at least one k-block (an empty range would leave its accumulator unwritten)
    const int kbp = (Kdim / tk + ksplit - 1) / ksplit;
    ksplit = (Kdim / tk + kbp - 1) / kbp;
    const int64_t tiles = (int64_t)((N + tn - 1) / tn) * ((M + TM - 1) / TM) * ksplit;
    static int sms_of[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& sms = sms_of[dev & 63];
    if (!sms) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // mixed: the B maps have 128-row boxes; the tiles of the last, partly filled round become half tiles when those fit
    // one round (2 R <= #SM; fewer than #SM/2 tiles in all: every tile is a half tile = the narrow configuration)
    int64_t wide = tiles;
    int bbox = tn;
    if (mixed) {
        if (tn != TN || N % TN != 0) { rmn_set_error("tf32x3 gemm: mixed tiles need N to be a multiple of 256"); return RMN_ERR_PARAM; }
        bbox = 128;
        const int64_t rem = tiles % sms;
        if (rem > 0 && 2 * rem <= sms) wide = tiles - rem;
    }
    const int64_t launched = wide + 2 * (tiles - wide);
    dim3 grid((unsigned)(launched < sms ? launched : sms));    // persistent: one CTA per SM
    if (int rc = prepare_kernels()) return rc;
    long long* dbg = g_dbg_stamp;
    g_dbg_stamp = nullptr;
    if (bf16) tf32x3_gemm_kernel<1, true><<<grid, THREADS, SMEM_BYTES, st>>>(maps, M, N, Kdim, C, ldc, ksplit, kbp, split_stride, m_fastest, tn, bbox, wide, dbg);
    else if (passes == 1) tf32x3_gemm_kernel<1><<<grid, THREADS, SMEM_BYTES, st>>>(maps, M, N, Kdim, C, ldc, ksplit, kbp, split_stride, m_fastest, tn, bbox, wide, dbg);
    else tf32x3_gemm_kernel<3><<<grid, THREADS, SMEM_BYTES, st>>>(maps, M, N, Kdim, C, ldc, ksplit, kbp, split_stride, m_fastest, tn, bbox, wide, dbg);
    RMN_KERNEL_CHECK();
    return RMN_OK;
}

int launch_plain(const GemmMaps& maps, int64_t M, int N, int Kdim, float* C, int ldc, cudaStream_t st) {
    return launch_common(maps, M, N, Kdim, C, ldc, st);
}
// 3-pass product with 128-column tiles (B maps built with box_rows = 128)
int launch_plain_narrow(const GemmMaps& maps, int64_t M, int N, int Kdim, float* C, int ldc, cudaStream_t st) {
    return launch_common(maps, M, N, Kdim, C, ldc, st, 3, 1, 0, 0, 128);
}
// 3-pass product, 128 x 256 tiles with the last partly filled round cut into 128 x 128 tiles (B maps: box_rows = 128)
int launch_plain_mixed(const GemmMaps& maps, int64_t M, int N, int Kdim, float* C, int ldc, cudaStream_t st) {
    return launch_common(maps, M, N, Kdim, C, ldc, st, 3, 1, 0, 0, TN, true);
}
int launch_plain_tf32(const GemmMaps& maps, int64_t M, int N, int Kdim, float* C, int ldc, cudaStream_t st) {
    return launch_common(maps, M, N, Kdim, C, ldc, st, 1);
}
// single pass with bf16 operands (maps.ah / maps.bh from make_tmap_2d_bf16), fp32 accumulate and output
int launch_plain_bf16(const GemmMaps& maps, int64_t M, int N, int Kdim, float* C, int ldc, cudaStream_t st) {
    return launch_common(maps, M, N, Kdim, C, ldc, st, 1, 1, 0, 0, TN, false, true);
}
// Single-pass products (TF32, or bf16 when `bf16`) with the contraction split into `ksplit` ranges, partial s at
// C + s * split_stride: for outputs with far fewer tiles than SMs and a long contraction (the Fisher-metric GEMM of a
// small chain shard).  splits_used() tells the caller how many partials the launch writes.
int splits_used(int Kdim, int ksplit, bool bf16) {
    const int tk = bf16 ? 64 : TK;
    const int kb = Kdim / tk;
    if (ksplit < 1) ksplit = 1;
    if (ksplit > kb) ksplit = kb > 0 ? kb : 1;
    const int kbp = (kb + ksplit - 1) / ksplit;
    return kbp > 0 ? (kb + kbp - 1) / kbp : 1;
}
int launch_single_splitk(const GemmMaps& maps, int64_t M, int N, int Kdim, float* C, int ldc, int ksplit,
                         int64_t split_stride, bool bf16, cudaStream_t st) {
    return launch_common(maps, M, N, Kdim, C, ldc, st, 1, ksplit, split_stride, 0, TN, false, bf16);
}
// 3-pass product with the contraction split into `ksplit` ranges; returns the number of splits actually
// used through *used (<= ksplit); partial s is at C + s * split_stride
int launch_plain_splitk(const GemmMaps& maps, int64_t M, int N, int Kdim, float* C, int ldc, int ksplit,
                        int64_t split_stride, int* used, cudaStream_t st, int passes) {
    const int tk = (passes == 1) ? TK : TK3;
    const int kbp = (Kdim / tk + ksplit - 1) / ksplit;
    if (used) *used = (Kdim / tk + kbp - 1) / kbp;
    return launch_common(maps, M, N, Kdim, C, ldc, st, passes, ksplit, split_stride);
}
}  // namespace tc

// Validation entry of the single-pass mode: C[M][N] (fp32, ld = N) ~= A B^T with TF32 operands
// (tcgen05.mma.kind::tf32 reads fp32 and drops the low 13 mantissa bits); A is [M][K], B is [N][K].
extern "C" int rmn_tf32_gemm(int64_t M, int N, int Kdim, const float* d_A, const float* d_B, float* d_C, void* stream) {
    RMN_REQUIRE(M >= 1 && N >= 1 && Kdim >= 32 && Kdim % 32 == 0 && N % 4 == 0, "rmn_tf32_gemm: bad shape");
    RMN_REQUIRE(d_A && d_B && d_C, "rmn_tf32_gemm: null pointer");
    tc::GemmMaps maps;
    int rc;
    if ((rc = tc::make_tmap_2d(&maps.ah, d_A, M, Kdim, Kdim, tc::TM))) return rc;
    if ((rc = tc::make_tmap_2d(&maps.bh, d_B, N, Kdim, Kdim, tc::TN))) return rc;
    maps.al = maps.ah; maps.bl = maps.bh;
    return tc::launch_plain_tf32(maps, M, N, Kdim, d_C, N, (cudaStream_t)stream);
}

// Validation entry of the bf16 mode: C[M][N] (fp32, ld = N) = A B^T with bf16 operands A [M][K], B [N][K]; K % 64 == 0.
extern "C" int rmn_bf16_gemm(int64_t M, int N, int Kdim, const void* d_A, const void* d_B, float* d_C, void* stream) {
    RMN_REQUIRE(M >= 1 && N >= 1 && Kdim >= 64 && Kdim % 64 == 0 && N % 4 == 0, "rmn_bf16_gemm: bad shape");
    RMN_REQUIRE(d_A && d_B && d_C, "rmn_bf16_gemm: null pointer");
    tc::GemmMaps maps;
    int rc;
    if ((rc = tc::make_tmap_2d_bf16(&maps.ah, d_A, M, Kdim, Kdim, tc::TM))) return rc;
    if ((rc = tc::make_tmap_2d_bf16(&maps.bh, d_B, N, Kdim, Kdim, tc::TN))) return rc;
    maps.al = maps.ah; maps.bl = maps.bh;
    return tc::launch_plain_bf16(maps, M, N, Kdim, d_C, N, (cudaStream_t)stream);
}

// Validation entry of the split-K mode: C[ksplit][M][N] partial products (the caller sums them).
extern "C" int rmn_tf32x3_gemm_splitk(int64_t M, int N, int Kdim, int ksplit, const float* d_Ah, const float* d_Al,
                                      const float* d_Bh, const float* d_Bl, float* d_C, int* used_splits, void* stream) {
    RMN_REQUIRE(M >= 1 && N >= 1 && Kdim >= 32 && Kdim % 32 == 0 && N % 4 == 0 && ksplit >= 1, "rmn_tf32x3_gemm_splitk: bad shape");
    RMN_REQUIRE(d_Ah && d_Al && d_Bh && d_Bl && d_C, "rmn_tf32x3_gemm_splitk: null pointer");
    tc::GemmMaps maps;
    int rc;
    if ((rc = tc::make_tmap_2d(&maps.ah, d_Ah, M, Kdim, Kdim, tc::TM, tc::TK3))) return rc;
    if ((rc = tc::make_tmap_2d(&maps.al, d_Al, M, Kdim, Kdim, tc::TM, tc::TK3))) return rc;
    if ((rc = tc::make_tmap_2d(&maps.bh, d_Bh, N, Kdim, Kdim, tc::TN, tc::TK3))) return rc;
    if ((rc = tc::make_tmap_2d(&maps.bl, d_Bl, N, Kdim, Kdim, tc::TN, tc::TK3))) return rc;
    if (ksplit > Kdim / 32) ksplit = Kdim / 32;
    return tc::launch_plain_splitk(maps, M, N, Kdim, d_C, N, ksplit, (int64_t)M * N, used_splits, (cudaStream_t)stream, 3);
}

// Validation entry: C[M][N] (fp32, ld = N) ~= (Ah + Al)(Bh + Bl)^T; A* are [M][K], B* are [N][K], fp32, K % 16 == 0.
// N a multiple of 256 goes through the mixed tile list (the dense sampler's path), anything else through plain tiles.
extern "C" int rmn_tf32x3_gemm(int64_t M, int N, int Kdim, const float* d_Ah, const float* d_Al, const float* d_Bh,
                               const float* d_Bl, float* d_C, void* stream) {
    RMN_REQUIRE(M >= 1 && N >= 1 && Kdim >= 16 && Kdim % tc::TK3 == 0 && N % 4 == 0, "rmn_tf32x3_gemm: bad shape");
    RMN_REQUIRE(d_Ah && d_Al && d_Bh && d_Bl && d_C, "rmn_tf32x3_gemm: null pointer");
    tc::GemmMaps maps;
    int rc;
    if ((rc = tc::make_tmap_2d(&maps.ah, d_Ah, M, Kdim, Kdim, tc::TM, tc::TK3))) return rc;
    if ((rc = tc::make_tmap_2d(&maps.al, d_Al, M, Kdim, Kdim, tc::TM, tc::TK3))) return rc;
    const bool mixed = N % tc::TN == 0;
    if ((rc = tc::make_tmap_2d(&maps.bh, d_Bh, N, Kdim, Kdim, mixed ? 128 : tc::TN, tc::TK3))) return rc;
    if ((rc = tc::make_tmap_2d(&maps.bl, d_Bl, N, Kdim, Kdim, mixed ? 128 : tc::TN, tc::TK3))) return rc;
    if (mixed) return tc::launch_plain_mixed(maps, M, N, Kdim, d_C, N, (cudaStream_t)stream);
    return tc::launch_plain(maps, M, N, Kdim, d_C, N, (cudaStream_t)stream);
}
