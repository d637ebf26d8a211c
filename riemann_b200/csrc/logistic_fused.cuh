// riemann_b200 -- interface of the fused tcgen05 likelihood sweep (logistic_fused.cu), used by logistic.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lgf {

struct Geometry {
    int dp32;              // columns of the operand tiles: 64 (d <= 64) or 128 (d <= 128)
    int ldx;               // row pitch of the X copies in elements: d rounded up to 8 (fp32 rows start on 32-byte sectors, bf16 rows on 16 bytes);
                           // the columns from ldx to dp32 are never stored nor read -- TMA zero-fills them in the tile
    int nblk;              // chain blocks of 128
    int ns;                // row splits (grid = nblk x ns)
    int tps;               // 64-row tiles per split
    int64_t tiles_total;
    int64_t nys;           // entries of the label-mask array (a whole number of tiles)
};

constexpr int LLP_PER_SPLIT = 4;

struct Maps { CUtensorMap xh, xlb, xhb; };   // fp32 hi part; bf16 remainder; bf16 copy (GEMM1 correction K-major, GEMM2 MN-major)

struct SweepArgs {
    const uint32_t* ys;    // [nys] 0x80000000 where y = 1
    const double* Th;      // [2][K][dp] chain states (two slots)
    const int* cur;        // [K] current slot; the sweep evaluates slot cur ^ 1 unless fixed_slot >= 0
    int fixed_slot;
    int64_t K, N;
    int d, dp;
    double* llp;           // [LLP_PER_SPLIT ns][K]  log-likelihood partial sums (one per pointwise warpgroup and split)
    float* gp;             // [ns][K][dp32]  gradient partial sums
    float* W;              // [K][ldw] p (1 - p), or NULL
    int64_t ldw;           // (elements; a multiple of 64)
    int w_bf16;            // W points to bf16 storage (round to nearest): the operand of the bf16 metric GEMM
    long long* dbg;        // optional timeline of CTA 0 (clock64 stamps, RMN_LGF_TIMELINE=1; scripts/lgf_timeline.py), else NULL
    int nblk, tps;         // filled by sweep()
    int64_t tiles_total;
};

bool supported(int d);
int dp32_of(int d);
void make_geometry(Geometry* g, int64_t N, int d, int64_t K);
int prep_x(int64_t N, int d, const Geometry& g, const double* X, const double* y, float* Xh, float* Xl, uint32_t* ys,
           cudaStream_t st);
int make_maps(Maps* m, const Geometry& g, int64_t N, const float* Xh, const float* Xl);
int sweep(const Maps& m, const Geometry& g, SweepArgs a, cudaStream_t st);
int reduce(const Geometry& g, int64_t K, int dp, const double* llp, const float* gp, double* llpart, double* gpart,
           cudaStream_t st);

}  // namespace lgf
