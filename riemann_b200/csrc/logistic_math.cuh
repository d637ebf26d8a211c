// riemann_b200 -- fp64 sigmoid / softplus for the logistic likelihood at about half the instruction
// count of exp() + log() + a division from the CUDA math library.
//
// The pointwise stage of lg_eval_kernel shares the fp64 pipe with the DMMA products, and with library
// calls it executed ~95 instructions per (chain, data row).  Both transcendental arguments live in
// narrow ranges -- e = exp(-|z|) in (0, 1], m = 1 + e in (1, 2] -- so 256-entry tables bring the
// polynomial arguments below 2^-9 and short Taylor sums reach 1e-16:
//   exp(-a)  = 2^q * T2[j] * exp(r),        -a = (256 q + j) ln2/256 + r,  |r| <= ln2/512, degree 5
//   log(m)   = k ln2 + logc[j] + log1p(r),   r = m' invc[j] - 1, |r| < 2^-9,  degree 6
//   1/m      = 2^-k invc[j] / (1 + r),       geometric sum to r^6
// with m' = m 2^-k in [1,2), j = top 8 mantissa bits of m', c_j = 1 + (j + 1/2)/256.
// Absolute accuracy ~2e-16 on softplus and sigmoid (tests: device log-posterior vs numpy 1e-9 relative
// over up to 1e6 rows; tests/test_gpu_logistic.py::test_fast_sigmoid_softplus_accuracy sweeps z).
// NaN propagates; |z| beyond 64 is clamped (e < 2e-28 is invisible next to 1).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace lgmath {

constexpr int TAB_DOUBLES = 256 + 512;          // T2[256] | {invc, logc}[256]
constexpr int TAB_BYTES = TAB_DOUBLES * 8;

// host: fill the table (long double arithmetic, rounded once)
inline void fill_tables(double* tab) {
    for (int j = 0; j < 256; ++j) {
        tab[j] = (double)exp2l((long double)j / 256.0L);
        const long double c = 1.0L + ((long double)j + 0.5L) / 256.0L;
        tab[256 + 2 * j] = (double)(1.0L / c);
        tab[256 + 2 * j + 1] = (double)logl(c);
    }
}

// p = sigmoid(z), sp = softplus(z) = log(1 + e^z), pq = p (1 - p).
// NANPROP = false skips the NaN propagation (callers whose prior term already maps a NaN state to -inf).
template <bool NANPROP = true>
__device__ __forceinline__ void sigmoid_softplus(double zz, const double* __restrict__ tab, double& p, double& sp,
                                                 double& pq) {
    const double* T2 = tab;
    const double2* LC = reinterpret_cast<const double2*>(tab + 256);
    const double a = fmin(fabs(zz), 64.0);
    // ---- e = exp(-a)
    double kd = fma(a, -369.3299304675746, 6755399441055744.0);     // rint(-a 256/ln2) in the low mantissa bits
    const int ki = __double2loint(kd);
    kd -= 6755399441055744.0;
    double r = fma(kd, -0x1.62e42fefa0000p-9 /* ln2/256, high 36 bits: kd * hi is exact */, -a);
    r = fma(kd, -6.432011555819173e-15 /* ln2/256 - hi */, r);
    double pe = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    pe = fma(pe, r, 1.0 / 6.0);
    pe = fma(pe, r, 0.5);
    pe = fma(pe, r, 1.0);
    pe = fma(pe, r, 1.0);
    double e = T2[ki & 255] * pe;
    e = __hiloint2double(__double2hiint(e) + ((ki >> 8) << 20), __double2loint(e));
    // ---- m = 1 + e in (1, 2]:  log(m) and 1/m from one table entry
    const double m = 1.0 + e;
    const int hi = __double2hiint(m);
    const int k = (hi >> 20) - 1023;                                 // 0, or 1 when m == 2
    const double2 lc = LC[(hi >> 12) & 255];
    const double mn = __hiloint2double(hi - (k << 20), __double2loint(m));
    const double rr = fma(mn, lc.x, -1.0);
    double pl = fma(rr, -1.0 / 6.0, 0.2);
    pl = fma(pl, rr, -0.25);
    pl = fma(pl, rr, 1.0 / 3.0);
    pl = fma(pl, rr, -0.5);
    pl = fma(pl, rr, 1.0);
    const double lm = fma(rr, pl, fma((double)k, 0.6931471805599453, lc.y));
    double pr = 1.0 - rr;
    pr = fma(-rr, pr, 1.0);
    pr = fma(-rr, pr, 1.0);
    pr = fma(-rr, pr, 1.0);
    pr = fma(-rr, pr, 1.0);
    pr = fma(-rr, pr, 1.0);
    const double inv = (lc.x * pr) * (k ? 0.5 : 1.0);
    const double ei = e * inv;
    p = (zz >= 0.0) ? inv : ei;
    sp = fmax(zz, 0.0) + lm;
    pq = ei * inv;
    if (NANPROP && zz != zz) { p = zz; sp = zz; pq = zz; }
}

}  // namespace lgmath
