// riemann_b200 -- dense Gaussian target, fp32-accurate tensor-core mode ("tf32x3").
//
// Same MH step as dense.cu (MALA = VanillaHMC(eps, 1, grad), hamiltonian.py:76-91; random walk,
// randomwalk.py:21-26; MultiGaussianDist log-density and gradient, gaussian.py:49-58; accept and
// adapt, sampler.py:72-90 / adaptive.py:26-35) but the product V' = Y' P runs on the 5th-gen
// tensor cores: tcgen05.mma.kind::tf32 fed by TMA, accumulators in TMEM, three TF32 MMAs per
// operand pair (tc_gemm.cu).  Chain state is fp32; everything that enters the accept test
// (quadratic form, |p'|^2, log-posterior) is reduced and kept in fp64.
//
// DELTA form.  Tensor cores accumulate in fp32; forming P y' directly cancels 30 - 29.97 per row
// for the 0.1 I + 0.9 11^T target once the chain carries its O(1) common mode, and the error adds
// coherently over rows (0.4 in the log-posterior at d = 1000 -- measured).  MH proposals are local,
// so the kernels multiply the INCREMENT delta = theta' - theta (no common mode):
//     V' = V + P delta,      quad' - quad = delta . (2 V + P delta),      logpost' = logpost - (quad' - quad)/2
// with the dot products in fp64.  V and the log-posterior of the start state (and of every
// `refresh`-th step, to stop round-off drift) come from an exact fp64 pass.
//
// Accuracy budget (stated, and checked in tests/test_gpu_dense_tf32.py): the log-posterior
// DIFFERENCE that enters the accept test is within 2e-4 of the fp64 value at d = 1000 (fp32
// state, fp32 P); decisions agree with the fp64 reference except when log u is within that
// distance of the threshold; the carried log-posterior stays within 1e-2 of a fresh fp64
// evaluation between refreshes.
//
// TMA boxes cannot follow a per-row "current slot" bit, so this mode keeps ONE increment buffer
// and fuses the accept-update (y += delta, V += P delta) into the finish/propose pass.
#include <climits>
#include <string>
#include <vector>
#include "common.cuh"
#include "tc_gemm.cuh"
#include <stdlib.h>

namespace tc {
int launch_plain(const GemmMaps& maps, int64_t M, int N, int Kdim, float* C, int ldc, cudaStream_t st);
int launch_plain_narrow(const GemmMaps& maps, int64_t M, int N, int Kdim, float* C, int ldc, cudaStream_t st);
int launch_plain_mixed(const GemmMaps& maps, int64_t M, int N, int Kdim, float* C, int ldc, cudaStream_t st);
void set_debug_stamp(long long* p);
int prepare_kernels();
}

size_t dense_exact_scratch_bytes(int64_t K, int d);
int dense_exact_prepare(int64_t K, int d, const double* d_prec, void* scratch);
int dense_exact_pass(int64_t K, int d, int dpf, const float* Yf, float* Vf, double* lp, double c1, double c2, void* scratch,
                     cudaStream_t stream);

namespace {

constexpr int ND_MAX = 8;

struct TState {
    int64_t K; int d, dp;
    float* Y; float* V;            // current state (centred) and V = Y P           [K][dp]
    float* Yph; float* Ypl;        // increment delta: raw fp32, and its remainder delta - trunc_tf32(delta)  [K][dp]
    float* Vp;                     // P delta (GEMM output)                          [K][dp]
    const double* prec;            // fp64 P [d][d] (exact start / refresh pass)
    double* lp; double* k0; double* epsrow;
    double* scale; long long* nsamp; long long* nacc; long long* dacc;
    double* S1; double* S2;
    const double* mu;              // [dp]
    double mubar;                  // mean(mu): the tracked functionals are those of theta, not of the centred state
    const float* Ldiag;            // [dp]
};

struct TStep {
    int prop_kind, adapt, finish, propose, diag;
    double target, eps0, c1, c2;
    uint64_t seed; int64_t chain_offset, step_fin, step_prop;
    const int64_t* d_step_base;   // CUDA-graph replays: step_fin / step_prop are relative to *d_step_base (else NULL)
    int64_t row0, nrows;          // the chain rows [row0, row0 + nrows) this launch works on (nrows = 0: all K)
    long long* dbg;               // optional {first block start, last block end} globaltimer stamps (RMN_TF32_TIMELINE)
    const double* inj_xi; const double* inj_u;
    int64_t trace_slot;
    double* tr_theta; double* tr_logpost; double* tr_prop_lp; uint8_t* tr_acc; double* tr_lqr; double* tr_prop_theta;
};

// streaming loads of the row pass: L2 only (ld.global.cg).  Next to the GEMM's CTA (218 KB of shared memory) the SM keeps
// ~28 KB of L1, less than the pass's loads in flight; lines allocated there only throttle the load pipe.
#ifndef RMN_TF32_FP_LDCG
#define RMN_TF32_FP_LDCG 1
#endif
__device__ __forceinline__ float4 ldg4(const float* p) {
#if RMN_TF32_FP_LDCG
    return __ldcg(reinterpret_cast<const float4*>(p));
#else
    return *reinterpret_cast<const float4*>(p);
#endif
}

// One warp per chain row.  128-thread blocks inside a 48-register budget: next to the GEMM's persistent CTA (384 threads x
// 56 registers, all of the shared memory) an SM still takes 7 of these blocks, which is what lets the pass of one
// half-batch run under the GEMM of the other (DenseTF32Sampler below).
constexpr int FP_THREADS = 128;
#ifndef RMN_TF32_FP_MINB
#define RMN_TF32_FP_MINB 8
#endif
__global__ void __launch_bounds__(FP_THREADS, RMN_TF32_FP_MINB)
finish_propose_f32_kernel(TState st, TStep sp) {
    const int lane = threadIdx.x & 31;
    const int64_t rl = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (sp.dbg && threadIdx.x == 0) atomicMin(reinterpret_cast<unsigned long long*>(sp.dbg), (unsigned long long)rmn_globaltimer());
    if (rl >= (sp.nrows ? sp.nrows : st.K)) return;
    const int64_t r = sp.row0 + rl;
    const int dp = st.dp, d = st.d;
    const int64_t K = st.K;
    double lp = st.lp[r];
    bool acc = false;
    const RngKey rk(sp.seed, (uint64_t)(sp.chain_offset + r));
    const int64_t sbase = sp.d_step_base ? *sp.d_step_base : 0;
    const int64_t step_fin = sp.step_fin + sbase, step_prop = sp.step_prop + sbase;

    if (sp.finish) {
        // quad' - quad = delta . (2 V + P delta) and |p'|^2 from the row itself, in a fixed lane-strided order
        // (deterministic).  p' = p_half - eps/2 (v + P delta) with p_half = delta / eps (hamiltonian.py:27-40; the
        // increment actually applied, theta' = fl32(y + delta)), so the noise is not stored.  The row is re-read from L2
        // by the update loop below; the GEMM keeps its plain, store-only epilogue.
        double q = 0.0, k1 = 0.0;
        {
            const size_t ro0 = (size_t)r * dp;
            const double eps_old = st.epsrow[r];
            const double he = 0.5 * eps_old;
            const bool mala = sp.prop_kind == RMN_PROP_HMC;
            const double ie = mala ? 1.0 / eps_old : 0.0;
            // software pipeline: the loads of iteration i + 1 are issued before the arithmetic of iteration i
            const float* __restrict__ gD = st.Yph + ro0;
            const float* __restrict__ gV = st.V + ro0;
            const float* __restrict__ gP = st.Vp + ro0;
            float4 na = make_float4(0.f, 0.f, 0.f, 0.f), nw = na, npd = na;
            if (lane * 4 < dp) { na = ldg4(gD + lane * 4); nw = ldg4(gV + lane * 4); npd = ldg4(gP + lane * 4); }
            for (int j4 = lane * 4; j4 < dp; j4 += 128) {
                const float4 a = na, w = nw, pd = npd;                                      // delta (raw fp32), V, P delta
                if (j4 + 128 < dp) { na = ldg4(gD + j4 + 128); nw = ldg4(gV + j4 + 128); npd = ldg4(gP + j4 + 128); }
                const float dl[4] = {a.x, a.y, a.z, a.w};
                const float wv[4] = {w.x, w.y, w.z, w.w};
                const float pv[4] = {pd.x, pd.y, pd.z, pd.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const double w2 = (double)wv[e] + (double)pv[e];                        // V of the proposal
                    q += (double)dl[e] * ((double)wv[e] + w2);
                    if (mala) {
                        const double p1 = (double)dl[e] * ie - he * w2;                     // hamiltonian.py:40
                        k1 += p1 * p1;
                    }
                }
            }
        }
        q = group_sum<32>(q);
        k1 = group_sum<32>(k1);
        const double lpn = combine_logpost(0.0, lp - 0.5 * q);      // q = quad' - quad (gaussian.py:52)
        const double lqr = (sp.prop_kind == RMN_PROP_HMC) ? 0.5 * (k1 - st.k0[r]) : 0.0; // hamiltonian.py:89
        const double u = sp.inj_u ? sp.inj_u[r] : u01(rk.block((uint64_t)step_fin, RMN_BLOCK_ACCEPT).x);
        {   // mh_accept (common.cuh) without the fp64 log where the ratio is >= 1: log(u) < 0 for every u in [0, 1)
            const double dl = lpn - lp - lqr;
            acc = (dl < 0.0) ? (log(u) < dl) : (u < 1.0);
        }
        if (acc) lp = lpn;
        if (lane == 0) {
            if (acc) st.lp[r] = lp;
            st.dacc[r] += acc ? 1 : 0;
            if (sp.adapt) {
                AdaptState ad{st.scale[r], st.nsamp[r], st.nacc[r]};
                ad.update(acc, sp.target);
                st.scale[r] = ad.scale; st.nsamp[r] = ad.nsamples; st.nacc[r] = ad.naccepts;
            }
            if (sp.tr_prop_lp) sp.tr_prop_lp[r] = lpn;
            if (sp.tr_acc) sp.tr_acc[r] = acc ? 1 : 0;
            if (sp.tr_lqr) sp.tr_lqr[r] = lqr;
            if (sp.trace_slot >= 0 && sp.tr_logpost) sp.tr_logpost[sp.trace_slot * K + r] = lp;
        }
    }
    __syncwarp();

    const size_t ro = (size_t)r * dp;
    const double scale = sp.adapt ? st.scale[r] : 1.0;
    const double eps = (sp.prop_kind == RMN_PROP_HMC) ? scale * sp.eps0 : scale;
    double k0 = 0.0, rowsum = 0.0;
    const float eps_f = (float)eps, heps_f = (float)(0.5 * eps), scale_f = (float)scale;
    const bool want_trace = sp.finish && sp.trace_slot >= 0 && sp.tr_theta;

    // Same software pipeline: Y, V (and delta, P delta when the row moves or the proposal is traced) of iteration i + 1 are
    // requested before iteration i generates its noise; the stores of iteration i go to other addresses.
    const bool need_a = sp.finish && (acc || sp.tr_prop_theta);
    float4 nyc = make_float4(0.f, 0.f, 0.f, 0.f), nvc = nyc, nda = nyc, npd2 = nyc;
    if (lane * 4 < dp) {
        nyc = ldg4(st.Y + ro + lane * 4); nvc = ldg4(st.V + ro + lane * 4);
        if (need_a) { nda = ldg4(st.Yph + ro + lane * 4); if (acc) npd2 = ldg4(st.Vp + ro + lane * 4); }
    }
    for (int j4 = lane * 4; j4 < dp; j4 += 128) {
        // the row's (new) current state: the accepted proposal or the old state
        float4 yc = nyc, vc = nvc;
        const float4 a = nda, pd = npd2;
        if (j4 + 128 < dp) {
            nyc = ldg4(st.Y + ro + j4 + 128); nvc = ldg4(st.V + ro + j4 + 128);
            if (need_a) { nda = ldg4(st.Yph + ro + j4 + 128); if (acc) npd2 = ldg4(st.Vp + ro + j4 + 128); }
        }
        if (need_a) {
            // the proposal as a state: theta' = fl32(y + delta)
            const float4 yp = make_float4(yc.x + a.x, yc.y + a.y, yc.z + a.z, yc.w + a.w);
            if (sp.tr_prop_theta) {
                const float pv[4] = {yp.x, yp.y, yp.z, yp.w};
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (j4 + q < d) sp.tr_prop_theta[r * d + j4 + q] = (double)pv[q] + st.mu[j4 + q];
            }
            if (acc) {                                           // accept: y += delta, V += P delta
                yc = yp;
                vc = make_float4(vc.x + pd.x, vc.y + pd.y, vc.z + pd.z, vc.w + pd.w);
                *reinterpret_cast<float4*>(st.Y + ro + j4) = yc;
                *reinterpret_cast<float4*>(st.V + ro + j4) = vc;
            }
        }
        const float yv[4] = {yc.x, yc.y, yc.z, yc.w};
        const float vv[4] = {vc.x, vc.y, vc.z, vc.w};
        rowsum += (double)((yv[0] + yv[1]) + (yv[2] + yv[3]));
        if (want_trace) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (j4 + q < d) sp.tr_theta[(sp.trace_slot * K + r) * d + j4 + q] = (double)yv[q] + st.mu[j4 + q];
        }
        if (!sp.propose) continue;
        // The proposal in fp32: the noise is drawn in fp32 (Box-Muller on 32-bit uniforms), the increment is rounded to
        // the grid of y anyway (theta' = fl32(y + delta)), and the acceptance works from the increment actually applied
        // (p_half = delta / eps above), so fp64 here bought nothing but 6 conversions and 4 fp64 operations per element.
        float xf[4];
        if (sp.inj_xi) {
#pragma unroll
            for (int q = 0; q < 4; ++q) xf[q] = (j4 + q < d) ? (float)sp.inj_xi[r * d + j4 + q] : 0.0f;
        } else {
            const uint4 rb = rk.block((uint64_t)step_prop, (uint32_t)(j4 >> 2));
            box_muller(rb.x, rb.y, xf[0], xf[1]);
            box_muller(rb.z, rb.w, xf[2], xf[3]);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (j4 + q >= d) xf[q] = 0.0f;
        }
        float od[4], ol[4];
        float ksum = 0.0f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float dlt;                                                               // theta' - theta
            if (sp.prop_kind == RMN_PROP_HMC) dlt = eps_f * fmaf(-heps_f, vv[q], xf[q]);     // hamiltonian.py:27,30
            else dlt = scale_f * (st.Ldiag[j4 + q] * xf[q]);                                 // randomwalk.py:26
            // the increment actually applied is the fp32 one: theta' = fl32(y + delta)
            od[q] = (yv[q] + dlt) - yv[q];
            ksum = fmaf(xf[q], xf[q], ksum);
            float hi;
            tc::split_tf32(od[q], hi, ol[q]);            // remainder after the tensor core's own truncation of delta
        }
        k0 += (double)ksum;
        // the increment is stored ONCE, raw: kind::tf32 drops the low 13 mantissa bits of its operand itself, so the
        // GEMM's "hi" pass reads this array as is and its "lo" pass the remainder
        *reinterpret_cast<float4*>(st.Yph + ro + j4) = make_float4(od[0], od[1], od[2], od[3]);
        *reinterpret_cast<float4*>(st.Ypl + ro + j4) = make_float4(ol[0], ol[1], ol[2], ol[3]);
    }
    if (sp.propose) {
        k0 = group_sum<32>(k0);
        if (lane == 0) { st.k0[r] = k0; st.epsrow[r] = eps; }
    }
    if (sp.diag) {
        rowsum = group_sum<32>(rowsum);
        const int nd = min(d, ND_MAX - 1) + 1;
        if (lane < nd) {
            const double f = (lane == nd - 1) ? rowsum / (double)d + st.mubar : (double)st.Y[ro + lane] + st.mu[lane];
            st.S1[(int64_t)lane * K + r] += f;
            st.S2[(int64_t)lane * K + r] += f * f;
        }
    }
    if (sp.dbg && lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(sp.dbg + 1), (unsigned long long)rmn_globaltimer());
}

__global__ void tstep_base_kernel(int64_t* p, int64_t v) { *p = v; }

__global__ void tset_kernel(TState st, const double* __restrict__ theta) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= st.K * st.dp) return;
    const int64_t r = i / st.dp;
    const int j = (int)(i % st.dp);
    st.Y[i] = (j < st.d) ? (float)(theta[r * st.d + j] - st.mu[j]) : 0.0f;
    st.Yph[i] = 0.0f; st.Ypl[i] = 0.0f; st.Vp[i] = 0.0f; st.V[i] = 0.0f;
    if (j == 0) { st.k0[r] = 0.0; st.epsrow[r] = 0.0; }
}
__global__ void tget_kernel(TState st, double* theta, double* lp) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= st.K * st.d) return;
    const int64_t r = i / st.d;
    const int j = (int)(i % st.d);
    if (theta) theta[i] = (double)st.Y[(size_t)r * st.dp + j] + st.mu[j];
    if (lp && j == 0) lp[r] = st.lp[r];
}
__global__ void tget_adapt_kernel(TState st, double* scale, int64_t* ns, int64_t* na) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= st.K) return;
    if (scale) scale[c] = st.scale[c];
    if (ns) ns[c] = st.nsamp[c];
    if (na) na[c] = st.nacc[c];
}

struct DenseTF32Sampler : SamplerImpl {
    rmn_sampler* s;
    TState st{};
    tc::GemmMaps maps;
    tc::GemmMaps maps_narrow;     // the same operands with 128-row boxes of P: 128 x 128 output tiles
    bool narrow = false;          // fewer than #SM tiles of 128 x 256: use the narrow tile (RMN_TF32_NARROW=0|1 overrides)
    bool mixed = true;            // 128 x 256 tiles, the last partly filled round of the persistent loop cut into 128 x 128
                                  // halves (tc_gemm.cu); RMN_TF32_MIXED=0: one tile width per launch as in round 1
    float* d_Ph = nullptr; float* d_Pl = nullptr; float* d_Ldiag = nullptr; double* d_mupad = nullptr;
    // The MH loop is two short kernels per step (GEMM + finish/propose pass); for a plain Philox run without trace the
    // T steps of a call are captured ONCE into a CUDA graph and replayed (the step counter the Philox streams need comes
    // from device memory), which removes the per-launch gaps: 2,048 chains per GPU -- BASELINE's 16,384 over 8 -- is
    // 57 us per step with ~16 us of them between kernels.  RMN_TF32_GRAPH=0 turns it off.
    // Inside the graph the chains are processed as TWO half-batches on two captured branches.  The GEMM is tensor-bound
    // (93 registers x 384 threads, one persistent CTA per SM) and the finish/propose pass memory-bound (64 registers):
    // a block of the pass fits next to a GEMM CTA on every SM, so the GEMM of one half overlaps the pass of the other
    // instead of the two alternating on an otherwise idle resource: +5 % at 16,384 chains (237 us per step against 250,
    // gpurun r2j); below 8,192 chains it is neutral or worse and one branch is used.  Results do not change (rows are
    // independent and the Philox streams are keyed by the global chain id).
    tc::GemmMaps maps_h[2];
    bool narrow_h[2] = {false, false};
    int64_t row0_h[2] = {0, 0}, nrows_h[2] = {0, 0};
    cudaStream_t cap_stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool two_halves = true;       // RMN_TF32_HALVES=0: one branch over all rows
    bool use_graph = true;
    cudaGraphExec_t gexec = nullptr;
    int64_t g_T = -1;
    int64_t g_nodes = 0;
    int64_t* d_step_base = nullptr;
    cudaStream_t cap_stream = nullptr;
    // RMN_TF32_TIMELINE=<file>: every kernel of a captured graph stamps its start / end (globaltimer); after each replay the
    // list is appended to the file (scripts/dense_timeline.py reads it).  Measurement aid, off by default.
    std::string tl_path;
    long long* d_tl = nullptr;
    std::vector<std::string> tl_labels;
    int64_t refresh = 512;        // exact fp64 recomputation of V / log-posterior every this many steps
    int64_t since_refresh = 0;
    explicit DenseTF32Sampler(rmn_sampler* s_) : s(s_) {
        st.K = s->K; st.d = s->model->d; st.dp = (st.d + 31) / 32 * 32;
        if (const char* e = getenv("RMN_TF32_GRAPH")) use_graph = !(e[0] == '0');
        if (const char* e = getenv("RMN_TF32_HALVES")) two_halves = !(e[0] == '0');
        if (const char* e = getenv("RMN_TF32_MIXED")) mixed = !(e[0] == '0');
        if (const char* e = getenv("RMN_TF32_TIMELINE")) tl_path = e;
        if (getenv("RMN_TF32_NARROW") || st.dp % tc::TN != 0) mixed = false;
    }
    // contraction length handed to the GEMM: the k-blocks beyond d hold only padding zeros
    int kdim() const { return (st.d + tc::TK3 - 1) / tc::TK3 * tc::TK3; }
    ~DenseTF32Sampler() override {
        if (gexec) cudaGraphExecDestroy(gexec);
        if (cap_stream) cudaStreamDestroy(cap_stream);
        if (cap_stream2) cudaStreamDestroy(cap_stream2);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        cudaFree(d_Ph); cudaFree(d_Pl); cudaFree(d_Ldiag); cudaFree(d_mupad); cudaFree(d_step_base); cudaFree(d_tl); cudaFree(d_exact);
    }
    size_t rowb() const { return align256((size_t)st.K * st.dp * 4); }
    size_t workspace_bytes() const override {
        const size_t K = (size_t)st.K;
        return 5 * rowb() + 7 * align256(K * 8) +
               2 * align256(ND_MAX * K * 8) + 256;
    }
    int bind(void* ws) override {
        const size_t K = (size_t)st.K;
        char* p = (char*)ws;
        st.Y = (float*)p; p += rowb();   st.V = (float*)p; p += rowb();
        st.Yph = (float*)p; p += rowb(); st.Ypl = (float*)p; p += rowb();
        st.Vp = (float*)p; p += rowb();
        st.lp = (double*)p; p += align256(K * 8);
        st.k0 = (double*)p; p += align256(K * 8);
        st.epsrow = (double*)p; p += align256(K * 8);
        st.scale = (double*)p; p += align256(K * 8);
        st.nsamp = (long long*)p; p += align256(K * 8);
        st.nacc = (long long*)p; p += align256(K * 8);
        st.dacc = (long long*)p; p += align256(K * 8);
        st.S1 = (double*)p; p += align256(ND_MAX * K * 8);
        st.S2 = (double*)p; p += align256(ND_MAX * K * 8);
        RMN_CUDA(cudaMemset(ws, 0, workspace_bytes()));
        const int d = st.d, dp = st.dp;
        std::vector<double> tmp((size_t)d * d), hm(dp, 0.0);
        RMN_CUDA(cudaMemcpy(tmp.data(), s->model->d_prec, (size_t)d * d * 8, cudaMemcpyDeviceToHost));
        std::vector<float> ph((size_t)dp * dp, 0.0f), pl((size_t)dp * dp, 0.0f), ld(dp, 0.0f);
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) {
                float hi, lo;
                tc::split_tf32((float)tmp[(size_t)i * d + j], hi, lo);
                ph[(size_t)i * dp + j] = hi; pl[(size_t)i * dp + j] = lo;
            }
        for (int i = 0; i < d; ++i) hm[i] = s->model->h_mu[i];
        const rmn_proposal* pr = s->prop;
        if (pr->kind == RMN_PROP_RW)
            for (int i = 0; i < d; ++i) ld[i] = (float)pr->h_L[(size_t)i * d + i];
        RMN_CUDA(cudaMalloc(&d_Ph, ph.size() * 4)); RMN_CUDA(cudaMalloc(&d_Pl, pl.size() * 4));
        RMN_CUDA(cudaMalloc(&d_Ldiag, (size_t)dp * 4)); RMN_CUDA(cudaMalloc(&d_mupad, (size_t)dp * 8));
        RMN_CUDA(cudaMemcpy(d_Ph, ph.data(), ph.size() * 4, cudaMemcpyHostToDevice));
        RMN_CUDA(cudaMemcpy(d_Pl, pl.data(), pl.size() * 4, cudaMemcpyHostToDevice));
        RMN_CUDA(cudaMemcpy(d_Ldiag, ld.data(), (size_t)dp * 4, cudaMemcpyHostToDevice));
        RMN_CUDA(cudaMemcpy(d_mupad, hm.data(), (size_t)dp * 8, cudaMemcpyHostToDevice));
        st.mu = d_mupad; st.Ldiag = d_Ldiag; st.prec = s->model->d_prec;
        st.mubar = 0.0;
        for (int i = 0; i < d; ++i) st.mubar += hm[i] / (double)d;
        int rc;
        if ((rc = tc::make_tmap_2d(&maps.ah, st.Yph, st.K, dp, dp, tc::TM, tc::TK3))) return rc;
        if ((rc = tc::make_tmap_2d(&maps.al, st.Ypl, st.K, dp, dp, tc::TM, tc::TK3))) return rc;
        if ((rc = tc::make_tmap_2d(&maps.bh, d_Ph, dp, dp, dp, tc::TN, tc::TK3))) return rc;
        if ((rc = tc::make_tmap_2d(&maps.bl, d_Pl, dp, dp, dp, tc::TN, tc::TK3))) return rc;
        if ((rc = tc::prepare_kernels())) return rc;
        maps_narrow.ah = maps.ah; maps_narrow.al = maps.al;
        if ((rc = tc::make_tmap_2d(&maps_narrow.bh, d_Ph, dp, dp, dp, 128, tc::TK3))) return rc;
        if ((rc = tc::make_tmap_2d(&maps_narrow.bl, d_Pl, dp, dp, dp, 128, tc::TK3))) return rc;
        {
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const int64_t wide = ((st.K + tc::TM - 1) / tc::TM) * ((dp + tc::TN - 1) / tc::TN);
            narrow = wide < sms;
            if (const char* e = getenv("RMN_TF32_NARROW")) narrow = (e[0] == '1');
        }
        {
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            int64_t ma = ((st.K + tc::TM - 1) / tc::TM + 1) / 2;            // row tiles of the first part: half of them, or
            if (const char* e = getenv("RMN_TF32_SPLIT_MTILES")) ma = std::max<int64_t>(1, atoll(e));   // as given
            const int64_t k0 = std::min<int64_t>(st.K, ma * tc::TM);
            row0_h[0] = 0; nrows_h[0] = k0; row0_h[1] = k0; nrows_h[1] = st.K - k0;
            for (int h = 0; h < 2; ++h) {
                if (nrows_h[h] <= 0) continue;
                const size_t off = (size_t)row0_h[h] * dp;
                narrow_h[h] = ((nrows_h[h] + tc::TM - 1) / tc::TM) * ((dp + tc::TN - 1) / tc::TN) < sms;
                if (const char* e = getenv("RMN_TF32_NARROW")) narrow_h[h] = (e[0] == '1');
                const uint32_t brows = (narrow_h[h] || mixed) ? 128 : tc::TN;
                if ((rc = tc::make_tmap_2d(&maps_h[h].ah, st.Yph + off, nrows_h[h], dp, dp, tc::TM, tc::TK3))) return rc;
                if ((rc = tc::make_tmap_2d(&maps_h[h].al, st.Ypl + off, nrows_h[h], dp, dp, tc::TM, tc::TK3))) return rc;
                if ((rc = tc::make_tmap_2d(&maps_h[h].bh, d_Ph, dp, dp, dp, brows, tc::TK3))) return rc;
                if ((rc = tc::make_tmap_2d(&maps_h[h].bl, d_Pl, dp, dp, dp, brows, tc::TK3))) return rc;
            }
        }
        if ((rc = rmn_fill_f64(st.scale, st.K, 1.0, 0))) return rc;
        RMN_CUDA(cudaDeviceSynchronize());
        return RMN_OK;
    }
    unsigned row_grid() const { return (unsigned)((st.K * 32 + 255) / 256); }
    static constexpr int TL_MAX = 4096;
    long long* tl_slot(const char* what, int half, int64_t t) {
        if (!d_tl || (int)tl_labels.size() >= TL_MAX) return nullptr;
        char b[64];
        snprintf(b, sizeof b, "%s h=%d t=%lld", what, half, (long long)t);
        tl_labels.push_back(b);
        return d_tl + 2 * (tl_labels.size() - 1);
    }
    int tl_half = -1; int64_t tl_t = 0;
    void launch_fp(const TStep& sp_in, cudaStream_t stream) {
        TStep sp = sp_in;
        sp.dbg = tl_slot("pass", tl_half, tl_t);
        const int64_t n = sp.nrows ? sp.nrows : st.K;
        finish_propose_f32_kernel<<<(unsigned)((n * 32 + FP_THREADS - 1) / FP_THREADS), FP_THREADS, 0, stream>>>(st, sp);
    }
    int gemm_half(int h, cudaStream_t stream) {
        launches++;
        float* C = st.Vp + (size_t)row0_h[h] * st.dp;
        tc::set_debug_stamp(tl_slot("gemm", h, tl_t));
        if (mixed) return tc::launch_plain_mixed(maps_h[h], nrows_h[h], st.dp, kdim(), C, st.dp, stream);
        if (narrow_h[h]) return tc::launch_plain_narrow(maps_h[h], nrows_h[h], st.dp, kdim(), C, st.dp, stream);
        return tc::launch_plain(maps_h[h], nrows_h[h], st.dp, kdim(), C, st.dp, stream);
    }
    double c1() const { return st.d * log(2.0 * M_PI); }
    // The GEMM keeps a plain, store-only epilogue (V' only); the MH row reductions run in the finish/propose pass.  (A fused
    // MH epilogue was built and measured in round 1: its eight epilogue warps stalled on L2 latency, 235 us against
    // 139 + the row pass's share -- DESIGN.md section 4.)
    int gemm(cudaStream_t stream) {
        launches++;
        ktimer.begin("tf32x3_gemm_kernel", stream);
        tc::set_debug_stamp(tl_slot("gemm", -1, tl_t));
        int rc;
        if (mixed) rc = tc::launch_plain_mixed(maps_narrow, st.K, st.dp, kdim(), st.Vp, st.dp, stream);
        else if (narrow) rc = tc::launch_plain_narrow(maps_narrow, st.K, st.dp, kdim(), st.Vp, st.dp, stream);
        else rc = tc::launch_plain(maps, st.K, st.dp, kdim(), st.Vp, st.dp, stream);
        ktimer.end(stream);
        return rc;
    }
    int set_state(const double* d_theta, cudaStream_t stream) override {
        const int64_t n = st.K * st.dp;
        tset_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(st, d_theta);
        RMN_KERNEL_CHECK(); launches++;
        return exact(stream);
    }
    // Exact (fp64) V = Y P and log-posterior of the CURRENT fp32 states: start of a run and the periodic refresh, as one
    // fp64 DMMA GEMM on widened copies (dense.cu: dense_exact_pass)
    void* d_exact = nullptr;
    int exact(cudaStream_t stream) {
        if (!d_exact) {
            RMN_CUDA(cudaMalloc(&d_exact, dense_exact_scratch_bytes(st.K, st.d)));
            if (int rc = dense_exact_prepare(st.K, st.d, st.prec, d_exact)) return rc;
        }
        if (int rc = dense_exact_pass(st.K, st.d, st.dp, st.Y, st.V, st.lp, c1(), s->model->logdetC, d_exact, stream)) return rc;
        launches += 3;
        since_refresh = 0;
        return RMN_OK;
    }
    int get_state(double* d_theta, double* d_lp, cudaStream_t stream) override {
        const int64_t n = st.K * st.d;
        tget_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(st, d_theta, d_lp);
        RMN_KERNEL_CHECK(); launches++;
        return RMN_OK;
    }
    int run(int64_t T, const rmn_inject_t* inj, const rmn_trace_t* tr, cudaStream_t stream) override {
        const bool traced = tr && (tr->d_theta || tr->d_logpost || tr->d_prop_logpost || tr->d_accepted ||
                                   tr->d_logqratio || tr->d_prop_theta);
        const bool graphable = use_graph && !inj && !traced && !ktimer.on && T >= 2 &&
                               !(refresh > 0 && since_refresh + T >= refresh);
        if (!graphable) return run_steps(T, inj, tr, stream, nullptr, -1);
        if (!d_step_base) RMN_CUDA(cudaMalloc(&d_step_base, 8));
        if (!gexec || g_T != T) {
            // capture on a stream of our own (the caller's may be the legacy default stream, which cannot capture);
            // nothing executes during capture, and the graph is then launched into the CALLER's stream
            if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
            if (!cap_stream && cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
                cudaGetLastError(); use_graph = false;
                return run_steps(T, inj, tr, stream, nullptr, -1);
            }
            if (!cap_stream2 && two_halves) {
                if (cudaStreamCreateWithFlags(&cap_stream2, cudaStreamNonBlocking) != cudaSuccess ||
                    cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) != cudaSuccess) {
                    cudaGetLastError(); cap_stream2 = nullptr;
                }
            }
            if (!tl_path.empty()) {
                if (!d_tl) RMN_CUDA(cudaMalloc(&d_tl, (size_t)TL_MAX * 16));
                tl_labels.clear();
            }
            if (cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
                cudaGetLastError(); use_graph = false;
                return run_steps(T, inj, tr, stream, nullptr, -1);
            }
            const int64_t l0 = launches, s0 = step0, d0 = diag_steps, r0 = since_refresh;
            cudaGraph_t graph = nullptr;
            int rc = RMN_OK;
            if (nrows_h[1] >= 4096 && cap_stream2 && ev_fork && ev_join) {
                // fork: branch 0 on cap_stream, branch 1 on cap_stream2, joined before the capture ends
                cudaEventRecord(ev_fork, cap_stream);
                cudaStreamWaitEvent(cap_stream2, ev_fork, 0);
                rc = run_steps(T, nullptr, nullptr, cap_stream, d_step_base, 0);
                const int64_t l1 = launches, s1 = step0, d1 = diag_steps, r1 = since_refresh;
                step0 = s0; diag_steps = d0; since_refresh = r0;
                if (rc == RMN_OK) rc = run_steps(T, nullptr, nullptr, cap_stream2, d_step_base, 1);
                (void)l1; step0 = s1; diag_steps = d1; since_refresh = r1;
                cudaEventRecord(ev_join, cap_stream2);
                cudaStreamWaitEvent(cap_stream, ev_join, 0);
            } else {
                rc = run_steps(T, nullptr, nullptr, cap_stream, d_step_base, -1);
            }
            const cudaError_t ce = cudaStreamEndCapture(cap_stream, &graph);
            g_nodes = launches - l0;
            launches = l0; step0 = s0; diag_steps = d0; since_refresh = r0;       // nothing ran yet
            if (rc != RMN_OK || ce != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                use_graph = false;                                               // plain launches from now on
                return run_steps(T, inj, tr, stream, nullptr, -1);
            }
            const cudaError_t ie = cudaGraphInstantiate(&gexec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) { gexec = nullptr; cudaGetLastError(); use_graph = false; return run_steps(T, inj, tr, stream, nullptr, -1); }
            g_T = T;
        }
        tstep_base_kernel<<<1, 1, 0, stream>>>(d_step_base, step0);
        RMN_KERNEL_CHECK();
        if (d_tl) {
            std::vector<long long> init(2 * TL_MAX);
            for (int i = 0; i < TL_MAX; ++i) { init[2 * i] = LLONG_MAX; init[2 * i + 1] = 0; }
            RMN_CUDA(cudaMemcpyAsync(d_tl, init.data(), init.size() * 8, cudaMemcpyHostToDevice, stream));
            RMN_CUDA(cudaStreamSynchronize(stream));
        }
        RMN_CUDA(cudaGraphLaunch(gexec, stream));
        if (d_tl) {
            std::vector<long long> h(2 * tl_labels.size());
            RMN_CUDA(cudaStreamSynchronize(stream));
            RMN_CUDA(cudaMemcpy(h.data(), d_tl, h.size() * 8, cudaMemcpyDeviceToHost));
            if (FILE* f = fopen(tl_path.c_str(), "a")) {
                fprintf(f, "# replay T=%lld K=%lld\n", (long long)T, (long long)st.K);
                for (size_t i = 0; i < tl_labels.size(); ++i) fprintf(f, "%s %lld %lld\n", tl_labels[i].c_str(), h[2 * i], h[2 * i + 1]);
                fclose(f);
            }
        }
        launches += g_nodes + 1;
        step0 += T; diag_steps += T; since_refresh += T;
        return RMN_OK;
    }
    // the loop itself; d_base != NULL: step indices relative to *d_base (graph capture)
    // half = 0 | 1: only that half-batch of rows (graph branches); -1: all rows
    int run_steps(int64_t T, const rmn_inject_t* inj, const rmn_trace_t* tr, cudaStream_t stream, const int64_t* d_base,
                  int half) {
        const rmn_proposal* pr = s->prop;
        if (inj) RMN_REQUIRE(inj->d_xi && inj->d_u, "injected run needs d_xi and d_u");
        rmn_trace_t t0{};
        if (tr) t0 = *tr;
        if (t0.thin <= 0) t0.thin = 1;
        TStep sp{};
        sp.prop_kind = pr->kind; sp.adapt = pr->adapt; sp.target = pr->target; sp.eps0 = pr->eps;
        sp.c1 = c1(); sp.c2 = s->model->logdetC; sp.seed = s->seed; sp.chain_offset = s->chain_offset;
        sp.d_step_base = d_base;
        if (half >= 0) { sp.row0 = row0_h[half]; sp.nrows = nrows_h[half]; }
        const int64_t sb = d_base ? 0 : step0;
        const int64_t K = st.K;
        const int d = st.d;
        for (int64_t t = 0; t <= T; ++t) {
            tl_half = half; tl_t = t;
            sp.finish = (t > 0); sp.propose = (t < T); sp.diag = (t > 0);
            sp.step_fin = sb + t - 1; sp.step_prop = sb + t;
            sp.inj_u = (inj && t > 0) ? inj->d_u + (t - 1) * K : nullptr;
            sp.inj_xi = (inj && t < T) ? inj->d_xi + t * K * d : nullptr;
            sp.trace_slot = -1;
            sp.tr_theta = t0.d_theta; sp.tr_logpost = t0.d_logpost;
            sp.tr_prop_lp = (t0.d_prop_logpost && t > 0) ? t0.d_prop_logpost + (t - 1) * K : nullptr;
            sp.tr_acc = (t0.d_accepted && t > 0) ? t0.d_accepted + (t - 1) * K : nullptr;
            sp.tr_lqr = (t0.d_logqratio && t > 0) ? t0.d_logqratio + (t - 1) * K : nullptr;
            sp.tr_prop_theta = (t0.d_prop_theta && t > 0) ? t0.d_prop_theta + (t - 1) * K * d : nullptr;
            if (t > 0 && (t0.d_theta || t0.d_logpost)) {
                const int64_t i = t;
                if (i >= t0.first && (i - t0.first) % t0.thin == 0) sp.trace_slot = (i - t0.first) / t0.thin;
            }
            // the proposal of step t is drawn from (y, V): refresh those exactly first when due.
            // A pending proposal must be finished before V is overwritten, so split the pass.
            if (refresh > 0 && since_refresh >= refresh && sp.finish && sp.propose) {
                TStep fin = sp; fin.propose = 0;
                launch_fp(fin, stream);
                RMN_KERNEL_CHECK(); launches++;
                if (int rc = exact(stream)) return rc;
                TStep pro = sp; pro.finish = 0; pro.diag = 0; pro.trace_slot = -1;
                pro.tr_prop_lp = nullptr; pro.tr_acc = nullptr; pro.tr_lqr = nullptr; pro.tr_prop_theta = nullptr;
                launch_fp(pro, stream);
                RMN_KERNEL_CHECK(); launches++;
            } else {
                launch_fp(sp, stream);
                RMN_KERNEL_CHECK(); launches++;
            }
            if (t == T) break;
            since_refresh++;
            if (int rc = (half >= 0) ? gemm_half(half, stream) : gemm(stream)) return rc;
        }
        step0 += T; diag_steps += T;
        return RMN_OK;
    }
    int get_adapt(double* sc, int64_t* ns, int64_t* na, cudaStream_t stream) override {
        tget_adapt_kernel<<<(unsigned)((st.K + 127) / 128), 128, 0, stream>>>(st, sc, ns, na);
        RMN_KERNEL_CHECK(); launches++;
        return RMN_OK;
    }
    int set_adapt(const double* sc, const int64_t* ns, const int64_t* na, cudaStream_t stream) override {
        launches++;
        return rmn_copy_adapt(st.K, sc, ns, na, st.scale, st.nsamp, st.nacc, stream);
    }
    int diag_dim() const override { return (st.d < ND_MAX - 1 ? st.d : ND_MAX - 1) + 1; }
    int reset_diag(cudaStream_t stream) override {
        RMN_CUDA(cudaMemsetAsync(st.S1, 0, (size_t)ND_MAX * st.K * 8, stream));
        RMN_CUDA(cudaMemsetAsync(st.S2, 0, (size_t)ND_MAX * st.K * 8, stream));
        RMN_CUDA(cudaMemsetAsync(st.dacc, 0, (size_t)st.K * 8, stream));
        diag_steps = 0;
        return RMN_OK;
    }
    int chain_sums(const double** S1, const double** S2, int64_t* n) override {
        *S1 = st.S1; *S2 = st.S2; *n = diag_steps;
        return RMN_OK;
    }
    int reduce_diag(double* d_block, cudaStream_t stream) override {
        launches++;
        return rmn_reduce_diag_block(st.K, diag_dim(), diag_steps, diag_steps, st.S1, st.S2, st.dacc, nullptr, d_block, stream);
    }
};

}  // namespace

SamplerImpl* make_dense_tf32_sampler(rmn_sampler* s) {
    const rmn_proposal* p = s->prop;
    if (s->model->kind != RMN_MODEL_GAUSS) {
        rmn_set_error("tf32x3 precision is implemented for the dense Gaussian model");
        return nullptr;
    }
    if (p->kind == RMN_PROP_HMC && p->nsteps == 1 && !p->has_mass) return new DenseTF32Sampler(s);
    if (p->kind == RMN_PROP_RW) {
        const int d = s->model->d;
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < i; ++j)
                if (p->h_L[(size_t)i * d + j] != 0.0) {
                    rmn_set_error("tf32x3 random walk supports diagonal proposal covariances");
                    return nullptr;
                }
        return new DenseTF32Sampler(s);
    }
    rmn_set_error("tf32x3 precision supports RW (diagonal covariance) and MALA (VanillaHMC Nsteps=1, no mass matrix)");
    return nullptr;
}
