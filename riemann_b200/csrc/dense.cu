// riemann_b200 -- dense Gaussian target at any d (config 3: d = 1000, MALA, 16,384 chains).
//
// Per MH iteration and chain the reference does (fp64 numpy)
//   VanillaHMC.propose, Nsteps = 1 (= MALA)     riemann/proposals/hamiltonian.py:76-91, 13-52
//   MetropolisRandomWalk.propose               riemann/proposals/randomwalk.py:21-26
//   MultiGaussianDist.log_likelihood / grad    riemann/models/gaussian.py:49-58  (O(d^3) LU per call)
//   Sampler.sample accept/swap, adapt          riemann/samplers/sampler.py:72-90, adaptive.py:26-35
// Here one product  V' = Y' P  (Y' = proposed states minus mu, K x d;  P = C^-1, d x d,
// symmetric, L2-resident) gives both the gradient (-V') and the quadratic form
// (rowsum(Y' * V')) of every chain; gradient and log-posterior of the current state are
// cached, so a step costs ONE batched GEMM instead of the reference's three solves.
//
//   kernel 1  finish_propose : per row -- finish the previous step (sum the GEMM's row
//             partials, log-posterior, log q ratio, accept, adapt), then draw xi (Philox or
//             injected) and write the next proposal.  One warp per chain row, coalesced.
//             pCN (randomwalk.py:78-100): two extra GEMMs per step, theta' = rho theta + rho_c xi L^T and
//             the reverse residual (rho_c L)^-1 (theta - rho theta'), whose squared norm is the log q ratio.
//             Nsteps > 1 (leapfrog, hamiltonian.py:13-52): Nsteps - 1 extra [gradient GEMM,
//             leapfrog_mid_kernel] pairs advance the trajectory in place on the proposal slot.
//   kernel 2  gemm_abt       : C = A B^T in fp64 on the tensor cores (DMMA m8n8k4),
//             128x128x16 tiles, 3-stage cp.async pipeline, 8 warps.  The epilogue is fused:
//             it stores V', and reduces y'.v' and |p'|^2 over the tile's columns into
//             deterministic per-(row, column-block) partials (no atomics).
//
// Accepting a proposal never copies a row: every chain has two state slots and a one-bit
// "current slot" flag; the proposal is written to the other slot and accept flips the bit.
// Internal state is centred (y = theta - mu), row-major [K][dp], dp = d rounded up to 16
// (padding columns are exactly zero in Y, xi and P, so they contribute nothing).
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 3, LDT = 20;   // LDT: smem row stride (doubles)
constexpr int GEMM_THREADS = 256;
constexpr size_t GEMM_SMEM = (size_t)STAGES * (BM + BN) * LDT * 8;
constexpr int ND_MAX = 8;          // tracked functionals: first 7 coordinates + mean(theta)

enum { EPI_LOGPOST_RW = 0, EPI_LOGPOST_MALA = 1, EPI_RWPROP = 2, EPI_PCNPROP = 3, EPI_PCNREV = 4,
       EPI_MASS_P0 = 5, EPI_MASS_STEP = 6, EPI_LOGPOST_MASS = 7, EPI_MASS_W1 = 8 };

struct DenseState {
    int64_t K; int d, dp, nblk;    // nblk = column blocks of 32 (partials per row)
    double* Y;       // [2][K][dp]   centred states, two slots per chain
    double* V;       // [2][K][dp]   V = Y P  (gradient = -V)
    double* Xi;      // [K][dp]      noise of the pending proposal (MALA) / Xi L^T scratch (dense RW)
    double* partq;   // [nblk][K]    row partials of y'.v'
    double* partk;   // [nblk][K]    row partials of |p'|^2
    double* lp;      // [K]
    double* k0;      // [K]          |xi|^2 of the pending proposal
    double* epsrow;  // [K]          step size used by the pending proposal
    int* cur;        // [K]          current slot
    double* scale; long long* nsamp; long long* nacc;   // AdaptScale state
    long long* dacc;               // accepts since diagnostics reset
    double* S1; double* S2;        // [ND_MAX][K]
    const double* mu;              // [dp] (padded copy)
    double mubar;                  // mean(mu): the tracked functionals are those of theta, not of the centred state
    double rho, rho_c;             // pCN (randomwalk.py:83-86)
    double* Pm;                    // [K][dp]   momentum of the trajectory (HMC with a mass matrix)
    int mass_base_prop;            // EPI_MASS_STEP: the position step starts from the proposal slot (interior steps)
    // parallel tempering (ptsampler.py:11-127): nt = 0 off; rows l*nt .. l*nt + nt - 1 form ladder l
    int nt; double pswap;
    const double* betas;           // [nt]
    double* ll;                    // [K] untempered log-likelihood of the current state (lp = beta * ll)
    unsigned char* role;           // [K] role in the pending PT step: 0 within-chain step, 1 swap initiator, 2 partner
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(n));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// C[m][n] = sum_k A[m][k] B[n][k].  A rows are the PROPOSAL slots of the chains
// (EPI_LOGPOST_*) or the noise scratch (EPI_RWPROP); B is P (symmetric) or L.
template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_abt_kernel(DenseState st, const double* __restrict__ B) {
    extern __shared__ __align__(16) double sm[];
    double* As = sm;
    double* Bs = sm + (size_t)STAGES * BM * LDT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;              // 2 x 4 warps, warp tile 64 x 32
    const int64_t m0 = (int64_t)blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const int dp = st.dp;
    const int64_t K = st.K;
    const int KT = dp / BK;

    // each thread copies 4 A-chunks and 4 B-chunks (16 B each) per k-tile
    const double* a_src[4];
    const double* b_src[4];
    bool a_ok[4], b_ok[4];
    int soff[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = tid + GEMM_THREADS * i;
        const int row = q >> 3, c2 = (q & 7) * 2;
        soff[i] = row * LDT + c2;
        const int64_t m = m0 + row;
        a_ok[i] = m < K;
        const int64_t mm = a_ok[i] ? m : 0;
        if (EPI == EPI_RWPROP || EPI == EPI_PCNPROP || EPI == EPI_MASS_P0) a_src[i] = st.Xi + mm * dp + c2;
        else if (EPI == EPI_MASS_STEP || EPI == EPI_MASS_W1) a_src[i] = st.Pm + mm * dp + c2;
        else if (EPI == EPI_PCNREV) a_src[i] = st.V + ((int64_t)(st.cur[mm] ^ 1) * K + mm) * dp + c2;   // D, see below
        else a_src[i] = st.Y + ((int64_t)(st.cur[mm] ^ 1) * K + mm) * dp + c2;
        const int n = n0 + row;
        b_ok[i] = n < dp;
        b_src[i] = B + (int64_t)(b_ok[i] ? n : 0) * dp + c2;
    }
    auto load_tile = [&](int kt, int stage) {
        double* as = As + (size_t)stage * BM * LDT;
        double* bs = Bs + (size_t)stage * BN * LDT;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            cp_async16(as + soff[i], a_src[i] + kt * BK, a_ok[i]);
            cp_async16(bs + soff[i], b_src[i] + kt * BK, b_ok[i]);
        }
    };

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_tile(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nk = kt + STAGES - 1;
        if (nk < KT) load_tile(nk, nk % STAGES);
        cp_async_commit();
        const double* as = As + (size_t)(kt % STAGES) * BM * LDT + (wm * 64 + g) * LDT + t;
        const double* bs = Bs + (size_t)(kt % STAGES) * BN * LDT + (wn * 32 + g) * LDT + t;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = as[i * 8 * LDT + kk * 4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = bs[j * 8 * LDT + kk * 4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // ---- fused epilogue.  Thread owns rows m0 + wm*64 + i*8 + g, cols n0 + wn*32 + j*8 + 2t (+1)
    const int blk = blockIdx.x * 4 + wn;            // 32-column block index of this warp
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + wm * 64 + i * 8 + g;
        const bool rok = m < K;
        const int64_t mm = rok ? m : 0;
        const int c = st.cur[mm];
        double pq = 0.0, pk = 0.0;
        if (EPI == EPI_MASS_P0) {
            // p0 = chM xi (hamiltonian.py:81), then the initial half step p = p0 + eps/2 g(theta), g = -V (:27)
            const double he = 0.5 * st.epsrow[mm];
            const double* vc = st.V + ((int64_t)c * K + mm) * dp;
            double* pm = st.Pm + mm * dp;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + wn * 32 + j * 8 + 2 * t;
                if (rok && n < dp) {
                    const double2 w = *reinterpret_cast<const double2*>(vc + n);
                    *reinterpret_cast<double2*>(pm + n) = make_double2(acc[i][j][0] - he * w.x, acc[i][j][1] - he * w.y);
                }
            }
        } else if (EPI == EPI_MASS_STEP) {
            // position step q <- q + eps M^-1 p (hamiltonian.py:29-30, 36-37): acc = p Minv^T
            const double eps = st.epsrow[mm];
            const double* yb = st.Y + ((int64_t)(st.mass_base_prop ? (c ^ 1) : c) * K + mm) * dp;
            double* yp = st.Y + ((int64_t)(c ^ 1) * K + mm) * dp;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + wn * 32 + j * 8 + 2 * t;
                if (rok && n < dp) {
                    const double2 y = *reinterpret_cast<const double2*>(yb + n);
                    *reinterpret_cast<double2*>(yp + n) = make_double2((n < st.d) ? y.x + eps * acc[i][j][0] : 0.0,
                                                                       (n + 1 < st.d) ? y.y + eps * acc[i][j][1] : 0.0);
                }
            }
        } else if (EPI == EPI_MASS_W1) {
            // whitened final momentum solve(chM, p') (hamiltonian.py:87): acc = p' chMinv^T; only its squared norm is needed
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + wn * 32 + j * 8 + 2 * t;
                if (rok && n < dp) pk += acc[i][j][0] * acc[i][j][0] + acc[i][j][1] * acc[i][j][1];
            }
            pk += __shfl_xor_sync(0xffffffffu, pk, 1);
            pk += __shfl_xor_sync(0xffffffffu, pk, 2);
            if (rok && t == 0 && blk < st.nblk) st.partk[(int64_t)blk * K + m] = pk;
        } else if (EPI == EPI_PCNPROP) {
            // pCN (randomwalk.py:88-94): theta' = rho theta + rho_c L xi; the residual of the REVERSE move,
            // D = theta - rho theta', is parked in the proposal slot of V (free until the log-posterior GEMM)
            const double* yc = st.Y + ((int64_t)c * K + mm) * dp;
            double* yp = st.Y + ((int64_t)(c ^ 1) * K + mm) * dp;
            double* dv = st.V + ((int64_t)(c ^ 1) * K + mm) * dp;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + wn * 32 + j * 8 + 2 * t;
                if (rok && n < dp) {
                    const double2 y = *reinterpret_cast<const double2*>(yc + n);
                    const double2 mu = *reinterpret_cast<const double2*>(st.mu + n);
                    const double th0 = y.x + mu.x, th1 = y.y + mu.y;
                    const double tp0 = st.rho * th0 + st.rho_c * acc[i][j][0];
                    const double tp1 = st.rho * th1 + st.rho_c * acc[i][j][1];
                    const bool in0 = n < st.d, in1 = n + 1 < st.d;
                    *reinterpret_cast<double2*>(yp + n) = make_double2(in0 ? tp0 - mu.x : 0.0, in1 ? tp1 - mu.y : 0.0);
                    *reinterpret_cast<double2*>(dv + n) = make_double2(in0 ? th0 - st.rho * tp0 : 0.0,
                                                                       in1 ? th1 - st.rho * tp1 : 0.0);
                }
            }
        } else if (EPI == EPI_PCNREV) {
            // u_rev = (rho_c L)^-1 D  (randomwalk.py:98): acc = D Linv^T; only |u_rev|^2 is needed
            const double irc = 1.0 / st.rho_c;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + wn * 32 + j * 8 + 2 * t;
                if (rok && n < dp) {
                    const double u0 = acc[i][j][0] * irc, u1 = acc[i][j][1] * irc;
                    pk += u0 * u0 + u1 * u1;
                }
            }
            pk += __shfl_xor_sync(0xffffffffu, pk, 1);
            pk += __shfl_xor_sync(0xffffffffu, pk, 2);
            if (rok && t == 0 && blk < st.nblk) st.partk[(int64_t)blk * K + m] = pk;
        } else if (EPI == EPI_RWPROP) {
            const double sc = st.epsrow[mm];
            const double* yc = st.Y + ((int64_t)c * K + mm) * dp;
            double* yp = st.Y + ((int64_t)(c ^ 1) * K + mm) * dp;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + wn * 32 + j * 8 + 2 * t;
                if (rok && n < dp) {
                    const double2 y = *reinterpret_cast<const double2*>(yc + n);
                    double2 o;
                    o.x = (n < st.d) ? y.x + sc * acc[i][j][0] : 0.0;
                    o.y = (n + 1 < st.d) ? y.y + sc * acc[i][j][1] : 0.0;
                    *reinterpret_cast<double2*>(yp + n) = o;
                }
            }
        } else {
            const double* yp = st.Y + ((int64_t)(c ^ 1) * K + mm) * dp;
            double* vp = st.V + ((int64_t)(c ^ 1) * K + mm) * dp;
            const double* xi = st.Xi + mm * dp;
            const double he = 0.5 * st.epsrow[mm];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + wn * 32 + j * 8 + 2 * t;
                if (rok && n < dp) {
                    const double v0 = acc[i][j][0], v1 = acc[i][j][1];
                    *reinterpret_cast<double2*>(vp + n) = make_double2(v0, v1);
                    const double2 y = *reinterpret_cast<const double2*>(yp + n);
                    pq += y.x * v0 + y.y * v1;
                    if (EPI == EPI_LOGPOST_MASS) {
                        // final half step of the momentum, kept for the whitening GEMM: p' = p + eps/2 g' (:40)
                        double* pm = st.Pm + mm * dp;
                        const double2 p = *reinterpret_cast<const double2*>(pm + n);
                        *reinterpret_cast<double2*>(pm + n) = make_double2(p.x - he * v0, p.y - he * v1);
                    }
                    if (EPI == EPI_LOGPOST_MALA) {
                        // final half step  p' = p + eps/2 g',  g' = -V'   (hamiltonian.py:40); Xi holds p
                        const double2 x = *reinterpret_cast<const double2*>(xi + n);
                        const double p0 = x.x - he * v0;
                        const double p1 = x.y - he * v1;
                        pk += p0 * p0 + p1 * p1;
                    }
                }
            }
            pq += __shfl_xor_sync(0xffffffffu, pq, 1);
            pq += __shfl_xor_sync(0xffffffffu, pq, 2);
            pk += __shfl_xor_sync(0xffffffffu, pk, 1);
            pk += __shfl_xor_sync(0xffffffffu, pk, 2);
            if (rok && t == 0 && blk < st.nblk) {
                st.partq[(int64_t)blk * K + m] = pq;
                if (EPI == EPI_LOGPOST_MALA) st.partk[(int64_t)blk * K + m] = pk;
            }
        }
    }
}

struct DenseStep {
    int prop_kind;          // RMN_PROP_RW / RMN_PROP_HMC
    int adapt, rw_diag, finish, propose, diag, has_mass;
    int record;             // write trace / diagnostics of the state after the finished step (= finish, except under
                            // tempering, where a swap touches two rows and the record is taken by the NEXT launch)
    const double* inj_usel; // tempering: injected selection uniforms of the step being proposed
    double target, eps0, c1, c2;
    const double* Ldiag;    // [dp] diagonal of chol(C0) when the RW covariance is diagonal
    uint64_t seed; int64_t chain_offset;
    int64_t step_fin;       // global index of the step being finished
    int64_t step_prop;      // global index of the step being proposed
    const double* inj_xi; const double* inj_u;      // already offset to the respective step
    int64_t trace_slot;     // record slot for the state after the finished step, or -1
    double* tr_theta; double* tr_logpost; double* tr_prop_lp; uint8_t* tr_acc;   // offset to step
    double* tr_lqr; double* tr_prop_theta;                                       // offset to step
};

// One warp per chain row.
__global__ void __launch_bounds__(256)
finish_propose_kernel(DenseState st, DenseStep sp) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= st.K) return;
    const int dp = st.dp, d = st.d;
    const int64_t K = st.K;
    int c = st.cur[r];
    double lp = st.lp[r];
    const RngKey rk(sp.seed, (uint64_t)(sp.chain_offset + r));

    const bool pt = st.nt > 0;
    const int ti = pt ? (int)((sp.chain_offset + r) % st.nt) : 0;
    const double beta = pt ? st.betas[ti] : 1.0;
    const int role = (pt && sp.finish) ? st.role[r] : 0;
    if (sp.finish && role == 1) {
        // ---- swap proposal with the next rung (ptsampler.py:113-125); this warp moves BOTH rows
        const int64_t rj = r + 1;
        const double beta_j = st.betas[ti + 1];
        const double ll_i = st.ll[r], ll_j = st.ll[rj], lp_j = st.lp[rj];
        const double lp_ij = combine_logpost(0.0, ll_j * beta);       // model_i at theta_j
        const double lp_ji = combine_logpost(0.0, ll_i * beta_j);     // model_j at theta_i
        const double x = exp((lp_ji + lp_ij) - (lp + lp_j));
        const double mhr = (x < 1.0) ? x : 1.0;                        // Python min(1, x): nan -> 1
        const double u = sp.inj_u ? sp.inj_u[r] : u01(rk.block((uint64_t)sp.step_fin, RMN_BLOCK_ACCEPT).x);
        const bool sw = u < mhr;                                       // :121
        if (sp.tr_acc && lane == 0) { sp.tr_acc[r] = sw ? 1 : 0; sp.tr_acc[rj] = sw ? 1 : 0; }
        if (sp.tr_prop_lp && lane == 0) { sp.tr_prop_lp[r] = lp_ij; sp.tr_prop_lp[rj] = lp_ji; }
        if (sw) {
            const int cj = st.cur[rj];
            double* yi = st.Y + ((int64_t)c * K + r) * dp;  double* yj = st.Y + ((int64_t)cj * K + rj) * dp;
            double* vi = st.V + ((int64_t)c * K + r) * dp;  double* vj = st.V + ((int64_t)cj * K + rj) * dp;
            for (int j = lane; j < dp; j += 32) {
                const double a = yi[j], b = yj[j]; yi[j] = b; yj[j] = a;
                const double e = vi[j], f = vj[j]; vi[j] = f; vj[j] = e;
            }
            if (lane == 0) { st.ll[r] = ll_j; st.ll[rj] = ll_i; st.lp[r] = lp_ij; st.lp[rj] = lp_ji; }
        }
    } else if (sp.finish && role == 0) {
        double q = 0.0, k1 = 0.0;
        for (int b = lane; b < st.nblk; b += 32) {
            q += st.partq[(int64_t)b * K + r];
            if (sp.prop_kind != RMN_PROP_RW) k1 += st.partk[(int64_t)b * K + r];
        }
        // fixed-order reduction => bitwise reproducible log-posteriors
        q = group_sum<32>(q);
        k1 = group_sum<32>(k1);
        const double ll = -0.5 * ((q + sp.c1) + sp.c2);            // gaussian.py:52
        const double lpn = combine_logpost(0.0, pt ? ll * beta : ll);   // TemperedModel: logL * beta (ptsampler.py:33-34)
        // HMC: kinetic-energy difference (hamiltonian.py:89); pCN: -(|u_fwd|^2 - |u_rev|^2)/2 with
        // u_fwd = (rho_c L)^-1 (theta' - rho theta) = xi (randomwalk.py:95-100)
        const double lqr = (sp.prop_kind == RMN_PROP_HMC) ? 0.5 * (k1 - st.k0[r])
                         : ((sp.prop_kind == RMN_PROP_PCN) ? -0.5 * (st.k0[r] - k1) : 0.0);
        const double u = sp.inj_u ? sp.inj_u[r] : u01(rk.block((uint64_t)sp.step_fin, RMN_BLOCK_ACCEPT).x);
        const bool acc = mh_accept(lpn, lp, lqr, u);
        if (sp.tr_prop_theta) {                  // the proposal = what Proposal.propose returned
            const double* ypr = st.Y + ((int64_t)(c ^ 1) * K + r) * dp;
            for (int j = lane; j < d; j += 32) sp.tr_prop_theta[r * d + j] = ypr[j] + st.mu[j];
        }
        if (sp.tr_lqr && lane == 0) sp.tr_lqr[r] = lqr;
        if (acc) { c ^= 1; lp = lpn; }
        if (lane == 0) {
            if (acc) { st.cur[r] = c; st.lp[r] = lp; if (pt) st.ll[r] = ll; }
            st.dacc[r] += acc ? 1 : 0;
            if (sp.adapt) {
                AdaptState ad{st.scale[r], st.nsamp[r], st.nacc[r]};
                ad.update(acc, sp.target);       // a.s. identical to any(theta != last_theta)
                st.scale[r] = ad.scale; st.nsamp[r] = ad.nsamples; st.nacc[r] = ad.naccepts;
            }
            if (sp.tr_prop_lp) sp.tr_prop_lp[r] = lpn;
            if (sp.tr_acc) sp.tr_acc[r] = acc ? 1 : 0;
            if (!pt && sp.trace_slot >= 0 && sp.tr_logpost) sp.tr_logpost[sp.trace_slot * K + r] = lp;
        }
    }
    if (!sp.propose && !sp.record) return;           // tempering: the finish-only launch ends here
    if (pt && sp.record && lane == 0 && sp.trace_slot >= 0 && sp.tr_logpost) sp.tr_logpost[sp.trace_slot * K + r] = lp;

    // ---- tempering: roles of the step being proposed (ptsampler.py:102-112): chain i initiates a swap with i+1 iff it
    //      was not itself swapped by i-1, u <= Pswap and it is not the last rung; sequential along the ladder
    int role_next = 0;
    if (pt && sp.propose) {
        const int64_t r0 = r - ti;
        double usel = 1.0;
        if (lane < st.nt) {
            usel = sp.inj_usel ? sp.inj_usel[r0 + lane]
                               : u01(RngKey(sp.seed, (uint64_t)(sp.chain_offset + r0 + lane)).block((uint64_t)sp.step_prop, RMN_BLOCK_AUX).x);
        }
        const unsigned wmask = __ballot_sync(0xffffffffu, lane < st.nt - 1 && !(usel > st.pswap));
        unsigned init = 0;
        for (int b = 0; b < st.nt - 1; ++b)
            if (((wmask >> b) & 1u) && !(b > 0 && ((init >> (b - 1)) & 1u))) init |= 1u << b;
        role_next = ((init >> ti) & 1u) ? 1 : ((ti > 0 && ((init >> (ti - 1)) & 1u)) ? 2 : 0);
        if (lane == 0) st.role[r] = (unsigned char)role_next;
    }
    __syncwarp();

    const double* y = st.Y + ((int64_t)c * K + r) * dp;
    const double* v = st.V + ((int64_t)c * K + r) * dp;
    double* yp = st.Y + ((int64_t)(c ^ 1) * K + r) * dp;
    double* xo = st.Xi + r * dp;
    const double scale = sp.adapt ? st.scale[r] : 1.0;
    const double eps = (sp.prop_kind == RMN_PROP_HMC) ? scale * sp.eps0 : scale;
    double k0 = 0.0, rowsum = 0.0;
    const bool want_trace = sp.record && sp.trace_slot >= 0 && sp.tr_theta;
    const bool do_propose = sp.propose && role_next == 0;    // swap rows draw no proposal (ptsampler.py:106-125)

    for (int j4 = lane * 4; j4 < dp; j4 += 128) {
        const double2 ya = *reinterpret_cast<const double2*>(y + j4);
        const double2 yb = *reinterpret_cast<const double2*>(y + j4 + 2);
        const double yv[4] = {ya.x, ya.y, yb.x, yb.y};
        rowsum += (yv[0] + yv[1]) + (yv[2] + yv[3]);
        if (want_trace) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (j4 + q < d) sp.tr_theta[(sp.trace_slot * K + r) * d + j4 + q] = yv[q] + st.mu[j4 + q];
        }
        if (!do_propose) continue;
        double xi[4];
        if (sp.inj_xi) {
#pragma unroll
            for (int q = 0; q < 4; ++q) xi[q] = (j4 + q < d) ? sp.inj_xi[r * d + j4 + q] : 0.0;
        } else {
            normal4(rk.block((uint64_t)sp.step_prop, (uint32_t)(j4 >> 2)), xi);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (j4 + q >= d) xi[q] = 0.0;
        }
        double out[4];
        if (sp.prop_kind == RMN_PROP_HMC && !sp.has_mass) {
            const double2 va = *reinterpret_cast<const double2*>(v + j4);
            const double2 vb = *reinterpret_cast<const double2*>(v + j4 + 2);
            const double vv[4] = {va.x, va.y, vb.x, vb.y};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double ph = xi[q] + 0.5 * eps * (-vv[q]);        // hamiltonian.py:27
                out[q] = yv[q] + eps * ph;                             // :30
                k0 += xi[q] * xi[q];
                xi[q] = ph;
            }
            // the momentum after the initial half step rides in the Xi buffer through the trajectory
            *reinterpret_cast<double2*>(xo + j4) = make_double2(xi[0], xi[1]);
            *reinterpret_cast<double2*>(xo + j4 + 2) = make_double2(xi[2], xi[3]);
        } else if (sp.prop_kind == RMN_PROP_RW && sp.rw_diag) {
#pragma unroll
            for (int q = 0; q < 4; ++q) out[q] = yv[q] + scale * (sp.Ldiag[j4 + q] * xi[q]);   // randomwalk.py:26
        } else {
            // dense L: write xi; a GEMM (EPI_RWPROP / EPI_PCNPROP) forms y + scale * xi L^T resp. the pCN move
            *reinterpret_cast<double2*>(xo + j4) = make_double2(xi[0], xi[1]);
            *reinterpret_cast<double2*>(xo + j4 + 2) = make_double2(xi[2], xi[3]);
            k0 += (xi[0] * xi[0] + xi[1] * xi[1]) + (xi[2] * xi[2] + xi[3] * xi[3]);
            continue;
        }
        *reinterpret_cast<double2*>(yp + j4) = make_double2(out[0], out[1]);
        *reinterpret_cast<double2*>(yp + j4 + 2) = make_double2(out[2], out[3]);
    }
    if (do_propose) {
        k0 = group_sum<32>(k0);
        if (lane == 0) { st.k0[r] = k0; st.epsrow[r] = eps; }
    }
    if (sp.diag) {
        rowsum = group_sum<32>(rowsum);
        const int nd = min(d, ND_MAX - 1) + 1;
        if (lane < nd) {
            const double f = (lane == nd - 1) ? rowsum / (double)d + st.mubar : y[lane] + st.mu[lane];
            st.S1[(int64_t)lane * K + r] += f;
            st.S2[(int64_t)lane * K + r] += f * f;
        }
    }
}

// Interior leapfrog step (hamiltonian.py:33-37, Nsteps > 1): with V' = Y' P of the trajectory point just
// evaluated,  p <- p + eps (-V'),  y' <- y' + eps p,  in place on the chain's proposal slot.
__global__ void __launch_bounds__(256)
leapfrog_mid_kernel(DenseState st) {
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 2;
    if (i >= st.K * st.dp) return;
    const int64_t r = i / st.dp;
    const int j = (int)(i % st.dp);
    const double eps = st.epsrow[r];
    const int64_t o = ((int64_t)(st.cur[r] ^ 1) * st.K + r) * st.dp + j;
    const double2 v = *reinterpret_cast<const double2*>(st.V + o);
    double2 p = *reinterpret_cast<const double2*>(st.Xi + i);
    double2 y = *reinterpret_cast<const double2*>(st.Y + o);
    p.x = p.x + eps * (-v.x); p.y = p.y + eps * (-v.y);
    y.x = y.x + eps * p.x;    y.y = y.y + eps * p.y;
    *reinterpret_cast<double2*>(st.Xi + i) = p;
    *reinterpret_cast<double2*>(st.Y + o) = y;
}

// HMC with a mass matrix, interior step: p <- p + eps g(q), g = -V' (hamiltonian.py:35); the position step is a GEMM
__global__ void __launch_bounds__(256)
mass_mid_kernel(DenseState st) {
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 2;
    if (i >= st.K * st.dp) return;
    const int64_t r = i / st.dp;
    const int j = (int)(i % st.dp);
    const double eps = st.epsrow[r];
    const double2 v = *reinterpret_cast<const double2*>(st.V + ((int64_t)(st.cur[r] ^ 1) * st.K + r) * st.dp + j);
    double2 p = *reinterpret_cast<double2*>(st.Pm + i);
    p.x -= eps * v.x; p.y -= eps * v.y;
    *reinterpret_cast<double2*>(st.Pm + i) = p;
}

// state i/o: theta[K][d] <-> centred slot-0/current rows; V recomputed by a GEMM on set
__global__ void dense_set_kernel(DenseState st, const double* __restrict__ theta) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= st.K * st.dp) return;
    const int64_t r = i / st.dp;
    const int j = (int)(i % st.dp);
    // the GEMM reads the PROPOSAL slot: stage the new state there, cur = 0 => slot 1
    st.Y[((int64_t)1 * st.K + r) * st.dp + j] = (j < st.d) ? theta[r * st.d + j] - st.mu[j] : 0.0;
    st.Xi[i] = 0.0;
    if (j == 0) { st.cur[r] = 0; st.k0[r] = 0.0; st.epsrow[r] = 0.0; }
}
__global__ void dense_adopt_kernel(DenseState st, double c1, double c2) {
    // after the GEMM on the staged rows: make slot 1 current and set lp from the partials
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= st.K) return;
    double q = 0.0;
    for (int b = lane; b < st.nblk; b += 32) q += st.partq[(int64_t)b * st.K + r];
    q = group_sum<32>(q);
    if (lane == 0) {
        st.cur[r] = 1;
        const double ll = -0.5 * ((q + c1) + c2);
        if (st.nt > 0) {
            st.ll[r] = ll;
            st.lp[r] = combine_logpost(0.0, ll * st.betas[r % st.nt]);   // chain_offset is a multiple of nt (set_tempering)
        } else {
            st.lp[r] = combine_logpost(0.0, ll);
        }
    }
}
__global__ void dense_get_kernel(DenseState st, double* theta, double* lp) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= st.K * st.d) return;
    const int64_t r = i / st.d;
    const int j = (int)(i % st.d);
    if (theta) theta[i] = st.Y[((int64_t)st.cur[r] * st.K + r) * st.dp + j] + st.mu[j];
    if (lp && j == 0) lp[r] = st.lp[r];
}
__global__ void dense_get_adapt_kernel(DenseState st, double* scale, int64_t* ns, int64_t* na) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= st.K) return;
    if (scale) scale[c] = st.scale[c];
    if (ns) ns[c] = st.nsamp[c];
    if (na) na[c] = st.nacc[c];
}


// ---------------------------------------------------------------------------------------------------------------------
// Pooled covariance adaptation (rmn_proposal_rw_set_pooled_cov_adapt).  Three small kernels that run every
// t_adapt steps: (1) S1 += sum_chains y, S2 += sum_chains y y^T over the chains' CURRENT centred states (lower triangle,
// 32 x 32 tiles, fp64 atomics -- the order of the atomics makes the last bits of S2 run-dependent; the proposal
// covariance is not parity-bound to anything); (2) optional all-reduce over ranks; (3) one block forms
// sd (S2 / n - m m^T) + jitter I and factors it (right-looking Cholesky on the transposed factor so that the column
// sweeps are coalesced), then overwrites the padded L the dense random walk multiplies its noise with.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int PM_T = 32;            // tile of the second-moment matrix
constexpr int PM_ROWS = 1024;       // chain rows per block
__global__ void __launch_bounds__(256)
pool_moments_kernel(DenseState st, double* __restrict__ S1, double* __restrict__ S2) {
    __shared__ double a[PM_T][PM_T + 1], b[PM_T][PM_T + 1];       // [row][col] of the two column blocks
    const int d = st.d, dp = st.dp;
    // lower-triangular tile pair (ti >= tj) from the linear block index
    int ti = 0, rem = blockIdx.x;
    while (rem > ti) { rem -= ti + 1; ++ti; }
    const int tj = rem;
    const int i0 = ti * PM_T, j0 = tj * PM_T;
    const int64_t r0 = (int64_t)blockIdx.y * PM_ROWS, r1 = min(st.K, r0 + PM_ROWS);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;       // 16 x 16 threads, 2 x 2 outputs each
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    double colsum = 0.0;                                          // S1 of column i0 + (tid & 31), by the tj == 0 blocks
    for (int64_t rb = r0; rb < r1; rb += PM_T) {
        for (int q = threadIdx.x; q < PM_T * PM_T; q += 256) {
            const int rr = q / PM_T, cc = q % PM_T;
            const int64_t r = rb + rr;
            double va = 0.0, vb = 0.0;
            if (r < r1) {
                const double* y = st.Y + ((int64_t)st.cur[r] * st.K + r) * dp;
                if (i0 + cc < d) va = y[i0 + cc];
                if (j0 + cc < d) vb = y[j0 + cc];
            }
            a[rr][cc] = va; b[rr][cc] = vb;
        }
        __syncthreads();
#pragma unroll 8
        for (int rr = 0; rr < PM_T; ++rr) {
            const double a0 = a[rr][2 * ty], a1 = a[rr][2 * ty + 1], b0 = b[rr][2 * tx], b1 = b[rr][2 * tx + 1];
            acc[0][0] += a0 * b0; acc[0][1] += a0 * b1; acc[1][0] += a1 * b0; acc[1][1] += a1 * b1;
        }
        if (tj == 0 && threadIdx.x < PM_T)
            for (int rr = 0; rr < PM_T; ++rr) colsum += a[rr][threadIdx.x];
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = i0 + 2 * ty + p, j = j0 + 2 * tx + q;
            if (i < d && j < d && j <= i) atomicAdd(&S2[(size_t)i * d + j], acc[p][q]);
        }
    if (tj == 0 && threadIdx.x < PM_T && i0 + threadIdx.x < d) atomicAdd(&S1[i0 + threadIdx.x], colsum);
}

// one block: C = sd (S2/n - m m^T) + jitter I (lower), U = chol(C)^T row-major (U[k][i] = L[i][k]); on success L -> Lpad
__global__ void __launch_bounds__(1024)
pool_chol_kernel(int d, int dp, const double* __restrict__ S1, const double* __restrict__ S2, double n, double sd,
                 double jitter, double* __restrict__ U, double* __restrict__ Lpad, int* __restrict__ status) {
    __shared__ double s_piv;
    __shared__ int s_bad;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) s_bad = 0;
    const double inv = 1.0 / n;
    for (int64_t q = tid; q < (int64_t)d * d; q += nt) {
        const int i = (int)(q / d), j = (int)(q % d);
        if (j <= i) {
            const double c = sd * (S2[q] * inv - (S1[i] * inv) * (S1[j] * inv)) + ((i == j) ? jitter : 0.0);
            U[(size_t)j * d + i] = c;                               // transposed: row j of U holds column j of C
        }
    }
    __syncthreads();
    for (int j = 0; j < d; ++j) {
        // column j of L: U[j][i] (i >= j) currently holds C[i][j]; subtract sum_k L[i][k] L[j][k] = sum_k U[k][i] U[k][j]
        for (int i = j + tid; i < d; i += nt) {
            double v = U[(size_t)j * d + i];
            for (int k = 0; k < j; ++k) v -= U[(size_t)k * d + i] * U[(size_t)k * d + j];
            U[(size_t)j * d + i] = v;
        }
        __syncthreads();
        if (tid == 0) {
            const double piv = U[(size_t)j * d + j];
            if (!(piv > 0.0) || !isfinite(piv)) s_bad = 1;
            s_piv = sqrt(piv);
        }
        __syncthreads();
        if (s_bad) break;
        const double rp = 1.0 / s_piv;
        for (int i = j + tid; i < d; i += nt) U[(size_t)j * d + i] = (i == j) ? s_piv : U[(size_t)j * d + i] * rp;
        __syncthreads();
    }
    if (s_bad) { if (tid == 0) atomicAdd(status, 1); return; }     // not positive definite: keep the previous factor
    for (int64_t q = tid; q < (int64_t)d * d; q += nt) {
        const int i = (int)(q / d), j = (int)(q % d);
        Lpad[(size_t)i * dp + j] = (j <= i) ? U[(size_t)j * d + i] : 0.0;
    }
}
__global__ void pool_cov_out_kernel(int d, const double* __restrict__ S1, const double* __restrict__ S2, double n,
                                    const double* __restrict__ mu, double* cov, double* mean, double* count) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const double inv = n > 0 ? 1.0 / n : 0.0;
    if (q < (int64_t)d * d) {
        const int i = (int)(q / d), j = (int)(q % d);
        const int hi = i > j ? i : j, lo = i > j ? j : i;
        cov[q] = S2[(size_t)hi * d + lo] * inv - (S1[i] * inv) * (S1[j] * inv);
    }
    if (q < d) mean[q] = S1[q] * inv + mu[q];
    if (q == 0) *count = n;
}

// ---------------------------------------------------------------------------------------------------------------------
// Exact (fp64) V = Y P and log-density of fp32 chain states, for the tensor-core precision mode (dense_tf32.cu): its start /
// periodic refresh pass as ONE DMMA GEMM instead of a SIMT row kernel (6.2 ms -> 1.6 ms at 16,384 chains x d = 1000).
// The fp32 states are widened into a scratch DenseState whose "proposal slot" is slot 0 (cur = 1 everywhere).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void exact_widen_kernel(int64_t K, int d, int dpf, int dp, const float* __restrict__ Yf, double* __restrict__ Yd,
                                   int* __restrict__ cur) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= K * dp) return;
    const int64_t r = i / dp;
    const int j = (int)(i % dp);
    Yd[i] = (j < d) ? (double)Yf[r * dpf + j] : 0.0;
    if (j == 0) cur[r] = 1;
}
__global__ void __launch_bounds__(256)
exact_narrow_kernel(DenseState st, int dpf, float* __restrict__ Vf, double* __restrict__ lp, double c1, double c2) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= st.K) return;
    double q = 0.0;
    for (int b = lane; b < st.nblk; b += 32) q += st.partq[(int64_t)b * st.K + r];
    q = group_sum<32>(q);
    if (lane == 0) lp[r] = combine_logpost(0.0, -0.5 * ((q + c1) + c2));
    const double* v = st.V + r * st.dp;
    for (int j = lane; j < dpf; j += 32) Vf[r * dpf + j] = (j < st.d) ? (float)v[j] : 0.0f;
}

}  // namespace

size_t dense_exact_scratch_bytes(int64_t K, int d) {
    const int dp = (d + 15) / 16 * 16, nblk = (dp + 31) / 32;
    return 2 * align256((size_t)K * dp * 8) + align256((size_t)nblk * K * 8) + align256((size_t)K * 4) + align256((size_t)dp * dp * 8);
}
// h_prec: the model's precision matrix [d][d] on the HOST (copied once into the scratch, zero padded)
int dense_exact_prepare(int64_t K, int d, const double* d_prec, void* scratch) {
    const int dp = (d + 15) / 16 * 16, nblk = (dp + 31) / 32;
    char* p = (char*)scratch + 2 * align256((size_t)K * dp * 8) + align256((size_t)nblk * K * 8) + align256((size_t)K * 4);
    RMN_CUDA(cudaMemset(p, 0, (size_t)dp * dp * 8));
    RMN_CUDA(cudaMemcpy2D(p, (size_t)dp * 8, d_prec, (size_t)d * 8, (size_t)d * 8, (size_t)d, cudaMemcpyDeviceToDevice));
    RMN_RAISE_SMEM(gemm_abt_kernel<EPI_LOGPOST_RW>, (int)GEMM_SMEM);
    return RMN_OK;
}
int dense_exact_pass(int64_t K, int d, int dpf, const float* Yf, float* Vf, double* lp, double c1, double c2, void* scratch,
                     cudaStream_t stream) {
    DenseState st{};
    st.K = K; st.d = d; st.dp = (d + 15) / 16 * 16; st.nblk = (st.dp + 31) / 32;
    char* p = (char*)scratch;
    st.Y = (double*)p; p += align256((size_t)K * st.dp * 8);
    st.V = (double*)p; p += align256((size_t)K * st.dp * 8);
    st.partq = (double*)p; p += align256((size_t)st.nblk * K * 8);
    st.cur = (int*)p; p += align256((size_t)K * 4);
    const double* Ppad = (const double*)p;
    const int64_t n = K * st.dp;
    exact_widen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(K, d, dpf, st.dp, Yf, st.Y, st.cur);
    RMN_KERNEL_CHECK();
    dim3 grid((st.dp + BN - 1) / BN, (unsigned)((K + BM - 1) / BM));
    gemm_abt_kernel<EPI_LOGPOST_RW><<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(st, Ppad);
    RMN_KERNEL_CHECK();
    exact_narrow_kernel<<<(unsigned)((K * 32 + 255) / 256), 256, 0, stream>>>(st, dpf, Vf, lp, c1, c2);
    RMN_KERNEL_CHECK();
    return RMN_OK;
}

namespace {

struct DenseGaussSampler : SamplerImpl {
    rmn_sampler* s;
    DenseState st{};
    double* d_Ppad = nullptr;     // [dp][dp] zero-padded precision
    double* d_Lpad = nullptr;     // [dp][dp] zero-padded chol(C0) (dense RW, pCN) or nullptr
    double* d_Linvpad = nullptr;  // [dp][dp] zero-padded chol(C0)^-1 (pCN)
    double* d_chM = nullptr; double* d_Minv = nullptr; double* d_chMinv = nullptr;   // [dp][dp] HMC mass matrix factors
    bool has_mass = false;
    double* d_Ldiag = nullptr;    // [dp]
    double* d_mupad = nullptr;
    bool rw_diag = true;
    bool pending = false;         // a proposal is in flight (written, GEMM done, not finished)
    // parallel tempering
    double* d_betas = nullptr;
    int set_tempering(int nt, const double* betas, double pswap) override {
        const rmn_proposal* pr = s->prop;
        RMN_REQUIRE(nt >= 2 && nt <= 32 && betas, "set_tempering: need 2 <= nt <= 32 temperatures");
        RMN_REQUIRE(pswap > 0.0 && pswap < 1.0, "Pswap must be a number between 0 and 1");
        RMN_REQUIRE(s->K % nt == 0, "set_tempering: the number of chains (%lld) must be a multiple of nt = %d", (long long)s->K, nt);
        RMN_REQUIRE(s->chain_offset % nt == 0, "set_tempering: chain_offset must be a multiple of nt");
        RMN_REQUIRE(!pr->adapt && !pr->pool_cov, "parallel tempering supports non-adaptive proposals only (the reference shares ONE "
                                                 "proposal object between all temperatures, ptsampler.py:81)");
        RMN_REQUIRE(pr->kind != RMN_PROP_HMC, "parallel tempering: the HMC gradient is not tempered in the reference; use RW or pCN");
        for (int i = 0; i < nt; ++i) RMN_REQUIRE(betas[i] >= 0.0 && betas[i] <= 1.0, "beta = %g must be a number between 0 and 1", betas[i]);
        if (!d_betas) RMN_CUDA(cudaMalloc(&d_betas, 32 * 8));
        RMN_CUDA(cudaMemcpy(d_betas, betas, (size_t)nt * 8, cudaMemcpyHostToDevice));
        if (!st.ll) {
            RMN_CUDA(cudaMalloc(&st.ll, (size_t)st.K * 8));
            RMN_CUDA(cudaMalloc(&st.role, (size_t)st.K));
            RMN_CUDA(cudaMemset(st.role, 0, (size_t)st.K));
        }
        st.betas = d_betas; st.nt = nt; st.pswap = pswap;
        return RMN_OK;
    }
    // pooled covariance adaptation
    double* d_pS1 = nullptr; double* d_pS2 = nullptr; double* d_pU = nullptr; int* d_pstatus = nullptr;
    double pool_n = 0.0;          // samples in the sums (all ranks)
    int64_t pool_updates = 0;
    RowComm poolc;                // the other ranks' chains join the pool through this communicator (optional)
    int set_row_comm(const void* id, size_t nbytes, int rank, int world) override {
        if (!s->prop->pool_cov) return unsupported("a communicator on the dense Gaussian sampler (pooled covariance adaptation only)");
        return rmn_rowcomm_init(&poolc, id, nbytes, rank, world);
    }
    explicit DenseGaussSampler(rmn_sampler* s_) : s(s_) {
        st.K = s->K; st.d = s->model->d; st.dp = (st.d + 15) / 16 * 16; st.nblk = (st.dp + 31) / 32;
        has_mass = s->prop->kind == RMN_PROP_HMC && s->prop->has_mass;
    }
    ~DenseGaussSampler() override {
        cudaFree(d_Ppad); cudaFree(d_Lpad); cudaFree(d_Linvpad); cudaFree(d_Ldiag); cudaFree(d_mupad);
        cudaFree(d_chM); cudaFree(d_Minv); cudaFree(d_chMinv);
        cudaFree(d_pS1); cudaFree(d_pS2); cudaFree(d_pU); cudaFree(d_pstatus); cudaFree(d_pS1g); cudaFree(d_pS2g);
        cudaFree(d_betas); cudaFree(st.ll); cudaFree(st.role);
        rmn_rowcomm_destroy(&poolc);
    }
    size_t row_bytes() const { return align256((size_t)st.K * st.dp * 8); }
    size_t workspace_bytes() const override {
        const size_t K = (size_t)st.K;
        return (has_mass ? 6 : 5) * row_bytes() + 2 * align256((size_t)st.nblk * K * 8) + 7 * align256(K * 8) +
               align256(K * 4) + 2 * align256(ND_MAX * K * 8) + 256;
    }
    int bind(void* ws) override {
        const size_t K = (size_t)st.K;
        char* p = (char*)ws;
        st.Y = (double*)p; p += 2 * row_bytes();
        st.V = (double*)p; p += 2 * row_bytes();
        st.Xi = (double*)p; p += row_bytes();
        st.partq = (double*)p; p += align256((size_t)st.nblk * K * 8);
        st.partk = (double*)p; p += align256((size_t)st.nblk * K * 8);
        st.lp = (double*)p; p += align256(K * 8);
        st.k0 = (double*)p; p += align256(K * 8);
        st.epsrow = (double*)p; p += align256(K * 8);
        st.scale = (double*)p; p += align256(K * 8);
        st.nsamp = (long long*)p; p += align256(K * 8);
        st.nacc = (long long*)p; p += align256(K * 8);
        st.dacc = (long long*)p; p += align256(K * 8);
        st.cur = (int*)p; p += align256(K * 4);
        st.S1 = (double*)p; p += align256(ND_MAX * K * 8);
        st.S2 = (double*)p; p += align256(ND_MAX * K * 8);
        if (has_mass) { st.Pm = (double*)p; p += row_bytes(); }
        // Y and V slots are contiguous pairs: slot b of array X is X + b*K*dp (row_bytes may pad)
        // -> keep the two slots exactly K*dp apart by laying them out inside one 2*row_bytes block
        RMN_CUDA(cudaMemset(ws, 0, workspace_bytes()));
        const int d = st.d, dp = st.dp;
        std::vector<double> hp((size_t)dp * dp, 0.0), hm(dp, 0.0);
        std::vector<double> tmp((size_t)d * d);
        RMN_CUDA(cudaMemcpy(tmp.data(), s->model->d_prec, (size_t)d * d * 8, cudaMemcpyDeviceToHost));
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) hp[(size_t)i * dp + j] = tmp[(size_t)i * d + j];
        for (int i = 0; i < d; ++i) hm[i] = s->model->h_mu[i];
        RMN_CUDA(cudaMalloc(&d_Ppad, hp.size() * 8));
        RMN_CUDA(cudaMemcpy(d_Ppad, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice));
        RMN_CUDA(cudaMalloc(&d_mupad, (size_t)dp * 8));
        RMN_CUDA(cudaMemcpy(d_mupad, hm.data(), (size_t)dp * 8, cudaMemcpyHostToDevice));
        st.mu = d_mupad;
        st.mubar = 0.0;
        for (int i = 0; i < d; ++i) st.mubar += hm[i] / (double)d;
        const rmn_proposal* pr = s->prop;
        if (pr->kind == RMN_PROP_RW) {
            rw_diag = true;
            for (int i = 0; i < d && rw_diag; ++i)
                for (int j = 0; j < i; ++j)
                    if (pr->h_L[(size_t)i * d + j] != 0.0) { rw_diag = false; break; }
            std::vector<double> ld(dp, 0.0);
            for (int i = 0; i < d; ++i) ld[i] = pr->h_L[(size_t)i * d + i];
            RMN_CUDA(cudaMalloc(&d_Ldiag, (size_t)dp * 8));
            RMN_CUDA(cudaMemcpy(d_Ldiag, ld.data(), (size_t)dp * 8, cudaMemcpyHostToDevice));
            if (pr->pool_cov) {
                rw_diag = false;                   // the adapted factor is dense whatever C0 was
                RMN_CUDA(cudaMalloc(&d_pS1, (size_t)d * 8));
                RMN_CUDA(cudaMalloc(&d_pS2, (size_t)d * d * 8));
                RMN_CUDA(cudaMalloc(&d_pU, (size_t)d * d * 8));
                RMN_CUDA(cudaMalloc(&d_pstatus, 4));
                RMN_CUDA(cudaMemset(d_pS1, 0, (size_t)d * 8));
                RMN_CUDA(cudaMemset(d_pS2, 0, (size_t)d * d * 8));
                RMN_CUDA(cudaMemset(d_pstatus, 0, 4));
            }
            if (!rw_diag) {
                std::vector<double> hl((size_t)dp * dp, 0.0);
                for (int i = 0; i < d; ++i)
                    for (int j = 0; j <= i; ++j) hl[(size_t)i * dp + j] = pr->h_L[(size_t)i * d + j];
                RMN_CUDA(cudaMalloc(&d_Lpad, hl.size() * 8));
                RMN_CUDA(cudaMemcpy(d_Lpad, hl.data(), hl.size() * 8, cudaMemcpyHostToDevice));
            }
        }
        if (has_mass) {
            auto up = [&](const std::vector<double>& full, double** dst, bool lower) -> int {
                std::vector<double> h((size_t)dp * dp, 0.0);
                for (int i = 0; i < d; ++i)
                    for (int j = 0; j < (lower ? i + 1 : d); ++j) h[(size_t)i * dp + j] = full[(size_t)i * d + j];
                RMN_CUDA(cudaMalloc(dst, h.size() * 8));
                RMN_CUDA(cudaMemcpy(*dst, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
                return RMN_OK;
            };
            if (int rc = up(pr->h_chM, &d_chM, true)) return rc;
            if (int rc = up(pr->h_Minv, &d_Minv, false)) return rc;
            if (int rc = up(pr->h_chMinv, &d_chMinv, true)) return rc;
            RMN_RAISE_SMEM(gemm_abt_kernel<EPI_MASS_P0>, (int)GEMM_SMEM);
            RMN_RAISE_SMEM(gemm_abt_kernel<EPI_MASS_STEP>, (int)GEMM_SMEM);
            RMN_RAISE_SMEM(gemm_abt_kernel<EPI_LOGPOST_MASS>, (int)GEMM_SMEM);
            RMN_RAISE_SMEM(gemm_abt_kernel<EPI_MASS_W1>, (int)GEMM_SMEM);
        }
        if (pr->kind == RMN_PROP_PCN) {
            rw_diag = false;
            st.rho = pr->rho; st.rho_c = sqrt(1.0 - pr->rho * pr->rho);
            std::vector<double> hl((size_t)dp * dp, 0.0), hi((size_t)dp * dp, 0.0);
            for (int i = 0; i < d; ++i)
                for (int j = 0; j <= i; ++j) {
                    hl[(size_t)i * dp + j] = pr->h_L[(size_t)i * d + j];
                    hi[(size_t)i * dp + j] = pr->h_Linv[(size_t)i * d + j];
                }
            RMN_CUDA(cudaMalloc(&d_Lpad, hl.size() * 8));
            RMN_CUDA(cudaMemcpy(d_Lpad, hl.data(), hl.size() * 8, cudaMemcpyHostToDevice));
            RMN_CUDA(cudaMalloc(&d_Linvpad, hi.size() * 8));
            RMN_CUDA(cudaMemcpy(d_Linvpad, hi.data(), hi.size() * 8, cudaMemcpyHostToDevice));
            RMN_RAISE_SMEM(gemm_abt_kernel<EPI_PCNPROP>, (int)GEMM_SMEM);
            RMN_RAISE_SMEM(gemm_abt_kernel<EPI_PCNREV>, (int)GEMM_SMEM);
        }
        RMN_RAISE_SMEM(gemm_abt_kernel<EPI_LOGPOST_RW>, (int)GEMM_SMEM);
        RMN_RAISE_SMEM(gemm_abt_kernel<EPI_LOGPOST_MALA>, (int)GEMM_SMEM);
        RMN_RAISE_SMEM(gemm_abt_kernel<EPI_RWPROP>, (int)GEMM_SMEM);
        int rc = rmn_fill_f64(st.scale, st.K, 1.0, 0);
        if (rc) return rc;
        RMN_CUDA(cudaDeviceSynchronize());
        return RMN_OK;
    }
    dim3 gemm_grid() const { return dim3((st.dp + BN - 1) / BN, (unsigned)((st.K + BM - 1) / BM)); }
    unsigned row_grid() const { return (unsigned)((st.K * 32 + 255) / 256); }
    double c1() const { return st.d * log(2.0 * M_PI); }

    template <int EPI>
    int gemm(const double* B, cudaStream_t stream) {
        ktimer.begin("gemm_abt_kernel", stream);
        gemm_abt_kernel<EPI><<<gemm_grid(), GEMM_THREADS, GEMM_SMEM, stream>>>(st, B);
        ktimer.end(stream);
        RMN_KERNEL_CHECK(); launches++;
        return RMN_OK;
    }

    int set_state(const double* d_theta, cudaStream_t stream) override {
        const int64_t n = st.K * st.dp;
        dense_set_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(st, d_theta);
        RMN_KERNEL_CHECK(); launches++;
        if (int rc = gemm<EPI_LOGPOST_RW>(d_Ppad, stream)) return rc;
        dense_adopt_kernel<<<row_grid(), 256, 0, stream>>>(st, c1(), s->model->logdetC);
        RMN_KERNEL_CHECK(); launches++;
        pending = false;
        return RMN_OK;
    }
    int get_state(double* d_theta, double* d_lp, cudaStream_t stream) override {
        const int64_t n = st.K * st.d;
        dense_get_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(st, d_theta, d_lp);
        RMN_KERNEL_CHECK(); launches++;
        return RMN_OK;
    }
    int run(int64_t T, const rmn_inject_t* inj, const rmn_trace_t* tr, cudaStream_t stream) override {
        const rmn_proposal* pr = s->prop;
        if (inj) RMN_REQUIRE(inj->d_xi && inj->d_u, "injected run needs d_xi and d_u");
        if (inj && st.nt > 0) RMN_REQUIRE(inj->d_usel, "injected tempered run needs d_usel (selection uniforms)");
        rmn_trace_t t0{};
        if (tr) t0 = *tr;
        if (t0.thin <= 0) t0.thin = 1;
        DenseStep sp{};
        sp.prop_kind = pr->kind; sp.adapt = pr->adapt; sp.rw_diag = rw_diag ? 1 : 0;
        sp.target = pr->target; sp.eps0 = pr->eps; sp.c1 = c1(); sp.c2 = s->model->logdetC;
        sp.Ldiag = d_Ldiag; sp.seed = s->seed; sp.chain_offset = s->chain_offset;
        sp.has_mass = has_mass ? 1 : 0;
        if (has_mass) sp.rw_diag = 0;                 // the propose pass only writes xi; the trajectory is a GEMM chain
        const int64_t K = st.K;
        const int d = st.d;
        // iteration t: [finish step t-1 | propose step t] ; GEMM(s).  A last call finishes step T-1.
        for (int64_t t = 0; t <= T; ++t) {
            sp.finish = (t > 0); sp.propose = (t < T); sp.diag = (t > 0); sp.record = sp.finish;
            sp.step_fin = step0 + t - 1; sp.step_prop = step0 + t;
            sp.inj_u = (inj && t > 0) ? inj->d_u + (t - 1) * K : nullptr;
            sp.inj_xi = (inj && t < T) ? inj->d_xi + t * K * d : nullptr;
            sp.inj_usel = (inj && st.nt > 0 && t < T) ? inj->d_usel + t * K : nullptr;
            sp.trace_slot = -1;
            sp.tr_theta = t0.d_theta; sp.tr_logpost = t0.d_logpost;
            sp.tr_prop_lp = (t0.d_prop_logpost && t > 0) ? t0.d_prop_logpost + (t - 1) * K : nullptr;
            sp.tr_acc = (t0.d_accepted && t > 0) ? t0.d_accepted + (t - 1) * K : nullptr;
            sp.tr_lqr = (t0.d_logqratio && t > 0) ? t0.d_logqratio + (t - 1) * K : nullptr;
            sp.tr_prop_theta = (t0.d_prop_theta && t > 0) ? t0.d_prop_theta + (t - 1) * K * d : nullptr;
            if (t > 0 && (t0.d_theta || t0.d_logpost)) {
                const int64_t i = t;                      // history index of the state after step t-1
                if (i >= t0.first && (i - t0.first) % t0.thin == 0) sp.trace_slot = (i - t0.first) / t0.thin;
            }
            const int64_t gstep = step0 + t;       // index of the step about to be proposed
            const bool adapt_now = pr->pool_cov && t < T && gstep > 0 && gstep % pr->pool_t_adapt == 0 &&
                                   (pr->pool_stop == 0 || gstep <= pr->pool_stop);
            if (st.nt > 0) {
                // tempering: a swap moves TWO rows, so the step is finished by one launch (the initiator's warp exchanges
                // both rows) and recorded / followed by the next proposal in a second one
                if (sp.finish) {
                    DenseStep fin = sp; fin.propose = 0; fin.record = 0; fin.diag = 0;
                    finish_propose_kernel<<<row_grid(), 256, 0, stream>>>(st, fin);
                    RMN_KERNEL_CHECK(); launches++;
                }
                DenseStep pro = sp; pro.finish = 0;
                pro.tr_prop_lp = nullptr; pro.tr_acc = nullptr; pro.tr_lqr = nullptr; pro.tr_prop_theta = nullptr;
                finish_propose_kernel<<<row_grid(), 256, 0, stream>>>(st, pro);
                RMN_KERNEL_CHECK(); launches++;
            } else if (adapt_now && sp.finish) {
                // the new factor must be in place before step gstep is proposed: finish step gstep-1 on its own first
                DenseStep fin = sp; fin.propose = 0;
                finish_propose_kernel<<<row_grid(), 256, 0, stream>>>(st, fin);
                RMN_KERNEL_CHECK(); launches++;
                if (int rc = pool_adapt(stream)) return rc;
                DenseStep pro = sp; pro.finish = 0; pro.diag = 0; pro.record = 0; pro.trace_slot = -1;
                pro.tr_prop_lp = nullptr; pro.tr_acc = nullptr; pro.tr_lqr = nullptr; pro.tr_prop_theta = nullptr;
                finish_propose_kernel<<<row_grid(), 256, 0, stream>>>(st, pro);
                RMN_KERNEL_CHECK(); launches++;
            } else {
                if (adapt_now) if (int rc = pool_adapt(stream)) return rc;
                finish_propose_kernel<<<row_grid(), 256, 0, stream>>>(st, sp);
                RMN_KERNEL_CHECK(); launches++;
            }
            if (t == T) break;
            if (pr->kind == RMN_PROP_RW && !rw_diag)
                if (int rc = gemm<EPI_RWPROP>(d_Lpad, stream)) return rc;
            if (pr->kind == RMN_PROP_PCN) {
                // theta' = rho theta + rho_c xi L^T (and D = theta - rho theta'), then |(rho_c L)^-1 D|^2
                if (int rc = gemm<EPI_PCNPROP>(d_Lpad, stream)) return rc;
                if (int rc = gemm<EPI_PCNREV>(d_Linvpad, stream)) return rc;
            }
            if (pr->kind == RMN_PROP_HMC && has_mass) {
                // leapfrog with a mass matrix (hamiltonian.py:13-52, 76-91): p0 = chM xi, positions move by
                // eps M^-1 p, and the kinetic energies use the whitened momenta chM^-1 p -- one GEMM each
                const int64_t n2 = st.K * st.dp / 2;
                if (int rc = gemm<EPI_MASS_P0>(d_chM, stream)) return rc;
                st.mass_base_prop = 0;
                if (int rc = gemm<EPI_MASS_STEP>(d_Minv, stream)) return rc;
                for (int l = 1; l < pr->nsteps; ++l) {
                    if (int rc = gemm<EPI_LOGPOST_RW>(d_Ppad, stream)) return rc;
                    mass_mid_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, stream>>>(st);
                    RMN_KERNEL_CHECK(); launches++;
                    st.mass_base_prop = 1;
                    if (int rc = gemm<EPI_MASS_STEP>(d_Minv, stream)) return rc;
                }
                if (int rc = gemm<EPI_LOGPOST_MASS>(d_Ppad, stream)) return rc;
                if (int rc = gemm<EPI_MASS_W1>(d_chMinv, stream)) return rc;
            } else if (pr->kind == RMN_PROP_HMC) {
                // Nsteps - 1 interior leapfrog steps, each one gradient GEMM + an in-place update
                for (int l = 1; l < pr->nsteps; ++l) {
                    if (int rc = gemm<EPI_LOGPOST_RW>(d_Ppad, stream)) return rc;
                    const int64_t n2 = st.K * st.dp / 2;
                    leapfrog_mid_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, stream>>>(st);
                    RMN_KERNEL_CHECK(); launches++;
                }
                if (int rc = gemm<EPI_LOGPOST_MALA>(d_Ppad, stream)) return rc;
            } else {
                if (int rc = gemm<EPI_LOGPOST_RW>(d_Ppad, stream)) return rc;
            }
        }
        step0 += T; diag_steps += T;
        return RMN_OK;
    }
    int pool_adapt(cudaStream_t stream) {
        const rmn_proposal* pr = s->prop;
        const int d = st.d;
        const int nt = (d + PM_T - 1) / PM_T;
        dim3 grid((unsigned)(nt * (nt + 1) / 2), (unsigned)((st.K + PM_ROWS - 1) / PM_ROWS));
        pool_moments_kernel<<<grid, 256, 0, stream>>>(st, d_pS1, d_pS2);
        RMN_KERNEL_CHECK(); launches++;
        pool_n += (double)st.K * (poolc.comm ? poolc.world : 1);
        const double* S1 = d_pS1; const double* S2 = d_pS2;
        if (poolc.comm) {
            // the sums are LOCAL running sums; the factor is built from their all-reduced copies (kept in d_pU's tail is not
            // possible -- U is the work matrix -- so reduce into scratch copies)
            if (!d_pS1g) { RMN_CUDA(cudaMalloc(&d_pS1g, (size_t)d * 8)); RMN_CUDA(cudaMalloc(&d_pS2g, (size_t)d * d * 8)); }
            RMN_CUDA(cudaMemcpyAsync(d_pS1g, d_pS1, (size_t)d * 8, cudaMemcpyDeviceToDevice, stream));
            RMN_CUDA(cudaMemcpyAsync(d_pS2g, d_pS2, (size_t)d * d * 8, cudaMemcpyDeviceToDevice, stream));
            double* bufs[2] = {d_pS1g, d_pS2g};
            const size_t counts[2] = {(size_t)d, (size_t)d * d};
            if (int rc = rmn_rowcomm_allreduce_f64(&poolc, bufs, counts, 2, stream)) return rc;
            S1 = d_pS1g; S2 = d_pS2g;
        }
        pool_chol_kernel<<<1, 1024, 0, stream>>>(d, st.dp, S1, S2, pool_n, pr->pool_sd, pr->pool_jitter, d_pU, d_Lpad, d_pstatus);
        RMN_KERNEL_CHECK(); launches++;
        pool_updates++;
        return RMN_OK;
    }
    double* d_pS1g = nullptr; double* d_pS2g = nullptr;
    int get_pooled_cov(double* d_cov, double* d_mean, double* d_count, cudaStream_t stream) override {
        if (!s->prop->pool_cov) return unsupported("get_pooled_cov (PooledAdaptCovRandomWalk only)");
        const int d = st.d;
        const double* S1 = d_pS1; const double* S2 = d_pS2;
        if (poolc.comm && d_pS1g && pool_updates > 0) { S1 = d_pS1g; S2 = d_pS2g; }   // the last all-reduced sums
        pool_cov_out_kernel<<<(unsigned)(((int64_t)d * d + 255) / 256), 256, 0, stream>>>(d, S1, S2, pool_n, st.mu, d_cov, d_mean, d_count);
        RMN_KERNEL_CHECK(); launches++;
        return RMN_OK;
    }
    int get_adapt(double* sc, int64_t* ns, int64_t* na, cudaStream_t stream) override {
        dense_get_adapt_kernel<<<(unsigned)((st.K + 127) / 128), 128, 0, stream>>>(st, sc, ns, na);
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

// one warp per point: v = P y, quad = y.v, grad = -v      (gaussian.py:49-58 via P = C^-1)
__global__ void __launch_bounds__(128)
gauss_point_kernel(int d, const double* __restrict__ mu, const double* __restrict__ prec, double c1,
                   double c2, int which, int64_t n, const double* __restrict__ theta,
                   double* __restrict__ out, double* __restrict__ grad) {
    const int64_t pt = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pt >= n) return;
    const double* th = theta + pt * d;
    double quad = 0.0;
    for (int i = 0; i < d; ++i) {
        double s = 0.0;
        for (int j = lane; j < d; j += 32) s += prec[(size_t)i * d + j] * (th[j] - mu[j]);
        s = group_sum<32>(s);
        if (grad && lane == 0) grad[pt * d + i] = -s;
        quad += s * (th[i] - mu[i]);
    }
    if (out && lane == 0) {
        const double ll = -0.5 * ((quad + c1) + c2);
        out[pt] = (which == 2) ? 0.0 : ((which == 1) ? ll : combine_logpost(0.0, ll));
    }
}

}  // namespace

SamplerImpl* make_dense_gauss_sampler(rmn_sampler* s) {
    const rmn_proposal* p = s->prop;
    (void)p;
    return new DenseGaussSampler(s);
}

int gauss_pointwise(rmn_model* m, int which, int64_t n, const double* d_theta, double* d_out,
                    double* d_grad, cudaStream_t st) {
    if (n <= 0) return RMN_OK;
    gauss_point_kernel<<<(unsigned)((n * 32 + 127) / 128), 128, 0, st>>>(
        m->d, m->d_mu, m->d_prec, m->d * log(2.0 * M_PI), m->logdetC, which, n, d_theta, d_out, d_grad);
    RMN_KERNEL_CHECK();
    return RMN_OK;
}
