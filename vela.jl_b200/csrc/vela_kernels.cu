// vela_kernels.cu — the CUDA kernels of the batched posterior hot path (sm_100a).
//
//   K0 sample_prologue   read_params scatter + per-sample hoists + TZR phase        (1 thread / walker)
//   K1 toa_chain         per (walker, TOA): component chain -> r, sigma^2; block-reduced
//                        chi^2 / log-det partials; optional emit of y, w           (1 thread / TOA,
//                        loops over a group of walkers whose parameter rows sit in shared memory)
//   K1b finish_white     sums the per-tile partials, assembles ln L / chi^2 (+ ln prior)
//   K3 sigma_contract    Sigma^-1 blocks = M^T diag(w) M (lower block triangle) and u = M^T (w.y)
//                        as an FP64 tensor-core (DMMA m8n8k4) contraction over the TOA axis
//   K4 chol_finish       + diag(Phi^-1); batched in-place Cholesky, log-det, forward solve,
//                        Woodbury ln L assembly (+ ln prior)
//
// Host-side orchestration and the C ABI are in vela_api.cu.
#include "vela_kernels.cuh"

#include <cstdio>

namespace vela {

// ------------------------------------------------------------------------------------------
// K0  sample prologue
// ------------------------------------------------------------------------------------------
// read_params (parameter.jl:158-169) + hoists + calc_tzr_phase (residuals.jl:29-32)
__global__ void __launch_bounds__(128) sample_prologue_kernel(const __grid_constant__ ProgramDev prog, PrologueArgs a) {
    prologue_body(prog, a);
}

// ------------------------------------------------------------------------------------------
// Kp  device-side priors: lnprior (prior.jl:13-19) and prior_transform (prior.jl:31-32)
// ------------------------------------------------------------------------------------------
__device__ inline double prior_logpdf(const PriorDev& pr, double x) {
    const double kLog2Pi = 1.8378770664093453;
    switch (pr.kind) {
        case VELA_PRIOR_UNIFORM: return (x >= pr.a && x <= pr.b) ? pr.norm : -CUDART_INF;
        case VELA_PRIOR_NORMAL: { double z = (x - pr.a) / pr.b; return -(z * z + kLog2Pi) / 2 - pr.norm; }
        case VELA_PRIOR_LOGNORMAL: {
            if (!(x > 0)) return -CUDART_INF;
            double lx = log(x), z = (lx - pr.a) / pr.b;
            return -(z * z + kLog2Pi) / 2 - pr.norm - lx;
        }
        case VELA_PRIOR_LOGUNIFORM: return (x >= pr.a && x <= pr.b) ? -log(x) - pr.norm : -CUDART_INF;
        case VELA_PRIOR_KIN: return (x >= 0 && x <= CUDART_PI) ? log(sin(x) / 2) : -CUDART_INF;
        case VELA_PRIOR_SINI: return (x >= 0 && x <= 1) ? log(x / sqrt(1 - x * x)) : -CUDART_INF;
        case VELA_PRIOR_STIGMA: { double t = 1 + x * x; return (x >= 0 && x <= 1) ? log(4 * x / (t * t)) : -CUDART_INF; }
        case VELA_PRIOR_SHAPMAX: return (x >= 0) ? log((1 - exp(-x)) / sqrt(2 * exp(x) - 1)) : -CUDART_INF;
        case VELA_PRIOR_PX: { double x2 = x * x; return (x >= 1 / pr.a) ? log(3 / (pr.a * pr.a * pr.a * x2 * x2)) : -CUDART_INF; }
        case VELA_PRIOR_LATITUDE: return (x >= -CUDART_PI / 2 && x <= CUDART_PI / 2) ? log(cos(x) / 2) : -CUDART_INF;
        case VELA_PRIOR_TRUNCATED_NORMAL: {
            if (!(x >= pr.lo && x <= pr.hi)) return -CUDART_INF;
            double z = (x - pr.a) / pr.b;
            return -(z * z + kLog2Pi) / 2 - pr.norm;
        }
    }
    return CUDART_NAN;
}

__device__ inline double prior_quantile(const PriorDev& pr, double q) {
    switch (pr.kind) {
        case VELA_PRIOR_UNIFORM: return pr.a + q * (pr.b - pr.a);
        case VELA_PRIOR_NORMAL: return pr.a + pr.b * normcdfinv(q);
        case VELA_PRIOR_LOGNORMAL: return exp(pr.a + pr.b * normcdfinv(q));
        case VELA_PRIOR_LOGUNIFORM: return exp(log(pr.a) + q * log(pr.b / pr.a));
        case VELA_PRIOR_KIN: return acos(1 - 2 * q);
        case VELA_PRIOR_SINI: return sqrt((2 - q) * q);
        case VELA_PRIOR_STIGMA: return q / (2 - q);                  // as in known_priors.jl:61
        case VELA_PRIOR_SHAPMAX: return log((sqrt(2 * q - q * q) + 1) / ((q - 1) * (q - 1)));
        case VELA_PRIOR_PX: return 1 / (pr.a * cbrt(1 - q));
        case VELA_PRIOR_LATITUDE: return asin(2 * q - 1);
        case VELA_PRIOR_TRUNCATED_NORMAL: {
            double c0 = normcdf((pr.lo - pr.a) / pr.b), c1 = normcdf((pr.hi - pr.a) / pr.b);
            return pr.a + pr.b * normcdfinv(c0 + q * (c1 - c0));
        }
    }
    return CUDART_NAN;
}

// mode 0: out[b] = sum_k logpdf_k(params[b][k]);  mode 1: out[b][k] = quantile_k(params[b][k])
__global__ void prior_kernel(const PriorDev* __restrict__ priors, int nd, const double* __restrict__ params, long B,
                             int mode, double* __restrict__ out) {
    long b = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (mode == 0) {
        double lp = 0.0;
        for (int k = 0; k < nd; ++k) lp += prior_logpdf(priors[k], params[b * nd + k]);
        out[b] = lp;
    } else {
        for (int k = 0; k < nd; ++k) out[b * nd + k] = prior_quantile(priors[k], params[b * nd + k]);
    }
}

// ------------------------------------------------------------------------------------------
// K1  per-(walker, TOA) chain
// ------------------------------------------------------------------------------------------
// _wls_lnlike_term (wls_likelihood.jl:7-14, wideband_wls_likelihood.jl:1-21) and
// _calc_resids_and_Ninvdiag (gls_likelihood.jl:9-79).
//
// grid = (TOA tiles, walker groups); one thread owns one TOA (its record lives in registers for the
// whole kernel) and loops over the walkers of the group, whose extended parameter rows are staged in
// shared memory once per block (uniform-address LDS = broadcast).  For each walker the block reduces
// chi^2 and log-det terms over its TOAs with warp shuffles; per-tile partials go to `partial`
// [tile][walker][2] and are summed in a fixed order later (deterministic, no atomics).
__global__ void __launch_bounds__(kChainThreads, 2) toa_chain_kernel(const __grid_constant__ ProgramDev prog, ChainArgs a) {
    chain_body(prog, a);
}

// create(): one thread per TOA evaluates the chain at the default parameters (P: their extended row) and fills the C_DELTA
// columns; thread 0 also stores the TZR TOA's spin phase (default_phase_point, vela_chain.cuh)
__global__ void default_phase_kernel(const __grid_constant__ ProgramDev prog, const double* P, double* toa, long stride, long n_toas,
                                     const double* tzr_block, const uint32_t* index_masks, const uint8_t* bit_masks, double* tzr0) {
    const long j = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (j >= n_toas) return;
    double out3[3], z[2];
    default_phase_point(prog, P, toa, stride, j, tzr_block, index_masks, bit_masks, n_toas, out3, z);
#pragma unroll
    for (int c = 0; c < 3; ++c) toa[(size_t)(C_DELTA + c) * stride + j] = out3[c];
    if (j == 0) { tzr0[0] = z[0]; tzr0[1] = z[1]; }
}

// ------------------------------------------------------------------------------------------
// K1b  white-noise finish: -sum/2 (wls_likelihood.jl:49-56) or chi^2 (wls_chi2.jl:44-51), + prior
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double combine_prior(double lnl, const double* lnprior, long b, int with_prior) {
    if (isnan(lnl)) lnl = -CUDART_INF;
    if (!with_prior) return lnl;
    double lp = lnprior ? lnprior[b] : 0.0;
    return isfinite(lp) ? lp + lnl : lp;  // posterior.jl:9-12
}

__global__ void finish_white_kernel(const double* __restrict__ partial, int n_tiles, long B, int mode,
                                    const double* __restrict__ lnprior, int with_prior, double* __restrict__ out) {
    long b = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (b >= B) return;
    double chi = 0.0, lg = 0.0;
    for (int t = 0; t < n_tiles; ++t) {
        chi += partial[((size_t)t * B + b) * 2];
        lg += partial[((size_t)t * B + b) * 2 + 1];
    }
    if (mode == 1) { out[b] = chi; return; }
    out[b] = combine_prior(-(chi + lg) / 2, lnprior, b, with_prior);
}

// ------------------------------------------------------------------------------------------
// K2  ECORR group corrections (EcorrKernel)
// ------------------------------------------------------------------------------------------
// _ecorr_lnlike_group / _ecorr_chi2_group (ecorr_likelihood.jl:8-38, ecorr_chi2.jl:6-30): with r_u = sum_g y w,
// u_u = sum_g w over the TOAs of an ECORR group and W = ECORR^2, the block-diagonal covariance adds
//   chi^2   -= W r_u^2 / (1 + W u_u),      log det += log(1 + W u_u)
// to the white-noise sums the chain kernel already produced.  One warp = one walker x kEcorrChunksPerRow consecutive
// chunks (a chunk = 32 consecutive groups): the lanes read the walker's w / y rows in coalesced 64-TOA passes
// into a per-warp shared buffer, lane g then sums the TOAs of group g of the chunk; the group terms are reduced over
// the warp (log(1 + W u_u) as a mantissa product and an exponent sum: one logarithm per row) and written as one
// extra row of `partial`, so the finishing kernels need no change.  (The first version, one thread per walker
// walking its row alone, read w / y with a stride of one row per lane: 0.6 ms at 10k TOAs x 8192 walkers.)
__global__ void __launch_bounds__(256) ecorr_group_kernel(EcorrArgs a) {
    __shared__ double sY[8][64], sW[8][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long b = (long)blockIdx.x * 8 + warp;
    if (b >= a.B) return;                        // whole warps leave; the kernel uses no block barrier
    const double* P = a.Pext + (size_t)b * a.row;
    const double* y = a.Y + (size_t)b * a.ld_yw;
    const double* w = a.W + (size_t)b * a.ld_yw;
    const bool y_is_wy = a.y_is_wy != 0;
    double dchi = 0.0, lman = 1.0;
    int lex = 0;
    const int c_end = min((int)(blockIdx.y + 1) * kEcorrChunksPerRow, a.n_chunks);
    for (int c = blockIdx.y * kEcorrChunksPerRow; c < c_end; ++c) {
        const int g0 = a.chunk_start[c], g1 = a.chunk_start[c + 1];      // g1 - g0 <= 32 (create())
        int3 grp = make_int3(0x7fffffff, 0, 0);
        if (g0 + lane < g1) grp = a.groups[g0 + lane];   // first TOA (0-based), one-past-last TOA, ECORR index (0 = none)
        if (grp.z == 0) grp = make_int3(0x7fffffff, 0, 0);
        int t0 = grp.x, t1 = grp.y;                      // TOA span of the chunk's active groups
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            t0 = min(t0, __shfl_xor_sync(0xffffffffu, t0, off));
            t1 = max(t1, __shfl_xor_sync(0xffffffffu, t1, off));
        }
        double r_u = 0.0, u_u = 0.0;
        for (int p0 = t0; p0 < t1; p0 += 64) {
            __syncwarp();
#pragma unroll
            for (int i = lane; i < 64; i += 32)
                if (p0 + i < t1) { sY[warp][i] = y[p0 + i]; sW[warp][i] = w[p0 + i]; }
            __syncwarp();
            const int lo = max(grp.x, p0), hi = min(grp.y, p0 + 64);
            for (int j = lo; j < hi; ++j) {
                const double wj = sW[warp][j - p0];
                r_u += y_is_wy ? sY[warp][j - p0] : sY[warp][j - p0] * wj;
                u_u += wj;
            }
        }
        if (grp.z) {
            const double ecorr = P[a.par_ecorr + grp.z - 1];
            const double wt = ecorr * ecorr;
            const double denom = 1 + wt * u_u;
            dchi -= wt * r_u * r_u / denom;
            int ex;
            lman *= split_mantissa(denom, &ex);          // <= kEcorrChunksPerRow mantissas in [1/2, 1) per lane
            lex += ex;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        dchi += __shfl_xor_sync(0xffffffffu, dchi, off);
        lman *= __shfl_xor_sync(0xffffffffu, lman, off);
        lex += __shfl_xor_sync(0xffffffffu, lex, off);
    }
    if (lane == 0) {
        double* out = a.partial + ((size_t)(a.first_row + blockIdx.y) * a.B + b) * 2;
        out[0] = dchi;
        out[1] = log(lman) + lex * 0.6931471805599453;   // ln 2
    }
}

__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// ---- WoodburyKernel{EcorrKernel}: per-group Sherman-Morrison (_calc_Ninv_M, gls_ecorr_likelihood.jl:66-137).
// With P_g = sum_{i in g} w_i M_i (a p-vector), Q_g = sum w_i, r_g = sum w_i y_i and alpha_g = W/(1 + W Q_g):
//   M^T Nc^-1 M = M^T N^-1 M - sum_g alpha_g P_g P_g^T,      M^T Nc^-1 y = M^T N^-1 y - sum_g alpha_g r_g P_g.
// K2b builds V_b[g] = sqrt(alpha_g) [P_g | r_g] for every walker b; K2c subtracts V_b^T V_b (a per-walker SYRK on the
// FP64 tensor cores) from the Sigma / u that K3 produced for the white part.
//
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// K2b: block = 32 walkers x 8 column slices.  One staging pass brings the TOAs of SEVERAL consecutive ECORR groups
// (up to `chunk` TOAs: M rows, w, y) into shared memory, then every thread walks the groups of the pass with its
// accumulators (NACC = ceil(columns / 8)) in registers and writes one V row per group: two block barriers per ~8
// epochs instead of two per epoch, no padded staging.  A single group longer than the chunk is processed in pieces.
template <int NACC>
__global__ void __launch_bounds__(256, NACC <= 8 ? 4 : (NACC <= 16 ? 2 : 1)) ecorr_basis_kernel(EcorrBasisArgs a) {
    extern __shared__ double smem[];
    const int CH = a.chunk;                       // TOAs per staging pass: a power of two (64, or less when the rank is large)
    const int lgCH = 31 - __clz(CH);
    // smem row pitch (odd -> conflict-free).  Up to rank 128 the rows are padded with zeros to the 8 * NACC columns the
    // accumulators cover, so the inner loop carries no column predicate.
    constexpr bool PADDED = NACC <= 16;
    const int p = a.rank, pc = PADDED ? 8 * NACC + 1 : p + 1;
    double* sM = smem;                            // [CH TOAs][pc]
    double* sW = sM + CH * pc;                    // [32 walkers][CH + 1]
    double* sY = sW + 32 * (CH + 1);
    const int tid = threadIdx.x, s = tid >> 3, c = tid & 7;
    const long b0 = (long)blockIdx.y * 32;
    const long b = b0 + s;
    const bool live = b < a.B;
    const int g_begin = blockIdx.x * a.groups_per_block, g_end = min(g_begin + a.groups_per_block, a.n_groups);

    auto stage = [&](int t0, int nj) {            // TOAs [t0, t0 + nj) -> shared memory
        __syncthreads();
        // asynchronous 8-byte copies: every load of the pass is in flight at once and none occupies a register
        for (int e = tid; e < p * CH; e += 256) {
            const int col = e >> lgCH, i = e & (CH - 1);
            if (i < nj) cp_async8(&sM[i * pc + col], &a.M[(size_t)col * a.ldM + t0 + i]);
        }
        for (int e = tid; e < 32 * CH; e += 256) {
            const int ss = e >> lgCH, i = e & (CH - 1);
            if (i < nj) {
                const long br = min(b0 + ss, a.B - 1);    // rows past the batch repeat the last walker (never stored)
                cp_async8(&sW[ss * (CH + 1) + i], &a.W[(size_t)br * a.ld_yw + t0 + i]);
                cp_async8(&sY[ss * (CH + 1) + i], &a.Y[(size_t)br * a.ld_yw + t0 + i]);
            }
        }
        cp_async_commit_wait_all();
        __syncthreads();
    };
    if (PADDED)                                   // the pad columns stay zero: staging writes columns < rank only
        for (int e = tid; e < CH * pc; e += 256) sM[e] = 0.0;
    double acc[NACC];
    double Q, ru;
    auto reset = [&]() {
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
        Q = 0.0; ru = 0.0;
    };
    const bool y_is_wy = a.y_is_wy != 0;
    auto accumulate = [&](int i0, int i1) {        // staged TOAs [i0, i1)
        const double* wrow = sW + s * (CH + 1);
        const double* yrow = sY + s * (CH + 1);
        const double* mrow = sM + (size_t)i0 * pc + c;
        for (int i = i0; i < i1; ++i, mrow += pc) {
            const double w = wrow[i];
            Q += w;
            ru += y_is_wy ? yrow[i] : w * yrow[i];
#pragma unroll
            for (int k = 0; k < NACC; ++k)        // explicit fma: the file is built with -fmad=false
                if (PADDED || c + 8 * k < p) acc[k] = fma(mrow[8 * k], w, acc[k]);
        }
    };
    double dchi = 0.0, lman = 1.0;                 // scalar corrections of this block's groups (slice 0 only)
    int lex = 0;
    auto flush = [&](int g, int ecorr_index) {
        if (!live) return;
        const double ecorr = a.Pext[(size_t)b * a.row + a.par_ecorr + ecorr_index - 1];
        const double wt = ecorr * ecorr;
        const double denom = fma(wt, Q, 1.0);
        const double sa = fabs(ecorr) * rsqrt(denom);   // sqrt(wt / denom)
        if (c == 0 && a.partial) {                 // ecorr_likelihood.jl:8-76 terms of this group
            const double sr = sa * ru;             // wt ru^2 / denom without a second division
            dchi = fma(-sr, sr, dchi);
            int ex;                                // log(denom) summed as mantissa product + exponent sum: one
            lman *= split_mantissa(denom, &ex);    // logarithm per block (<= 32 groups: the product stays >= 2^-32)
            lex += ex;
        }
        double* v = a.V + ((size_t)b * a.g_pad + g) * a.ldv;
#pragma unroll
        for (int k = 0; k < NACC; ++k) {
            const int col = c + 8 * k;
            if (col < p) v[col] = sa * acc[k];
        }
        if (c == 0) v[p] = sa * ru;
    };

    int g = g_begin;
    while (g < g_end) {
        const int3 first = a.groups[g];            // [first, last) TOA, ECORR index >= 1
        if (first.y - first.x > CH) {              // one long group: pieces of CH TOAs
            reset();
            for (int j0 = first.x; j0 < first.y; j0 += CH) {
                const int nj = min(CH, first.y - j0);
                stage(j0, nj);
                accumulate(0, nj);
            }
            flush(g, first.z);
            ++g;
            continue;
        }
        int ge = g;                                // last group of this pass: the span first.x .. groups[ge].y fits
        while (ge + 1 < g_end && a.groups[ge + 1].y - first.x <= CH) ++ge;
        stage(first.x, a.groups[ge].y - first.x);
        for (int gg = g; gg <= ge; ++gg) {
            const int3 grp = a.groups[gg];
            reset();
            accumulate(grp.x - first.x, grp.y - first.x);
            flush(gg, grp.z);
        }
        g = ge + 1;
    }
    if (c == 0 && live && a.partial) {
        double* out = a.partial + ((size_t)(a.first_row + blockIdx.x) * a.B + b) * 2;
        out[0] = dchi;
        out[1] = log(lman) + lex * 0.6931471805599453;   // ln 2
    }
}
template __global__ void ecorr_basis_kernel<8>(EcorrBasisArgs);
template __global__ void ecorr_basis_kernel<16>(EcorrBasisArgs);
template __global__ void ecorr_basis_kernel<40>(EcorrBasisArgs);
template __global__ void ecorr_basis_kernel<128>(EcorrBasisArgs);

// K2c: C_b = V_b^T V_b over the ECORR groups, one warp = one walker x one 32 x 32 rectangle (4 x 4 m8n8k4 tiles) of
// the lower block triangle; fragments come straight from global/L2 (V_b rows are 64-byte segments per lane group).
// Row p of V (the r_g column) yields the correction to u.  The result is subtracted from slab 0 of Sigma / U.
__global__ void __launch_bounds__(256) ecorr_syrk_kernel(EcorrSyrkArgs a) {
    const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
    const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= a.B * a.n_rect) return;
    const long b = wid / a.n_rect;
    int rect = (int)(wid - b * a.n_rect);
    int ri = 0;
    while ((ri + 1) * (ri + 2) / 2 <= rect) ++ri;
    const int cj = rect - ri * (ri + 1) / 2;
    const double* V = a.V + (size_t)b * a.g_pad * a.ldv;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const double* pa = V + (size_t)tig * a.ldv + ri * 32 + gid;
    const double* pb = V + (size_t)tig * a.ldv + cj * 32 + gid;
    for (int g0 = 0; g0 < a.g_pad; g0 += 4) {
        double fa[4], fb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            fa[i] = pa[i * 8];
            fb[i] = pb[i * 8];
        }
        pa += 4 * a.ldv;
        pb += 4 * a.ldv;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
    }
    const int p = a.rank;
    double* sig = a.Sig + (size_t)b * a.ld_sig;
    double* u = a.U + (size_t)b * a.ld_u;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int pp = (ri * 4 + i) * 8 + gid, qq = (cj * 4 + j) * 8 + 2 * tig + r;
                if (pp < p && qq <= pp) sig[(size_t)pp * (pp + 1) / 2 + qq] -= acc[i][j][r];
                else if (pp == p && qq < p) u[qq] -= acc[i][j][r];
            }
}

// K2c for ranks up to 63: one warp = one walker holds the whole lower triangle of the 64 x 64 product (36 m8n8k4
// tiles = 72 accumulators), so V is read exactly once and no diagonal rectangle computes its upper half; the A
// fragment of tile row t and the B fragment of tile column t are the same registers, and the fragments of the next
// four groups are fetched before the 36 DMMAs of the current four are issued.
__global__ void __launch_bounds__(128) ecorr_syrk64_kernel(EcorrSyrkArgs a) {
    const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
    const long b = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= a.B) return;
    const double* V = a.V + (size_t)b * a.g_pad * a.ldv + (size_t)tig * a.ldv + gid;
    double acc[36][2];
#pragma unroll
    for (int t = 0; t < 36; ++t) acc[t][0] = acc[t][1] = 0.0;
    // fragments travel two steps ahead of the DMMAs (eight warps per SM: one step does not cover the HBM latency)
    double f[8], f1[8], f2[8];
    const long last = a.g_pad - 4;
    const double* v1 = V + (size_t)(4 < a.g_pad ? 4 : last) * a.ldv;
#pragma unroll
    for (int t = 0; t < 8; ++t) { f[t] = V[t * 8]; f1[t] = v1[t * 8]; }
    for (long g0 = 0; g0 < a.g_pad; g0 += 4) {
        const double* nxt = V + (size_t)(g0 + 8 < a.g_pad ? g0 + 8 : last) * a.ldv;
#pragma unroll
        for (int t = 0; t < 8; ++t) f2[t] = nxt[t * 8];
#pragma unroll
        for (int tr = 0; tr < 8; ++tr)
#pragma unroll
            for (int tc = 0; tc <= tr; ++tc) dmma8x8x4(acc[tr * (tr + 1) / 2 + tc][0], acc[tr * (tr + 1) / 2 + tc][1], f[tr], f[tc]);
#pragma unroll
        for (int t = 0; t < 8; ++t) { f[t] = f1[t]; f1[t] = f2[t]; }
    }
    const int p = a.rank;
    double* sig = a.Sig + (size_t)b * a.ld_sig;
    double* u = a.U + (size_t)b * a.ld_u;
#pragma unroll
    for (int tr = 0; tr < 8; ++tr)
#pragma unroll
        for (int tc = 0; tc <= tr; ++tc)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int pp = tr * 8 + gid, qq = tc * 8 + 2 * tig + r;
                if (pp < p && qq <= pp) sig[(size_t)pp * (pp + 1) / 2 + qq] -= acc[tr * (tr + 1) / 2 + tc][r];
                else if (pp == p && qq < p) u[qq] -= acc[tr * (tr + 1) / 2 + tc][r];
            }
}

// ------------------------------------------------------------------------------------------
// K3  Sigma^-1 contraction on the FP64 tensor cores
// ------------------------------------------------------------------------------------------
// _calc_Ninv_M + _calc_Sigmainv__and__MT_Ninv_y (gls_likelihood.jl:101-199), restructured as one
// GEMM over the TOA axis:   C[t(p,q), b] = sum_j (M[j,p] M[j,q]) * w[b,j],   u[p, b] = sum_j M[j,p] (w y)[b,j].
//
// GEMM rows are the PACKED lower triangle t = p(p+1)/2 + q (q <= p < rank): no padding of the rank and
// no redundant upper/diagonal-block work.  Rows are cut into m8 tiles; a warp owns a "unit" of MT=4
// consecutive m-tiles x NT=4 n-tiles (32 walkers): 16 m8n8k4 accumulators = 32 FP64 registers, and each
// k4 step issues 16 independent DMMAs for 12 LDS.64 + 4 DMUL.  The A operand is never materialised: every
// A element is the product of two entries of the shared-memory M tile (row gid of m-tile i <-> its (p,q)).
// "u" units use M rows directly against the w.y product.  A CTA is 16 warps = 16 consecutive units over
// the same 32 walkers (one CTA per SM, 4 warps per scheduler); all `rank` columns of M plus the W and Y
// rows of the 32 walkers are staged with cp.async (LDGSTS) in a STAGES-deep ring, K-contiguous with a
// (BK+4)-double row pitch (conflict-free fragment reads).  Stage hand-off uses mbarriers instead of
// __syncthreads: copies complete on a per-stage "full" barrier (cp.async.mbarrier.arrive.noinc), readers
// release a stage through its "empty" barrier, and the refill runs STAGES-2 tiles ahead - one tile behind
// the slowest reader - so warps drift freely.
// Measured alternatives (B200, config 3, ms per launch): 8x8-block tiling 16.3;
// packed rows, 8x4 warp tile x 8 warps: __syncthreads ring 11.32, mbarrier ring 11.10, 1-D bulk copies
// (cp.async.bulk; UBLKCP is a uniform-datapath instruction, so per-lane rows serialise) 13.96;
// 4x4 warp tile: 8 warps x 2 CTAs/SM 10.61, 16 warps x 1 CTA/SM (this kernel) 10.50.

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// mbarrier helpers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// ---- K3, decoupled cp.async pipeline: LDGSTS data movement as in sigma_contract_kernel, but completion is
// tracked with per-stage mbarriers (cp.async.mbarrier.arrive.noinc) and stages are released through "empty"
// mbarriers, so there is no block-wide barrier in the main loop; the refill runs S-2 tiles ahead, one tile
// behind the slowest reader, so neither wait normally blocks.
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

template <int BK, int STAGES, int MT, int NT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (MT * NT * WARPS <= 128) ? 2 : 1) sigma_contract_kernel(SigmaArgs a) {
    constexpr int kThreads = WARPS * 32;
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) unsigned long long full_bar[STAGES], empty_bar[STAGES];
    constexpr int PITCH = BK + 4;
    constexpr int NS = NT * 8;
    constexpr int AHEAD = STAGES > 2 ? STAGES - 2 : 1;   // prefetch distance in tiles
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const long b0 = (long)blockIdx.y * NS;
    const int P = a.rank;
    const int stage_rows = P + 2 * NS;
    const size_t stage_doubles = (size_t)stage_rows * PITCH;
    constexpr int kChunks = BK / 2;
    const int total_chunks = stage_rows * kChunks;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], kThreads); mbar_init(&empty_bar[s], WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    // split-K: this CTA covers k-tiles [kt0, kt0 + n_ktiles) and writes its own partial Sigma slab
    const int kt0 = blockIdx.z * a.kt_per_split;
    const int n_ktiles = min(a.kt_per_split, (int)(a.n_rows_pad / BK) - kt0);
    auto issue = [&](int t) {   // t: local tile index
        const int s = t % STAGES;
        double* base = smem + (size_t)s * stage_doubles;
        const long j0 = (long)(kt0 + t) * BK;
        for (int c = tid; c < total_chunks; c += kThreads) {
            int r = c / kChunks, ch = c - r * kChunks;
            const double* src;
            if (r < P) src = a.M + (size_t)r * a.ldM + j0 + ch * 2;
            else if (r < P + NS) src = a.W + (size_t)(b0 + (r - P)) * a.ld_yw + j0 + ch * 2;
            else src = a.Y + (size_t)(b0 + (r - P - NS)) * a.ld_yw + j0 + ch * 2;
            cp_async16(base + (size_t)r * PITCH + ch * 2, src);
        }
        cp_async_mbar_arrive(&full_bar[s]);
    };

    const int u = blockIdx.x * WARPS + warp;
    const bool have_unit = u < a.n_units;
    const bool is_pair = u < a.n_pair_units;
    const int tile0 = is_pair ? u * MT : (u - a.n_pair_units) * MT;
    int offp[MT], offq[MT];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        int p = 0, q = 0;
        if (have_unit) {
            long t = (long)(tile0 + i) * 8 + gid;
            if (is_pair) {
                if (t < a.n_pairs) {
                    p = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
                    while ((long)(p + 1) * (p + 2) / 2 <= t) ++p;
                    while ((long)p * (p + 1) / 2 > t) --p;
                    q = (int)(t - (long)p * (p + 1) / 2);
                }
            } else if (t < P) {
                p = (int)t;
            }
        }
        offp[i] = p * PITCH;
        offq[i] = q * PITCH;
    }

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[i][n][0] = acc[i][n][1] = 0.0;

    for (int t = 0; t < AHEAD && t < n_ktiles; ++t) issue(t);

    for (int kt = 0; kt < n_ktiles; ++kt) {
        const int nk = kt + AHEAD;
        if (nk < n_ktiles) {
            if (nk >= STAGES) mbar_wait(&empty_bar[nk % STAGES], ((nk / STAGES) - 1) & 1);
            issue(nk);
        }
        const int s = kt % STAGES;
        mbar_wait(&full_bar[s], (kt / STAGES) & 1);
        if (have_unit) {
            const double* sM = smem + (size_t)s * stage_doubles;
            const double* sW = sM + (size_t)P * PITCH;
            const double* sY = sW + (size_t)NS * PITCH;
#pragma unroll
            for (int kk = 0; kk < BK / 4; ++kk) {
                const int k = kk * 4 + tig;
                double bw[NT], af[MT];
#pragma unroll
                for (int n = 0; n < NT; ++n) bw[n] = sW[(n * 8 + gid) * PITCH + k];
                if (is_pair) {
#pragma unroll
                    for (int i = 0; i < MT; ++i) af[i] = sM[offp[i] + k] * sM[offq[i] + k];
                } else {
#pragma unroll
                    for (int n = 0; n < NT; ++n) bw[n] *= sY[(n * 8 + gid) * PITCH + k];
#pragma unroll
                    for (int i = 0; i < MT; ++i) af[i] = sM[offp[i] + k];
                }
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int n = 0; n < NT; ++n) dmma8x8x4(acc[i][n][0], acc[i][n][1], af[i], bw[n]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

    if (!have_unit) return;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const long t = (long)(tile0 + i) * 8 + gid;
        const bool ok = is_pair ? (t < a.n_pairs) : (t < P);
        if (!ok) continue;
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                long b = b0 + n * 8 + 2 * tig + r;
                if (b < a.B) {
                    if (is_pair) a.Sig[(size_t)blockIdx.z * a.split_stride_sig + (size_t)b * a.ld_sig + t] = acc[i][n][r];
                    else a.U[(size_t)blockIdx.z * a.split_stride_u + (size_t)b * a.ld_u + t] = acc[i][n][r];
                }
            }
    }
}

#define VELA_SIGMA_INST(BK, ST, MT, NT, W) template __global__ void sigma_contract_kernel<BK, ST, MT, NT, W>(SigmaArgs);
// 4x4 warp tiles, 16 warps sharing one ring, one CTA per SM
VELA_SIGMA_INST(64, 3, 4, 4, 16)
VELA_SIGMA_INST(32, 4, 4, 4, 16)
VELA_SIGMA_INST(32, 3, 4, 4, 16)
VELA_SIGMA_INST(16, 4, 4, 4, 16)
VELA_SIGMA_INST(16, 3, 4, 4, 16)
VELA_SIGMA_INST(16, 2, 4, 4, 16)
VELA_SIGMA_INST(8, 2, 4, 4, 16)
// 3x4 warp tiles: same CTA shape, used when 24-row units quantise the packed rows better (plan_sigma)
VELA_SIGMA_INST(64, 3, 3, 4, 16)
VELA_SIGMA_INST(32, 4, 3, 4, 16)
VELA_SIGMA_INST(32, 3, 3, 4, 16)
VELA_SIGMA_INST(16, 4, 3, 4, 16)
VELA_SIGMA_INST(16, 3, 3, 4, 16)
VELA_SIGMA_INST(16, 2, 3, 4, 16)
VELA_SIGMA_INST(8, 2, 3, 4, 16)

// ------------------------------------------------------------------------------------------
// K3s  column sums for the structured Sigma path
// ------------------------------------------------------------------------------------------
// When the basis columns are harmonics of one phase per row (Fourier GPs: M[j, sin_k] = s_j sin(k x_j), ...), every
// product M[j,p] M[j,q] is half the sum / difference of two entries of a small trigonometric table
// (sin a sin b = (cos(a-b) - cos(a+b)) / 2, ...), so Sigma^-1's p(p+1)/2 contractions over the TOAs collapse to
// the 2 (K_a + K_c) + 1 sums  C(m) = sum_j w_j s_j cos(m x_j),  S(m) = sum_j w_j s_j sin(m x_j)  per pair of
// column groups; products with non-harmonic columns are stored as explicit columns.  The host builds that table
// T at create() (vela_api.cu: analyse_basis); this kernel is the remaining GEMM  R = T^T w  (and u = M^T (w y),
// whose columns are appended to T), on the same DMMA tiling and mbarrier ring as the dense kernel: a warp owns
// 32 columns x 32 walkers, a CTA is UC column units x WC walker tiles (16 warps) and stages only its own
// 32 UC columns of T.
// SPLIT = true: the chain kernel stored w.y in the Y buffer (ChainArgs::emit == 3), and every CTA column is
// homogeneous - either 32 UC table columns against W or 32 UC columns of M against w.y - so a stage holds one walker
// operand instead of two and the two kinds of units pad separately.
template <int BK, int STAGES, int UC, int WC, bool SPLIT>
__global__ void __launch_bounds__(UC * WC * 32, (UC * WC <= 8) ? 2 : 1) colsum_kernel(ColsumArgs a) {
    constexpr int WARPS = UC * WC, kThreads = WARPS * 32, MT = 4, NT = 4;
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) unsigned long long full_bar[STAGES], empty_bar[STAGES];
    constexpr int PITCH = BK + 4;
    constexpr int NCOL = UC * 32, NS = WC * 32;
    constexpr int AHEAD = STAGES > 2 ? STAGES - 2 : 1;
    constexpr int stage_rows = NCOL + (SPLIT ? NS : 2 * NS);
    constexpr size_t stage_doubles = (size_t)stage_rows * PITCH;
    constexpr int kChunks = BK / 2;
    constexpr int total_chunks = stage_rows * kChunks;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int uc = warp % UC, wc = warp / UC;
    const long b0 = (long)blockIdx.y * NS;
    const int n_units = a.n_cols / 32;
    // first unit of this CTA column and whether its units multiply w or w.y
    int u_first, u_end;
    bool use_y;
    if (SPLIT) {
        const int n_t_ctas = (a.n_w_units + UC - 1) / UC;
        use_y = (int)blockIdx.x >= n_t_ctas;
        u_first = use_y ? a.n_w_units + ((int)blockIdx.x - n_t_ctas) * UC : (int)blockIdx.x * UC;
        u_end = use_y ? n_units : a.n_w_units;
    } else {
        u_first = blockIdx.x * UC;
        u_end = n_units;
        use_y = false;
    }
    int c0 = u_first * 32, col_clamp = a.n_cols - 1, out_col0 = 0, n_valid = a.n_valid;
    const double* Tbase = a.T;
    const double* walker_src = (SPLIT && use_y) ? a.Y : a.W;
    if (a.units) {   // stage-2 job table: this CTA column's own matrix, moment plane and output run
        const ColsumUnit cu = a.units[blockIdx.x];
        u_first = 0; u_end = (cu.n_valid + 31) / 32; c0 = 0; col_clamp = u_end * 32 - 1;
        Tbase = a.T + cu.t_off; walker_src = a.W + (size_t)cu.w_off * a.w_plane; out_col0 = cu.r_col; n_valid = cu.n_valid;
    }

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], kThreads); mbar_init(&empty_bar[s], WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const int kt0 = blockIdx.z * a.kt_per_split;
    const int n_ktiles = min(a.kt_per_split, (int)(a.n_rows_pad / BK) - kt0);
    auto issue = [&](int t) {
        const int s = t % STAGES;
        double* base = smem + (size_t)s * stage_doubles;
        const long j0 = (long)(kt0 + t) * BK;
        for (int c = tid; c < total_chunks; c += kThreads) {
            int r = c / kChunks, ch = c - r * kChunks;
            const double* src;
            if (r < NCOL) src = Tbase + (size_t)min(c0 + r, col_clamp) * a.ldT + j0 + ch * 2;
            else if (r < NCOL + NS) src = walker_src + (size_t)min(b0 + (r - NCOL), a.B - 1) * a.ld_yw + j0 + ch * 2;
            else src = a.Y + (size_t)min(b0 + (r - NCOL - NS), a.B - 1) * a.ld_yw + j0 + ch * 2;
            cp_async16(base + (size_t)r * PITCH + ch * 2, src);
        }
        cp_async_mbar_arrive(&full_bar[s]);
    };

    const int u = u_first + uc;
    const bool have_unit = u < u_end;
    const bool mul_y = !SPLIT && u >= a.n_w_units;

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[i][n][0] = acc[i][n][1] = 0.0;

    for (int t = 0; t < AHEAD && t < n_ktiles; ++t) issue(t);

    for (int kt = 0; kt < n_ktiles; ++kt) {
        const int nk = kt + AHEAD;
        if (nk < n_ktiles) {
            if (nk >= STAGES) mbar_wait(&empty_bar[nk % STAGES], ((nk / STAGES) - 1) & 1);
            issue(nk);
        }
        const int s = kt % STAGES;
        mbar_wait(&full_bar[s], (kt / STAGES) & 1);
        if (have_unit) {
            const double* sT = smem + (size_t)s * stage_doubles + (size_t)(uc * 32) * PITCH;
            const double* sW = smem + (size_t)s * stage_doubles + (size_t)(NCOL + wc * 32) * PITCH;
            const double* sY = sW + (size_t)NS * PITCH;
#pragma unroll
            for (int kk = 0; kk < BK / 4; ++kk) {
                const int k = kk * 4 + tig;
                double bw[NT], af[MT];
#pragma unroll
                for (int n = 0; n < NT; ++n) bw[n] = sW[(n * 8 + gid) * PITCH + k];
                if (mul_y) {
#pragma unroll
                    for (int n = 0; n < NT; ++n) bw[n] *= sY[(n * 8 + gid) * PITCH + k];
                }
#pragma unroll
                for (int i = 0; i < MT; ++i) af[i] = sT[(i * 8 + gid) * PITCH + k];
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int n = 0; n < NT; ++n) dmma8x8x4(acc[i][n][0], acc[i][n][1], af[i], bw[n]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

    if (!have_unit) return;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int col = out_col0 + u * 32 + i * 8 + gid;
        if (n_valid && col - out_col0 >= n_valid) continue;
        // normal mode: one slab per split-K slice; moment mode: walker-major planes per 32-column block, slice = cell
        double* dst = a.plane_stride ? a.R + (size_t)u * a.plane_stride + (size_t)blockIdx.z * 32 + (i * 8 + gid)
                                     : a.R + (size_t)blockIdx.z * a.split_stride_r + col;
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                long b = b0 + wc * 32 + n * 8 + 2 * tig + r;
                if (b < a.B) dst[(size_t)b * a.ld_r] = acc[i][n][r];
            }
    }
}

template __global__ void colsum_kernel<16, 2, 16, 1, false>(ColsumArgs);
template __global__ void colsum_kernel<16, 3, 8, 2, false>(ColsumArgs);
template __global__ void colsum_kernel<16, 3, 4, 4, false>(ColsumArgs);
template __global__ void colsum_kernel<16, 2, 2, 8, false>(ColsumArgs);
template __global__ void colsum_kernel<8, 4, 16, 1, true>(ColsumArgs);
template __global__ void colsum_kernel<16, 4, 8, 2, true>(ColsumArgs);
template __global__ void colsum_kernel<16, 4, 4, 4, true>(ColsumArgs);
template __global__ void colsum_kernel<16, 4, 2, 8, true>(ColsumArgs);
template __global__ void colsum_kernel<32, 2, 2, 8, true>(ColsumArgs);
template __global__ void colsum_kernel<16, 2, 1, 16, true>(ColsumArgs);   // one column unit x 512 walkers: stage 1 of the two-stage sums
template __global__ void colsum_kernel<8, 4, 1, 16, true>(ColsumArgs);    // stage-1 variants (VELA_B200_STAGE1_SHAPE): deeper ring of 8-row tiles,
template __global__ void colsum_kernel<16, 2, 1, 8, true>(ColsumArgs);    // 256-walker CTAs, two resident per SM,
template __global__ void colsum_kernel<16, 4, 1, 8, true>(ColsumArgs);    // 256-walker CTAs with a four-stage ring

// ------------------------------------------------------------------------------------------
// K4  + Phi^-1, Cholesky, log-det, solve, ln L
// ------------------------------------------------------------------------------------------
// calc_noise_weights_inv (gls_likelihood.jl:3-7, gp_noise.jl:12-16,64-78, marginalized_timing_model.jl:24)
// and _gls_lnlike_serial (gls_likelihood.jl:201-226).  One CTA per walker.  The lower triangle of
// Sigma^-1 (row p, col q<=p) is factorised in place as L L^T (right-looking, one column per step),
// in shared memory when it fits, else directly in the L2-resident global buffer.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += scratch[w];
    return s;
}

// Sum of one entry over the split-K slabs.  Small batches run the column sums with up to 64 slabs and a single CTA then
// adds ~16k values: four accumulators and an unrolled body keep 16 L2 loads in flight per thread (the plain loop had 4).
__device__ __forceinline__ double slab_sum(const double* g, long stride, int n) {
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    int z = 0;
#pragma unroll 4
    for (; z + 4 <= n; z += 4) {
        v0 += g[(size_t)z * stride];
        v1 += g[(size_t)(z + 1) * stride];
        v2 += g[(size_t)(z + 2) * stride];
        v3 += g[(size_t)(z + 3) * stride];
    }
    for (; z < n; ++z) v0 += g[(size_t)z * stride];
    return (v0 + v1) + (v2 + v3);
}

// R[i] <- sum over the split-K slabs of R[z * stride + i], in slice order (deterministic); one thread per entry of the
// first slab.  Used when a small batch split K deeply (see enqueue_batch).
__global__ void slab_reduce_kernel(double* R, long stride, int n_slabs, long n) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) R[i] = slab_sum(R + i, stride, n_slabs);
}

// Row-packed lower triangle: entry (r, c <= r) of the walker's matrix (rows 0..p, row p = u^T and -z.z)
__device__ __forceinline__ size_t tri(int r, int c) { return (size_t)r * (r + 1) / 2 + c; }

// matrix entry (r, c <= r): row-packed triangle or pitched rows
#define AIDX(r, c) (PACKED ? tri((r), (c)) : (size_t)(r) * lda + (c))

// Cholesky of one NB x NB (NB <= 8) diagonal block by one warp, in registers: lane r owns row r (statically indexed
// after unrolling), columns travel by shuffle.  Rows >= nb are identity padding.  ~20 registers, no shared-memory
// round trips inside the dependent chain.
template <int NB, bool PACKED>
__device__ __forceinline__ void factor_diag_block_warp(double* A, int lda, int k0, int nb, double* dblk, double* rdiag, int* fail) {
    constexpr int PP = NB + 1;
    const int lane = threadIdx.x & 31;
    double row[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c)
        row[c] = (lane < nb && c <= lane) ? A[AIDX(k0 + lane, k0 + c)] : ((lane == c) ? 1.0 : 0.0);
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        double d = __shfl_sync(0xffffffffu, row[c], c);          // pivot, held by lane c
        if (!(d > 0.0) || !isfinite(d)) { if (lane == 0) *fail = 1; d = 1.0; }
        const double rd = rsqrt(d);
        if (lane == c) { row[c] = d * rd; rdiag[c] = rd; }
        else if (lane > c) row[c] *= rd;
#pragma unroll
        for (int q = c + 1; q < NB; ++q) {
            const double lqc = __shfl_sync(0xffffffffu, row[c], q);   // L[q][c]
            if (lane >= q) row[q] = fma(-row[c], lqc, row[q]);
        }
    }
    if (lane < NB) {
#pragma unroll
        for (int c = 0; c < NB; ++c)
            if (c <= lane) {
                dblk[lane * PP + c] = row[c];
                if (lane < nb) A[AIDX(k0 + lane, k0 + c)] = row[c];
            }
    }
}

// Blocked right-looking factorisation (block width NB): per block column (1) the NB x NB diagonal block is
// factorised by one warp (in registers for NB = 8, in a small shared tile for NB = 16), (2) one thread per row
// solves the panel below it against that block, (3) all threads apply the rank-NB update to the trailing lower
// triangle in 2x2 register micro-tiles fed from a shared copy of the panel.  u is carried as an extra matrix row p:
// the panel solve turns it into z = L^-1 u and the trailing update leaves -z.z in entry (p, p), so there is no
// separate forward solve.  Three block barriers per block of 8 columns instead of five per column.
// The dot products are written with fma(): this file is compiled with -fmad=false for the double-double chain, and
// the factorisation has no reference rounding order to mirror (LAPACK potrf uses fused multiply-adds as well).
template <int NB, bool PACKED>
__global__ void __launch_bounds__(NB <= 8 ? kCholLaunchThreads : kCholThreads, NB <= 8 ? 6 : 2) chol_finish_kernel(CholArgs a) {
    extern __shared__ double smem[];
    __shared__ double scratch[kCholThreads / 32];
    __shared__ int sFail;
    constexpr int PP = NB + 1;                    // pitch of the panel / diagonal tiles
    const long b = blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int p = a.p, pp = a.pp;

    double* phi = smem;                           // [pp]
    double* z = phi + pp;                         // [pp]   (conditional-mean back substitution only)
    double* dblk = z + pp;                        // [NB][PP] diagonal block
    double* rdiag = dblk + NB * PP;               // [NB]   1 / L_cc of the current block
    double* panel = rdiag + NB;                   // [p + 1][PP] rows below the diagonal block
    // matrix rows 0..p (row p = u^T): row-packed lower triangle in shared memory (PACKED), pitched rows in the
    // L2-resident global scratch of the large ranks
    const int lda = pp + 1;
    const size_t n_tri = PACKED ? (size_t)(p + 1) * (p + 2) / 2 : (size_t)(p + 1) * lda;
    double* A = PACKED ? panel + (size_t)(p + 1) * PP : a.Sq + (size_t)b * n_tri;
    // ---- Sigma: unpack the packed lower triangle t = r(r+1)/2 + c (dense path) or assemble it from the column sums
    //      through the pair table (structured path); split-K slabs are summed in slice order
    {
        const double* gA = a.Sig + (size_t)b * a.ld_sig;
        const long n_pairs = (long)p * (p + 1) / 2;
        double* sR = nullptr;
        if (a.table) {
            sR = panel + (size_t)(p + 1) * PP + (PACKED ? n_tri : 0);   // last shared slot
            const double* gR = a.R + (size_t)b * a.ld_r;
            for (int i = tid; i < a.n_r; i += nt) sR[i] = slab_sum(gR + i, a.split_stride_r, a.ksplit);
            __syncthreads();
        }
        if (PACKED && a.table) {
            // the packed triangle is indexed by the pair number itself; four table entries in flight per thread
#pragma unroll 4
            for (long t = tid; t < n_pairs; t += nt) {
                const PairTerm pt = a.table[t];
                double v = pt.c1 * sR[pt.i1];
                if (pt.c2 != 0.0) v = fma(pt.c2, sR[pt.i2], v);
                if (a.add_sig) v += gA[t];
                A[t] = v;
            }
        } else {
            int r = 0, c = tid;
            while (c > r) { c -= r + 1; ++r; }
            for (long t = tid; t < n_pairs; t += nt) {
                double v;
                if (a.table) {
                    const PairTerm pt = a.table[t];
                    v = pt.c1 * sR[pt.i1];
                    if (pt.c2 != 0.0) v = fma(pt.c2, sR[pt.i2], v);
                    if (a.add_sig) v += gA[t];
                } else {
                    v = slab_sum(gA + t, a.split_stride_sig, a.ksplit);
                }
                A[AIDX(r, c)] = v;
                c += nt;
                while (c > r) { c -= r + 1; ++r; }
            }
        }
        const double* gU = a.table ? a.R + a.u_off : a.U;
        const long ldu = a.table ? a.ld_r : a.ld_u, su = a.table ? a.split_stride_r : a.split_stride_u;
        for (int i = tid; i < p; i += nt) A[AIDX(p, i)] = slab_sum(gU + (size_t)b * ldu + i, su, a.ksplit);
    }
    if (tid == 0) { sFail = 0; A[AIDX(p, p)] = 0.0; }

    // ---- Phi^-1 for this walker
    const double* P = a.Pext + (size_t)b * a.row;
    for (int g = 0; g < a.n_gp; ++g) {
        GPDev gp = a.gp[g];
        if (gp.kind == VELA_GP_MARGINALIZED_TM) {
            for (int i = tid; i < gp.n; i += nt) phi[gp.col0 + i] = a.gp_data[gp.data0 + i];
        } else {
            // powerlaw(A, gamma, f1, f1) gp_noise.jl:12-16 with A = exp10(log10_A)
            const double fyr = 1.0 / 3600 / 24 / 365.25;
            const double denom = 12 * (CUDART_PI * CUDART_PI) * fyr * fyr * fyr;
            double amp = exp10(P[gp.par_log10_amp]), gamma = P[gp.par_gamma], f1 = P[gp.par_f1];
            double P1 = amp * amp / denom * pow(fyr / f1, gamma) * f1;
            for (int i = tid; i < gp.n; i += nt) {
                double v = 1 / (P1 * exp(-gamma * a.gp_data[gp.data0 + i]));
                phi[gp.col0 + i] = v;
                phi[gp.col0 + gp.n + i] = v;
            }
        }
    }
    __syncthreads();
    double lphi = 0.0;
    for (int i = tid; i < p; i += nt) {
        lphi += log(phi[i]);
        A[AIDX(i, i)] += phi[i];
    }
    const double logdetPhi = -block_sum(lphi, scratch);
    __syncthreads();

    // ---- Cholesky A = L L^T (isposdef + cholesky!, gls_likelihood.jl:211-217) and z = L^-1 u (ldiv!, :220)
    const int n1 = p + 1;                         // matrix rows including the u row
    // NB <= 8: look-ahead.  Warp 0 factorises the NEXT diagonal block during the trailing update (it first brings that
    // block up to date itself) while the other warps update the rest, so nobody waits for the serial factorisation and
    // a block step has two barriers instead of three; only the first block is factorised up front.
    if constexpr (NB <= 8) {
        if (tid < 32 && p > 0) factor_diag_block_warp<NB, PACKED>(A, lda, 0, min(NB, p), dblk, rdiag, &sFail);
        __syncthreads();
    }
    for (int k0 = 0; k0 < p; k0 += NB) {
        const int nb = min(NB, p - k0), k1 = k0 + nb;
        // (1) diagonal block
        if constexpr (NB > 8) {
            for (int e = tid; e < nb * nb; e += nt) {
                const int r = e / nb, c = e - r * nb;
                if (c <= r) dblk[r * PP + c] = A[AIDX(k0 + r, k0 + c)];
            }
            __syncthreads();
            if (tid < 32) {
                for (int c = 0; c < nb; ++c) {
                    if (tid == 0) {
                        double d = dblk[c * PP + c];
                        if (!(d > 0.0) || !isfinite(d)) { sFail = 1; d = 1.0; }
                        const double rd = rsqrt(d);
                        dblk[c * PP + c] = d * rd;
                        rdiag[c] = rd;
                    }
                    __syncwarp();
                    const double rd = rdiag[c];
                    for (int r = c + 1 + tid; r < nb; r += 32) dblk[r * PP + c] *= rd;
                    __syncwarp();
                    const int m = nb - c - 1;     // trailing block of the tile: entries (r, q), c < q <= r < nb
                    for (int e = tid; e < m * m; e += 32) {
                        const int r = c + 1 + e / m, q = c + 1 + e % m;
                        if (q <= r) dblk[r * PP + q] = fma(-dblk[r * PP + c], dblk[q * PP + c], dblk[r * PP + q]);
                    }
                    __syncwarp();
                }
                for (int e = tid; e < nb * nb; e += 32) {
                    const int r = e / nb, c = e - r * nb;
                    if (c <= r) A[AIDX(k0 + r, k0 + c)] = dblk[r * PP + c];
                }
            }
            __syncthreads();
        }
        // (2) panel: rows k1 .. p (row p is u^T) times L_kk^-T, one thread per row
        for (int i = k1 + tid; i < n1; i += nt) {
            double x[NB];
            double* row = A + AIDX(i, k0);
#pragma unroll
            for (int c = 0; c < NB; ++c) x[c] = c < nb ? row[c] : 0.0;
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                if (c < nb) {
                    double v = x[c];
#pragma unroll
                    for (int q = 0; q < c; ++q) v = fma(-x[q], dblk[c * PP + q], v);
                    x[c] = v * rdiag[c];
                }
            }
            double* prow = panel + (size_t)(i - k1) * PP;
#pragma unroll
            for (int c = 0; c < NB; ++c)
                if (c < nb) { row[c] = x[c]; prow[c] = x[c]; }
                else prow[c] = 0.0;
        }
        __syncthreads();
        // (3) trailing update A[i][j] -= sum_c L[i][k0+c] L[j][k0+c] for k1 <= j <= i <= p, 2x2 micro-tiles
        const int nbn = NB <= 8 ? max(0, min(NB, p - k1)) : 0;   // rows of the next diagonal block (look-ahead)
        if (NB <= 8 && tid < 32) {
            for (int e = tid; e < nbn * (nbn + 1) / 2; e += 32) {
                int i = 0;
                while ((i + 1) * (i + 2) / 2 <= e) ++i;
                const int j = e - i * (i + 1) / 2;
                double sacc = 0.0;
#pragma unroll
                for (int c = 0; c < NB; ++c) sacc = fma(panel[(size_t)i * PP + c], panel[(size_t)j * PP + c], sacc);
                A[AIDX(k1 + i, k1 + j)] -= sacc;
            }
            __syncwarp();
            if (nbn > 0) factor_diag_block_warp<NB, PACKED>(A, lda, k1, nbn, dblk, rdiag, &sFail);
        } else {
            const int wt = NB <= 8 ? tid - 32 : tid, wn = NB <= 8 ? nt - 32 : nt;   // worker index / count
            const int tx = wt & 15, ty = wt >> 4, ny = wn >> 4;
            const int m = n1 - k1;                // trailing size (rows k1..p)
            for (int i0 = nbn + ty; i0 < m; i0 += 2 * ny) {
                const int i1 = i0 + ny;
                double li0[NB], li1[NB];
#pragma unroll
                for (int c = 0; c < NB; ++c) {
                    li0[c] = panel[(size_t)i0 * PP + c];
                    li1[c] = i1 < m ? panel[(size_t)i1 * PP + c] : 0.0;
                }
                for (int j0 = tx; j0 <= min(i1, m - 1); j0 += 32) {
                    const int j1 = j0 + 16;
                    // the four matrix entries are fetched before the dot products so that their latency (L2 for the
                    // global-scratch variant) overlaps the FMAs
                    const bool w00 = j0 <= i0, w01 = j1 <= i0 && j1 < m, w10 = i1 < m && j0 <= i1, w11 = i1 < m && j1 <= i1 && j1 < m;
                    double* q00 = A + AIDX(k1 + i0, k1 + j0);
                    double* q01 = A + AIDX(k1 + i0, k1 + j1);
                    double* q10 = A + AIDX(k1 + i1, k1 + j0);
                    double* q11 = A + AIDX(k1 + i1, k1 + j1);
                    const double a00 = w00 ? *q00 : 0.0, a01 = w01 ? *q01 : 0.0, a10 = w10 ? *q10 : 0.0, a11 = w11 ? *q11 : 0.0;
                    double s00 = 0.0, s01 = 0.0, s10 = 0.0, s11 = 0.0;
#pragma unroll
                    for (int c = 0; c < NB; ++c) {
                        const double lj0 = panel[(size_t)j0 * PP + c];
                        const double lj1 = j1 < m ? panel[(size_t)j1 * PP + c] : 0.0;
                        s00 = fma(li0[c], lj0, s00); s01 = fma(li0[c], lj1, s01);
                        s10 = fma(li1[c], lj0, s10); s11 = fma(li1[c], lj1, s11);
                    }
                    if (w00) *q00 = a00 - s00;
                    if (w01) *q01 = a01 - s01;
                    if (w10) *q10 = a10 - s10;
                    if (w11) *q11 = a11 - s11;
                }
            }
        }
        __syncthreads();
    }

    double ld = 0.0;
    for (int i = tid; i < p; i += nt) ld += log(A[AIDX(i, i)]);
    const double logdetSig = 2 * block_sum(ld, scratch);
    // y^T N^-1 y and log det N: partial rows of the chain tiles (and ECORR blocks), summed by the whole block
    double py = 0.0, pl = 0.0;
    for (int t = tid; t < a.n_tiles; t += nt) {
        const double2 v = *reinterpret_cast<const double2*>(a.partial + ((size_t)t * a.B + b) * 2);
        py += v.x;
        pl += v.y;
    }
    const double yNy = block_sum(py, scratch), logdetN = block_sum(pl, scratch);
    if (tid == 0) {
        const double zz = -A[AIDX(p, p)];
        double lnl = -0.5 * (yNy - zz + logdetN + logdetPhi + logdetSig);
        if (sFail) lnl = -CUDART_INF;
        a.out[b] = combine_prior(lnl, a.lnprior, b, a.with_prior);
    }
    if (!a.ahat) return;
    // ---- conditional mean of the marginalised amplitudes: L^T ahat = z, i.e. ahat = Sigma u
    // (pyvela/pyvela/spnta.py:376-394: cho_solve((Sigmainv_cf, False), MT_Ninv_y)) and the upper factor L^T
    for (int i = tid; i < p; i += nt) z[i] = A[AIDX(p, i)];
    for (int k = p - 1; k >= 0; --k) {
        __syncthreads();
        if (tid == 0) z[k] /= A[AIDX(k, k)];
        __syncthreads();
        const double zk = z[k];
        for (int i = tid; i < k; i += nt) z[i] = fma(-A[AIDX(k, i)], zk, z[i]);
    }
    __syncthreads();
    const double bad = sFail ? CUDART_NAN : 0.0;
    for (int i = tid; i < p; i += nt) a.ahat[(size_t)b * p + i] = z[i] + bad;
    if (a.cf) {
        double* dst = a.cf + (size_t)b * p * p;
        for (int idx = tid; idx < p * p; idx += nt) {
            const int i = idx / p, j = idx - i * p;
            dst[idx] = j >= i ? A[AIDX(j, i)] + bad : 0.0;
        }
    }
}
#undef AIDX
template __global__ void chol_finish_kernel<8, true>(CholArgs);
template __global__ void chol_finish_kernel<16, true>(CholArgs);
template __global__ void chol_finish_kernel<16, false>(CholArgs);

// ------------------------------------------------------------------------------------------
// K4p  Cholesky finish for large ranks: 32-column panels in shared memory, trailing updates on the FP64 tensor cores
// ------------------------------------------------------------------------------------------
// Ranks whose triangle does not fit shared memory (p >~ 220; the 100k-TOA array pulsar has p = 300) keep the matrix
// (rows 0..p, row p = u^T, pitched rows) in an L2-resident global scratch.  chol_finish_kernel<16, false> updates that
// scratch with 2 x 2 register tiles 16 columns at a time: 16 FMAs per read-modify-write of the triangle, latency-bound on
// L2 (ncu: FP64 pipe 26 % active, 8.7 GB of DRAM traffic per 2048 walkers).  Here a block column is 32 wide, lives in
// shared memory while it is factorised, and the trailing update A22 -= L21 L21^T runs on DMMA (m8n8k4) straight from
// that shared panel: 32 FMAs per read-modify-write, half the passes over the triangle, and the FP64 pipe does the work.
//   one CTA of 8 warps per walker; per panel (columns k0 .. k0+31):
//     1. the panel's 32 x 32 top block L_kk (shared, Tb) and the inverses W_s of its four 8 x 8 diagonal factors (sW) are
//        ready when the panel starts: warp 0 made them during the previous panel's trailing update (look-ahead; first
//        panel: up front) with a register Cholesky, lane = row (factor_top_block)
//     2. rows below the top block, one warp per 8-row tile and no block barrier in between: global -> shared, blocked
//        forward substitution X_s = (R_s - sum_{s' < s} X_s' L_ss'^T) W_s^T on DMMA, X -> shared (L21) and -> global (L)
//     3. trailing update A22 -= L21 L21^T: 32 x 32 tiles of the lower triangle of rows k1..p handed out through a shared
//        counter, 4 x 4 m8n8k4 accumulators, 8 k-steps; the C tile is fetched before the k-loop and written once.  Warp 0
//        takes tile (0, 0) first - the next panel's top block - and factorises it while the others do the rest.
//   Two block barriers per panel.  The first version factorised the panel 8 columns at a time (diagonal block by warp 0,
//   thread-per-row substitution, rank-8 update: three barriers per 8 columns, ten per panel) and assembled Sigma in pair
//   order: 12.6 ms at 8192 walkers; this one 8.0 ms (tensor pipe 21 -> 34 % busy).  What is left is the look-ahead itself:
//   its scalar FP64 chain shares the FP64 pipe with the other warps' DMMAs and runs ~5 x slower than alone, so the
//   trailing tiles still wait for it on the later, smaller panels (profiles/r2_ncu_chol_panel_rank300.txt).
// u rides along as row p (z = L^-1 u, entry (p, p) = -z.z) exactly as in the other finishing kernels.
constexpr int kPanelW = 32;
constexpr int kPanelPitch = kPanelW + 4;
constexpr int kPanelThreads = 256;

// 1 / sqrt(d) from the hardware seed (rsqrt.approx.f64: ~22 bits) and one third-order step: e = 1 - d y^2,
// y <- y + y e (1/2 + 3/8 e), good to ~1e-18 relative - four dependent FP64 operations where the library routine's
// special-case scaffolding sits on the factorisation's critical path.  d <= 0, Inf or NaN gives NaN (0 . Inf, Inf . 0).
__device__ __forceinline__ double rsqrt_fast(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-(d * y), y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}

// Cholesky of the panel's top block (nbt <= 32 rows / columns, lower triangle in Tb, pitch kPanelPitch, zero elsewhere) by
// ONE warp, in registers: lane r owns row r (32 doubles, statically indexed after unrolling; lanes past nbt hold identity
// rows).  Column by column: every lane publishes its (unscaled) entry u_r of column c in a shared column buffer, reads the
// pivot d = u_c from it, and updates row[q] -= (u_r / d) u_q with the u_q as broadcast shared loads; L_rc = u_r / sqrt(d) is
// formed off the critical path.  No shuffles (a shuffled double is two SHFLs plus, inside a function, a convergence
// collective around each: the shuffle version ran 10 k instructions) and no predicates: entries of the upper triangle are
// updated with junk that nobody reads.  A pivot that is not positive and finite turns into NaN under rsqrt and poisons the
// walker's ln L (mapped to -Inf at the end), so there is no test in the dependent chain.  Then the four 8 x 8 diagonal
// factors are inverted at once (lane = 8 block + column) into sW[block], and L goes back to Tb.  This runs while the other
// warps do the trailing update, so its length decides how many of them end up waiting (the first version - 8 columns at
// a time: diagonal block, inverse, solve and update on DMMA through shared memory - took 22 k cycles).
// colbuf: 64 + 32 doubles (two column buffers, then the reciprocal pivots).  Returns this lane's log L_rr (0 past nbt).
__device__ __noinline__ double factor_top_block(double* Tb, int nbt, double* colbuf, double* sW) {
    const int lane = threadIdx.x & 31;
    double* rdiag = colbuf + 64;
    double row[32];
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
        const double2 v = *reinterpret_cast<const double2*>(Tb + lane * kPanelPitch + c);
        row[c] = v.x;
        row[c + 1] = v.y;
    }
    if (lane >= nbt) {
#pragma unroll
        for (int c = 0; c < 32; ++c) row[c] = (c == lane) ? 1.0 : 0.0;
    }
    double myrd = 1.0;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        double* cb = colbuf + (c & 1) * 32;        // two buffers: a fast lane's next column cannot overtake a slow lane's reads
        cb[lane] = row[c];
        __syncwarp();
        const double d = cb[c];
        const double t = row[c] * rcp_fast(d);
        const double rd = rsqrt_fast(d);
        row[c] *= rd;                              // lane c: d / sqrt(d) = L_cc; lanes below: L_rc
        myrd = (lane == c) ? rd : myrd;
#pragma unroll
        for (int q = c + 1; q < 32; ++q) row[q] = fma(-t, cb[q], row[q]);
    }
    rdiag[lane] = myrd;
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
        // (lower triangle only: the junk above the diagonal must not reach the fragment reads of the solve)
        const double2 v = make_double2(c <= lane ? row[c] : 0.0, c + 1 <= lane ? row[c + 1] : 0.0);
        *reinterpret_cast<double2*>(Tb + lane * kPanelPitch + c) = v;
    }
    __syncwarp();
    {   // W = L_bb^-1 for the four diagonal blocks at once: lane = 8 b + column, forward substitution down the rows
        const int blk = lane >> 3, col = lane & 7;
        const double* L = Tb + (blk * 8) * kPanelPitch + blk * 8;
        double w[8];
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
            double acc = (rr == col) ? 1.0 : 0.0;
#pragma unroll
            for (int q = 0; q < rr; ++q) acc = fma(-L[rr * kPanelPitch + q], w[q], acc);
            w[rr] = acc * rdiag[blk * 8 + rr];
        }
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) sW[blk * 64 + rr * 8 + col] = w[rr];
    }
    __syncwarp();
    return lane < nbt ? log(Tb[lane * kPanelPitch + lane]) : 0.0;
}

__global__ void __launch_bounds__(kPanelThreads, 2) chol_panel_kernel(CholArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double scratch[kPanelThreads / 32];
    __shared__ int sFail, sTile;
    const long b = blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int p = a.p, pp = a.pp, n1 = p + 1;
    const int lda = pp + 2;                                // even pitch of the global matrix rows
    double* phi = smem;                                    // [pp]
    double* dblk = phi + pp;                               // [96 + 8] factor_top_block: two column buffers, reciprocal pivots
    double* sW = dblk + 104;                               // [4][8][8] inverses of the top block's diagonal factors
    double* Tb = sW + 256;                                 // [32][kPanelPitch] the panel's top block (L_kk)
    double* P = Tb + kPanelW * kPanelPitch;                // [prow][kPanelPitch] rows below it (L21); the column sums alias it during assembly
    double* sR = P;
    double* A = a.Sq + (size_t)b * n1 * lda;
    if (tid == 0) sFail = 0;

    // ---- Sigma (lower triangle, rows 0..p-1) and u (row p) into the global scratch
    {
        const long n_pairs = (long)p * (p + 1) / 2;
        const double* gA = a.Sig + (size_t)b * a.ld_sig;
        if (a.table) {
            const double* gR = a.R + (size_t)b * a.ld_r;
            if (a.ctab) {   // pre-scaled sums S = {+R/2 | -R/2 | 0} for the two-index codes
                for (int i = tid; i < a.n_r; i += nt) {
                    const double v = 0.5 * slab_sum(gR + i, a.split_stride_r, a.ksplit);
                    sR[i] = v;
                    sR[a.n_r + i] = -v;
                }
                if (tid == 0) sR[2 * a.n_r] = 0.0;
            } else {
                for (int i = tid; i < a.n_r; i += nt) sR[i] = slab_sum(gR + i, a.split_stride_r, a.ksplit);
            }
            __syncthreads();
        }
        if (a.ctab) {
            // 4-byte recipe per pair (two 16-bit indices into S: entry = S[i1] + S[i2]), row by row: warp = row, lanes along
            // the row, so the recipe loads and the stores are coalesced and no thread tracks (row, column) through the
            // packed order (the pair-ordered loop below spent 62 thread-instructions per pair on bookkeeping and on
            // decoding coefficient codes: 20 % of the kernel)
            // (eight codes in flight per lane: the loop is a chain of dependent L2 / shared / global accesses per entry and
            // the CTA has eight warps to hide it with)
            const unsigned zero_code = (unsigned)(2 * a.n_r) * 0x10001u;      // S[zero] + S[zero]
            const bool add_sig = a.add_sig != 0;
            for (int r = warp; r < p; r += nwarps) {
                const size_t t0 = (size_t)r * (r + 1) / 2;
                const unsigned* ct = a.ctab + t0;
                double* dst = A + (size_t)r * lda;
                for (int c0 = lane; c0 <= r; c0 += 256) {
                    unsigned e[8];
                    double g[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const bool in = c0 + 32 * u <= r;
                        e[u] = in ? __ldg(ct + c0 + 32 * u) : zero_code;
                        g[u] = (in && add_sig) ? gA[t0 + c0 + 32 * u] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        double v = sR[e[u] & 0xffffu] + sR[e[u] >> 16];
                        if (add_sig) v += g[u];
                        if (c0 + 32 * u <= r) dst[c0 + 32 * u] = v;
                    }
                }
            }
        } else {
        int r = 0, c = tid;
        while (c > r) { c -= r + 1; ++r; }
#pragma unroll 4
        for (long t = tid; t < n_pairs; t += nt) {
            double v;
            if (a.table) {
                const PairTerm pt = a.table[t];
                v = pt.c1 * sR[pt.i1];
                if (pt.c2 != 0.0) v = fma(pt.c2, sR[pt.i2], v);
                if (a.add_sig) v += gA[t];
            } else {
                v = slab_sum(gA + t, a.split_stride_sig, a.ksplit);
            }
            A[(size_t)r * lda + c] = v;
            c += nt;
            while (c > r) { c -= r + 1; ++r; }
        }
        }
        const double* gU = a.table ? a.R + a.u_off : a.U;
        const long ldu = a.table ? a.ld_r : a.ld_u, su = a.table ? a.split_stride_r : a.split_stride_u;
        for (int i = tid; i < p; i += nt) A[(size_t)p * lda + i] = slab_sum(gU + (size_t)b * ldu + i, su, a.ksplit);
        if (tid == 0) A[(size_t)p * lda + p] = 0.0;
    }
    // ---- Phi^-1
    const double* Pw = a.Pext + (size_t)b * a.row;
    for (int g = 0; g < a.n_gp; ++g) {
        GPDev gp = a.gp[g];
        if (gp.kind == VELA_GP_MARGINALIZED_TM) {
            for (int i = tid; i < gp.n; i += nt) phi[gp.col0 + i] = a.gp_data[gp.data0 + i];
        } else {
            const double fyr = 1.0 / 3600 / 24 / 365.25;
            const double denom = 12 * (CUDART_PI * CUDART_PI) * fyr * fyr * fyr;
            double amp = exp10(Pw[gp.par_log10_amp]), gamma = Pw[gp.par_gamma], f1 = Pw[gp.par_f1];
            double P1 = amp * amp / denom * pow(fyr / f1, gamma) * f1;
            for (int i = tid; i < gp.n; i += nt) {
                double v = 1 / (P1 * exp(-gamma * a.gp_data[gp.data0 + i]));
                phi[gp.col0 + i] = v;
                phi[gp.col0 + gp.n + i] = v;
            }
        }
    }
    __syncthreads();                                       // phi complete; the triangle in global is complete (same block)
    double lphi = 0.0;
    for (int i = tid; i < p; i += nt) {
        lphi += log(phi[i]);
        A[(size_t)i * lda + i] += phi[i];
    }
    const double logdetPhi = -block_sum(lphi, scratch);
    __syncthreads();

    // ---- blocked factorisation
    const int gid = lane >> 2, tig = lane & 3;
    double ldsum = 0.0;                                    // sum of log L_cc over the columns this thread finalises
    // top block of the first panel: nobody has anything else to do yet
    if (warp == 0) {
        const int nb0 = min(kPanelW, p);
        for (int e = lane; e < kPanelW * (kPanelW / 2); e += 32) {
            const int i = e >> 4, c = (e & 15) * 2;
            double2 v = make_double2(0.0, 0.0);
            if (i < nb0 && c <= i) {
                v = *reinterpret_cast<const double2*>(A + (size_t)i * lda + c);
                if (c + 1 > i) v.y = 0.0;
            }
            *reinterpret_cast<double2*>(Tb + i * kPanelPitch + c) = v;
        }
        __syncwarp();
        ldsum += factor_top_block(Tb, nb0, dblk, sW);
        for (int e = lane; e < nb0 * (kPanelW / 2); e += 32) {
            const int i = e >> 4, c = (e & 15) * 2;
            if (c <= i) {
                const double2 v = *reinterpret_cast<const double2*>(Tb + i * kPanelPitch + c);
                double* dst = A + (size_t)i * lda + c;
                if (c + 1 <= i) *reinterpret_cast<double2*>(dst) = v;
                else dst[0] = v.x;
            }
        }
    }
    for (int k0 = 0; k0 < p; k0 += kPanelW) {
        const int nb = min(kPanelW, p - k0), k1 = k0 + nb;
        const int m2 = n1 - k1;                            // rows below the top block: global rows k1 .. p (row p = u^T)
        const int T = (m2 + 31) >> 5;                      // 32-row tiles of the trailing matrix
        const int nsb = (nb + 7) >> 3;                     // 8-column blocks of this panel
        __syncthreads();                                   // Tb / sW of this panel are complete (first panel: above; later: look-ahead)
        if (tid == 0) sTile = 1;                           // trailing tiles are handed out dynamically; tile 0 is warp 0's
        // 1. rows below the top block, one warp per 8-row tile, no block barrier: R -> shared, then the blocked forward
        //    substitution X_s = (R_s - sum_{s' < s} X_s' L_ss'^T) W_s^T on DMMA (W_s = inverse of the s-th 8 x 8 diagonal
        //    factor), X -> shared (L21 for the trailing update) and -> global (it is L).  Rows are independent of each other,
        //    so the whole panel is solved between two block barriers (the first version went 8 columns at a time with
        //    three barriers each and every warp waiting for warp 0's diagonal block: 29 % of the kernel).
        for (int r0 = warp * 8; r0 < T * 32; r0 += nwarps * 8) {
            double* tile = P + r0 * kPanelPitch;
#pragma unroll
            for (int q = 0; q < 4; ++q) {                  // 8 rows x 16 column pairs, 4 per lane; a row is one 256-byte run
                const int i = q * 2 + (lane >> 4), c = (lane & 15) * 2;
                double2 v = make_double2(0.0, 0.0);
                if (r0 + i < m2 && c < nb) {
                    v = *reinterpret_cast<const double2*>(A + (size_t)(k1 + r0 + i) * lda + k0 + c);
                    if (c + 1 >= nb) v.y = 0.0;
                }
                *reinterpret_cast<double2*>(tile + i * kPanelPitch + c) = v;
            }
            __syncwarp();
            if (r0 >= m2) continue;                        // zero padding below the matrix
            double* xrow = tile + gid * kPanelPitch;
#pragma unroll
            for (int sb = 0; sb < kPanelW / 8; ++sb) {
                if (sb >= nsb) break;
                double2 acc = *reinterpret_cast<const double2*>(xrow + sb * 8 + 2 * tig);
#pragma unroll
                for (int sp = 0; sp < sb; ++sp) {
                    const double* lrow = Tb + (sb * 8 + gid) * kPanelPitch + sp * 8 + tig;
                    dmma8x8x4(acc.x, acc.y, -xrow[sp * 8 + tig], lrow[0]);
                    dmma8x8x4(acc.x, acc.y, -xrow[sp * 8 + 4 + tig], lrow[4]);
                }
                __syncwarp();                              // the tile's block sb has been read by every lane
                *reinterpret_cast<double2*>(xrow + sb * 8 + 2 * tig) = acc;
                __syncwarp();
                const double* w = sW + sb * 64 + gid * 8 + tig;
                double x0 = 0.0, x1 = 0.0;
                dmma8x8x4(x0, x1, xrow[sb * 8 + tig], w[0]);
                dmma8x8x4(x0, x1, xrow[sb * 8 + 4 + tig], w[4]);
                __syncwarp();
                *reinterpret_cast<double2*>(xrow + sb * 8 + 2 * tig) = make_double2(x0, x1);
                __syncwarp();
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = q * 2 + (lane >> 4), c = (lane & 15) * 2;
                if (r0 + i < m2 && c < nb) {
                    const double2 v = *reinterpret_cast<const double2*>(tile + i * kPanelPitch + c);
                    double* dst = A + (size_t)(k1 + r0 + i) * lda + k0 + c;
                    if (c + 1 < nb) *reinterpret_cast<double2*>(dst) = v;
                    else dst[0] = v.x;
                }
            }
        }
        __syncthreads();                                   // L21 complete in shared memory
        // 2. trailing update on DMMA: rows / columns k1 .. p, 32 x 32 tiles of the lower triangle, one warp each.  Warp 0
        //    takes tile (0, 0) - it holds the next panel's top block - copies the result into Tb and factorises it while the
        //    other warps work through the remaining tiles (look-ahead across panels: the serial diagonal work is hidden
        //    behind the trailing update whenever that has more than a few tiles).
        const int n_tiles = T * (T + 1) / 2;
        const int nb_next = min(kPanelW, p - k1);
        bool first = warp == 0;
        for (;;) {
            int tile;
            if (first) tile = 0;
            else {
                tile = 0;
                if (lane == 0) tile = atomicAdd(&sTile, 1);
                tile = __shfl_sync(0xffffffffu, tile, 0);
            }
            if (tile >= n_tiles) break;
            int ti = 0;
            while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti;
            const int tj = tile - ti * (ti + 1) / 2;
            // the C tile is the accumulators' initial value (A enters negated): its global loads are in flight while the
            // fragments come from shared memory, and no second register tile is needed
            double acc[4][4][2];
            // tiles strictly below the diagonal with all 32 rows inside the matrix (most of them) need no predicates
            const bool interior = ti > tj && ti * 32 + 32 <= m2;
            double* ctile = A + (size_t)(k1 + ti * 32 + gid) * lda + k1 + tj * 32 + 2 * tig;
            if (interior) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double2 v = *reinterpret_cast<const double2*>(ctile + (size_t)(i * 8) * lda + j * 8);
                        acc[i][j][0] = v.x;
                        acc[i][j][1] = v.y;
                    }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int rr = ti * 32 + i * 8 + gid, cc = tj * 32 + j * 8 + 2 * tig;
                        acc[i][j][0] = acc[i][j][1] = 0.0;
                        if (rr < m2) {
                            const double* src = A + (size_t)(k1 + rr) * lda + k1 + cc;
                            if (cc + 1 <= rr) {
                                const double2 v = *reinterpret_cast<const double2*>(src);
                                acc[i][j][0] = v.x;
                                acc[i][j][1] = v.y;
                            } else if (cc <= rr) {
                                acc[i][j][0] = src[0];
                            }
                        }
                    }
            }
            const double* pa = P + (ti * 32 + gid) * kPanelPitch + tig;
            const double* pb = P + (tj * 32 + gid) * kPanelPitch + tig;
#pragma unroll
            for (int kk = 0; kk < kPanelW / 4; ++kk) {
                double fa[4], fb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    fa[i] = -pa[i * 8 * kPanelPitch + kk * 4];
                    fb[i] = pb[i * 8 * kPanelPitch + kk * 4];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
            }
            if (interior) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<double2*>(ctile + (size_t)(i * 8) * lda + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int rr = ti * 32 + i * 8 + gid, cc = tj * 32 + j * 8 + 2 * tig;
                        if (rr < m2) {
                            double* dst = A + (size_t)(k1 + rr) * lda + k1 + cc;
                            if (cc + 1 <= rr) *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
                            else if (cc <= rr) dst[0] = acc[i][j][0];
                        }
                    }
            }
            if (first) {
                first = false;
                if (nb_next > 0) {
                    // the next top block (rows / columns k1 .. k1 + nb_next - 1): lower triangle into Tb, zero elsewhere.
                    // Tb and sW of the current panel were last read before the barrier above.
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int rr = i * 8 + gid, cc = j * 8 + 2 * tig;
                            double2 v = make_double2(0.0, 0.0);
                            if (rr < nb_next) {
                                if (cc <= rr) v.x = acc[i][j][0];
                                if (cc + 1 <= rr) v.y = acc[i][j][1];
                            }
                            *reinterpret_cast<double2*>(Tb + rr * kPanelPitch + cc) = v;
                        }
                    __syncwarp();
                    ldsum += factor_top_block(Tb, nb_next, dblk, sW);
                    for (int e = lane; e < nb_next * (kPanelW / 2); e += 32) {
                        const int i = e >> 4, c = (e & 15) * 2;
                        if (c <= i) {
                            const double2 v = *reinterpret_cast<const double2*>(Tb + i * kPanelPitch + c);
                            double* dst = A + (size_t)(k1 + i) * lda + k1 + c;
                            if (c + 1 <= i) *reinterpret_cast<double2*>(dst) = v;
                            else dst[0] = v.x;
                        }
                    }
                }
            }
        }
    }
    __syncthreads();

    const double logdetSig = 2 * block_sum(ldsum, scratch);
    double py = 0.0, pl = 0.0;
    for (int t = tid; t < a.n_tiles; t += nt) {
        const double2 v = *reinterpret_cast<const double2*>(a.partial + ((size_t)t * a.B + b) * 2);
        py += v.x;
        pl += v.y;
    }
    const double yNy = block_sum(py, scratch), logdetN = block_sum(pl, scratch);
    if (tid == 0) {
        const double zz = -A[(size_t)p * lda + p];
        double lnl = -0.5 * (yNy - zz + logdetN + logdetPhi + logdetSig);
        if (sFail) lnl = -CUDART_INF;
        a.out[b] = combine_prior(lnl, a.lnprior, b, a.with_prior);
    }
}

// ------------------------------------------------------------------------------------------
// K4w  Cholesky finish, one WARP per walker (ranks up to 63, structured Sigma path)
// ------------------------------------------------------------------------------------------
// Same quantities as chol_finish_kernel, for matrices of at most 64 rows (p + 1 <= 64, row p = u^T): the whole
// factorisation runs inside one warp, so there is no block barrier anywhere (the CTA-per-walker kernel waits on its
// barriers a third of the time at rank 60) and a walker occupies 32 threads and ~18 KB of shared memory.
//   * lane l owns matrix rows l and l + 32; the lower triangle sits in shared memory column by column, column k
//     holding rows (k & ~3) .. NR-1 (NR = p + 1 rounded up to even), so the four columns of a block share one row
//     range, the own-row reads of a warp are one contiguous row segment and the four entries L[j0..j0+3][k] every lane
//     needs are two aligned 16-byte broadcasts (layout: cw_col_base / chol_warp_layout, vela_kernels.cuh);
//   * assembly through a LAYOUT table built at create(): one 4-byte code per shared-memory slot (zero | c1 R[i1] +
//     c2 R[i2] | copy of u_i), so the loop that fills the matrix is a coalesced table read, two shared loads and one
//     store per slot, with no index arithmetic and no separate zero fill;
//   * left-looking, four columns at a time: the block's entries of the lane's two rows are 8 accumulators; for every
//     earlier column k: acc[s][c] -= L[row_s][k] L[j0 + c][k] (own rows: 2 LDS.64, the block's rows: 2 broadcast
//     LDS.128, 8 DFMA; column pointers advance by additions only); then the 4 x 4 diagonal block is factorised in
//     place in the accumulators, pivots and multipliers travelling by shuffle - no separate panel solve and no
//     shared-memory round trip in the dependent chain; the columns are stored once, finished;
//   * u rides along as row p (column p takes no pivot): its row ends as z = L^-1 u and entry (p, p) as -z.z;
//   * log det Sigma^-1 = sum log(pivot) and log det Phi are mantissa products with exponent sums: two logarithms per
//     walker instead of 2 p.
template <bool LOW, bool TWO>
__device__ __forceinline__ void cw_block_step(double* A, int NR, int jb, int p, int lane, int& fail, double& lman, int& lex) {
    const int j0 = 4 * jb;
    const int lenj = NR - j0;
    double* cj = A + cw_col_base(jb, NR) - j0;            // cj[c * lenj + row] = entry (row, j0 + c)
    double acc0[4], acc1[4];
    const bool own0 = LOW && lane >= j0 && lane < NR, own1 = TWO && lane + 32 >= j0 && lane + 32 < NR;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        acc0[c] = own0 ? cj[c * lenj + lane] : 0.0;
        acc1[c] = own1 ? cj[c * lenj + lane + 32] : 0.0;
    }
    {
        // cp[row] = entry (row, 4 m) of the earlier column block m; the next column is `len` doubles further, the
        // next block 4 len - 4
        const double* cp = A;
        int len = NR;
#pragma unroll 2
        for (int m = 0; m < jb; ++m) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double* cc = cp + r * len;
                const double2 b01 = *reinterpret_cast<const double2*>(cc + j0);
                const double2 b23 = *reinterpret_cast<const double2*>(cc + j0 + 2);
                if (LOW) {
                    const double a0 = cc[lane];   // rows above j0 read finished entries (or a neighbouring column): never stored
                    acc0[0] = fma(-a0, b01.x, acc0[0]); acc0[1] = fma(-a0, b01.y, acc0[1]);
                    acc0[2] = fma(-a0, b23.x, acc0[2]); acc0[3] = fma(-a0, b23.y, acc0[3]);
                }
                if (TWO) {
                    const double a1 = cc[lane + 32];
                    acc1[0] = fma(-a1, b01.x, acc1[0]); acc1[1] = fma(-a1, b01.y, acc1[1]);
                    acc1[2] = fma(-a1, b23.x, acc1[2]); acc1[3] = fma(-a1, b23.y, acc1[3]);
                }
            }
            cp += 4 * len - 4;
            len -= 4;
        }
    }
    // 4 x 4 diagonal block + panel, in the accumulators: the rows j0 .. j0 + 3 belong to lanes (j0 & 31) .. + 3
    const int dl = j0 & 31;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int col = j0 + c;
        if (col >= p) break;                               // column p (u) takes no pivot; columns past it do not exist
        double d = __shfl_sync(0xffffffffu, LOW ? acc0[c] : acc1[c], dl + c);
        if (!(d > 0.0) || !isfinite(d)) { fail = 1; d = 1.0; }
        int ex;
        lman *= split_mantissa(d, &ex);
        lex += ex;
        const double rd = rsqrt(d);
        if (LOW) acc0[c] *= rd;
        if (TWO) acc1[c] *= rd;
#pragma unroll
        for (int c2 = c + 1; c2 < 4; ++c2) {
            const double lv = __shfl_sync(0xffffffffu, LOW ? acc0[c] : acc1[c], dl + c2);   // L[j0 + c2][col]
            if (LOW) acc0[c2] = fma(-acc0[c], lv, acc0[c2]);
            if (TWO) acc1[c2] = fma(-acc1[c], lv, acc1[c2]);
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (own0) cj[c * lenj + lane] = acc0[c];
        if (own1) cj[c * lenj + lane + 32] = acc1[c];
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32) chol_warp_kernel(CholArgs a) {
    extern __shared__ __align__(16) double smem[];
    const long b = blockIdx.x;
    const int lane = threadIdx.x;
    const int p = a.p, n1 = p + 1;
    const int NR = (n1 + 1) & ~1, nblk = (n1 + 3) / 4;
    const int nA = cw_col_base(nblk, NR);
    double* A = smem;
    double* sR = A + nA + kCholWarpSlack;                  // column sums, then u
    // entry (row, col <= row)
    auto at = [&](int row, int col) -> double& { return A[cw_col_base(col >> 2, NR) + (col & 3) * (NR - (col & ~3)) + row - (col & ~3)]; };

    // ---- the column sums and u; split-K slabs are summed in slice order
    {
        const double* gR = a.R + (size_t)b * a.ld_r;
        const int n_ru = a.n_r + p;
        // every load of the walker's row is independent: unrolled so that they are all in flight together (one warp
        // has nobody to hide a global-memory round trip behind)
        if (a.ksplit == 1) {
#pragma unroll 8
            for (int i = lane; i < a.n_r; i += 32) sR[i] = gR[i];
            for (int i = lane; i < p; i += 32) sR[a.n_r + i] = gR[a.u_off + i];
        } else {
#pragma unroll 4
            for (int i = lane; i < a.n_r; i += 32) sR[i] = slab_sum(gR + i, a.split_stride_r, a.ksplit);
            for (int i = lane; i < p; i += 32) sR[a.n_r + i] = slab_sum(gR + a.u_off + i, a.split_stride_r, a.ksplit);
        }
        (void)n_ru;
    }
    __syncwarp();
    // ---- Sigma and the u row through the layout table: one code per shared-memory slot
#pragma unroll 8
    for (int i = lane; i < nA; i += 32) {
        const unsigned e = __ldg(a.ltab + i);              // the same 8 KB for every walker: read-only path, L1-resident
        const unsigned kind = e >> 29;
        double v = 0.0;
        if (kind == 1) {
            v = ((e >> 26) & 1 ? 1.0 : 0.5) * sR[e & 0x1fff];
            const unsigned c2 = (e >> 27) & 3;
            if (c2) v = fma(c2 == 1 ? 0.5 : -0.5, sR[(e >> 13) & 0x1fff], v);
        } else if (kind == 2) {
            v = sR[e & 0x1fff];
        }
        A[i] = v;
    }
    for (int i = lane; i < kCholWarpSlack; i += 32) A[nA + i] = 0.0;
    __syncwarp();
    // ---- + Phi^-1 on the diagonal (calc_noise_weights_inv); log det Phi as a mantissa product
    const double* P = a.Pext + (size_t)b * a.row;
    double pman = 1.0;
    int pex = 0;
    for (int g = 0; g < a.n_gp; ++g) {
        const GPDev gp = a.gp[g];
        if (gp.kind == VELA_GP_MARGINALIZED_TM) {
            for (int i = lane; i < gp.n; i += 32) {
                const double v = a.gp_data[gp.data0 + i];
                at(gp.col0 + i, gp.col0 + i) += v;
                int ex;
                pman *= split_mantissa(v, &ex);
                pex += ex;
            }
        } else {
            // powerlaw(A, gamma, f1, f1) gp_noise.jl:12-16 with A = exp10(log10_A)
            const double fyr = 1.0 / 3600 / 24 / 365.25;
            const double denom = 12 * (CUDART_PI * CUDART_PI) * fyr * fyr * fyr;
            const double amp = exp10(P[gp.par_log10_amp]), gamma = P[gp.par_gamma], f1 = P[gp.par_f1];
            const double P1 = amp * amp / denom * pow(fyr / f1, gamma) * f1;
            for (int i = lane; i < gp.n; i += 32) {
                const double v = 1 / (P1 * exp(-gamma * a.gp_data[gp.data0 + i]));
                at(gp.col0 + i, gp.col0 + i) += v;
                at(gp.col0 + gp.n + i, gp.col0 + gp.n + i) += v;
                int ex;
                const double mv = split_mantissa(v, &ex);   // the weight serves the sine and the cosine column
                pman *= mv * mv;
                pex += 2 * ex;
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {               // <= 32 lanes x a few mantissas each: far from underflow
        pman *= __shfl_xor_sync(0xffffffffu, pman, off);
        pex += __shfl_xor_sync(0xffffffffu, pex, off);
    }
    const double logdetPhi = -(log(pman) + pex * 0.6931471805599453);   // log det Phi = -sum log Phi^-1
    __syncwarp();

    // ---- Cholesky (isposdef + cholesky!, gls_likelihood.jl:211-217) with z = L^-1 u riding along (ldiv!, :220)
    int fail = 0, lex = 0;
    double lman = 1.0;
    const bool two = NR > 32;
    for (int jb = 0; jb < nblk; ++jb) {
        if (4 * jb < 32) {
            if (two) cw_block_step<true, true>(A, NR, jb, p, lane, fail, lman, lex);
            else cw_block_step<true, false>(A, NR, jb, p, lane, fail, lman, lex);
        } else {
            cw_block_step<false, true>(A, NR, jb, p, lane, fail, lman, lex);
        }
        if ((jb & 7) == 7) {                               // 32 pivots: fold the mantissa product into the exponent sum
            int ex;
            lman = split_mantissa(lman, &ex);
            lex += ex;
        }
    }
    const double logdetSig = log(lman) + lex * 0.6931471805599453;     // sum log(pivot) = 2 sum log L_cc
    // y^T N^-1 y and log det N: partial rows of the chain tiles (and ECORR blocks)
    double py = 0.0, pl = 0.0;
    for (int t = lane; t < a.n_tiles; t += 32) {
        const double2 v = *reinterpret_cast<const double2*>(a.partial + ((size_t)t * a.B + b) * 2);
        py += v.x;
        pl += v.y;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        py += __shfl_xor_sync(0xffffffffu, py, off);
        pl += __shfl_xor_sync(0xffffffffu, pl, off);
    }
    if (lane == 0) {
        const double zz = -at(p, p);
        double lnl = -0.5 * (py - zz + pl + logdetPhi + logdetSig);
        if (fail) lnl = -CUDART_INF;
        a.out[b] = combine_prior(lnl, a.lnprior, b, a.with_prior);
    }
}

// ------------------------------------------------------------------------------------------
// K4d  Cholesky finish on the FP64 tensor cores, one WARP per walker, matrix in registers (ranks up to 64)
// ------------------------------------------------------------------------------------------
// Same quantities as chol_finish_kernel.  Sigma^-1 is cut into 8 x 8 blocks (NBK block rows, identity padding past
// rank p) and u^T is one more block row; every block lives in two registers per lane, in the operand layout of
// mma.sync.m8n8k4.f64 with rows AND columns permuted by pi(n) = 4 (n & 1) + (n >> 1):
//     lane l = 4 r + t, register e  holds  X[pi(r)][4 e + t].
// In that layout the two registers of a block ARE its A fragments (k-halves e = 0, 1) and its B fragments, and a
// product A.B^T of two such blocks comes out of the DMMA in the same layout again, so the blocked right-looking
// factorisation needs no data movement between its steps:
//   * diagonal block (k, k): its rows are gathered by shuffle (lane & 7 = row, four replicas), factorised in registers
//     as in factor_diag_block_warp, and the inverse W of the 8 x 8 factor is formed column by column (lane = column;
//     the factor's entries are broadcast shared-memory reads);
//   * panel: L(I, k) = A(I, k) W^T - two DMMAs per block, W^T read from shared memory in B layout;
//   * trailing update: A(I, J) -= L(I, k) L(J, k)^T - two DMMAs per block, operands = the panel registers.
// All blocks are stored NEGATED, so the panel product yields -L and the update adds (-L)(-L)^T: no negations anywhere.
// The u row rides along as block row NBK: its panel blocks end as -z^T (z = L^-1 u), z.z is accumulated from them.
// Per walker: ~6 k warp instructions against ~21 k of the CTA-per-walker kernel, and a dependent chain of 8 block
// steps with no barrier and no shared-memory matrix traffic.  The blocks are filled through a fragment table built at
// create(): per lane and register two 16-bit indices into the walker's pre-scaled column sums (entry = S[i1] + S[i2]).
// The factorisation of the diagonal block and its inverse are ONE copy of code inside a rolled loop over the block
// steps (the first version unrolled everything: 64 KB of straight-line code, and the warps waited for instructions
// more than for anything else); only the DMMA steps are unrolled, selected by a switch.
// Step K of the blocked factorisation on the register blocks: panel products against W^T (wf0 / wf1: its two k-halves
// in B layout) and the trailing update.  Every block index is a compile-time constant.
template <int NBK, int K>
__device__ __forceinline__ void cd_step(double (&m)[NBK * (NBK + 1) / 2 + NBK][2], double wf0, double wf1, double& zz) {
    constexpr int NT0 = NBK * (NBK + 1) / 2;
    // panel: -L(I, K) = (-A(I, K)) W^T; the u row's blocks end as -z^T
#pragma unroll
    for (int I = K + 1; I <= NBK; ++I) {
        const int blk = I < NBK ? I * (I + 1) / 2 + K : NT0 + K;
        double x0 = 0.0, x1 = 0.0;
        dmma8x8x4(x0, x1, m[blk][0], wf0);
        dmma8x8x4(x0, x1, m[blk][1], wf1);
        m[blk][0] = x0; m[blk][1] = x1;
        if (I == NBK) zz = fma(x0, x0, fma(x1, x1, zz));
    }
    // trailing update: (-A(I, J)) += (-L(I, K)) (-L(J, K))^T
#pragma unroll
    for (int I = K + 1; I <= NBK; ++I) {
        const int bi = I < NBK ? I * (I + 1) / 2 + K : NT0 + K;
#pragma unroll
        for (int J = K + 1; J <= (I < NBK ? I : NBK - 1); ++J) {
            const int bj = J * (J + 1) / 2 + K;
            const int bij = I < NBK ? I * (I + 1) / 2 + J : NT0 + J;
            dmma8x8x4(m[bij][0], m[bij][1], m[bi][0], m[bj][0]);
            dmma8x8x4(m[bij][0], m[bij][1], m[bi][1], m[bj][1]);
        }
    }
}

template <int NBK>
__global__ void __launch_bounds__(32) chol_dmma_kernel(CholArgs a) {
    constexpr int NT0 = NBK * (NBK + 1) / 2, NT = NT0 + NBK;
    constexpr unsigned kFull = 0xffffffffu;
    extern __shared__ __align__(16) double smem[];
    const long b = blockIdx.x;
    const int lane = threadIdx.x, r = lane >> 2, t = lane & 3;
    const int p = a.p, n_ru = a.n_r + p;
    double* sW = smem;                 // [64] inverse of the current diagonal factor, row-major
    double* sL = sW + 64;              // [64] the current diagonal factor, row-major, reciprocal pivots on the diagonal
    double* sPhi = sL + 72;            // [8 NBK] Phi^-1 (zero past p)
    double* sS = sPhi + 8 * NBK;       // [2 n_ru + 2]: -R / 2, -u / 2 | +R / 2, +u / 2 | 0 | -1 / 2

    // ---- the column sums and u (split-K slabs summed in slice order), pre-scaled for the assembly below; u follows the
    //      column sums both in R (u_off == n_r rounded to 32: read through a second index) and in S
    {
        const double* gR = a.R + (size_t)b * a.ld_r;
#pragma unroll 4
        for (int i = lane; i < n_ru; i += 32) {
            const double* g = gR + (i < a.n_r ? i : a.u_off + (i - a.n_r));
            // (one to four slabs at full size: plain loads, all in flight; emcee-sized batches split K up to 64 ways)
            double v = a.ksplit <= 4 ? g[0] : slab_sum(g, a.split_stride_r, a.ksplit);
            if (a.ksplit <= 4) {
                const double v1 = a.ksplit > 1 ? g[a.split_stride_r] : 0.0, v2 = a.ksplit > 2 ? g[2 * a.split_stride_r] : 0.0,
                             v3 = a.ksplit > 3 ? g[3 * a.split_stride_r] : 0.0;
                v = ((v + v1) + v2) + v3;
            }
            v *= 0.5;
            sS[i] = -v; sS[n_ru + i] = v;
        }
        if (lane == 0) { sS[2 * n_ru] = 0.0; sS[2 * n_ru + 1] = -0.5; }
    }
    // ---- Phi^-1 (calc_noise_weights_inv); log det Phi as a mantissa product
    for (int i = lane; i < 8 * NBK; i += 32) sPhi[i] = 0.0;
    __syncwarp();
    const double* P = a.Pext + (size_t)b * a.row;
    double pman = 1.0;
    int pex = 0;
    for (int g = 0; g < a.n_gp; ++g) {
        const GPDev gp = a.gp[g];
        if (gp.kind == VELA_GP_MARGINALIZED_TM) {
            for (int i = lane; i < gp.n; i += 32) {
                const double v = a.gp_data[gp.data0 + i];
                sPhi[gp.col0 + i] = v;
                int ex;
                pman *= split_mantissa(v, &ex);
                pex += ex;
            }
        } else {
            // powerlaw(A, gamma, f1, f1) gp_noise.jl:12-16 with A = exp10(log10_A)
            const double fyr = 1.0 / 3600 / 24 / 365.25;
            const double denom = 12 * (CUDART_PI * CUDART_PI) * fyr * fyr * fyr;
            const double amp = exp10(P[gp.par_log10_amp]), gamma = P[gp.par_gamma], f1 = P[gp.par_f1];
            const double P1 = amp * amp / denom * pow(fyr / f1, gamma) * f1;
            for (int i = lane; i < gp.n; i += 32) {
                const double v = 1 / (P1 * exp(-gamma * a.gp_data[gp.data0 + i]));
                sPhi[gp.col0 + i] = v;
                sPhi[gp.col0 + gp.n + i] = v;
                int ex;
                const double mv = split_mantissa(v, &ex);   // the weight serves the sine and the cosine column
                pman *= mv * mv;
                pex += 2 * ex;
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        pman *= __shfl_xor_sync(kFull, pman, off);
        pex += __shfl_xor_sync(kFull, pex, off);
    }
    const double logdetPhi = -(log(pman) + pex * 0.6931471805599453);   // log det Phi = -sum log Phi^-1
    __syncwarp();

    // ---- the blocks, negated: entry = S[i1] + S[i2] with the two 16-bit indices of the fragment table (four entries
    //      per 16-byte load; 0.5 (R1 +- R2), R1 = 0.5 (R1 + R1), the identity padding and zero all take this one form)
    double m[NT][2];
    {
        const uint4* ft = reinterpret_cast<const uint4*>(a.ftab) + lane;
#pragma unroll
        for (int g = 0; g < (2 * NT + 3) / 4; ++g) {
            const uint4 c4 = __ldg(ft + g * 32);
            const unsigned cs[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (4 * g + q < 2 * NT) m[(4 * g + q) >> 1][(4 * g + q) & 1] = sS[cs[q] & 0xffffu] + sS[cs[q] >> 16];
        }
    }
    if (t == (r >> 1)) {   // this lane holds a diagonal entry of every diagonal block: row pi(r), register r & 1
#pragma unroll
        for (int I = 0; I < NBK; ++I) {
            const double ph = sPhi[8 * I + cd_pi(r)];
            if (r & 1) m[I * (I + 1) / 2 + I][1] -= ph;
            else m[I * (I + 1) / 2 + I][0] -= ph;
        }
    }

    // ---- blocked Cholesky (isposdef + cholesky!, gls_likelihood.jl:211-217) with z = L^-1 u riding along (ldiv!, :220).
    // A pivot that is not positive and finite turns into NaN under rsqrt and poisons everything after it, ln L included:
    // no test inside the dependent chain (combine_prior maps NaN to -Inf).
    int lex = 0;
    double lman = 1.0, zz = 0.0;
    const int i8 = lane & 7;
    const int src0 = 4 * (2 * (i8 & 3) + (i8 >> 2));   // lanes holding row i8 of a block: 4 pi^-1(i8) + (column & 3)
#pragma unroll 1
    for (int k = 0; k < NBK; ++k) {
        double d0, d1;
        switch (k) {   // the diagonal block's registers (every case a compile-time index)
            case 0: d0 = m[0][0]; d1 = m[0][1]; break;
            case 1: d0 = m[NBK > 1 ? 2 : 0][0]; d1 = m[NBK > 1 ? 2 : 0][1]; break;
            case 2: d0 = m[NBK > 2 ? 5 : 0][0]; d1 = m[NBK > 2 ? 5 : 0][1]; break;
            case 3: d0 = m[NBK > 3 ? 9 : 0][0]; d1 = m[NBK > 3 ? 9 : 0][1]; break;
            case 4: d0 = m[NBK > 4 ? 14 : 0][0]; d1 = m[NBK > 4 ? 14 : 0][1]; break;
            case 5: d0 = m[NBK > 5 ? 20 : 0][0]; d1 = m[NBK > 5 ? 20 : 0][1]; break;
            case 6: d0 = m[NBK > 6 ? 27 : 0][0]; d1 = m[NBK > 6 ? 27 : 0][1]; break;
            default: d0 = m[NBK > 7 ? 35 : 0][0]; d1 = m[NBK > 7 ? 35 : 0][1]; break;
        }
        {   // (1) diagonal block: gather rows (lane & 7 = row, four replicas), factorise in registers
            double row[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) row[c] = -__shfl_sync(kFull, c < 4 ? d0 : d1, src0 + (c & 3));
            double prod = 1.0, myrd = 0.0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const double d = __shfl_sync(kFull, row[c], c);      // pivot, held by the lanes of row c
                prod *= d;
                const double rd = rsqrt(d);
                // No predicates: the lanes of rows above c scale / update entries of the block's upper triangle, which
                // nobody reads (pivots and multipliers come from the lower triangle of the lanes that own them), and for
                // the pivot's own lane row[c] == d, so row[c] * rd is d * rd = L[c][c].
                row[c] *= rd;
                myrd = (i8 == c) ? rd : myrd;      // the reciprocal pivot of this lane's own row: stored on the diagonal below
#pragma unroll
                for (int q = c + 1; q < 8; ++q) {
                    const double lqc = __shfl_sync(kFull, row[c], q);   // L[q][c]
                    row[q] = fma(-row[c], lqc, row[q]);
                }
            }
            int ex;                                                  // 8 pivots of ~1e14 at most: no overflow
            lman *= split_mantissa(prod, &ex);
            lex += ex;
            if (lane < 8) {
#pragma unroll
                for (int c = 0; c < 8; ++c) sL[lane * 8 + c] = (c == lane) ? myrd : row[c];   // diagonal: 1 / L[c][c], all the inverse needs
            }
        }
        __syncwarp();
        {   // (2) W = L_kk^-1, lane & 7 = column: w[rr] = (delta - sum_q L[rr][q] w[q]) / L[rr][rr]
            double w[8];
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                double acc = (rr == i8) ? 1.0 : 0.0;
#pragma unroll
                for (int q = 0; q < rr; ++q) acc = fma(-sL[rr * 8 + q], w[q], acc);
                w[rr] = acc * sL[rr * 9];
            }
            if (lane < 8) {
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) sW[rr * 8 + lane] = w[rr];
            }
        }
        __syncwarp();
        const double wf0 = sW[cd_pi(r) * 8 + t], wf1 = sW[cd_pi(r) * 8 + 4 + t];   // W^T as the B operand, columns permuted
        __syncwarp();
        switch (k) {   // (3) + (4) panel products and trailing update of this step
            case 0: cd_step<NBK, 0>(m, wf0, wf1, zz); break;
            case 1: cd_step<NBK, (NBK > 1 ? 1 : 0)>(m, wf0, wf1, zz); break;
            case 2: cd_step<NBK, (NBK > 2 ? 2 : 0)>(m, wf0, wf1, zz); break;
            case 3: cd_step<NBK, (NBK > 3 ? 3 : 0)>(m, wf0, wf1, zz); break;
            case 4: cd_step<NBK, (NBK > 4 ? 4 : 0)>(m, wf0, wf1, zz); break;
            case 5: cd_step<NBK, (NBK > 5 ? 5 : 0)>(m, wf0, wf1, zz); break;
            case 6: cd_step<NBK, (NBK > 6 ? 6 : 0)>(m, wf0, wf1, zz); break;
            default: cd_step<NBK, (NBK > 7 ? 7 : 0)>(m, wf0, wf1, zz); break;
        }
    }
    const double logdetSig = log(lman) + lex * 0.6931471805599453;     // sum log(pivot) = 2 sum log L_cc
    // y^T N^-1 y and log det N: partial rows of the chain tiles (and ECORR blocks)
    double py = 0.0, pl = 0.0;
    for (int tt = lane; tt < a.n_tiles; tt += 32) {
        const double2 v = *reinterpret_cast<const double2*>(a.partial + ((size_t)tt * a.B + b) * 2);
        py += v.x;
        pl += v.y;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        py += __shfl_xor_sync(kFull, py, off);
        pl += __shfl_xor_sync(kFull, pl, off);
        zz += __shfl_xor_sync(kFull, zz, off);
    }
    if (lane == 0) {
        const double lnl = -0.5 * (py - zz + pl + logdetPhi + logdetSig);   // NaN (failed factorisation) -> -Inf below
        a.out[b] = combine_prior(lnl, a.lnprior, b, a.with_prior);
    }
}
template __global__ void chol_dmma_kernel<2>(CholArgs);
template __global__ void chol_dmma_kernel<4>(CholArgs);
template __global__ void chol_dmma_kernel<6>(CholArgs);
template __global__ void chol_dmma_kernel<8>(CholArgs);

// Phi^-1 only (vela_b200_noise_weights_inv): one block, walker 0
__global__ void phiinv_kernel(CholArgs a, double* __restrict__ out) {
    const double* P = a.Pext;
    for (int g = 0; g < a.n_gp; ++g) {
        GPDev gp = a.gp[g];
        if (gp.kind == VELA_GP_MARGINALIZED_TM) {
            for (int i = threadIdx.x; i < gp.n; i += blockDim.x) out[gp.col0 + i] = a.gp_data[gp.data0 + i];
        } else {
            const double fyr = 1.0 / 3600 / 24 / 365.25;
            const double denom = 12 * (CUDART_PI * CUDART_PI) * fyr * fyr * fyr;
            double amp = exp10(P[gp.par_log10_amp]), gamma = P[gp.par_gamma], f1 = P[gp.par_f1];
            double P1 = amp * amp / denom * pow(fyr / f1, gamma) * f1;
            for (int i = threadIdx.x; i < gp.n; i += blockDim.x) {
                double v = 1 / (P1 * exp(-gamma * a.gp_data[gp.data0 + i]));
                out[gp.col0 + i] = v;
                out[gp.col0 + gp.n + i] = v;
            }
        }
    }
}

// test hook for div3_by / div_fast / rcp_fast: counts quotients that differ from the IEEE division in any bit
__global__ void div3_check_kernel(unsigned long long seed, long n, unsigned long long* mismatches) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long st = seed + 0x9E3779B97F4A7C15ULL * (unsigned long long)(i + 1);
    auto next = [&]() {   // splitmix64
        st += 0x9E3779B97F4A7C15ULL;
        unsigned long long z = st;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    };
    auto unit = [&]() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); };
    // a unit vector plus a proper-motion-sized offset, as evaluate_proper_motion sees it; every 4th sample a wide range
    double a[3], b;
    const bool wide = (i & 3) == 3;
    for (int k = 0; k < 3; ++k) a[k] = 2.0 * unit() - 1.0;
    if (wide) b = 0.26 + 3.7 * unit();
    else b = sqrt(1.0 + 1e-6 * (2.0 * unit() - 1.0));
    double q[3];
    div3_by(a[0], a[1], a[2], b, q);
    unsigned long long bad = 0;
    for (int k = 0; k < 3; ++k) bad += __double_as_longlong(q[k]) != __double_as_longlong(a[k] / b);
    // div_fast / rcp_fast over the range the chain kernel's operands live in (and well beyond): 2^-130 .. 2^130
    const double num = (1.0 + unit()) * exp2((double)((int)(next() % 261) - 130)) * ((next() & 1) ? 1.0 : -1.0);
    const double den = (1.0 + unit()) * exp2((double)((int)(next() % 261) - 130)) * ((next() & 1) ? 1.0 : -1.0);
    bad += __double_as_longlong(div_fast(num, den)) != __double_as_longlong(num / den);
    bad += __double_as_longlong(rcp_fast(den)) != __double_as_longlong(1.0 / den);
    if (bad) atomicAdd(mismatches, bad);
}

// test hook for sincos_cr
__global__ void sincos_cr_kernel(const double* x, int n, double* s, double* c) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sincos_cr(x[i], s + i, c + i);
}

// ------------------------------------------------------------------------------------------
// FP64 peak micro-benchmarks (roofline denominators)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double a = 1.0000001, c = 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
    double af = 1.0 + 1e-9 * threadIdx.x, bf = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma8x8x4(acc[i][0], acc[i][1], af, bf);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Co-issue probe: the even warps of every block run the register-resident DFMA loop (mode & 1), the odd warps the DMMA
// loop (mode & 2).  If scalar FP64 and the FP64 tensor path were separate pipes, mode 3 would deliver the sum of modes 1
// and 2; on B200 it delivers what either delivers alone (profiles/r2_fp64_coissue.txt), i.e. one shared datapath - which
// is why the chain kernel and the column-sum kernel are not fused into one warp-specialised kernel.
__global__ void __launch_bounds__(256) fp64_coissue_kernel(double* out, int iters, int mode) {
    const bool dfma_warp = ((threadIdx.x >> 5) & 1) == 0;
    double s = 0.0;
    if (dfma_warp) {
        if (mode & 1) {
            double x[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = 1.0 + 1e-9 * (threadIdx.x + i);
            const double a = 1.0000001, c = 1e-7;
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, c);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) s += x[i];
        }
    } else if (mode & 2) {
        double acc[16][2];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
        const double af = 1.0 + 1e-9 * threadIdx.x, bf = 1.0 - 1e-9 * threadIdx.x;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma8x8x4(acc[i][0], acc[i][1], af, bf);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
    }
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA loop with `nmul` interleaved DMULs and `nlds` shared-memory loads per 16 DMMAs (pipe-interference probe)
__global__ void __launch_bounds__(256) dmma_mix_kernel(double* out, int iters, int nmul, int nlds) {
    __shared__ double sbuf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sbuf[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
    double af[4], bf = 1.0 - 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) af[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    double m0 = 1.0000001;
    int idx = threadIdx.x & 255;
    for (int it = 0; it < iters; ++it) {
        if (nlds == 8) {
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = sbuf[(idx + 32 * i + it) & 1023] * sbuf[(idx + 7 * i + 2 * it) & 1023];
        } else if (nmul == 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = af[i] * m0;
        } else if (nmul == 1) {
            af[0] = af[0] * m0;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma8x8x4(acc[i][0], acc[i][1], af[i & 3], bf);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace vela
