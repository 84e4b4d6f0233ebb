// vela_chain.cuh - body of the per-(TOA, walker) chain kernel (K1), shared by the generic kernel in
// vela_kernels.cu (program passed as a __grid_constant__ parameter and interpreted) and by the per-model
// specialised kernel that vela_api.cu compiles at create() time with NVRTC (program baked in as a constant, so
// the component switch and every parameter offset fold away).  Same source, same flags, same arithmetic.
#pragma once
#include "vela_device.cuh"

namespace vela {

constexpr int kChainThreads = 256;  // max threads per block of the chain kernel (one TOA each)

struct ChainArgs {
    const double* Pext;        // [B][row]
    const double* tzr_phase;   // [B][2]
    const double* toa;         // SoA table
    long toa_stride;
    long n_toas;
    const uint32_t* index_masks;
    const uint8_t* bit_masks;
    long B;
    int group;                 // walkers per block
    int emit;                  // 1: write y, w (Woodbury path); 2: write phase (hi, lo) and spin frequency; 3: write w.y, w
    double* Y;                 // [B_pad][ld_yw]
    double* W;
    double* F;                 // emit == 2 only
    long ld_yw;
    double* partial;           // [tiles][B][2]
};

struct PrologueArgs {
    const double* defaults;    // [np + n_aux]: default parameter values, then the components' constant tables
    const int* free_idx;       // [nd] 0-based positions of the free parameters
    const double* params;      // [B][nd] samples
    long B;
    const double* tzr_block;   // the TZR TOA as C_NCOL columns with stride 1
    const uint32_t* index_masks;
    const uint8_t* bit_masks;
    long n_toas;
    double* Pext;              // [B][row]
    double* tzr_phase;         // [B][2]
};

// K0 body: read_params (parameter.jl:158-169) + per-walker hoists + calc_tzr_phase (residuals.jl:29-32); one thread
// per walker.  Shared by the generic kernel and the per-model NVRTC build like chain_body below.
__device__ __forceinline__ void prologue_body(const ProgramDev& prog, const PrologueArgs& a) {
    long b = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    double* P = a.Pext + b * prog.row;
    for (int i = 0; i < prog.n_params; ++i) P[i] = a.defaults[i];
    for (int i = 0; i < prog.n_free; ++i) P[a.free_idx[i]] = a.params[b * prog.n_free + i];
    double* D = P + prog.n_params;
    const double* aux = a.defaults + prog.n_params;
#ifdef VELA_SPEC_CALLS
#define VELA_APPLY(i) fill_derived(prog.comp[i], P, D + prog.comp[i].der, aux);
    VELA_SPEC_CALLS
#undef VELA_APPLY
#else
    for (int ic = 0; ic < prog.n_comp; ++ic) fill_derived(prog.comp[ic], P, D + prog.comp[ic].der, aux);
#endif
    ChainEnv e;
    e.toa_base = a.tzr_block; e.stride = 1; e.j = 0;
    e.index_masks = a.index_masks; e.bit_masks = a.bit_masks; e.n_toas = a.n_toas;
    e.tzr = true; e.wide = false;
    // the TZR block is laid out column-after-column with stride 1 (C_NCOL doubles)
    ToaRegs t = load_toa(a.tzr_block, 1, 0, false);
    Corr k = run_chain(prog, P, t, e);
    a.tzr_phase[2 * b] = k.bad ? CUDART_NAN : k.phase.hi;
    a.tzr_phase[2 * b + 1] = k.phase.lo;
}

// frexp for the log-det accumulation: normal positive inputs (every sane err^2) are split with three integer
// operations; zero, subnormal, infinite and NaN inputs go through the library routine.
__device__ __forceinline__ double split_mantissa(double v, int* ex) {
    const long long bits = __double_as_longlong(v);
    const int field = (int)((bits >> 52) & 0x7ff);
    if (bits > 0 && field != 0 && field != 0x7ff) {
        *ex = field - 1022;
        return __longlong_as_double((bits & 0x800fffffffffffffLL) | 0x3fe0000000000000LL);   // mantissa in [1/2, 1)
    }
    return frexp(v, ex);
}

__device__ __forceinline__ void chain_body(const ProgramDev& prog, const ChainArgs& a) {
    extern __shared__ double smem[];
    const int row = prog.row;
    double* sP = smem;                                  // [group][row]
    double* sTzr = sP + (size_t)a.group * row;          // [group][2]
    double* sRed = sTzr + 2 * a.group;                  // [warps][group][3]: chi^2, mantissa product, exponent sum

    const long b0 = (long)blockIdx.y * a.group;
    const int nb = (int)min((long)a.group, a.B - b0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < nb * row; i += blockDim.x) sP[i] = a.Pext[b0 * row + i];
    for (int i = tid; i < 2 * nb; i += blockDim.x) sTzr[i] = a.tzr_phase[2 * b0 + i];

    const long j = (long)blockIdx.x * blockDim.x + tid;
    const bool active = j < a.n_toas;
    const long jj = active ? j : a.n_toas - 1;
    ToaRegs t = load_toa(a.toa, a.toa_stride, jj, false);
    ChainEnv e;
    e.toa_base = a.toa; e.stride = a.toa_stride; e.j = jj;
    e.index_masks = a.index_masks; e.bit_masks = a.bit_masks; e.n_toas = a.n_toas;
    e.tzr = false; e.wide = prog.wideband != 0;
    __syncthreads();

    for (int s = 0; s < nb; ++s) {
        const double* P = sP + (size_t)s * row;
        Corr k = run_chain(prog, P, t, e);
        // form_residual residuals.jl:8-12: Float64((phase - pulse_number) - tzrphase) / (f (1 + doppler))
        // (pulse numbers are integers below 2^53, so pn_lo is normally 0: dd - double is the same bits for 11 adds less)
        dd dphase = t.pn_lo == 0.0 ? dd_add_d(k.phase, -t.pn_hi) : dd_sub(k.phase, dd{t.pn_hi, t.pn_lo});
        dphase = dd_sub(dphase, dd{sTzr[2 * s], sTzr[2 * s + 1]});
        double f = k.spin_frequency * (1 + k.doppler);
        double tres = dphase.hi / f;
        double err2 = (t.err2 + k.equad2) * k.efac * k.efac;  // scaled_toa_error_sqr toa.jl:122-123
        // log det N = sum_j log(err2_j) is accumulated as a PRODUCT of mantissas and a sum of binary exponents, so the
        // logarithm is taken once per (tile, walker) in the block epilogue instead of once per (TOA, walker)
        double chi, w = 0.0;
        int lex;
        double lman = split_mantissa(err2, &lex);
        if (a.emit) {  // Woodbury: y^T N^-1 y and log det N from Ninvdiag, gls_likelihood.jl:81-99
            w = 1.0 / err2;
            chi = tres * tres * w;
        } else {
            chi = tres * tres / err2;
        }
        double dmres = 0.0, wdm = 0.0;
        if (e.wide) {  // wideband_residuals.jl:6-17, wideband_toa.jl:68-69
            dmres = t.dm - k.model_dm;
            double dmerr2 = (t.dmerr2 + k.dmequad2) * k.dmefac * k.dmefac;
            int dex;
            lman *= split_mantissa(dmerr2, &dex);
            lex += dex;
            if (a.emit) {
                wdm = 1.0 / dmerr2;
                chi += dmres * dmres * wdm;
            } else {
                chi += dmres * dmres / dmerr2;
            }
        }
        if (k.bad || k.spin_frequency == 0.0) chi = CUDART_NAN;
        if (!active) { chi = 0.0; lman = 1.0; lex = 0; }
        if (a.emit == 2 && active) {  // phase mode: Double64 phase residual and doppler-shifted spin frequency
            const size_t o = (size_t)(b0 + s) * a.ld_yw + j;
            a.Y[o] = dphase.hi;
            a.W[o] = dphase.lo;
            a.F[o] = f;
        } else if (a.emit && active) {
            const size_t o = (size_t)(b0 + s) * a.ld_yw + j;
            a.Y[o] = a.emit == 3 ? w * tres : tres;       // emit 3: the structured contraction wants w.y
            a.W[o] = w;
            if (e.wide) {
                a.Y[o + a.n_toas] = a.emit == 3 ? wdm * dmres : dmres;
                a.W[o + a.n_toas] = wdm;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            chi += __shfl_xor_sync(0xffffffffu, chi, off);
            lman *= __shfl_xor_sync(0xffffffffu, lman, off);    // 32 mantissas in [1/4, 1): no underflow
            lex += __shfl_xor_sync(0xffffffffu, lex, off);
        }
        if (lane == 0) {
            double* r = sRed + ((size_t)warp * a.group + s) * 3;
            r[0] = chi; r[1] = lman; r[2] = (double)lex;
        }
    }
    __syncthreads();
    const int nwarps = blockDim.x >> 5;
    for (int i = tid; i < nb; i += blockDim.x) {
        double chi = 0.0, man = 1.0, ex = 0.0;
        for (int w = 0; w < nwarps; ++w) {
            const double* r = sRed + ((size_t)w * a.group + i) * 3;
            chi += r[0]; man *= r[1]; ex += r[2];
        }
        double* out = a.partial + ((size_t)blockIdx.x * a.B + b0 + i) * 2;
        out[0] = chi;
        out[1] = log(man) + ex * 0.6931471805599453;   // ln 2
    }
}

}  // namespace vela
