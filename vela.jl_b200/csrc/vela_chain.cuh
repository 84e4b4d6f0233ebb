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
    int delta;                 // the C_DELTA columns of the TOA table are filled and the launch wants residuals (emit != 2)
    int n_index_masks;         // rows of index_masks
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
    int rows_in_smem;          // 1: the launch carries blockDim.x * row doubles of dynamic shared memory for the working rows
    const double* tzr0;        // Double64 spin phase of the TZR TOA at the default parameters (kFlagDelta), or null
};

// K0 body: read_params (parameter.jl:158-169) + per-walker hoists + calc_tzr_phase (residuals.jl:29-32); one thread
// per walker.  Shared by the generic kernel and the per-model NVRTC build like chain_body below.
__device__ __forceinline__ void prologue_row(const ProgramDev& prog, const PrologueArgs& a, long b, double* P);
// Spindown with kFlagDelta: derived slots 2 (TZR phase minus its default-parameter value) and 3.. (dF_n); a no-op for every
// other component (folded away in the per-model build).
__device__ __forceinline__ void fill_delta_slots(const CompDev& c, const double* P, double* Ds, const double* defaults,
                                                 const double* tzr0, dd ph, bool bad) {
    if (c.kind != VELA_COMP_SPINDOWN || !(c.flags & kFlagDelta)) return;
    const int nf = c.len[0];
    const dd s0 = two_sum(defaults[c.par[0]], nf > 0 ? defaults[c.par[1]] : 0.0);
    const dd ds = dd_sub(dd{Ds[0], Ds[1]}, s0);
    Ds[3] = ds.hi + ds.lo;
    for (int n = 1; n < nf; ++n) Ds[3 + n] = P[c.par[1] + n] - defaults[c.par[1] + n];
    const dd dz = tzr0 ? dd_sub(ph, dd{tzr0[0], tzr0[1]}) : dd{CUDART_NAN, 0.0};
    Ds[2] = bad ? CUDART_NAN : dz.hi + dz.lo;
}
__device__ __forceinline__ void prologue_body(const ProgramDev& prog, const PrologueArgs& a) {
    long b = blockIdx.x * (long)blockDim.x + threadIdx.x;
#ifndef VELA_HOST_EMUL
    // The walker's extended row is written and then read back dozens of times along one dependent chain (derived values,
    // the TZR fold); in global memory every one of those reads is an L2 round trip (stores do not allocate in L1): the
    // kernel waited on them 8 cycles out of 9.  The working row lives in shared memory (odd pitch: conflict-free) and
    // goes out in one coalesced pass at the end.
    extern __shared__ double sRows[];
    if (a.rows_in_smem) {
        const long b0 = blockIdx.x * (long)blockDim.x;
        const int nrow = (int)min((long)blockDim.x, a.B - b0);
        if (b < a.B) prologue_row(prog, a, b, sRows + (size_t)threadIdx.x * prog.row);
        __syncthreads();
        for (int i = threadIdx.x; i < nrow * prog.row; i += blockDim.x) a.Pext[b0 * prog.row + i] = sRows[i];
        return;
    }
#endif
    if (b >= a.B) return;
    prologue_row(prog, a, b, a.Pext + b * prog.row);
}

__device__ __forceinline__ void prologue_row(const ProgramDev& prog, const PrologueArgs& a, long b, double* P) {
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
    e.tzr = true; e.wide = false; e.delta = false; e.d0[0] = e.d0[1] = e.d0[2] = 0.0; e.sun_regs = false;
    // the TZR block is laid out column-after-column with stride 1 (C_NCOL doubles)
    ToaRegs t = load_toa(a.tzr_block, 1, 0, false);
    Corr k = run_chain(prog, P, t, e);
    const dd ph = dd_add_d(k.phase, k.phase_small);
    a.tzr_phase[2 * b] = k.bad ? CUDART_NAN : ph.hi;
    a.tzr_phase[2 * b + 1] = ph.lo;
    // kFlagDelta: this walker's spin parameters and TZR phase as differences from the default-parameter values
#ifdef VELA_SPEC_CALLS
#define VELA_APPLY(i) fill_delta_slots(prog.comp[i], P, D + prog.comp[i].der, a.defaults, a.tzr0, ph, k.bad);
    VELA_SPEC_CALLS
#undef VELA_APPLY
#else
    for (int ic = 0; ic < prog.n_comp; ++ic) fill_delta_slots(prog.comp[ic], P, D + prog.comp[ic].der, a.defaults, a.tzr0, ph, k.bad);
#endif
}

// frexp for the log-det accumulation: normal positive inputs (every sane err^2) are split with three integer
// operations; zero, subnormal, infinite and NaN inputs go through the library routine.
__device__ __forceinline__ double split_mantissa(double v, int* ex, bool spec, bool& redo) {
    const int hi = __double2hiint(v);
    if (spec) {   // speculative evaluation (Corr::spec): the integer split unconditionally, the range test deferred
        redo |= !((unsigned)(hi - 0x00100000) < 0x7fe00000u);
        *ex = (hi >> 20) - 1022;
        return __hiloint2double((hi & 0x000fffff) | 0x3fe00000, __double2loint(v));
    }
    // positive and normal <=> sign 0 and exponent field in 1 .. 0x7fe: one unsigned compare on the high word
    if ((unsigned)(hi - 0x00100000) < 0x7fe00000u) {
        *ex = (hi >> 20) - 1022;
        return __hiloint2double((hi & 0x000fffff) | 0x3fe00000, __double2loint(v));   // mantissa in [1/2, 1)
    }
    return frexp(v, ex);
}
__device__ __forceinline__ double split_mantissa(double v, int* ex) {
    bool unused = false;
    return split_mantissa(v, ex, false, unused);
}

// What one (TOA, walker) pair contributes: form_residual (residuals.jl:8-12, wideband_residuals.jl:6-17), the scaled
// errors (toa.jl:122-123, wideband_toa.jl:68-69) and the terms of chi^2 / log det N.  Shared by the chain kernel and
// the host-side arithmetic check (tests/host_emul).
struct ResidTerms {
    double tres, w, dmres, wdm;   // residual [s], 1 / sigma^2 (emit != 0 only), DM residual and its weight (wideband)
    double chi;                   // r^2 / sigma^2 (+ DM term); NaN for a failed chain
    double lman; int lex;         // prod sigma^2 = lman 2^lex
    dd dphase;                    // phase mode (emit == 2) only
    double f;                     // doppler-shifted spin frequency
    bool redo;                    // speculative evaluation: a gate failed (in the chain or here), evaluate again with branches
};

__device__ __forceinline__ ResidTerms residual_terms(const Corr& k, const ToaRegs& t, bool wide, double tzr_hi, double tzr_lo,
                                                     int emit) {
    ResidTerms r;
    // Float64((phase - pulse_number) - tzrphase) / (f (1 + doppler)).
    // phase and tzrphase are ~1e10 cycles and differ by the pulse number to within a fraction of a cycle: their
    // leading words are subtracted error-free, the pulse number (an integer) then cancels exactly, and what is
    // left - trailing words and the small phase terms - is of the size of the residual itself, so plain FP64
    // keeps it to 1e-16 relative.  12 operations where the Double64 route takes 40.
    double tres_phase;
    r.dphase = dd{0.0, 0.0};
    if (emit == 2) {   // phase mode wants the Double64 value itself
        dd dphase = dd_add_d(k.phase, k.phase_small);
        dphase = t.pn_lo == 0.0 ? dd_add_d(dphase, -t.pn_hi) : dd_sub(dphase, dd{t.pn_hi, t.pn_lo});
        dphase = dd_sub(dphase, dd{tzr_hi, tzr_lo});
        r.dphase = dphase;
        tres_phase = dphase.hi;
    } else if (k.delta) {   // kFlagDelta: (phase - pulse number) - TZR phase arrived as one residual-sized double
        tres_phase = k.dph + k.phase_small;
    } else {
        const dd lead = two_sum(k.phase.hi, -tzr_hi);
        const double tail = ((k.phase.lo - tzr_lo) + lead.lo) + (k.phase_small - t.pn_lo);
        tres_phase = (lead.hi - t.pn_hi) + tail;
    }
    if (k.fast) {   // the quotient is residual-sized (1 ulp = 1e-22 s): a reciprocal and a product do
        r.f = __fma_rn(k.spin_frequency, k.doppler, k.spin_frequency);
        r.tres = tres_phase * rcp_fast(r.f);
    } else {
        r.f = k.spin_frequency * (1 + k.doppler);
        r.tres = div_fast(tres_phase, r.f);
    }
    const double err2 = (t.err2 + k.equad2) * k.efac * k.efac;  // scaled_toa_error_sqr toa.jl:122-123
    // log det N = sum_j log(err2_j) is accumulated as a PRODUCT of mantissas and a sum of binary exponents, so the
    // logarithm is taken once per (tile, walker) in the block epilogue instead of once per (TOA, walker)
    r.w = 0.0;
    r.redo = k.redo;
    r.lman = split_mantissa(err2, &r.lex, k.spec, r.redo);
    if (emit) {  // Woodbury: y^T N^-1 y and log det N from Ninvdiag, gls_likelihood.jl:81-99
        r.w = rcp_fast(err2);
        r.chi = r.tres * r.tres * r.w;
    } else {
        r.chi = div_fast(r.tres * r.tres, err2);
    }
    r.dmres = 0.0; r.wdm = 0.0;
    if (wide) {  // wideband_residuals.jl:6-17, wideband_toa.jl:68-69
        r.dmres = t.dm - k.model_dm;
        const double dmerr2 = (t.dmerr2 + k.dmequad2) * k.dmefac * k.dmefac;
        int dex;
        r.lman *= split_mantissa(dmerr2, &dex, k.spec, r.redo);
        r.lex += dex;
        if (emit) {
            r.wdm = rcp_fast(dmerr2);
            r.chi += r.dmres * r.dmres * r.wdm;
        } else {
            r.chi += div_fast(r.dmres * r.dmres, dmerr2);
        }
    }
    if (k.bad || k.spin_frequency == 0.0) r.chi = CUDART_NAN;
    return r;
}

// create(): the C_DELTA columns of one TOA (kFlagDelta) from the chain at the DEFAULT parameters - P is the extended row
// the prologue made of them.  tzr0: Double64 spin phase of the TZR TOA at those parameters (every thread computes it;
// the caller keeps one copy for the prologue).  Same code as the per-walker evaluation, so delay0 is the very double a
// walker sitting at the defaults would carry.
__device__ inline void default_phase_point(const ProgramDev& prog, const double* P, const double* toa, long stride, long j,
                                           const double* tzr_block, const uint32_t* index_masks, const uint8_t* bit_masks,
                                           long n_toas, double* out3, double* tzr0) {
    ChainEnv e;
    e.index_masks = index_masks; e.bit_masks = bit_masks; e.n_toas = n_toas; e.wide = false; e.delta = false;
    e.d0[0] = e.d0[1] = e.d0[2] = 0.0; e.sun_regs = false;
    e.toa_base = tzr_block; e.stride = 1; e.j = 0; e.tzr = true;
    const Corr kz = run_chain(prog, P, load_toa(tzr_block, 1, 0, false), e);
    tzr0[0] = kz.phase.hi; tzr0[1] = kz.phase.lo;
    e.toa_base = toa; e.stride = stride; e.j = j; e.tzr = false; e.wide = prog.wideband != 0;
    const ToaRegs t = load_toa(toa, stride, j, false);
    const Corr k = run_chain(prog, P, t, e);
    const dd dt0 = dd_add_d(dd{t.t_hi, t.t_lo}, -k.delay);
    const dd r0 = dd_sub(dd_sub(k.phase, dd{t.pn_hi, t.pn_lo}), kz.phase);
    const bool bad = k.bad || kz.bad;
    out3[0] = bad ? CUDART_NAN : k.delay;
    out3[1] = dt0.hi;
    out3[2] = r0.hi + r0.lo;
}

#ifndef VELA_HOST_EMUL   // the kernel body proper: warp shuffles and shared memory have no host stand-in
// Walkers whose (chi^2, mantissa product, exponent sum) terms wait in the per-warp ring before one pass reduces them
// over the 32 TOAs of the warp.
constexpr int kRedRing = 8;
constexpr int kRedPitch = 33;   // doubles / ints per ring row: conflict-free for the column-wise pass
// shared doubles one warp needs for its ring: two double rows and one int row per slot
constexpr int kRedRingDoubles = kRedRing * kRedPitch * 2 + (kRedRing * kRedPitch + 1) / 2;

__host__ __device__ inline size_t chain_smem_doubles(int group, int row, int warps) {
    return (size_t)group * row + 2 * (size_t)group + (size_t)warps * group * 3 + (size_t)warps * kRedRingDoubles;
}

// Reduce `cnt` ring slots (walkers s0 .. s0 + cnt - 1 of the block's group) over the warp's 32 TOAs.  Lane l sums
// eight consecutive TOAs of slot l & 7; two shuffle steps join the four parts; the lanes of part 0 store the warp's
// partial for their walker.  25 shuffles + 10 FP64 operations per walker in the butterfly this replaces; here 1.25 + 2.25.
__device__ __forceinline__ void ring_reduce(const double* rChi, const double* rMan, const int* rExp, int cnt, int lane,
                                            double* out, int s0) {
    __syncwarp();
    const int slot = lane & (kRedRing - 1), part = lane >> 3;
    double chi = 0.0, man = 1.0;
    int ex = 0;
    if (slot < cnt) {
        const double* pc = rChi + slot * kRedPitch + part * 8;
        const double* pm = rMan + slot * kRedPitch + part * 8;
        const int* pe = rExp + slot * kRedPitch + part * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) { chi += pc[i]; man *= pm[i]; ex += pe[i]; }   // 8 mantissas in [1/2, 1): >= 2^-8
    }
#pragma unroll
    for (int off = 8; off <= 16; off <<= 1) {
        chi += __shfl_xor_sync(0xffffffffu, chi, off);
        man *= __shfl_xor_sync(0xffffffffu, man, off);    // 32 mantissas: >= 2^-32, no underflow
        ex += __shfl_xor_sync(0xffffffffu, ex, off);
    }
    if (part == 0 && slot < cnt) {
        double* r = out + (size_t)(s0 + slot) * 3;
        r[0] = chi; r[1] = man; r[2] = (double)ex;
    }
    __syncwarp();
}

#ifndef VELA_CHAIN_UNROLL
#define VELA_CHAIN_UNROLL 1
#endif
constexpr int kChainUnroll = VELA_CHAIN_UNROLL;   // walkers in flight per thread (per-model build: VELA_B200_JIT_UNROLL)

__device__ __forceinline__ void chain_body(const ProgramDev& prog, const ChainArgs& a) {
    extern __shared__ double smem[];
    const int row = prog.row;
    const int nwarps = blockDim.x >> 5;
    double* sP = smem;                                  // [group][row]
    double* sTzr = sP + (size_t)a.group * row;          // [group][2]
    double* sRed = sTzr + 2 * a.group;                  // [warps][group][3]: chi^2, mantissa product, exponent sum
    double* sRing = sRed + (size_t)nwarps * a.group * 3;   // [warps][kRedRingDoubles]

    const long b0 = (long)blockIdx.y * a.group;
    const int nb = (int)min((long)a.group, a.B - b0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* rChi = sRing + (size_t)warp * kRedRingDoubles;
    double* rMan = rChi + kRedRing * kRedPitch;
    int* rExp = reinterpret_cast<int*>(rMan + kRedRing * kRedPitch);
    double* wRed = sRed + (size_t)warp * a.group * 3;

    {   // the group's parameter rows -> shared memory, 16 bytes per load and eight loads in flight per thread (one element per
        // iteration left the block waiting on a dozen dependent round trips: 9 % of the kernel's stall samples); b0 is a
        // multiple of the group size, so b0 * row doubles is a multiple of 16 bytes
        const int n = nb * row, n2 = n >> 1;
        const double2* src = reinterpret_cast<const double2*>(a.Pext + b0 * row);
        double2* dst = reinterpret_cast<double2*>(sP);
#pragma unroll 8
        for (int i = tid; i < n2; i += blockDim.x) dst[i] = src[i];
        if ((n & 1) && tid == 0) sP[n - 1] = a.Pext[b0 * row + n - 1];
    }
    for (int i = tid; i < 2 * nb; i += blockDim.x) sTzr[i] = a.tzr_phase[2 * b0 + i];

    const long j = (long)blockIdx.x * blockDim.x + tid;
    const bool active = j < a.n_toas;
    const long jj = active ? j : a.n_toas - 1;
    ToaRegs t = load_toa(a.toa, a.toa_stride, jj, false);
    ChainEnv e;
    e.toa_base = a.toa; e.stride = a.toa_stride; e.j = jj;
    e.index_masks = a.index_masks; e.bit_masks = a.bit_masks; e.n_toas = a.n_toas;
    e.tzr = false; e.wide = prog.wideband != 0;
#ifdef VELA_SPEC_DELTA   // per-model build: whether the handle's C_DELTA columns are filled and this emit mode takes residuals is a
    e.delta = VELA_SPEC_DELTA != 0;   // compile-time fact
#else
    e.delta = a.delta != 0;
#endif
    // the three C_DELTA words ride in registers through the walker loop (as loads inside it they missed L1 - five CTAs'
    // parameter rows leave it ~40 KB - and sat on the critical path: 0.95 ms against 0.87 without the route)
#pragma unroll
    for (int c = 0; c < 3; ++c) e.d0[c] = e.delta ? __ldg(a.toa + (size_t)(C_DELTA + c) * a.toa_stride + jj) : 0.0;
    // so do the Sun's Shapiro expansion point and the first four index masks (EFAC / EQUAD groups, exclusive jumps ...):
    // TOA-only words that every walker step used to fetch again (config 3 chain 0.775 -> 0.734 ms)
    e.sun_regs = true;
#pragma unroll
    for (int c = 0; c < 3; ++c) e.sh0[c] = __ldg(a.toa + (size_t)(C_SHAP + c) * a.toa_stride + jj);
#pragma unroll
    for (int m = 0; m < 4; ++m) e.mreg[m] = m < a.n_index_masks ? __ldg(a.index_masks + (long)m * a.n_toas + jj) : 0u;
    __syncthreads();

    // the launch's emit mode; the per-model build knows it at compile time (VELA_SPEC_EMIT), so the mode tests fold away
#ifdef VELA_SPEC_EMIT
    constexpr int emit = VELA_SPEC_EMIT;
#else
    const int emit = a.emit;
#endif
    // running pointers instead of per-walker index arithmetic (64-bit multiplies in the innermost loop otherwise)
    const double* P = sP;
    const double* tz = sTzr;
    double* rc = rChi + lane;
    double* rm = rMan + lane;
    int* re = rExp + lane;
    // what one evaluated (TOA, walker) pair leaves behind: its rows of Y / W (/ F) and its terms in ring slot `slot`
    auto deposit = [&](const ResidTerms& r, size_t off, int slot) {   // off: (b0 + s) ld_yw + j
        double* yp = a.Y + off;
        double* wp = a.W + off;
        if (emit == 2) {  // phase mode: Double64 phase residual and doppler-shifted spin frequency
            if (active) {
                *yp = r.dphase.hi;
                *wp = r.dphase.lo;
                a.F[off] = r.f;
            }
        } else if (emit && active) {
            *yp = emit == 3 ? r.w * r.tres : r.tres;       // emit 3: the structured contraction wants w.y
            *wp = r.w;
            if (e.wide) {
                yp[a.n_toas] = emit == 3 ? r.wdm * r.dmres : r.dmres;
                wp[a.n_toas] = r.wdm;
            }
        }
        rc[slot * kRedPitch] = active ? r.chi : 0.0;
        rm[slot * kRedPitch] = active ? r.lman : 1.0;   // wideband: a product of two mantissas, >= 1/4
        re[slot * kRedPitch] = active ? r.lex : 0;
    };
#ifdef VELA_SPEC_SPECULATE
    // Per-model fast build: every walker is first evaluated straight-line (Corr::spec); walkers for which some lane of the
    // warp fell outside a gate are noted in a warp-uniform mask and evaluated again, through the branching code, in a
    // second loop after this one (rare: prior draws far from the default sky position, barycentred TOAs).  The hot loop
    // holds no trace of the branching code: it is one basic block plus the ring pass.
    unsigned long long redo_mask = 0;
#endif
    size_t off = (size_t)b0 * a.ld_yw + j;   // running offset instead of per-walker index arithmetic (64-bit multiplies otherwise)
#pragma unroll kChainUnroll
    for (int s = 0; s < nb; ++s, P += row, tz += 2, off += a.ld_yw) {
#ifdef VELA_SPEC_SPECULATE
        const Corr k = run_chain(prog, P, t, e, true);
        const ResidTerms r = residual_terms(k, t, e.wide, tz[0], tz[1], emit);
        if (__any_sync(0xffffffffu, r.redo)) redo_mask |= 1ULL << s;
#else
        const Corr k = run_chain(prog, P, t, e);
        const ResidTerms r = residual_terms(k, t, e.wide, tz[0], tz[1], emit);
#endif
        // this walker's terms wait in the warp's ring; every kRedRing walkers one pass reduces them over the TOAs
        const int slot = s & (kRedRing - 1);
        deposit(r, off, slot);
        if (slot == kRedRing - 1) ring_reduce(rChi, rMan, rExp, kRedRing, lane, wRed, s - (kRedRing - 1));
    }
    if (nb & (kRedRing - 1)) ring_reduce(rChi, rMan, rExp, nb & (kRedRing - 1), lane, wRed, nb & ~(kRedRing - 1));
#ifdef VELA_SPEC_SPECULATE
    while (redo_mask) {
        const int s = __ffsll((long long)redo_mask) - 1;
        redo_mask &= redo_mask - 1;
        ToaRegs tr = t;
#if defined(VELA_SPEC_DELTA) && VELA_SPEC_DELTA   // the hot loop no longer reads the words of the double-double route: fetched here, they cost it no register
        tr.t_lo = __ldg(a.toa + (size_t)C_T_LO * a.toa_stride + jj);
        tr.pn_hi = __ldg(a.toa + (size_t)C_PN_HI * a.toa_stride + jj);
        tr.pn_lo = __ldg(a.toa + (size_t)C_PN_LO * a.toa_stride + jj);
#endif
        const Corr k = run_chain(prog, sP + (size_t)s * row, tr, e, false);
        const ResidTerms r = residual_terms(k, tr, e.wide, sTzr[2 * s], sTzr[2 * s + 1], emit);
        deposit(r, (size_t)(b0 + s) * a.ld_yw + j, 0);
        ring_reduce(rChi, rMan, rExp, 1, lane, wRed, s);
    }
#endif
    __syncthreads();
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

#endif  // !VELA_HOST_EMUL

}  // namespace vela
