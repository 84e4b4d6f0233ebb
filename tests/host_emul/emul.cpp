// emul.cpp - host-side arithmetic check of the chain kernel (TEST INFRASTRUCTURE ONLY).
//
// Compiles the product's own device headers (vela_device.cuh, vela_chain.cuh) and host preparation (vela_host_prep.h)
// with g++ through cuda_shim.h, and runs, one "thread" at a time, exactly what the CUDA kernels run per walker
// (prologue_body) and per (TOA, walker) (run_chain + residual_terms).  What it cannot cover is the kernel's
// parallel plumbing (shared memory, shuffles, grid mapping) - the `-m gpu` tests do that.  Nothing in the product
// links or loads this file.
#define VELA_HOST_EMUL 1
#include "../../vela.jl_b200/csrc/vela_chain.cuh"
#include "../../vela.jl_b200/csrc/vela_host_prep.h"

#include <vector>

using namespace vela;

// kFlagDelta (vela_device.cuh): fills the C_DELTA columns of the SoA table and the TZR spin phase at the default parameters
// the way create() does on the device - prologue for the default vector, then default_phase_point per TOA.
static bool prepare_delta(const VelaModelDesc* desc, const ProgramDev& pg, std::vector<double>& soa, long stride,
                          const std::vector<double>& tzr, const std::vector<double>& defaults, const std::vector<int>& free_idx,
                          double* tzr0) {
    bool want = false;
    for (int ic = 0; ic < pg.n_comp; ++ic) want = want || (pg.comp[ic].flags & kFlagDelta);
    if (!want) return false;
    const long nt = desc->n_toas;
    std::vector<double> x0((size_t)std::max(desc->n_free, 1), 0.0), P0((size_t)pg.row, 0.0), tz(2, 0.0);
    for (int i = 0; i < desc->n_free; ++i) x0[i] = desc->default_values[desc->free_indices[i]];
    PrologueArgs pa;
    pa.defaults = defaults.data(); pa.free_idx = free_idx.data(); pa.params = x0.data(); pa.B = 1;
    pa.tzr_block = tzr.data(); pa.index_masks = desc->index_masks; pa.bit_masks = desc->bit_masks; pa.n_toas = nt;
    pa.Pext = P0.data(); pa.tzr_phase = tz.data(); pa.rows_in_smem = 0; pa.tzr0 = nullptr;
    blockIdx = {0, 0, 0}; blockDim = {1, 1, 1}; threadIdx = {0, 0, 0};
    prologue_body(pg, pa);
    for (long j = 0; j < nt; ++j) {
        double out3[3];
        default_phase_point(pg, P0.data(), soa.data(), stride, j, tzr.data(), desc->index_masks, desc->bit_masks, nt, out3, tzr0);
        for (int c = 0; c < 3; ++c) soa[(size_t)(C_DELTA + c) * stride + j] = out3[c];
    }
    return true;
}

extern "C" __attribute__((visibility("default")))
int emul_chain(const VelaModelDesc* desc, const double* params, long B, int emit, int fast, double* Y, double* W, double* F,
               double* chi2, double* logdetN) {
    ProgramDev pg;
    build_program(desc, pg, fast != 0);   // fast: the fast-arithmetic build (kFlagFast), else the strict one
    const long nt = desc->n_toas, stride = (nt + 31) / 32 * 32;
    const long n_rows = desc->wideband ? 2 * nt : nt;
    std::vector<double> soa((size_t)C_NCOL * stride, 0.0), tzr(C_NCOL, 0.0);
    for (long j = 0; j < nt; ++j) {
        fill_toa_columns((const double*)((const char*)desc->toas + j * (long)desc->toa_stride_bytes), desc->wideband,
                         soa.data(), stride, j);
        fill_shapiro_columns(desc, soa.data(), stride, j);
    }
    fill_toa_columns((const double*)desc->tzr_toa, 0, tzr.data(), 1, 0);
    fill_shapiro_columns(desc, tzr.data(), 1, 0);
    std::vector<double> defaults(desc->default_values, desc->default_values + desc->n_params);
    if (desc->n_aux > 0) defaults.insert(defaults.end(), desc->aux_values, desc->aux_values + desc->n_aux);
    std::vector<int> free_idx(desc->free_indices, desc->free_indices + desc->n_free);
    std::vector<double> Pext((size_t)B * pg.row, 0.0), tzr_phase((size_t)B * 2, 0.0);
    double tzr0[2] = {0.0, 0.0};
    const bool delta = prepare_delta(desc, pg, soa, stride, tzr, defaults, free_idx, tzr0);

    PrologueArgs pa;
    pa.defaults = defaults.data(); pa.free_idx = free_idx.data(); pa.params = params; pa.B = B;
    pa.tzr_block = tzr.data(); pa.index_masks = desc->index_masks; pa.bit_masks = desc->bit_masks; pa.n_toas = nt;
    pa.Pext = Pext.data(); pa.tzr_phase = tzr_phase.data(); pa.rows_in_smem = 0; pa.tzr0 = delta ? tzr0 : nullptr;
    blockIdx = {0, 0, 0}; blockDim = {1, 1, 1};
    for (long b = 0; b < B; ++b) {
        threadIdx = {(unsigned)b, 0, 0};
        prologue_body(pg, pa);
    }
    for (long b = 0; b < B; ++b) {
        const double* P = Pext.data() + (size_t)b * pg.row;
        double chi = 0.0, lg = 0.0;
        for (long j = 0; j < nt; ++j) {
            ToaRegs t = load_toa(soa.data(), stride, j, false);
            ChainEnv e;
            e.toa_base = soa.data(); e.stride = stride; e.j = j;
            e.index_masks = desc->index_masks; e.bit_masks = desc->bit_masks; e.n_toas = nt;
            e.tzr = false; e.wide = desc->wideband != 0; e.delta = delta && emit != 2; e.sun_regs = false;
            for (int c = 0; c < 3; ++c) e.d0[c] = e.delta ? soa[(size_t)(C_DELTA + c) * stride + j] : 0.0;
            Corr k = run_chain(pg, P, t, e);
            ResidTerms r = residual_terms(k, t, e.wide, tzr_phase[2 * b], tzr_phase[2 * b + 1], emit);
            chi += r.chi;
            lg += std::log(r.lman) + r.lex * 0.6931471805599453;
            const size_t o = (size_t)b * n_rows + j;
            if (emit == 2) {
                if (Y) Y[o] = r.dphase.hi;
                if (W) W[o] = r.dphase.lo;
                if (F) F[o] = r.f;
            } else {
                if (Y) Y[o] = r.tres;
                if (W) W[o] = emit ? r.w : 0.0;
                if (e.wide) {
                    if (Y) Y[o + nt] = r.dmres;
                    if (W) W[o + nt] = emit ? r.wdm : 0.0;
                }
            }
        }
        if (chi2) chi2[b] = chi;
        if (logdetN) logdetN[b] = lg;
    }
    return 0;
}

// Speculative evaluation (Corr::spec, the per-model fast build's hot loop) against the branching evaluation of the same
// (TOA, walker) pairs: wherever the speculative pass raises no `redo`, every output must be the same bits; where it does,
// the kernel evaluates again with branches, so nothing is to be compared.  counts = {pairs, redo raised, mismatches}.
extern "C" __attribute__((visibility("default")))
int emul_spec_check(const VelaModelDesc* desc, const double* params, long B, int emit, long* counts) {
    ProgramDev pg;
    build_program(desc, pg, true);
    const long nt = desc->n_toas, stride = (nt + 31) / 32 * 32;
    std::vector<double> soa((size_t)C_NCOL * stride, 0.0), tzr(C_NCOL, 0.0);
    for (long j = 0; j < nt; ++j) {
        fill_toa_columns((const double*)((const char*)desc->toas + j * (long)desc->toa_stride_bytes), desc->wideband,
                         soa.data(), stride, j);
        fill_shapiro_columns(desc, soa.data(), stride, j);
    }
    fill_toa_columns((const double*)desc->tzr_toa, 0, tzr.data(), 1, 0);
    fill_shapiro_columns(desc, tzr.data(), 1, 0);
    std::vector<double> defaults(desc->default_values, desc->default_values + desc->n_params);
    if (desc->n_aux > 0) defaults.insert(defaults.end(), desc->aux_values, desc->aux_values + desc->n_aux);
    std::vector<int> free_idx(desc->free_indices, desc->free_indices + desc->n_free);
    std::vector<double> Pext((size_t)B * pg.row, 0.0), tzr_phase((size_t)B * 2, 0.0);
    double tzr0[2] = {0.0, 0.0};
    const bool delta = prepare_delta(desc, pg, soa, stride, tzr, defaults, free_idx, tzr0);
    PrologueArgs pa;
    pa.defaults = defaults.data(); pa.free_idx = free_idx.data(); pa.params = params; pa.B = B;
    pa.tzr_block = tzr.data(); pa.index_masks = desc->index_masks; pa.bit_masks = desc->bit_masks; pa.n_toas = nt;
    pa.Pext = Pext.data(); pa.tzr_phase = tzr_phase.data(); pa.rows_in_smem = 0; pa.tzr0 = delta ? tzr0 : nullptr;
    blockIdx = {0, 0, 0}; blockDim = {1, 1, 1};
    for (long b = 0; b < B; ++b) {
        threadIdx = {(unsigned)b, 0, 0};
        prologue_body(pg, pa);
    }
    counts[0] = counts[1] = counts[2] = 0;
    auto same = [](double x, double y) { return memcmp(&x, &y, 8) == 0 || (x != x && y != y); };
    for (long b = 0; b < B; ++b) {
        const double* P = Pext.data() + (size_t)b * pg.row;
        for (long j = 0; j < nt; ++j) {
            ToaRegs t = load_toa(soa.data(), stride, j, false);
            ChainEnv e;
            e.toa_base = soa.data(); e.stride = stride; e.j = j;
            e.index_masks = desc->index_masks; e.bit_masks = desc->bit_masks; e.n_toas = nt;
            e.tzr = false; e.wide = desc->wideband != 0; e.delta = delta && emit != 2; e.sun_regs = false;
            for (int c = 0; c < 3; ++c) e.d0[c] = e.delta ? soa[(size_t)(C_DELTA + c) * stride + j] : 0.0;
            Corr ks = run_chain(pg, P, t, e, true);
            ResidTerms rs = residual_terms(ks, t, e.wide, tzr_phase[2 * b], tzr_phase[2 * b + 1], emit);
            Corr kb = run_chain(pg, P, t, e, false);
            ResidTerms rb = residual_terms(kb, t, e.wide, tzr_phase[2 * b], tzr_phase[2 * b + 1], emit);
            ++counts[0];
            if (rs.redo) { ++counts[1]; continue; }
            if (!(same(rs.tres, rb.tres) && same(rs.w, rb.w) && same(rs.chi, rb.chi) && same(rs.lman, rb.lman) && rs.lex == rb.lex &&
                  same(rs.dmres, rb.dmres) && same(rs.wdm, rb.wdm) && same(rs.f, rb.f) && same(rs.dphase.hi, rb.dphase.hi) &&
                  same(rs.dphase.lo, rb.dphase.lo)))
                ++counts[2];
        }
    }
    return 0;
}
