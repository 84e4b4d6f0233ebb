// vela_kernels.cuh — launch-argument structs and tile constants shared by vela_kernels.cu and vela_api.cu
#pragma once
#include "vela_chain.cuh"

namespace vela {

// K3 tiling
constexpr int kSigMaxSamples = 64;  // walkers per CTA: 8 x (n-tiles per warp)
constexpr int kSigMaxRank = 1024;

constexpr int kEcorrChunk = 64;         // TOAs staged per pass of the ECORR basis kernel
constexpr int kCholThreads = 256;        // sizing of the per-warp scratch
constexpr int kCholLaunchThreads = 128;  // threads per walker CTA (measured best across ranks 60..300)

struct PriorDev {
    int kind;
    double a, b, lo, hi;
    double norm;               // kind-dependent precomputed constant (log normalisation)
};

struct SigmaArgs {
    const double* M;           // column-major, leading dimension ldM (rows zero-padded to n_rows_pad)
    long ldM;
    int rank;                  // p
    long n_pairs;              // p (p + 1) / 2 packed lower-triangle rows
    int n_pair_units;          // units of MT m-tiles covering the pairs
    int n_units;               // pair units + u units
    long n_rows_pad;
    const double* W;
    const double* Y;
    long ld_yw;
    long B;
    double* Sig;               // [B][ld_sig], packed lower triangle t = p(p+1)/2 + q
    long ld_sig;
    double* U;                 // [B][ld_u]
    long ld_u;
    int kt_per_split;          // k-tiles per split-K slice (gridDim.z slices)
    long split_stride_sig;     // doubles between the Sigma slabs of consecutive slices
    long split_stride_u;
};

// One CTA column of a stage-2 launch of the two-stage column sums (ColsumArgs::units): up to 32 UC output columns fed by
// one block's cell moments.
struct ColsumUnit {
    long t_off;                // first of its columns inside ColsumArgs::T (doubles)
    long w_off;                // its stage-1 block: the moments start at ColsumArgs::W + w_off * ColsumArgs::w_plane
    int r_col;                 // first output column in R
    int n_valid;               // output columns of this CTA column (<= 32 UC); the rest of the tile is padding
};

// K3s: column sums  R[b, c] = sum_j T[j, c] * (c < n_w_cols ? w[b, j] : w[b, j] y[b, j])  (structured Sigma path)
struct ColsumArgs {
    const double* T;           // column-major [n_rows_pad x n_cols], leading dimension ldT; n_cols multiple of 32
    long ldT;
    int n_cols;
    int n_w_units;             // 32-column units that use w; the remaining units use w.y
    long n_rows_pad;
    const double* W;
    const double* Y;
    long ld_yw;
    long B;
    double* R;                 // [B][ld_r]
    long ld_r;
    int kt_per_split;
    long split_stride_r;       // doubles between the slabs of consecutive split-K slices
    // two-stage evaluation (vela_structured.h): stage 1 stores the cell moments walker-major per 32-column block,
    //   R[(col / 32) * plane_stride + b * ld_r + slice * 32 + col % 32]      (plane_stride > 0; slice = cell)
    // and stage 2 stores only its first n_valid columns (0 = all): its output runs are not multiples of 32
    long plane_stride = 0;
    int n_valid = 0;
    // stage 2 as ONE launch: blockIdx.x indexes this table instead of the 32 UC-column units of T (nullptr otherwise)
    const ColsumUnit* units = nullptr;
    long w_plane = 0;          // doubles between the moment planes of consecutive blocks
};

// Sigma[p][q] = c1 R[i1] + c2 R[i2]  (c2 == 0: single term) for the packed pair t = p(p+1)/2 + q
struct PairTerm {
    int i1, i2;
    double c1, c2;
};

constexpr int kEcorrChunksPerRow = 8;   // chunks one warp of ecorr_group_kernel walks = chunks per `partial` row
struct EcorrArgs {
    const double* Pext;
    int row;
    const double* Y;
    const double* W;
    long ld_yw;
    long B;
    const int3* groups;        // [n_groups] (first TOA 0-based, one past last, ECORR index)
    const int* chunk_start;    // [n_chunks + 1] group ranges (at most 32 groups each)
    int n_chunks;
    int par_ecorr;
    int first_row;             // first `partial` row owned by the ECORR corrections (= number of TOA tiles)
    double* partial;
    int y_is_wy;               // the Y buffer holds w.y (ChainArgs::emit == 3)
};

struct EcorrBasisArgs {
    const double* M;           // column-major noise basis
    long ldM;
    int rank;
    const double* Pext;
    int row;
    int par_ecorr;
    const double* Y;
    const double* W;
    long ld_yw;
    long B;
    const int3* groups;        // ECORR-active groups only (index >= 1)
    int n_groups;
    int groups_per_block;
    int chunk;                 // TOAs staged per pass (kEcorrChunk unless the rank is too large for shared memory)
    int y_is_wy;               // the Y buffer holds w.y (ChainArgs::emit == 3)
    double* partial;           // when set: the scalar ECORR corrections (chi^2, log det) of this block's groups go to
    int first_row;             //   partial[first_row + blockIdx.x][b][2] and ecorr_group_kernel is not needed
    double* V;                 // [B][g_pad][ldv]: sqrt(alpha_g) [P_g | r_g | 0...]
    long g_pad;
    long ldv;
};

struct EcorrSyrkArgs {
    const double* V;
    long g_pad;
    long ldv;
    int rank;
    int n_rect;                // 32 x 32 rectangles of the lower block triangle of (rank + 1)
    long B;
    double* Sig;
    long ld_sig;
    double* U;
    long ld_u;
};

struct GPDev {
    int kind, n, par_log10_amp, par_gamma, par_f1;
    int col0;                  // first basis column of this GP component
    int data0;                 // offset into gp_data (ln_js or prior_weights_inv)
};

struct CholArgs {
    double* Sig;               // packed lower triangle per walker
    long ld_sig;
    long ld_u;
    double* Sq;                // [B][p + 1][pp + 1] scratch, only when the matrix does not fit shared memory
    int ksplit;                // number of split-K slabs to sum
    long split_stride_sig;
    long split_stride_u;
    const double* U;
    const double* Pext;
    int row;
    int p, pp;
    int n_gp;
    const GPDev* gp;
    const double* gp_data;
    const double* partial;
    int n_tiles;
    long B;
    const double* lnprior;
    int with_prior;
    double* out;
    int use_smem;
    // structured path: Sigma is assembled from the column sums R through the pair table (Sig then only carries the
    // ECORR correction, or is unused)
    const PairTerm* table;     // null = dense path (Sig holds the packed triangle)
    const double* R;           // [B][ld_r]: n_r column sums, then u at column u_off
    long ld_r;
    long split_stride_r;
    int n_r, u_off;
    int add_sig;               // add Sig[t] (ECORR correction) to the assembled entry
    const unsigned* ctab;      // compact pair table (chol_panel_kernel): per packed pair i1 | i2 << 16 into S = {+R/2 | -R/2 | 0}, null = use `table`
    const unsigned* ltab;      // chol_warp_kernel: layout table (one code per shared-memory slot), null = not available
    const unsigned* ftab;      // chol_dmma_kernel: fragment table (one code per block register and lane), null = not available
    double* ahat;              // optional [B][p]: conditional mean of the marginalised amplitudes, Sigma u
    double* cf;                // optional [B][p][p] row-major upper Cholesky factor of Sigma^-1 (L^T)
};

__global__ void sample_prologue_kernel(const __grid_constant__ ProgramDev prog, PrologueArgs a);
__global__ void toa_chain_kernel(const __grid_constant__ ProgramDev prog, ChainArgs a);
// create(): C_DELTA columns of the TOA table and the TZR spin phase at the default parameters (kFlagDelta)
__global__ void default_phase_kernel(const __grid_constant__ ProgramDev prog, const double* P, double* toa, long stride, long n_toas,
                                     const double* tzr_block, const uint32_t* index_masks, const uint8_t* bit_masks, double* tzr0);
__global__ void prior_kernel(const PriorDev* priors, int nd, const double* params, long B, int mode, double* out);
__global__ void ecorr_group_kernel(EcorrArgs a);
template <int NACC>
__global__ void ecorr_basis_kernel(EcorrBasisArgs a);
__global__ void ecorr_syrk_kernel(EcorrSyrkArgs a);
__global__ void ecorr_syrk64_kernel(EcorrSyrkArgs a);   // ranks <= 63 (ldv == 64)
__global__ void finish_white_kernel(const double* partial, int n_tiles, long B, int mode, const double* lnprior,
                                    int with_prior, double* out);
template <int BK, int STAGES, int MT, int NT, int WARPS>
__global__ void sigma_contract_kernel(SigmaArgs a);

template <int NB, bool PACKED>
__global__ void chol_finish_kernel(CholArgs a);
constexpr int kCholSmallRank = 96;   // block width 8 up to this rank, 16 above
inline size_t chol_smem_doubles(int p, int pp, int nb, bool matrix_in_smem, int n_r) {
    return (size_t)2 * pp + (size_t)nb * (nb + 1) + nb + (size_t)(p + 1) * (nb + 1) +
           (matrix_in_smem ? (size_t)(p + 1) * (p + 2) / 2 : 0) + (size_t)n_r;   // n_r: column sums of the structured path
}
// K4p: large ranks, 32-column panels in shared memory + DMMA trailing updates on the global scratch (vela_kernels.cu)
__global__ void chol_panel_kernel(CholArgs a);
constexpr int kPanelThreadsHost = 256;   // = kPanelThreads in vela_kernels.cu
inline size_t chol_panel_smem_bytes(int p, int pp, int n_ru) {
    // phi[pp] | column buffers[96] | [8] | sW[4][64] | Tb[32][36] | P[prow][36]: prow = rows below the first panel's top block,
    // rounded up to the 32-row tiles of the trailing update
    const int n1 = p + 1, prow = (n1 - (p < 32 ? p : 32) + 31) & ~31;
    const size_t panel = (size_t)prow * 36;                // kPanelPitch
    const size_t sums = 2 * (size_t)n_ru + 2;              // S = {+R/2 | -R/2 | 0} aliases the panel during assembly
    return ((size_t)pp + 96 + 8 + 256 + 32 * 36 + (panel > sums ? panel : sums)) * sizeof(double);
}
inline size_t chol_panel_scratch_doubles(int p, int pp) { return (size_t)(p + 1) * (pp + 2); }   // per walker

// K4w: one warp per walker, ranks up to 63 (vela_kernels.cu).  Shared-memory layout of the walker's matrix (rows 0..p,
// row p = u^T): columns in blocks of four, column k holding rows (k & ~3) .. NR-1, NR = p + 1 rounded up to even.
constexpr int kCholWarpMaxRows = 64;
constexpr int kCholWarpSlack = 64;       // zeroed doubles behind the matrix: lanes past the last row read them
__host__ __device__ inline int cw_col_base(int m, int NR) { return 4 * m * NR - 8 * m * (m - 1); }   // first entry of column block m
inline int cw_slot(int row, int col, int NR) { return cw_col_base(col >> 2, NR) + (col & 3) * (NR - (col & ~3)) + row - (col & ~3); }
inline size_t chol_warp_smem_bytes(int p, int n_ru) {
    const int n1 = p + 1, NR = (n1 + 1) & ~1, nblk = (n1 + 3) / 4;
    return ((size_t)cw_col_base(nblk, NR) + kCholWarpSlack + (size_t)((n_ru + 1) & ~1)) * sizeof(double);
}
// layout-table codes (one per shared-memory slot): [12:0] i1, [25:13] i2, [26] c1 is 1.0 (else 0.5), [28:27] c2
// (0 none, 1 +0.5, 2 -0.5), [31:29] kind (0 zero, 1 c1 R[i1] + c2 R[i2], 2 copy R[i1])
constexpr int kCholWarpMaxCols = 8191;
__global__ void chol_warp_kernel(CholArgs a);
// K4d (vela_kernels.cu): tensor-core Cholesky finish, one warp per walker, ranks up to kCholDmmaMaxRank
constexpr int kCholDmmaMaxRank = 64;
__host__ __device__ constexpr int cd_pi(int n) { return 4 * (n & 1) + (n >> 1); }   // row / column permutation of its block layout
template <int NBK>
__global__ void chol_dmma_kernel(CholArgs a);
inline int chol_dmma_nbk(int p) { const int n = ((p + 7) / 8 + 1) & ~1; return n < 2 ? 2 : n; }   // block rows: 2, 4, 6 or 8
inline size_t chol_dmma_smem_bytes(int nbk, int n_ru) { return (size_t)(64 + 72 + 8 * nbk + 2 * n_ru + 2) * sizeof(double); }
__global__ void slab_reduce_kernel(double* R, long stride, int n_slabs, long n);
__global__ void phiinv_kernel(CholArgs a, double* out);
__global__ void div3_check_kernel(unsigned long long seed, long n, unsigned long long* mismatches);
__global__ void sincos_cr_kernel(const double* x, int n, double* s, double* c);
__global__ void dfma_peak_kernel(double* out, int iters);
__global__ void dmma_peak_kernel(double* out, int iters);
__global__ void dmma_mix_kernel(double* out, int iters, int nmul, int nlds);
__global__ void fp64_coissue_kernel(double* out, int iters, int mode);

template <int BK, int STAGES, int UC, int WC, bool SPLIT>
__global__ void colsum_kernel(ColsumArgs a);
inline size_t colsum_smem_bytes(int bk, int stages, int uc, int wc, bool split) {
    return (size_t)stages * (uc * 32 + (split ? 1 : 2) * wc * 32) * (bk + 4) * sizeof(double);
}

inline size_t sigma_smem_bytes(int rank, int bk, int stages, int nt) {
    return (size_t)stages * (rank + 2 * 8 * nt) * (bk + 4) * sizeof(double);
}

}  // namespace vela
