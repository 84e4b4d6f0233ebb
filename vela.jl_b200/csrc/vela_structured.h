// vela_structured.h - host-side analysis for the structured Sigma path (see colsum_kernel in vela_kernels.cu)
#pragma once
#include <vector>

#include "vela_kernels.cuh"

namespace vela {

// The table of trigonometric / product columns T, M's columns appended for u, and the per-pair recipe that rebuilds
// Sigma from the column sums.
struct StructPlan {
    bool active = false;
    int n_r = 0;               // T columns (multiple of 32)
    int n_cols = 0;            // n_r + round_up(p, 32)
    int n_groups = 0;          // harmonic column groups found
    std::vector<double> T;     // n_rows_pad x n_cols, column-major (host copy dropped after upload)
    std::vector<PairTerm> table;

    // ---- two-stage evaluation of the same column sums (large batches), see build_two_stage() in vela_structured.cu
    // Rows are cut into cells of `cp` consecutive rows.  Within a cell the phase x_j stays within +-h_c of its centre g_c,
    // so  exp(i m x_j) = exp(i m g_c) sum_q a_q(m h_c) T_q(tau_j),  tau_j = (x_j - g_c) / h_c,  with Chebyshev polynomials
    // T_q and a_0 = J_0, a_q = 2 i^q J_q (Jacobi-Anger); 32 terms carry it to 1e-18 for m h_c <= 7.  Stage 1 contracts the
    // walkers' weights with the 32 Chebyshev columns of every distinct per-row scale function ("block") cell by cell - 32
    // columns per block instead of 2 (K_a + K_c) + 1 per group pair; stage 2 maps the cell moments to the column sums
    // with small constant matrices.  Both stages are launches of colsum_kernel.
    struct Stage2Job {
        int block;             // stage-1 block whose moments feed this run of output columns
        int r_off;             // first output column (index into R, u columns included: n_r + basis column)
        int n_out;             // number of output columns
        long t2_off;           // its matrix inside T2: [k2_pad x round_up(n_out, 32)], column-major
    };
    bool two_stage = false;
    int cp = 0;                // rows per cell (multiple of 32)
    int n_cells = 0;
    int n_blocks = 0, n_w_blocks = 0;   // 32-column blocks of A1; the first n_w_blocks contract with w, the rest with w.y
    long k2_pad = 0;           // padded length of a walker's moment vector per block (n_cells * 32 rounded up to 64)
    std::vector<double> A1;    // n_rows_pad x (32 n_blocks), column-major
    std::vector<double> T2;
    std::vector<Stage2Job> jobs;
};

// Fills `sp` and returns true when the basis of `desc` (n_rows x p, Woodbury kernels) has enough harmonic structure
// for the column-sum path; false leaves the handle on the dense contraction.  VELA_B200_STRUCTURED=0 forces false.
// n_rows_pad may grow to a multiple of the two-stage cell size (the caller sizes every row-indexed buffer with it).
bool analyse_basis(const VelaModelDesc* desc, long n_rows, long& n_rows_pad, int p, StructPlan& sp);

}  // namespace vela
