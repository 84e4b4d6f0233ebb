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
};

// Fills `sp` and returns true when the basis of `desc` (n_rows x p, Woodbury kernels) has enough harmonic structure
// for the column-sum path; false leaves the handle on the dense contraction.  VELA_B200_STRUCTURED=0 forces false.
bool analyse_basis(const VelaModelDesc* desc, long n_rows, long n_rows_pad, int p, StructPlan& sp);

}  // namespace vela
