// vela_structured.cu - host-side analysis of the noise basis for the structured Sigma path.
//
// For bases whose columns are harmonics of one phase per row (every power-law Fourier GP), every product
// M[j,p] M[j,q] is half the sum / difference of two entries of a small trigonometric table, so Sigma^-1's p(p+1)/2
// contractions over the TOAs collapse to 2 (K_a + K_c) + 1 column sums per pair of column groups.
// Nothing here runs on the device; the result (table T, pair recipe) is uploaded by vela_api.cu.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "vela_structured.h"

namespace vela {

namespace {
inline long round_up(long x, long m) { return (x + m - 1) / m * m; }
}  // namespace

// Looks for column groups of the form  M[j, sin_k] = s_j sin(k x_j),  M[j, cos_k] = s_j cos(k x_j)  (the power-law
// Fourier GPs; s_j is the per-row chromatic / DM-row scale, x_j one phase per row shared by all groups) and VERIFIES
// the form on the numbers of M itself (relative 2e-13), so an arbitrary basis simply stays on the dense path.
// Harmonic x harmonic products are then served by 2 (K_a + K_c) + 1 trigonometric columns per group pair, any
// product with another column by an explicit product column.  Returns false when that is not at least ~2x fewer
// rows than the packed triangle.
bool analyse_basis(const VelaModelDesc* desc, long n_rows, long n_rows_pad, int p, StructPlan& sp) {
    const char* env = getenv("VELA_B200_STRUCTURED");
    if (env && atoi(env) == 0) return false;
    const double* M = desc->noise_basis;
    const long ld = desc->basis_ld;
    auto Mc = [&](int c) { return M + (size_t)c * ld; };
    std::vector<int> group(p, -1), kh(p, 0), is_cos(p, 0);
    struct Grp { int sin1 = -1, cos1 = -1, kmax = 0; std::vector<int> cols; std::vector<double> s; bool ok = false; };
    std::vector<Grp> groups;
    int col = 0;
    for (int g = 0; g < desc->n_gp; ++g) {
        const VelaGPDesc& gp = desc->gp[g];
        if (gp.kind == VELA_GP_MARGINALIZED_TM) { col += gp.n; continue; }
        Grp G;
        const int gi = (int)groups.size();
        for (int i = 0; i < gp.n; ++i) {
            const double j = std::exp(gp.ln_js[i]);
            const long kk = std::lround(j);
            if (kk >= 1 && std::fabs(j - (double)kk) <= 1e-6 * (double)kk) {
                for (int t = 0; t < 2; ++t) {
                    const int c = col + t * gp.n + i;
                    group[c] = gi; kh[c] = (int)kk; is_cos[c] = t;
                    G.cols.push_back(c);
                }
                if (kk == 1) { G.sin1 = col + i; G.cos1 = col + gp.n + i; }
                G.kmax = std::max(G.kmax, (int)kk);
            }
        }
        col += 2 * gp.n;
        groups.push_back(std::move(G));
    }
    // one phase per row, from the first group whose fundamental is non-zero there
    std::vector<double> x(n_rows, 0.0);
    std::vector<char> have(n_rows, 0);
    for (auto& G : groups) {
        if (G.sin1 < 0) continue;
        const double *sn = Mc(G.sin1), *cs = Mc(G.cos1);
        for (long j = 0; j < n_rows; ++j)
            if (!have[j] && (sn[j] != 0.0 || cs[j] != 0.0)) { x[j] = std::atan2(sn[j], cs[j]); have[j] = 1; }
    }
    int n_ok = 0;
    for (size_t gi = 0; gi < groups.size(); ++gi) {
        Grp& G = groups[gi];
        if (G.sin1 < 0) continue;
        const double *sn = Mc(G.sin1), *cs = Mc(G.cos1);
        G.s.resize(n_rows);
        for (long j = 0; j < n_rows; ++j) G.s[j] = sn[j] * std::sin(x[j]) + cs[j] * std::cos(x[j]);   // signed scale
        bool ok = true;
        for (int c : G.cols) {
            const double* m = Mc(c);
            for (long j = 0; j < n_rows && ok; ++j) {
                const double arg = (double)kh[c] * x[j];
                const double want = G.s[j] * (is_cos[c] ? std::cos(arg) : std::sin(arg));
                if (!(std::fabs(m[j] - want) <= 2e-13 * std::fabs(G.s[j]) + 1e-300)) ok = false;
            }
            if (!ok) break;
        }
        G.ok = ok;
        n_ok += ok;
    }
    for (int c = 0; c < p; ++c)
        if (group[c] >= 0 && !groups[group[c]].ok) group[c] = -1;
    if (n_ok == 0) return false;

    // column layout of T: per unordered pair of good groups (a <= c): C(0..K), S(1..K), K = kmax_a + kmax_c
    std::vector<int> good;
    for (size_t gi = 0; gi < groups.size(); ++gi) if (groups[gi].ok) good.push_back((int)gi);
    const int ng = (int)good.size();
    std::vector<int> gslot(groups.size(), -1);
    for (int i = 0; i < ng; ++i) gslot[good[i]] = i;
    std::vector<int> baseC((size_t)ng * ng, -1), baseS((size_t)ng * ng, -1);
    long n_r = 0;
    // group pairs whose scale products coincide row by row share their columns (red x chromatic and DM x DM both
    // scale as (1400 MHz / nu)^4, for instance); "owner" marks the pair that fills them
    std::vector<char> owner((size_t)ng * ng, 0);
    for (int a = 0; a < ng; ++a)
        for (int c = a; c < ng; ++c) {
            const Grp &Ga = groups[good[a]], &Gc = groups[good[c]];
            const int K = Ga.kmax + Gc.kmax;
            bool shared = false;
            for (int a2 = 0; a2 <= a && !shared; ++a2)
                for (int c2 = a2; c2 < ng && !shared; ++c2) {
                    if (!owner[(size_t)a2 * ng + c2] || groups[good[a2]].kmax + groups[good[c2]].kmax < K) continue;
                    const Grp &Ha = groups[good[a2]], &Hc = groups[good[c2]];
                    bool same = true;
                    for (long j = 0; j < n_rows && same; ++j) {
                        const double u = Ga.s[j] * Gc.s[j], v = Ha.s[j] * Hc.s[j];
                        same = std::fabs(u - v) <= 1e-14 * std::max(std::fabs(u), std::fabs(v));
                    }
                    if (same) {
                        baseC[(size_t)a * ng + c] = baseC[(size_t)c * ng + a] = baseC[(size_t)a2 * ng + c2];
                        baseS[(size_t)a * ng + c] = baseS[(size_t)c * ng + a] = baseS[(size_t)a2 * ng + c2];
                        shared = true;
                    }
                }
            if (shared) continue;
            owner[(size_t)a * ng + c] = 1;
            baseC[(size_t)a * ng + c] = baseC[(size_t)c * ng + a] = (int)n_r; n_r += K + 1;
            baseS[(size_t)a * ng + c] = baseS[(size_t)c * ng + a] = (int)n_r - 1; n_r += K;   // S(m) at baseS + m, m >= 1
        }
    const long n_trig = n_r;
    const long n_pairs = (long)p * (p + 1) / 2;
    std::vector<std::pair<int, int>> products;
    sp.table.assign(n_pairs, PairTerm{0, 0, 0.0, 0.0});
    for (int r = 0; r < p; ++r)
        for (int c = 0; c <= r; ++c) {
            PairTerm& pt = sp.table[(size_t)r * (r + 1) / 2 + c];
            if (group[r] >= 0 && group[c] >= 0) {
                const int a = gslot[group[r]], b = gslot[group[c]];
                const int bc = baseC[(size_t)a * ng + b], bs = baseS[(size_t)a * ng + b];
                const int kr = kh[r], kc = kh[c], sum = kr + kc, dif = std::abs(kr - kc);
                if (is_cos[r] == is_cos[c]) {        // sin sin = (C(|d|) - C(sum)) / 2,  cos cos = (C(|d|) + C(sum)) / 2
                    pt = PairTerm{bc + dif, bc + sum, 0.5, is_cos[r] ? 0.5 : -0.5};
                } else {                              // sin_k cos_l = (S(k + l) + S(k - l)) / 2, S(-m) = -S(m), S(0) = 0
                    const int d = is_cos[r] ? (kc - kr) : (kr - kc);   // (harmonic of the sine) - (harmonic of the cosine)
                    pt = PairTerm{bs + sum, d == 0 ? bs + sum : bs + std::abs(d), 0.5, d == 0 ? 0.0 : (d > 0 ? 0.5 : -0.5)};
                }
            } else {
                pt = PairTerm{(int)(n_trig + (long)products.size()), 0, 1.0, 0.0};
                products.emplace_back(r, c);
            }
        }
    n_r = n_trig + (long)products.size();
    if ((double)n_r > 0.5 * (double)n_pairs) { sp.table.clear(); return false; }
    sp.n_r = (int)round_up(n_r, 32);
    sp.n_cols = sp.n_r + (int)round_up(p, 32);
    sp.n_groups = ng;
    sp.T.assign((size_t)n_rows_pad * sp.n_cols, 0.0);
    auto Tc = [&](long c) { return sp.T.data() + (size_t)c * n_rows_pad; };
    {   // trigonometric columns: cos(m x_j), sin(m x_j) once per row block, reused by every group pair
        int mmax = 0;
        for (int a = 0; a < ng; ++a) mmax = std::max(mmax, 2 * groups[good[a]].kmax);
        constexpr long JB = 128;
        std::vector<double> cm((size_t)JB * (mmax + 1)), sm((size_t)JB * (mmax + 1));
        for (long j0 = 0; j0 < n_rows; j0 += JB) {
            const long nj = std::min<long>(JB, n_rows - j0);
            for (long jj = 0; jj < nj; ++jj)
                for (int m = 0; m <= mmax; ++m) {
                    const double arg = (double)m * x[j0 + jj];
                    cm[(size_t)m * JB + jj] = std::cos(arg);
                    sm[(size_t)m * JB + jj] = std::sin(arg);
                }
            for (int a = 0; a < ng; ++a)
                for (int c = a; c < ng; ++c) {
                    if (!owner[(size_t)a * ng + c]) continue;
                    const Grp &Ga = groups[good[a]], &Gc = groups[good[c]];
                    const int K = Ga.kmax + Gc.kmax, bc = baseC[(size_t)a * ng + c], bs = baseS[(size_t)a * ng + c];
                    for (int m = 0; m <= K; ++m) {
                        double* tc = Tc(bc + m) + j0;
                        double* ts = m >= 1 ? Tc(bs + m) + j0 : nullptr;
                        for (long jj = 0; jj < nj; ++jj) {
                            const double ss = Ga.s[j0 + jj] * Gc.s[j0 + jj];
                            tc[jj] = ss * cm[(size_t)m * JB + jj];
                            if (ts) ts[jj] = ss * sm[(size_t)m * JB + jj];
                        }
                    }
                }
        }
    }
    for (size_t i = 0; i < products.size(); ++i) {
        const double *a = Mc(products[i].first), *b = Mc(products[i].second);
        double* t = Tc(n_trig + (long)i);
        for (long j = 0; j < n_rows; ++j) t[j] = a[j] * b[j];
    }
    for (int c = 0; c < p; ++c) memcpy(Tc(sp.n_r + c), Mc(c), sizeof(double) * n_rows);
    sp.active = true;
    return true;
}


}  // namespace vela
