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


namespace {

// J_q(z), q = 0 .. 31, for 0 <= z <= ~8: ascending series in long double (terms decay like (z/2)^(2k+q) / (k! (k+q)!))
void bessel_j(long double z, long double* J, int nq) {
    const long double h = z / 2;
    for (int q = 0; q < nq; ++q) {
        long double term = 1.0L;
        for (int i = 1; i <= q; ++i) term *= h / i;          // (z/2)^q / q!
        long double sum = term;
        for (int k = 1; k < 60; ++k) {
            term *= -(h * h) / ((long double)k * (k + q));
            sum += term;
            if (fabsl(term) < 1e-22L * fabsl(sum)) break;
        }
        J[q] = sum;
    }
}

constexpr int kCheb = 32;            // Chebyshev terms per block = columns of one colsum unit
constexpr long double kZMax = 7.0L;   // largest m h_c: 2 J_32(7) < 2e-18

struct Block {                        // one 32-column unit of the stage-1 matrix
    bool with_y = false;              // contracts with w.y (u columns) instead of w
    bool plain = false;               // explicit columns (no expansion): cols[i] is the per-row vector of column i
    std::vector<double> f;            // harmonic block: per-row scale function
    std::vector<std::vector<double>> cols;
};

struct OutCol {                       // where one column of R (or u) comes from
    int block = -1;
    int m = 0;                        // harmonic number (harmonic blocks) or column slot (plain blocks)
    bool is_sin = false;
};

bool same_vec(const std::vector<double>& a, const std::vector<double>& b) {
    for (size_t j = 0; j < a.size(); ++j)
        if (std::fabs(a[j] - b[j]) > 1e-14 * std::max(std::fabs(a[j]), std::fabs(b[j]))) return false;
    return true;
}

}  // namespace

// Two-stage plan for the column sums described by `out` (one entry per column of R followed by the p columns of u):
// cells of `cp` consecutive rows, stage-1 matrix A1 and the stage-2 jobs.  Returns false (plan left inactive) when the
// rows are not ordered well enough for small cells (some cell's phases span more than 2 kZMax / K), or nothing is gained.
static bool build_two_stage(long n_rows, long& n_rows_pad, const std::vector<double>& x, int kmax_all,
                            const std::vector<Block>& blocks, const std::vector<OutCol>& out, int n_r, StructPlan& sp) {
    if (const char* e = getenv("VELA_B200_TWO_STAGE")) { if (atoi(e) == 0) return false; }
    if (blocks.empty() || kmax_all <= 0) return false;
    const long double two_pi = 6.283185307179586476925286766559L;
    // rows per cell: the largest multiple of 32 (<= 4096) whose every cell keeps K h_c <= kZMax
    auto cell_geometry = [&](int cp, std::vector<long double>* g, std::vector<long double>* h) -> bool {
        const long nc = (n_rows + cp - 1) / cp;
        if (g) { g->assign(nc, 0.0L); h->assign(nc, 1.0L); }
        for (long c = 0; c < nc; ++c) {
            const long j0 = c * cp, j1 = std::min<long>(n_rows, j0 + cp);
            const long double ref = x[j0];
            long double lo = 0.0L, hi = 0.0L;                 // offsets from the first row's phase, unwrapped onto (-pi, pi]
            for (long j = j0; j < j1; ++j) {
                long double d = (long double)x[j] - ref;
                d -= two_pi * roundl(d / two_pi);
                lo = std::min(lo, d);
                hi = std::max(hi, d);
            }
            const long double half = (hi - lo) / 2;
            if (half * kmax_all > kZMax) return false;
            if (g) { (*g)[c] = ref + (lo + hi) / 2; (*h)[c] = std::max(half, 1e-12L); }
        }
        return true;
    };
    int cp = 0;
    for (int trial = 4096; trial >= 32; trial -= 32)
        if (trial <= std::max<long>(32, n_rows) && cell_geometry(trial, nullptr, nullptr)) { cp = trial; break; }
    if (cp == 0) return false;
    // keep at least ~24 cells when the data allow it: the cells are the stage-1 grid's third dimension
    while (cp > 32 && (n_rows + cp - 1) / cp < 24 && cell_geometry(cp - 32, nullptr, nullptr)) cp -= 32;
    const long lcm = (long)cp * 64 / std::__gcd((long)cp, 64L);
    const long rows_pad = round_up(n_rows, lcm);
    const int n_cells = (int)(rows_pad / cp);
    std::vector<long double> g, h;
    cell_geometry(cp, &g, &h);
    g.resize(n_cells, 0.0L);
    h.resize(n_cells, 1.0L);

    // work estimate: stage 1 + stage 2 against the direct contraction; demand a clear gain
    const int n_blocks = (int)blocks.size();
    const double direct = (double)rows_pad * (double)sp.n_cols;
    double two = (double)rows_pad * 32.0 * n_blocks;
    for (const OutCol& o : out) if (o.block >= 0) two += (double)n_cells * 32.0;
    if (two > 0.75 * direct) return false;

    sp.cp = cp; sp.n_cells = n_cells; sp.n_blocks = n_blocks;
    sp.k2_pad = round_up((long)n_cells * kCheb, 64);
    n_rows_pad = rows_pad;
    // ---- stage-1 matrix: W blocks first
    std::vector<int> order;
    for (int pass = 0; pass < 2; ++pass)
        for (int b = 0; b < n_blocks; ++b) if ((int)blocks[b].with_y == pass) order.push_back(b);
    std::vector<int> slot(n_blocks, 0);
    sp.n_w_blocks = 0;
    for (int i = 0; i < n_blocks; ++i) { slot[order[i]] = i; if (!blocks[order[i]].with_y) ++sp.n_w_blocks; }
    sp.A1.assign((size_t)rows_pad * kCheb * n_blocks, 0.0);
    std::vector<long double> tq((size_t)n_rows * kCheb);
    for (long j = 0; j < n_rows; ++j) {
        const long c = j / cp;
        long double d = (long double)x[j] - g[c];
        d -= two_pi * roundl(d / two_pi);
        long double tau = d / h[c];
        tau = std::max(-1.0L, std::min(1.0L, tau));
        const long double th = acosl(tau);
        for (int q = 0; q < kCheb; ++q) tq[(size_t)j * kCheb + q] = cosl(q * th);
    }
    for (int b = 0; b < n_blocks; ++b) {
        const Block& B = blocks[b];
        for (int q = 0; q < kCheb; ++q) {
            double* col = sp.A1.data() + ((size_t)slot[b] * kCheb + q) * rows_pad;
            if (B.plain) {
                if (q < (int)B.cols.size()) for (long j = 0; j < n_rows; ++j) col[j] = B.cols[q][j];
            } else {
                for (long j = 0; j < n_rows; ++j) col[j] = (double)((long double)B.f[j] * tq[(size_t)j * kCheb + q]);
            }
        }
    }
    // ---- stage-2 jobs: maximal runs of consecutive output columns fed by one block
    const int n_out_all = (int)out.size();
    sp.jobs.clear();
    sp.T2.clear();
    std::vector<long double> J(kCheb);
    for (int c0 = 0; c0 < n_out_all;) {
        if (out[c0].block < 0) { ++c0; continue; }             // padding columns of R
        int c1 = c0 + 1;
        while (c1 < n_out_all && out[c1].block == out[c0].block && c1 - c0 < 4096) ++c1;
        StructPlan::Stage2Job job;
        job.block = slot[out[c0].block];
        job.r_off = c0;                                      // `out` is indexed like R: padded column sums, then u
        job.n_out = c1 - c0;
        job.t2_off = (long)sp.T2.size();
        const long ncol_pad = round_up(job.n_out, 32);
        sp.T2.resize(sp.T2.size() + (size_t)sp.k2_pad * ncol_pad, 0.0);
        double* t2 = sp.T2.data() + job.t2_off;
        const Block& B = blocks[out[c0].block];
        for (int oc = 0; oc < job.n_out; ++oc) {
            const OutCol& o = out[c0 + oc];
            double* col = t2 + (size_t)oc * sp.k2_pad;
            for (int c = 0; c < n_cells; ++c) {
                if (B.plain) { col[(size_t)c * kCheb + o.m] = 1.0; continue; }
                bessel_j((long double)o.m * h[c], J.data(), kCheb);
                const long double ang = (long double)o.m * g[c];
                const long double cg = cosl(ang), sg = sinl(ang);
                for (int q = 0; q < kCheb; ++q) {
                    // a_q = (q ? 2 : 1) i^q J_q;  exp(i m g) a_q = (cg + i sg) a_q
                    const long double amp = (q ? 2.0L : 1.0L) * J[q];
                    long double re, im;
                    switch (q & 3) {
                        case 0: re = amp * cg; im = amp * sg; break;
                        case 1: re = -amp * sg; im = amp * cg; break;
                        case 2: re = -amp * cg; im = -amp * sg; break;
                        default: re = amp * sg; im = -amp * cg; break;
                    }
                    col[(size_t)c * kCheb + q] = (double)(o.is_sin ? im : re);
                }
            }
        }
        sp.jobs.push_back(job);
        c0 = c1;
    }
    sp.two_stage = true;
    return true;
}

// Looks for column groups of the form  M[j, sin_k] = s_j sin(k x_j),  M[j, cos_k] = s_j cos(k x_j)  (the power-law
// Fourier GPs; s_j is the per-row chromatic / DM-row scale, x_j one phase per row shared by all groups) and VERIFIES
// the form on the numbers of M itself (relative 2e-13), so an arbitrary basis simply stays on the dense path.
// Harmonic x harmonic products are then served by 2 (K_a + K_c) + 1 trigonometric columns per group pair, any
// product with another column by an explicit product column.  Returns false when that is not at least ~2x fewer
// rows than the packed triangle.
bool analyse_basis(const VelaModelDesc* desc, long n_rows, long& n_rows_pad, int p, StructPlan& sp) {
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
    {   // ---- two-stage plan: which scale function ("block") and harmonic every column of R and u comes from
        std::vector<Block> blocks;
        std::vector<OutCol> out((size_t)sp.n_r + p);
        int kmax_all = 0;
        auto harmonic_block = [&](bool with_y, const std::vector<double>& f) -> int {
            for (size_t b = 0; b < blocks.size(); ++b)
                if (!blocks[b].plain && blocks[b].with_y == with_y && same_vec(blocks[b].f, f)) return (int)b;
            Block B;
            B.with_y = with_y; B.f = f;
            blocks.push_back(std::move(B));
            return (int)blocks.size() - 1;
        };
        auto plain_slot = [&](bool with_y, std::vector<double> col, int* slot) -> int {
            int b = -1;
            for (size_t i = 0; i < blocks.size(); ++i)
                if (blocks[i].plain && blocks[i].with_y == with_y && (int)blocks[i].cols.size() < kCheb) b = (int)i;
            if (b < 0) {
                Block B;
                B.with_y = with_y; B.plain = true;
                blocks.push_back(std::move(B));
                b = (int)blocks.size() - 1;
            }
            *slot = (int)blocks[b].cols.size();
            blocks[b].cols.push_back(std::move(col));
            return b;
        };
        std::vector<double> f(n_rows);
        for (int a = 0; a < ng; ++a)
            for (int c = a; c < ng; ++c) {
                if (!owner[(size_t)a * ng + c]) continue;
                const Grp &Ga = groups[good[a]], &Gc = groups[good[c]];
                const int K = Ga.kmax + Gc.kmax, bc = baseC[(size_t)a * ng + c], bs = baseS[(size_t)a * ng + c];
                for (long j = 0; j < n_rows; ++j) f[j] = Ga.s[j] * Gc.s[j];
                const int b = harmonic_block(false, f);
                for (int m = 0; m <= K; ++m) out[bc + m] = OutCol{b, m, false};
                for (int m = 1; m <= K; ++m) out[bs + m] = OutCol{b, m, true};
                kmax_all = std::max(kmax_all, K);
            }
        for (size_t i = 0; i < products.size(); ++i) {
            const int r = products[i].first, c = products[i].second;
            const int hc = group[r] >= 0 ? r : (group[c] >= 0 ? c : -1), v = hc == r ? c : r;
            const double *mr = Mc(r), *mc = Mc(c);
            if (hc >= 0) {      // (harmonic column) x (other column): harmonic in x with the scale  s_g M[:, v]
                const Grp& G = groups[group[hc]];
                const double* mv = Mc(v);
                for (long j = 0; j < n_rows; ++j) f[j] = G.s[j] * mv[j];
                out[n_trig + i] = OutCol{harmonic_block(false, f), kh[hc], !is_cos[hc]};
                kmax_all = std::max(kmax_all, kh[hc]);
            } else {
                std::vector<double> col(n_rows);
                for (long j = 0; j < n_rows; ++j) col[j] = mr[j] * mc[j];
                int slot = 0;
                const int b = plain_slot(false, std::move(col), &slot);
                out[n_trig + i] = OutCol{b, slot, false};
            }
        }
        for (int c = 0; c < p; ++c) {
            if (group[c] >= 0) {
                const int b = harmonic_block(true, groups[group[c]].s);
                out[sp.n_r + c] = OutCol{b, kh[c], !is_cos[c]};
                kmax_all = std::max(kmax_all, kh[c]);
            } else {
                int slot = 0;
                const int b = plain_slot(true, std::vector<double>(Mc(c), Mc(c) + n_rows), &slot);
                out[sp.n_r + c] = OutCol{b, slot, false};
            }
        }
        // rows no harmonic group touches carry no phase: give them their neighbour's, they weigh nothing
        std::vector<double> xc(x);
        for (long j = 1; j < n_rows; ++j) if (!have[j]) xc[j] = xc[j - 1];
        for (long j = n_rows - 2; j >= 0; --j) if (!have[j] && !have[j + 1] && j + 1 < n_rows) xc[j] = xc[j + 1];
        build_two_stage(n_rows, n_rows_pad, xc, kmax_all, blocks, out, sp.n_r, sp);
    }
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
