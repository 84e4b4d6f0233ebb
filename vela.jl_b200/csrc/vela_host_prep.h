// vela_host_prep.h - host-side preparation of the device inputs, shared by vela_api.cu (create()) and by the host-side
// arithmetic check under tests/host_emul: the SoA TOA table with its TOA-only hoists, the component program, and the
// expansion points of the Shapiro-delay logarithms.
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "vela_device.cuh"

namespace vela {

// TOA record (Julia isbits layout) -> SoA columns with the TOA-only hoists
inline void fill_toa_columns(const double* rec, int wideband, double* cols, long stride, long j) {
    auto put = [&](int c, double v) { cols[(size_t)c * stride + j] = v; };
    put(C_T_HI, rec[0]); put(C_T_LO, rec[1]);
    put(C_ERR2, rec[2] * rec[2]);  // toa.error * toa.error, toa.jl:123
    put(C_FREQ, rec[3]); put(C_PN_HI, rec[4]); put(C_PN_LO, rec[5]);
    for (int i = 0; i < 3; ++i) { put(C_POS + i, rec[6 + i]); put(C_VEL + i, rec[9 + i]); put(C_SUN + i, rec[12 + i]); }
    const double* s = rec + 12;
    const double* R = rec + 6;
    put(C_RSUN, std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]));
    put(C_R2, R[0] * R[0] + R[1] * R[1] + R[2] * R[2]);
    put(C_DM, wideband ? rec[31] : 0.0);
    put(C_DMERR2, wideband ? rec[32] * rec[32] : 0.0);
    {   // fast build: chromatic powers and FD logarithms start from ln(frequency), kept to ~1e-19 in two words
        const long double lf = logl((long double)rec[3]);
        const double hi = (double)lf;
        put(C_LOG_FREQ, hi);
        put(C_LOG_FREQ + 1, std::isfinite(hi) ? (double)(lf - (long double)hi) : 0.0);
    }
    for (int b = 0; b < 5; ++b) {
        const double* r = rec + 15 + 3 * b;
        for (int i = 0; i < 3; ++i) put(C_PLANET + 3 * b + i, r[i]);
        put(C_RPLANET + b, std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]));
    }
}

// Component program of a validated descriptor (offsets of the hoisted per-sample values included).
// `fast`: the fast-arithmetic build of the chain (kFlagFast, vela_device.cuh); strict_arithmetic() is the switch.
inline bool strict_arithmetic() {
    const char* e = getenv("VELA_B200_STRICT");
    return e && atoi(e) != 0;
}
inline void build_program(const VelaModelDesc* desc, ProgramDev& pg, bool fast = !strict_arithmetic()) {
    memset(&pg, 0, sizeof pg);
    pg.n_comp = desc->n_components;
    pg.n_params = desc->n_params;
    pg.n_free = desc->n_free;
    pg.wideband = desc->wideband;
    int nder = 0;
    for (int ic = 0; ic < desc->n_components; ++ic) {
        const VelaComponentDesc& c = desc->components[ic];
        CompDev& cd = pg.comp[ic];
        cd.kind = c.kind; cd.flags = (c.flags & ~kFlagFast) | (fast ? kFlagFast : 0);
        memcpy(cd.par, c.par, sizeof cd.par);
        memcpy(cd.len, c.len, sizeof cd.len);
        memcpy(cd.mask, c.mask, sizeof cd.mask);
        memcpy(cd.dval, c.dval, sizeof cd.dval);
        cd.der = nder;
        nder += derived_count(c.kind, c.len);
    }
    // kFlagDelta (vela_device.cuh): spin phase as a difference from the default parameters - when Spindown is the model's
    // only spin-phase component and nothing after it touches the delay
    if (fast && !(getenv("VELA_B200_DELTA") && atoi(getenv("VELA_B200_DELTA")) == 0)) {
        int isp = -1, nsp = 0;
        for (int ic = 0; ic < pg.n_comp; ++ic) if (pg.comp[ic].kind == VELA_COMP_SPINDOWN) { isp = ic; ++nsp; }
        bool ok = nsp == 1;
        for (int ic = isp + 1; ok && ic < pg.n_comp; ++ic) {
            const int kd = pg.comp[ic].kind;
            ok = kd == VELA_COMP_PHASE_OFFSET || kd == VELA_COMP_PHASE_JUMP || kd == VELA_COMP_PHASE_JUMP_EXCLUSIVE ||
                 kd == VELA_COMP_GLITCH || kd == VELA_COMP_MEASUREMENT_NOISE || kd == VELA_COMP_DM_MEASUREMENT_NOISE ||
                 kd == VELA_COMP_DM_JUMP || kd == VELA_COMP_DM_JUMP_EXCLUSIVE;
        }
        if (ok) pg.comp[isp].flags |= kFlagDelta;
    }
    pg.n_derived = nder;
    pg.row = pg.n_params + nder;
    if ((pg.row & 1) == 0) pg.row += 1;  // odd pitch: conflict-free column reads across rows
}

// Expansion points of shapiro_log() (vela_device.cuh): for every TOA, x0 = r - L0.r for the Sun and the five planets with
// L0 the pulsar direction at the DEFAULT parameter values (evaluate_proper_motion, solarsystem.jl:45-56, plain libm: the
// point only has to be close), 1 / x0 and log(x0 / AU).  Where the geometry is degenerate (x0 not positive or not
// finite) or the component lacks parameters, 1 / x0 is NaN, which sends the device down the general branch.
inline void fill_shapiro_columns(const VelaModelDesc* desc, double* cols, long stride, long j) {
    const VelaComponentDesc* ss = nullptr;
    for (int ic = 0; ic < desc->n_components; ++ic)
        if (desc->components[ic].kind == VELA_COMP_SOLAR_SYSTEM) { ss = &desc->components[ic]; break; }
    if (!ss) return;
    // 1 / x0 = NaN marks "no expansion point": d = (x - x0) * NaN fails |d| < 1e-4 and the device takes log(x / AU)
    for (int b = 0; b < 6; ++b) cols[(size_t)(C_SHAP + 3 * b + 1) * stride + j] = std::nan("");
    cols[(size_t)(C_SWG + 1) * stride + j] = std::nan("");   // solar_wind_geometry(): no expansion point either
    const double* P = desc->default_values;
    for (int k = 0; k < 6; ++k) if (ss->par[k] < 0) return;
    const double lon = P[ss->par[0]], lat = P[ss->par[1]], pma = P[ss->par[2]], pmd = P[ss->par[3]], posepoch = P[ss->par[5]];
    const double sa = std::sin(lon), ca = std::cos(lon), sd = std::sin(lat), cd = std::cos(lat);
    auto at = [&](int c) { return cols[(size_t)c * stride + j]; };
    const double dt = at(C_T_HI) - posepoch;
    double L[3] = {ca * cd + (-sa * pma - ca * sd * pmd) * dt, sa * cd + (ca * pma - sa * sd * pmd) * dt, sd + cd * pmd * dt};
    const double mag = std::sqrt(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
    for (double& v : L) v /= mag;
    if (ss->flags & VELA_FLAG_ECLIPTIC) {
        const double se = 0.397776969112606, ce = 0.9174821430652418;
        const double y = L[1], z = L[2];
        L[1] = ce * y - se * z;
        L[2] = se * y + ce * z;
    }
    {   // solar-wind geometry AU^2 rho / (r sin rho) as a quadratic in cos(rho) around the default sky position
        const double r = at(C_RSUN);
        const double c0 = -(L[0] * at(C_SUN) + L[1] * at(C_SUN + 1) + L[2] * at(C_SUN + 2)) / r;
        const long double c = c0, w = 1.0L - c * c;
        if (r > 0.0 && std::isfinite(c0) && w > 1e-12L) {
            const long double au2_r = 499.00478383615643L * 499.00478383615643L / (long double)r;
            const long double g = acosl(c) / sqrtl(w), g1 = (c * g - 1.0L) / w, g2 = (g + 3.0L * c * g1) / w;
            cols[(size_t)(C_SWG + 0) * stride + j] = c0;
            cols[(size_t)(C_SWG + 1) * stride + j] = (double)w;
            cols[(size_t)(C_SWG + 2) * stride + j] = (double)(au2_r * g);
            cols[(size_t)(C_SWG + 3) * stride + j] = (double)(au2_r * g1);
            cols[(size_t)(C_SWG + 4) * stride + j] = (double)(au2_r * g2 / 2);
        }
    }
    for (int b = 0; b < 6; ++b) {
        const int c0 = b == 0 ? C_SUN : C_PLANET + 3 * (b - 1);
        const double r = b == 0 ? at(C_RSUN) : at(C_RPLANET + b - 1);
        const double x0 = r - (L[0] * at(c0) + L[1] * at(c0 + 1) + L[2] * at(c0 + 2));
        if (!(x0 > 0.0) || !std::isfinite(x0)) continue;
        cols[(size_t)(C_SHAP + 3 * b) * stride + j] = x0;
        cols[(size_t)(C_SHAP + 3 * b + 1) * stride + j] = 1.0 / x0;
        cols[(size_t)(C_SHAP + 3 * b + 2) * stride + j] = std::log(x0 / 499.00478383615643);
    }
}

}  // namespace vela
