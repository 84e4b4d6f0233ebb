/*
 * vela_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the reference algorithm (Vela.jl v0.1.6) for the batched
 * posterior hot path.  It exists ONLY to check the CUDA library: it may be built and
 * called from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs, and from nowhere else.  libvela_b200.so never links or calls it.
 *
 * Pinning: the reference's own implementation cannot run here (no Julia, no PINT), and
 * its tests hold identities and bounds rather than golden numbers.  This oracle is pinned
 * by tests/test_oracle_*.py against (i) the five JLSO fixtures of the reference's test
 * suite with the bounds its tests assert (test/test_NGC6440E.jl:120-124,143,
 * test/test_sim_sw_wb.jl:92-97,108, test/test_sim3_gp.jl:132-134, ...), (ii) PINT's CHI2
 * recorded in pyvela/examples/NGC6440E.par:19-20, (iii) the known-answer identities of
 * test/test_woodbury_kernel.jl:37-92, test/test_dd.jl:2-11, test/test_gp_noise.jl:23-34 and
 * (iv) an independent numpy.longdouble restatement (tests/longdouble_check.py).
 *
 * Third-party arithmetic that is not in the reference tree (unpinned in Project.toml:9-21):
 *   - GeometricUnits.jl `taylor_horner(x, cs) = sum cs[n] x^n / n!` and
 *     `taylor_horner_integral(x, cs, c0) = c0 + sum cs[n] x^(n+1)/(n+1)!`: restated in Horner form.
 *   - DoubleFloats.jl `Double64`: restated as standard double-double (Dekker/Knuth error-free
 *     transforms, QD-library algorithms).
 *   - LAPACK potrf/trtrs: restated as an unblocked Cholesky + forward substitution.
 *
 * Each function cites the reference file:line it follows (paths relative to the reference tree).
 * Build with -ffp-contract=off (Julia does not contract a*b+c).
 */
#include "../include/vela_b200.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Double-double (DoubleFloats.Double64)
 * ---------------------------------------------------------------------------------------- */
typedef struct { double hi, lo; } dd;

static inline dd two_sum(double a, double b) {
    double s = a + b, v = s - a;
    dd r = { s, (a - (s - v)) + (b - v) };
    return r;
}
static inline dd quick_two_sum(double a, double b) {
    double s = a + b;
    dd r = { s, b - (s - a) };
    return r;
}
static inline dd two_prod(double a, double b) {
    double p = a * b;
    dd r = { p, fma(a, b, -p) };
    return r;
}
static inline dd dd_from(double a) { dd r = { a, 0.0 }; return r; }
static inline dd dd_neg(dd a) { dd r = { -a.hi, -a.lo }; return r; }
static inline dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi), t = two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return quick_two_sum(s.hi, s.lo);
}
static inline dd dd_sub(dd a, dd b) { return dd_add(a, dd_neg(b)); }
static inline dd dd_add_d(dd a, double b) {
    dd s = two_sum(a.hi, b);
    s.lo += a.lo;
    return quick_two_sum(s.hi, s.lo);
}
static inline dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return quick_two_sum(p.hi, p.lo);
}
static inline dd dd_mul_d(dd a, double b) {
    dd p = two_prod(a.hi, b);
    p.lo += a.lo * b;
    return quick_two_sum(p.hi, p.lo);
}
static inline dd dd_div_d(dd a, double b) {
    double q1 = a.hi / b;
    dd p = two_prod(q1, b);
    dd s = two_sum(a.hi, -p.hi);
    double e = s.lo + a.lo - p.lo;
    double q2 = (s.hi + e) / b;
    return quick_two_sum(q1, q2);
}

/* GeometricUnits.taylor_horner(x, cs) = sum_n cs[n] x^n / n!  (FP64) */
static double taylor_horner(double x, const double* cs, int n) {
    double result = 0.0;
    for (int i = n; i >= 1; --i) result = cs[i - 1] + result * x / (double)i;
    return result;
}
/* GeometricUnits.taylor_horner_integral(x, cs, c0) = c0 + sum_n cs[n] x^(n+1)/(n+1)!  (FP64) */
static double taylor_horner_integral(double x, const double* cs, int n, double c0) {
    double result = 0.0;
    for (int i = n; i >= 1; --i) result = cs[i - 1] + result * x / (double)(i + 1);
    return c0 + result * x;
}
/* same with a Double64 argument (used by Spindown only, src/model/spindown.jl:15) */
static dd taylor_horner_integral_dd(dd x, const double* cs, int n) {
    dd result = dd_from(0.0);
    for (int i = n; i >= 1; --i)
        result = dd_add_d(dd_div_d(dd_mul(result, x), (double)(i + 1)), cs[i - 1]);
    return dd_mul(result, x);
}

/* ------------------------------------------------------------------------------------------
 * TOA records and the accumulated correction
 * ---------------------------------------------------------------------------------------- */
typedef struct {          /* src/toa/toa.jl:39-50, src/toa/wideband_toa.jl:18-21,85-88 */
    dd value;
    double error, freq;
    dd pulse_number;
    const double* eph;    /* 24 doubles: ssb_obs_pos, ssb_obs_vel, sun, jupiter, saturn, venus, uranus, neptune */
    uint64_t index;
    int wideband;
    double dm, dmerr;
} Toa;

static Toa toa_view(const void* rec, int wideband) {
    const double* w = (const double*)rec;
    Toa t;
    t.value.hi = w[0]; t.value.lo = w[1];
    t.error = w[2]; t.freq = w[3];
    t.pulse_number.hi = w[4]; t.pulse_number.lo = w[5];
    t.eph = w + 6;
    memcpy(&t.index, w + 30, 8);
    t.wideband = wideband;
    t.dm = wideband ? w[31] : 0.0;
    t.dmerr = wideband ? w[32] : 0.0;
    return t;
}

typedef struct {          /* src/toa/toa.jl:85-115, src/toa/wideband_toa.jl:27-33 */
    double delay;
    dd phase;
    double efac, equad2, spin_frequency, doppler;
    double ssb_psr_pos[3];
    double model_dm, dmefac, dmequad2;
    int bad;              /* an @assert of the reference would have fired */
} Corr;

static Corr corr_init(void) {
    Corr c;
    memset(&c, 0, sizeof c);
    c.efac = 1.0;
    c.dmefac = 1.0;
    return c;
}

static inline int toa_is_tzr(const Toa* t) { return t->index == 0; }                 /* toa.jl:57 */
static inline int toa_is_barycentered(const Toa* t) {                                /* toa.jl:64 */
    return t->eph[0] == 0.0 && t->eph[1] == 0.0 && t->eph[2] == 0.0;
}
static inline int corr_is_barycentered(const Corr* c) {                              /* toa.jl:105 */
    return !(c->ssb_psr_pos[0] == 0.0 && c->ssb_psr_pos[1] == 0.0 && c->ssb_psr_pos[2] == 0.0);
}
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
/* corrected_toa_value(toa, toacorr, Float64) toa.jl:144-145 */
static inline double t64(const Toa* t, const Corr* c) { return t->value.hi - c->delay; }
/* doppler_corrected_observing_frequency toa.jl:138-139 */
static inline double nu_c(const Toa* t, const Corr* c) { return t->freq * (1 - c->doppler); }

typedef struct {
    const VelaModelDesc* d;
    const double* P;      /* full parameter vector */
} Ctx;

#define PAR(comp, k) (ctx->P[(comp)->par[k]])

/* ------------------------------------------------------------------------------------------
 * Components
 * ---------------------------------------------------------------------------------------- */
static const double OBL = 0.4090926006005829;      /* solarsystem.jl:3 */
static const double AU = 499.00478383615643;       /* solarsystem.jl:5 */
static const double M_SUN = 4.92549094830932e-06;  /* solarsystem.jl:7-15 */
static const double M_PLANETS[5] = { 4.702819050227708e-09, 1.408128810019423e-09, 1.205680558494223e-11,
                                     2.1505895513637613e-10, 2.5373119991867603e-10 }; /* jup, sat, ven, ura, nep */

/* solar_system_shapiro_delay solarsystem.jl:31-37 */
static double ss_shapiro(double M, const double* obs_obj_pos, const double* Lhat) {
    double r = sqrt(dot3(obs_obj_pos, obs_obj_pos));
    double Ldotr = dot3(Lhat, obs_obj_pos);
    return -2 * M * log((r - Ldotr) / AU);
}

/* Correctly rounded sin/cos (double-double Taylor series after reduction by pi/2).  The sky-position unit vector
 * multiplies the ~500 s Roemer delay of EVERY TOA, so a 1-ulp library difference in these four values is a coherent
 * ~1e-13 s term in all residuals; libm (like Julia's own sincos) is faithful to < 1 ulp but not always correctly
 * rounded (glibc: ~0.3 % of arguments), which would make the oracle itself platform dependent at the 1e-9 level of
 * ln L for walkers far from the maximum.  Rounding the exact value removes that ambiguity. */
static void sincos_cr(double x, double* s_out, double* c_out) {
    if (!isfinite(x) || fabs(x) > 1e6) { *s_out = sin(x); *c_out = cos(x); return; }
    const dd pio2 = { 1.5707963267948966, 6.123233995736766e-17 };
    double k = rint(x * 0.6366197723675814);
    dd r = dd_sub(dd_from(x), dd_mul_d(pio2, k));
    r = dd_add_d(r, -k * -1.4973849048591698e-33);
    dd r2 = dd_mul(r, r);
    dd st = r, ct = { 1.0, 0.0 }, ssum = r, csum = { 1.0, 0.0 };
    for (int n = 1; n <= 15; ++n) {
        ct = dd_div_d(dd_mul(ct, r2), -(double)((2 * n - 1) * (2 * n)));
        st = dd_div_d(dd_mul(st, r2), -(double)((2 * n) * (2 * n + 1)));
        csum = dd_add(csum, ct);
        ssum = dd_add(ssum, st);
    }
    int q = ((int)(long long)k) & 3;
    double sv = ssum.hi, cv = csum.hi;
    *s_out = (q == 0) ? sv : (q == 1) ? cv : (q == 2) ? -sv : -cv;
    *c_out = (q == 0) ? cv : (q == 1) ? -sv : (q == 2) ? -cv : sv;
}

/* evaluate_proper_motion solarsystem.jl:45-64 */
static void evaluate_proper_motion(double a0, double d0, double pma, double pmd, double dt, double* out) {
    double sa, ca, sd, cd;
    sincos_cr(a0, &sa, &ca);
    sincos_cr(d0, &sd, &cd);
    double x0[3] = { ca * cd, sa * cd, sd };
    if (pma == 0.0 && pmd == 0.0) { out[0] = x0[0]; out[1] = x0[1]; out[2] = x0[2]; return; }
    double xa[3] = { -sa * pma, ca * pma, 0.0 };
    double xd[3] = { -ca * sd * pmd, -sa * sd * pmd, cd * pmd };
    double x1[3];
    for (int i = 0; i < 3; ++i) x1[i] = x0[i] + (xa[i] + xd[i]) * dt;
    double mag = sqrt(dot3(x1, x1));
    for (int i = 0; i < 3; ++i) out[i] = x1[i] / mag;
}

/* correct_toa(::SolarSystem, ...) solarsystem.jl:67-134 */
static void comp_solar_system(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    if (toa_is_barycentered(toa) || corr_is_barycentered(k)) return;
    int ecl = c->flags & VELA_FLAG_ECLIPTIC;
    double dt = t64(toa, k) - PAR(c, 5);
    double Lhat[3];
    evaluate_proper_motion(PAR(c, 0), PAR(c, 1), PAR(c, 2), PAR(c, 3), dt, Lhat);
    if (ecl) {
        double se = sin(OBL), ce = cos(OBL);
        double y = Lhat[1], z = Lhat[2];
        Lhat[1] = ce * y - se * z;
        Lhat[2] = se * y + ce * z;
    }
    const double* Rvec = toa->eph;
    double LdotR = dot3(Lhat, Rvec);
    double delay = -LdotR;
    double px = PAR(c, 4);
    if (px != 0.0) {
        double R2 = dot3(Rvec, Rvec);
        double Rperp2 = R2 - LdotR * LdotR;
        delay += 0.5 * px * Rperp2;
    }
    delay += ss_shapiro(M_SUN, toa->eph + 6, Lhat);
    if (c->flags & VELA_FLAG_PLANET_SHAPIRO)
        for (int i = 0; i < 5; ++i) delay += ss_shapiro(M_PLANETS[i], toa->eph + 9 + 3 * i, Lhat);
    double doppler = dot3(Lhat, toa->eph + 3);
    k->delay += delay;
    k->doppler += doppler;
    k->ssb_psr_pos[0] = Lhat[0]; k->ssb_psr_pos[1] = Lhat[1]; k->ssb_psr_pos[2] = Lhat[2];
}

/* DispersionComponent: delay component.jl:70-79; wideband also accumulates model_dm wideband_model.jl:16-27 */
static void apply_dispersion(const Toa* toa, Corr* k, double dm) {
    double nu = nu_c(toa, k);
    k->delay += dm / (nu * nu);
    if (toa->wideband) k->model_dm += dm;
}

/* sun_angle_and_distance solarwind.jl:10-17 + dispersion_slope solarwind.jl:37-51 */
static void comp_solar_wind(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    double dm = 0.0;
    if (!toa_is_barycentered(toa)) {
        if (!corr_is_barycentered(k)) { k->bad = 1; return; }   /* @assert solarwind.jl:12 */
        const double* rvec = toa->eph + 6;
        double r = sqrt(dot3(rvec, rvec));
        double cos_rho = -dot3(k->ssb_psr_pos, rvec) / r;
        double rho = acos(cos_rho);
        double t = t64(toa, k);
        double ne_sw = taylor_horner(t - PAR(c, 0), &PAR(c, 1), c->len[0]);
        dm = ne_sw * AU * AU * rho / (r * sin(rho));
    }
    apply_dispersion(toa, k, dm);
}

/* DispersionTaylor dispersion.jl:15-26 */
static void comp_dispersion_taylor(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    double t = t64(toa, k);
    double dm = taylor_horner(t - PAR(c, 0), &PAR(c, 1), c->len[0]);
    apply_dispersion(toa, k, dm);
}

static uint32_t mask_at(const Ctx* ctx, int mask, const Toa* toa) {
    return ctx->d->index_masks[(int64_t)mask * ctx->d->n_toas + (int64_t)(toa->index - 1)];
}

/* DispersionPiecewise dispersion.jl:43-54 */
static void comp_dmx(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    double dm = 0.0;
    if (!toa_is_tzr(toa)) {
        uint32_t idx = mask_at(ctx, c->mask[0], toa);
        if (idx != 0) dm = ctx->P[c->par[0] + idx - 1];
    }
    apply_dispersion(toa, k, dm);
}

/* ChromaticComponent delay chromatic.jl:10-12 */
static void apply_chromatic(const Toa* toa, Corr* k, double cm, double alpha) {
    k->delay += cm * pow(1e6 / nu_c(toa, k), alpha);
}
/* ChromaticTaylor chromatic.jl:23-34 */
static void comp_chromatic_taylor(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    double t = t64(toa, k);
    double cm = taylor_horner(t - PAR(c, 0), &PAR(c, 1), c->len[0]);
    apply_chromatic(toa, k, cm, PAR(c, 2));
}
/* ChromaticPiecewise chromatic.jl:46-57 */
static void comp_cmx(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    double cm = 0.0;
    if (!toa_is_tzr(toa)) {
        uint32_t idx = mask_at(ctx, c->mask[0], toa);
        if (idx != 0) cm = ctx->P[c->par[0] + idx - 1];
    }
    apply_chromatic(toa, k, cm, PAR(c, 2));
}

static const double TWO_PI = 6.283185307179586;

/* mean_anomaly / mean_motion binary/orbit.jl:3-12 */
static double mean_anomaly(const Ctx* ctx, const VelaComponentDesc* c, double dt) {
    if (c->flags & VELA_FLAG_USE_FBX) return TWO_PI * taylor_horner_integral(dt, &PAR(c, 3), c->len[0], 0.0);
    double x = dt / PAR(c, 1);
    return TWO_PI * (x - 0.5 * PAR(c, 2) * (x * x));
}
static double mean_motion(const Ctx* ctx, const VelaComponentDesc* c, double dt) {
    if (c->flags & VELA_FLAG_USE_FBX) return TWO_PI * taylor_horner(dt, &PAR(c, 3), c->len[0]);
    return TWO_PI * (1 / (PAR(c, 1) + PAR(c, 2) * dt));
}

/* shapiro_delay_params for the orthometric (H3, STIGMA) parametrisation: binary_ell1h.jl:22-28, binary_ddh.jl:18-24 */
static void shapiro_from_h3_stigma(double h3, double stigma, double* m2, double* sini) {
    *m2 = h3 / (stigma * stigma * stigma);
    *sini = 2 * stigma / (1 + stigma * stigma);
}

/* BinaryELL1 / ELL1H / ELL1k: binary/binary_ell1_base.jl:36-180, binary_ell1.jl:15, binary_ell1h.jl:22-46,
 * binary_ell1k.jl:17-49 */
static void comp_ell1(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    double dt = t64(toa, k) - PAR(c, 0);
    double a1 = PAR(c, 4) + dt * PAR(c, 5);
    double e1, e2;
    if (c->kind == VELA_COMP_BINARY_ELL1K) {
        double omdot = PAR(c, 8), lnedot = PAR(c, 9), e10 = PAR(c, 6), e20 = PAR(c, 7);
        double so = sin(omdot * dt), co = cos(omdot * dt);
        e1 = (1 + lnedot * dt) * (e10 * co + e20 * so);
        e2 = (1 + lnedot * dt) * (e20 * co - e10 * so);
    } else {
        e1 = PAR(c, 6) + dt * PAR(c, 8);
        e2 = PAR(c, 7) + dt * PAR(c, 9);
    }
    double Phi = mean_anomaly(ctx, c, dt);
    double s1 = sin(Phi), c1 = cos(Phi), s2 = sin(2 * Phi), c2 = cos(2 * Phi);
    double s3 = sin(3 * Phi), c3 = cos(3 * Phi), s4 = sin(4 * Phi), c4 = cos(4 * Phi);
    double n = mean_motion(ctx, c, dt);
    double m2, sini;
    if (c->kind == VELA_COMP_BINARY_ELL1H) shapiro_from_h3_stigma(PAR(c, 10), PAR(c, 11), &m2, &sini);
    else { m2 = PAR(c, 10); sini = PAR(c, 11); }
    double e1_2 = e1 * e1, e2_2 = e2 * e2, e1_3 = e1 * e1 * e1, e2_3 = e2 * e2 * e2;

    double dR = a1 * (s1 + 0.5 * (e2 * s2 - e1 * c2) -
                      (1.0 / 8) * ((5 * e2_2 + 3 * e1_2) * s1 - 2 * e2 * e1 * c1 + (-3 * e2_2 + 3 * e1_2) * s3 +
                                   6 * e2 * e1 * c3) -
                      (1.0 / 12) * ((5 * e2_3 + 3 * e1_2 * e2) * s2 + (-6 * e1 * e2_2 - 4 * e1_3) * c2 +
                                    (-4 * e2_3 + 12 * e1_2 * e2) * s4 + (12 * e1 * e2_2 - 4 * e1_3) * c4));
    if (c->kind == VELA_COMP_BINARY_ELL1K) dR = dR - 1.5 * a1 * e1;        /* binary_ell1k.jl:46-49 */
    double dRp = a1 * (c1 + e1 * s2 + e2 * c2 -
                       (1.0 / 8) * ((5 * e2_2 + 3 * e1_2) * c1 + 2 * e1 * e2 * s1 + (-9 * e2_2 + 9 * e1_2) * c3 -
                                    18 * e1 * e2 * s3) -
                       (1.0 / 12) * ((10 * e2_3 + 6 * e1_2 * e2) * c2 + (12 * e1 * e2_2 + 8 * e1_3) * s2 +
                                     (-16 * e2_3 + 48 * e1_2 * e2) * c4 + (-48 * e1 * e2_2 + 16 * e1_3) * s4));
    double dRp2 = a1 * (-s1 + 2 * e1 * c2 - 2 * e2 * s2 -
                        (1.0 / 8) * ((-5 * e2_2 - 3 * e1_2) * s1 + 2 * e1 * e2 * c1 + (27 * e2_2 - 27 * e1_2) * s3 -
                                     54 * e1 * e2 * c3) -
                        (1.0 / 12) * ((-20 * e2_3 - 12 * e1_2 * e2) * s2 + (24 * e1 * e2_2 + 16 * e1_3) * c2 +
                                      (64 * e2_3 - 192 * e1_2 * e2) * s4 + (-192 * e1 * e2_2 + 64 * e1_3) * c4));
    double dR_inv = dR * (1 - n * dRp + n * n * dRp * dRp + 0.5 * n * n * dR * dRp2);
    double dS = -2 * m2 * log(1 - sini * s1);
    if (c->kind == VELA_COMP_BINARY_ELL1H) {                              /* binary_ell1h.jl:30-46 */
        double cbar = sqrt(1 - sini * sini);
        double sg = sini / (1 + cbar);
        double a0 = -log(1 + sg * sg), b1 = -2 * sg, a2 = sg * sg;
        double dS012 = -2 * m2 * (a0 + b1 * s1 + a2 * c2);
        dS = dS - dS012;
    }
    k->delay += dR_inv + dS;
    k->doppler += -dRp * n;
}

/* mikkola binary/orbit.jl:21-74 */
static double mikkola(double l, double e, int* bad) {
    if (e == 0.0 || l == 0.0) return l;
    if (!(0.0 < e && e < 1.0)) { *bad = 1; return NAN; }
    double sgn = (l > 0) - (l < 0);
    l *= sgn;
    double ncycles = floor(l / TWO_PI);
    l -= TWO_PI * ncycles;
    int flag = l > M_PI;
    if (flag) l = TWO_PI - l;
    double alpha = (1 - e) / (4 * e + 0.5);
    double alpha3 = alpha * alpha * alpha;
    double beta = (l / 2) / (4 * e + 0.5);
    double beta2 = beta * beta;
    double z = (beta > 0) ? cbrt(beta + sqrt(alpha3 + beta2)) : cbrt(beta - sqrt(alpha3 + beta2));
    double s = z - alpha / z;
    double s5 = s * s * s * s * s;
    double w = s - (0.078 * s5) / (1 + e);
    double w3 = w * w * w;
    double E0 = l + e * (3 * w - 4 * w3);
    double u = E0;
    double su = sin(u), cu = cos(u);
    double esu = e * su, ecu = e * cu;
    double fu = u - esu - l, f1u = 1 - ecu, f2u = esu, f3u = ecu, f4u = -esu;
    double u1 = -fu / f1u;
    double u2 = -fu / (f1u + f2u * u1 / 2);
    double u3 = -fu / (f1u + f2u * u2 / 2 + f3u * (u2 * u2) / 6.0);
    double u4 = -fu / (f1u + f2u * u3 / 2 + f3u * (u3 * u3) / 6.0 + f4u * (u3 * u3 * u3) / 24.0);
    double xi = E0 + u4;
    double sol = flag ? (TWO_PI - xi) : xi;
    return sgn * (sol + ncycles * TWO_PI);
}

/* kopeikin_I0_J0 + kopeikin_corrections binary/binary_ddk.jl:72-120 */
static void kopeikin_corrections(double dt, double x, double kin, double kom, double mua, double mud, double px,
                                 const double* ssb_obs_pos, const double* L, double* dx, double* dom, double* dkin) {
    double si = sin(kin), ci = cos(kin), sO = sin(kom), cO = cos(kom);
    double cot = ci / si, csc = 1 / si;
    double dkin_pm = (-mua * sO + mud * cO) * dt;
    double dx_pm = x * cot * dkin_pm;
    double dom_pm = csc * (mua * cO + mud * sO) * dt;
    double sind = L[2];
    double cosd = sqrt(1 - sind * sind);
    double cosa = L[0] / cosd, sina = L[1] / cosd;
    double I0[3] = { -sina, cosa, 0.0 };
    double J0[3] = { -cosa * sind, -sina * sind, cosd };
    double dI0 = dot3(ssb_obs_pos, I0), dJ0 = dot3(ssb_obs_pos, J0);
    double dx_px = x * cot * px * (dI0 * sO - dJ0 * cO);
    double dom_px = -csc * px * (dI0 * cO + dJ0 * sO);
    *dx = dx_pm + dx_px;
    *dom = dom_pm + dom_px;
    *dkin = dkin_pm;
}

/* BinaryDD / DDH / DDS / DDK: binary/binary_dd_base.jl:36-162, binary_dd.jl:15, binary_ddh.jl:18-24,
 * binary_dds.jl:18-23, binary_ddk.jl:20-66 */
static void comp_dd(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    double dt = t64(toa, k) - PAR(c, 0);
    double n = mean_motion(ctx, c, dt);
    double l = mean_anomaly(ctx, c, dt);
    double et = PAR(c, 6) + dt * PAR(c, 7);
    double er = et * (1 + PAR(c, 10));
    double ephi = et * (1 + PAR(c, 11));
    double u = mikkola(l, et, &k->bad);
    double sinu = sin(u), cosu = cos(u);
    double eta = sqrt(1 - ephi * ephi);
    double bphi = (1 - eta) / ephi;
    double v_u = 2 * atan2(bphi * sinu, 1 - bphi * cosu);
    double v = v_u + u;
    double a1 = PAR(c, 4) + dt * PAR(c, 5);
    double m2, sini, dom = 0.0;
    if (c->kind == VELA_COMP_BINARY_DDH) shapiro_from_h3_stigma(PAR(c, 13), PAR(c, 14), &m2, &sini);
    else if (c->kind == VELA_COMP_BINARY_DDS) { m2 = PAR(c, 13); sini = 1 - exp(-PAR(c, 14)); }
    else if (c->kind == VELA_COMP_BINARY_DDK) {
        double dx, dkin, kin = PAR(c, 14);
        kopeikin_corrections(dt, a1, kin, PAR(c, 15), PAR(c, 16), PAR(c, 17), PAR(c, 18), toa->eph, k->ssb_psr_pos,
                             &dx, &dom, &dkin);
        a1 += dx;
        kin += dkin;
        m2 = PAR(c, 13);
        sini = sin(kin);
    } else { m2 = PAR(c, 13); sini = PAR(c, 14); }
    double kk = PAR(c, 9) / n;
    double om = PAR(c, 8) + kk * v + dom;
    double sinw = sin(om), cosw = cos(om);
    double alpha = a1 * sinw;
    double beta = a1 * eta * cosw;
    double gamma = PAR(c, 12);

    double dRE = alpha * (cosu - er) + (beta + gamma) * sinu;
    double dREp = -alpha * sinu + (beta + gamma) * cosu;
    double dREp2 = -alpha * cosu - (beta + gamma) * sinu;
    double nhat = n / (1 - et * cosu);
    double dRE_inv = dRE * (1 - nhat * dREp + (nhat * nhat * dREp * dREp) + 0.5 * nhat * nhat * dRE * dREp2 -
                            0.5 * et * sinu / (1 - et * cosu) * nhat * nhat * dRE * dREp);
    double dS = -2 * m2 * log(1 - et * cosu - (sini / a1) * (alpha * (cosu - er) + beta * sinu));
    k->delay += dRE_inv + dS;
    k->doppler += dREp * nhat;
}

/* FrequencyDependent frequency_dependent.jl:22-34 / FrequencyDependentJump :46-81 */
static double int_pow(double x, unsigned p) { double r = 1.0; for (unsigned i = 0; i < p; ++i) r *= x; return r; }
static void comp_fd(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    double lam = log(nu_c(toa, k) / 1e9);
    double sum = 0.0;
    for (int p = 1; p <= c->len[0]; ++p) sum += ctx->P[c->par[0] + p - 1] * int_pow(lam, (unsigned)p);
    k->delay += sum;
}
static void comp_fdjump(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    if (toa_is_tzr(toa)) return;
    double lam = log(nu_c(toa, k) / 1e9);
    double delay = 0.0;
    for (int g = 0; g < c->len[0]; ++g) {
        if (ctx->d->bit_masks[(int64_t)(c->mask[0] + g) * ctx->d->n_toas + (int64_t)(toa->index - 1)]) {
            unsigned p = ctx->d->index_masks[(int64_t)c->mask[1] * ctx->d->n_toas + g];
            delay += ctx->P[c->par[0] + g] * int_pow(lam, p);
        }
    }
    k->delay += delay;
}

/* evaluate_xwavex wavex.jl:28-42 */
static double xwavex(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, const Corr* k) {
    double t_t0 = t64(toa, k) - PAR(c, 0);
    double kk = 2 * M_PI * t_t0;
    double sum = 0.0;
    for (int i = 0; i < c->len[0]; ++i) {
        double phi = kk * ctx->P[c->par[3] + i];
        sum += ctx->P[c->par[1] + i] * sin(phi) + ctx->P[c->par[2] + i] * cos(phi);
    }
    return sum;
}

static double powerlaw(double A, double gamma, double f, double f1);

/* evaluate_powerlaw_red_noise_gp gp_noise.jl:23-58 with the free (un-marginalised) amplitudes.
 * aux holds ln_js[0..N) then js[0..N) = exp(ln_j) evaluated in the field's storage type (include/vela_b200.h). */
static double powerlaw_gp(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, const Corr* k) {
    const int N = c->len[0], nlog = c->len[1];
    const double* ln_js = ctx->d->aux_values + c->len[2];
    const double* js = ln_js + N;
    const double* alpha = ctx->P + c->par[2];
    const double* beta = ctx->P + c->par[3];
    double A = pow(10.0, PAR(c, 0)), gamma = PAR(c, 1), f1 = PAR(c, 4);
    double dt = t64(toa, k) - PAR(c, 5);
    double phi1 = 2 * M_PI * f1 * dt;
    double e_re = cos(phi1), e_im = sin(phi1);      /* exp(im * phi1) */
    double sigma1 = sqrt(powerlaw(A, gamma, f1, f1));
    double result = 0.0;
    for (int i = 0; i < nlog; ++i) {
        double jfac = exp(-(gamma / 2) * ln_js[i]);
        double ph = js[i] * phi1;
        result += jfac * (alpha[i] * sin(ph) + beta[i] * cos(ph));
    }
    double re = e_re, im = e_im;
    for (int i = nlog; i < N; ++i) {
        double jfac = exp(-(gamma / 2) * ln_js[i]);
        result += jfac * (alpha[i] * im + beta[i] * re);
        double nre = re * e_re - im * e_im, nim = re * e_im + im * e_re;   /* Complex * Complex */
        re = nre; im = nim;
    }
    return sigma1 * result;
}

/* SolarWindDispersionPiecewise solarwind.jl:72-91 */
static void comp_swx(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    double dm = 0.0;
    if (!toa_is_tzr(toa) && !toa_is_barycentered(toa)) {
        uint32_t idx = mask_at(ctx, c->mask[0], toa);
        if (idx != 0) {
            if (!corr_is_barycentered(k)) { k->bad = 1; return; }
            const double* rvec = toa->eph + 6;
            double r = sqrt(dot3(rvec, rvec));
            double rho = acos(-dot3(k->ssb_psr_pos, rvec) / r);
            double geometry = AU * AU * rho / (r * sin(rho));
            dm = ctx->P[c->par[0] + idx - 1] * (geometry - c->dval[1]) / (c->dval[0] - c->dval[1]);
        }
    }
    apply_dispersion(toa, k, dm);
}

/* DispersionOffset / ExclusiveDispersionOffset (FDJUMPDM) fdjumpdm.jl:14-66: a full DispersionComponent */
static void comp_dmoffset(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    double dm = 0.0;
    if (!toa_is_tzr(toa)) {
        if (c->kind == VELA_COMP_DM_OFFSET_EXCLUSIVE) {
            uint32_t idx = mask_at(ctx, c->mask[0], toa);
            if (idx != 0) dm = -ctx->P[c->par[0] + idx - 1];
        } else {
            double dot = 0.0;
            for (int g = 0; g < c->len[0]; ++g)
                dot += ctx->P[c->par[0] + g] *
                       (double)ctx->d->bit_masks[(int64_t)(c->mask[0] + g) * ctx->d->n_toas + (int64_t)(toa->index - 1)];
            dm = -dot;
        }
    }
    apply_dispersion(toa, k, dm);
}

/* Glitch glitch.jl:14-38 */
static void comp_glitch(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    double t = t64(toa, k);
    double sum = 0.0;
    for (int g = 0; g < c->len[0]; ++g) {
        double t_gl = ctx->P[c->par[0] + g], ph = ctx->P[c->par[1] + g], f0 = ctx->P[c->par[2] + g];
        double f1 = ctx->P[c->par[3] + g], f2 = ctx->P[c->par[4] + g], fd = ctx->P[c->par[5] + g];
        double tau = ctx->P[c->par[6] + g];
        double dt = t - t_gl;
        double term = (t < t_gl) ? 0.0
                                 : (ph + dt * (f0 + dt * f1 / 2 + dt * dt * f2 / 6) + fd * tau * (1 - exp(-dt / tau)));
        sum += term;
    }
    k->phase = dd_add_d(k->phase, sum);
}

/* ChromaticExponentialDip expdip.jl:11-31 */
static void comp_expdip(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    double t = t64(toa, k);
    double f = nu_c(toa, k);
    double fref = PAR(c, 0), eps = PAR(c, 1);
    double sum = 0.0;
    for (int g = 0; g < c->len[0]; ++g) {
        double T = ctx->P[c->par[2] + g], A = ctx->P[c->par[3] + g], gam = ctx->P[c->par[4] + g];
        double tau = ctx->P[c->par[5] + g];
        double dt = t - T;
        sum += -A * pow(f / fref, gam) * pow(tau / eps, eps / tau) * pow(tau / (tau - eps), (tau - eps) / tau) *
               exp(-dt / tau) / (1 + exp(-dt / eps));
    }
    k->delay += sum;
}

/* Spindown spindown.jl:15-33 */
static void comp_spindown(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    dd dt = dd_add_d(toa->value, -k->delay);               /* corrected_toa_value toa.jl:142-143 */
    double f_ = PAR(c, 0);
    const double* fs = &PAR(c, 1);
    dd phase = dd_add(dd_mul_d(dt, f_), taylor_horner_integral_dd(dt, fs, c->len[0]));
    double dsf = f_ + taylor_horner(dt.hi, fs, c->len[0]);
    k->phase = dd_add(k->phase, phase);
    k->spin_frequency += dsf;
}

/* PhaseOffset phase_offset.jl:12-13 */
static void comp_phoff(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    double ph = -(double)(!toa_is_tzr(toa)) * PAR(c, 0);
    k->phase = dd_add_d(k->phase, ph);
}

/* ExclusivePhaseJump jump.jl:31-45 */
static void comp_jump_excl(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    double ph = 0.0;
    if (!toa_is_tzr(toa)) {
        uint32_t idx = mask_at(ctx, c->mask[0], toa);
        if (idx != 0) ph = ctx->P[c->par[0] + idx - 1] * (PAR(c, 1) + PAR(c, 2));
    }
    k->phase = dd_add_d(k->phase, ph);
}
/* PhaseJump (BitMatrix) jump.jl:14-22 */
static void comp_jump_bits(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    double ph = 0.0;
    if (!toa_is_tzr(toa)) {
        double dot = 0.0;
        for (int g = 0; g < c->len[0]; ++g)
            dot += ctx->P[c->par[0] + g] *
                   (double)ctx->d->bit_masks[(int64_t)(c->mask[0] + g) * ctx->d->n_toas + (int64_t)(toa->index - 1)];
        ph = dot * (PAR(c, 1) + PAR(c, 2));
    }
    k->phase = dd_add_d(k->phase, ph);
}

/* MeasurementNoise measurement_noise.jl:20-38 */
static void comp_measurement_noise(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    if (toa_is_tzr(toa)) return;
    uint32_t ie = mask_at(ctx, c->mask[0], toa), iq = mask_at(ctx, c->mask[1], toa);
    double efac = (ie == 0) ? 1.0 : ctx->P[c->par[0] + ie - 1];
    double equad = (iq == 0) ? 0.0 : ctx->P[c->par[1] + iq - 1];
    k->efac *= efac;
    k->equad2 += equad * equad;
}

/* DispersionJump / ExclusiveDispersionJump dmjump.jl:5-86: DM only, no delay */
static void comp_dmjump_excl(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    if (!toa->wideband || toa_is_tzr(toa)) return;
    uint32_t idx = mask_at(ctx, c->mask[0], toa);
    if (idx != 0) k->model_dm += -ctx->P[c->par[0] + idx - 1];
}
static void comp_dmjump_bits(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    if (!toa->wideband || toa_is_tzr(toa)) return;
    double dot = 0.0;
    for (int g = 0; g < c->len[0]; ++g)
        dot += ctx->P[c->par[0] + g] *
               (double)ctx->d->bit_masks[(int64_t)(c->mask[0] + g) * ctx->d->n_toas + (int64_t)(toa->index - 1)];
    k->model_dm += -dot;
}

/* DispersionMeasurementNoise dispersion_measurement_noise.jl:19-46 (no-op for narrowband TOAs) */
static void comp_dm_noise(const Ctx* ctx, const VelaComponentDesc* c, const Toa* toa, Corr* k) {
    if (!toa->wideband) return;
    uint32_t ie = mask_at(ctx, c->mask[0], toa), iq = mask_at(ctx, c->mask[1], toa);
    double dmefac = (ie == 0) ? 1.0 : ctx->P[c->par[0] + ie - 1];
    double dmequad = (iq == 0) ? 0.0 : ctx->P[c->par[1] + iq - 1];
    k->dmefac *= dmefac;
    k->dmequad2 += dmequad * dmequad;
}

/* correct_toa(components::Tuple, ...) timing_model.jl:85-98 */
static Corr correct_toa(const Ctx* ctx, const Toa* toa) {
    Corr k = corr_init();
    for (int i = 0; i < ctx->d->n_components; ++i) {
        const VelaComponentDesc* c = &ctx->d->components[i];
        switch (c->kind) {
            case VELA_COMP_SOLAR_SYSTEM: comp_solar_system(ctx, c, toa, &k); break;
            case VELA_COMP_SOLAR_WIND: comp_solar_wind(ctx, c, toa, &k); break;
            case VELA_COMP_DISPERSION_TAYLOR: comp_dispersion_taylor(ctx, c, toa, &k); break;
            case VELA_COMP_DISPERSION_PIECEWISE: comp_dmx(ctx, c, toa, &k); break;
            case VELA_COMP_CHROMATIC_TAYLOR: comp_chromatic_taylor(ctx, c, toa, &k); break;
            case VELA_COMP_CHROMATIC_PIECEWISE: comp_cmx(ctx, c, toa, &k); break;
            case VELA_COMP_BINARY_ELL1: case VELA_COMP_BINARY_ELL1H: case VELA_COMP_BINARY_ELL1K:
                comp_ell1(ctx, c, toa, &k); break;
            case VELA_COMP_BINARY_DD: case VELA_COMP_BINARY_DDH: case VELA_COMP_BINARY_DDS: case VELA_COMP_BINARY_DDK:
                comp_dd(ctx, c, toa, &k); break;
            case VELA_COMP_FD: comp_fd(ctx, c, toa, &k); break;
            case VELA_COMP_FD_JUMP: comp_fdjump(ctx, c, toa, &k); break;
            case VELA_COMP_WAVEX: k.delay += xwavex(ctx, c, toa, &k); break;
            case VELA_COMP_DM_WAVEX: apply_dispersion(toa, &k, xwavex(ctx, c, toa, &k)); break;
            case VELA_COMP_CM_WAVEX: apply_chromatic(toa, &k, xwavex(ctx, c, toa, &k), PAR(c, 4)); break;
            case VELA_COMP_POWERLAW_RED_GP: k.delay += powerlaw_gp(ctx, c, toa, &k); break;               /* gp_noise.jl:120-130 */
            case VELA_COMP_POWERLAW_DM_GP:                                                                /* gp_noise.jl:165-183 */
                apply_dispersion(toa, &k, (1.4e9 * 1.4e9) * powerlaw_gp(ctx, c, toa, &k)); break;
            case VELA_COMP_POWERLAW_CHROM_GP:                                                             /* gp_noise.jl:219-235 */
                apply_chromatic(toa, &k, pow(1400.0, PAR(c, 6)) * powerlaw_gp(ctx, c, toa, &k), PAR(c, 6)); break;
            case VELA_COMP_SOLAR_WIND_PIECEWISE: comp_swx(ctx, c, toa, &k); break;
            case VELA_COMP_DM_OFFSET_EXCLUSIVE: case VELA_COMP_DM_OFFSET: comp_dmoffset(ctx, c, toa, &k); break;
            case VELA_COMP_GLITCH: comp_glitch(ctx, c, toa, &k); break;
            case VELA_COMP_EXP_DIP: comp_expdip(ctx, c, toa, &k); break;
            case VELA_COMP_SPINDOWN: comp_spindown(ctx, c, toa, &k); break;
            case VELA_COMP_PHASE_OFFSET: comp_phoff(ctx, c, toa, &k); break;
            case VELA_COMP_PHASE_JUMP_EXCLUSIVE: comp_jump_excl(ctx, c, toa, &k); break;
            case VELA_COMP_PHASE_JUMP: comp_jump_bits(ctx, c, toa, &k); break;
            case VELA_COMP_MEASUREMENT_NOISE: comp_measurement_noise(ctx, c, toa, &k); break;
            case VELA_COMP_DM_JUMP_EXCLUSIVE: comp_dmjump_excl(ctx, c, toa, &k); break;
            case VELA_COMP_DM_JUMP: comp_dmjump_bits(ctx, c, toa, &k); break;
            case VELA_COMP_DM_MEASUREMENT_NOISE: comp_dm_noise(ctx, c, toa, &k); break;
            default: k.bad = 1; break;
        }
    }
    /* TOACorrection constructor asserts toa.jl:95-99 */
    if (!(fabs(k.doppler) < 1.0) || !(k.spin_frequency >= 0.0)) k.bad = 1;
    return k;
}

/* read_params parameter.jl:158-169 */
static void read_params(const VelaModelDesc* d, const double* free_values, int64_t stride, double* P) {
    memcpy(P, d->default_values, sizeof(double) * d->n_params);
    for (int i = 0; i < d->n_free; ++i) P[d->free_indices[i]] = free_values[i * stride];
}

/* calc_tzr_phase residuals.jl:29-32 */
static dd calc_tzr_phase(const Ctx* ctx, int* bad) {
    Toa tzr = toa_view(ctx->d->tzr_toa, 0);
    Corr k = correct_toa(ctx, &tzr);
    *bad |= k.bad;
    return k.phase;
}

typedef struct { double tres, err2, dmres, dmerr2; dd dphase; double f; int bad; } Resid;

/* form_residual residuals.jl:8-12; wideband_residuals.jl:6-17; scaled errors toa.jl:122-123, wideband_toa.jl:68-69 */
static Resid form_residual(const Ctx* ctx, const Toa* toa, dd tzrphase) {
    Corr k = correct_toa(ctx, toa);
    Resid r;
    r.dphase = dd_sub(dd_sub(k.phase, toa->pulse_number), tzrphase);
    r.f = k.spin_frequency * (1 + k.doppler);              /* doppler_shifted_spin_frequency toa.jl:131-134 */
    r.tres = r.dphase.hi / r.f;
    r.err2 = (toa->error * toa->error + k.equad2) * k.efac * k.efac;
    r.dmres = toa->dm - k.model_dm;
    r.dmerr2 = (toa->dmerr * toa->dmerr + k.dmequad2) * k.dmefac * k.dmefac;
    r.bad = k.bad || k.spin_frequency == 0.0;
    return r;
}

static const void* toa_rec(const VelaModelDesc* d, int64_t j) {
    return (const char*)d->toas + j * (int64_t)d->toa_stride_bytes;
}

/* _calc_resids_and_Ninvdiag gls_likelihood.jl:9-79 */
static int resids_and_ninvdiag(const Ctx* ctx, double* y, double* w) {
    const VelaModelDesc* d = ctx->d;
    int bad = 0;
    dd tzr = calc_tzr_phase(ctx, &bad);
    int64_t n = d->n_toas;
    for (int64_t j = 0; j < n; ++j) {
        Toa toa = toa_view(toa_rec(d, j), d->wideband);
        Resid r = form_residual(ctx, &toa, tzr);
        bad |= r.bad;
        y[j] = r.tres;
        w[j] = 1.0 / r.err2;
        if (d->wideband) { y[n + j] = r.dmres; w[n + j] = 1.0 / r.dmerr2; }
    }
    return bad;
}

/* EcorrGroup {start, stop, index}, kernel.jl:25-29 (1-based, inclusive) */
typedef struct { int64_t start, stop, index; } EGroup;
static EGroup egroup(const VelaModelDesc* d, int g) {
    EGroup e = { (int64_t)d->ecorr_groups[3 * g], (int64_t)d->ecorr_groups[3 * g + 1], (int64_t)d->ecorr_groups[3 * g + 2] };
    return e;
}
static double ecorr_weight(const VelaModelDesc* d, const double* P, const EGroup* e) {
    double ecorr = (e->index == 0) ? 0.0 : P[d->par_ecorr + e->index - 1];
    return ecorr * ecorr;
}
static int has_ecorr(const VelaModelDesc* d) {
    return d->kernel_kind == VELA_KERNEL_ECORR || d->kernel_kind == VELA_KERNEL_WOODBURY_ECORR;
}

/* powerlaw gp_noise.jl:12-16 */
static double powerlaw(double A, double gamma, double f, double f1) {
    double fyr = 1.0 / 3600 / 24 / 365.25;
    double denom = 12 * (M_PI * M_PI) * fyr * fyr * fyr;
    return A * A / denom * pow(fyr / f, gamma) * f1;
}

/* calc_noise_weights_inv gls_likelihood.jl:3-7; gp_noise.jl:64-78; marginalized_timing_model.jl:24 */
static void noise_weights_inv(const Ctx* ctx, double* phiinv) {
    const VelaModelDesc* d = ctx->d;
    int64_t off = 0;
    for (int g = 0; g < d->n_gp; ++g) {
        const VelaGPDesc* gp = &d->gp[g];
        if (gp->kind == VELA_GP_MARGINALIZED_TM) {
            for (int i = 0; i < gp->n; ++i) phiinv[off + i] = gp->weights_inv[i];
            off += gp->n;
        } else {
            double A = pow(10.0, ctx->P[gp->par_log10_amp]);   /* exp10 */
            double gamma = ctx->P[gp->par_gamma], f1 = ctx->P[gp->par_f1];
            double P1 = powerlaw(A, gamma, f1, f1);
            for (int i = 0; i < gp->n; ++i) {
                double v = 1 / (P1 * exp(-gamma * gp->ln_js[i]));
                phiinv[off + i] = v;
                phiinv[off + gp->n + i] = v;
            }
            off += 2 * (int64_t)gp->n;
        }
    }
}

/* _calc_Ninv_M gls_likelihood.jl:171-199 + _calc_Sigmainv__and__MT_Ninv_y gls_likelihood.jl:101-169.
 * Sigma is returned as a full column-major p x p with only the upper triangle (q <= p) meaningful. */
static void sigma_and_u(const VelaModelDesc* d, const double* Pfull, const double* w, const double* phiinv, const double* y,
                        double* NinvM, double* Sig, double* u) {
    int64_t n = d->basis_rows, p = d->basis_cols, ld = d->basis_ld;
    const double* M = d->noise_basis;
    if (!has_ecorr(d)) {
        for (int64_t c = 0; c < p; ++c)
            for (int64_t j = 0; j < n; ++j) NinvM[c * n + j] = M[c * ld + j] * w[j];
    } else {
        /* _calc_Ninv_M(::EcorrKernel, ...) gls_ecorr_likelihood.jl:66-137: per-group Sherman-Morrison */
        for (int g = 0; g < d->n_ecorr_groups; ++g) {
            EGroup e = egroup(d, g);
            double wt = ecorr_weight(d, Pfull, &e);
            double Q = 0.0;
            for (int64_t i = e.start - 1; i < e.stop; ++i) Q += w[i];
            double alpha = wt / (1 + wt * Q);
            for (int64_t c = 0; c < p; ++c) {
                double Pp = 0.0;
                for (int64_t i = e.start - 1; i < e.stop; ++i) Pp += M[c * ld + i] * w[i];
                double R = Pp * alpha;
                for (int64_t i = e.start - 1; i < e.stop; ++i) NinvM[c * n + i] = (M[c * ld + i] - R) * w[i];
            }
        }
    }
    for (int64_t c = 0; c < p; ++c)
        for (int64_t q = 0; q <= c; ++q) {
            double s = (c == q) ? phiinv[c] : 0.0;
            const double* mc = M + c * ld;
            const double* nq = NinvM + q * n;
            for (int64_t j = 0; j < n; ++j) s += mc[j] * nq[j];
            Sig[c * p + q] = s;
        }
    for (int64_t c = 0; c < p; ++c) {
        double s = 0.0;
        const double* nc = NinvM + c * n;
        for (int64_t j = 0; j < n; ++j) s += nc[j] * y[j];
        u[c] = s;
    }
}

/* cholesky!(Symmetric(Sigma, :U)) + logdet + ldiv!(L, u) + dot — gls_likelihood.jl:211-221.
 * Upper-stored column-major: factor U^T U in place; returns 0 if not positive definite. */
static int chol_upper(double* A, int64_t p) {
    for (int64_t j = 0; j < p; ++j) {
        double s = A[j * p + j];
        for (int64_t k = 0; k < j; ++k) s -= A[j * p + k] * A[j * p + k];
        if (!(s > 0.0) || !isfinite(s)) return 0;
        double ujj = sqrt(s);
        A[j * p + j] = ujj;
        for (int64_t i = j + 1; i < p; ++i) {
            double t = A[i * p + j];
            for (int64_t k = 0; k < j; ++k) t -= A[i * p + k] * A[j * p + k];
            A[i * p + j] = t / ujj;
        }
    }
    return 1;
}

/* _gls_lnlike_serial gls_likelihood.jl:201-226 */
static double gls_lnlike(const VelaModelDesc* d, const double* Pfull, const double* w, const double* phiinv, const double* y,
                         double* NinvM, double* Sig, double* u) {
    int64_t n = d->basis_rows, p = d->basis_cols;
    sigma_and_u(d, Pfull, w, phiinv, y, NinvM, Sig, u);
    double yNy = 0.0, logdetN = 0.0;
    if (!has_ecorr(d)) {
        /* _calc_y_Ninv_y__and__logdet_N gls_likelihood.jl:81-99 */
        for (int64_t j = 0; j < n; ++j) yNy += y[j] * y[j] * w[j];
        for (int64_t j = 0; j < n; ++j) logdetN += log(w[j]);
        logdetN = -logdetN;
    } else {
        /* _calc_y_Ninv_y__and__logdet_N(::EcorrKernel, ...) gls_ecorr_likelihood.jl:1-64 (serial branch) */
        for (int g = 0; g < d->n_ecorr_groups; ++g) {
            EGroup e = egroup(d, g);
            double wt = ecorr_weight(d, Pfull, &e);
            double r_r = 0.0, r_u = 0.0, u_u = 0.0, ldn = 0.0;
            for (int64_t ii = e.start - 1; ii < e.stop; ++ii) {
                double Ninv_ii = w[ii];
                double r_u_ii = y[ii] * Ninv_ii;
                r_r += y[ii] * r_u_ii;
                r_u += r_u_ii;
                u_u += Ninv_ii;
                ldn -= log(w[ii]);
            }
            double denom = 1 + wt * u_u;
            yNy += r_r - wt * r_u * r_u / denom;
            logdetN += ldn + log(denom);
        }
    }
    if (!chol_upper(Sig, p)) return -INFINITY;
    double logdetS = 0.0;
    for (int64_t j = 0; j < p; ++j) logdetS += log(Sig[j * p + j]);
    logdetS *= 2;
    /* solve U^T z = u (L = U^T) */
    double zz = 0.0;
    for (int64_t i = 0; i < p; ++i) {
        double t = u[i];
        for (int64_t k = 0; k < i; ++k) t -= Sig[i * p + k] * u[k];
        u[i] = t / Sig[i * p + i];
        zz += u[i] * u[i];
    }
    double logdetPhi = 0.0;
    for (int64_t j = 0; j < p; ++j) logdetPhi += log(phiinv[j]);
    logdetPhi = -logdetPhi;
    return -0.5 * (yNy - zz + logdetN + logdetPhi + logdetS);
}

/* _ecorr_lnlike_group ecorr_likelihood.jl:8-38 / _ecorr_chi2_group ecorr_chi2.jl:6-30, summed over groups
 * (calc_lnlike_serial ecorr_likelihood.jl:62-76, calc_chi2_serial ecorr_chi2.jl:58-72) */
static double ecorr_white_sum(const Ctx* ctx, int chi2_only, int* bad) {
    const VelaModelDesc* d = ctx->d;
    dd tzr = calc_tzr_phase(ctx, bad);
    double total = 0.0;
    for (int g = 0; g < d->n_ecorr_groups; ++g) {
        EGroup e = egroup(d, g);
        double w = ecorr_weight(d, ctx->P, &e);
        double r_r = 0.0, r_u = 0.0, u_u = 0.0, norm = 0.0;
        for (int64_t ii = e.start; ii <= e.stop; ++ii) {
            Toa toa = toa_view(toa_rec(d, ii - 1), 0);
            Resid r = form_residual(ctx, &toa, tzr);
            *bad |= r.bad;
            r_r += r.tres * r.tres / r.err2;
            r_u += r.tres / r.err2;
            u_u += 1 / r.err2;
            norm += log(r.err2);
        }
        double term = r_r - w * r_u * r_u / (1 + w * u_u);
        if (!chi2_only) term += norm + log(1 + w * u_u);
        total += term;
    }
    return total;
}

typedef struct { double *P, *y, *w, *phiinv, *NinvM, *Sig, *u; } Work;

static int work_alloc(const VelaModelDesc* d, Work* k) {
    int64_t n = d->wideband ? 2 * d->n_toas : d->n_toas;
    int64_t p = (d->kernel_kind == VELA_KERNEL_WOODBURY_WHITE || d->kernel_kind == VELA_KERNEL_WOODBURY_ECORR) ? d->basis_cols : 0;
    memset(k, 0, sizeof *k);
    k->P = malloc(sizeof(double) * (d->n_params + 1));
    k->y = malloc(sizeof(double) * (n + 1));
    k->w = malloc(sizeof(double) * (n + 1));
    k->phiinv = malloc(sizeof(double) * (p + 1));
    k->NinvM = malloc(sizeof(double) * (n * p + 1));
    k->Sig = malloc(sizeof(double) * (p * p + 1));
    k->u = malloc(sizeof(double) * (p + 1));
    return k->P && k->y && k->w && k->phiinv && k->NinvM && k->Sig && k->u;
}
static void work_free(Work* k) {
    free(k->P); free(k->y); free(k->w); free(k->phiinv); free(k->NinvM); free(k->Sig); free(k->u);
}

/* calc_lnlike_serial: wls_likelihood.jl:7-14,49-56; wideband_wls_likelihood.jl:1-21; gls_likelihood.jl:249-262 */
static double lnlike_one(const VelaModelDesc* d, Work* k, int chi2_only) {
    Ctx ctx = { d, k->P };
    int64_t n = d->n_toas;
    if (d->kernel_kind == VELA_KERNEL_ECORR) {
        int bad = 0;
        double sum = ecorr_white_sum(&ctx, chi2_only, &bad);
        double res = chi2_only ? sum : -sum / 2;
        if (bad || isnan(res)) return chi2_only ? NAN : -INFINITY;
        return res;
    }
    if (d->kernel_kind == VELA_KERNEL_WHITE || chi2_only) {
        int bad = 0;
        dd tzr = calc_tzr_phase(&ctx, &bad);
        double sum = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            Toa toa = toa_view(toa_rec(d, j), d->wideband);
            Resid r = form_residual(&ctx, &toa, tzr);
            bad |= r.bad;
            double term = r.tres * r.tres / r.err2 + (chi2_only ? 0.0 : log(r.err2));
            if (d->wideband) term += r.dmres * r.dmres / r.dmerr2 + (chi2_only ? 0.0 : log(r.dmerr2));
            sum += term;
        }
        double res = chi2_only ? sum : -sum / 2;
        if (bad || isnan(res)) return chi2_only ? NAN : -INFINITY;
        return res;
    }
    int bad = resids_and_ninvdiag(&ctx, k->y, k->w);
    noise_weights_inv(&ctx, k->phiinv);
    double res = gls_lnlike(d, k->P, k->w, k->phiinv, k->y, k->NinvM, k->Sig, k->u);
    if (bad || isnan(res)) return -INFINITY;
    return res;
}

/* ------------------------------------------------------------------------------------------
 * Exported entry points (ctypes)
 * ---------------------------------------------------------------------------------------- */
/* calc_lnlike_vectorized gls_likelihood.jl:298-309: @threads :static over rows, serial kernel per row */
ORACLE_API int oracle_lnlike_batch(const VelaModelDesc* d, const double* params, int64_t B, int64_t row_stride,
                                   int64_t col_stride, double* out, int nthreads, int chi2_only) {
    int fail = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        Work k;
        int ok = work_alloc(d, &k);
        if (!ok) {
#pragma omp atomic write
            fail = 1;
        }
#pragma omp for schedule(static)
        for (int64_t b = 0; b < B; ++b) {
            if (!ok) continue;
            read_params(d, params + b * row_stride, col_stride, k.P);
            out[b] = lnlike_one(d, &k, chi2_only);
        }
        work_free(&k);
    }
    return fail ? VELA_ERR_OOM : VELA_OK;
}

/* _calc_resids_and_Ninvdiag for one parameter vector; y, w: n_rows */
ORACLE_API int oracle_residuals(const VelaModelDesc* d, const double* params, double* y, double* w) {
    Work k;
    if (!work_alloc(d, &k)) return VELA_ERR_OOM;
    read_params(d, params, 1, k.P);
    Ctx ctx = { d, k.P };
    int bad = resids_and_ninvdiag(&ctx, k.y, k.w);
    int64_t n = d->wideband ? 2 * d->n_toas : d->n_toas;
    if (y) memcpy(y, k.y, sizeof(double) * n);
    if (w) memcpy(w, k.w, sizeof(double) * n);
    work_free(&k);
    return bad ? 1 : 0;
}

/* Double64 phase (phase - pulse_number - tzrphase), doppler-shifted spin frequency and total delay per TOA.
 * Used by the synthetic-TOA generator in tests (phase connection) and for intermediate parity checks. */
ORACLE_API int oracle_phases(const VelaModelDesc* d, const double* params, double* ph_hi, double* ph_lo, double* f,
                             double* delay, double* tzr_hi_lo) {
    Work k;
    if (!work_alloc(d, &k)) return VELA_ERR_OOM;
    read_params(d, params, 1, k.P);
    Ctx ctx = { d, k.P };
    int bad = 0;
    dd tzr = calc_tzr_phase(&ctx, &bad);
    if (tzr_hi_lo) { tzr_hi_lo[0] = tzr.hi; tzr_hi_lo[1] = tzr.lo; }
    for (int64_t j = 0; j < d->n_toas; ++j) {
        Toa toa = toa_view(toa_rec(d, j), d->wideband);
        Corr c = correct_toa(&ctx, &toa);
        dd dphase = dd_sub(dd_sub(c.phase, toa.pulse_number), tzr);
        ph_hi[j] = dphase.hi; ph_lo[j] = dphase.lo;
        if (f) f[j] = c.spin_frequency * (1 + c.doppler);
        if (delay) delay[j] = c.delay;
        bad |= c.bad;
    }
    work_free(&k);
    return bad ? 1 : 0;
}

ORACLE_API int oracle_noise_weights_inv(const VelaModelDesc* d, const double* params, double* phiinv) {
    Work k;
    if (!work_alloc(d, &k)) return VELA_ERR_OOM;
    read_params(d, params, 1, k.P);
    Ctx ctx = { d, k.P };
    noise_weights_inv(&ctx, phiinv);
    work_free(&k);
    return 0;
}

/* Sigma^-1 (column-major p x p, upper triangle q<=p valid) and u = M^T N^-1 y for one parameter vector */
ORACLE_API int oracle_sigma(const VelaModelDesc* d, const double* params, double* Sig, double* u) {
    Work k;
    if (!work_alloc(d, &k)) return VELA_ERR_OOM;
    read_params(d, params, 1, k.P);
    Ctx ctx = { d, k.P };
    int bad = resids_and_ninvdiag(&ctx, k.y, k.w);
    noise_weights_inv(&ctx, k.phiinv);
    sigma_and_u(d, k.P, k.w, k.phiinv, k.y, k.NinvM, k.Sig, k.u);
    int64_t p = d->basis_cols;
    memcpy(Sig, k.Sig, sizeof(double) * p * p);
    memcpy(u, k.u, sizeof(double) * p);
    work_free(&k);
    return bad ? 1 : 0;
}

/* _gls_lnlike_serial on caller-supplied (M, Ninvdiag, Phiinv, y), optionally with an EcorrKernel given as
 * groups[ngroups][3] and the ECORR values: test_woodbury_kernel.jl:37-92 (white) and :96-160 (ECORR) */
ORACLE_API double oracle_gls_lnlike_raw(const double* M, int64_t n, int64_t p, const double* w, const double* phiinv,
                                        const double* y, const uint64_t* groups, int ngroups, const double* ecorr) {
    VelaModelDesc d;
    memset(&d, 0, sizeof d);
    d.basis_rows = n; d.basis_cols = p; d.basis_ld = n; d.noise_basis = M;
    d.kernel_kind = ngroups > 0 ? VELA_KERNEL_WOODBURY_ECORR : VELA_KERNEL_WOODBURY_WHITE;
    d.n_ecorr_groups = ngroups; d.ecorr_groups = groups; d.par_ecorr = 0;
    double* NinvM = malloc(sizeof(double) * (n * p + 1));
    double* Sig = malloc(sizeof(double) * (p * p + 1));
    double* u = malloc(sizeof(double) * (p + 1));
    double r = gls_lnlike(&d, ecorr, w, phiinv, y, NinvM, Sig, u);
    free(NinvM); free(Sig); free(u);
    return r;
}

ORACLE_API void oracle_sincos_cr(double x, double* s, double* c) { sincos_cr(x, s, c); }
ORACLE_API double oracle_mikkola(double l, double e) { int bad = 0; return mikkola(l, e, &bad); }
ORACLE_API double oracle_powerlaw(double A, double gamma, double f, double f1) { return powerlaw(A, gamma, f, f1); }
ORACLE_API int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
