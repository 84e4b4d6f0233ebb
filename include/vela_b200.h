/*
 * vela_b200.h — C ABI of libvela_b200.so: the B200 (sm_100a) implementation of
 * Vela.jl's batched posterior hot path.
 *
 * The reference (Vela.jl v0.1.6) has no FFI layer for this path; the seam it
 * replaces is Julia multiple dispatch on (TimingModel, Vector{TOA}|Vector{WidebandTOA},
 * params).  Every entry point below names the reference function(s) whose
 * result it reproduces (paths relative to the reference tree):
 *
 *   vela_b200_lnlike_batch   <- calc_lnlike_vectorized   src/likelihood/gls_likelihood.jl:298-309
 *                               calc_lnlike_serial       src/likelihood/wls_likelihood.jl:49-56,
 *                                                        src/likelihood/gls_likelihood.jl:249-262
 *                               calc_lnlike              src/likelihood/wls_likelihood.jl:24-47,
 *                                                        src/likelihood/gls_likelihood.jl:264-296
 *   vela_b200_lnpost_batch   <- calc_lnpost_vectorized   src/likelihood/posterior.jl:14-25
 *                               calc_lnpost(_serial)     src/likelihood/posterior.jl:4-12
 *   vela_b200_chi2_batch     <- calc_chi2_serial         src/likelihood/wls_chi2.jl:44-51
 *   vela_b200_residuals      <- form_residuals           src/residuals/residuals.jl:19-26,
 *                                                        src/residuals/wideband_residuals.jl:24-27
 *                               _calc_resids_and_Ninvdiag src/likelihood/gls_likelihood.jl:9-79
 *   vela_b200_noise_weights_inv <- calc_noise_weights_inv src/likelihood/gls_likelihood.jl:3-7
 *   vela_b200_create/destroy <- Pulsar(model, toas)      src/pulsar/pulsar.jl:9-12
 *                               (deep copy of the constant model/TOA data to the GPU(s))
 *
 * Conventions
 *  - plain C types only; all pointers are HOST pointers owned by the caller and
 *    only need to stay valid for the duration of the call (create deep-copies).
 *  - all quantities are in Vela's internal units (seconds^n; docs/src/quantities.md).
 *  - return value 0 = success, negative = VelaStatus error; message through
 *    vela_b200_last_error() (thread-local).  No exceptions, no aborts, no signal
 *    handlers.  Per-sample numerical failure (non positive-definite Sigma^-1,
 *    src/likelihood/gls_likelihood.jl:211-213; invalid eccentricity,
 *    src/model/binary/orbit.jl:26; NaN anywhere) yields -Inf for that sample
 *    and the call still returns 0.
 *  - there is NO CPU fallback: every compute entry point fails with
 *    VELA_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef VELA_B200_H
#define VELA_B200_H

#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VELA_API __attribute__((visibility("default")))
#else
#define VELA_API
#endif

#define VELA_B200_ABI_VERSION 3
#define VELA_MAX_PAR 24

typedef enum VelaStatus {
    VELA_OK = 0,
    VELA_ERR_INVALID_ARG = -1,   /* bad descriptor / shape mismatch (cf. parameter.jl:162) */
    VELA_ERR_NO_DEVICE = -2,     /* no usable CUDA device: the library has no CPU path */
    VELA_ERR_CUDA = -3,          /* CUDA runtime error; see vela_b200_last_error() */
    VELA_ERR_UNSUPPORTED = -4,   /* component / kernel outside the implemented hot path */
    VELA_ERR_OOM = -5
} VelaStatus;

/* Timing-model components, in the order-dependent chain of
 * correct_toa(components::Tuple, ...)  src/model/timing_model.jl:85-98. */
typedef enum VelaComponentKind {
    VELA_COMP_SOLAR_SYSTEM = 1,        /* src/model/solarsystem.jl:67-134 */
    VELA_COMP_SOLAR_WIND = 2,          /* SolarWindDispersion, src/model/solarwind.jl:37-51 */
    VELA_COMP_DISPERSION_TAYLOR = 3,   /* src/model/dispersion.jl:15-26 */
    VELA_COMP_DISPERSION_PIECEWISE = 4,/* DMX, src/model/dispersion.jl:43-54 */
    VELA_COMP_CHROMATIC_TAYLOR = 5,    /* src/model/chromatic.jl:23-34 */
    VELA_COMP_CHROMATIC_PIECEWISE = 6, /* CMX, src/model/chromatic.jl:46-57 */
    VELA_COMP_BINARY_ELL1 = 7,         /* src/model/binary/binary_ell1_base.jl:159-180 */
    VELA_COMP_BINARY_DD = 8,           /* src/model/binary/binary_dd_base.jl:134-162 */
    VELA_COMP_GLITCH = 9,              /* src/model/glitch.jl:14-38 */
    VELA_COMP_EXP_DIP = 10,            /* ChromaticExponentialDip, src/model/expdip.jl:11-31 */
    VELA_COMP_SPINDOWN = 11,           /* src/model/spindown.jl:15-33 */
    VELA_COMP_PHASE_OFFSET = 12,       /* src/model/phase_offset.jl:12-13 */
    VELA_COMP_PHASE_JUMP_EXCLUSIVE = 13, /* ExclusivePhaseJump, src/model/jump.jl:31-45 */
    VELA_COMP_PHASE_JUMP = 14,         /* PhaseJump (BitMatrix), src/model/jump.jl:14-22 */
    VELA_COMP_MEASUREMENT_NOISE = 15,  /* EFAC/EQUAD, src/model/measurement_noise.jl:20-38 */
    VELA_COMP_DM_JUMP_EXCLUSIVE = 16,  /* ExclusiveDispersionJump, src/model/dmjump.jl:62-86 */
    VELA_COMP_DM_JUMP = 17,            /* DispersionJump (BitMatrix), src/model/dmjump.jl:28-47 */
    VELA_COMP_DM_MEASUREMENT_NOISE = 18, /* DMEFAC/DMEQUAD, src/model/dispersion_measurement_noise.jl:19-46 */
    VELA_COMP_BINARY_ELL1H = 19,       /* src/model/binary/binary_ell1h.jl:18-46 */
    VELA_COMP_BINARY_ELL1K = 20,       /* src/model/binary/binary_ell1k.jl:13-49 */
    VELA_COMP_BINARY_DDH = 21,         /* src/model/binary/binary_ddh.jl:14-24 */
    VELA_COMP_BINARY_DDS = 22,         /* src/model/binary/binary_dds.jl:14-23 */
    VELA_COMP_BINARY_DDK = 23,         /* src/model/binary/binary_ddk.jl:15-120 (Kopeikin corrections) */
    VELA_COMP_FD = 24,                 /* FrequencyDependent, src/model/frequency_dependent.jl:22-34 */
    VELA_COMP_FD_JUMP = 25,            /* FrequencyDependentJump, src/model/frequency_dependent.jl:46-81 */
    VELA_COMP_WAVEX = 26,              /* src/model/wavex.jl:23-62 */
    VELA_COMP_DM_WAVEX = 27,           /* src/model/wavex.jl:78-87 */
    VELA_COMP_CM_WAVEX = 28,           /* src/model/wavex.jl:97-106 */
    VELA_COMP_SOLAR_WIND_PIECEWISE = 29, /* SolarWindDispersionPiecewise (SWX), src/model/solarwind.jl:64-91 */
    VELA_COMP_DM_OFFSET_EXCLUSIVE = 30,  /* ExclusiveDispersionOffset (FDJUMPDM), src/model/fdjumpdm.jl:44-66 */
    VELA_COMP_DM_OFFSET = 31,          /* DispersionOffset (FDJUMPDM, BitMatrix), src/model/fdjumpdm.jl:14-37 */
    /* Power-law Fourier GPs with FREE (un-marginalised) amplitudes, used as ordinary delay components when
     * pyvela is run with marginalize_gp_noise=False (pyvela/pyvela/model.py:693-700): */
    VELA_COMP_POWERLAW_RED_GP = 32,    /* PowerlawRedNoiseGP delay, src/model/gp_noise.jl:110-130 */
    VELA_COMP_POWERLAW_DM_GP = 33,     /* PowerlawDispersionNoiseGP dispersion_slope, src/model/gp_noise.jl:155-183 */
    VELA_COMP_POWERLAW_CHROM_GP = 34   /* PowerlawChromaticNoiseGP chromatic_slope, src/model/gp_noise.jl:207-235 */
} VelaComponentKind;
#define VELA_COMP_KIND_MAX 34

/* VelaComponentDesc.flags */
#define VELA_FLAG_ECLIPTIC        1   /* SolarSystem.ecliptic_coordinates */
#define VELA_FLAG_PLANET_SHAPIRO  2   /* SolarSystem.planet_shapiro */
#define VELA_FLAG_USE_FBX         4   /* BinaryELL1/BinaryDD.use_fbx */

/*
 * One timing-model component.  `par[]` holds 0-based offsets into the FULL
 * parameter vector (the flattened `_default_params_tuple`,
 * src/parameter/parameter.jl:87-92,151-153: non-noise singles, non-noise multis,
 * noise singles, noise multis); -1 = unused.  For tuple-valued parameters the
 * offset is that of the first element, the length is in `len[]`.
 *
 * Slot meaning per kind (par / len / mask):
 *  SOLAR_SYSTEM   par: 0 LONG(RAJ|ELONG) 1 LAT(DECJ|ELAT) 2 PMLONG 3 PMLAT 4 PX 5 POSEPOCH
 *  SOLAR_WIND     par: 0 SWEPOCH 1 NE_SW[0]            len: 0 #NE_SW
 *  DISPERSION_TAYLOR par: 0 DMEPOCH 1 DM[0]            len: 0 #DM
 *  DISPERSION_PIECEWISE par: 0 DMX_[0]                 len: 0 #DMX_   mask: 0 dmx index mask
 *  CHROMATIC_TAYLOR par: 0 CMEPOCH 1 CM[0] 2 TNCHROMIDX len: 0 #CM
 *  CHROMATIC_PIECEWISE par: 0 CMX_[0] 2 TNCHROMIDX      len: 0 #CMX_   mask: 0 cmx index mask
 *  BINARY_ELL1    par: 0 TASC 1 PB 2 PBDOT 3 FB[0] 4 A1 5 A1DOT 6 EPS1 7 EPS2 8 EPS1DOT
 *                      9 EPS2DOT 10 M2 11 SINI          len: 0 #FB
 *  BINARY_DD      par: 0 T0 1 PB 2 PBDOT 3 FB[0] 4 A1 5 A1DOT 6 ECC 7 EDOT 8 OM 9 OMDOT
 *                      10 DR 11 DTH 12 GAMMA 13 M2 14 SINI  len: 0 #FB
 *  GLITCH         par: 0 GLEP_[0] 1 GLPH_[0] 2 GLF0_[0] 3 GLF1_[0] 4 GLF2_[0] 5 GLF0D_[0]
 *                      6 GLTD_[0]                       len: 0 #glitches
 *  EXP_DIP        par: 0 EXPDIPFREF 1 EXPDIPEPS 2 EXPDIPEP_[0] 3 EXPDIPAMP_[0] 4 EXPDIPIDX_[0]
 *                      5 EXPDIPTAU_[0]                  len: 0 #dips
 *  SPINDOWN       par: 0 F_ 1 F[0]                      len: 0 #F
 *  PHASE_OFFSET   par: 0 PHOFF
 *  PHASE_JUMP_EXCLUSIVE par: 0 JUMP[0] 1 F_ 2 F[0]      len: 0 #JUMP  mask: 0 jump index mask
 *  PHASE_JUMP     par: 0 JUMP[0] 1 F_ 2 F[0]            len: 0 #JUMP  mask: 0 first bit-mask row
 *  MEASUREMENT_NOISE par: 0 EFAC[0] 1 EQUAD[0]          len: 0 #EFAC 1 #EQUAD
 *                                                       mask: 0 efac index mask 1 equad index mask
 *  DM_JUMP_EXCLUSIVE par: 0 DMJUMP[0]                   len: 0 #DMJUMP mask: 0 index mask
 *  DM_JUMP        par: 0 DMJUMP[0]                      len: 0 #DMJUMP mask: 0 first bit-mask row
 *  DM_MEASUREMENT_NOISE par: 0 DMEFAC[0] 1 DMEQUAD[0]   len: 0 #DMEFAC 1 #DMEQUAD
 *                                                       mask: 0 dmefac mask 1 dmequad mask
 *  BINARY_ELL1H   as BINARY_ELL1 with par: 10 H3 11 STIGMA
 *  BINARY_ELL1K   as BINARY_ELL1 with par: 8 OMDOT 9 LNEDOT (instead of EPS1DOT/EPS2DOT)
 *  BINARY_DDH     as BINARY_DD with par: 13 H3 14 STIGMA
 *  BINARY_DDS     as BINARY_DD with par: 13 M2 14 SHAPMAX
 *  BINARY_DDK     as BINARY_DD with par: 13 M2 14 KIN 15 KOM 16 PMLONG(PMRA|PMELONG) 17 PMLAT(PMDEC|PMELAT) 18 PX;
 *                 flags also carry VELA_FLAG_ECLIPTIC
 *  FD             par: 0 FD[0]                          len: 0 #FD
 *  FD_JUMP        par: 0 FDJUMP[0]                      len: 0 #FDJUMP mask: 0 first bit-mask row,
 *                                                       1 index-mask row whose first #FDJUMP entries are the exponents
 *  WAVEX          par: 0 WXEPOCH 1 WXSIN_[0] 2 WXCOS_[0] 3 WXFREQ_[0]          len: 0 #harmonics
 *  DM_WAVEX       par: 0 DMWXEPOCH 1 DMWXSIN_[0] 2 DMWXCOS_[0] 3 DMWXFREQ_[0]  len: 0 #harmonics
 *  CM_WAVEX       par: 0 CMWXEPOCH 1 CMWXSIN_[0] 2 CMWXCOS_[0] 3 CMWXFREQ_[0] 4 TNCHROMIDX  len: 0 #harmonics
 *  SOLAR_WIND_PIECEWISE par: 0 SWXDM_[0]                len: 0 #SWXDM_  mask: 0 swx index mask
 *                                                       dval: 0 conjunction_geometry 1 opposition_geometry
 *  DM_OFFSET_EXCLUSIVE par: 0 FDJUMPDM[0]               len: 0 #FDJUMPDM mask: 0 index mask
 *  DM_OFFSET      par: 0 FDJUMPDM[0]                    len: 0 #FDJUMPDM mask: 0 first bit-mask row
 *  POWERLAW_RED_GP   par: 0 TNREDAMP 1 TNREDGAM 2 PLREDSIN_[0] 3 PLREDCOS_[0] 4 PLREDFREQ 5 PLREDEPOCH
 *  POWERLAW_DM_GP    par: 0 TNDMAMP 1 TNDMGAM 2 PLDMSIN_[0] 3 PLDMCOS_[0] 4 PLDMFREQ 5 PLDMEPOCH
 *  POWERLAW_CHROM_GP par: 0 TNCHROMAMP 1 TNCHROMGAM 2 PLCHROMSIN_[0] 3 PLCHROMCOS_[0] 4 PLCHROMFREQ 5 PLCHROMEPOCH
 *                         6 TNCHROMIDX
 *                    all three: len: 0 #harmonics N, 1 number of leading log-spaced harmonics
 *                    (findfirst(iszero, ln_js) - 1, gp_noise.jl:37), 2 offset into VelaModelDesc.aux_values of
 *                    2N doubles: ln_js[0..N) (the struct field, widened from its storage type) followed by
 *                    js[0..N) = exp(ln_js[i]) evaluated IN THE STORAGE TYPE (Float32 for the red-noise GP).
 */
typedef struct VelaComponentDesc {
    int32_t kind;
    int32_t flags;
    int32_t par[VELA_MAX_PAR];
    int32_t len[4];
    int32_t mask[2];
    double dval[2];              /* per-component constants that are not parameters (struct fields in the reference) */
} VelaComponentDesc;

/* Likelihood kernels, src/model/kernel.jl:20-92. */
typedef enum VelaKernelKind {
    VELA_KERNEL_WHITE = 0,            /* WhiteNoiseKernel */
    VELA_KERNEL_ECORR = 1,            /* EcorrKernel: white noise + ECORR, block-diagonal covariance
                                         (src/likelihood/ecorr_likelihood.jl, ecorr_chi2.jl); narrowband only */
    VELA_KERNEL_WOODBURY_WHITE = 2,   /* WoodburyKernel{WhiteNoiseKernel} */
    VELA_KERNEL_WOODBURY_ECORR = 3    /* WoodburyKernel{EcorrKernel} (src/likelihood/gls_ecorr_likelihood.jl) */
} VelaKernelKind;

/* Amplitude-marginalised Gaussian-process components of a WoodburyKernel. */
typedef enum VelaGPKind {
    VELA_GP_POWERLAW_RED = 1,    /* PowerlawRedNoiseGP, ln_js are Float32: src/model/gp_noise.jl:110-115 */
    VELA_GP_POWERLAW_DM = 2,     /* PowerlawDispersionNoiseGP, src/model/gp_noise.jl:155-160 */
    VELA_GP_POWERLAW_CHROM = 3,  /* PowerlawChromaticNoiseGP, src/model/gp_noise.jl:207-212 */
    VELA_GP_MARGINALIZED_TM = 4  /* MarginalizedTimingModel, src/model/marginalized_timing_model.jl:13-26 */
} VelaGPKind;

typedef struct VelaGPDesc {
    int32_t kind;
    int32_t n;               /* power-law GPs: number of harmonics (2n basis columns, [SIN x n, COS x n]);
                                MARGINALIZED_TM: number of columns */
    int32_t par_log10_amp;   /* TNREDAMP | TNDMAMP | TNCHROMAMP */
    int32_t par_gamma;       /* TNREDGAM | TNDMGAM | TNCHROMGAM */
    int32_t par_f1;          /* PLREDFREQ | PLDMFREQ | PLCHROMFREQ */
    int32_t _pad;
    const double* ln_js;        /* n values; RED: the Float32 tuple widened to double (exact) */
    const double* weights_inv;  /* MARGINALIZED_TM: prior_weights_inv[n] */
} VelaGPDesc;

/*
 * TOA records are passed in the memory layout of the reference's isbits structs
 * so that a Julia caller hands over `pointer(toas)` unchanged:
 *   TOA (248 bytes, src/toa/toa.jl:39-50 + src/toa/ephemeris.jl:14-23), 31 little-endian 8-byte words:
 *     0 value.hi 1 value.lo 2 error 3 observing_frequency 4 pulse_number.hi 5 pulse_number.lo
 *     6-8 ssb_obs_pos 9-11 ssb_obs_vel 12-14 obs_sun_pos 15-17 obs_jupiter_pos 18-20 obs_saturn_pos
 *     21-23 obs_venus_pos 24-26 obs_uranus_pos 27-29 obs_neptune_pos 30 index (UInt64; 0 = TZR TOA)
 *   WidebandTOA (264 bytes, src/toa/wideband_toa.jl:18-21,85-88): the TOA record followed by
 *     31 dminfo.value 32 dminfo.error
 * The TZR TOA is always a narrowband TOA record (src/model/timing_model.jl:24).
 */
#define VELA_TOA_WORDS 31
#define VELA_WTOA_WORDS 33

typedef struct VelaModelDesc {
    int32_t abi_version;          /* VELA_B200_ABI_VERSION */
    int32_t n_components;
    const VelaComponentDesc* components;

    int32_t n_params;             /* length of the full parameter vector */
    int32_t n_free;               /* number of free parameters (ndim of the sampler) */
    const double* default_values; /* ParamHandler._default_values, parameter.jl:130 */
    const int32_t* free_indices;  /* ParamHandler._free_indices converted to 0-based */

    int64_t n_toas;
    int32_t wideband;             /* 0: Vector{TOA}; 1: Vector{WidebandTOA} */
    int32_t toa_stride_bytes;     /* 248 or 264 */
    const void* toas;             /* n_toas records */
    const void* tzr_toa;          /* one 248-byte TOA record, index == 0 */

    int32_t n_index_masks;        /* exclusive ("index") masks, each n_toas long */
    int32_t n_bit_mask_rows;      /* rows of all BitMatrix masks, each n_toas long */
    const uint32_t* index_masks;  /* [n_index_masks][n_toas]; entry = 1-based group, 0 = none */
    const uint8_t* bit_masks;     /* [n_bit_mask_rows][n_toas]; entry = 0/1 */

    int32_t kernel_kind;          /* VelaKernelKind */
    int32_t n_gp;
    const VelaGPDesc* gp;         /* WoodburyKernel.gp_components, in order */
    int64_t basis_rows;           /* n_toas (narrowband) or 2*n_toas (wideband) */
    int64_t basis_cols;           /* p = sum of GP columns */
    int64_t basis_ld;             /* leading dimension (column-major, Julia Matrix{Float64}) */
    const double* noise_basis;    /* WoodburyKernel.noise_basis, kernel.jl:77 */

    /* ECORR groups of an EcorrKernel (kernel kinds 1 and 3): the reference's isbits EcorrGroup records
     * {start, stop, index} as 3 x UInt64 each (src/model/kernel.jl:25-29), 1-based inclusive TOA ranges that
     * partition the (ECORR-sorted, pyvela/pyvela/ecorr.py:31-75) TOAs; index 0 = no ECORR, else 1-based into ECORR. */
    int32_t n_ecorr_groups;
    int32_t par_ecorr;            /* offset of ECORR[1] in the full parameter vector (-1 if none) */
    const uint64_t* ecorr_groups; /* [n_ecorr_groups][3] */

    int32_t n_devices;            /* 0: use the current CUDA device */
    int32_t _pad;
    const int32_t* devices;       /* CUDA device ordinals; walkers are split evenly across them */

    /* Per-component constant tables that are struct fields in the reference (ln_js of the un-marginalised
     * power-law GPs); components index into it through len[2].  May be NULL when n_aux == 0. */
    int64_t n_aux;
    const double* aux_values;
} VelaModelDesc;

typedef struct VelaPulsar* VelaHandle;

#ifndef __CUDACC_RTC__   /* the entry points are host functions: hidden from the device-only NVRTC pass */

/* Number of usable CUDA devices (0 if none / no driver). */
VELA_API int vela_b200_device_count(void);

/* Thread-local description of the last error returned on this thread. */
VELA_API const char* vela_b200_last_error(void);

/* Deep-copies the model and TOA data onto the device(s). */
VELA_API int vela_b200_create(const VelaModelDesc* desc, VelaHandle* out);
VELA_API int vela_b200_destroy(VelaHandle h);

/* Model shape queries. */
VELA_API int vela_b200_n_free(VelaHandle h);
VELA_API int64_t vela_b200_n_rows(VelaHandle h);   /* rows of y: n_toas or 2*n_toas */
VELA_API int64_t vela_b200_n_basis(VelaHandle h);  /* p */

/*
 * ln L for a matrix of parameter samples.  Element (b, k) of the (B x n_free)
 * matrix is params[b*row_stride + k*col_stride] (strides in elements): a
 * row-major numpy array uses (n_free, 1); a Julia Matrix uses (1, B).
 * out_lnlike[B].
 */
VELA_API int vela_b200_lnlike_batch(VelaHandle h, const double* params, int64_t B, int64_t n_free,
                           int64_t row_stride, int64_t col_stride, double* out_lnlike);

/*
 * ln posterior = lnprior[b] + ln L_b, with the reference's short-circuit: when
 * lnprior[b] is not finite the result is lnprior[b] itself (posterior.jl:9-12).
 * lnprior == NULL means a flat (zero) prior (or the device-side priors, once vela_b200_set_priors has been called).
 * A sample whose likelihood evaluates to NaN (an assertion of the reference would have thrown: orbit.jl:26,
 * toa.jl:95-99) is reported as -Inf, never as NaN and never as an error of the call - a deliberate difference from the
 * reference, whose exception would abort the whole batch.
 */
VELA_API int vela_b200_lnpost_batch(VelaHandle h, const double* params, int64_t B, int64_t n_free,
                           int64_t row_stride, int64_t col_stride, const double* lnprior,
                           double* out_lnpost);

/* chi^2 = sum r^2/sigma^2 (+ DM term for wideband) for white-noise kernels (wls_chi2.jl:44-51), the
 * ECORR-corrected chi^2 for EcorrKernel (ecorr_chi2.jl:6-72); not defined for Woodbury kernels. */
VELA_API int vela_b200_chi2_batch(VelaHandle h, const double* params, int64_t B, int64_t n_free,
                         int64_t row_stride, int64_t col_stride, double* out_chi2);

/*
 * Residuals y[n_rows] (seconds; wideband: [time residuals; DM residuals]) and
 * N^-1 diagonal w[n_rows] = 1/sigma^2 for ONE parameter vector.
 * Either output may be NULL.
 */
VELA_API int vela_b200_residuals(VelaHandle h, const double* params, int64_t n_free, double* y, double* w);

/*
 * Device-side priors (optional).  The reference evaluates one univariate prior per free parameter on the host
 * (lnprior, src/prior/prior.jl:13-19; SimplePrior / SimplePriorMulti, src/prior/simple_prior.jl:17-55; custom
 * distributions, src/prior/known_priors.jl:33-100; defaults pyvela/pyvela/priors.py:30-76).  After
 * vela_b200_set_priors the library evaluates them itself: vela_b200_lnpost_batch with lnprior == NULL then returns
 * lnprior + lnlike with the reference's -Inf short-circuit (posterior.jl:9-12) instead of assuming a flat prior.
 * One descriptor per free parameter, in free-parameter order.
 */
typedef enum VelaPriorKind {
    VELA_PRIOR_UNIFORM = 1,        /* Uniform(a, b) */
    VELA_PRIOR_NORMAL = 2,         /* Normal(mu = a, sigma = b) */
    VELA_PRIOR_LOGNORMAL = 3,      /* LogNormal(mu = a, sigma = b) */
    VELA_PRIOR_LOGUNIFORM = 4,     /* LogUniform(a, b) */
    VELA_PRIOR_KIN = 5,            /* KINPriorDistribution, known_priors.jl:33-37 */
    VELA_PRIOR_SINI = 6,           /* SINIPriorDistribution, known_priors.jl:45-49 */
    VELA_PRIOR_STIGMA = 7,         /* STIGMAPriorDistribution, known_priors.jl:57-61 */
    VELA_PRIOR_SHAPMAX = 8,        /* SHAPMAXPriorDistribution, known_priors.jl:69-73 */
    VELA_PRIOR_PX = 9,             /* PXPriorDistribution(Rmax = a), known_priors.jl:82-88 */
    VELA_PRIOR_LATITUDE = 10,      /* LatitudePriorDistribution, known_priors.jl:96-100 */
    VELA_PRIOR_TRUNCATED_NORMAL = 11 /* truncated(Normal(a, b), lo, hi) */
} VelaPriorKind;

typedef struct VelaPriorDesc {
    int32_t kind;
    int32_t _pad;
    double a, b;                  /* distribution arguments, see VelaPriorKind */
    double lo, hi;                /* truncation bounds (TRUNCATED_NORMAL only) */
} VelaPriorDesc;

VELA_API int vela_b200_set_priors(VelaHandle h, const VelaPriorDesc* priors, int32_t n_free);
/* sum of logpdf over the free parameters for each row of params (calc_lnprior, prior.jl:18-19) */
VELA_API int vela_b200_lnprior_batch(VelaHandle h, const double* params, int64_t B, int64_t n_free,
                                     int64_t row_stride, int64_t col_stride, double* out_lnprior);
/* quantile of each prior at each unit-cube coordinate (prior_transform, prior.jl:31-32); cube and out are
 * row-major (B x n_free) */
VELA_API int vela_b200_prior_transform_batch(VelaHandle h, const double* cube, int64_t B, int64_t n_free, double* out);

/*
 * Phase residuals before rounding to Float64, for ONE parameter vector: the Double64
 * `phase_residual(toa, ctoa) - tzrphase` (src/toa/toa.jl:151, src/residuals/residuals.jl:10) as (hi, lo)
 * and `doppler_shifted_spin_frequency(ctoa)` (src/toa/toa.jl:131-134), n_toas values each.
 * This is what pulse-number assignment needs (pyvela/pyvela/model.py:699-700 asks PINT for it).
 */
VELA_API int vela_b200_phases(VelaHandle h, const double* params, int64_t n_free, double* phase_hi,
                              double* phase_lo, double* spin_frequency);

/* Phi^-1 (p values) for ONE parameter vector; Woodbury kernels only. */
VELA_API int vela_b200_noise_weights_inv(VelaHandle h, const double* params, int64_t n_free, double* phiinv);

/*
 * Sigma^-1 = diag(Phi^-1) + M^T N^-1 M (full symmetric, column-major p x p) and u = M^T N^-1 y
 * for ONE parameter vector: _calc_Sigmainv__and__MT_Ninv_y, src/likelihood/gls_likelihood.jl:101-169
 * (pyvela reads the same quantities in spnta.py:376-394).  Woodbury kernels only.
 */
VELA_API int vela_b200_sigma_and_u(VelaHandle h, const double* params, int64_t n_free, double* Sigma, double* u);

/*
 * Conditional Gaussian of the analytically marginalised amplitudes for a batch of samples, computed from
 * the same Sigma^-1 / u / Cholesky factor as ln L (no second factorisation):
 *   ahat[b, :]   = Sigma u                        (B x p, row-major)   conditional mean offsets
 *   cf[b, :, :]  = upper factor R, Sigma^-1 = R^T R (B x p x p, row-major; NULL to skip)
 *   lnlike[b]    = ln L of the same sample          (NULL to skip)
 * Replaces SPNTA.get_marginalized_param_offset_mean_and_covinvcf, pyvela/pyvela/spnta.py:376-394
 * (scipy cholesky(lower=False) + cho_solve on the host); the draws in spnta.py:419-437 need only
 * R and ahat.  Rows whose Sigma^-1 is not positive definite are NaN.  With an ECORR inner kernel
 * Sigma^-1 carries the exact N^-1 (the reference helper keeps only its diagonal there).
 * Woodbury kernels only; runs on devices[0].
 */
/*
 * Per-model specialisation of the chain kernel.  `create` compiles the component chain of THIS model with NVRTC
 * (program baked in as constants; same source and flags as the generic kernel, bit-identical results) unless
 * VELA_B200_JIT=0 or libnvrtc cannot be loaded, in which case the generic interpreting kernel is used.
 *   vela_b200_chain_specialized(h, -1) -> 1 / 0: which kernel is in use;  (h, 0 | 1) switches (A/B tests).
 *   vela_b200_jit_log(h): compiler log, or why there is no specialised kernel.
 *   vela_b200_debug_spec_source / _compile: the generated translation unit and a device-free compile of it
 *   (returns the cubin size), for tests and tooling.
 */
VELA_API int vela_b200_chain_specialized(VelaHandle h, int on);
/* Structured Sigma path (Fourier-harmonic bases): host-only evaluation for tests, see vela_api.cu. */
VELA_API int64_t vela_b200_debug_structured_sigma(const VelaModelDesc* desc, const double* w, double* sigma_packed, int32_t* info);
/* 1 if this handle assembles Sigma from column sums (structured path), 0 if it runs the dense pair contraction. */
VELA_API int vela_b200_sigma_structured(VelaHandle h);
/* What the Sigma stage executes for a batch of B walkers: returns the FP64 operations per walker of the column-sum launches
 * (0 on the dense path); info[8] = {structured, two-stage taken, table columns, padded rows, stage-1 columns, stage-2
 * columns as launched, moments per block and walker, cells}.  Measurement bookkeeping only. */
VELA_API double vela_b200_sigma_info(VelaHandle h, int64_t B, int64_t* info);
/* Host-only check of the two-stage column sums (cell moments + stage-2 matrices) against the direct ones, see vela_api.cu. */
VELA_API int64_t vela_b200_debug_two_stage(const VelaModelDesc* desc, const double* w, const double* wy, double* R_direct,
                                           double* R_two, int64_t* info);
VELA_API const char* vela_b200_jit_log(VelaHandle h);
VELA_API int64_t vela_b200_debug_spec_source(const VelaModelDesc* desc, char* buf, int64_t cap);
VELA_API int64_t vela_b200_debug_spec_compile(const VelaModelDesc* desc);

VELA_API int vela_b200_marginalized_mean_batch(VelaHandle h, const double* params, int64_t B, int64_t n_free,
                                               int64_t row_stride, int64_t col_stride, double* ahat, double* cf,
                                               double* lnlike);

/*
 * Device-resident variant used by benchmarks and by callers that already hold
 * the samples on the GPU: `d_params` is a row-major (B x n_free) DEVICE buffer on
 * device `devices[0]`, `d_out` a DEVICE buffer of B doubles.  Work is enqueued on
 * `stream` (a cudaStream_t passed as void*; NULL = the handle's own stream) and
 * the call returns without synchronising.  The handle's scratch is one set per device:
 * every call (this one, the host-buffer batches, residuals, ...) orders itself after the
 * previous call's work with an event, whatever streams the two used, so calls on one
 * handle may be issued back to back on different streams; d_params / d_out / d_lnprior
 * are the caller's and must stay valid until the stream has run the work.
 */
VELA_API int vela_b200_lnlike_batch_device(VelaHandle h, const double* d_params, int64_t B,
                                  double* d_out, void* stream);

/* Same, for the log-posterior (with_prior != 0): d_lnprior is a DEVICE buffer of B doubles.  d_lnprior == NULL: the
 * priors set with vela_b200_set_priors are evaluated on the device; without them NULL means a flat (zero) prior. */
VELA_API int vela_b200_lnpost_batch_device(VelaHandle h, const double* d_params, int64_t B, const double* d_lnprior,
                                           int with_prior, double* d_out, void* stream);

/* Number of kernel launches the handle has issued so far (all devices). */
VELA_API int64_t vela_b200_launch_count(VelaHandle h);

/* Timing hooks: device time (ms) of the stages of the last batch on device 0, measured with CUDA
 * events on the stream the kernels were launched on.  stage: 0 sample prologue, 1 TOA chain,
 * 2 Sigma contraction (white kernels: finish), 3 Cholesky/finish, 4 all kernels.  Returns <0 if stage
 * timing is not enabled. */
VELA_API int vela_b200_set_stage_timing(VelaHandle h, int enabled);
VELA_API double vela_b200_last_stage_ms(VelaHandle h, int stage);

/* FP64 micro-benchmarks used for the roofline denominators (TFLOP/s):
 * kind 0 = register-resident DFMA loop, kind 1 = mma.sync m8n8k4 f64 (DMMA) loop. */
VELA_API double vela_b200_fp64_peak_tflops(int device, int kind);

/* Probe: do scalar FP64 (DFMA) and the FP64 tensor path (DMMA) issue side by side?  tflops[0] DFMA in half the warps
 * alone, [1] DMMA in the other half alone, [2] / [3] the two parts when both halves run in one launch. */
VELA_API int vela_b200_debug_fp64_coissue(int device, double* tflops);

/* Test hook: quotients of the shared-reciprocal division (proper-motion normalisation) that differ from IEEE division
 * over n pseudo-random operand sets; 0 expected. */
VELA_API int64_t vela_b200_debug_div3(int64_t n, uint64_t seed);

/* Test hook: the correctly rounded sin/cos (double-double Taylor) used for the sky-position unit vector. */
VELA_API int vela_b200_debug_sincos(const double* x, int n, double* s, double* c);

#endif /* !__CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* VELA_B200_H */
