# VelaB200.jl — the Julia host side of the drop-in: Vela.jl's own API names, backed by libvela_b200.so.
#
# How a maintainer wires it in (see INTEGRATION.md): `include("gpu/VelaB200.jl")` at the end of
# src/Vela.jl (after src/pulsar/pulsar.jl:25) and `using .VelaB200`.  Nothing else in Vela.jl or pyvela
# changes: `SPNTA.pulsar = VelaB200.B200Pulsar(model, toas)` makes `vl.calc_lnpost_vectorized(self.pulsar, …)`
# (pyvela/pyvela/spnta.py:273-276) dispatch to the GPU.
#
# NOTE: Julia is not installed in the build/CI image of this repository, so this file is NOT executed by
# the test-suite; every field it fills is covered through the identical Python/ctypes path
# (vela.jl_b200/descriptor.py), which is written against the same header, include/vela_b200.h.
# It is kept deliberately thin: it only flattens structs and forwards pointers.

module VelaB200

using ..Vela
using ..Vela: TimingModel, TOA, WidebandTOA, TOABase, Pulsar, ParamHandler, WhiteNoiseKernel, EcorrKernel, EcorrGroup,
    WoodburyKernel,
    SolarSystem, SolarWindDispersion, DispersionTaylor, DispersionPiecewise, ChromaticTaylor, ChromaticPiecewise,
    BinaryELL1, BinaryELL1H, BinaryELL1k, BinaryDD, BinaryDDH, BinaryDDS, BinaryDDK, FrequencyDependent,
    FrequencyDependentJump, WaveX, DMWaveX, CMWaveX, SolarWindDispersionPiecewise, DispersionOffset,
    ExclusiveDispersionOffset, Glitch, ChromaticExponentialDip, Spindown, PhaseOffset, PhaseJump, ExclusivePhaseJump,
    MeasurementNoise, DispersionJump, ExclusiveDispersionJump, DispersionMeasurementNoise, PowerlawRedNoiseGP,
    PowerlawDispersionNoiseGP, PowerlawChromaticNoiseGP, MarginalizedTimingModel, calc_lnprior,
    read_param_values_to_vector

export B200Pulsar

const LIB = get(ENV, "VELA_B200_LIB", "libvela_b200")
const VELA_MAX_PAR = 24

# ---- mirrors of the C structs in include/vela_b200.h (field order and types must match) ----------------
struct VelaComponentDesc
    kind::Int32
    flags::Int32
    par::NTuple{VELA_MAX_PAR,Int32}
    len::NTuple{4,Int32}
    mask::NTuple{2,Int32}
    dval::NTuple{2,Float64}
end

struct VelaGPDesc
    kind::Int32
    n::Int32
    par_log10_amp::Int32
    par_gamma::Int32
    par_f1::Int32
    _pad::Int32
    ln_js::Ptr{Float64}
    weights_inv::Ptr{Float64}
end

struct VelaModelDesc
    abi_version::Int32
    n_components::Int32
    components::Ptr{VelaComponentDesc}
    n_params::Int32
    n_free::Int32
    default_values::Ptr{Float64}
    free_indices::Ptr{Int32}
    n_toas::Int64
    wideband::Int32
    toa_stride_bytes::Int32
    toas::Ptr{Cvoid}
    tzr_toa::Ptr{Cvoid}
    n_index_masks::Int32
    n_bit_mask_rows::Int32
    index_masks::Ptr{UInt32}
    bit_masks::Ptr{UInt8}
    kernel_kind::Int32
    n_gp::Int32
    gp::Ptr{VelaGPDesc}
    basis_rows::Int64
    basis_cols::Int64
    basis_ld::Int64
    noise_basis::Ptr{Float64}
    n_ecorr_groups::Int32
    par_ecorr::Int32
    ecorr_groups::Ptr{UInt64}      # Vector{EcorrGroup} is isbits: 3 x UInt64 per group (kernel.jl:25-29)
    n_devices::Int32
    _pad::Int32
    devices::Ptr{Int32}
    n_aux::Int64
    aux_values::Ptr{Float64}
end

# ---- parameter offsets: the flattened `_default_params_tuple` (src/parameter/parameter.jl:87-92,151-153) ----
function param_offsets(ph::ParamHandler)
    offs = Dict{Symbol,Tuple{Int32,Int32}}()
    idx = 0
    for (name, val) in pairs(ph._default_params_tuple)
        n = val isa Tuple ? length(val) : 1
        offs[name] = (Int32(idx), Int32(n))      # 0-based offset, length
        idx += n
    end
    return offs
end

mutable struct Builder
    offs::Dict{Symbol,Tuple{Int32,Int32}}
    ntoa::Int
    index_masks::Vector{Vector{UInt32}}
    bit_rows::Vector{Vector{UInt8}}
    aux::Vector{Float64}
end

function comp(kind; flags = 0, par = (), len = (), mask = (), dval = (0.0, 0.0))
    p = fill(Int32(-1), VELA_MAX_PAR)
    for (slot, off) in par
        p[slot+1] = off
    end
    l = zeros(Int32, 4)
    for (slot, n) in len
        l[slot+1] = n
    end
    m = fill(Int32(-1), 2)
    for (slot, i) in mask
        m[slot+1] = i
    end
    return VelaComponentDesc(Int32(kind), Int32(flags), Tuple(p), Tuple(l), Tuple(m), Float64.(dval))
end

off(b::Builder, s::Symbol) = b.offs[s][1]
len(b::Builder, s::Symbol) = b.offs[s][2]
has(b::Builder, s::Symbol) = haskey(b.offs, s)
function add_index_mask!(b::Builder, mask::Vector{UInt})
    push!(b.index_masks, UInt32.(mask))
    return Int32(length(b.index_masks) - 1)
end
function add_bit_mask!(b::Builder, mask::BitMatrix)
    first = length(b.bit_rows)
    for r = 1:size(mask, 1)
        push!(b.bit_rows, UInt8.(mask[r, :]))
    end
    return Int32(first)
end

# one method per hot-path component; slot tables are documented in include/vela_b200.h
describe(b::Builder, c::SolarSystem) =
    let n = c.ecliptic_coordinates ? (:ELONG, :ELAT, :PMELONG, :PMELAT) : (:RAJ, :DECJ, :PMRA, :PMDEC)
        comp(1; flags = (c.ecliptic_coordinates ? 1 : 0) | (c.planet_shapiro ? 2 : 0),
            par = [(i - 1, off(b, s)) for (i, s) in enumerate((n..., :PX, :POSEPOCH))])
    end
describe(b::Builder, ::SolarWindDispersion) =
    comp(2; par = [(0, off(b, :SWEPOCH)), (1, off(b, :NE_SW))], len = [(0, len(b, :NE_SW))])
describe(b::Builder, ::DispersionTaylor) =
    comp(3; par = [(0, off(b, :DMEPOCH)), (1, off(b, :DM))], len = [(0, len(b, :DM))])
describe(b::Builder, c::DispersionPiecewise) =
    comp(4; par = [(0, off(b, :DMX_))], len = [(0, len(b, :DMX_))], mask = [(0, add_index_mask!(b, c.dmx_mask))])
describe(b::Builder, ::ChromaticTaylor) =
    comp(5; par = [(0, off(b, :CMEPOCH)), (1, off(b, :CM)), (2, off(b, :TNCHROMIDX))], len = [(0, len(b, :CM))])
describe(b::Builder, c::ChromaticPiecewise) =
    comp(6; par = [(0, off(b, :CMX_)), (2, off(b, :TNCHROMIDX))], len = [(0, len(b, :CMX_))],
        mask = [(0, add_index_mask!(b, c.cmx_mask))])
function binary_par(b::Builder, first::Symbol, use_fbx::Bool, rest)
    par = [(0, off(b, first))]
    use_fbx ? push!(par, (3, off(b, :FB))) : append!(par, [(1, off(b, :PB)), (2, off(b, :PBDOT))])
    append!(par, [(3 + i, off(b, s)) for (i, s) in enumerate(rest)])
    return par, (use_fbx ? [(0, len(b, :FB))] : [])
end
describe(b::Builder, c::BinaryELL1) =
    let (par, l) = binary_par(b, :TASC, c.use_fbx, (:A1, :A1DOT, :EPS1, :EPS2, :EPS1DOT, :EPS2DOT, :M2, :SINI))
        comp(7; flags = c.use_fbx ? 4 : 0, par = par, len = l)
    end
describe(b::Builder, c::BinaryDD) =
    let (par, l) = binary_par(b, :T0, c.use_fbx,
            (:A1, :A1DOT, :ECC, :EDOT, :OM, :OMDOT, :DR, :DTH, :GAMMA, :M2, :SINI))
        comp(8; flags = c.use_fbx ? 4 : 0, par = par, len = l)
    end
describe(b::Builder, c::BinaryELL1H) =
    let (par, l) = binary_par(b, :TASC, c.use_fbx, (:A1, :A1DOT, :EPS1, :EPS2, :EPS1DOT, :EPS2DOT, :H3, :STIGMA))
        comp(19; flags = c.use_fbx ? 4 : 0, par = par, len = l)
    end
describe(b::Builder, c::BinaryELL1k) =
    let (par, l) = binary_par(b, :TASC, c.use_fbx, (:A1, :A1DOT, :EPS1, :EPS2, :OMDOT, :LNEDOT, :M2, :SINI))
        comp(20; flags = c.use_fbx ? 4 : 0, par = par, len = l)
    end
const DD_COMMON = (:A1, :A1DOT, :ECC, :EDOT, :OM, :OMDOT, :DR, :DTH, :GAMMA)
describe(b::Builder, c::BinaryDDH) =
    let (par, l) = binary_par(b, :T0, c.use_fbx, (DD_COMMON..., :H3, :STIGMA))
        comp(21; flags = c.use_fbx ? 4 : 0, par = par, len = l)
    end
describe(b::Builder, c::BinaryDDS) =
    let (par, l) = binary_par(b, :T0, c.use_fbx, (DD_COMMON..., :M2, :SHAPMAX))
        comp(22; flags = c.use_fbx ? 4 : 0, par = par, len = l)
    end
describe(b::Builder, c::BinaryDDK) =
    let pm = c.ecliptic_coordinates ? (:PMELONG, :PMELAT) : (:PMRA, :PMDEC),
        (par, l) = binary_par(b, :T0, c.use_fbx, (DD_COMMON..., :M2, :KIN, :KOM, pm..., :PX))

        comp(23; flags = (c.use_fbx ? 4 : 0) | (c.ecliptic_coordinates ? 1 : 0), par = par, len = l)
    end
describe(b::Builder, ::FrequencyDependent) = comp(24; par = [(0, off(b, :FD))], len = [(0, len(b, :FD))])
function describe(b::Builder, c::FrequencyDependentJump)
    exps = zeros(UInt, b.ntoa)
    exps[1:length(c.exponents)] .= c.exponents
    bits = add_bit_mask!(b, c.jump_mask)
    return comp(25; par = [(0, off(b, :FDJUMP))], len = [(0, len(b, :FDJUMP))], mask = [(0, bits), (1, add_index_mask!(b, exps))])
end
xwavex(b::Builder, kind, pre) =
    comp(kind; par = vcat([(0, off(b, Symbol(pre, :EPOCH))), (1, off(b, Symbol(pre, :SIN_))), (2, off(b, Symbol(pre, :COS_))),
                (3, off(b, Symbol(pre, :FREQ_)))], kind == 28 ? [(4, off(b, :TNCHROMIDX))] : []),
        len = [(0, len(b, Symbol(pre, :SIN_)))])
describe(b::Builder, ::WaveX) = xwavex(b, 26, :WX)
describe(b::Builder, ::DMWaveX) = xwavex(b, 27, :DMWX)
describe(b::Builder, ::CMWaveX) = xwavex(b, 28, :CMWX)
# power-law GPs with free amplitudes used as delay components (gp_noise.jl:120-130,165-183,219-235): the aux table
# carries ln_js widened to Float64 and js = exp(ln_j) evaluated in the field's own type (Float32 for the red GP)
function free_gp(b::Builder, kind, tn, pre, ln_js; idx = false)
    nlog = findfirst(iszero, ln_js) - 1
    aux0 = length(b.aux)
    append!(b.aux, Float64.(collect(ln_js)))
    append!(b.aux, Float64.(map(exp, collect(ln_js))))
    par = [(0, off(b, Symbol(tn, :AMP))), (1, off(b, Symbol(tn, :GAM))), (2, off(b, Symbol(pre, :SIN_))),
        (3, off(b, Symbol(pre, :COS_))), (4, off(b, Symbol(pre, :FREQ))), (5, off(b, Symbol(pre, :EPOCH)))]
    idx && push!(par, (6, off(b, :TNCHROMIDX)))
    return comp(kind; par = par, len = [(0, length(ln_js)), (1, nlog), (2, aux0)])
end
describe(b::Builder, c::PowerlawRedNoiseGP) = free_gp(b, 32, :TNRED, :PLRED, c.ln_js)
describe(b::Builder, c::PowerlawDispersionNoiseGP) = free_gp(b, 33, :TNDM, :PLDM, c.ln_js)
describe(b::Builder, c::PowerlawChromaticNoiseGP) = free_gp(b, 34, :TNCHROM, :PLCHROM, c.ln_js; idx = true)
describe(b::Builder, c::SolarWindDispersionPiecewise) =
    comp(29; par = [(0, off(b, :SWXDM_))], len = [(0, len(b, :SWXDM_))], mask = [(0, add_index_mask!(b, c.swx_mask))],
        dval = (c.conjunction_geometry.x, c.opposition_geometry.x))
describe(b::Builder, c::ExclusiveDispersionOffset) =
    comp(30; par = [(0, off(b, :FDJUMPDM))], len = [(0, len(b, :FDJUMPDM))], mask = [(0, add_index_mask!(b, c.jump_mask))])
describe(b::Builder, c::DispersionOffset) =
    comp(31; par = [(0, off(b, :FDJUMPDM))], len = [(0, len(b, :FDJUMPDM))], mask = [(0, add_bit_mask!(b, c.jump_mask))])
describe(b::Builder, ::Glitch) =
    comp(9; par = [(i - 1, off(b, s)) for (i, s) in enumerate((:GLEP_, :GLPH_, :GLF0_, :GLF1_, :GLF2_, :GLF0D_, :GLTD_))],
        len = [(0, len(b, :GLEP_))])
describe(b::Builder, ::ChromaticExponentialDip) =
    comp(10; par = [(0, off(b, :EXPDIPFREF)), (1, off(b, :EXPDIPEPS)), (2, off(b, :EXPDIPEP_)), (3, off(b, :EXPDIPAMP_)),
            (4, off(b, :EXPDIPIDX_)), (5, off(b, :EXPDIPTAU_))], len = [(0, len(b, :EXPDIPEP_))])
describe(b::Builder, ::Spindown) = comp(11; par = [(0, off(b, :F_)), (1, off(b, :F))], len = [(0, len(b, :F))])
describe(b::Builder, ::PhaseOffset) = comp(12; par = [(0, off(b, :PHOFF))])
describe(b::Builder, c::ExclusivePhaseJump) =
    comp(13; par = [(0, off(b, :JUMP)), (1, off(b, :F_)), (2, off(b, :F))], len = [(0, len(b, :JUMP))],
        mask = [(0, add_index_mask!(b, c.jump_mask))])
describe(b::Builder, c::PhaseJump) =
    comp(14; par = [(0, off(b, :JUMP)), (1, off(b, :F_)), (2, off(b, :F))], len = [(0, len(b, :JUMP))],
        mask = [(0, add_bit_mask!(b, c.jump_mask))])
describe(b::Builder, c::MeasurementNoise) =
    comp(15; par = vcat(has(b, :EFAC) ? [(0, off(b, :EFAC))] : [], has(b, :EQUAD) ? [(1, off(b, :EQUAD))] : []),
        len = vcat(has(b, :EFAC) ? [(0, len(b, :EFAC))] : [], has(b, :EQUAD) ? [(1, len(b, :EQUAD))] : []),
        mask = [(0, add_index_mask!(b, c.efac_index_mask)), (1, add_index_mask!(b, c.equad_index_mask))])
describe(b::Builder, c::ExclusiveDispersionJump) =
    comp(16; par = [(0, off(b, :DMJUMP))], len = [(0, len(b, :DMJUMP))], mask = [(0, add_index_mask!(b, c.jump_mask))])
describe(b::Builder, c::DispersionJump) =
    comp(17; par = [(0, off(b, :DMJUMP))], len = [(0, len(b, :DMJUMP))], mask = [(0, add_bit_mask!(b, c.jump_mask))])
describe(b::Builder, c::DispersionMeasurementNoise) =
    comp(18; par = vcat(has(b, :DMEFAC) ? [(0, off(b, :DMEFAC))] : [], has(b, :DMEQUAD) ? [(1, off(b, :DMEQUAD))] : []),
        len = vcat(has(b, :DMEFAC) ? [(0, len(b, :DMEFAC))] : [], has(b, :DMEQUAD) ? [(1, len(b, :DMEQUAD))] : []),
        mask = [(0, add_index_mask!(b, c.dmefac_index_mask)), (1, add_index_mask!(b, c.dmequad_index_mask))])
describe(::Builder, c) = error("VelaB200: component $(typeof(c)) is outside the implemented hot path")

gp_kind(::PowerlawRedNoiseGP) = (1, :TNREDAMP, :TNREDGAM, :PLREDFREQ)
gp_kind(::PowerlawDispersionNoiseGP) = (2, :TNDMAMP, :TNDMGAM, :PLDMFREQ)
gp_kind(::PowerlawChromaticNoiseGP) = (3, :TNCHROMAMP, :TNCHROMGAM, :PLCHROMFREQ)

# ---- device-side priors (include/vela_b200.h: VelaPriorDesc); families of pyvela/pyvela/priors.py:30-76 -----------
struct VelaPriorDesc
    kind::Int32
    _pad::Int32
    a::Float64
    b::Float64
    lo::Float64
    hi::Float64
end
using Distributions: Uniform, Normal, LogNormal, LogUniform, Truncated
prior_desc(d::Uniform) = VelaPriorDesc(1, 0, d.a, d.b, 0.0, 0.0)
prior_desc(d::Normal) = VelaPriorDesc(2, 0, d.μ, d.σ, 0.0, 0.0)
prior_desc(d::LogNormal) = VelaPriorDesc(3, 0, d.μ, d.σ, 0.0, 0.0)
prior_desc(d::LogUniform) = VelaPriorDesc(4, 0, d.a, d.b, 0.0, 0.0)
prior_desc(::Vela.KINPriorDistribution) = VelaPriorDesc(5, 0, 0.0, 0.0, 0.0, 0.0)
prior_desc(::Vela.SINIPriorDistribution) = VelaPriorDesc(6, 0, 0.0, 0.0, 0.0, 0.0)
prior_desc(::Vela.STIGMAPriorDistribution) = VelaPriorDesc(7, 0, 0.0, 0.0, 0.0, 0.0)
prior_desc(::Vela.SHAPMAXPriorDistribution) = VelaPriorDesc(8, 0, 0.0, 0.0, 0.0, 0.0)
prior_desc(d::Vela.PXPriorDistribution) = VelaPriorDesc(9, 0, d.Rmax, 0.0, 0.0, 0.0)
prior_desc(::Vela.LatitudePriorDistribution) = VelaPriorDesc(10, 0, 0.0, 0.0, 0.0, 0.0)
prior_desc(d::Truncated{<:Normal}) = VelaPriorDesc(11, 0, d.untruncated.μ, d.untruncated.σ, d.lower, d.upper)
prior_desc(::Any) = nothing     # no device twin: lnprior stays on the host for this model

# ---- the wrapped pulsar --------------------------------------------------------------------------------
mutable struct B200Pulsar{M<:TimingModel,T<:TOABase}
    model::M
    toas::Vector{T}
    handle::Ptr{Cvoid}
    device_priors::Bool
end

last_error() = unsafe_string(ccall((:vela_b200_last_error, LIB), Cstring, ()))
check(rc, what) = rc == 0 ? nothing : error("$what failed ($rc): $(last_error())")

function B200Pulsar(model::TimingModel, toas::Vector{T}; devices::Vector{Int32} = Int32[]) where {T<:TOABase}
    ph = model.param_handler
    b = Builder(param_offsets(ph), length(toas), Vector{UInt32}[], Vector{UInt8}[], Float64[])
    comps = VelaComponentDesc[describe(b, c) for c in model.components]
    defaults = copy(ph._default_values)
    free0 = Int32.(ph._free_indices .- 1)
    index_masks = isempty(b.index_masks) ? UInt32[] : reduce(vcat, b.index_masks)
    bit_masks = isempty(b.bit_rows) ? UInt8[] : reduce(vcat, b.bit_rows)
    wide = T <: WidebandTOA
    tzr = [model.tzr_toa]
    aux = b.aux

    kernel = model.kernel
    gps = VelaGPDesc[]
    keep = Any[]            # arrays the descriptor points into
    basis = Matrix{Float64}(undef, 0, 0)
    egroups = EcorrGroup[]
    par_ecorr = Int32(-1)
    kernel_kind = 0
    inner = kernel isa WoodburyKernel ? kernel.inner_kernel : kernel
    if inner isa EcorrKernel
        egroups = inner.ecorr_groups
        has(b, :ECORR) && (par_ecorr = off(b, :ECORR))
    elseif !(inner isa WhiteNoiseKernel)
        error("VelaB200: kernel $(typeof(kernel)) is not on the GPU path yet")
    end
    if kernel isa WoodburyKernel
        kernel_kind = inner isa EcorrKernel ? 3 : 2
        basis = kernel.noise_basis
        for gp in kernel.gp_components
            if gp isa MarginalizedTimingModel
                w = copy(gp.prior_weights_inv)
                push!(keep, w)
                push!(gps, VelaGPDesc(4, length(w), -1, -1, -1, 0, C_NULL, pointer(w)))
            else
                kind, amp, gam, frq = gp_kind(gp)
                lj = Float64.(collect(gp.ln_js))      # Float32 -> Float64 is exact (gp_noise.jl:111)
                push!(keep, lj)
                push!(gps, VelaGPDesc(kind, length(lj), off(b, amp), off(b, gam), off(b, frq), 0, pointer(lj), C_NULL))
            end
        end
    else
        kernel_kind = inner isa EcorrKernel ? 1 : 0
    end

    handle = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve comps defaults free0 toas tzr index_masks bit_masks gps keep basis egroups devices aux begin
        desc = VelaModelDesc(3, length(comps), pointer(comps), length(defaults), length(free0), pointer(defaults),
            pointer(free0), length(toas), wide ? 1 : 0, sizeof(T), pointer(toas), pointer(tzr),
            length(b.index_masks), length(b.bit_rows), pointer(index_masks), pointer(bit_masks), kernel_kind,
            length(gps), pointer(gps), size(basis, 1), size(basis, 2), max(size(basis, 1), 1), pointer(basis),
            length(egroups), par_ecorr, Ptr{UInt64}(pointer(egroups)), length(devices), 0, pointer(devices),
            length(aux), pointer(aux))
        check(ccall((:vela_b200_create, LIB), Cint, (Ref{VelaModelDesc}, Ref{Ptr{Cvoid}}), desc, handle), "vela_b200_create")
    end
    psr = B200Pulsar{typeof(model),T}(model, toas, handle[], false)
    finalizer(p -> ccall((:vela_b200_destroy, LIB), Cint, (Ptr{Cvoid},), p.handle), psr)
    descs = [prior_desc(p.distribution) for p in model.priors]
    if !isempty(descs) && all(!isnothing, descs)
        pd = VelaPriorDesc[descs...]
        check(ccall((:vela_b200_set_priors, LIB), Cint, (Ptr{Cvoid}, Ptr{VelaPriorDesc}, Int32), psr.handle, pd, length(pd)),
            "vela_b200_set_priors")
        psr.device_priors = true
    end
    return psr
end

B200Pulsar(psr::Pulsar; kw...) = B200Pulsar(psr.model, psr.toas; kw...)

# free vector or NamedTuple (wls_likelihood.jl:46-47)
free_vector(psr::B200Pulsar, params::NamedTuple) = read_param_values_to_vector(psr.model, params)
free_vector(::B200Pulsar, params) = collect(Float64, params)

function batch(sym::Symbol, psr::B200Pulsar, paramss::AbstractMatrix, extra...)
    B, nd = size(paramss)
    A = paramss isa StridedMatrix{Float64} ? paramss : Matrix{Float64}(paramss)
    rs, cs = strides(A)                      # Julia Matrix: (1, B); a PyArray from numpy: (nd, 1)
    out = Vector{Float64}(undef, B)
    # the host-side lnprior vector (when there is one) is rooted for the duration of the call like A and out
    lnpr = (sym === :lnpost && !(extra[1] isa Ptr)) ? convert(Vector{Float64}, extra[1]) : nothing
    GC.@preserve A out lnpr begin
        rc = if sym === :lnpost
            ccall((:vela_b200_lnpost_batch, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}),
                psr.handle, pointer(A), B, nd, rs, cs, lnpr === nothing ? extra[1] : pointer(lnpr), pointer(out))
        else
            ccall((:vela_b200_lnlike_batch, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Int64, Ptr{Float64}),
                psr.handle, pointer(A), B, nd, rs, cs, pointer(out))
        end
        check(rc, "vela_b200_$(sym)_batch")
    end
    return out
end

# ---- Vela's API names (src/pulsar/pulsar.jl:14-25, src/likelihood/posterior.jl:4-46, wls_likelihood.jl:75-104) ----
Vela.calc_lnlike_vectorized(psr::B200Pulsar, paramss) = batch(:lnlike, psr, paramss)
Vela.calc_lnlike(psr::B200Pulsar, params) = batch(:lnlike, psr, reshape(free_vector(psr, params), 1, :))[1]
Vela.calc_lnlike_serial(psr::B200Pulsar, params) = Vela.calc_lnlike(psr, params)
Vela.calc_lnprior(psr::B200Pulsar, params) = calc_lnprior(psr.model, params)
Vela.prior_transform(psr::B200Pulsar, cube) = Vela.prior_transform(psr.model, cube)

function Vela.calc_lnpost_vectorized(psr::B200Pulsar, paramss)
    # device-side priors when every family has a device twin (lnprior == C_NULL); otherwise priors stay on the
    # host exactly as in posterior.jl:9-12.  Either way the library applies the -Inf short-circuit.
    psr.device_priors && return batch(:lnpost, psr, paramss, Ptr{Float64}(C_NULL))
    lnpr = Float64[calc_lnprior(psr.model, view(paramss, ii, :)) for ii = 1:size(paramss, 1)]
    return batch(:lnpost, psr, paramss, lnpr)
end
Vela.calc_lnpost(psr::B200Pulsar, params) =
    Vela.calc_lnpost_vectorized(psr, reshape(free_vector(psr, params), 1, :))[1]
Vela.calc_lnpost_serial(psr::B200Pulsar, params) = Vela.calc_lnpost(psr, params)

Vela.get_lnlike_serial_func(psr::B200Pulsar) = params -> Vela.calc_lnlike(psr, params)
Vela.get_lnlike_parallel_func(psr::B200Pulsar) = params -> Vela.calc_lnlike(psr, params)
Vela.get_lnlike_func(psr::B200Pulsar) = params -> Vela.calc_lnlike(psr, params)
Vela.get_lnpost_func(psr::B200Pulsar, vectorize::Bool = false) =
    vectorize ? (paramss -> Vela.calc_lnpost_vectorized(psr, paramss)) : (params -> Vela.calc_lnpost(psr, params))

"""(y, Ninvdiag) — `_calc_resids_and_Ninvdiag`, gls_likelihood.jl:9-79; pyvela calls it at spnta.py:383."""
function Vela._calc_resids_and_Ninvdiag(psr::B200Pulsar, params)
    v = free_vector(psr, params)
    n = ccall((:vela_b200_n_rows, LIB), Int64, (Ptr{Cvoid},), psr.handle)
    y = Vector{Float64}(undef, n)
    w = Vector{Float64}(undef, n)
    check(ccall((:vela_b200_residuals, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}),
            psr.handle, v, length(v), y, w), "vela_b200_residuals")
    return y, w
end

"""
    marginalized_offset_mean_and_covinvcf(psr, paramss) -> (ahat::Matrix (p x B), cf::Array (p x p x B))

Conditional mean of the analytically marginalised amplitudes and the Cholesky factor of Sigma^-1 for a batch, from the
factorisation the likelihood already computed (what pyvela derives on the host with scipy in spnta.py:376-394).
`cf[:, :, b]` is the row-major upper factor of the C ABI read column-major, i.e. the LOWER factor L with
Sigma^-1 = L L'; `UpperTriangular(cf[:, :, b]')` is scipy's `cholesky(lower=False)`.
"""
function marginalized_offset_mean_and_covinvcf(psr::B200Pulsar, paramss::AbstractMatrix)
    B, nd = size(paramss)
    A = paramss isa StridedMatrix{Float64} ? paramss : Matrix{Float64}(paramss)
    rs, cs = strides(A)
    p = ccall((:vela_b200_n_basis, LIB), Int64, (Ptr{Cvoid},), psr.handle)
    ahat = Matrix{Float64}(undef, p, B)
    cf = Array{Float64,3}(undef, p, p, B)
    GC.@preserve A check(ccall((:vela_b200_marginalized_mean_batch, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            psr.handle, pointer(A), B, nd, rs, cs, ahat, cf, C_NULL), "vela_b200_marginalized_mean_batch")
    return ahat, cf
end

end # module
