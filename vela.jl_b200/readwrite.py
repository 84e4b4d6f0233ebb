"""`load_pulsar_data`: read a Vela.jl JLSO pulsar file into the host-side mirror types.

Mirror of `Vela.load_pulsar_data(filename) -> (model, toas)` (src/readwrite/readwrite.jl:9-12), built on
the PINT-free container reader in `jlso.py`.
"""

from __future__ import annotations

from typing import Any, Tuple

import numpy as np

from . import model as M
from . import priors as P
from .jlso import JLStruct, load_jlso


def _gq(v) -> float:
    """GQ{d,Float64} -> float; GQ{d,Double64} -> hi (use _gq_dd for the pair)."""
    x = v.x
    if isinstance(x, JLStruct):
        return float(x.hi)
    return float(x)


def _gq_dd(v) -> Tuple[float, float]:
    x = v.x
    if isinstance(x, JLStruct):
        return float(x.hi), float(x.lo)
    return float(x), 0.0


def _vec3(t) -> Tuple[float, float, float]:
    return tuple(_gq(x) for x in t)


def _toa(js: JLStruct) -> M.TOA:
    e = js.ephem
    ephem = M.SolarSystemEphemeris(*[_vec3(getattr(e, n)) for n in (
        "ssb_obs_pos", "ssb_obs_vel", "obs_sun_pos", "obs_jupiter_pos", "obs_saturn_pos",
        "obs_venus_pos", "obs_uranus_pos", "obs_neptune_pos")])
    return M.TOA(_gq_dd(js.value), _gq(js.error), _gq(js.observing_frequency),
                 _gq_dd(js.pulse_number), ephem, int(js.index))


def _parameter(js: JLStruct) -> M.Parameter:
    return M.Parameter(js.name.name, float(js.default_value), int(js.dimension), bool(js.frozen),
                       str(js.original_units), float(js.unit_conversion_factor), bool(js.noise))


def _mask(a) -> np.ndarray:
    return np.asarray(a).astype(np.uint32)


def _bitmatrix(js) -> np.ndarray:
    """BitMatrix (chunks::Vector{UInt64}, len, dims) -> bool (rows, cols), column-major bit order."""
    if isinstance(js, np.ndarray):
        return js.astype(bool)
    chunks = np.asarray(js.chunks, dtype=np.uint64)
    nbits = int(js.len)
    bits = np.unpackbits(chunks.view(np.uint8), bitorder="little")[:nbits]
    dims = tuple(int(d) for d in js.dims)
    return bits.reshape(dims, order="F").astype(bool)


def _component(js: JLStruct):
    n = js.typename
    if n == "SolarSystem":
        return M.SolarSystem(bool(js.ecliptic_coordinates), bool(js.planet_shapiro))
    if n in ("SolarWindDispersion", "DispersionTaylor", "ChromaticTaylor", "Spindown", "PhaseOffset",
             "Glitch", "ChromaticExponentialDip"):
        return getattr(M, n)()
    if n == "DispersionPiecewise":
        return M.DispersionPiecewise(_mask(js.dmx_mask))
    if n == "ChromaticPiecewise":
        return M.ChromaticPiecewise(_mask(js.cmx_mask))
    if n in ("BinaryELL1", "BinaryDD", "BinaryELL1H", "BinaryELL1k", "BinaryDDH", "BinaryDDS"):
        return getattr(M, n)(bool(js.use_fbx))
    if n == "BinaryDDK":
        return M.BinaryDDK(bool(js.use_fbx), bool(js.ecliptic_coordinates))
    if n in ("FrequencyDependent", "WaveX", "DMWaveX", "CMWaveX"):
        return getattr(M, n)()
    if n == "FrequencyDependentJump":
        return M.FrequencyDependentJump(_bitmatrix(js.jump_mask), _mask(js.exponents))
    if n == "SolarWindDispersionPiecewise":
        return M.SolarWindDispersionPiecewise(_mask(js.swx_mask), _gq(js.conjunction_geometry), _gq(js.opposition_geometry))
    if n == "ExclusiveDispersionOffset":
        return M.ExclusiveDispersionOffset(_mask(js.jump_mask))
    if n == "DispersionOffset":
        return M.DispersionOffset(_bitmatrix(js.jump_mask))
    if n == "ExclusivePhaseJump":
        return M.ExclusivePhaseJump(_mask(js.jump_mask))
    if n == "PhaseJump":
        return M.PhaseJump(_bitmatrix(js.jump_mask))
    if n == "ExclusiveDispersionJump":
        return M.ExclusiveDispersionJump(_mask(js.jump_mask))
    if n == "DispersionJump":
        return M.DispersionJump(_bitmatrix(js.jump_mask))
    if n == "MeasurementNoise":
        return M.MeasurementNoise(_mask(js.efac_index_mask), _mask(js.equad_index_mask))
    if n == "DispersionMeasurementNoise":
        return M.DispersionMeasurementNoise(_mask(js.dmefac_index_mask), _mask(js.dmequad_index_mask))
    raise NotImplementedError(f"component {n} is outside the implemented hot path")


def _gp(js: JLStruct):
    n = js.typename
    if n in ("PowerlawRedNoiseGP", "PowerlawDispersionNoiseGP", "PowerlawChromaticNoiseGP"):
        return getattr(M, n)(ln_js=np.asarray(js.ln_js, dtype=np.float64))
    if n == "MarginalizedTimingModel":
        return M.MarginalizedTimingModel(prior_weights_inv=np.asarray(js.prior_weights_inv),
                                         pnames=[str(s) for s in js.marginalized_param_names])
    raise NotImplementedError(f"GP component {n}")


def _kernel(js: JLStruct):
    n = js.typename
    if n == "WhiteNoiseKernel":
        return M.WhiteNoiseKernel()
    if n == "EcorrKernel":
        g = js.ecorr_groups
        data = g.fields["data"] if isinstance(g, JLStruct) else np.zeros((0, 3))
        return M.EcorrKernel(np.ascontiguousarray(data).view(np.uint64).reshape(-1, 3))
    if n == "WoodburyKernel":
        return M.WoodburyKernel(_kernel(js.inner_kernel), [_gp(g) for g in js.gp_components],
                                np.asarray(js.noise_basis))
    raise NotImplementedError(f"kernel {n}")


def _distribution(js: JLStruct):
    n = js.typename
    f = js.fields
    if n == "Uniform":
        return P.Uniform(float(f["a"]), float(f["b"]))
    if n == "LogUniform":
        return P.LogUniform(float(f["a"]), float(f["b"]))
    if n == "Normal":
        return P.Normal(float(f["mu"]), float(f["sigma"]))
    if n == "LogNormal":
        return P.LogNormal(float(f["mu"]), float(f["sigma"]))
    if n == "PXPriorDistribution":
        return P.PXPriorDistribution(float(f["Rmax"]))
    if n == "Truncated":
        # Distributions.truncated(Normal(mu, sigma), lower, upper): what pyvela writes for bounded user priors
        # (pyvela/pyvela/priors.py:245-256); a missing bound (`nothing`) is infinite
        base = f["untruncated"]
        if not (isinstance(base, JLStruct) and base.typename == "Normal"):
            raise NotImplementedError(f"truncated {getattr(base, 'typename', base)} prior")

        def bound(v, default):
            return default if v is None or isinstance(v, JLStruct) else float(v)

        return P.TruncatedNormal(float(base.fields["mu"]), float(base.fields["sigma"]),
                                 bound(f.get("lower"), -np.inf), bound(f.get("upper"), np.inf))
    if hasattr(P, n):
        return getattr(P, n)()
    raise NotImplementedError(f"prior distribution {n}")


def _prior(js: JLStruct):
    tp = js.type.params
    if js.typename == "SimplePrior":
        return P.SimplePrior(tp[0].name, _distribution(js.distribution), int(js.source_type))
    return P.SimplePriorMulti(tp[0].name, int(tp[1]), _distribution(js.distribution), int(js.source_type))


def model_from_jl(js: JLStruct) -> M.TimingModel:
    ph = js.param_handler
    handler = M.ParamHandler([_parameter(p) for p in ph.single_params],
                             [M.MultiParameter(mp.name.name, [_parameter(p) for p in mp.parameters])
                              for mp in ph.multi_params])
    # the stored flattening must agree with the one recomputed from parameter.jl:87-119
    stored = np.asarray(ph._default_values, dtype=np.float64)
    if stored.shape != handler._default_values.shape or not np.array_equal(stored, handler._default_values):
        raise ValueError("ParamHandler._default_values in the file disagree with the recomputed ordering")
    if not np.array_equal(np.asarray(ph._free_indices), handler._free_indices):
        raise ValueError("ParamHandler._free_indices in the file disagree with the recomputed ordering")
    return M.TimingModel(
        pulsar_name=str(js.pulsar_name), ephem=str(js.ephem), clock=str(js.clock), units=str(js.units),
        epoch=_gq(js.epoch), components=tuple(_component(c) for c in js.components),
        kernel=_kernel(js.kernel), param_handler=handler, tzr_toa=_toa(js.tzr_toa),
        priors=tuple(_prior(p) for p in js.priors))


def toas_from_jl(obj: Any) -> M.TOAs:
    if isinstance(obj, JLStruct) and obj.typename == "RawArray":
        return M.TOAs(obj.fields["data"])
    raise ValueError("unexpected TOA container in JLSO file")


def load_pulsar_data(filename: str) -> Tuple[M.TimingModel, M.TOAs]:
    """readwrite.jl:9-12"""
    data = load_jlso(filename)
    return model_from_jl(data["model"]), toas_from_jl(data["toas"])


# ---------------------------------------------------------------------------------------------
# Portable (numpy .npz) form of a (model, toas) pair — the container used for tests/golden/ and for
# exchanging synthetic datasets.  One JSON document (structure) + named arrays (bulk data).
# ---------------------------------------------------------------------------------------------
import json  # noqa: E402

_MASK_FIELDS = {
    "DispersionPiecewise": ["dmx_mask"], "ChromaticPiecewise": ["cmx_mask"], "PhaseJump": ["jump_mask"],
    "ExclusivePhaseJump": ["jump_mask"], "DispersionJump": ["jump_mask"],
    "ExclusiveDispersionJump": ["jump_mask"], "MeasurementNoise": ["efac_index_mask", "equad_index_mask"],
    "DispersionMeasurementNoise": ["dmefac_index_mask", "dmequad_index_mask"],
    "FrequencyDependentJump": ["jump_mask", "exponents"], "SolarWindDispersionPiecewise": ["swx_mask"],
    "DispersionOffset": ["jump_mask"], "ExclusiveDispersionOffset": ["jump_mask"],
}


def _dist_spec(d) -> dict:
    spec = {"type": type(d).__name__}
    spec.update({k: (float(v) if np.isfinite(v) else ("inf" if v > 0 else "-inf")) for k, v in vars(d).items()
                 if isinstance(v, (int, float))})
    return spec


def _dist_from(spec: dict):
    kw = {k: (float(v) if isinstance(v, str) else v) for k, v in spec.items() if k != "type"}   # "inf" / "-inf": open bounds
    kw.pop("lb", None), kw.pop("ub", None)
    return getattr(P, spec["type"])(**kw)


def save_pulsar_npz(path: str, model: M.TimingModel, toas: M.TOAs, extra: dict = None) -> None:
    arrays = {"toas": toas.records, "tzr_toa": model.tzr_toa.record()}
    comps = []
    for i, c in enumerate(model.components):
        name = type(c).__name__
        spec = {"type": name}
        for f in _MASK_FIELDS.get(name, []):
            key = f"comp{i}_{f}"
            arrays[key] = np.asarray(getattr(c, f))
            spec[f] = key
        for f in ("ecliptic_coordinates", "planet_shapiro", "use_fbx"):
            if hasattr(c, f):
                spec[f] = bool(getattr(c, f))
        for f in ("conjunction_geometry", "opposition_geometry"):
            if hasattr(c, f):
                spec[f] = float(getattr(c, f))
        comps.append(spec)
    k = model.kernel
    if isinstance(k, M.WoodburyKernel):
        gps = []
        for i, g in enumerate(k.gp_components):
            if isinstance(g, M.MarginalizedTimingModel):
                arrays[f"gp{i}_weights_inv"] = g.prior_weights_inv
                gps.append({"type": "MarginalizedTimingModel", "names": g.marginalized_param_names})
            else:
                arrays[f"gp{i}_ln_js"] = g.ln_js
                gps.append({"type": type(g).__name__})
        arrays["noise_basis"] = np.asarray(k.noise_basis)
        if isinstance(k.inner_kernel, M.EcorrKernel):
            arrays["ecorr_groups"] = k.inner_kernel.ecorr_groups
        kspec = {"type": "WoodburyKernel", "inner": type(k.inner_kernel).__name__, "gp": gps}
    elif isinstance(k, M.EcorrKernel):
        arrays["ecorr_groups"] = k.ecorr_groups
        kspec = {"type": "EcorrKernel"}
    else:
        kspec = {"type": "WhiteNoiseKernel"}
    ph = model.param_handler
    pspec = lambda p: [p.name, p.default_value, p.dimension, p.frozen, p.original_units,  # noqa: E731
                       p.unit_conversion_factor, p.noise]
    priors = []
    for pr in model.priors:
        priors.append({"name": pr.name, "index": getattr(pr, "index", None), "source": pr.source_type,
                       "dist": _dist_spec(pr.distribution)})
    doc = {
        "format": "vela_b200_pulsar/1", "pulsar_name": model.pulsar_name, "ephem": model.ephem,
        "clock": model.clock, "units": model.units, "epoch": model.epoch, "components": comps,
        "kernel": kspec, "single_params": [pspec(p) for p in ph.single_params],
        "multi_params": [[mp.name, [pspec(p) for p in mp.parameters]] for mp in ph.multi_params],
        "priors": priors, "extra": extra or {},
    }
    np.savez_compressed(path, __spec__=np.frombuffer(json.dumps(doc).encode(), dtype=np.uint8), **arrays)


def load_pulsar_npz(path: str):
    """Inverse of `save_pulsar_npz`; returns (model, toas, extra)."""
    z = np.load(path)
    doc = json.loads(bytes(z["__spec__"]).decode())
    comps = []
    for spec in doc["components"]:
        cls = getattr(M, spec["type"])
        kw = {}
        for f in _MASK_FIELDS.get(spec["type"], []):
            kw[f] = z[spec[f]]
        for f in ("ecliptic_coordinates", "planet_shapiro", "use_fbx", "conjunction_geometry", "opposition_geometry"):
            if f in spec:
                kw[f] = spec[f]
        comps.append(cls(**kw))
    ks = doc["kernel"]
    if ks["type"] == "WoodburyKernel":
        gps = []
        for i, g in enumerate(ks["gp"]):
            if g["type"] == "MarginalizedTimingModel":
                gps.append(M.MarginalizedTimingModel(prior_weights_inv=z[f"gp{i}_weights_inv"], pnames=g["names"]))
            else:
                gps.append(getattr(M, g["type"])(ln_js=z[f"gp{i}_ln_js"]))
        inner = M.EcorrKernel(z["ecorr_groups"]) if ks["inner"] == "EcorrKernel" else getattr(M, ks["inner"])()
        kernel = M.WoodburyKernel(inner, gps, z["noise_basis"])
    elif ks["type"] == "EcorrKernel":
        kernel = M.EcorrKernel(z["ecorr_groups"])
    else:
        kernel = M.WhiteNoiseKernel()
    mkp = lambda a: M.Parameter(a[0], float(a[1]), int(a[2]), bool(a[3]), a[4], float(a[5]), bool(a[6]))  # noqa: E731
    handler = M.ParamHandler([mkp(a) for a in doc["single_params"]],
                             [M.MultiParameter(n, [mkp(a) for a in ps]) for n, ps in doc["multi_params"]])
    priors = []
    for pr in doc["priors"]:
        d = _dist_from(pr["dist"])
        priors.append(P.SimplePrior(pr["name"], d, pr["source"]) if pr["index"] is None
                      else P.SimplePriorMulti(pr["name"], pr["index"], d, pr["source"]))
    rec = z["tzr_toa"]
    tzr = M.TOA((rec[0], rec[1]), rec[2], rec[3], (rec[4], rec[5]),
                M.SolarSystemEphemeris(*[tuple(rec[6 + 3 * i: 9 + 3 * i]) for i in range(8)]),
                int(rec[30:31].view(np.uint64)[0]))
    model = M.TimingModel(doc["pulsar_name"], doc["ephem"], doc["clock"], doc["units"], doc["epoch"],
                          tuple(comps), kernel, handler, tzr, tuple(priors))
    return model, M.TOAs(z["toas"]), doc.get("extra", {})
