"""Host-side mirror of the reference's data types for the posterior hot path.

Same names, fields and ordering rules as the Julia structs, so that parity tests read like the
reference's own tests.  Everything here is one-time setup data: no timing arithmetic happens in
Python — the arithmetic lives in libvela_b200.so (CUDA).

Reference definitions mirrored (paths relative to the reference tree):
  TOA / SolarSystemEphemeris / WidebandTOA      src/toa/toa.jl:39-72, src/toa/ephemeris.jl:14-49,
                                                src/toa/wideband_toa.jl:18-21,85-88
  Parameter / MultiParameter / ParamHandler     src/parameter/parameter.jl:24-169
  components                                    src/model/*.jl, src/model/binary/*.jl
  kernels                                       src/model/kernel.jl:20-92
  GP components                                 src/model/gp_noise.jl:81-115,155-160,207-212,
                                                src/model/marginalized_timing_model.jl:13-26
  TimingModel                                   src/model/timing_model.jl:13-58
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

TOA_WORDS = 31
WTOA_WORDS = 33


# ---------------------------------------------------------------------------------------------
# TOAs: stored in the reference's isbits record layout (see include/vela_b200.h)
# ---------------------------------------------------------------------------------------------
@dataclass
class SolarSystemEphemeris:
    ssb_obs_pos: Tuple[float, float, float]
    ssb_obs_vel: Tuple[float, float, float]
    obs_sun_pos: Tuple[float, float, float]
    obs_jupiter_pos: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    obs_saturn_pos: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    obs_venus_pos: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    obs_uranus_pos: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    obs_neptune_pos: Tuple[float, float, float] = (0.0, 0.0, 0.0)

    def __post_init__(self):
        v = np.asarray(self.ssb_obs_vel, dtype=float)
        if not float(v @ v) < 1:  # ephemeris.jl:34
            raise AssertionError("Magnitude of ssb_obs_vel should be less than 1.")

    def words(self) -> np.ndarray:
        return np.concatenate([np.asarray(x, dtype=np.float64) for x in (
            self.ssb_obs_pos, self.ssb_obs_vel, self.obs_sun_pos, self.obs_jupiter_pos,
            self.obs_saturn_pos, self.obs_venus_pos, self.obs_uranus_pos, self.obs_neptune_pos)])


@dataclass
class TOA:
    """A single narrowband TOA.  `value` and `pulse_number` are (hi, lo) double-double pairs."""
    value: Tuple[float, float]
    error: float
    observing_frequency: float
    pulse_number: Tuple[float, float]
    ephem: SolarSystemEphemeris
    index: int

    def record(self) -> np.ndarray:
        rec = np.zeros(TOA_WORDS, dtype=np.float64)
        rec[0:2] = self.value
        rec[2] = self.error
        rec[3] = self.observing_frequency
        rec[4:6] = self.pulse_number
        rec[6:30] = self.ephem.words()
        rec[30:31].view(np.uint64)[0] = np.uint64(self.index)
        return rec


def is_tzr(toa: TOA) -> bool:
    return toa.index == 0


def make_tzr_toa(tzrtdb: Tuple[float, float], tzrfreq: float, tzrephem: SolarSystemEphemeris) -> TOA:
    """toa.jl:71-72"""
    return TOA(tzrtdb, 0.0, tzrfreq, (0.0, 0.0), tzrephem, 0)


@dataclass
class DMInfo:
    value: float
    error: float


@dataclass
class WidebandTOA:
    toa: TOA
    dminfo: DMInfo

    def record(self) -> np.ndarray:
        return np.concatenate([self.toa.record(), [self.dminfo.value, self.dminfo.error]])


class TOAs:
    """`Vector{TOA}` / `Vector{WidebandTOA}`: an (n, 31|33) float64 array of records."""

    def __init__(self, records: np.ndarray):
        records = np.ascontiguousarray(records, dtype=np.float64)
        if records.ndim != 2 or records.shape[1] not in (TOA_WORDS, WTOA_WORDS):
            raise ValueError("TOA records must be (n, 31) or (n, 33) float64")
        self.records = records

    @classmethod
    def from_list(cls, toas: Sequence[Union[TOA, WidebandTOA]]) -> "TOAs":
        return cls(np.stack([t.record() for t in toas]))

    @property
    def wideband(self) -> bool:
        return self.records.shape[1] == WTOA_WORDS

    def __len__(self) -> int:
        return self.records.shape[0]

    @property
    def index(self) -> np.ndarray:
        return self.records[:, 30].copy().view(np.uint64)

    # convenient column views
    value_hi = property(lambda s: s.records[:, 0])
    value_lo = property(lambda s: s.records[:, 1])
    error = property(lambda s: s.records[:, 2])
    observing_frequency = property(lambda s: s.records[:, 3])
    pulse_number_hi = property(lambda s: s.records[:, 4])
    pulse_number_lo = property(lambda s: s.records[:, 5])
    dm = property(lambda s: s.records[:, 31])
    dm_error = property(lambda s: s.records[:, 32])


# ---------------------------------------------------------------------------------------------
# Parameters
# ---------------------------------------------------------------------------------------------
@dataclass
class Parameter:
    name: str
    default_value: float
    dimension: int
    frozen: bool
    original_units: str = ""
    unit_conversion_factor: float = 1.0
    noise: bool = False


@dataclass
class MultiParameter:
    name: str
    parameters: List[Parameter]

    def __post_init__(self):
        if len({p.noise for p in self.parameters}) > 1:  # parameter.jl:63
            raise AssertionError("all members of a MultiParameter must agree on `noise`")

    @property
    def noise(self) -> bool:
        return self.parameters[0].noise


class ParamHandler:
    """Free-vector <-> full-vector mapping; ordering per parameter.jl:87-119:
    non-noise singles, non-noise multis, noise singles, noise multis."""

    def __init__(self, single_params: List[Parameter], multi_params: List[MultiParameter]):
        self.single_params = list(single_params)
        self.multi_params = list(multi_params)
        self._offsets: Dict[str, Tuple[int, int]] = {}
        values: List[float] = []
        free: List[int] = []
        self._flat: List[Tuple[Union[Parameter, MultiParameter], Parameter]] = []
        for noise in (False, True):
            for sp in self.single_params:
                if sp.noise == noise:
                    self._offsets[sp.name] = (len(values), 1)
                    if not sp.frozen:
                        free.append(len(values))
                    values.append(sp.default_value)
                    self._flat.append((sp, sp))
            for mp in self.multi_params:
                if mp.noise == noise:
                    self._offsets[mp.name] = (len(values), len(mp.parameters))
                    for p in mp.parameters:
                        if not p.frozen:
                            free.append(len(values))
                        values.append(p.default_value)
                        self._flat.append((mp, p))
        self._default_values = np.asarray(values, dtype=np.float64)
        self._free_indices0 = np.asarray(free, dtype=np.int32)   # 0-based
        self._nfree = len(free)

    # Julia-compatible 1-based view
    @property
    def _free_indices(self) -> np.ndarray:
        return self._free_indices0 + 1

    def offset(self, name: str) -> int:
        return self._offsets[name][0]

    def length(self, name: str) -> int:
        return self._offsets[name][1]

    def has(self, name: str) -> bool:
        return name in self._offsets

    def read_params(self, free_values) -> np.ndarray:
        """parameter.jl:158-169 — returns the full parameter vector (the flattened NamedTuple)."""
        free_values = np.asarray(free_values, dtype=np.float64)
        if free_values.shape != (self._nfree,):
            raise AssertionError(f"expected {self._nfree} free values, got {free_values.shape}")
        values = self._default_values.copy()
        values[self._free_indices0] = free_values
        return values

    def _free_attr(self, fn):
        return [fn(m, p) for (m, p) in self._flat if not p.frozen]

    def get_free_param_names(self) -> List[str]:
        return self._free_attr(lambda m, p: p.name)

    def get_free_param_prefixes(self) -> List[str]:
        return self._free_attr(lambda m, p: m.name)

    def get_free_param_units(self) -> List[str]:
        return self._free_attr(lambda m, p: p.original_units)

    def get_free_param_labels(self, delim: str = "\n") -> List[str]:
        return self._free_attr(lambda m, p: p.name if p.original_units in ("", "1")
                               else f"{p.name}{delim}({p.original_units})")

    def get_scale_factors(self) -> np.ndarray:
        return np.asarray(self._free_attr(lambda m, p: p.unit_conversion_factor))

    def get_num_timing_params(self) -> int:
        return sum(1 for (m, p) in self._flat if not p.frozen and not p.noise)

    def read_param_values_to_vector(self, params: Optional[np.ndarray] = None) -> np.ndarray:
        """parameter.jl:234-261 — free values out of a full parameter vector (default: defaults)."""
        full = self._default_values if params is None else np.asarray(params, dtype=np.float64)
        return full[self._free_indices0].copy()


# ---------------------------------------------------------------------------------------------
# Components (hot-path subset; see include/vela_b200.h for the slot tables)
# ---------------------------------------------------------------------------------------------
class Component:
    pass


@dataclass
class SolarSystem(Component):
    ecliptic_coordinates: bool
    planet_shapiro: bool


@dataclass
class SolarWindDispersion(Component):
    pass


@dataclass
class DispersionTaylor(Component):
    pass


@dataclass
class DispersionPiecewise(Component):
    dmx_mask: np.ndarray


@dataclass
class ChromaticTaylor(Component):
    pass


@dataclass
class ChromaticPiecewise(Component):
    cmx_mask: np.ndarray


@dataclass
class BinaryELL1(Component):
    use_fbx: bool


@dataclass
class BinaryDD(Component):
    use_fbx: bool


@dataclass
class BinaryELL1H(Component):
    use_fbx: bool


@dataclass
class BinaryELL1k(Component):
    use_fbx: bool


@dataclass
class BinaryDDH(Component):
    use_fbx: bool


@dataclass
class BinaryDDS(Component):
    use_fbx: bool


@dataclass
class BinaryDDK(Component):
    use_fbx: bool
    ecliptic_coordinates: bool


@dataclass
class FrequencyDependent(Component):
    pass


@dataclass
class FrequencyDependentJump(Component):
    jump_mask: np.ndarray            # bool (njump, ntoa)
    exponents: np.ndarray            # uint (njump,)


@dataclass
class WaveX(Component):
    pass


@dataclass
class DMWaveX(Component):
    pass


@dataclass
class CMWaveX(Component):
    pass


@dataclass
class SolarWindDispersionPiecewise(Component):
    swx_mask: np.ndarray
    conjunction_geometry: float
    opposition_geometry: float


@dataclass
class DispersionOffset(Component):
    jump_mask: np.ndarray            # bool (njump, ntoa)


@dataclass
class ExclusiveDispersionOffset(Component):
    jump_mask: np.ndarray            # uint (ntoa,)


@dataclass
class Glitch(Component):
    pass


@dataclass
class ChromaticExponentialDip(Component):
    pass


@dataclass
class Spindown(Component):
    pass


@dataclass
class PhaseOffset(Component):
    pass


@dataclass
class PhaseJump(Component):
    jump_mask: np.ndarray            # bool (njump, ntoa)


@dataclass
class ExclusivePhaseJump(Component):
    jump_mask: np.ndarray            # uint (ntoa,)


@dataclass
class MeasurementNoise(Component):
    efac_index_mask: np.ndarray
    equad_index_mask: np.ndarray


@dataclass
class DispersionJump(Component):
    jump_mask: np.ndarray


@dataclass
class ExclusiveDispersionJump(Component):
    jump_mask: np.ndarray


@dataclass
class DispersionMeasurementNoise(Component):
    dmefac_index_mask: np.ndarray
    dmequad_index_mask: np.ndarray


# ---------------------------------------------------------------------------------------------
# Kernels and GP components
# ---------------------------------------------------------------------------------------------
class Kernel:
    pass


@dataclass
class WhiteNoiseKernel(Kernel):
    pass


@dataclass
class EcorrKernel(Kernel):
    """Parsed for completeness (sim2 fixture); the ECORR likelihood is a 'next' row."""
    ecorr_groups: np.ndarray


def _calc_ln_js(Nlin: int, Nlog: int, logfac: float) -> np.ndarray:
    """gp_noise.jl:81-85"""
    log_js_lin = np.log(np.arange(1, Nlin + 1, dtype=np.float64))
    log_js_log = (1.0 / logfac) ** np.arange(Nlog, 0, -1, dtype=np.float64)
    return np.concatenate([log_js_log, log_js_lin])


class _PowerlawGP(Component):
    float32_ln_js = False

    def __init__(self, Nlin: int = 0, Nlog: int = 0, logfac: float = 2.0, ln_js=None):
        if ln_js is None:
            ln_js = _calc_ln_js(Nlin, Nlog, logfac)
        ln_js = np.asarray(ln_js)
        if self.float32_ln_js:
            ln_js = ln_js.astype(np.float32)      # NTuple{N,Float32}, gp_noise.jl:111
        self.ln_js = ln_js.astype(np.float64)     # widened exactly; what Julia's promotion does

    def get_gp_npars(self) -> int:
        return 2 * len(self.ln_js)

    def js(self) -> np.ndarray:
        """exp(ln_j) as the reference evaluates it: in the storage type of the field (gp_noise.jl:41)."""
        if self.float32_ln_js:
            return np.exp(self.ln_js.astype(np.float32)).astype(np.float64)
        return np.exp(self.ln_js)

    def __repr__(self):
        return f"{type(self).__name__}({len(self.ln_js)} harmonics)"


class PowerlawRedNoiseGP(_PowerlawGP):
    float32_ln_js = True
    amp, gam, freq, prefix = "TNREDAMP", "TNREDGAM", "PLREDFREQ", "PLRED"


class PowerlawDispersionNoiseGP(_PowerlawGP):
    amp, gam, freq, prefix = "TNDMAMP", "TNDMGAM", "PLDMFREQ", "PLDM"


class PowerlawChromaticNoiseGP(_PowerlawGP):
    amp, gam, freq, prefix = "TNCHROMAMP", "TNCHROMGAM", "PLCHROMFREQ", "PLCHROM"


class MarginalizedTimingModel(Component):
    def __init__(self, weights=None, pnames=(), prior_weights_inv=None):
        if prior_weights_inv is None:
            prior_weights_inv = 1.0 / np.asarray(weights, dtype=np.float64)   # marginalized_timing_model.jl:19
        self.prior_weights_inv = np.asarray(prior_weights_inv, dtype=np.float64)
        self.marginalized_param_names = list(pnames)
        assert len(self.prior_weights_inv) == len(self.marginalized_param_names)

    def get_gp_npars(self) -> int:
        return len(self.prior_weights_inv)

    def __repr__(self):
        return f"MarginalizedTimingModel({len(self.prior_weights_inv)} parameters)"


def marginalized_param_names_for_gp_noise(prefix: str, N: int) -> List[str]:
    """gp_noise.jl:87-97"""
    return [f"{prefix}{fname}_{ii:04d}" for fname in ("SIN", "COS") for ii in range(1, N + 1)]


def get_marginalized_param_names(gp) -> List[str]:
    if isinstance(gp, MarginalizedTimingModel):
        return list(gp.marginalized_param_names)
    return marginalized_param_names_for_gp_noise(gp.prefix, len(gp.ln_js))


class WoodburyKernel(Kernel):
    """kernel.jl:74-92"""

    def __init__(self, inner_kernel: Kernel, gp_components: Sequence[Any], noise_basis: np.ndarray):
        self.inner_kernel = inner_kernel
        self.gp_components = tuple(gp_components)
        self.noise_basis = np.asfortranarray(noise_basis, dtype=np.float64)
        if sum(g.get_gp_npars() for g in self.gp_components) != self.noise_basis.shape[1]:
            raise AssertionError("sum of GP parameter counts must equal the number of basis columns")

    def __repr__(self):
        return f"WoodburyKernel({self.inner_kernel}, {self.gp_components})"


# ---------------------------------------------------------------------------------------------
# TimingModel
# ---------------------------------------------------------------------------------------------
@dataclass(eq=False)
class TimingModel:
    pulsar_name: str
    ephem: str
    clock: str
    units: str
    epoch: float
    components: Tuple[Component, ...]
    kernel: Kernel
    param_handler: ParamHandler
    tzr_toa: TOA
    priors: Tuple[Any, ...] = field(default_factory=tuple)

    def __post_init__(self):
        if not is_tzr(self.tzr_toa):                        # timing_model.jl:37
            raise AssertionError("Invalid TZR TOA.")
        if len(self.priors) not in (0, self.param_handler._nfree):   # timing_model.jl:40
            raise AssertionError("The number of prior distributions must match the number of free parameters.")

    # forwarding methods, timing_model.jl:58-83
    def read_params(self, values):
        return self.param_handler.read_params(values)

    def get_free_param_names(self):
        return self.param_handler.get_free_param_names()

    def get_free_param_units(self):
        return self.param_handler.get_free_param_units()

    def get_free_param_labels(self, delim="\n"):
        return self.param_handler.get_free_param_labels(delim)

    def get_free_param_prefixes(self):
        return self.param_handler.get_free_param_prefixes()

    def get_num_timing_params(self):
        return self.param_handler.get_num_timing_params()

    def get_scale_factors(self):
        return self.param_handler.get_scale_factors()

    def read_param_values_to_vector(self, params=None):
        return self.param_handler.read_param_values_to_vector(params)

    def get_marginalized_param_names(self) -> List[str]:
        if isinstance(self.kernel, WoodburyKernel):
            out: List[str] = []
            for gp in self.kernel.gp_components:
                out += get_marginalized_param_names(gp)
            return out
        return []

    def __repr__(self):
        return f"TimingModel[{self.components}; {self.kernel}]"
