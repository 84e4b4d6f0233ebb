"""PINT-free reader for Vela.jl's JLSO pulsar files (model + TOAs).

The reference stores a `(model::TimingModel, toas::Vector{TOA|WidebandTOA})` pair with
`JLSO.save(filename, :model => model, :toas => toas)` (src/readwrite/readwrite.jl:1-12).  A JLSO file is a
BSON document whose `objects` are gzip-compressed Julia `Serialization` streams.  Neither Julia nor
JLSO.jl/BSON.jl exist in this environment, so this module restates the two public container
formats:

* BSON (bsonspec.org v1.1) — only the element types JLSO writes (documents, arrays, strings,
  binary, int32/int64, bool, null, double);
* Julia `Serialization` stream format (stdlib Serialization, `ser_version` byte 0x1a as written by
  Julia 1.11.2, per the file's own metadata) — the tag table and the back-reference slot rules of
  `Serialization.handle_deserialize` / `deserialize_datatype` / `deserialize_array`.

Struct field layouts are NOT stored in a Serialization stream; they come from the reference's
struct definitions, listed in `_FIELDS` with the file:line of each definition.

This is a data-format reader ("PINT-free JLSO -> descriptor loader"); it does no
timing arithmetic.
"""

from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass, field
from typing import Any, Dict, List, Tuple

import numpy as np

__all__ = ["JLStruct", "JLType", "JLSymbol", "load_jlso", "read_bson", "deserialize"]


# --------------------------------------------------------------------------------------------
# BSON
# --------------------------------------------------------------------------------------------
def _bson_cstring(buf: bytes, pos: int) -> Tuple[str, int]:
    end = buf.index(b"\x00", pos)
    return buf[pos:end].decode("utf8"), end + 1


def _bson_document(buf: bytes, pos: int, as_array: bool = False):
    (size,) = struct.unpack_from("<i", buf, pos)
    end = pos + size
    pos += 4
    out: Dict[str, Any] = {}
    while pos < end - 1:
        etype = buf[pos]
        pos += 1
        key, pos = _bson_cstring(buf, pos)
        if etype == 0x01:
            (val,) = struct.unpack_from("<d", buf, pos)
            pos += 8
        elif etype == 0x02:
            (n,) = struct.unpack_from("<i", buf, pos)
            val = buf[pos + 4 : pos + 4 + n - 1].decode("utf8")
            pos += 4 + n
        elif etype == 0x03:
            val, pos = _bson_document(buf, pos)
        elif etype == 0x04:
            val, pos = _bson_document(buf, pos, as_array=True)
        elif etype == 0x05:
            (n,) = struct.unpack_from("<i", buf, pos)
            val = buf[pos + 5 : pos + 5 + n]
            pos += 5 + n
        elif etype == 0x08:
            val = bool(buf[pos])
            pos += 1
        elif etype == 0x0A:
            val = None
        elif etype == 0x10:
            (val,) = struct.unpack_from("<i", buf, pos)
            pos += 4
        elif etype == 0x12:
            (val,) = struct.unpack_from("<q", buf, pos)
            pos += 8
        else:
            raise ValueError(f"unsupported BSON element type 0x{etype:02x} at {pos}")
        out[key] = val
    if as_array:
        return [out[str(i)] for i in range(len(out))], end
    return out, end


def read_bson(buf: bytes) -> Dict[str, Any]:
    doc, _ = _bson_document(buf, 0)
    return doc


def _bson_jl(obj):
    """Collapse BSON.jl's tagged encoding of Julia values (`{tag: ...}` documents) for the few
    shapes JLSO uses: Dict{Symbol/String, ...}, Vector{UInt8}, strings, numbers."""
    if isinstance(obj, dict):
        tag = obj.get("tag")
        if tag == "array" and "data" in obj:
            data = obj["data"]
            if isinstance(data, (bytes, bytearray)):
                return bytes(data)
            return [_bson_jl(x) for x in data]
        if tag == "symbol":
            return obj["name"]
        if tag == "struct":
            # Dict: data = [slots?]; BSON.jl lowers Dict to struct(type, data=[keys, vals])
            data = obj.get("data")
            if isinstance(data, list) and len(data) == 2:
                keys = _bson_jl(data[0])
                vals = _bson_jl(data[1])
                if isinstance(keys, list) and isinstance(vals, list) and len(keys) == len(vals):
                    return {str(k): v for k, v in zip(keys, vals)}
            return {k: _bson_jl(v) for k, v in obj.items()}
        return {k: _bson_jl(v) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_bson_jl(x) for x in obj]
    return obj


# --------------------------------------------------------------------------------------------
# Julia Serialization stream
# --------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class JLSymbol:
    name: str

    def __str__(self) -> str:  # pragma: no cover - cosmetic
        return self.name


@dataclass
class JLType:
    name: str
    module: Tuple[str, ...] = ()
    params: Tuple[Any, ...] = ()

    def __hash__(self):
        return hash((self.name, self.module))


@dataclass
class JLStruct:
    type: JLType
    fields: Dict[str, Any] = field(default_factory=dict)

    @property
    def typename(self) -> str:
        return self.type.name

    def __getattr__(self, item):
        try:
            return self.__dict__["fields"][item]
        except KeyError as exc:
            raise AttributeError(item) from exc


# Tag table (stdlib/Serialization/src/Serialization.jl `TAGS`), 1-based like Julia.
_T = {
    "SYM": 0x01, "INT8": 0x02, "UINT8": 0x03, "INT16": 0x04, "UINT16": 0x05, "INT32": 0x06,
    "UINT32": 0x07, "INT64": 0x08, "UINT64": 0x09, "INT128": 0x0A, "UINT128": 0x0B,
    "FLOAT16": 0x0C, "FLOAT32": 0x0D, "FLOAT64": 0x0E, "CHAR": 0x0F, "DATATYPE": 0x10,
    "UNION": 0x11, "UNIONALL": 0x12, "TYPENAME": 0x13, "TUPLE": 0x14, "ARRAY": 0x15,
    "MODULE": 0x1F, "STRING": 0x21, "SIMPLEVECTOR": 0x22,
    "UNDEFREF": 0x29, "BACKREF": 0x2A, "LONGBACKREF": 0x2B, "SHORTBACKREF": 0x2C,
    "LONGTUPLE": 0x2D, "LONGSYMBOL": 0x2E, "LONGEXPR": 0x2F, "LONGSTRING": 0x30,
    "SHORTINT64": 0x31, "FULL_DATATYPE": 0x32, "WRAPPER_DATATYPE": 0x33, "OBJECT": 0x34,
    "REF_OBJECT": 0x35, "FULL_GLOBALREF": 0x36, "HEADER": 0x37, "IDDICT": 0x38,
    "SHARED_REF": 0x39,
}
_VALUE_TAGS = 68  # first "value" tag: the empty tuple ()
_PRIM_TYPES = {
    0x02: ("Int8", "<b", 1), 0x03: ("UInt8", "<B", 1), 0x04: ("Int16", "<h", 2),
    0x05: ("UInt16", "<H", 2), 0x06: ("Int32", "<i", 4), 0x07: ("UInt32", "<I", 4),
    0x08: ("Int64", "<q", 8), 0x09: ("UInt64", "<Q", 8), 0x0D: ("Float32", "<f", 4),
    0x0E: ("Float64", "<d", 8),
}
_NP_DTYPE = {"Int8": "i1", "UInt8": "u1", "Int16": "<i2", "UInt16": "<u2", "Int32": "<i4",
             "UInt32": "<u4", "Int64": "<i8", "UInt64": "<u8", "Float32": "<f4",
             "Float64": "<f8", "Bool": "u1"}
# value tags >= 68 that matter; symbol constants we do not need are returned as JLSymbol("#tagN")
_VALUE_CONSTS = {68: (), 69: JLType("Bool", ("Core",)), 70: JLType("Any", ("Core",)), 76: False,
                 77: True, 78: None, 79: JLSymbol("Any"), 80: JLSymbol("Array"),
                 83: JLSymbol("Tuple"), 155: JLSymbol("Core"), 156: JLSymbol("Base")}

# number of type parameters of the parametric types that occur in Vela pulsar files
_NPARAMS = {
    "TimingModel": 3, "Tuple": None, "GQ": 2, "DoubleFloat": 1, "SimplePrior": 2,
    "SimplePriorMulti": 3, "Uniform": 1, "LogNormal": 1, "LogUniform": 1, "Normal": 1,
    "Truncated": 5, "ParamHandler": 1, "NamedTuple": 2, "WoodburyKernel": 2,
    "PowerlawRedNoiseGP": 1, "PowerlawDispersionNoiseGP": 1, "PowerlawChromaticNoiseGP": 1,
    "Array": 2, "BitArray": 1, "Pulsar": 2,
}

# struct field names, from the reference definitions
_FIELDS = {
    # src/model/timing_model.jl:13-24
    "TimingModel": ["pulsar_name", "ephem", "clock", "units", "epoch", "components", "kernel",
                    "param_handler", "tzr_toa", "priors"],
    "SolarSystem": ["ecliptic_coordinates", "planet_shapiro"],   # src/model/solarsystem.jl:25-28
    "SolarWindDispersion": [], "DispersionTaylor": [], "ChromaticTaylor": [], "Spindown": [],
    "PhaseOffset": [], "Glitch": [], "ChromaticExponentialDip": [], "WhiteNoiseKernel": [],
    "DispersionPiecewise": ["dmx_mask"],                        # src/model/dispersion.jl:39-41
    "ChromaticPiecewise": ["cmx_mask"],                         # src/model/chromatic.jl:42-44
    "BinaryELL1": ["use_fbx"], "BinaryDD": ["use_fbx"],        # binary_ell1.jl:11-13, binary_dd.jl:11-13
    "BinaryELL1H": ["use_fbx"], "BinaryELL1k": ["use_fbx"], "BinaryDDH": ["use_fbx"], "BinaryDDS": ["use_fbx"],
    "BinaryDDK": ["use_fbx", "ecliptic_coordinates"],          # binary_ddk.jl:15-18
    "FrequencyDependent": [], "WaveX": [], "DMWaveX": [], "CMWaveX": [],
    "FrequencyDependentJump": ["jump_mask", "exponents"],      # frequency_dependent.jl:46-49
    "SolarWindDispersionPiecewise": ["swx_mask", "conjunction_geometry", "opposition_geometry"],  # solarwind.jl:64-68
    "DispersionOffset": ["jump_mask"], "ExclusiveDispersionOffset": ["jump_mask"],                # fdjumpdm.jl:14-16,44-46
    "PhaseJump": ["jump_mask"], "ExclusivePhaseJump": ["jump_mask"],   # src/model/jump.jl:14-16,31-33
    "DispersionJump": ["jump_mask"], "ExclusiveDispersionJump": ["jump_mask"],  # dmjump.jl:28-30,62-64
    "MeasurementNoise": ["efac_index_mask", "equad_index_mask"],        # measurement_noise.jl:14-17
    "DispersionMeasurementNoise": ["dmefac_index_mask", "dmequad_index_mask"],
    "EcorrKernel": ["ecorr_groups"], "EcorrGroup": ["start", "stop", "index"],  # kernel.jl:25-44
    "WoodburyKernel": ["inner_kernel", "gp_components", "noise_basis"],  # kernel.jl:74-77
    "PowerlawRedNoiseGP": ["ln_js"], "PowerlawDispersionNoiseGP": ["ln_js"],
    "PowerlawChromaticNoiseGP": ["ln_js"],                      # gp_noise.jl:110-115,155-160,207-212
    "MarginalizedTimingModel": ["prior_weights_inv", "marginalized_param_names"],
    # src/parameter/parameter.jl:24-32,58-61,126-133
    "Parameter": ["name", "default_value", "dimension", "frozen", "original_units",
                  "unit_conversion_factor", "noise"],
    "MultiParameter": ["name", "parameters"],
    "ParamHandler": ["single_params", "multi_params", "_default_params_tuple", "_default_values",
                     "_free_indices", "_nfree"],
    # src/toa/toa.jl:39-45, src/toa/ephemeris.jl:14-23, src/toa/wideband_toa.jl:18-21,85-88
    "TOA": ["value", "error", "observing_frequency", "pulse_number", "ephem", "index"],
    "SolarSystemEphemeris": ["ssb_obs_pos", "ssb_obs_vel", "obs_sun_pos", "obs_jupiter_pos",
                             "obs_saturn_pos", "obs_venus_pos", "obs_uranus_pos",
                             "obs_neptune_pos"],
    "WidebandTOA": ["toa", "dminfo"], "DMInfo": ["value", "error"],
    "GQ": ["x"], "DoubleFloat": ["hi", "lo"],
    # src/prior/simple_prior.jl:17-20,39-42
    "SimplePrior": ["distribution", "source_type"],
    "SimplePriorMulti": ["distribution", "source_type"],
    # Distributions.jl public struct fields
    "Uniform": ["a", "b"], "LogUniform": ["a", "b"], "LogNormal": ["mu", "sigma"],
    "Normal": ["mu", "sigma"],
    "Truncated": ["untruncated", "lower", "upper", "loglcdf", "lcdf", "ucdf", "tp", "logtp"],
    "KINPriorDistribution": [], "SINIPriorDistribution": [], "STIGMAPriorDistribution": [],
    "SHAPMAXPriorDistribution": [], "LatitudePriorDistribution": [],
    "PXPriorDistribution": ["Rmax"],                            # src/prior/known_priors.jl:82-84
    "BitArray": ["chunks", "len", "dims"],
}
_PRIMITIVE_ENUMS = {"PriorSourceType": 4}   # @enum => 32-bit primitive, src/prior/simple_prior.jl:3
# isbits element types whose arrays are written as raw memory: name -> number of 8-byte words
_ISBITS_WORDS = {"TOA": 31, "WidebandTOA": 33, "EcorrGroup": 3, "GQ": 1}


class _Reader:
    def __init__(self, buf: bytes):
        self.buf = buf
        self.pos = 0
        self.table: Dict[int, Any] = {}
        self.counter = 0

    # -- primitives
    def u8(self) -> int:
        v = self.buf[self.pos]
        self.pos += 1
        return v

    def unpack(self, fmt: str, n: int):
        (v,) = struct.unpack_from(fmt, self.buf, self.pos)
        self.pos += n
        return v

    def take(self, n: int) -> bytes:
        v = self.buf[self.pos : self.pos + n]
        if len(v) != n:
            raise EOFError("truncated Serialization stream")
        self.pos += n
        return v

    def new_slot(self) -> int:
        s = self.counter
        self.counter += 1
        return s

    # -- entry
    def deserialize(self):
        return self.handle(self.u8())

    def symbol(self, n: int) -> JLSymbol:
        sym = JLSymbol(self.take(n).decode("utf8"))
        if n > 7:  # Serialization.deserialize_symbol -> resolve_ref_immediately
            self.table[self.new_slot()] = sym
        return sym

    def handle(self, b: int):
        T = _T
        if b == 0:  # a tag written "as a value" (write_as_tag)
            return self.desertag(self.u8())
        if b >= _VALUE_TAGS:
            return self.desertag(b)
        if b == T["TUPLE"]:
            return tuple(self.deserialize() for _ in range(self.u8()))
        if b == T["LONGTUPLE"]:
            return tuple(self.deserialize() for _ in range(self.unpack("<i", 4)))
        if b == T["SHORTBACKREF"]:
            return self.table[self.unpack("<H", 2)]
        if b == T["BACKREF"]:
            return self.table[self.unpack("<i", 4)]
        if b == T["LONGBACKREF"]:
            return self.table[self.unpack("<q", 8)]
        if b == T["ARRAY"]:
            return self.array()
        if b == T["DATATYPE"]:
            return self.datatype()
        if b == T["OBJECT"]:
            t = self.deserialize()
            return self.object(t)
        if b == T["REF_OBJECT"]:
            slot = self.new_slot()
            t = self.deserialize()
            obj = self.object(t)
            self.table[slot] = obj
            return obj
        if b == T["SHARED_REF"]:
            slot = self.new_slot()
            obj = self.deserialize()
            self.table[slot] = obj
            return obj
        if b == T["SYM"]:
            return self.symbol(self.u8())
        if b == T["LONGSYMBOL"]:
            return self.symbol(self.unpack("<i", 4))
        if b == T["STRING"]:
            return self.take(self.u8()).decode("utf8")
        if b == T["LONGSTRING"]:
            return self.take(self.unpack("<q", 8)).decode("utf8")
        if b == T["SHORTINT64"]:
            return self.unpack("<i", 4)
        if b == T["HEADER"]:
            self.header()
            return self.deserialize()
        if b == T["MODULE"]:
            return self.module()
        if b in _PRIM_TYPES:
            _, fmt, n = _PRIM_TYPES[b]
            return self.unpack(fmt, n)
        if b == T["UINT128"]:
            return int.from_bytes(self.take(16), "little")
        if b == T["INT128"]:
            return int.from_bytes(self.take(16), "little", signed=True)
        raise ValueError(f"unsupported Serialization tag 0x{b:02x} at offset {self.pos - 1}")

    def desertag(self, b: int):
        if b in _PRIM_TYPES:
            return JLType(_PRIM_TYPES[b][0], ("Core",))
        if b in _VALUE_CONSTS:
            return _VALUE_CONSTS[b]
        if 190 <= b <= 222:
            return b - 190  # Int32(0..32)
        if 223 <= b <= 255:
            return b - 223  # Int64(0..32)
        if b == _T["TUPLE"]:
            return JLType("Tuple", ("Core",))
        if b == _T["SYM"]:
            return JLType("Symbol", ("Core",))
        if b == _T["STRING"]:
            return JLType("String", ("Core",))
        if b == _T["DATATYPE"]:
            return JLType("DataType", ("Core",))
        if b == _T["ARRAY"]:
            return JLType("Array", ("Core",))
        return JLSymbol(f"#tag{b}")

    def header(self):
        # "JL" + version byte + flags byte + 3 reserved (Serialization.writeheader); HEADER_TAG consumed
        magic = self.take(2)
        if magic != b"JL":
            raise ValueError("not a Julia Serialization stream")
        self.version = self.u8()
        self.take(4)

    def module(self):
        names: List[str] = []
        while True:
            b = self.u8()
            if b == 68:  # EMPTYTUPLE_TAG terminates serialize_mod_names
                break
            v = self.handle(b)
            if isinstance(v, JLSymbol):
                names.append(v.name)
            # uuid (UInt128) / nothing are skipped
        return tuple(names)

    def datatype(self) -> JLType:
        slot = self.new_slot()
        name = self.deserialize()
        mod = self.deserialize()
        if not isinstance(name, JLSymbol):
            raise ValueError(f"datatype name is not a symbol: {name!r}")
        params: Tuple[Any, ...] = ()
        if name.name in _NPARAMS:
            np_ = self.unpack("<i", 4)
            params = tuple(self.deserialize() for _ in range(np_))
        elif name.name not in _FIELDS and name.name not in _PRIMITIVE_ENUMS and \
                name.name not in ("String", "Symbol", "Float64", "Int64", "UInt64", "Bool"):
            raise ValueError(
                f"unknown Julia type {'.'.join(mod) if mod else ''}.{name.name}: add its parameter "
                "count/field list to vela jlso._NPARAMS/_FIELDS")
        t = JLType(name.name, tuple(mod) if mod else (), params)
        self.table[slot] = t
        return t

    def object(self, t):
        if not isinstance(t, JLType):
            raise ValueError(f"object of non-type {t!r} at {self.pos}")
        if t.name in _PRIMITIVE_ENUMS:
            return int.from_bytes(self.take(_PRIMITIVE_ENUMS[t.name]), "little", signed=True)
        if t.name in ("Float64",):
            return self.unpack("<d", 8)
        if t.name == "NamedTuple":
            names = [s.name for s in t.params[0]]
            return JLStruct(t, {n: self.deserialize() for n in names})
        if t.name == "Tuple":
            return tuple(self.deserialize() for _ in range(len(t.params)))
        if t.name not in _FIELDS:
            raise ValueError(f"no field list for Julia struct {t.name}")
        return JLStruct(t, {n: self.deserialize() for n in _FIELDS[t.name]})

    def array(self):
        slot = self.new_slot()
        d1 = self.deserialize()   # Serialization.deserialize_array: a Type here is the eltype,
        if isinstance(d1, JLType):  # otherwise eltype is UInt8 and d1 is already the length/dims
            elty = d1
            dims = self.deserialize()
        else:
            elty = JLType("UInt8", ("Core",))
            dims = d1
        shape = (dims,) if isinstance(dims, int) else tuple(dims)
        n = int(np.prod(shape)) if len(shape) else 1
        name = elty.name if isinstance(elty, JLType) else str(elty)
        if name == "Bool":
            out = np.empty(n, dtype=bool)  # run-length encoded (serialize_array_data)
            i = 0
            while i < n:
                c = self.u8()
                cnt = c & 0x7F
                out[i : i + cnt] = bool(c >> 7)
                i += cnt
            arr: Any = out.reshape(shape, order="F")
        elif name in _NP_DTYPE:
            dt = np.dtype(_NP_DTYPE[name])
            arr = np.frombuffer(self.take(n * dt.itemsize), dtype=dt).reshape(shape, order="F").copy()
        elif name in _ISBITS_WORDS:
            w = _ISBITS_WORDS[name]
            raw = np.frombuffer(self.take(n * w * 8), dtype="<f8").reshape(n, w).copy()
            arr = JLStruct(JLType("RawArray", (), (elty,)), {"data": raw, "eltype": name})
        else:
            arr = [None] * n
            self.table[slot] = arr
            for i in range(n):
                b = self.u8()
                arr[i] = None if b == _T["UNDEFREF"] else self.handle(b)
        self.table[slot] = arr
        return arr


def deserialize(buf: bytes):
    """Deserialize one Julia Serialization stream."""
    return _Reader(buf).deserialize()


def _bson_dict(doc) -> Dict[str, Any]:
    """A Julia Dict lowered by BSON.jl: {tag: struct, type: ..Dict.., data: [keys, vals]}."""
    if isinstance(doc, dict) and doc.get("tag") == "struct" and isinstance(doc.get("data"), list) \
            and len(doc["data"]) == 2 and isinstance(doc["data"][0], list):
        keys, vals = doc["data"]
        return {str(_bson_jl(k)): v for k, v in zip(keys, vals)}
    return doc


def _bson_int_array(obj) -> List[int]:
    if isinstance(obj, dict) and obj.get("tag") == "array":
        data = obj["data"]
        if isinstance(data, (bytes, bytearray)):
            return [int(v) for v in np.frombuffer(bytes(data), dtype="<i8")]
        return [int(v) for v in data]
    return [int(v) for v in obj]


def load_jlso(path: str) -> Dict[str, Any]:
    """Read a JLSO v4 file; returns {object name: deserialized value} (here 'model', 'toas').

    Layout (JLSO.jl file format 4.0.0, as found in test/datafiles/*.jlso): a BSON document holding a
    Dict with `object-names`, `object-nbytes`, `object-pos` (0-based byte offsets) and `metadata`,
    followed by the (gzip-compressed) serialized objects at those offsets."""
    with open(path, "rb") as fh:
        raw = fh.read()
    (doc_len,) = struct.unpack_from("<i", raw, 0)
    top = _bson_dict(read_bson(raw[:doc_len]))
    names = [str(_bson_jl(n)) for n in top["object-names"]]
    nbytes = _bson_int_array(top["object-nbytes"])
    pos = _bson_int_array(top["object-pos"])
    meta = _bson_dict(top.get("metadata", {}))
    meta = {k: (v if not isinstance(v, (bytes, bytearray)) else bytes(v)) for k, v in meta.items()}
    comp = meta.get("compression", "gzip")
    fmt = meta.get("format", "julia_serialize")
    if fmt != "julia_serialize":
        raise ValueError(f"unsupported JLSO object format {fmt!r}")
    out: Dict[str, Any] = {}
    for name, nb, p0 in zip(names, nbytes, pos):
        blob = raw[p0 : p0 + nb]
        if str(comp).startswith("gzip"):
            blob = zlib.decompress(blob, 31)
        elif comp not in ("none", None, ""):
            raise ValueError(f"unsupported JLSO compression {comp!r}")
        out[name] = deserialize(blob)
    out["__metadata__"] = {k: v for k, v in meta.items() if k != "manifest"}
    return out
