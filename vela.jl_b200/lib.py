"""ctypes binding of libvela_b200.so (include/vela_b200.h).

The library is the product: if it cannot be loaded, or there is no CUDA device, every compute call
raises — there is no CPU fallback (BASELINE.json north_star).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VELA_B200_LIB", os.path.join(_HERE, "libvela_b200.so"))
_LIB = None

EXPORTS = [
    "vela_b200_device_count", "vela_b200_last_error", "vela_b200_create", "vela_b200_destroy",
    "vela_b200_n_free", "vela_b200_n_rows", "vela_b200_n_basis", "vela_b200_lnlike_batch",
    "vela_b200_lnpost_batch", "vela_b200_chi2_batch", "vela_b200_residuals", "vela_b200_phases",
    "vela_b200_noise_weights_inv", "vela_b200_sigma_and_u", "vela_b200_lnlike_batch_device", "vela_b200_lnpost_batch_device", "vela_b200_launch_count",
    "vela_b200_set_stage_timing", "vela_b200_last_stage_ms", "vela_b200_fp64_peak_tflops",
    "vela_b200_debug_sincos", "vela_b200_set_priors", "vela_b200_lnprior_batch", "vela_b200_prior_transform_batch",
    "vela_b200_marginalized_mean_batch", "vela_b200_chain_specialized", "vela_b200_jit_log",
    "vela_b200_debug_spec_source", "vela_b200_debug_spec_compile", "vela_b200_debug_structured_sigma",
    "vela_b200_sigma_structured", "vela_b200_debug_div3", "vela_b200_debug_fp64_coissue", "vela_b200_debug_two_stage",
    "vela_b200_sigma_info",
]


class VelaB200Error(RuntimeError):
    pass


def build(force: bool = False) -> str:
    """Compile libvela_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    so = os.path.join(_HERE, "libvela_b200.so")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(_HERE, "..", "include", "vela_b200.h"))
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", csrc] + (["-B"] if force else []), check=True, stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise VelaB200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    dp, i64, vp = C.POINTER(C.c_double), C.c_int64, C.c_void_p
    L.vela_b200_device_count.restype = C.c_int
    L.vela_b200_last_error.restype = C.c_char_p
    L.vela_b200_create.argtypes = [vp, C.POINTER(vp)]
    L.vela_b200_destroy.argtypes = [vp]
    L.vela_b200_n_free.argtypes = [vp]
    L.vela_b200_n_rows.argtypes = [vp]
    L.vela_b200_n_rows.restype = i64
    L.vela_b200_n_basis.argtypes = [vp]
    L.vela_b200_n_basis.restype = i64
    L.vela_b200_lnlike_batch.argtypes = [vp, dp, i64, i64, i64, i64, dp]
    L.vela_b200_lnpost_batch.argtypes = [vp, dp, i64, i64, i64, i64, dp, dp]
    L.vela_b200_chi2_batch.argtypes = [vp, dp, i64, i64, i64, i64, dp]
    L.vela_b200_residuals.argtypes = [vp, dp, i64, dp, dp]
    L.vela_b200_phases.argtypes = [vp, dp, i64, dp, dp, dp]
    L.vela_b200_noise_weights_inv.argtypes = [vp, dp, i64, dp]
    L.vela_b200_sigma_and_u.argtypes = [vp, dp, i64, dp, dp]
    L.vela_b200_lnlike_batch_device.argtypes = [vp, vp, i64, vp, vp]
    L.vela_b200_lnpost_batch_device.argtypes = [vp, vp, i64, vp, C.c_int, vp, vp]
    L.vela_b200_launch_count.argtypes = [vp]
    L.vela_b200_launch_count.restype = i64
    L.vela_b200_set_stage_timing.argtypes = [vp, C.c_int]
    L.vela_b200_last_stage_ms.argtypes = [vp, C.c_int]
    L.vela_b200_last_stage_ms.restype = C.c_double
    L.vela_b200_set_priors.argtypes = [vp, vp, C.c_int32]
    L.vela_b200_lnprior_batch.argtypes = [vp, dp, i64, i64, i64, i64, dp]
    L.vela_b200_prior_transform_batch.argtypes = [vp, dp, i64, i64, dp]
    L.vela_b200_marginalized_mean_batch.argtypes = [vp, dp, i64, i64, i64, i64, dp, dp, dp]
    L.vela_b200_chain_specialized.argtypes = [vp, C.c_int]
    L.vela_b200_jit_log.argtypes = [vp]
    L.vela_b200_jit_log.restype = C.c_char_p
    L.vela_b200_debug_spec_source.argtypes = [vp, C.c_char_p, i64]
    L.vela_b200_debug_spec_source.restype = i64
    L.vela_b200_debug_spec_compile.argtypes = [vp]
    L.vela_b200_debug_spec_compile.restype = i64
    L.vela_b200_debug_structured_sigma.argtypes = [vp, dp, dp, C.POINTER(C.c_int32)]
    L.vela_b200_debug_structured_sigma.restype = i64
    L.vela_b200_sigma_structured.argtypes = [vp]
    L.vela_b200_debug_sincos.argtypes = [dp, C.c_int, dp, dp]
    L.vela_b200_debug_div3.argtypes = [i64, C.c_uint64]
    L.vela_b200_debug_div3.restype = i64
    L.vela_b200_debug_fp64_coissue.argtypes = [C.c_int, dp]
    L.vela_b200_debug_two_stage.argtypes = [vp, dp, dp, dp, dp, C.POINTER(C.c_int64)]
    L.vela_b200_debug_two_stage.restype = i64
    L.vela_b200_sigma_info.argtypes = [vp, i64, C.POINTER(C.c_int64)]
    L.vela_b200_sigma_info.restype = C.c_double
    L.vela_b200_fp64_peak_tflops.argtypes = [C.c_int, C.c_int]
    L.vela_b200_fp64_peak_tflops.restype = C.c_double
    _LIB = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().vela_b200_last_error()
        raise VelaB200Error(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")
