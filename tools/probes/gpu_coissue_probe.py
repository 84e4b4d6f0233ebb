"""FP64 co-issue probe: can DFMA (scalar FP64) and DMMA (FP64 tensor path) warps of one SM run side by side?"""
import ctypes as C
import sys
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vb = load_package()
L = vb.lib.lib()
out = (C.c_double * 4)()
rc = L.vela_b200_debug_fp64_coissue(0, out)
full_dfma = L.vela_b200_fp64_peak_tflops(0, 0)
full_dmma = L.vela_b200_fp64_peak_tflops(0, 1)
print(f"rc {rc}")
print(f"all warps DFMA: {full_dfma:.2f} TFLOP/s   all warps DMMA: {full_dmma:.2f} TFLOP/s")
print(f"half the warps DFMA, the rest idle: {out[0]:.2f} TFLOP/s")
print(f"half the warps DMMA, the rest idle: {out[1]:.2f} TFLOP/s")
print(f"both halves at once: DFMA {out[2]:.2f} + DMMA {out[3]:.2f} = {out[2] + out[3]:.2f} TFLOP/s "
      f"({(out[2] + out[3]) / max(out[0], out[1], 1e-9):.2f} x the better half alone)")
