"""Perf probe: K3s launch shapes (VELA_B200_COLSUM_SHAPE) and split-K (VELA_B200_SIGMA_KSPLIT)."""
import os, sys, subprocess
exec(open('tools/probes/gpu_plan_probe.py').read().split("cases =")[0])   # writes gpurun_out/_probe.py
cases = [("config3_gp", 10000, 8192), ("config3_gp", 10000, 64), ("config4_dd_wb", 20000, 4096), ("config5_array", 20000, 2048)]
for name, ntoa, B in cases:
    for shape, ks in [(-1, 0), (0, 1), (1, 1), (2, 1), (3, 1), (1, 2), (2, 2), (3, 2), (3, 4)]:
        env = dict(os.environ)
        if shape >= 0:
            env["VELA_B200_COLSUM_SHAPE"] = str(shape)
            env["VELA_B200_SIGMA_KSPLIT"] = str(ks)
        r = subprocess.run([sys.executable, 'gpurun_out/_probe.py', name, str(ntoa), str(B)], env=env, capture_output=True, text=True, timeout=600)
        print(name, ntoa, B, 'shape', shape, 'ks', ks, r.stdout.strip() or r.stderr[-300:], flush=True)
