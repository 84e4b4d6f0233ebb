"""Probe: split-K of the column-sum kernel at the benchmark size."""
import os, sys, subprocess
exec(open('tools/probes/gpu_plan_probe.py').read().split("cases =")[0])   # writes gpurun_out/_probe.py
for name, ntoa, B in [("config3_gp", 10000, 8192), ("config3_gp", 10000, 4096), ("config4_dd_wb", 20000, 4096)]:
    for ks in (0, 1, 2, 3, 4, 5, 6, 8):
        env = dict(os.environ)
        if ks:
            env["VELA_B200_SIGMA_KSPLIT"] = str(ks)
        r = subprocess.run([sys.executable, 'gpurun_out/_probe.py', name, str(ntoa), str(B)], env=env, capture_output=True, text=True, timeout=600)
        print(name, ntoa, B, 'ks', ks, r.stdout.strip() or r.stderr[-300:], flush=True)
