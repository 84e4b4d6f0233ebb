"""Perf probe: K4 thread count (VELA_B200_CHOL_THREADS) at several ranks."""
import os, sys, subprocess
exec(open('tools/probes/gpu_plan_probe.py').read().split("cases =")[0])   # writes gpurun_out/_probe.py
cases = [("config3_gp", 10000, 8192), ("config3_gp", 10000, 64), ("config4_dd_wb", 20000, 4096), ("config5_array", 20000, 2048)]
for name, ntoa, B in cases:
    for th in (0, 64, 128, 256):
        env = dict(os.environ)
        if th:
            env["VELA_B200_CHOL_THREADS"] = str(th)
        r = subprocess.run([sys.executable, 'gpurun_out/_probe.py', name, str(ntoa), str(B)], env=env, capture_output=True, text=True, timeout=600)
        print(name, ntoa, B, 'threads', th, r.stdout.strip() or r.stderr[-300:], flush=True)
