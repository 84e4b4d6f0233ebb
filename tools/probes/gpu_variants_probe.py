"""A/B stage times of one synthetic config under environment-selected kernel variants (one subprocess per variant, so
every handle is created under its own environment): name ntoa walkers.  Prints one line per variant."""
import os
import subprocess
import sys

name, ntoa, B = sys.argv[1], sys.argv[2], sys.argv[3]
VARIANTS = [
    ("default", {}),
    ("chol_warp", {"VELA_B200_CHOL_WARP": "1"}),
    ("mb4", {"VELA_B200_JIT_MINBLOCKS": "4"}),
    ("mb6", {"VELA_B200_JIT_MINBLOCKS": "6"}),
    ("unroll2_mb5", {"VELA_B200_JIT_UNROLL": "2"}),
    ("unroll2_mb4", {"VELA_B200_JIT_UNROLL": "2", "VELA_B200_JIT_MINBLOCKS": "4"}),
    ("unroll2_mb3", {"VELA_B200_JIT_UNROLL": "2", "VELA_B200_JIT_MINBLOCKS": "3"}),
    ("t256_mb2", {"VELA_B200_CHAIN_THREADS": "256", "VELA_B200_JIT_MINBLOCKS": "2"}),
    ("t256_mb3", {"VELA_B200_CHAIN_THREADS": "256", "VELA_B200_JIT_MINBLOCKS": "3"}),
    ("t64_mb10", {"VELA_B200_CHAIN_THREADS": "64", "VELA_B200_JIT_MINBLOCKS": "10"}),
    ("group16", {"VELA_B200_CHAIN_GROUP": "16"}),
    ("group64", {"VELA_B200_CHAIN_GROUP": "64"}),
]
only = sys.argv[4].split(",") if len(sys.argv) > 4 else None
here = os.path.dirname(os.path.abspath(__file__))
for tag, env in VARIANTS:
    if only and tag not in only:
        continue
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(here, "gpu_stage_probe.py"), name, ntoa, B], env=e,
                       capture_output=True, text=True, timeout=600)
    out = (r.stdout.strip().splitlines() or ["(no output)"])[-1]
    print(f"{tag:14s} {out}" + ("" if r.returncode == 0 else f"  [rc {r.returncode}] {r.stderr[-300:]}"), flush=True)
