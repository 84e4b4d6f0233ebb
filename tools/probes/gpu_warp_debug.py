"""Debug: chol_warp_kernel vs the CTA kernel across batch sizes, graphs on/off, in separate processes."""
import os, subprocess, sys
CODE = r'''
import os, sys, numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vb = load_package()
name, ntoa, B, order = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t), ntoa=ntoa)
psr = vb.Pulsar(ds.model, ds.toas)
W = ds.walkers(B, seed=11)
out = {}
for v in order:
    if v == 'w': os.environ['VELA_B200_CHOL_WARP'] = '1'
    else: os.environ.pop('VELA_B200_CHOL_WARP', None)
    out[v] = psr.lnlike_vectorized(W)
    out[v] = psr.lnlike_vectorized(W)
if len(out) == 2:
    print('max rel diff', np.max(np.abs(out['w'] - out['c']) / np.abs(out['c'])))
else:
    print('ok', list(out.values())[0][:2])
'''
for graphs in ("1", "0"):
    for B in (3, 70, 1500):
        for order in ("w", "cw", "wc"):
            e = dict(os.environ, VELA_B200_GRAPHS=graphs)
            r = subprocess.run([sys.executable, "-c", CODE, "config3_gp", "1500", str(B), order], env=e, capture_output=True, text=True, timeout=300)
            tail = (r.stdout.strip().splitlines() or [""])[-1] if r.returncode == 0 else r.stderr.strip().splitlines()[-1][-200:]
            print(f"graphs={graphs} B={B} order={order}: rc={r.returncode} {tail}", flush=True)
