"""Perf probe (not a test): per-config stage times at the BASELINE.json sizes, plus small-batch latency."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vb = load_package()
names = sys.argv[1:] or ["config2_ell1", "config3_gp", "config4_dd_wb", "config5_array"]
for name in names:
    cfg = vb.synth.CONFIGS[name]
    t0 = time.time()
    ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t))
    psr = vb.Pulsar(ds.model, ds.toas)
    B = min(cfg["walkers"], 8192)
    W = ds.walkers(B, seed=1)
    psr.set_stage_timing(True)
    out = psr.lnlike_vectorized(W)            # warm-up (allocations)
    ts = []
    for _ in range(3):
        t1 = time.perf_counter(); out = psr.lnlike_vectorized(W); ts.append(time.perf_counter() - t1)
    st = psr.last_stage_ms()
    nt, p = len(ds.toas), ds.info["rank"]
    nr = nt * (2 if ds.toas.wideband else 1)
    fl_sigma = (nr * p * (p + 1) + 3 * nr * p) * B
    print(f"{name}: Nt={nt} p={p} nfree={ds.info['nfree']} B={B} gen {t1 - t0:.1f}s | e2e {min(ts) * 1e3:.2f} ms -> {B / min(ts):.0f} evals/s | "
          f"stages ms prologue {st[0]:.3f} chain {st[1]:.3f} sigma {st[2]:.3f} chol {st[3]:.3f} | finite {np.isfinite(out).mean():.3f}"
          + (f" | sigma {fl_sigma / (st[2] * 1e-3) / 1e12:.1f} TF" if p else ""))
    for b in (1, 64, 256):
        ts = []
        for _ in range(5):
            t1 = time.perf_counter(); psr.lnlike_vectorized(W[:b]); ts.append(time.perf_counter() - t1)
        print(f"    B={b}: {min(ts) * 1e3:.3f} ms per call")
