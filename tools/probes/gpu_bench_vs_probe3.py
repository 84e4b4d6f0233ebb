"""Chain stage of config 3 under variants of process set-up (what does importing torch change?)."""
import os, sys
mode = sys.argv[1]
if mode == "lazyenv":
    os.environ["CUDA_MODULE_LOADING"] = "LAZY"
if mode == "eagerenv_torch":
    os.environ["CUDA_MODULE_LOADING"] = "EAGER"
import numpy as np
sys.path.insert(0, '.')
if mode in ("import_only", "torch_after", "eagerenv_torch", "torch_init_noalloc"):
    import torch
if mode == "torch_init_noalloc":
    torch.cuda.init()
from __graft_entry__ import load_package
vb = load_package()
ds = vb.synth.make_dataset("config3_gp", lambda m, t: vb.Pulsar(m, t))
psr = vb.Pulsar(ds.model, ds.toas)
if mode in ("torch_after", "eagerenv_torch"):
    torch.cuda.init(); torch.zeros(1, device="cuda")
W = ds.walkers(8192, seed=1000)
psr.set_stage_timing(True)
out = []
for i in range(5):
    psr.lnlike_vectorized(W); out.append(psr.last_stage_ms()[1])
print(mode, " ".join(f"{t:.3f}" for t in out), "CUDA_MODULE_LOADING =", os.environ.get("CUDA_MODULE_LOADING"))
