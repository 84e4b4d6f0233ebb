"""World-size-2 test of the walker sharding + all-gather host logic on CPU (gloo), `-m "not gpu"`.

Each rank evaluates its contiguous walker block (with the oracle standing in for the CUDA library, which
cannot run here) and all ranks must end up with the same full vector as a single-process evaluation.
"""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch.distributed as dist
from __graft_entry__ import load_package
vb = load_package()
import oracle as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
model, toas, _ = vb.readwrite.load_pulsar_npz(os.path.join(ROOT, "tests", "golden", "NGC6440E.npz"))
md = vb.ModelDescriptor(model, toas)
W = np.load(os.path.join(ROOT, "tests", "golden", "NGC6440E.npz"))["golden_walkers"][:37]   # ragged: 37 = 19 + 18
def lnpost(block):
    lp = vb.priors.lnprior(model.priors, block)
    ll = O.lnlike_batch(md, block, nthreads=1)
    return np.where(np.isfinite(lp), lp + ll, lp)
full = vb.parallel.lnpost_sharded(lnpost, W)
np.save(os.path.join(OUT, f"rank{rank}.npy"), full)
lo, hi = vb.parallel.shard_range(len(W), world, rank)
assert (lo, hi) == ((0, 19) if rank == 0 else (19, 37))
dist.destroy_process_group()
'''


def test_two_ranks_gather_matches_single_process(tmp_path, vb, oracle, golden):
    script = tmp_path / "worker.py"
    script.write_text(f"ROOT = {ROOT!r}\nOUT = {str(tmp_path)!r}\n" + WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    g = golden("NGC6440E")
    W = g.walkers[:37]
    lp = vb.priors.lnprior(g.model.priors, W)
    want = np.where(np.isfinite(lp), lp + oracle.lnlike_batch(g.md, W), lp)
    for rank in range(2):
        got = np.load(tmp_path / f"rank{rank}.npy")
        np.testing.assert_array_equal(got, want)
