"""save_pulsar_npz -> load_pulsar_npz for every component the kernels support, and the Truncated prior of JLSO files
(reference: src/readwrite/readwrite.jl:1-12 and pyvela/pyvela/priors.py:245-256 for what gets written)."""
import ctypes as C

import numpy as np
import pytest

from helpers import oracle_eval_factory

CASES = ["extra_ell1h_fd_wavex", "extra_ell1k_swx", "extra_ddk_wb", "extra_dmjump_bits_wb", "extra_ell1_fbx", "extra_ecorr_gp"]


@pytest.mark.parametrize("name", CASES)
def test_npz_roundtrip_keeps_the_descriptor(name, vb, oracle, tmp_path):
    ds = vb.synth.make_dataset(name, oracle_eval_factory(vb, oracle), ntoa=300)
    path = str(tmp_path / "psr.npz")
    vb.readwrite.save_pulsar_npz(path, ds.model, ds.toas)
    model, toas, _ = vb.readwrite.load_pulsar_npz(path)
    assert [type(c).__name__ for c in model.components] == [type(c).__name__ for c in ds.model.components]
    md0, md1 = vb.ModelDescriptor(ds.model, ds.toas), vb.ModelDescriptor(model, toas)
    W = ds.walkers(3, seed=1)
    np.testing.assert_array_equal(oracle.lnlike_batch(md1, W), oracle.lnlike_batch(md0, W))
    for f in ("n_index_masks", "n_bit_mask_rows", "n_components", "n_params", "n_free", "kernel_kind"):
        assert getattr(md0.desc, f) == getattr(md1.desc, f)


def test_truncated_normal_prior_from_jlso_struct(vb):
    J = vb.jlso
    normal = J.JLStruct(J.JLType("Normal", ()), {"mu": 1.5, "sigma": 0.25})
    tr = J.JLStruct(J.JLType("Truncated", ()), {"untruncated": normal, "lower": 1.0, "upper": None})
    d = vb.readwrite._distribution(tr)
    assert isinstance(d, vb.priors.TruncatedNormal) and d.lower == 1.0 and d.upper == np.inf
    assert np.isfinite(d.logpdf(1.2)) and d.logpdf(0.9) == -np.inf
    # and through the portable container, open bound included
    spec = vb.readwrite._dist_spec(d)
    d2 = vb.readwrite._dist_from(spec)
    assert d2 == d
    # the device descriptor accepts it
    descs = vb.descriptor.prior_descriptors((vb.priors.SimplePrior("X", d),))
    assert descs is not None
