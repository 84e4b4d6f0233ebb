"""The PINT-free JLSO reader against the reference's own fixture files (`-m "not gpu"`; skipped where
/root/reference is absent, e.g. on the GPU box)."""
import os

import numpy as np
import pytest

REF = "/root/reference/test/datafiles"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference fixtures not available")


@pytest.mark.parametrize("name,ntoa,words,kernel", [
    ("NGC6440E", 62, 31, "WhiteNoiseKernel"), ("sim2", 2000, 31, "EcorrKernel"), ("sim3.gp", 2000, 31, "WoodburyKernel"),
    ("sim_dmgp_wb", 500, 33, "WoodburyKernel"), ("sim_sw.wb", 500, 33, "WhiteNoiseKernel")])
def test_load_pulsar_data(name, ntoa, words, kernel, vb):
    model, toas = vb.load_pulsar_data(os.path.join(REF, name + ".jlso"))
    assert toas.records.shape == (ntoa, words)                       # test_*.jl "read_toas"
    np.testing.assert_array_equal(toas.index, np.arange(1, ntoa + 1))
    assert np.all(np.modf(toas.pulse_number_hi)[0] == 0) and np.all(toas.error > 0)
    assert type(model.kernel).__name__ == kernel
    assert model.tzr_toa.index == 0
    assert len(model.priors) == model.param_handler._nfree           # timing_model.jl:40
    assert np.isfinite(vb.priors.lnprior(model.priors, model.read_param_values_to_vector()))
    vb.ModelDescriptor(model, toas)      # every reference fixture maps onto the C-ABI descriptor


def test_golden_npz_is_the_jlso_content(vb, golden):
    model, toas = vb.load_pulsar_data(os.path.join(REF, "sim3.gp.jlso"))
    g = golden("sim3_gp")
    np.testing.assert_array_equal(toas.records, g.toas.records)
    np.testing.assert_array_equal(np.asarray(model.kernel.noise_basis), np.asarray(g.model.kernel.noise_basis))
    np.testing.assert_array_equal(model.param_handler._default_values, g.model.param_handler._default_values)
    # basis layout: [red sin x30, red cos x30, DM sin, DM cos, chrom sin, chrom cos]; DM = red x (1400 MHz/nu)^2
    Mb = np.asarray(model.kernel.noise_basis)
    ratio = Mb[:, 60] / Mb[:, 0]
    good = np.abs(Mb[:, 0]) > 0.1
    assert np.all(ratio[good] > 0.8) and np.all(ratio[good] < 8.5)


def test_sim2_ecorr_groups(vb):
    model, _ = vb.load_pulsar_data(os.path.join(REF, "sim2.jlso"))
    grp = model.kernel.ecorr_groups
    assert grp.shape[1] == 3 and grp[0].tolist() == [1, 8, 1]         # EcorrGroup(start, stop, index), kernel.jl:25-29
    assert set(model.get_free_param_names()) == {"RAJ", "DECJ", "PHOFF", "F0", "F1", "EFAC1", "ECORR1"}
