"""Synthetic dataset generator, driven by the oracle evaluator (`-m "not gpu"`)."""
import numpy as np
import pytest

from helpers import oracle_eval_factory


@pytest.mark.parametrize("name,ntoa", [("config2_ell1", 600), ("config3_gp", 800), ("config4_dd_wb", 400),
                                       ("config5_array", 500), ("extra_ell1_fbx", None), ("extra_dd_fbx_wb", None),
                                       ("extra_dmjump_bits_wb", None), ("extra_ell1h_fd_wavex", None),
                                       ("extra_ell1k_swx", None), ("extra_ddk_wb", None), ("extra_ddh", None),
                                       ("extra_dds", None), ("extra_ecorr", None), ("extra_ecorr_gp", None),
                                       ("extra_free_gp", None), ("extra_free_gp_wb", None), ("extra_ecorr_gp_r80", None),
                                       ("extra_ecorr_gp_r140", None), ("extra_mtm_only", None)])
def test_datasets_are_phase_connected_and_finite(name, ntoa, vb, oracle):
    ds = vb.synth.make_dataset(name, oracle_eval_factory(vb, oracle), ntoa=ntoa)
    assert ds.info["connect_rms_s"] < 1e-12          # Newton iterations converge far below 1 ns
    md = vb.ModelDescriptor(ds.model, ds.toas)
    y, w = oracle.residuals(md, ds.truth)
    n = len(ds.toas)
    assert np.all(w > 0)
    assert np.sqrt(np.mean(y[:n] ** 2)) < 1e-4        # white + GP noise, not phase wraps
    assert np.all(ds.toas.pulse_number_lo == 0) and np.all(np.modf(ds.toas.pulse_number_hi)[0] == 0)
    W = ds.walkers(32, seed=3)
    ll = oracle.lnlike_batch(md, W)
    assert np.all(np.isfinite(ll))
    # posterior-width scatter in the timing parameters, narrow priors in the noise hyper-parameters
    assert oracle.lnlike(md, ds.truth) - ll.min() < 1e4


def test_generator_is_deterministic(vb, oracle):
    a = vb.synth.make_dataset("config3_gp", oracle_eval_factory(vb, oracle), ntoa=300)
    b = vb.synth.make_dataset("config3_gp", oracle_eval_factory(vb, oracle), ntoa=300)
    np.testing.assert_array_equal(a.toas.records, b.toas.records)
    np.testing.assert_array_equal(a.walkers(8, 1), b.walkers(8, 1))


def test_config_shapes_match_baseline_json(vb):
    c = vb.synth.CONFIGS
    assert c["config2_ell1"]["ntoa"] == 5000 and c["config2_ell1"]["walkers"] == 4096
    assert c["config3_gp"]["ntoa"] == 10000 and c["config3_gp"]["red"] == 30 and c["config3_gp"]["walkers"] == 8192
    assert c["config4_dd_wb"]["ntoa"] == 20000 and c["config4_dd_wb"]["wideband"] and c["config4_dd_wb"]["walkers"] == 8192
    assert c["config5_array"]["ntoa"] == 100000 and c["config5_array"]["walkers"] == 65536
    assert 2 * (c["config5_array"]["red"] + c["config5_array"]["dmgp"] + c["config5_array"]["chromgp"]) == 300
