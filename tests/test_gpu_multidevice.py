"""One handle, several devices (what the single-process Julia host uses): walkers are split into contiguous
blocks across the devices named in the descriptor; results must equal the single-device ones bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_devices_in_one_handle(vb, golden):
    if vb.lib.lib().vela_b200_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    for name in ("NGC6440E", "sim3_gp"):
        g = golden(name)
        one = vb.Pulsar(g.model, g.toas, devices=[0])
        two = vb.Pulsar(g.model, g.toas, devices=[0, 1])
        W = np.resize(g.walkers, (101, g.walkers.shape[1]))            # ragged split: 51 + 50
        # each device sees a smaller block than the single-device run (possibly another split-K factor): equal to rounding
        np.testing.assert_allclose(two.lnlike_vectorized(W), one.lnlike_vectorized(W), rtol=1e-13)
        np.testing.assert_allclose(two.lnpost_vectorized(W[:3]), one.lnpost_vectorized(W[:3]), rtol=1e-13)
        np.testing.assert_array_equal(two.lnlike_vectorized(W[:1]), one.lnlike_vectorized(W[:1]))


def test_bad_device_is_an_error(vb, golden):
    g = golden("NGC6440E")
    with pytest.raises(vb.VelaB200Error, match="not available"):
        vb.Pulsar(g.model, g.toas, devices=[99])
