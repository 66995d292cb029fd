"""Drop-in check at the caller's level: a scripted Neural-Photo-Editor session (the reference's Tk event handlers
restated headlessly in oracle/npe_replay.py: infer, paint x3, scroll, sample, paint-on-sample, Reset) is run once over
the CUDA-backed API.IAN and once over the float64 oracle behind the same surface.  Same call sequence, dtypes
(float64 Z after sample(), integral-float box coordinates) and blends; the two traces must agree."""
import numpy as np
import pytest

from oracle import npe_replay as nr

pytestmark = pytest.mark.gpu


def _zclose(a, b, k):
    return (np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)) <= k * (1.0 + np.abs(b))).all()


def test_scripted_npe_session(model, weights, golden):
    img = golden["images"][0]                                    # CelebAValid[420], NPE's default image (NPE.py:44)
    ref = nr.scripted_session(nr.OracleModel(weights), img)
    got = nr.scripted_session(model, img)
    assert got["z_dtype_after_sample"] == ref["z_dtype_after_sample"] == "float64"
    assert _zclose(got["z_infer"], ref["z_infer"], 2e-4)
    # uint8(from_tanh(.)) truncates: a value within 1e-5 of an integer may land one step lower on either side
    for k in ("recon", "im_paint"):
        d = np.abs(got[k].astype(np.int32) - ref[k].astype(np.int32))
        assert d.max() <= 1 and d.mean() <= 0.02, (k, d.max(), d.mean())
    # latents after strokes: gradient steps are ~5e-3 in size; allow the oracle tolerance plus 3 % of the stroke
    moved = np.abs(ref["z_paint"] - ref["z_infer"]).max()
    assert moved > 1e-3
    assert np.abs(got["z_paint"] - ref["z_paint"]).max() <= 2e-4 + 0.03 * moved
    assert np.abs(got["z_scroll"] - ref["z_scroll"]).max() <= 2e-4 + 0.03 * np.abs(ref["z_scroll"] - ref["z_paint"]).max() + 0.03 * moved
    # the sampled branch starts from the same float64 noise on both sides
    assert np.abs(got["z_sample_paint"] - ref["z_sample_paint"]).max() <= 1e-3
    assert _zclose(got["z_reset"], ref["z_reset"], 2e-4)
    assert got["display_reset"].shape == (256, 256, 3) and np.array_equal(got["display_reset"], ref["display_reset"])
