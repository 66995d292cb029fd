"""GPU parity of the FULL IAN graph (reference IAN.py: MADE/IAF latent, MDC blocks, RGB-Beta head) through
API.IAN('IAN.py') -> C-ABI, against the float64 oracle's golden vectors.  Tolerances (float32 path):
  x_hat max-abs <= 2e-4 ; latents |dz| <= 3e-4 * (1 + |z|) (the IAF divides by exp(MADE_ls), |z| reaches ~10)."""
import os

import numpy as np
import pytest

from oracle import ian_full_numpy as fn
from oracle import ian_numpy as on
from oracle import weights as ow

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "ian_full_golden.npz")))


@pytest.fixture(scope="module")
def PF(gold):
    return ow.make_full_weights(int(gold["weight_seed"]))


@pytest.fixture(scope="module")
def full_model(npe, PF):
    m = npe.IAN("IAN.py", dnn=True, weights=PF, device=0)
    yield m
    m.close()


@pytest.fixture(params=["tc", "simt"])
def fm(full_model, request):
    full_model.set_path(request.param)
    yield full_model
    full_model.set_path("tc")


def _zclose(z, ref, k=3e-4):
    return (np.abs(z - ref) <= k * (1.0 + np.abs(ref))).all()


def test_full_encode_golden(fm, gold):
    x = on.to_tanh(gold["images"].astype(np.float64)).astype(np.float32)
    z = fm.encode_images(x)
    assert z.shape == (2, 100) and _zclose(z, gold["z"]), np.abs(z - gold["z"]).max()
    zs = fm.encode(x, eps=gold["eps"])
    amp = np.abs(np.exp(gold["logsigma"]) * gold["eps"])
    assert (np.abs(zs - gold["z_sample"]) <= 1e-3 * (1.0 + np.abs(gold["z_sample"]) + amp)).all()


def test_full_decode_golden(fm, gold):
    xh = fm.sample_at(gold["z_rand"])
    assert xh.shape == (2, 3, 64, 64)
    assert np.abs(xh - gold["xhat_rand"]).max() <= 2e-4
    xh = fm.sample_at(gold["z"].astype(np.float32))
    assert np.abs(xh - gold["xhat"]).max() <= 2e-4


def test_full_reconstruct_and_ordering(full_model, gold, PF):
    assert np.array_equal(full_model.made_ordering, gold["ordering"])
    rng = np.random.default_rng(4)
    x = rng.uniform(-1, 1, (3, 3, 64, 64)).astype(np.float32)
    xh, z = full_model.reconstruct(x, return_z=True)
    masks = fn.made_masks(gold["ordering"].astype(np.float32))
    zr = fn.full_encode(PF, x[:1], masks)
    assert _zclose(z[:1], zr)
    assert np.abs(xh[:1] - fn.full_decode(PF, z[:1])).max() <= 2e-4
    assert np.abs(xh - full_model.sample_at(z)).max() <= 5e-5


def test_full_model_has_no_brush_yet(full_model, npe):
    with pytest.raises(npe.IanError):
        full_model.imgrad(1, 1, 5, 5, np.zeros((1, 100), np.float32))


def test_custom_ordering_changes_masks(npe, PF):
    o = np.arange(100, dtype=np.int32)[::-1].copy()
    m = npe.IAN("IAN.py", True, weights=PF, made_ordering=o)
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, (1, 3, 64, 64)).astype(np.float32)
    z = m.encode_images(x)
    zr = fn.full_encode(PF, x, fn.made_masks(o.astype(np.float32)))
    m.close()
    assert _zclose(z, zr)
    with pytest.raises(npe.IanError):
        npe.IAN("IAN.py", True, weights=PF, made_ordering=np.zeros(100, np.int32))
