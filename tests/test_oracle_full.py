"""CPU tests of the full-IAN oracle (reference IAN.py graph): numpy-f64 vs torch agree, the KATs derivable from
the reference text hold (MDCL sum-of-convs == composite kernel, layers.py:138-150 vs 207-258; MADE mask structure,
SURVEY Appendix D), golden vectors reproduce."""
import os

import numpy as np
import pytest
import torch

from oracle import ian_full_numpy as fn
from oracle import ian_torch as ot
from oracle import weights as ow

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "ian_full_golden.npz")))


@pytest.fixture(scope="module")
def PF(gold):
    return ow.make_full_weights(int(gold["weight_seed"]))


def test_made_mask_structure():
    o = fn.made_ordering()
    assert sorted(o.astype(int).tolist()) == list(range(100))
    M0, M1, Md = fn.made_masks(o)
    first = int(np.where(o == 0)[0][0])
    assert M0.sum() == 100 and np.all(M0[first] == 1)               # only the first-in-order latent feeds the hidden layer
    assert np.array_equal(M1, np.broadcast_to((o >= 1)[None, :], (100, 100)).astype(np.float32))
    assert np.array_equal(Md, (o[:, None] < o[None, :]).astype(np.float32))     # strict autoregressive direct links
    assert Md.sum() == 4950 and set(np.unique(M0)) <= {0.0, 1.0}


def test_made_is_autoregressive(PF):
    """the MADE stack itself (made_core): output j depends only on inputs i with ordering[i] < ordering[j]."""
    o = fn.made_ordering()
    masks = fn.made_masks(o)
    rng = np.random.default_rng(0)
    z = rng.standard_normal((1, 100))
    base = fn.made_core(PF, "l_IAF_mu", z, masks)
    i = int(np.where(o == 50)[0][0])
    z2 = z.copy()
    z2[0, i] += 1.0
    d = np.abs(fn.made_core(PF, "l_IAF_mu", z2, masks) - base)[0]
    assert np.all(d[o <= 50] == 0) and np.any(d[o > 50] > 0)


def test_made_as_wired_in_the_reference_graph(PF):
    """as wired (layers.py:769: the MADE layer is fed its own input MaskedLayer, see made_forward) the layer's output is
    made_core(relu(z W0*M0 + b0)); with one hidden layer of all-ones connectivity M0 keeps only the input whose
    ordering is 0, so the whole flow reads a single latent -- what the executed reference shows, restated here."""
    o = fn.made_ordering()
    masks = fn.made_masks(o)
    rng = np.random.default_rng(0)
    z = rng.standard_normal((1, 100))
    base = fn.made_forward(PF, "l_IAF_mu", z, masks)
    assert np.abs(base - fn.made_core(PF, "l_IAF_mu", fn.made_input_layer(PF, "l_IAF_mu", z, masks), masks)).max() == 0
    i0 = int(np.where(o == 0)[0][0])
    z2 = z.copy()
    z2[0, np.arange(100) != i0] += 1.0
    assert np.abs(fn.made_forward(PF, "l_IAF_mu", z2, masks) - base).max() == 0
    z3 = z.copy()
    z3[0, i0] += 1.0
    assert np.abs(fn.made_forward(PF, "l_IAF_mu", z3, masks) - base).max() > 0


def test_mdcl_equals_composite_kernel(PF):
    P64 = ot.to_torch(PF, torch.float64)
    g = torch.Generator().manual_seed(3)
    for name, C, scales in (("dec_conv4a", 128, [0, 2, 3]), ("R", 128, [2, 3, 4]), ("B_b", 4, [2, 3, 4])):
        x = torch.randn(1, C, 12, 12, dtype=torch.float64, generator=g)
        a = ot.mdcl(P64, name, x, scales)
        b = ot.mdcl_composite(P64, name, x, scales)
        assert (a - b).abs().max().item() < 1e-12
        c = fn.mdcl(PF, name, x.numpy(), scales)
        assert np.abs(c - a.numpy()).max() < 1e-12


def test_numpy_vs_torch_and_golden(PF, gold):
    masks = fn.made_masks(gold["ordering"].astype(np.float32))
    P64 = ot.to_torch(PF, torch.float64)
    mt = [torch.from_numpy(m).double() for m in masks]
    z = ot.full_latent(P64, torch.from_numpy(gold["mu"][:1]), mt).numpy()
    assert np.abs(z - gold["z"][:1]).max() < 1e-11
    xh = ot.full_decode(P64, torch.from_numpy(gold["z_rand"][:1]).double()).numpy()
    assert np.abs(xh - gold["xhat_rand"][:1]).max() < 1e-6          # golden stored as float32
    xn = fn.full_decode(PF, gold["z_rand"][:1])
    assert np.abs(xn - xh).max() < 1e-12


def test_v1_numpy_vs_torch():
    P1 = ow.make_v1_weights(0)
    z = np.random.default_rng(3).standard_normal((1, 100)).astype(np.float32)
    a = fn.v1_decode(P1, z)
    b = ot.v1_decode(ot.to_torch(P1, torch.float64), torch.from_numpy(z).double()).numpy()
    assert np.abs(a - b).max() < 1e-12 and a.shape == (1, 3, 64, 64)
