"""CPU tests of the oracle itself: the two independent restatements agree, the self-consistency KATs
derivable from the reference text hold, and the committed golden vectors reproduce."""
import numpy as np
import pytest
import torch

from oracle import ian_numpy as on
from oracle import ian_torch as ot


@pytest.fixture(scope="module")
def Pt(weights):
    return ot.to_torch(weights, torch.float64)


def test_numpy_vs_torch_encode_decode(weights, Pt, golden):
    x = on.to_tanh(golden["images"][:2].astype(np.float64)).astype(np.float32)
    mu, ls = on.simple_encode_mu_ls(weights, x)
    mu_t, ls_t = ot.encode_mu_ls(Pt, torch.from_numpy(x).double())
    assert np.abs(mu - mu_t.numpy()).max() < 1e-11
    assert np.abs(ls - ls_t.numpy()).max() < 1e-11
    xh = on.simple_decode(weights, mu)
    xh_t = ot.decode(Pt, torch.from_numpy(mu))
    assert np.abs(xh - xh_t.numpy()).max() < 1e-12
    # float32 torch (the timed CPU baseline) stays within fp32 noise of the float64 definition
    P32 = ot.to_torch(weights, torch.float32)
    xh32 = ot.decode(P32, ot.encode(P32, torch.from_numpy(x)))
    assert np.abs(xh - xh32.numpy()).max() < 2e-5


def test_golden_reproduces(weights, golden):
    x = on.to_tanh(golden["images"][:2].astype(np.float64)).astype(np.float32)
    mu, ls = on.simple_encode_mu_ls(weights, x)
    assert np.array_equal(mu, golden["mu"][:2]) or np.abs(mu - golden["mu"][:2]).max() < 1e-12
    assert np.abs(ls - golden["logsigma"][:2]).max() < 1e-12
    xh = on.simple_decode(weights, golden["z_rand"][:2])
    assert np.abs(xh - golden["xhat_rand"][:2]).max() < 1e-6      # golden stored as float32


def test_kat_deconv_two_reference_formulations_agree(Pt):
    """IAN_simple.py:141-181 (cuDNN GradI) vs :183-223 (TransposedConv2D crop=1 + slice [1:,1:])."""
    z = torch.randn(2, 100, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    a = ot.decode(Pt, z, ot.deconv)
    b = ot.decode(Pt, z, ot.deconv_tc2d_slice)
    assert (a - b).abs().max().item() < 1e-12


def test_kat_deconv_is_input_gradient_of_true_convolution():
    """layers.py:470-481: GpuDnnConvGradI of a conv_mode='conv' (flipped kernel) stride-2 pad-2 conv."""
    g = torch.Generator().manual_seed(1)
    W = torch.randn(6, 4, 5, 5, dtype=torch.float64, generator=g)     # (Cin_deconv, Cout_deconv, 5, 5)
    x = torch.randn(2, 6, 4, 4, dtype=torch.float64, generator=g)
    y = torch.zeros(2, 4, 8, 8, dtype=torch.float64, requires_grad=True)
    # true convolution = correlation with the flipped kernel; W acts as (out=6, in=4)
    fwd = torch.nn.functional.conv2d(y, W.flip(2, 3), stride=2, padding=2)
    (gi,) = torch.autograd.grad(fwd, y, grad_outputs=x)
    mine = on.deconv5x5_s2(x.numpy(), W.numpy())
    assert np.abs(mine - gi.numpy()).max() < 1e-12
    assert np.abs(mine - ot.deconv(x, W).numpy()).max() < 1e-12


def test_kat_deconv_backward_is_adjoint():
    rng = np.random.default_rng(2)
    W = rng.standard_normal((5, 3, 5, 5))
    x = rng.standard_normal((2, 5, 4, 4))
    dy = rng.standard_normal((2, 3, 8, 8))
    lhs = (on.deconv5x5_s2(x, W) * dy).sum()
    rhs = (x * on.deconv5x5_s2_bwd_data(dy, W)).sum()
    assert abs(lhs - rhs) < 1e-9 * max(1.0, abs(lhs))


def test_gradients_match_autograd_and_finite_differences(weights, Pt, golden):
    z = golden["z_rand"][:2].astype(np.float64)
    c1, r1, c2, r2 = [int(v) for v in golden["boxes"][0]]
    rgb = np.broadcast_to(golden["rgb"][0].reshape(1, 3, 1, 1), (1, 3, 64, 64)).astype(np.float64)
    g = on.simple_imgradRGB(weights, c1, r1, c2, r2, rgb, z)
    gt = ot.imgradRGB(Pt, c1, r1, c2, r2, torch.from_numpy(rgb.copy()), torch.from_numpy(z))
    assert np.abs(g - gt.numpy()).max() < 1e-14
    assert np.all(g[1] == 0)                               # only sample 0 is differentiated (API.py:64)
    gl = on.simple_imgrad(weights, c1, r1, c2, r2, z)
    glt = ot.imgrad(Pt, c1, r1, c2, r2, torch.from_numpy(z))
    assert np.abs(gl - glt.numpy()).max() < 1e-14
    # central finite differences of the float64 forward on a few coordinates

    def loss(zz):
        xh = on.simple_decode(weights, zz)
        return ((rgb[0, :, r1:r2, c1:c2] - xh[0, :, r1:r2, c1:c2]) ** 2).mean()
    for j in (0, 17, 99):
        e = np.zeros_like(z)
        e[0, j] = 1e-5
        fd = (loss(z + e) - loss(z - e)) / 2e-5
        assert abs(fd - g[0, j]) < 1e-7 + 1e-5 * abs(g[0, j])


def test_batched_grad_is_vmap_of_single(weights, golden):
    z, boxes, rgb = golden["z_rand"][:3], golden["boxes"][:3], golden["rgb"][:3]
    gb = on.simple_grad_batched(weights, z, boxes, rgb)
    for k in range(3):
        frame = np.broadcast_to(rgb[k].reshape(1, 3, 1, 1), (1, 3, 64, 64))
        gk = on.simple_imgradRGB(weights, *boxes[k], frame, z[k:k + 1])
        assert np.abs(gb[k] - gk[0]).max() < 1e-15


def test_box_scalar_semantics(weights, golden):
    z = golden["z_rand"][:1]
    a = on.simple_imgrad(weights, 3, 4, 9, 10, z)
    b = on.simple_imgrad(weights, 3.0, 4.0, 9.0, 10.0, z)     # NPE.py:202 passes integral floats
    assert np.array_equal(a, b)
    with pytest.raises(TypeError):
        on.simple_imgrad(weights, 3.5, 4, 9, 10, z)


def test_tanh_roundtrip_and_display_truncation():
    u = np.arange(256, dtype=np.float64)
    assert np.abs(on.from_tanh(on.to_tanh(u)) - u).max() < 1e-12
    assert on.to_tanh(np.array([0.0, 255.0])).tolist() == [-1.0, 1.0]


def test_edit_loop_torch_vs_numpy(weights, golden):
    P32 = ot.to_torch(weights, torch.float32)
    z = torch.from_numpy(golden["z_rand"][:2])
    boxes = torch.from_numpy(golden["boxes"][:2].astype(np.int64))
    rgb = torch.from_numpy(golden["rgb"][:2])
    zt = ot.edit_loop(P32, z, boxes, rgb, n_steps=2, weight=0.05).numpy()
    zn = on.simple_edit_loop(weights, golden["z_rand"][:2], golden["boxes"][:2], golden["rgb"][:2], n_steps=2)
    assert np.abs(zt - zn).max() < 1e-5
