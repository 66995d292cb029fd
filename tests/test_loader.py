"""The checkpoint contract of IAN.__init__ (reference API.py:18-36, GANcheckpoints.py:33-57).

CPU part (no GPU needed: the C-ABI's handle-free queries): the library's own parameter list equals the list the
EXECUTED reference loader was handed (tests/golden/ref_exec_*.npz: param_names / param_shapes), and the MADE masks the
library derives from the ordering equal the reference MaskGenerator's bit for bit.

GPU part: a checkpoint written to disk in GANcheckpoints.save_weights format -- including the keys a reference trainer
adds that the graph does not own (log_sigma_theta, train_IAN_simple.py:300,564; discriminator weights; metadata) --
loads through `IAN('<dir>/IAN_simple.py', dnn=True)` with no `weights=` argument.
"""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

from oracle import ian_full_numpy as fn
from oracle import ian_numpy as on
from oracle import weights as ow

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
DISCRIMINATOR = {"discrimi.W", "minibatch_discrim.b", "minibatch_discrim.log_weight_scale", "minibatch_discrim.theta"}
KINDS = {"simple": 0, "full": 1, "v1": 2}


@pytest.mark.parametrize("which", ["simple", "v1", "full"])
def test_library_parameter_list_equals_the_executed_reference_loaders(npe, which):
    api = importlib.import_module(npe.__name__ + ".API")
    ref = np.load(os.path.join(GOLD, "ref_exec_%s.npz" % which))
    theirs = {str(n): tuple(int(d) for d in str(s).split()) for n, s in zip(ref["param_names"], ref["param_shapes"])}
    mine = dict(api.model_param_specs(KINDS[which]))
    assert len(mine) == len(api.model_param_specs(KINDS[which]))          # unique names (GANcheckpoints.py:14-16)
    assert set(theirs) - set(mine) == (DISCRIMINATOR & set(theirs))       # the discriminator head is not on the path
    assert set(mine) <= set(theirs)
    for n, shp in mine.items():
        assert theirs[n] == shp, n
    with pytest.raises(ValueError):
        api.model_param_specs(7)


@pytest.mark.parametrize("which", ["v1", "full"])
def test_made_masks_of_the_library_equal_the_reference_maskgenerator_bit_for_bit(npe, which):
    """north_star: 'bit-exact for the MADE mask indexing' -- the C++ derivation itself, not only its float outputs."""
    api = importlib.import_module(npe.__name__ + ".API")
    lib = npe.load()
    ref = np.load(os.path.join(GOLD, "ref_exec_%s.npz" % which))
    o = api.made_ordering()
    assert np.array_equal(o, ref["ordering_mu"].astype(np.int32)) and np.array_equal(o, ref["ordering_ls"].astype(np.int32))
    for which_mask, key in enumerate(("mask_input", "mask_output_W", "mask_output_D")):
        m = np.empty((100, 100), np.uint8)
        assert lib.ian_made_mask(o.ctypes.data_as(C.POINTER(C.c_int32)), 100, which_mask, m.ctypes.data_as(C.c_void_p)) == 0
        theirs = ref[key]
        assert set(np.unique(theirs)) <= {0.0, 1.0}
        assert np.array_equal(m, theirs.astype(np.uint8)), key
        assert np.array_equal(m, np.asarray(fn.made_masks(o)[which_mask]).astype(np.uint8))
    bad = np.zeros(100, np.int32)
    assert lib.ian_made_mask(bad.ctypes.data_as(C.POINTER(C.c_int32)), 99, 0, m.ctypes.data_as(C.c_void_p)) < 0
    assert lib.ian_made_mask(bad.ctypes.data_as(C.POINTER(C.c_int32)), 100, 3, m.ctypes.data_as(C.c_void_p)) < 0


def _trainer_style_checkpoint(path, P):
    """what train_IAN_simple.py writes: the graph's parameters + log_sigma_theta + the discriminator's + metadata"""
    extra = dict(P)
    rng = np.random.default_rng(5)
    extra["log_sigma_theta"] = rng.normal(0, 1, (3, 64, 64)).astype(np.float32)
    extra["discrimi.W"] = rng.normal(0, 0.02, (1024, 1)).astype(np.float32)
    extra["minibatch_discrim.theta"] = rng.normal(0, 0.02, (16384, 100, 5)).astype(np.float16).astype(np.float32)[:16]
    ow.save_checkpoint(path, extra, metadata={"epoch": 3, "itr": 1234})


@pytest.mark.gpu
def test_on_disk_checkpoint_loads_like_the_reference(npe, weights, golden, tmp_path):
    cfg = tmp_path / "IAN_simple.py"                       # API.py:20: weights_fname = config_path[:-3] + '.npz'
    cfg.write_text("# stand-in for the reference config module; the product reads the graph from its file name\n")
    _trainer_style_checkpoint(str(tmp_path / "IAN_simple.npz"), weights)
    m = npe.IAN(str(cfg), dnn=True)
    try:
        assert m.weights_fname == str(tmp_path / "IAN_simple.npz")
        assert set(m.ignored_keys) == {"log_sigma_theta", "discrimi.W", "minibatch_discrim.theta", "metadata"}
        x = on.to_tanh(golden["images"].astype(np.float64)).astype(np.float32)
        z = m.encode_images(x)
        assert np.abs(z - golden["mu"]).max() <= 2e-4
        assert np.abs(m.sample_at(golden["mu"].astype(np.float32)) - golden["xhat"]).max() <= 1e-4
    finally:
        m.close()
    # a file that lacks one of the graph's own parameters is an error here (the reference only logs a warning)
    short = {k: v for k, v in weights.items() if k != "dec_conv2.W"}
    ow.save_checkpoint(str(tmp_path / "IAN_simple.npz"), short)
    with pytest.raises(npe.IanError, match="dec_conv2.W"):
        npe.IAN(str(cfg), dnn=True)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["v1", "full"])
def test_uploaded_made_weights_are_w_times_mask_bit_for_bit(npe, which):
    ref = np.load(os.path.join(GOLD, "ref_exec_%s.npz" % which))
    gold = np.load(os.path.join(GOLD, "ian_%s_golden.npz" % which))
    P = (ow.make_v1_weights if which == "v1" else ow.make_full_weights)(int(gold["weight_seed"]))
    m = npe.IAN("IANv1.py" if which == "v1" else "IAN.py", True, weights=P)
    try:
        got = np.empty((2, 3, 100, 100), np.float32)
        m._check(m._lib.ian_debug_made_weights(m._h, got.ctypes.data_as(C.POINTER(C.c_float))))
        for a, net in enumerate(("l_IAF_mu", "l_IAF_ls")):
            for b, (sub, key) in enumerate((("_input", "mask_input"), ("_output_W", "mask_output_W"), ("_output_D", "mask_output_D"))):
                keep = ref[key] != 0
                W = P[net + sub + ".W"]
                assert np.array_equal(got[a, b].view(np.uint32)[keep], W.view(np.uint32)[keep]), (net, sub)   # kept weights: same bits
                assert not got[a, b][~keep].any(), (net, sub)                                              # masked weights: exactly zero
    finally:
        m.close()


@pytest.mark.gpu
def test_device_resident_boxes_are_clamped_and_empty_boxes_give_nan(model, weights):
    """*_dev entry points cannot validate boxes on the host: the kernels clamp to the frame (numpy slice semantics) and an
    empty box yields NaN for that sample (mean of an empty slice) -- never an out-of-bounds read or a silent inf."""
    import torch
    rng = np.random.default_rng(17)
    z = rng.standard_normal((4, 100)).astype(np.float32)
    boxes = np.array([[8, 8, 24, 24], [50, 50, 80, 90], [5, 5, 5, 9], [-4, -3, 12, 10]], np.int32)
    clamped = np.array([[8, 8, 24, 24], [50, 50, 64, 64], [0, 0, 1, 1], [0, 0, 12, 10]], np.int32)
    rgb = rng.uniform(-1, 1, (4, 3)).astype(np.float32)
    want = model.grad(z, clamped, rgb)
    zt, bt, rt = torch.from_numpy(z).cuda(), torch.from_numpy(boxes).cuda(), torch.from_numpy(rgb).cuda()
    gt = torch.empty(4, 100, device="cuda")
    model.grad_dev(zt.data_ptr(), bt.data_ptr(), rt.data_ptr(), 0, 4, gt.data_ptr())
    torch.cuda.synchronize()
    g = gt.cpu().numpy()
    for k in (0, 1, 3):
        assert np.abs(g[k] - want[k]).max() <= 1e-5 * np.abs(want[k]).max(), k
    assert np.isnan(g[2]).all()
    with pytest.raises(npe_error(model)):
        model.grad(z, boxes, rgb)                                  # the host entry point rejects them outright


def npe_error(model):
    return importlib.import_module("neural-photo-editor_b200").IanError
