"""Training-mode pieces next to the IAN hot path, through the C-ABI of libian_b200.so (include/ian_b200.h):

  * batch_norm_train  = lasagne BatchNormLayer.get_output_for(deterministic=False), i.e. `BN(...)` of the reference graphs
    (IAN_simple.py:12,84-170; layers.py:411-416) in training, with the running `mean` / `inv_std` update; statistics
    can be synchronised over a torch.distributed group (sum / sum-of-squares all-reduce between the two library calls).
  * minibatch_layer   = reference layers.py:486-524 (MinibatchLayer forward).

Inputs and outputs are torch CUDA tensors (float32); torch only carries the device memory and the optional all-reduce.
The trainers themselves (train_IAN*.py) are out of scope and stay the reference's.
"""
from __future__ import annotations

import contextlib


@contextlib.contextmanager
def _lib_stream(model, x):
    """The raw CUDA stream the library call is enqueued on, ordered with torch's current stream.  A non-default torch
    stream is used as is; the legacy default stream (handle 0, which the C-ABI reads as "the handle's own stream") is
    bridged through a side stream with wait_stream on both sides."""
    import torch
    cur = torch.cuda.current_stream(x.device)
    if cur.cuda_stream != 0:
        yield cur.cuda_stream
        return
    side = getattr(model, "_train_stream", None)
    if side is None:
        side = model._train_stream = torch.cuda.Stream(device=x.device)
    side.wait_stream(cur)
    yield side.cuda_stream
    cur.wait_stream(side)


def batch_norm_train(model, x, gamma, beta, running_mean=None, running_inv_std=None, eps=1e-4, alpha=0.1, group=None):
    """x (n, c, ...) float32 CUDA, contiguous.  Returns y; running_mean / running_inv_std are updated IN PLACE.
    `model`: any finalized IAN handle on x's device (it owns the workspace).  `group`: torch.distributed group (or True
    for the default group) to synchronise the batch statistics across ranks."""
    import torch
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() >= 2
    n, c = int(x.shape[0]), int(x.shape[1])
    hw = int(x[0, 0].numel()) if x.dim() > 2 else 1
    sums = torch.empty(2, c, dtype=torch.float64, device=x.device)
    with _lib_stream(model, x) as st:
        model._check(model._lib.ian_bn_batch_stats_dev(model._h, x.data_ptr(), n, c, hw, sums[0].data_ptr(), sums[1].data_ptr(), st))
    count = float(n * hw)
    if group is not None:                                   # cross-GPU synchronised BN: one all-reduce of 2*c float64
        import torch.distributed as dist
        g = None if group is True else group
        dist.all_reduce(sums, group=g)
        count *= dist.get_world_size(g)
    y = torch.empty_like(x)
    ptr = lambda t: t.data_ptr() if t is not None else None
    with _lib_stream(model, x) as st:
        model._check(model._lib.ian_bn_train_normalize_dev(model._h, x.data_ptr(), n, c, hw, sums[0].data_ptr(), sums[1].data_ptr(),
                                                           count, ptr(gamma), ptr(beta), float(eps), float(alpha),
                                                           ptr(running_mean), ptr(running_inv_std), y.data_ptr(), st))
    return y


def minibatch_layer(model, x, theta, log_weight_scale, b):
    """x (n, ...) float32 CUDA (flattened to (n, d) like layers.py:504-507); theta (d, K, P); log_weight_scale (K, P); b (K).
    Returns (n, d + K) = [x | f]."""
    import torch
    x2 = x.reshape(x.shape[0], -1).contiguous()
    n, d = int(x2.shape[0]), int(x2.shape[1])
    K, P = int(theta.shape[1]), int(theta.shape[2])
    assert tuple(theta.shape) == (d, K, P) and tuple(log_weight_scale.shape) == (K, P) and tuple(b.shape) == (K,)
    out = torch.empty(n, d + K, dtype=torch.float32, device=x.device)
    th, lw, bb = theta.contiguous(), log_weight_scale.contiguous(), b.contiguous()
    with _lib_stream(model, x2) as st:
        model._check(model._lib.ian_minibatch_discrim_dev(model._h, x2.data_ptr(), n, d, th.data_ptr(), lw.data_ptr(), bb.data_ptr(),
                                                          K, P, out.data_ptr(), st))
    return out
