"""CPU ORACLE (test infrastructure) for the training-mode pieces of SURVEY 8(f) rank 4: float64 restatements.

  * batch_norm_train: lasagne BatchNormLayer.get_output_for(deterministic=False) -- what every `BN(...)` /
    `lasagne.layers.batch_norm` of the reference graphs (IAN_simple.py:12,84-170; layers.py:411-416) evaluates in training.
    Third-party semantics, restated from Lasagne 0.2.dev1 `lasagne/layers/normalization.py`:
    statistics over all axes but the second, biased variance, epsilon inside the square root, running averages
    r <- (1 - alpha) r + alpha * batch value for `mean` AND `inv_std`.
  * minibatch_layer: reference layers.py:486-524 (MinibatchLayer.get_output_for, init=False); pinned by executing the
    reference class itself (tests/golden/make_golden_train.py -> tests/golden/ref_exec_train.npz).
Only tests/ may import this module.
"""
import numpy as np

F64 = np.float64


def batch_norm_train(x, gamma, beta, running_mean, running_inv_std, eps=1e-4, alpha=0.1):
    """x (n, c, ...) -> (y, new_running_mean, new_running_inv_std, batch_mean, batch_inv_std)."""
    x = np.asarray(x, F64)
    axes = (0,) + tuple(range(2, x.ndim))
    mean = x.mean(axes)
    inv_std = 1.0 / np.sqrt(x.var(axes) + eps)
    shape = [1, -1] + [1] * (x.ndim - 2)
    y = (x - mean.reshape(shape)) * (np.asarray(gamma, F64) * inv_std).reshape(shape) + np.asarray(beta, F64).reshape(shape)
    return (y, (1 - alpha) * np.asarray(running_mean, F64) + alpha * mean,
            (1 - alpha) * np.asarray(running_inv_std, F64) + alpha * inv_std, mean, inv_std)


def minibatch_layer(x, theta, log_weight_scale, b):
    """layers.py:495 (W), :503-524 (forward, init=False): x (n, d) [flattened if needed] -> (n, d + K)."""
    x = np.asarray(x, F64).reshape(len(x), -1)                                                  # :504-507 flatten(2)
    theta, lws, b = np.asarray(theta, F64), np.asarray(log_weight_scale, F64), np.asarray(b, F64)
    W = theta * (np.exp(lws) / np.sqrt(np.sum(np.square(theta), axis=0)))[None]                 # :495
    act = np.tensordot(x, W, [[1], [0]])                                                        # :509  (n, K, P)
    abs_dif = (np.sum(np.abs(act[:, :, :, None] - act.transpose(1, 2, 0)[None]), axis=2)
               + 1e6 * np.eye(len(x))[:, None, :])                                              # :510-511  (n, K, n)
    f = np.sum(np.exp(-abs_dif), axis=2) + b[None]                                              # :518, :524
    return np.concatenate([x, f], axis=1)                                                       # :526
