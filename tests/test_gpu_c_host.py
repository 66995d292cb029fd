"""The C-ABI from a plain C host program (examples/c_host/ian_cli.c): no Python and no torch in the process that
computes.  The test only writes the inputs, runs the binary and checks its outputs against the float64 oracle and,
bit for bit, against the same library driven through the Python mirror."""
import importlib
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import ian_numpy as on

pytestmark = pytest.mark.gpu


def _write_weights(path, weights):
    items = [(k, np.ascontiguousarray(v, dtype=np.float32)) for k, v in weights.items()
             if k != "metadata" and not k.startswith(("minibatch_discrim.", "discrimi."))]   # like API.IAN (API.py:41-47)
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(items)))
        for name, arr in items:
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)) + nb + struct.pack("<I", arr.ndim))
            f.write(struct.pack("<%dq" % arr.ndim, *arr.shape))
            f.write(arr.tobytes())


@pytest.mark.parametrize("n", [1, 9])
def test_c_host_reconstruct(model, weights, golden, tmp_path, n):
    exe = importlib.import_module("neural-photo-editor_b200.build").build_c_host()
    x = on.to_tanh(np.resize(golden["images"].astype(np.float64), (n, 3, 64, 64))).astype(np.float32)
    x[1:] += np.random.default_rng(n).uniform(-0.05, 0.05, x[1:].shape).astype(np.float32)
    x = np.clip(x, -1, 1)
    _write_weights(tmp_path / "w.bin", weights)
    x.tofile(tmp_path / "x.f32")
    out = subprocess.run([exe, str(tmp_path / "w.bin"), str(tmp_path / "x.f32"), str(n), str(tmp_path / "xh.f32"),
                          str(tmp_path / "z.f32")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "reconstructed %d image(s), zdim 100" % n in out.stdout
    xh = np.fromfile(tmp_path / "xh.f32", np.float32).reshape(n, 3, 64, 64)
    z = np.fromfile(tmp_path / "z.f32", np.float32).reshape(n, 100)
    zr = on.simple_encode(weights, x[:2].astype(np.float64))
    assert np.abs(z[:2] - zr).max() <= 2e-4
    assert np.abs(xh[:2] - on.simple_decode(weights, z[:2])).max() <= 1e-4
    xp, zp = model.reconstruct(x, return_z=True)           # same library, same batch size: bit-identical
    assert np.array_equal(xh, xp) and np.array_equal(z, zp)
