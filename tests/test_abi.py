"""CPU tests of the boundary: the C-ABI library builds/loads, exports every symbol include/ian_b200.h
declares, fails loudly without a GPU, and the Python mirror of API.IAN filters inputs like the
reference's theano functions (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ian_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ian_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree(npe):
    assert _declared_symbols() == sorted(npe.SIGNATURES)


def test_library_exports_every_declared_symbol(npe):
    import importlib
    build = importlib.import_module("neural-photo-editor_b200.build")
    build.build()                                   # no-op when up to date; nvcc cross-compiles on CPU
    lib = npe.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), name


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_create_fails_loudly_without_gpu(npe, weights):
    lib = npe.load()
    h = C.c_void_p()
    rc = lib.ian_create(0, 0, C.byref(h))
    assert rc < 0 and b"no CPU path" in lib.ian_last_error(None)
    with pytest.raises(npe.IanError):
        npe.IAN("IAN_simple.py", True, weights=weights)


def test_input_filtering_matches_theano_rules(npe):
    api = __import__("importlib").import_module("neural-photo-editor_b200.API")
    assert api._int_scalar(3.0, "c1") == 3 and api._int_scalar(np.float64(7.0), "c1") == 7
    with pytest.raises(TypeError):
        api._int_scalar(3.5, "c1")
    with pytest.raises(TypeError):
        api._f32(np.zeros((1, 100)), 2, "z")            # float64 is rejected like theano does
    with pytest.raises(TypeError):
        api._f32(np.zeros((100,), np.float32), 2, "z")
    assert api._f32(np.zeros((2, 100), np.float32), 2, "z").flags["C_CONTIGUOUS"]


def test_unknown_config_is_rejected(npe, weights):
    with pytest.raises(NotImplementedError):
        npe.IAN("IAN_v2.py", True, weights=weights)


def test_made_ordering_host_logic_matches_oracle(npe):
    """bit-exact MADE mask indexing starts from the same integer ordering on both sides."""
    api = __import__("importlib").import_module("neural-photo-editor_b200.API")
    from oracle import ian_full_numpy as fn
    o = api.made_ordering()
    assert o.dtype == np.int32 and sorted(o.tolist()) == list(range(100))
    assert np.array_equal(o, fn.made_ordering().astype(np.int32))
    assert o[:12].tolist() == [52, 79, 87, 45, 24, 71, 82, 80, 34, 36, 89, 77] and o[80] == 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "neural-photo-editor_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def _c_host():
    import importlib
    return importlib.import_module("neural-photo-editor_b200.build").build_c_host()


def test_c_host_program_links_every_entry_point():
    """examples/c_host/ian_cli.c is strict C99 (-Wall -Wextra -Werror -pedantic) against include/ian_b200.h: the header
    is a C header, and taking every entry point by address makes a missing export a link error."""
    import subprocess
    exe = _c_host()
    out = subprocess.run([exe, "--symbols"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split()[0] == str(len(_declared_symbols()))
    src = open(os.path.join(ROOT, "examples", "c_host", "ian_cli.c")).read()
    for name in _declared_symbols():
        assert "(anyfn)%s," % name in src or "(anyfn)%s}" % name in src, name


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_c_host_program_fails_loudly_without_gpu(tmp_path):
    import subprocess
    out = subprocess.run([_c_host(), str(tmp_path / "w.bin"), str(tmp_path / "x.f32"), "1", str(tmp_path / "o.f32")],
                         capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "no CPU path" in out.stderr
