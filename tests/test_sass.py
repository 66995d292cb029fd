"""The shipped library IS the tcgen05 / TMA / PDL code the design describes: SASS of the in-tree libian_b200.so, read with
cuobjdump (no GPU needed).  Guards against a silent rebuild onto another code path (a recompiled mma.sync kernel, a plain
store epilogue, plain launches) and keeps profiles/r2_sass_summary.txt -- the evidence file the docs cite -- in step with
the build that is tested."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "neural-photo-editor_b200", "libian_b200.so")


def _summary():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    if not os.path.exists(LIB):
        pytest.skip("library not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_summary.py")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    rows = {}
    for line in out.stdout.splitlines():
        if line.startswith("#"):
            continue
        f = [c.strip() for c in line.split("|")]
        rows[f[0]] = dict(kv.rsplit(" x", 1) for kv in f[4].split(", ") if " x" in kv)
    return out.stdout, rows


def test_tensor_core_kernels_are_tcgen05_tma_pdl():
    text, rows = _summary()
    pair = [k for k in rows if k.startswith("tapgemm_tc2_kernel<")]
    one = [k for k in rows if k.startswith("tapgemm_tc_kernel<")]
    assert len(pair) >= 4 and len(one) >= 4, sorted(rows)
    for k in pair:                                           # CTA-pair tap-GEMM: cta_group::2 MMA, TMA loads, multicast commits
        r = rows[k]
        assert "UTCHMMA.2CTA" in r and "UTCBAR.2CTA.MULTICAST" in r and "LDTM" in r, (k, r)
        assert any(m.startswith("UTMALDG") and m.endswith(".2CTA") for m in r), (k, r)
        assert "STG.E.ENL2.256" in r, (k, r)                  # 32-byte epilogue stores
        assert "PREEXIT" in r and "ACQBULK" in r, (k, r)      # griddepcontrol.launch_dependents / .wait
    for k in one:
        r = rows[k]
        assert "UTCHMMA" in r and "LDTM" in r and any(m.startswith("UTMALDG") for m in r), (k, r)
        assert "PREEXIT" in r and "ACQBULK" in r, (k, r)
    c1 = rows["conv1_tc_kernel"]
    assert "UTCHMMA" in c1 and "UTMASTG" in c1 and "STG.E.ENL2.256" not in c1, c1    # epilogue leaves through TMA stores
    for k in ("decout_tc_kernel", "head_tc_kernel<1>", "head_tc_kernel<3>"):
        assert "UTCHMMA" in rows[k] and any(m.startswith("UTMALDG") for m in rows[k]), (k, rows[k])
    assert not any("HMMA" in m and not m.startswith("UTC") for r in rows.values() for m in r), "an mma.sync kernel is in the library"
    for k in ("splitk_finalize_kernel", "splitk_finalize8_kernel", "brush_seed_bwd_kernel", "brush_update_kernel", "sample_kernel"):
        assert "PREEXIT" in rows[k] and "ACQBULK" in rows[k], (k, rows[k])


def test_committed_sass_summary_matches_the_build():
    text, _ = _summary()
    path = os.path.join(ROOT, "profiles", "r2_sass_summary.txt")
    committed = open(path).read()

    def key(t):                                              # kernel | Blackwell mnemonics (registers may differ across toolkits)
        out = []
        for line in t.splitlines():
            if line.startswith("#"):
                continue
            f = [c.strip() for c in line.split("|")]
            out.append((f[0], f[4]))
        return out
    assert key(text) == key(committed), "profiles/r2_sass_summary.txt is stale: run `python tools/sass_summary.py > profiles/r2_sass_summary.txt`"
