"""CPU tests of the drop-in boundary: the built C-ABI library loads and exports every symbol that
include/b200nn.h declares (no compute calls -- there is no GPU here), and it refuses to run
without a CUDA device instead of falling back to anything."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from cvt_b200 import build, capi
    build.build_lib()
    return capi.load()


def _declared():
    hdr = open(os.path.join(ROOT, "include", "b200nn.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200nn_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_all_exported(lib):
    from cvt_b200 import capi
    declared = _declared()
    assert len(declared) >= 40
    assert sorted(capi.SYMBOLS) == declared, "cvt_b200/capi.py SYMBOLS must list exactly what the header declares"
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/b200nn.h but not exported"


def test_library_is_sm100a_native_code():
    from cvt_b200 import capi
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(5|6|7|8|9)\d", out)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cvt_b200 import capi
    with pytest.raises(capi.B200nnError, match="no CUDA device"):
        capi.Context(0)
    assert b"sm_100a" in lib.b200nn_version()


def test_product_does_not_touch_oracle():
    # the oracle is test infrastructure: nothing under cvt_b200/, include/ or tools/ may reference it
    for base in ("cvt_b200", "include", "tools"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            if "/lib" in dp or "/bin" in dp:
                continue
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    assert "oracle" not in txt.lower() or f == "quick_scan_bench.py", f"{dp}/{f} mentions the oracle"
