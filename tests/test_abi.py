"""The C-ABI library loads here (no GPU) and exports every symbol include/tops_b200.h declares; without a device the
product refuses to run instead of falling back to a CPU path."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tops_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tops_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    import tensor_ops_b200._lib as L
    names = _declared()
    assert len(names) >= 50
    out = subprocess.check_output(["nm", "-D", "--defined-only", L.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (tops_[a-z0-9_]+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    # and every declared symbol has a ctypes prototype
    assert sorted(L.PROTOTYPES) == names


def test_library_contains_blackwell_sass():
    import tensor_ops_b200._lib as L
    sass = subprocess.run(["cuobjdump", "-sass", L.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR"):
        assert mnemonic in sass, f"{mnemonic} missing: the tcgen05/TMA path was not compiled in"
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", L.LIB_PATH], capture_output=True, text=True).stdout


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import tensor_ops_b200 as tb
    with pytest.raises(tb.TopsError) as ei:
        tb.Context(0)
    assert ei.value.code == 6   # TOPS_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tensor_ops_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read(), f
