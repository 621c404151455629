"""CPU: libmyzkp_b200.so loads and exports every symbol include/myzkp_b200.h declares
(no compute calls - there is no GPU here), and the product fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "myzkp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(myzkp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from myzkp_b200 import _lib

    lib = _lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/myzkp_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "python binding and header disagree"


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import myzkp_b200

    with pytest.raises(RuntimeError):
        myzkp_b200.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "myzkp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "myzkp_oracle" not in text and "oracle/" not in text.replace("the oracle", ""), f


def test_header_is_plain_c_and_mirrors_agree():
    """The boundary is a C ABI: the header must compile as C99, and the Rust `-sys` crate (shipped as source) must
    declare exactly the header's entry points."""
    import subprocess
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write('#include "myzkp_b200.h"\nint main(void) { return 0; }\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only",
                               "-I", os.path.join(ROOT, "include"), src])
    rs = open(os.path.join(ROOT, "rust", "myzkp-b200-sys", "src", "lib.rs")).read()
    rust = sorted(set(re.findall(r"pub fn (myzkp_[a-z0-9_]+)\s*\(", rs)))
    missing = [n for n in _declared() if n not in rust]
    extra = [n for n in rust if n not in _declared()]
    assert not missing and not extra, (missing, extra)
