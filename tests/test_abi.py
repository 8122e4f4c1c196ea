"""The C-ABI library loads on a CPU-only box and exports every symbol include/voroffset_b200.h declares.
No compute calls are made here."""
import ctypes
import os
import re

import pytest

from voroffset_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "voroffset_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vo_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == names


def test_version_and_span_size():
    lib = _lib.load()
    assert b"sm_100a" in lib.vo_version()
    assert lib.vo_span_bytes() == 16


def test_null_arguments_are_rejected_without_a_device():
    lib = _lib.load()
    assert lib.vo_create(0, None) == 1                       # VO_ERR_ARG
    assert lib.vo_last_error(None) == b"no context"
    lib.vo_free(None)                                         # no-op
    assert lib.vo_launch_count(None) == 0


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.VoroffsetError) as e:
        _lib.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "voroffset_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
                assert "libvoroffset_ref" not in src, f


def test_every_option_of_the_library_is_listed_in_the_header():
    """vo_set_option takes free-form keys: the header is the only place a caller of the C ABI can learn them from."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "voroffset_b200", "csrc", "vo_lib.cu")).read()
    hdr = open(os.path.join(root, "include", "voroffset_b200.h")).read()
    keys = sorted(set(re.findall(r'strcmp\(key, "([a-z0-9_]+)"\)', src)))
    assert len(keys) > 20
    missing = [k for k in keys if f'"{k}"' not in hdr]
    assert not missing, missing
