"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and
exports every symbol include/abm_b200.h declares.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "abm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(abm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built_lib):
    names = _declared_symbols()
    assert "abm_vf_step" in names and "abm_vf_projection_field" in names
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in include/abm_b200.h but not exported"


def test_python_binding_covers_header(built_lib):
    from abm_b200 import _lib
    assert sorted(_lib.SYMBOLS) == _declared_symbols()


def test_struct_sizes_match_header_layout():
    from abm_b200 import _lib
    assert C.sizeof(_lib.VFConfig) == 19 * 4
    assert C.sizeof(_lib.VFProjArgs) == 8 + 6 * 8 + 8 + 3 * 8 + 8 + 3 * 8
    assert C.sizeof(_lib.CSProjArgs) == 8 + 6 * 8 + 8 + 2 * 8 + 8


def test_version_and_words(built_lib):
    assert built_lib.abm_version() == 100
    assert built_lib.abm_field_words(1200) == 38
    assert built_lib.abm_field_words(2400) == 75
    assert built_lib.abm_field_words(8) == 1


def test_no_cpu_fallback_without_device(built_lib):
    """Without a usable device the engine must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from abm_b200 import AbmError, VFEngine
    with pytest.raises(AbmError) as ei:
        VFEngine(1, 4, resolution=64)
    assert ei.value.code in (-2, -3)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under abm_b200/ may reference it."""
    pkg = os.path.join(ROOT, "abm_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "ref_shim" not in src, f
