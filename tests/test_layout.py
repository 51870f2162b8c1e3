"""Repo-layout contract (CPU): the C-ABI library exports every symbol include/avb200.h declares and
the ctypes table types each of them; the product package never touches oracle/ or a CPU fallback."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "avb200.h")).read()
    return sorted(set(re.findall(r"AVB_API\s+[\w\s\*]+?\b(avb_\w+)\s*\(", src)))


def test_header_symbols_exported_and_typed():
    from animal_vision_b200 import _abi
    names = _declared()
    assert len(names) >= 10
    lib = _abi.load()                      # builds with nvcc if needed; no GPU required to load
    for n in names:
        assert hasattr(lib, n), f"libavb200.so does not export {n}"
        assert n in _abi.SIGNATURES, f"_abi.SIGNATURES does not type {n}"
    assert sorted(_abi.SIGNATURES) == names, "ctypes table and header disagree"
    assert lib.avb_version() == _abi.AVB_VERSION
    raw = ctypes.CDLL(_abi.lib_path())
    assert raw.avb_version() == _abi.AVB_VERSION


def test_product_never_imports_oracle_or_reference():
    pkg = os.path.join(ROOT, "animal_vision_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M) or "/root/reference" in txt:
                    bad.append(os.path.join(d, f))
    assert not bad, f"product code must not import the oracle or read the reference: {bad}"


def test_compute_fails_loudly_without_gpu():
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from animal_vision_b200._abi import AvbError
    from animal_vision_b200.animals import Dog
    with pytest.raises(AvbError):
        Dog().visualize(np.zeros((4, 4, 3), np.uint8))


def test_documented_runtime_switches_exist_in_the_sources():
    """Every AVB_* environment switch INTEGRATION.md documents is read somewhere in csrc/ (and vice versa for getenv calls)."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    section = doc[doc.index("## Run-time switches"):]
    documented = set(re.findall(r"`(AVB_[A-Z0-9_]+)(?:=[^`]*)?`", section))
    src = "".join(open(f).read() for f in glob.glob(os.path.join(root, "animal_vision_b200", "csrc", "*.cu")))
    read = set(re.findall(r'getenv\("(AVB_[A-Z0-9_]+)"\)', src))
    assert documented, "no switches parsed from INTEGRATION.md"
    assert documented <= read, f"documented but never read: {sorted(documented - read)}"
    assert read <= documented, f"read but undocumented: {sorted(read - documented)}"
