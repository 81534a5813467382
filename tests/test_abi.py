"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol include/svgpu.h declares; the product package never touches the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fhgr-stereo-vision_b200")


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "svgpu.h")).read()
    return re.findall(r"SVGPU_API\s+[\w\s\*]+?\b(svgpu_\w+)\s*\(", txt)


def test_header_and_binding_agree():
    import svgpu
    syms = header_symbols()
    assert len(syms) >= 25
    assert sorted(syms) == sorted(svgpu._abi.EXPORTS)


def test_library_exports_every_declared_symbol():
    import svgpu
    lib = svgpu._abi.load()
    for s in header_symbols():
        assert hasattr(lib, s), s
    assert b"sm_100a" in lib.svgpu_version()


def test_missing_library_fails_loudly(monkeypatch):
    import svgpu
    monkeypatch.setattr(svgpu._abi, "_lib", None)
    monkeypatch.setattr(svgpu._abi, "LIB_PATH", "/nonexistent/libsvgpu.so")
    with pytest.raises(ImportError):
        svgpu._abi.load()


def test_no_gpu_means_error_not_fallback():
    import torch
    import svgpu
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(svgpu.SvgpuError):
        svgpu.Context(64, 48)


def test_product_never_imports_the_oracle():
    pat = re.compile(r"oracle|cv2")
    for dp, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                for n, line in enumerate(open(os.path.join(dp, f), errors="ignore"), 1):
                    if re.search(r"^\s*(import|from|#include)\b.*(oracle|cv2)", line):
                        raise AssertionError(f"{f}:{n}: {line.strip()}")
    assert pat  # silence linters
