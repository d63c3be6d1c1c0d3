"""CPU tests of the C-ABI boundary: the library builds/loads without a GPU, exports every symbol include/vitb200.h
declares, the ctypes table matches the header, and compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "vitb200.h")).read()
    return sorted(set(re.findall(r"VB_API\s+[\w\s\*]+?\b(vb_\w+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    from vitb200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vitb200.h but not exported"
    assert sorted(_lib.SIGNATURES.keys()) == syms, "ctypes signature table and header disagree"
    assert lib.vb_version() == 1


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of the POD descriptors as gcc sees them in the header == the ctypes mirrors."""
    import subprocess
    from vitb200 import _lib
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "vitb200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(VbGemmDesc), offsetof(VbGemmDesc, bias), '
        'offsetof(VbGemmDesc, debug_direct_store), sizeof(VbAttnDesc), offsetof(VbAttnDesc, key_padding_mask), '
        'offsetof(VbAttnDesc, lddv));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    vals = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    G, A = _lib.VbGemmDesc, _lib.VbAttnDesc
    assert vals == [ctypes.sizeof(G), G.bias.offset, G.debug_direct_store.offset, ctypes.sizeof(A), A.key_padding_mask.offset,
                    A.lddv.offset]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less container")
    from vitb200 import _lib
    lib = _lib.load()
    assert lib.vb_device_check(0) != 0
    assert b"CUDA" in lib.vb_last_error() or b"device" in lib.vb_last_error()
    d = _lib.VbGemmDesc()
    assert lib.vb_gemm_bf16(ctypes.byref(d), None) != 0      # no device => error code, never a silent CPU path


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "vitb200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src, f"vitb200/{fn} references the oracle"
