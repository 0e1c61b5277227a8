"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/genfer_taylor.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from genfer_b200 import build as B
    return B.build()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "genfer_taylor.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gt[piu]_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_surface():
    syms = _declared_symbols()
    for needed in ("gtp_mul", "gtp_div", "gtp_exp", "gtp_log", "gtp_pow", "gtp_subst_var", "gtp_shift_down",
                   "gtp_derivative", "gtp_taylor_expansion_of_coeff", "gtp_coefficients_of_term", "gtp_gather_axis",
                   "gtp_mul_rows_raw", "gtu_mul", "gtu_div", "gtu_exp", "gtu_log"):
        assert needed in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_table_matches_header(built_lib):
    import genfer_b200
    assert sorted(genfer_b200.SYMBOLS) == _declared_symbols()
    genfer_b200.load()  # sets restype/argtypes for every symbol; raises if one is absent


def test_rust_shim_declares_every_symbol():
    """rust/genfer-taylor-sys/src/lib.rs is generated from the header (tools/gen_rust_externs.py): no entry point missing,
    and its build script compiles every translation unit the in-tree build compiles."""
    import subprocess
    import sys
    lib_rs = open(os.path.join(ROOT, "rust", "genfer-taylor-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (gt[piu]_[a-z0-9_]+)\(", lib_rs))
    assert sorted(declared) == _declared_symbols()
    assert subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_externs.py"), "--check"]).returncode == 0
    from genfer_b200 import build as B
    csrc = os.path.join(ROOT, "genfer_b200", "csrc")
    on_disk = sorted(f for f in os.listdir(csrc) if f.endswith((".cu", ".cpp")))
    assert on_disk == sorted(B.SOURCES), "build.py SOURCES and csrc/ (globbed by build.rs) disagree"


def test_mac_count_needs_no_gpu(built_lib):
    """gtp_mul_macs is pure integer work: (D(D+1)/2)^n for dense cubes."""
    import genfer_b200
    for n, d in ((4, 32), (5, 16), (6, 12), (6, 16)):
        assert genfer_b200.mul_macs([d] * n, [d] * n, [d] * n) == float((d * (d + 1) // 2) ** n)
    from oracle import oracle as O
    assert genfer_b200.mul_macs([3, 5], [4, 2], [6, 6]) == O.mul_macs([3, 5], [4, 2], [6, 6])


def test_no_cpu_fallback_without_device(built_lib):
    """Without a CUDA device the product path must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import genfer_b200
    with pytest.raises(genfer_b200.TaylorError):
        genfer_b200.Context(0)
