"""CPU: the C-ABI library loads and exports every symbol include/b200svd.h declares; the ctypes prototype table
covers exactly those symbols.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200svd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200svd_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header():
    import __graft_entry__ as g
    g.build()
    from streamingt2v_b200 import _lib
    lib = ctypes.CDLL(str(_lib.lib_path()))
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in b200svd.h but not exported"
    table = set(_lib.PROTOTYPES) | {"b200svd_last_error", "b200svd_version", "b200svd_init",
                                    "b200svd_gn_scratch_doubles", "b200svd_gemm_pair_mode", "b200svd_flash_attn_variant"}
    assert table == set(names), (sorted(table - set(names)), sorted(set(names) - table))
    assert lib.b200svd_version() >= 100


def test_gemm_params_struct_layout():
    """The ctypes mirror of b200svd_gemm_params must have the C layout (checked against a compiled sizeof)."""
    import subprocess
    import tempfile
    from streamingt2v_b200._lib import GemmParams
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write('#include <stdio.h>\n#include <stddef.h>\n#include "b200svd.h"\nint main(){printf("%zu %zu %zu %zu",'
                             'sizeof(b200svd_gemm_params), offsetof(b200svd_gemm_params, tap_off),'
                             'offsetof(b200svd_gemm_params, out), offsetof(b200svd_gemm_params, bn));return 0;}')
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        size, o_tap, o_out, o_bn = (int(v) for v in subprocess.check_output([exe]).split())
    assert ctypes.sizeof(GemmParams) == size
    assert GemmParams.tap_off.offset == o_tap and GemmParams.out.offset == o_out and GemmParams.bn.offset == o_bn


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from streamingt2v_b200 import _lib
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setenv("B200SVD_LIB", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(_lib.B200Error, match="no fallback"):
        _lib.load()
