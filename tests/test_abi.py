"""The C-ABI boundary (include/cora_b200.h <-> libcora_b200.so <-> cora_b200/_lib.py), no GPU needed:
every function the header declares is exported by the built library and bound by the ctypes
table with the same number of arguments; calls that need no device behave."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cora_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|long long|const char\*)\s+(cora_b200_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(1)] = n
    return decls


@pytest.fixture(scope="module")
def lib():
    from cora_b200 import _lib, build

    build.build()
    return _lib.load()


def test_header_declares_the_hot_path():
    d = _declared()
    for name in ("cora_b200_cl_fill_21cm", "cora_b200_cl_fill_sck", "cora_b200_root_batched", "cora_b200_draw_apply",
                 "cora_b200_alm2map", "cora_b200_alm2map_spin2", "cora_b200_draw_apply_peers", "cora_b200_cl_fill_21cm_tiles",
                 "cora_b200_peer_barrier"):
        assert name in d
    assert len(d) >= 35


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(os.path.join(ROOT, "cora_b200", "libcora_b200.so"))
    for name in _declared():
        assert hasattr(raw, name), "libcora_b200.so does not export %s" % name


def test_ctypes_table_matches_header(lib):
    from cora_b200 import _lib

    decl = _declared()
    assert set(decl) == set(_lib.SIGNATURES), set(decl) ^ set(_lib.SIGNATURES)
    for name, nargs in decl.items():
        assert len(_lib.SIGNATURES[name][1]) == nargs, name


def test_deviceless_calls(lib):
    assert lib.cora_b200_version() >= 100
    assert lib.cora_b200_timing_kinds() >= 9
    names = [lib.cora_b200_timing_name(i).decode() for i in range(lib.cora_b200_timing_kinds())]
    assert "sht_legendre" in names and "cl_fill" in names
    # argument validation happens before any CUDA call and reports through last_error
    rc = lib.cora_b200_root_batched(None, 0, 0, 0.0, 0.0, None, None, None, None, 0, None)
    assert rc != 0 and b"root_batched" in lib.cora_b200_last_error()
    assert lib.cora_b200_draw_apply_workspace_bytes(4, 7, 2) > 0
    assert lib.cora_b200_root_workspace_bytes(3, 8) > 0


def test_no_cpu_fallback_without_device():
    import torch

    from cora_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(_lib.CoraB200Error):
        _lib.require_cuda()
