"""The C-ABI library loads and exports every symbol include/plsb200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "plsb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(plsb200_\w+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    import __graft_entry__
    __graft_entry__.build()
    from plspy_b200 import _lib
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in plsb200.h but not exported by libplsb200.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in plspy_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)
    assert _lib.lib.plsb200_abi_version() == 3


def test_argument_errors_are_reported_without_a_gpu():
    """Size queries and argument checks run on the host."""
    from plspy_b200 import _lib
    lib = _lib.lib
    assert lib.plsb200_gram_f64_workspace(300, 200000) > 0
    assert lib.plsb200_boot_coef_bytes(300, 12, 5000) == 2500 * 76 * 3 * 32 * 8
    assert lib.plsb200_boot_coef_bytes(300, 40, 10) == 0          # K > 24 must be split by the caller
    rc = lib.plsb200_gram_f64(None, 10, 10, 10, None, None, 0, None)
    assert rc == -1 and b"null pointer" in lib.plsb200_last_error()
