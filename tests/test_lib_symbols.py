"""The C-ABI library loads and exports every symbol include/kurosiwo_b200.h declares (no compute calls)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_exported():
    from kurosiwo_b200 import build, lib
    build.build()
    header = (ROOT / "include" / "kurosiwo_b200.h").read_text()
    declared = set(re.findall(r"\b(ks_[a-z0-9_]+)\s*\(", header))
    declared -= {"ks_view"}
    assert declared, "no declarations parsed"
    handle = ctypes.CDLL(str(build.LIB))
    missing = [s for s in sorted(declared) if not hasattr(handle, s)]
    assert not missing, missing
    assert set(lib.EXPORTED_SYMBOLS) == declared
    assert handle.ks_version() == 100
    handle.ks_error_string.restype = ctypes.c_char_p
    assert handle.ks_error_string(-2).decode().startswith("unsupported")


def test_argument_errors_without_gpu():
    from kurosiwo_b200 import lib
    h = lib.load()
    assert h.ks_set_option(b"no_such_option", 1) == -1
    assert h.ks_set_option(b"tc_mt", 0) == 0
    # null pointers are rejected before any launch
    assert h.ks_bn_finalize(0, ctypes.c_double(1.0), None, None, None, ctypes.c_float(1e-5), ctypes.c_float(0.1), None, None, None, None, None, None, None) == -1
