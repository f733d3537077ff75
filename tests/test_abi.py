"""The C-ABI library loads on a CPU-only host and exports every symbol that
include/ssd_b200.h declares (no compute calls here); argument errors are
reported through the documented negative codes."""

import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ssd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ssd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from tf_ssd_b200 import _ffi
    lib = _ffi.lib()
    names = _declared()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/ssd_b200.h but not exported"
    assert lib.ssd_abi_version() == 1


def test_every_export_is_declared_and_bound():
    """No undeclared ssd_* export, and the ctypes table covers the whole header."""
    import subprocess
    from tf_ssd_b200 import _ffi, _ffi_conv
    out = subprocess.run(["nm", "-D", "--defined-only", _ffi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r"\bT (ssd_[a-z0-9_]+)$", out, flags=re.M)))
    assert exported == _declared()
    lib = _ffi.lib()
    for name in exported:
        assert getattr(lib, name).argtypes is not None or name in ("ssd_abi_version", "ssd_last_error"), name
    assert set(_ffi_conv.SIGNATURES) <= set(exported)


def test_argument_errors_without_gpu():
    from tf_ssd_b200 import _ffi
    lib = _ffi.lib()
    assert lib.ssd_iou_map(None, None, 1, 1, 1, 0, None, None) == -1            # SSD_ERR_NULL
    assert b"NULL" in lib.ssd_last_error()
    fm = (C.c_int * 2)(3, 1)
    cnt = (C.c_int * 2)(3, 3)
    assert lib.ssd_prior_box_count(fm, 2, cnt) == 3 * 3 * 4 + 4
    assert lib.ssd_loss_workspace_bytes(2, 100, 21) >= 2 * 100 * 14
    assert lib.ssd_decode_nms_workspace_bytes(2, 100, 21, 200, 0) > 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tf_ssd_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, fn)
