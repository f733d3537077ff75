"""ctypes binding of ``libssd_b200.so`` (C ABI declared in ``include/ssd_b200.h``).

PyTorch is used for device memory and streams only: every compute call goes
through the C ABI with raw device pointers (``tensor.data_ptr()``) and the
current CUDA stream handle.  There is NO CPU fallback: a missing library or a
missing CUDA device raises immediately.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssd_b200.so")

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
vp = C.c_void_p


class SsdB200Error(RuntimeError):
    """Raised for any non-zero status returned by the C ABI."""


def _declare(lib: C.CDLL) -> None:
    i, f, sz, i64 = C.c_int, C.c_float, C.c_size_t, C.c_int64
    sig = {
        "ssd_abi_version": (i, []),
        "ssd_last_error": (C.c_char_p, []),
        "ssd_device_info": (i, [C.c_char_p, i, c_int_p, c_int_p]),
        "ssd_prior_boxes": (i, [c_int_p, i, c_float_p, c_int_p, vp, i, vp]),
        "ssd_prior_box_count": (i, [c_int_p, i, c_int_p]),
        "ssd_iou_map": (i, [vp, vp, i, i, i, i, vp, vp]),
        "ssd_match_encode": (i, [vp, vp, vp, i, i, i, i, f, c_float_p, vp, vp, vp, vp, vp]),
        "ssd_encode_deltas": (i, [vp, vp, i, i, i, vp, vp]),
        "ssd_decode_boxes": (i, [vp, vp, i, i, i, vp, vp]),
        "ssd_loss_workspace_bytes": (sz, [i, i, i]),
        "ssd_loss_fwd": (i, [vp, vp, vp, vp, i, i, i, f, f, i, vp, vp, vp, sz, vp]),
        "ssd_loss_bwd": (i, [vp, vp, vp, vp, i, i, i, f, f, vp, vp, vp, sz, vp]),
        "ssd_softmax": (i, [vp, i64, i, vp, vp]),
        "ssd_decode_nms_workspace_bytes": (sz, [i, i, i, i, i]),
        "ssd_decode_nms": (i, [vp, vp, vp, i, i, i, c_float_p, i, f, f, i, i, vp, vp, vp, vp, vp, sz, vp]),
        "ssd_combined_nms_workspace_bytes": (sz, [i, i, i, i, i, i]),
        "ssd_combined_nms": (i, [vp, vp, i, i, i, i, i, i, f, f, i, i, vp, vp, vp, vp, vp, sz, vp]),
    }
    try:
        from . import _ffi_conv  # conv / network entry points (declared in the same header)
        sig.update(_ffi_conv.SIGNATURES)
    except ImportError:
        pass
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SsdB200Error(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C tf_ssd_b200/csrc`. There is no CPU fallback."
            )
        handle = C.CDLL(LIB_PATH)
        _declare(handle)
        if handle.ssd_abi_version() != 1:
            raise SsdB200Error("libssd_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().ssd_last_error().decode("utf-8", "replace")
        raise SsdB200Error(f"{what or 'libssd_b200'} failed with status {status}: {msg}")


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise SsdB200Error("tf_ssd_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


_checked_devices = set()


def check_device() -> None:
    """Fail loudly when the current device is not an sm_100 part."""
    dev = require_cuda()
    if dev.index in _checked_devices:
        return
    name = C.create_string_buffer(128)
    sms, cc = C.c_int(0), C.c_int(0)
    check(lib().ssd_device_info(name, 128, C.byref(sms), C.byref(cc)), "ssd_device_info")
    if cc.value // 10 != 10:
        raise SsdB200Error(f"device {name.value.decode()} is sm_{cc.value}; libssd_b200 is built for sm_100a only")
    _checked_devices.add(dev.index)


def stream() -> vp:
    return vp(torch.cuda.current_stream().cuda_stream)


def to_dev(x, dtype=torch.float32) -> torch.Tensor:
    """NumPy array / CPU tensor / CUDA tensor -> contiguous CUDA tensor of ``dtype``."""
    dev = require_cuda()
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    if t.device.type != "cuda":
        t = t.to(dev, non_blocking=False)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def ptr(t: Optional[torch.Tensor]) -> vp:
    return vp(0) if t is None else vp(t.data_ptr())


def f32_array(values: Sequence[float]):
    arr = (C.c_float * len(values))(*[float(v) for v in values])
    return arr


def i32_array(values: Sequence[int]):
    arr = (C.c_int * len(values))(*[int(v) for v in values])
    return arr


def workspace(nbytes: int) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=require_cuda())
