"""ctypes binding of the C ABI in include/phaze_b200.h (libphaze_b200.so).

The shared object holds the hand-written sm_100a kernels.  There is no Python or CPU
implementation behind it: if the library is missing, import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PVB_LIBRARY: another build of the same library (kernel experiments: profiles/*.sh)
LIB_PATH = os.environ.get("PVB_LIBRARY") or os.path.join(_HERE, "libphaze_b200.so")

PVB_OK, PVB_ERR_BAD_SIZE, PVB_ERR_BAD_ARG, PVB_ERR_CUDA, PVB_ERR_NOMEM = 0, -1, -2, -3, -4
# pvb_set_option (include/phaze_b200.h)
PVB_OPT_KERNEL, PVB_OPT_LAUNCH_MODE, PVB_OPT_INPUTS_READY, PVB_OPT_PEAK_GUARD, PVB_OPT_MANY_MODE = 1, 2, 3, 4, 5
KERNEL_AUTO, KERNEL_RING, KERNEL_WARP, KERNEL_CTA, KERNEL_GENERIC = 0, 1, 2, 3, 4


class PvbConfig(C.Structure):
    """struct pvb_config (include/phaze_b200.h)"""
    _fields_ = [("frame_size", C.c_int32), ("hop_size", C.c_int32),
                ("num_channels", C.c_int32), ("device", C.c_int32)]


class PhazeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[pvb {code}] {message}")
        self.code = code


_f32p = C.POINTER(C.c_float)
_SIGNATURES = {
    "pvb_version": (C.c_int32, []),
    "pvb_error_string": (C.c_char_p, [C.c_int32]),
    "pvb_create": (C.c_int32, [C.POINTER(PvbConfig), C.POINTER(C.c_void_p)]),
    "pvb_destroy": (None, [C.c_void_p]),
    "pvb_last_error": (C.c_char_p, [C.c_void_p]),
    "pvb_process": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]),
    "pvb_process_device": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "pvb_process_many_device": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_void_p]),
    "pvb_process_many": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float]),
    "pvb_process_pf": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pvb_process_pf_device": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pvb_sync": (C.c_int32, [C.c_void_p]),
    "pvb_set_option": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int64]),
    "pvb_get_option": (C.c_int64, [C.c_void_p, C.c_int32]),
    "pvb_resize": (C.c_int32, [C.c_void_p, C.c_int32]),
    "pvb_reset": (C.c_int32, [C.c_void_p]),
    "pvb_frame_size": (C.c_int32, [C.c_void_p]),
    "pvb_hop_size": (C.c_int32, [C.c_void_p]),
    "pvb_num_channels": (C.c_int32, [C.c_void_p]),
    "pvb_time_cursor": (C.c_double, [C.c_void_p]),
    "pvb_set_time_cursor": (C.c_int32, [C.c_void_p, C.c_double]),
    "pvb_kernel_launches": (C.c_int64, [C.c_void_p]),
    "pvb_ring_stuck_count": (C.c_int64, [C.c_void_p]),
    "pvb_peak_guard_count": (C.c_int64, [C.c_void_p]),
    "pvb_kernel_name": (C.c_char_p, [C.c_void_p, C.c_float]),
    "pvb_state_bytes": (C.c_size_t, [C.c_void_p]),
    "pvb_get_state": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "pvb_set_state": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "pvb_multi_create": (C.c_int32, [C.POINTER(PvbConfig), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p)]),
    "pvb_multi_destroy": (None, [C.c_void_p]),
    "pvb_multi_last_error": (C.c_char_p, [C.c_void_p]),
    "pvb_multi_num_devices": (C.c_int32, [C.c_void_p]),
    "pvb_multi_num_channels": (C.c_int32, [C.c_void_p]),
    "pvb_multi_shard": (C.c_void_p, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pvb_multi_set_option": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int64]),
    "pvb_multi_process": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]),
    "pvb_multi_process_many": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float]),
    "pvb_multi_process_root": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float]),
    "pvb_alloc_host": (C.c_void_p, [C.c_size_t]),
    "pvb_free_host": (None, [C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libphaze_b200.so and type every entry point.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C phaze_b200/csrc`.  phaze_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(handle, code: int):
    if code == PVB_OK:
        return
    lib = load()
    msg = lib.pvb_last_error(handle) if handle else lib.pvb_last_error(None)
    text = (msg or b"").decode() or lib.pvb_error_string(code).decode()
    raise PhazeError(code, text)
