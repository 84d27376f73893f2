"""ctypes binding of libwfe.so (C ABI: include/wfe.h).  Fails loudly — there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WFE_LIB_OVERRIDE") or os.path.join(_HERE, "libwfe.so")  # override: timing what-if builds

WFE_PCM_F32, WFE_PCM_I16, WFE_PCM_F16 = 0, 1, 2
WFE_OUT_F32, WFE_OUT_F16, WFE_OUT_BF16 = 0, 1, 2

# every symbol include/wfe.h declares (tests/test_abi_cpu.py checks the library exports each one)
SYMBOLS = [
    "wfe_create", "wfe_destroy", "wfe_last_error", "wfe_abi_version", "wfe_launch_count",
    "wfe_logmel_scratch_bytes", "wfe_n_frames", "wfe_logmel", "wfe_clip_stats", "wfe_collate", "wfe_extract_host",
    "wfe_logmel_ex", "wfe_extract_host_ex", "wfe_uses_tensor_cores", "wfe_debug_scratch_error",
]


class WfeConfig(C.Structure):
    _fields_ = [("n_mel", C.c_int32), ("n_fft", C.c_int32), ("hop_length", C.c_int32), ("n_samples", C.c_int32),
                ("sampling_rate", C.c_int32), ("device", C.c_int32)]


class WfeError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """dlopen libwfe.so (built in-tree by `__graft_entry__.build()` / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the sm_100a extension first (python -c 'import __graft_entry__ as g; "
            "g.build()' or `make -C asr-finetune_b200/csrc`). There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.wfe_create.argtypes = [C.POINTER(WfeConfig), vp, C.POINTER(vp)]
    lib.wfe_create.restype = C.c_int
    lib.wfe_destroy.argtypes = [vp]
    lib.wfe_destroy.restype = None
    lib.wfe_last_error.argtypes = []
    lib.wfe_last_error.restype = C.c_char_p
    lib.wfe_abi_version.argtypes = []
    lib.wfe_abi_version.restype = C.c_int
    lib.wfe_launch_count.argtypes = []
    lib.wfe_launch_count.restype = C.c_uint64
    lib.wfe_logmel_scratch_bytes.argtypes = [vp, i32]
    lib.wfe_logmel_scratch_bytes.restype = C.c_size_t
    lib.wfe_n_frames.argtypes = [vp]
    lib.wfe_n_frames.restype = i32
    lib.wfe_logmel.argtypes = [vp, vp, i32, f32, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.wfe_logmel.restype = C.c_int
    lib.wfe_logmel_ex.argtypes = [vp, vp, i32, f32, vp, vp, i32, vp, vp, i32, vp, vp, vp]
    lib.wfe_logmel_ex.restype = C.c_int
    lib.wfe_uses_tensor_cores.argtypes = [vp]
    lib.wfe_uses_tensor_cores.restype = i32
    lib.wfe_debug_scratch_error.argtypes = [vp, vp, i32]
    lib.wfe_debug_scratch_error.restype = i32
    lib.wfe_extract_host_ex.argtypes = [vp, vp, vp, i32, i32, f32, i32, vp, i32, vp, C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_uint64)]
    lib.wfe_extract_host_ex.restype = C.c_int
    lib.wfe_clip_stats.argtypes = [vp, vp, i32, f32, vp, vp, i32, vp, vp]
    lib.wfe_clip_stats.restype = C.c_int
    lib.wfe_collate.argtypes = [vp, vp, vp, i32, i32, i64, i64, vp, vp, vp, i64, vp, vp]
    lib.wfe_collate.restype = C.c_int
    lib.wfe_extract_host.argtypes = [vp, vp, vp, i32, i32, f32, i32, vp, vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.wfe_extract_host.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().wfe_last_error().decode("utf-8", "replace")
        raise WfeError(f"{what} failed ({rc}): {msg}")


def launch_count() -> int:
    return int(load().wfe_launch_count())
