"""ctypes binding of the C ABI declared in ``include/slowfast_b200.h``.

The product path has no CPU fallback: if the native library is missing, or is asked to run without a CUDA
device, the call fails loudly (``NativeLibraryError``).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from .build import lib_path


class NativeLibraryError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    """Mirror of ``sfb_conv_desc`` (include/slowfast_b200.h)."""

    _fields_ = [
        ("a_hi", C.c_void_p), ("a_lo", C.c_void_p),
        ("n", C.c_int32), ("d", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
        ("c_pitch", C.c_int64),
        ("b_hi", C.c_void_p), ("b_lo", C.c_void_p),
        ("cout", C.c_int32),
        ("kt", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
        ("dil_t", C.c_int32), ("dil_h", C.c_int32), ("dil_w", C.c_int32),
        ("str_t", C.c_int32), ("str_h", C.c_int32), ("str_w", C.c_int32),
        ("low_t", C.c_int32), ("low_h", C.c_int32), ("low_w", C.c_int32),
        ("out_t", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32),
        ("out", C.c_void_p),
        ("os_n", C.c_int64), ("os_t", C.c_int64), ("os_h", C.c_int64), ("os_w", C.c_int64),
        ("accumulate", C.c_int32),
        ("stats", C.c_void_p),
        ("nsplit", C.c_int32),
    ]


class WgradDesc(C.Structure):
    """Mirror of ``sfb_wgrad_desc``."""

    _fields_ = [
        ("x_hi", C.c_void_p), ("x_lo", C.c_void_p),
        ("n", C.c_int32), ("d", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
        ("c_pitch", C.c_int64),
        ("dy_hi", C.c_void_p), ("dy_lo", C.c_void_p),
        ("cout", C.c_int32), ("dy_pitch", C.c_int64),
        ("kt", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
        ("dil_t", C.c_int32), ("dil_h", C.c_int32), ("dil_w", C.c_int32),
        ("str_t", C.c_int32), ("str_h", C.c_int32), ("str_w", C.c_int32),
        ("low_t", C.c_int32), ("low_h", C.c_int32), ("low_w", C.c_int32),
        ("out_t", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32),
        ("dw", C.c_void_p),
        ("nsplit", C.c_int32),
    ]


_LIB = None

# every symbol include/slowfast_b200.h declares: (name, restype, argtypes)
_SIGNATURES = [
    ("sfb_last_error", C.c_char_p, []),
    ("sfb_abi_version", C.c_int, []),
    ("sfb_build_arch", C.c_char_p, []),
    ("sfb_conv_m_tiles", C.c_int64, [C.POINTER(ConvDesc)]),
    ("sfb_conv_igemm", C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    ("sfb_conv_wgrad", C.c_int, [C.POINTER(WgradDesc), C.c_void_p]),
]


def exported_symbols():
    return [s[0] for s in _SIGNATURES]


def load() -> C.CDLL:
    """Load ``libsfb200.so`` (built in-tree by ``slowfast_b200.build.build_native``)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = Path(lib_path())
    if not path.exists():
        raise NativeLibraryError(
            f"{path} is missing: build it with `python -m slowfast_b200.build` (nvcc, sm_100a). "
            "slowfast_b200 has no CPU fallback.")
    lib = C.CDLL(str(path))
    for name, res, args in _SIGNATURES:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().sfb_last_error()
        raise NativeLibraryError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")
