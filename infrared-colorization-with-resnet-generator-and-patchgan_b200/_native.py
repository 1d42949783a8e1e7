"""ctypes binding of libirc_sm100.so (include/irc_b200.h).

There is no fallback: if the library is missing, or the device is not sm_100, every entry
point raises.  Nothing here touches ``oracle/``."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libirc_sm100.so")
MAX_TAPS = 64


class IrcError(RuntimeError):
    pass


class ConvGemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_rows", C.c_longlong), ("a_ld", C.c_int), ("a_chan_off", C.c_int), ("cin", C.c_int),
        ("ntaps", C.c_int), ("taps", C.c_int * MAX_TAPS),
        ("w", C.c_void_p), ("n_out", C.c_int),
        ("out", C.c_void_p), ("out_ld", C.c_longlong), ("out_chan_off", C.c_int), ("out_fp32", C.c_int),
        ("bias", C.c_void_p), ("act", C.c_int), ("slope", C.c_float),
        ("row_img", C.c_void_p),
        ("mask", C.c_void_p), ("mask_ld", C.c_longlong), ("mask_chan_off", C.c_int), ("mask_slope", C.c_float),
        ("bn", C.c_int),
    ]


class TnGemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_rows", C.c_longlong), ("a_ld", C.c_int), ("a_chan_off", C.c_int), ("m", C.c_int),
        ("b", C.c_void_p), ("b_rows", C.c_longlong), ("b_ld", C.c_int), ("b_chan_off", C.c_int), ("n", C.c_int),
        ("k_rows", C.c_longlong),
        ("ntaps", C.c_int), ("a_shift", C.c_int * MAX_TAPS), ("b_shift", C.c_int * MAX_TAPS),
        ("out", C.c_void_p),
        ("out_tap_stride", C.c_longlong), ("out_m_stride", C.c_longlong), ("out_n_stride", C.c_longlong),
        ("out_split_stride", C.c_longlong),
        ("splits", C.c_int), ("bn", C.c_int),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises IrcError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IrcError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU or PyTorch fallback for the hot path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.irc_last_error.restype = C.c_char_p
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise IrcError(f"libirc_sm100 error {rc}: {lib().irc_last_error().decode()}")


def arch_check() -> None:
    check(lib().irc_arch_check())
