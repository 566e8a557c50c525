"""torchac 0.9.3 contract restated -- TEST INFRASTRUCTURE.

PARITY UNPINNED against real torchac bytes (its source is not under
/root/reference, ``README.md:20``).  ``encode_float_cdf`` / ``decode_float_cdf``
mirror the two calls the reference makes (``entropy_model.py:174,192``).

Two implementations of the same algorithm live here so they can check each
other: ``rc_ref.c`` (fast, via ctypes; built by ``oracle/Makefile``) and the
pure-Python loops ``py_encode`` / ``py_decode`` for small cases.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
PRECISION = 16


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "librc_ref.so")
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        lib.rc_ref_encode.restype = ctypes.c_size_t
        lib.rc_ref_encode.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_int64, ctypes.c_void_p, ctypes.c_size_t]
        lib.rc_ref_decode.restype = None
        lib.rc_ref_decode.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int64]
        _LIB = lib
    return _LIB


def cdf_float_to_u16(cdf_float: np.ndarray) -> np.ndarray:
    """Appendix B.1 (torchac ``_convert_to_int_and_normalize`` with
    ``needs_normalization=True``): ``round(cdf * (2^16 - (Lp-1)))`` cast to
    int16 (wraps), ``+ arange(Lp)``, bits reinterpreted as uint16."""
    cdf_float = np.asarray(cdf_float, dtype=np.float32)
    Lp = cdf_float.shape[-1]
    scaled = np.round(cdf_float * np.float32(2 ** PRECISION - (Lp - 1)))      # half-to-even like torch.round
    as_i16 = scaled.astype(np.int64).astype(np.uint16)                        # two's-complement wrap
    return (as_i16 + np.arange(Lp, dtype=np.uint16)).astype(np.uint16)


def _check(cdf_float, sym):
    if np.any(cdf_float < 0) or np.any(cdf_float > 1):
        raise ValueError("cdf_float out of [0, 1]")
    Lp = cdf_float.shape[-1]
    if np.any(sym < 0) or np.any(sym > Lp - 2):
        raise ValueError("symbol out of range")


def _table_and_rows(cdf_float: np.ndarray):
    """Collapse a tiled ``[N, C, Lp]`` table to its unique leading rows."""
    cdf_float = np.asarray(cdf_float, dtype=np.float32)
    Lp = cdf_float.shape[-1]
    flat = cdf_float.reshape(-1, Lp)
    n_sym = flat.shape[0]
    if cdf_float.ndim == 3 and cdf_float.shape[0] > 1 and (cdf_float == cdf_float[:1]).all():
        C = cdf_float.shape[1]
        table = cdf_float_to_u16(cdf_float[0])
        rows = np.tile(np.arange(C, dtype=np.int32), cdf_float.shape[0])
    else:
        table = cdf_float_to_u16(flat)
        rows = np.arange(n_sym, dtype=np.int32)
    return np.ascontiguousarray(table), np.ascontiguousarray(rows), Lp, n_sym


def encode_float_cdf(cdf_float, sym, check_input_bounds=False) -> bytes:
    cdf_float = np.asarray(cdf_float, dtype=np.float32)
    sym = np.ascontiguousarray(np.asarray(sym, dtype=np.int16))
    assert cdf_float.shape[:-1] == sym.shape
    if check_input_bounds:
        _check(cdf_float, sym)
    table, rows, Lp, n_sym = _table_and_rows(cdf_float)
    return encode_u16(table, rows, sym.reshape(-1))


def decode_float_cdf(cdf_float, data: bytes) -> np.ndarray:
    cdf_float = np.asarray(cdf_float, dtype=np.float32)
    table, rows, Lp, n_sym = _table_and_rows(cdf_float)
    return decode_u16(table, rows, data).reshape(cdf_float.shape[:-1])


def encode_u16(table: np.ndarray, rows: np.ndarray, sym: np.ndarray) -> bytes:
    table = np.ascontiguousarray(table, dtype=np.uint16)
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    sym = np.ascontiguousarray(sym, dtype=np.int16)
    n_sym = sym.size
    cap = 4 * n_sym + 64
    out = np.empty(cap, dtype=np.uint8)
    n = _lib().rc_ref_encode(table.ctypes.data, table.shape[-1], rows.ctypes.data, sym.ctypes.data,
                             n_sym, out.ctypes.data, cap)
    assert n <= cap
    return out[:n].tobytes()


def decode_u16(table: np.ndarray, rows: np.ndarray, data: bytes) -> np.ndarray:
    table = np.ascontiguousarray(table, dtype=np.uint16)
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    buf = np.frombuffer(data, dtype=np.uint8)
    out = np.empty(rows.size, dtype=np.int16)
    _lib().rc_ref_decode(table.ctypes.data, table.shape[-1], rows.ctypes.data,
                         buf.ctypes.data if buf.size else None, buf.size, out.ctypes.data, rows.size)
    return out


# ---------------------------------------------------------------- pure Python (small cases)

def py_encode(table, rows, sym) -> bytes:
    """Appendix B.2 transcribed as Python integer arithmetic."""
    M32 = 0xFFFFFFFF
    Lp = table.shape[-1]
    low, high, pending = 0, M32, 0
    bits = []

    def emit(b):
        nonlocal pending
        bits.append(b)
        bits.extend([1 - b] * pending)
        pending = 0

    for r, s in zip(rows.tolist(), sym.tolist()):
        span = high - low + 1
        c_low = int(table[r, s])
        c_high = 0x10000 if s == Lp - 2 else int(table[r, s + 1])
        high = (low - 1 + ((span * c_high) >> 16)) & M32
        low = (low + ((span * c_low) >> 16)) & M32
        while True:
            if high < 0x80000000:
                emit(0)
                low = (low << 1) & M32
                high = ((high << 1) | 1) & M32
            elif low >= 0x80000000:
                emit(1)
                low = (low << 1) & M32
                high = ((high << 1) | 1) & M32
            elif low >= 0x40000000 and high < 0xC0000000:
                pending += 1
                low = (low << 1) & 0x7FFFFFFF
                high = ((high << 1) | 0x80000001) & M32
            else:
                break
    pending += 1
    emit(0 if low < 0x40000000 else 1)
    while len(bits) % 8:
        bits.append(0)
    return np.packbits(np.asarray(bits, dtype=np.uint8)).tobytes()


def py_decode(table, rows, data: bytes) -> np.ndarray:
    """Appendix B.3 transcribed as Python integer arithmetic."""
    M32 = 0xFFFFFFFF
    Lp = table.shape[-1]
    bits = np.unpackbits(np.frombuffer(data, dtype=np.uint8)).tolist()
    pos = 0

    def nxt():
        nonlocal pos
        b = bits[pos] if pos < len(bits) else 0
        pos += 1
        return b

    low, high, value = 0, M32, 0
    for _ in range(32):
        value = (value << 1) | nxt()
    out = np.empty(len(rows), dtype=np.int16)
    n = len(rows)
    for i, r in enumerate(rows.tolist()):
        span = high - low + 1
        count = (((value - low + 1) << 16) - 1) // span & 0xFFFF
        row = table[r]
        s = 0
        for cand in range(Lp - 1):                      # linear scan: largest s with row[s] <= count
            if int(row[cand]) <= count:
                s = cand
        out[i] = s
        if i == n - 1:
            break
        c_low = int(row[s])
        c_high = 0x10000 if s == Lp - 2 else int(row[s + 1])
        high = (low - 1 + ((span * c_high) >> 16)) & M32
        low = (low + ((span * c_low) >> 16)) & M32
        while True:
            if low >= 0x80000000 or high < 0x80000000:
                low = (low << 1) & M32
                high = ((high << 1) | 1) & M32
                value = ((value << 1) | nxt()) & M32
            elif low >= 0x40000000 and high < 0xC0000000:
                low = (low << 1) & 0x7FFFFFFF
                high = ((high << 1) | 0x80000001) & M32
                value = (value - 0x40000000) & M32
                value = ((value << 1) | nxt()) & M32
            else:
                break
    return out
