"""ctypes loader for oracle/liboracle.so — the CHECKER. Only tests/, smoke() and bench.py's CPU legs use this."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None

STREAM_END, STREAM_TRUNCATED, STREAM_BAD, STREAM_OUTPUT_FULL = 0, 1, 2, 3
CHUNK = 65535


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(ORACLE_DIR, "zwz_oracle.c")):
            subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True, capture_output=True)
        L = C.CDLL(so)
        u8p = C.POINTER(C.c_uint8)
        L.oracle_adler32.restype = C.c_uint32
        L.oracle_adler32.argtypes = [C.c_void_p, C.c_size_t]
        L.oracle_md5_hex.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p]
        L.oracle_ref_md5_hex.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p]
        L.oracle_inflate.restype = C.c_int
        L.oracle_inflate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
        L.oracle_ref_deflate_chunk.restype = C.c_long
        L.oracle_ref_deflate_chunk.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.oracle_ref_deflate_bound_size.restype = C.c_long
        L.oracle_ref_deflate_bound_size.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
        L.oracle_ref_inflate_chunk.restype = C.c_longlong
        L.oracle_ref_inflate_chunk.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.oracle_ref_deflate_batch.restype = C.c_uint64
        L.oracle_ref_deflate_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.oracle_ref_inflate_batch.restype = C.c_uint64
        L.oracle_ref_inflate_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_ref_md5_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        _LIB = L
    return _LIB


def _buf(b):
    a = np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b
    return a, a.ctypes.data if a.size else None


def adler32(b) -> int:
    a, p = _buf(b)
    return lib().oracle_adler32(p, a.size)


def md5_hex(b) -> str:
    a, p = _buf(b)
    out = C.create_string_buffer(32)
    lib().oracle_md5_hex(p, a.size, out)
    return out.raw.decode()


def ref_md5_hex(b) -> str:
    a, p = _buf(b)
    out = C.create_string_buffer(32)
    lib().oracle_ref_md5_hex(p, a.size, out)
    return out.raw.decode()


def inflate(comp, cap=1 << 20):
    """-> (bytes produced (clipped to cap), status, full length)"""
    a, p = _buf(comp)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    n = C.c_uint64(0)
    st = lib().oracle_inflate(p, a.size, out.ctypes.data, cap, C.byref(n))
    return out[:min(n.value, cap)].tobytes(), st, n.value


def ref_deflate_chunk(raw) -> bytes:
    """compression.cpp:119-134 — including the silent truncation at 65 535 bytes."""
    a, p = _buf(raw)
    out = np.empty(CHUNK, dtype=np.uint8)
    n = lib().oracle_ref_deflate_chunk(p, a.size, out.ctypes.data)
    return out[:n].tobytes()


def ref_deflate_size(raw, level=6) -> int:
    a, p = _buf(raw)
    return lib().oracle_ref_deflate_bound_size(p, a.size, level)


def ref_inflate_chunk(comp, cap=1 << 21) -> bytes:
    """decompression.cpp:11-37 with avail_in = len(comp)."""
    a, p = _buf(comp)
    out = np.empty(cap, dtype=np.uint8)
    n = lib().oracle_ref_inflate_chunk(p, a.size, out.ctypes.data, cap)
    assert n >= 0
    return out[:n].tobytes()
