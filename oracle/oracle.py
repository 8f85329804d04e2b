"""ctypes view of oracle/liboracle.so (the C restatement of the reference algorithm).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product (entreepy_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OK, ERR_QUEUE_EMPTY, ERR_NO_SPACE, ERR_CORRUPT, ERR_HANG = 0, 1, 2, 3, 4


class OracleError(Exception):
    def __init__(self, code):
        super().__init__({1: "QueueEmpty", 2: "NoSpaceLeft", 3: "Corrupt", 4: "ReferenceDecoderHang"}.get(code, str(code)))
        self.code = code


class Code(ctypes.Structure):
    _fields_ = [("data", ctypes.c_uint32), ("length", ctypes.c_uint8)]


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        u8p, u64p, szp = ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)
        L.oracle_histogram.argtypes = [u8p, ctypes.c_size_t, u64p]
        L.oracle_histogram.restype = None
        L.oracle_sort_symbols.argtypes = [u64p, u8p]
        L.oracle_sort_symbols.restype = ctypes.c_int
        L.oracle_build_dictionary.argtypes = [u64p, ctypes.POINTER(Code)]
        L.oracle_build_dictionary.restype = ctypes.c_int
        L.oracle_encode.argtypes = [u8p, ctypes.c_size_t, u8p, ctypes.c_size_t, szp, ctypes.POINTER(Code)]
        L.oracle_encode.restype = ctypes.c_int
        for f in (L.oracle_decode, L.oracle_decode_ref):
            f.argtypes = [u8p, ctypes.c_size_t, u8p, ctypes.c_size_t, szp]
            f.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _arr(data):
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    return np.ascontiguousarray(a, dtype=np.uint8)


def histogram(data):
    a = _arr(data)
    occ = np.zeros(256, dtype=np.uint64)
    lib().oracle_histogram(a.ctypes.data, a.size, occ.ctypes.data)
    return occ


def sort_symbols(occ):
    occ = np.ascontiguousarray(occ, dtype=np.uint64)
    out = np.zeros(256, dtype=np.uint8)
    n = lib().oracle_sort_symbols(occ.ctypes.data, out.ctypes.data)
    return out[:n].copy(), n


def build_dictionary(occ):
    """-> (data[256] uint32, length[256] uint8); raises OracleError(QueueEmpty) on all-zero counts."""
    occ = np.ascontiguousarray(occ, dtype=np.uint64)
    d = (Code * 256)()
    rc = lib().oracle_build_dictionary(occ.ctypes.data, d)
    if rc:
        raise OracleError(rc)
    return (np.array([c.data for c in d], dtype=np.uint32), np.array([c.length for c in d], dtype=np.uint8))


def encode(data, cap=None):
    """Whole .et file (magic included) as a numpy uint8 array — encode.zig:25."""
    a = _arr(data)
    if cap is None:
        cap = 7200 + a.size  # encode.zig:253-254
    out = np.empty(cap, dtype=np.uint8)
    n = ctypes.c_size_t(0)
    rc = lib().oracle_encode(a.ctypes.data, a.size, out.ctypes.data, cap, ctypes.byref(n), None)
    if rc:
        raise OracleError(rc)
    return out[: n.value].copy()


def _dec(fn, et_after_magic, cap):
    a = _arr(et_after_magic)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    n = ctypes.c_size_t(0)
    rc = fn(a.ctypes.data, a.size, out.ctypes.data, cap, ctypes.byref(n))
    return rc, out[: n.value].copy()


def decode(et_after_magic, cap):
    """Original bytes from file[4..] — correct bit-serial decode of the .et layout."""
    rc, out = _dec(lib().oracle_decode, et_after_magic, cap)
    if rc:
        raise OracleError(rc)
    return out


def decode_ref(et_after_magic, cap):
    """decode.zig:13 restated with its defects; returns (rc, bytes)."""
    return _dec(lib().oracle_decode_ref, et_after_magic, cap)
