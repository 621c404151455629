"""ctypes wrapper of oracle/_build/liboracle.so (the C restatement, oracle/oracle.c).
TEST INFRASTRUCTURE ONLY - see the header of oracle.c."""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build():
    src = os.path.join(_HERE, "oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(src) > os.path.getmtime(_SO):
        subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _b32(v: int) -> bytes:
    return int(v).to_bytes(32, "little")


def _pt(p) -> bytes:
    return bytes(64) if p is None else _b32(p[0]) + _b32(p[1])


def _unpt(b: bytes):
    return None if b == bytes(64) else (int.from_bytes(b[:32], "little"), int.from_bytes(b[32:], "little"))


def fe_op(field: int, op: int, a: int, b: int = 0) -> int:
    out = ctypes.create_string_buffer(32)
    assert lib().oracle_fe_op(field, op, _b32(a), _b32(b), out) == 0
    return int.from_bytes(out.raw, "little")


def g1_add(a, b):
    out = ctypes.create_string_buffer(64)
    lib().oracle_g1_add(_pt(a), _pt(b), out)
    return _unpt(out.raw)


def g1_mul(p, k: int):
    out = ctypes.create_string_buffer(64)
    lib().oracle_g1_mul(_pt(p), _b32(k), out)
    return _unpt(out.raw)


def setup_kzg_bytes(alpha: int, n: int, threads: int = 1) -> bytes:
    out = ctypes.create_string_buffer(64 * max(n, 1))
    lib().oracle_setup_kzg(_b32(alpha), ctypes.c_size_t(n), out, threads)
    return out.raw[: 64 * n]


def commit_kzg_bytes(coefs: bytes, points: bytes, n: int, threads: int = 1):
    out = ctypes.create_string_buffer(64)
    lib().oracle_commit_kzg(coefs, points, ctypes.c_size_t(n), out, threads)
    return _unpt(out.raw)


def open_kzg_bytes(coefs: bytes, n: int, u: int, points: bytes, threads: int = 1):
    y = ctypes.create_string_buffer(32)
    w = ctypes.create_string_buffer(64)
    lib().oracle_open_kzg(coefs, ctypes.c_size_t(n), _b32(u), points, y, w, threads)
    return int.from_bytes(y.raw, "little"), _unpt(w.raw)


def fr_eval_bytes(coefs: bytes, n: int, u: int) -> int:
    y = ctypes.create_string_buffer(32)
    lib().oracle_fr_eval(coefs, ctypes.c_size_t(n), _b32(u), y)
    return int.from_bytes(y.raw, "little")


def quotient_bytes(coefs: bytes, n: int, u: int):
    y = ctypes.create_string_buffer(32)
    q = ctypes.create_string_buffer(32 * max(n - 1, 1))
    lib().oracle_quotient(coefs, ctypes.c_size_t(n), _b32(u), y, q)
    return int.from_bytes(y.raw, "little"), [int.from_bytes(q.raw[32 * i : 32 * i + 32], "little") for i in range(max(n - 1, 0))]


def fold_bytes(coefs: bytes, n_out: int, rho: int) -> bytes:
    out = ctypes.create_string_buffer(32 * max(n_out, 1))
    lib().oracle_fold(coefs, ctypes.c_size_t(n_out), _b32(rho), out)
    return out.raw[: 32 * n_out]
