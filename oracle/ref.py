"""
ORACLE (test infrastructure, NOT product code) -- ctypes wrapper around oracle/libkzg_ref.so, the C
restatement of the KZG hot path (see the header of kzg_ref.c for what it restates and how it is
pinned).  Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Sequence, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkzg_ref.so")
_lib = None


def build() -> str:
    src = os.path.join(_HERE, "kzg_ref.c")
    if not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        c = ctypes
        _lib.ref_ntt.argtypes = [c.c_char_p, c.c_char_p, c.c_size_t, c.c_int]
        _lib.ref_eval.argtypes = [c.c_char_p, c.c_size_t, c.c_char_p, c.c_char_p]
        _lib.ref_msm.argtypes = [c.c_char_p, c.c_char_p, c.c_size_t, c.c_char_p, c.c_int]
        _lib.ref_quotient_evals.argtypes = [c.c_char_p, c.c_size_t, c.c_char_p, c.c_char_p, c.c_char_p]
        _lib.ref_g1_mul_gen.argtypes = [c.c_char_p, c.c_char_p]
        _lib.ref_srs.argtypes = [c.c_char_p, c.c_char_p, c.c_size_t, c.c_int, c.c_char_p, c.c_int]
        _lib.ref_fr_dot.argtypes = [c.c_char_p, c.c_char_p, c.c_size_t, c.c_char_p]
        _lib.ref_lagrange_scalars.argtypes = [c.c_char_p, c.c_char_p, c.c_size_t, c.c_char_p]
        _lib.ref_random_scalars.argtypes = [c.c_uint64, c.c_size_t, c.c_char_p]
        _lib.ref_random_scalars.restype = None
        _lib.ref_random_scalars_ctr.argtypes = [c.c_uint64, c.c_uint64, c.c_size_t, c.c_char_p]
        _lib.ref_random_scalars_ctr.restype = None
    return _lib


def _chk(rc: int, what: str) -> None:
    if rc != 0:
        raise ValueError(f"oracle {what} failed: rc={rc}")


def fr_be(v: int) -> bytes:
    return int(v).to_bytes(32, "big")


def ntt(vals_be: bytes, inverse: bool = False) -> bytes:
    n = len(vals_be) // 32
    out = ctypes.create_string_buffer(32 * n)
    _chk(lib().ref_ntt(vals_be, out, n, int(inverse)), "ntt")
    return out.raw


def eval_coeffs(coeffs_be: bytes, x_be: bytes) -> bytes:
    out = ctypes.create_string_buffer(32)
    _chk(lib().ref_eval(coeffs_be, len(coeffs_be) // 32, x_be, out), "eval")
    return out.raw


def msm(points96: bytes, scalars_be: bytes, threads: int = 1) -> bytes:
    n = len(scalars_be) // 32
    assert len(points96) >= 96 * n
    out = ctypes.create_string_buffer(48)
    _chk(lib().ref_msm(points96, scalars_be, n, out, threads), "msm")
    return out.raw


def quotient_evals(evals_be: bytes, x_be: bytes) -> Tuple[bytes, bytes]:
    """(y, q evals) for evaluation-form f on the size-n domain."""
    n = len(evals_be) // 32
    y = ctypes.create_string_buffer(32)
    q = ctypes.create_string_buffer(32 * n)
    _chk(lib().ref_quotient_evals(evals_be, n, x_be, y, q), "quotient")
    return y.raw, q.raw


def open_evals(evals_be: bytes, x_be: bytes, srs_lagrange96: bytes, threads: int = 1) -> Tuple[bytes, bytes]:
    y, q = quotient_evals(evals_be, x_be)
    return y, msm(srs_lagrange96, q, threads)


def g1_mul_gen(k_be: bytes) -> bytes:
    out = ctypes.create_string_buffer(48)
    _chk(lib().ref_g1_mul_gen(k_be, out), "g1_mul_gen")
    return out.raw


def srs(n: int, tau: int, kind: str = "lagrange", scale: int = 1, threads: int = 0) -> bytes:
    """ZCash-uncompressed (96 B) points [scale*tau^j]G ("monomial") or [scale*L_j(tau)]G ("lagrange")."""
    out = ctypes.create_string_buffer(96 * n)
    threads = threads or (os.cpu_count() or 1)
    _chk(lib().ref_srs(fr_be(tau), fr_be(scale), n, 0 if kind == "monomial" else 1, out, threads), "srs")
    return out.raw


def fr_dot(a_be: bytes, b_be: bytes) -> bytes:
    out = ctypes.create_string_buffer(32)
    _chk(lib().ref_fr_dot(a_be, b_be, len(a_be) // 32, out), "fr_dot")
    return out.raw


def lagrange_scalars(n: int, tau: int, scale: int = 1) -> bytes:
    out = ctypes.create_string_buffer(32 * n)
    _chk(lib().ref_lagrange_scalars(fr_be(tau), fr_be(scale), n, out), "lagrange_scalars")
    return out.raw


def random_scalars(seed: int, n: int) -> bytes:
    out = ctypes.create_string_buffer(32 * n)
    lib().ref_random_scalars(seed, n, out)
    return out.raw


def random_scalars_ctr(seed: int, first: int, n: int) -> bytes:
    """elements [first, first + n) of the counter-based stream of the product's device generator (zkp_random_poly)"""
    out = ctypes.create_string_buffer(32 * n)
    lib().ref_random_scalars_ctr(seed, first, n, out)
    return out.raw


def split32(buf: bytes) -> List[int]:
    return [int.from_bytes(buf[i:i + 32], "big") for i in range(0, len(buf), 32)]


def join32(vals: Sequence[int]) -> bytes:
    return b"".join(int(v).to_bytes(32, "big") for v in vals)
