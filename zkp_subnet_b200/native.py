"""ctypes binding of libzkp_b200.so (the C ABI declared in include/zkp_b200.h).

The library is CUDA-only.  There is no CPU fallback: if the shared object is missing this module
raises at import of the first symbol, and if no GPU is present `Context()` raises `ZkpError`.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZKP_B200_LIB") or os.path.join(_HERE, "libzkp_b200.so")  # env override: tuning builds

ZKP_OK = 0
ZKP_ERR_ARG = -1
ZKP_ERR_ENCODING = -2
ZKP_ERR_CUDA = -3
ZKP_ERR_STATE = -4
ZKP_ERR_IO = -5


class ZkpError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"zkp_b200 error {code}: {message}")
        self.code = code


_lib: Optional[ctypes.CDLL] = None

_u8p = ctypes.c_char_p
_ctxp = ctypes.c_void_p

# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGNATURES = {
    "zkp_ctx_create": [ctypes.c_int, ctypes.POINTER(_ctxp)],
    "zkp_ctx_fork": [_ctxp, ctypes.POINTER(_ctxp)],
    "zkp_ctx_destroy": [_ctxp],
    "zkp_last_error": [],
    "zkp_device_count": [],
    "zkp_host_alloc": [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)],
    "zkp_host_free": [ctypes.c_void_p],
    "zkp_srs_generate": [_ctxp, _u8p, _u8p, ctypes.c_uint32, ctypes.c_uint32],
    "zkp_srs_generate_monomial": [_ctxp, _u8p, ctypes.c_uint32],
    "zkp_srs_generate_shard": [_ctxp, _u8p, _u8p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32],
    "zkp_g1_sum": [_u8p, ctypes.c_size_t, _u8p],
    "zkp_g1_sum_checked": [_u8p, ctypes.c_size_t, _u8p],
    "zkp_g1_uncompress": [_u8p, _u8p],
    "zkp_last_points_uncompressed": [_ctxp, _u8p],
    "zkp_last_points_jacobian": [_ctxp, _u8p],
    "zkp_g1_sum_jacobian": [_u8p, ctypes.c_size_t, ctypes.c_size_t, _u8p],
    "zkp_g1_sum_uncompressed": [_u8p, ctypes.c_size_t, _u8p],
    "zkp_shard_eval_partial": [_ctxp, ctypes.c_uint32, _u8p, ctypes.c_size_t, _u8p, _u8p],
    "zkp_shard_eval_combine": [_u8p, ctypes.c_size_t, ctypes.c_uint32, _u8p, _u8p],
    "zkp_shard_open_partial": [_ctxp, ctypes.c_uint32, _u8p, ctypes.c_size_t, _u8p, _u8p, _u8p],
    "zkp_srs_set_shape": [_ctxp, ctypes.c_uint32, ctypes.c_uint32],
    "zkp_srs_import_row": [_ctxp, ctypes.c_uint32, _u8p, ctypes.c_size_t, _u8p],
    "zkp_srs_import_g2_tau": [_ctxp, _u8p],
    "zkp_srs_import_g2_tau_y": [_ctxp, _u8p],
    "zkp_master_open_y": [_ctxp, _u8p, ctypes.c_size_t, _u8p, _u8p, _u8p],
    "zkp_master_verify": [_ctxp, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, ctypes.POINTER(ctypes.c_int)],
    "zkp_srs_export_row": [_ctxp, ctypes.c_uint32, _u8p, ctypes.c_size_t],
    "zkp_srs_import_row_compressed": [_ctxp, ctypes.c_uint32, _u8p, ctypes.c_size_t, _u8p],
    "zkp_srs_export_row_compressed": [_ctxp, ctypes.c_uint32, _u8p, ctypes.c_size_t],
    "zkp_srs_import_g2": [_ctxp, ctypes.c_int, _u8p],
    "zkp_srs_export_g2": [_ctxp, ctypes.c_int, _u8p],
    "zkp_srs_export_scale_point": [_ctxp, ctypes.c_uint32, _u8p],
    "zkp_srs_set_shard": [_ctxp, ctypes.c_uint32, ctypes.c_uint32],
    "zkp_srs_generate_monomial2": [_ctxp, _u8p, _u8p, ctypes.c_uint32, ctypes.c_uint32],
    "zkp_srs_monomial_to_lagrange": [_ctxp],
    "zkp_srs_save": [_ctxp, ctypes.c_char_p],
    "zkp_srs_load": [_ctxp, ctypes.c_char_p],
    "zkp_srs_shape": [_ctxp, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)],
    "zkp_worker_commit": [_ctxp, ctypes.c_uint32, _u8p, ctypes.c_size_t, _u8p],
    "zkp_worker_open": [_ctxp, ctypes.c_uint32, _u8p, ctypes.c_size_t, _u8p, _u8p, _u8p],
    "zkp_worker_open_resident": [_ctxp, ctypes.c_uint32, ctypes.c_size_t, _u8p, _u8p, _u8p],
    "zkp_worker_commit_open": [_ctxp, ctypes.c_uint32, _u8p, ctypes.c_size_t, _u8p, _u8p, _u8p, _u8p],
    "zkp_resident_generation": [_ctxp, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_size_t)],
    "zkp_stage_begin": [_ctxp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64)],
    "zkp_stage_chunk": [_ctxp, ctypes.c_size_t, _u8p, ctypes.c_size_t],
    "zkp_stage_end": [_ctxp, ctypes.c_uint64],
    "zkp_worker_commit_resident": [_ctxp, ctypes.c_uint32, ctypes.c_size_t, ctypes.c_uint64, _u8p],
    "zkp_worker_commit_open_resident": [_ctxp, ctypes.c_uint32, ctypes.c_size_t, ctypes.c_uint64, _u8p, _u8p, _u8p, _u8p],
    "zkp_worker_open_resident_gen": [_ctxp, ctypes.c_uint32, ctypes.c_size_t, ctypes.c_uint64, _u8p, _u8p, _u8p],
    "zkp_worker_commit_open_batch": [_ctxp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_void_p),
                                     ctypes.c_size_t, _u8p, _u8p, _u8p, _u8p, ctypes.POINTER(ctypes.c_int)],
    "zkp_worker_verify": [_ctxp, ctypes.c_uint32, _u8p, _u8p, _u8p, _u8p, ctypes.POINTER(ctypes.c_int)],
    "zkp_worker_verify_batch": [_ctxp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32), _u8p, _u8p, _u8p, _u8p, ctypes.POINTER(ctypes.c_int)],
    "zkp_fft": [_ctxp, _u8p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, _u8p],
    "zkp_eval": [_ctxp, _u8p, ctypes.c_size_t, _u8p, _u8p],
    "zkp_challenge_evals": [_ctxp, _u8p, ctypes.c_size_t, ctypes.c_size_t, _u8p, _u8p],
    "zkp_random_poly": [_ctxp, ctypes.c_uint64, _u8p, ctypes.c_size_t],
    "zkp_random_point": [_ctxp, ctypes.c_uint64, _u8p],
    "zkp_random_poly_range": [_ctxp, ctypes.c_uint64, ctypes.c_uint64, _u8p, ctypes.c_size_t],
    "zkp_mgpu_create": [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(_ctxp)],
    "zkp_mgpu_destroy": [_ctxp],
    "zkp_mgpu_device_count": [_ctxp],
    "zkp_mgpu_ctx": [_ctxp, ctypes.c_int],
    "zkp_mgpu_set_layout": [_ctxp, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32],
    "zkp_mgpu_srs_generate": [_ctxp, _u8p, _u8p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int],
    "zkp_mgpu_prebuild_tables": [_ctxp],
    "zkp_mgpu_msm_g1": [_ctxp, ctypes.c_uint32, _u8p, ctypes.c_size_t, ctypes.c_int, _u8p],
    "zkp_mgpu_commit_open": [_ctxp, ctypes.c_uint32, _u8p, ctypes.c_size_t, _u8p, ctypes.c_int, _u8p, _u8p, _u8p],
    "zkp_mgpu_pianist_commit_open": [_ctxp, ctypes.POINTER(ctypes.c_uint32), ctypes.c_size_t, _u8p, ctypes.c_size_t, _u8p,
                                     ctypes.c_int, _u8p, _u8p, _u8p, _u8p, _u8p],
    "zkp_b64_decode_fr": [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, _u8p],
    "zkp_b64_encode_fr": [_u8p, ctypes.c_size_t, ctypes.c_char_p],
    "zkp_msm_g1": [_ctxp, ctypes.c_uint32, _u8p, ctypes.c_size_t, _u8p],
    "zkp_bench_msm": [_ctxp, ctypes.c_uint32, _u8p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                      ctypes.POINTER(ctypes.c_float), _u8p],
    "zkp_bench_commit_open": [_ctxp, ctypes.c_uint32, _u8p, ctypes.c_size_t, _u8p, ctypes.c_int, ctypes.c_int,
                              ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float),
                              ctypes.POINTER(ctypes.c_uint32), _u8p, _u8p, _u8p],
    "zkp_bench_trace": [_ctxp, ctypes.c_uint32, _u8p, ctypes.c_size_t, _u8p, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t],
    "zkp_bench_ntt": [_ctxp, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float)],
    "zkp_bench_last_kernel_ms": [_ctxp, ctypes.POINTER(ctypes.c_float)],
    "zkp_bench_flush_l2": [_ctxp],
    "zkp_bench_peaks": [_ctxp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)],
    "zkp_set_msm_mode": [_ctxp, ctypes.c_int],
    "zkp_set_msm_window": [_ctxp, ctypes.c_uint32],
    "zkp_set_msm_sort": [_ctxp, ctypes.c_int],
    "zkp_set_msm_affine_rounds": [_ctxp, ctypes.c_int],
    "zkp_msm_info": [_ctxp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32),
                     ctypes.POINTER(ctypes.c_uint64)],
    "zkp_set_fuse": [_ctxp, ctypes.c_int],
    "zkp_set_open_coset": [_ctxp, ctypes.c_int],
    "zkp_set_rowcol_coop": [_ctxp, ctypes.c_int],
    "zkp_set_ntt_tma": [_ctxp, ctypes.c_int],
    "zkp_set_poly_form": [_ctxp, ctypes.c_int],
    "zkp_srs_prebuild_tables": [_ctxp, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32)],
    "zkp_set_table_budget": [_ctxp, ctypes.c_size_t],
    "zkp_srs_table_stats": [_ctxp, ctypes.POINTER(ctypes.c_uint64)],
    "zkp_pairing_check": [_u8p, _u8p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int)],
}
_RESTYPES = {"zkp_ctx_destroy": None, "zkp_last_error": ctypes.c_char_p, "zkp_mgpu_destroy": None, "zkp_mgpu_ctx": _ctxp}
LAYOUT_ROWS, LAYOUT_POINT_RANGE = 1, 2
MGPU_RESIDENT = 1

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib() -> ctypes.CDLL:
    """Load libzkp_b200.so; fails loudly if it has not been built (`make` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ZkpError(ZKP_ERR_STATE, f"{LIB_PATH} not built; run `make` (nvcc, sm_100a). No CPU fallback exists.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, args in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().zkp_last_error()
    return msg.decode() if msg else ""


def check(rc: int) -> None:
    if rc != ZKP_OK:
        raise ZkpError(rc, last_error())


class PinnedBuffer:
    """Page-locked host staging buffer (zkp_host_alloc).  Pass it wherever a Context method takes polynomial
    bytes: the host->device copy then runs as one asynchronous DMA.  `len()` is the number of bytes in use."""

    def __init__(self, nbytes: int):
        self._p = ctypes.c_void_p()
        check(lib().zkp_host_alloc(nbytes, ctypes.byref(self._p)))
        self.capacity = nbytes
        self.used = nbytes
        self.buf = (ctypes.c_char * nbytes).from_address(self._p.value)

    def __len__(self) -> int:
        return self.used

    def write(self, data: bytes, offset: int = 0) -> "PinnedBuffer":
        if offset + len(data) > self.capacity:
            raise ValueError("PinnedBuffer overflow")
        ctypes.memmove(self._p.value + offset, data, len(data))
        self.used = offset + len(data)
        return self

    def tobytes(self) -> bytes:
        return ctypes.string_at(self._p.value, self.used)

    def close(self) -> None:
        if self._p:
            self.buf = None
            lib().zkp_host_free(self._p)
            self._p = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _arg(b):
    """bytes or PinnedBuffer -> what ctypes passes as const uint8_t*"""
    return b.buf if isinstance(b, PinnedBuffer) else b


class Context:
    """Owns one zkp_ctx (one GPU, one resident SRS).  Mirrors the lifetime of the reference's prover
    process started by Client.start() and killed by Client.stop() (reference base/miner.py:73-84,155)."""

    def __init__(self, device: int = 0, _handle=None, _borrowed: bool = False):
        self._borrowed = _borrowed
        self.device = device
        if _handle is not None:
            self._h = _handle
            return
        self._h = _ctxp()
        check(lib().zkp_ctx_create(device, ctypes.byref(self._h)))

    def fork(self) -> "Context":
        """A further context on the same device sharing this one's resident SRS and tables (own streams and
        workspaces): one per request-handling thread."""
        h = _ctxp()
        check(lib().zkp_ctx_fork(self._h, ctypes.byref(h)))
        return Context(self.device, _handle=h)

    def close(self) -> None:
        if self._h and not self._borrowed:
            lib().zkp_ctx_destroy(self._h)
        self._h = _ctxp()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- SRS
    def srs_generate(self, tau_x: int, tau_y: int, log_n: int, log_machines: int) -> None:
        check(lib().zkp_srs_generate(self._h, tau_x.to_bytes(32, "big"), tau_y.to_bytes(32, "big"), log_n, log_machines))

    def srs_generate_monomial(self, tau_x: int, log_n: int) -> None:
        check(lib().zkp_srs_generate_monomial(self._h, tau_x.to_bytes(32, "big"), log_n))

    def srs_generate_shard(self, tau_x: int, tau_y: int, log_n: int, log_machines: int, shard: int, log_shards: int) -> None:
        check(lib().zkp_srs_generate_shard(self._h, tau_x.to_bytes(32, "big"), tau_y.to_bytes(32, "big"), log_n,
                                           log_machines, shard, log_shards))

    def srs_set_shape(self, log_n: int, log_machines: int) -> None:
        check(lib().zkp_srs_set_shape(self._h, log_n, log_machines))

    def srs_import_row(self, row: int, points96: bytes, scale_point48: Optional[bytes] = None) -> None:
        check(lib().zkp_srs_import_row(self._h, row, points96, len(points96) // 96, scale_point48))

    def srs_import_g2_tau(self, tau_x: int) -> None:
        check(lib().zkp_srs_import_g2_tau(self._h, tau_x.to_bytes(32, "big")))

    def srs_import_g2_tau_y(self, tau_y: int) -> None:
        check(lib().zkp_srs_import_g2_tau_y(self._h, tau_y.to_bytes(32, "big")))

    def srs_export_row(self, row: int, n: int) -> bytes:
        out = ctypes.create_string_buffer(96 * n)
        check(lib().zkp_srs_export_row(self._h, row, out, n))
        return out.raw

    def srs_import_row_compressed(self, row: int, points48: bytes, scale_point48: Optional[bytes] = None) -> None:
        check(lib().zkp_srs_import_row_compressed(self._h, row, points48, len(points48) // 48, scale_point48))

    def srs_export_row_compressed(self, row: int, n: int) -> bytes:
        out = ctypes.create_string_buffer(48 * n)
        check(lib().zkp_srs_export_row_compressed(self._h, row, out, n))
        return out.raw

    def srs_import_g2(self, which: int, g2_192: bytes) -> None:
        check(lib().zkp_srs_import_g2(self._h, which, g2_192))

    def srs_export_g2(self, which: int) -> bytes:
        out = ctypes.create_string_buffer(192)
        check(lib().zkp_srs_export_g2(self._h, which, out))
        return out.raw

    def srs_export_scale_point(self, row: int) -> bytes:
        out = ctypes.create_string_buffer(48)
        check(lib().zkp_srs_export_scale_point(self._h, row, out))
        return out.raw

    def srs_set_shard(self, log_domain: int, shard: int) -> None:
        check(lib().zkp_srs_set_shard(self._h, log_domain, shard))

    def srs_generate_monomial2(self, tau_x: int, tau_y: int, log_n: int, log_machines: int) -> None:
        check(lib().zkp_srs_generate_monomial2(self._h, tau_x.to_bytes(32, "big"), tau_y.to_bytes(32, "big"), log_n, log_machines))

    def srs_monomial_to_lagrange(self) -> None:
        """monomial rows [tau_x^j tau_y^i]_1 -> Lagrange rows + scale points, in place, without the trapdoor"""
        check(lib().zkp_srs_monomial_to_lagrange(self._h))

    def srs_save(self, path: str) -> None:
        check(lib().zkp_srs_save(self._h, path.encode()))

    def srs_load(self, path: str) -> None:
        check(lib().zkp_srs_load(self._h, path.encode()))

    def srs_shape(self) -> Tuple[int, int]:
        a, b = ctypes.c_uint32(), ctypes.c_uint32()
        check(lib().zkp_srs_shape(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    # ---- hot path (bytes in, bytes out)
    def worker_commit(self, i: int, poly_be: bytes) -> bytes:
        out = ctypes.create_string_buffer(48)
        check(lib().zkp_worker_commit(self._h, i, _arg(poly_be), len(poly_be) // 32, out))
        return out.raw

    def worker_open(self, i: int, poly_be: bytes, x_be: bytes) -> Tuple[bytes, bytes]:
        y = ctypes.create_string_buffer(32)
        proof = ctypes.create_string_buffer(48)
        check(lib().zkp_worker_open(self._h, i, _arg(poly_be), len(poly_be) // 32, x_be, y, proof))
        return y.raw, proof.raw

    def worker_open_resident(self, i: int, n: int, x_be: bytes) -> Tuple[bytes, bytes]:
        """Opening of the polynomial the previous worker_commit / worker_open / worker_commit_open call left on the
        device (no upload); ZkpError(ZKP_ERR_STATE) when none of n elements is resident."""
        y = ctypes.create_string_buffer(32)
        proof = ctypes.create_string_buffer(48)
        check(lib().zkp_worker_open_resident(self._h, i, n, x_be, y, proof))
        return y.raw, proof.raw

    def last_points_uncompressed(self) -> bytes:
        """commitment || proof of the last commit+open, 2 x 96 bytes uncompressed (cross-process combine)"""
        out = ctypes.create_string_buffer(192)
        check(lib().zkp_last_points_uncompressed(self._h, out))
        return out.raw

    def stage_list(self, strs, staging: "PinnedBuffer", chunk: int = 1 << 18) -> int:
        """Decode a List[str] of base64 field elements into `staging` (page-locked) in chunks, every finished chunk being
        copied to the device while the next one is decoded; returns the generation of the upload, to be handed to
        worker_commit_resident / worker_open_resident_gen / worker_commit_open_resident.  ValueError on a malformed element."""
        if not isinstance(strs, (list, tuple)):
            strs = list(strs)
        n = len(strs)
        if staging.capacity < 32 * n:
            raise ValueError("PinnedBuffer too small")
        gen = ctypes.c_uint64()
        check(lib().zkp_stage_begin(self._h, n, ctypes.byref(gen)))
        cb = ctypes.cast(lib().zkp_stage_chunk, ctypes.c_void_p)
        rc = wire().zkp_wire_decode_list_chunked(strs, ctypes.addressof(staging.buf), staging.capacity, chunk, cb, self._h)
        if rc != n:
            if rc <= -(1 << 40):
                raise ValueError("wire decode: bad argument" if rc > -(1 << 40) - 2 else f"staged upload failed: {last_error()}")
            raise ValueError(f"element {-1 - rc} is not a base64 field element of 43/44 characters")
        check(lib().zkp_stage_end(self._h, gen.value))
        staging.used = 32 * n
        return gen.value

    def worker_commit_resident(self, i: int, n: int, generation: int) -> bytes:
        out = ctypes.create_string_buffer(48)
        check(lib().zkp_worker_commit_resident(self._h, i, n, generation, out))
        return out.raw

    def worker_commit_open_resident(self, i: int, n: int, generation: int, x_be: bytes) -> Tuple[bytes, bytes, bytes]:
        com = ctypes.create_string_buffer(48)
        y = ctypes.create_string_buffer(32)
        proof = ctypes.create_string_buffer(48)
        check(lib().zkp_worker_commit_open_resident(self._h, i, n, generation, x_be, com, y, proof))
        return com.raw, y.raw, proof.raw

    def last_points_jacobian(self) -> bytes:
        """commitment || proof of the last commit+open as Jacobian points (2 x 144 bytes, internal limbs; no inversion)"""
        out = ctypes.create_string_buffer(288)
        check(lib().zkp_last_points_jacobian(self._h, out))
        return out.raw

    def resident_generation(self) -> Tuple[int, int]:
        """(generation, n) of the polynomial the last upload left on the device; see worker_open_resident_gen."""
        g, n = ctypes.c_uint64(), ctypes.c_size_t()
        check(lib().zkp_resident_generation(self._h, ctypes.byref(g), ctypes.byref(n)))
        return g.value, n.value

    def worker_open_resident_gen(self, i: int, n: int, generation: int, x_be: bytes) -> Tuple[bytes, bytes]:
        """worker_open_resident bound to one upload: ZkpError(ZKP_ERR_STATE) if anything has rewritten the staged
        polynomial since `generation` was read."""
        y = ctypes.create_string_buffer(32)
        proof = ctypes.create_string_buffer(48)
        check(lib().zkp_worker_open_resident_gen(self._h, i, n, generation, x_be, y, proof))
        return y.raw, proof.raw

    def worker_commit_open_batch(self, rows, polys, xs_be: bytes):
        """`len(rows)` commit+open requests in one launch set.  polys: bytes or PinnedBuffer objects of n x 32 bytes
        each; xs_be: 32 bytes per request.  Returns [(status, commitment, eval, proof)] (status 0 = OK)."""
        k = len(rows)
        n = len(polys[0]) // 32
        keep = [_arg(p) for p in polys]
        ptrs = (ctypes.c_void_p * k)(*[ctypes.cast(ctypes.c_char_p(b) if isinstance(b, bytes) else b, ctypes.c_void_p) for b in keep])
        idx = (ctypes.c_uint32 * k)(*rows)
        coms = ctypes.create_string_buffer(48 * k)
        ys = ctypes.create_string_buffer(32 * k)
        proofs = ctypes.create_string_buffer(48 * k)
        status = (ctypes.c_int * k)()
        check(lib().zkp_worker_commit_open_batch(self._h, k, idx, ptrs, n, xs_be, coms, ys, proofs, status))
        return [(status[r], coms.raw[48 * r:48 * r + 48], ys.raw[32 * r:32 * r + 32], proofs.raw[48 * r:48 * r + 48]) for r in range(k)]

    def worker_commit_open(self, i: int, poly_be: bytes, x_be: bytes) -> Tuple[bytes, bytes, bytes]:
        com = ctypes.create_string_buffer(48)
        y = ctypes.create_string_buffer(32)
        proof = ctypes.create_string_buffer(48)
        check(lib().zkp_worker_commit_open(self._h, i, _arg(poly_be), len(poly_be) // 32, x_be, com, y, proof))
        return com.raw, y.raw, proof.raw

    def worker_verify(self, i: int, proof48: bytes, alpha_be: bytes, eval_be: bytes, commitment48: bytes) -> bool:
        valid = ctypes.c_int(0)
        check(lib().zkp_worker_verify(self._h, i, proof48, alpha_be, eval_be, commitment48, ctypes.byref(valid)))
        return bool(valid.value)

    def worker_verify_batch(self, indices, proofs48: bytes, alpha_be: bytes, evals_be: bytes, commitments48: bytes):
        """Verify the responses of one challenge together; returns a list of bools (same answers as worker_verify)."""
        n = len(indices)
        idx = (ctypes.c_uint32 * n)(*indices)
        valid = (ctypes.c_int * n)()
        check(lib().zkp_worker_verify_batch(self._h, n, idx, proofs48, alpha_be, evals_be, commitments48, valid))
        return [bool(v) for v in valid]

    # ---- opening split by point range over several GPUs (see include/zkp_b200.h)
    def shard_eval_partial(self, i: int, slice_be: bytes, x_be: bytes) -> bytes:
        out = ctypes.create_string_buffer(32)
        check(lib().zkp_shard_eval_partial(self._h, i, _arg(slice_be), len(slice_be) // 32, x_be, out))
        return out.raw

    def shard_open_partial(self, i: int, slice_be: bytes, x_be: bytes, y_be: bytes) -> bytes:
        out = ctypes.create_string_buffer(48)
        check(lib().zkp_shard_open_partial(self._h, i, _arg(slice_be), len(slice_be) // 32, x_be, y_be, out))
        return out.raw

    # ---- Pianist master node (aggregation is g1_sum; the Y-direction opening and the bivariate check are here)
    def master_open_y(self, worker_evals_be: bytes, beta_be: bytes) -> Tuple[bytes, bytes]:
        z = ctypes.create_string_buffer(32)
        proof = ctypes.create_string_buffer(48)
        check(lib().zkp_master_open_y(self._h, worker_evals_be, len(worker_evals_be) // 32, beta_be, z, proof))
        return z.raw, proof.raw

    def master_verify(self, commitment48: bytes, proof_x48: bytes, proof_y48: bytes, alpha_be: bytes, beta_be: bytes,
                      z_be: bytes) -> bool:
        valid = ctypes.c_int(0)
        check(lib().zkp_master_verify(self._h, commitment48, proof_x48, proof_y48, alpha_be, beta_be, z_be, ctypes.byref(valid)))
        return bool(valid.value)

    def fft(self, vals_be: bytes, left: bool = True, inverse: bool = False) -> bytes:
        out = ctypes.create_string_buffer(len(vals_be))
        check(lib().zkp_fft(self._h, _arg(vals_be), len(vals_be) // 32, int(left), int(inverse), out))
        return out.raw

    def eval(self, coeffs_be: bytes, x_be: bytes) -> bytes:
        out = ctypes.create_string_buffer(32)
        check(lib().zkp_eval(self._h, _arg(coeffs_be), len(coeffs_be) // 32, x_be, out))
        return out.raw

    def challenge_evals(self, polys_be: bytes, rows: int, alpha_be: bytes) -> bytes:
        """f_i(alpha) for `rows` concatenated rows of evaluations (rows x 32 bytes out)."""
        out = ctypes.create_string_buffer(32 * rows)
        check(lib().zkp_challenge_evals(self._h, _arg(polys_be), rows, len(polys_be) // 32 // rows, alpha_be, out))
        return out.raw

    def random_poly(self, seed: int, count: int) -> bytes:
        out = ctypes.create_string_buffer(32 * count)
        check(lib().zkp_random_poly(self._h, seed, out, count))
        return out.raw

    def random_point(self, seed: int) -> bytes:
        out = ctypes.create_string_buffer(32)
        check(lib().zkp_random_point(self._h, seed, out))
        return out.raw

    def random_poly_range(self, seed: int, first: int, count: int) -> bytes:
        """elements [first, first + count) of the stream random_poly(seed, ...) yields"""
        out = ctypes.create_string_buffer(32 * count)
        check(lib().zkp_random_poly_range(self._h, seed, first, out, count))
        return out.raw

    def msm_g1(self, row: int, scalars_be: bytes) -> bytes:
        out = ctypes.create_string_buffer(48)
        check(lib().zkp_msm_g1(self._h, row, _arg(scalars_be), len(scalars_be) // 32, out))
        return out.raw

    # ---- bench / tuning
    def set_msm_window(self, c: int) -> None:
        check(lib().zkp_set_msm_window(self._h, c))

    def set_msm_sort(self, mode) -> None:
        """0 / False: cub::DeviceRadixSort, 1 / True: hand-written bucket sort, 2: by size (default)"""
        check(lib().zkp_set_msm_sort(self._h, int(mode)))

    def set_msm_mode(self, fixed_base_tables: bool) -> None:
        check(lib().zkp_set_msm_mode(self._h, int(fixed_base_tables)))

    def set_poly_form(self, coefficients: bool) -> None:
        """worker_* polynomials are evaluations (False, default) or coefficients (True)"""
        check(lib().zkp_set_poly_form(self._h, int(coefficients)))

    def set_ntt_tma(self, on: bool) -> None:
        """NTT pass-2 tile by one bulk async copy (TMA) instead of per-thread loads (experiment; same results)"""
        check(lib().zkp_set_ntt_tma(self._h, int(on)))

    def set_fuse(self, mode: int) -> None:
        """commit+open as one grouped launch set: 1 always, 0 never, -1 by row length (default)"""
        check(lib().zkp_set_fuse(self._h, mode))

    def set_rowcol_coop(self, on: bool) -> None:
        """row / column sums over a small bucket array with four lanes per share (default) or the large-array kernels"""
        check(lib().zkp_set_rowcol_coop(self._h, int(on)))

    def set_open_coset(self, on: bool) -> None:
        """single-request opening on cosets with the inversion on the host (default) or the general device form"""
        check(lib().zkp_set_open_coset(self._h, int(on)))

    def prebuild_tables(self, first_row: int = 0, count: int = 1 << 30) -> int:
        built = ctypes.c_uint32()
        check(lib().zkp_srs_prebuild_tables(self._h, first_row, min(count, 0xFFFFFFFF), ctypes.byref(built)))
        return built.value

    def set_table_budget(self, nbytes: int) -> None:
        check(lib().zkp_set_table_budget(self._h, nbytes))

    def table_stats(self) -> dict:
        out = (ctypes.c_uint64 * 6)()
        check(lib().zkp_srs_table_stats(self._h, out))
        return dict(zip(("resident", "slots", "arena_bytes", "builds", "evictions", "fallbacks"), [int(v) for v in out]))

    def set_msm_affine_rounds(self, rounds: int) -> None:
        check(lib().zkp_set_msm_affine_rounds(self._h, rounds))

    def msm_info(self, n: int) -> Tuple[int, int, int]:
        c, w, m = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint64()
        check(lib().zkp_msm_info(self._h, n, ctypes.byref(c), ctypes.byref(w), ctypes.byref(m)))
        return c.value, w.value, m.value

    def bench_msm(self, row: int, scalars_be: bytes, reps: int, flush_l2: bool = True) -> Tuple[float, bytes]:
        ms = ctypes.c_float()
        out = ctypes.create_string_buffer(48)
        check(lib().zkp_bench_msm(self._h, row, _arg(scalars_be), len(scalars_be) // 32, reps, int(flush_l2), ctypes.byref(ms), out))
        return ms.value, out.raw

    def bench_commit_open(self, row: int, poly_be: bytes, x_be: bytes, reps: int, flush_l2: bool = True):
        ms, ms_k, launches = ctypes.c_float(), ctypes.c_float(), ctypes.c_uint32()
        com = ctypes.create_string_buffer(48)
        y = ctypes.create_string_buffer(32)
        proof = ctypes.create_string_buffer(48)
        check(lib().zkp_bench_commit_open(self._h, row, _arg(poly_be), len(poly_be) // 32, x_be, reps, int(flush_l2),
                                          ctypes.byref(ms), ctypes.byref(ms_k), ctypes.byref(launches), com, y, proof))
        return ms.value, ms_k.value, launches.value, com.raw, y.raw, proof.raw

    def bench_trace(self, row: int, poly_be: bytes, x_be: bytes, warm: int = 2):
        """[(lane, stage, ms since request start)] of one commit+open, plus ("host", "total", ms)."""
        buf = ctypes.create_string_buffer(8192)
        check(lib().zkp_bench_trace(self._h, row, _arg(poly_be), len(poly_be) // 32, x_be, warm, buf, len(buf)))
        rows = []
        for line in buf.value.decode().splitlines():
            a, b, c = line.split()
            rows.append((a, b, float(c)))
        return rows

    def bench_flush_l2(self) -> None:
        check(lib().zkp_bench_flush_l2(self._h))

    def bench_last_kernel_ms(self) -> float:
        ms = ctypes.c_float()
        check(lib().zkp_bench_last_kernel_ms(self._h, ctypes.byref(ms)))
        return ms.value

    def bench_peaks(self) -> Tuple[float, float]:
        a, b = ctypes.c_double(), ctypes.c_double()
        check(lib().zkp_bench_peaks(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def bench_ntt(self, n: int, reps: int, inverse: bool = False) -> float:
        ms = ctypes.c_float()
        check(lib().zkp_bench_ntt(self._h, n, reps, int(inverse), ctypes.byref(ms)))
        return ms.value


class MultiContext:
    """zkp_mgpu: the GPUs of one box behind one handle (one context + one host thread per device inside the library;
    no torch, no NCCL).  layout LAYOUT_ROWS: whole SRS on every device, sub-polynomial k on device k mod G (Pianist);
    LAYOUT_POINT_RANGE: one polynomial split by point range."""

    def __init__(self, devices=None):
        if devices is None:
            devices = list(range(lib().zkp_device_count()))
        self.devices = list(devices)
        arr = (ctypes.c_int * len(self.devices))(*self.devices)
        self._h = _ctxp()
        check(lib().zkp_mgpu_create(arr, len(self.devices), ctypes.byref(self._h)))

    def close(self) -> None:
        if self._h:
            lib().zkp_mgpu_destroy(self._h)
            self._h = _ctxp()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ctx(self, k: int) -> Context:
        """the context of device k, borrowed (owned by this MultiContext)"""
        h = lib().zkp_mgpu_ctx(self._h, k)
        if not h:
            raise ZkpError(ZKP_ERR_ARG, "no such device in the MultiContext")
        return Context(self.devices[k], _handle=_ctxp(h), _borrowed=True)

    def set_layout(self, layout: int, log_n: int, log_machines: int) -> None:
        check(lib().zkp_mgpu_set_layout(self._h, layout, log_n, log_machines))

    def srs_generate(self, tau_x: int, tau_y: int, log_n: int, log_machines: int, layout: int) -> None:
        check(lib().zkp_mgpu_srs_generate(self._h, tau_x.to_bytes(32, "big"), tau_y.to_bytes(32, "big"), log_n, log_machines, layout))

    def prebuild_tables(self) -> None:
        check(lib().zkp_mgpu_prebuild_tables(self._h))

    def msm_g1(self, row: int, scalars_be, flags: int = 0) -> bytes:
        out = ctypes.create_string_buffer(48)
        check(lib().zkp_mgpu_msm_g1(self._h, row, _arg(scalars_be), len(scalars_be) // 32, flags, out))
        return out.raw

    def commit_open(self, row: int, poly_be, x_be: bytes, flags: int = 0) -> Tuple[bytes, bytes, bytes]:
        com = ctypes.create_string_buffer(48)
        y = ctypes.create_string_buffer(32)
        proof = ctypes.create_string_buffer(48)
        check(lib().zkp_mgpu_commit_open(self._h, row, _arg(poly_be), len(poly_be) // 32, x_be, flags, com, y, proof))
        return com.raw, y.raw, proof.raw

    def pianist_commit_open(self, rows, polys_be, alpha_be: bytes, flags: int = 0):
        """-> (commitments [48 B each], evals [32 B each], proofs [48 B each], aggregated commitment, aggregated proof)"""
        k = len(rows)
        n = len(polys_be) // 32 // k
        idx = (ctypes.c_uint32 * k)(*rows)
        coms = ctypes.create_string_buffer(48 * k)
        ys = ctypes.create_string_buffer(32 * k)
        proofs = ctypes.create_string_buffer(48 * k)
        agg_c = ctypes.create_string_buffer(48)
        agg_p = ctypes.create_string_buffer(48)
        check(lib().zkp_mgpu_pianist_commit_open(self._h, idx, k, _arg(polys_be), n, alpha_be, flags, coms, ys, proofs, agg_c, agg_p))
        split = lambda b, w: [b[w * r:w * r + w] for r in range(k)]
        return split(coms.raw, 48), split(ys.raw, 32), split(proofs.raw, 48), agg_c.raw, agg_p.raw


def b64_decode_fr(strs: bytes, stride: int, count: int, out: Optional[PinnedBuffer] = None):
    """Batch-decode `count` base64 field elements; into `out` (a PinnedBuffer, returned) when given."""
    if out is not None:
        if out.capacity < 32 * count:
            raise ValueError("PinnedBuffer too small")
        check(lib().zkp_b64_decode_fr(strs, stride, count, out.buf))
        out.used = 32 * count
        return out
    buf = ctypes.create_string_buffer(32 * count)
    check(lib().zkp_b64_decode_fr(strs, stride, count, buf))
    return buf.raw


def b64_encode_fr(vals_be: bytes) -> bytes:
    count = len(vals_be) // 32
    out = ctypes.create_string_buffer(43 * count)
    check(lib().zkp_b64_encode_fr(vals_be, count, out))
    return out.raw


# ---- CPython-side wire codec (zkp_subnet_b200/_zkp_wire.so, csrc/wire_py.cpp): List[str] <-> bytes without a
#      Python-level join/split; PyDLL keeps the GIL for the call, the decode itself runs on host threads
WIRE_PATH = os.path.join(_HERE, "_zkp_wire.so")
_wire = None


def wire() -> ctypes.PyDLL:
    global _wire
    if _wire is None:
        if not os.path.exists(WIRE_PATH):
            raise ZkpError(ZKP_ERR_STATE, f"{WIRE_PATH} not built; run `make`")
        h = ctypes.PyDLL(WIRE_PATH)
        h.zkp_wire_decode_list.argtypes = [ctypes.py_object, ctypes.c_void_p, ctypes.c_size_t]
        h.zkp_wire_decode_list.restype = ctypes.c_longlong
        h.zkp_wire_decode_list_cmp.argtypes = [ctypes.py_object, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                               ctypes.POINTER(ctypes.c_int)]
        h.zkp_wire_decode_list_cmp.restype = ctypes.c_longlong
        h.zkp_wire_decode_list_chunked.argtypes = [ctypes.py_object, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                                   ctypes.c_void_p, ctypes.c_void_p]
        h.zkp_wire_decode_list_chunked.restype = ctypes.c_longlong
        h.zkp_wire_encode_list.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        h.zkp_wire_encode_list.restype = ctypes.py_object
        _wire = h
    return _wire


def wire_decode_list(strs, out: Optional[PinnedBuffer] = None, same_as: Optional[PinnedBuffer] = None):
    """List[str] (43/44-char base64 field elements) -> n x 32 bytes, into `out` (returned) when given.
    With `same_as` (a buffer holding n x 32 bytes) the result is the pair (out, decoded bytes == same_as)."""
    if not isinstance(strs, (list, tuple)):
        strs = list(strs)
    n = len(strs)
    if same_as is not None:
        if out is None or out.capacity < 32 * n or same_as.capacity < 32 * n:
            raise ValueError("PinnedBuffer too small")
        same = ctypes.c_int(0)
        rc = wire().zkp_wire_decode_list_cmp(strs, ctypes.addressof(out.buf), out.capacity, ctypes.addressof(same_as.buf),
                                             ctypes.byref(same))
        if rc == n:
            out.used = 32 * n
            return out, bool(same.value)
    elif out is not None:
        if out.capacity < 32 * n:
            raise ValueError("PinnedBuffer too small")
        rc = wire().zkp_wire_decode_list(strs, ctypes.addressof(out.buf), out.capacity)
    else:
        buf = ctypes.create_string_buffer(32 * n)
        rc = wire().zkp_wire_decode_list(strs, ctypes.addressof(buf), 32 * n)
    if rc != n:
        if rc <= -(1 << 40):
            raise ValueError("wire decode: bad argument")
        raise ValueError(f"element {-1 - rc} is not a base64 field element of 43/44 characters")
    if out is not None:
        out.used = 32 * n
        return out
    return buf.raw


def wire_encode_list(vals_be: bytes):
    """n x 32 bytes -> list of n unpadded base64 strings (43 chars each)."""
    return wire().zkp_wire_encode_list(vals_be, len(vals_be) // 32)


def shard_eval_combine(partials_be: bytes, log_n: int, x_be: bytes) -> bytes:
    out = ctypes.create_string_buffer(32)
    check(lib().zkp_shard_eval_combine(partials_be, len(partials_be) // 32, log_n, x_be, out))
    return out.raw


def g1_sum(points48: bytes) -> bytes:
    out = ctypes.create_string_buffer(48)
    check(lib().zkp_g1_sum(points48, len(points48) // 48, out))
    return out.raw


def g1_sum_checked(points48: bytes) -> bytes:
    """sum of points received from other parties: each is checked to be in the prime-order subgroup"""
    out = ctypes.create_string_buffer(48)
    check(lib().zkp_g1_sum_checked(points48, len(points48) // 48, out))
    return out.raw


def g1_sum_jacobian(points: bytes, count: int, stride: int) -> bytes:
    """sum of `count` Jacobian points (records of `stride` >= 144 bytes, see Context.last_points_jacobian) -> compressed"""
    out = ctypes.create_string_buffer(48)
    check(lib().zkp_g1_sum_jacobian(points, count, stride, out))
    return out.raw


def g1_uncompress(point48: bytes) -> bytes:
    out = ctypes.create_string_buffer(96)
    check(lib().zkp_g1_uncompress(point48, out))
    return out.raw


def g1_sum_uncompressed(points96: bytes) -> bytes:
    out = ctypes.create_string_buffer(48)
    check(lib().zkp_g1_sum_uncompressed(points96, len(points96) // 96, out))
    return out.raw


def pairing_check(g1_48: bytes, g2_192: bytes) -> bool:
    pairs = len(g1_48) // 48
    res = ctypes.c_int(0)
    check(lib().zkp_pairing_check(g1_48, g2_192, pairs, ctypes.byref(res)))
    return bool(res.value)
