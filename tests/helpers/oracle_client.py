"""Test double: `fourier.Client` with the GPU context replaced by the CPU oracle.

Everything the reference's neurons see -- the constructor keywords, the ten methods, the Response objects, the wire
codec (csrc/wire_py.cpp), the error conventions -- is the product's `zkp_subnet_b200.client.Client`; only the object
behind `native.Context` is swapped for `OracleContext`, which answers the same calls from oracle/ (C restatement +
big-int pairing).  That lets the UNMODIFIED reference classes (/root/reference, absent on the GPU box) be driven in this
GPU-less container; the bytes they obtain are pinned to tests/golden/vectors.json, which the `-m gpu` tests pin the
CUDA path to as well.  Test infrastructure only (the oracle never backs the product).
"""
from __future__ import annotations

from oracle import bls12_381 as o
from oracle import ref
from zkp_subnet_b200 import native
from zkp_subnet_b200.client import Client

R = o.R


def _canon(buf: bytes) -> None:
    if len(buf) % 32:
        raise native.ZkpError(native.ZKP_ERR_ARG, "length not a multiple of 32")
    for i in range(0, len(buf), 32):
        if int.from_bytes(buf[i:i + 32], "big") >= R:
            raise native.ZkpError(native.ZKP_ERR_ENCODING, "non-canonical field element")


class OracleContext:
    def __init__(self, device: int = 0):
        self.rows = None

    # ---- SRS
    def srs_generate(self, tau_x: int, tau_y: int, log_n: int, log_m: int) -> None:
        self.log_n, self.log_m, self.tau_x, self.tau_y = log_n, log_m, tau_x, tau_y
        n, m = 1 << log_n, 1 << log_m
        Rs = ref.split32(ref.lagrange_scalars(m, tau_y)) if m > 1 else [1]
        self.rows = [ref.srs(n, tau_x, "lagrange", scale=Rs[i]) for i in range(m)]
        self.scale = [o.g1_mul(o.G1_GEN, Rs[i]) for i in range(m)]
        self.tau_g2 = o.g2_mul(o.G2_GEN, tau_x)

    def srs_shape(self):
        return self.log_n, self.log_m

    def prebuild_tables(self, *a):
        return 0

    def set_poly_form(self, coefficients: bool) -> None:
        self.coeffs = bool(coefficients)

    def fork(self):
        return self

    def close(self) -> None:
        pass

    def resident_generation(self):
        return 0, 0

    def _row(self, i: int, n: int, exact: bool = False):
        if self.rows is None:
            raise native.ZkpError(native.ZKP_ERR_STATE, "SRS not loaded")
        if not 0 <= i < len(self.rows):
            raise native.ZkpError(native.ZKP_ERR_ARG, "worker index out of range")
        if n == 0 or n > (1 << self.log_n) or (exact and n != (1 << self.log_n)):
            raise native.ZkpError(native.ZKP_ERR_ARG, "polynomial length does not fit the SRS row")
        return self.rows[i]

    def _evals(self, poly: bytes) -> bytes:
        _canon(poly)
        return ref.ntt(poly, False) if getattr(self, "coeffs", False) else poly

    # ---- hot path
    def worker_commit(self, i: int, poly: bytes) -> bytes:
        poly = bytes(poly)
        return ref.msm(self._row(i, len(poly) // 32), self._evals(poly), 4)

    def worker_open(self, i: int, poly: bytes, x: bytes):
        poly = bytes(poly)
        _canon(x)
        return ref.open_evals(self._evals(poly), x, self._row(i, len(poly) // 32, True), 4)

    def worker_commit_open(self, i: int, poly: bytes, x: bytes):
        return (self.worker_commit(i, poly),) + tuple(self.worker_open(i, poly, x))

    def worker_verify(self, i: int, proof: bytes, alpha: bytes, y: bytes, commitment: bytes) -> bool:
        if not 0 <= i < len(self.rows):
            raise native.ZkpError(native.ZKP_ERR_ARG, "worker index out of range")
        try:
            pi, com = o.g1_decompress(proof), o.g1_decompress(commitment)
        except (ValueError, AssertionError):
            return False
        a, yy = int.from_bytes(alpha, "big"), int.from_bytes(y, "big")
        if a >= R or yy >= R:
            return False
        return o.kzg_verify(com, pi, a, yy, self.tau_g2, self.scale[i])

    def worker_verify_batch(self, indices, proofs: bytes, alpha: bytes, evals: bytes, commitments: bytes):
        return [self.worker_verify(i, proofs[48 * k:48 * k + 48], alpha, evals[32 * k:32 * k + 32], commitments[48 * k:48 * k + 48])
                for k, i in enumerate(indices)]

    def fft(self, vals: bytes, left: bool = True, inverse: bool = False) -> bytes:
        vals = bytes(vals)
        _canon(vals)
        return ref.ntt(vals, inverse)

    def eval(self, coeffs: bytes, x: bytes) -> bytes:
        coeffs = bytes(coeffs)
        _canon(coeffs + x)
        return ref.eval_coeffs(coeffs, x)

    def challenge_evals(self, polys: bytes, rows: int, alpha: bytes) -> bytes:
        polys = bytes(polys)
        n = len(polys) // 32 // rows
        return b"".join(ref.quotient_evals(polys[32 * n * r:32 * n * (r + 1)], alpha)[0] for r in range(rows))

    def random_poly(self, seed: int, count: int) -> bytes:
        return ref.random_scalars_ctr(seed, 0, count)

    def random_point(self, seed: int) -> bytes:
        return ref.random_scalars_ctr(seed ^ 0x706F696E74, 0, 1)


class OracleClient(Client):
    """fourier.Client whose device contexts are OracleContext objects"""

    def __init__(self, *a, **k):
        k.setdefault("test_srs", True)
        k.setdefault("contexts", 1)
        super().__init__(*a, **k)
        self._pinned = False

    def _make_root(self, device: int):
        return OracleContext(device)
