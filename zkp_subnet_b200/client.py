"""Drop-in replacement for `fourier.Client` -- the only object through which the reference's miner and
validator reach their prover (reference base/miner.py:26,73-84; base/validator.py:28,80-91).

Same constructor keywords, same ten methods, same return convention: every call returns an object
usable as `with client.method(...) as response:` exposing `response.status_code` (200 = OK) and
`response.json()` (reference neurons/miner.py:38-54, neurons/validator.py:58-104).  Instead of
HTTP -> Rust process, the methods call libzkp_b200.so (CUDA, sm_100a) through ctypes.  There is no
CPU fallback: `start()` raises if the library or a GPU is missing.

Wire format (reference base/protocol.py:35-60, tests/test_miner.py:33-55): field elements are
unpadded standard base64 of 32 big-endian bytes; G1 points are base64 of the 48-byte ZCash
compressed encoding.

Concurrency (reference base/miner.py:66-70 hands `forward` to the axon's threads): the client keeps a POOL of
contexts -- `contexts` per device, all sharing one resident SRS and one set of fixed-base tables per device
(zkp_ctx_fork) -- and every call borrows one, so concurrent requests overlap on the GPU instead of queueing on a lock.
With `devices=[...]` the pool spans several GPUs (each holds the SRS); `multi_gpu="split"` instead splits EVERY
request by point range over the GPUs (zkp_mgpu_commit_open).
"""
from __future__ import annotations

import base64
import os
import queue
import secrets
import sys
import threading
from typing import Any, Dict, List, Optional, Sequence

from . import native, srsfile

# Public test trapdoors for a throw-away SRS (the reference's tests also generate one: tests/conftest.py:50-65).
# Anyone can forge openings against an SRS whose trapdoor is known, so it is only ever used behind an explicit
# opt-in (Client(test_srs=True) or ZKP_B200_TEST_SRS=1), it is announced on stderr, and it is never written to
# `setup_path`.  A deployment loads the ceremony files.
TEST_TAU_X = 1927409816240961209460912649124
TEST_TAU_Y = 0x1234567890ABCDEF1234567890ABCDEF


class Response:
    """Mimics the slice of `requests.Response` the reference touches."""

    def __init__(self, status_code: int, payload: Dict[str, Any]):
        self.status_code = status_code
        self._payload = payload

    def json(self) -> Dict[str, Any]:
        return self._payload

    def __enter__(self) -> "Response":
        return self

    def __exit__(self, *exc) -> bool:
        return False


def _b64_point(raw: bytes) -> str:
    return base64.b64encode(raw).decode()


def _b64_fr(raw: bytes) -> str:
    return base64.b64encode(raw).decode().rstrip("=")


def _decode_any(s: str, size: int) -> bytes:
    raw = base64.b64decode(s + "=" * (-len(s) % 4), validate=True)
    if len(raw) != size:
        raise ValueError(f"expected {size} bytes, got {len(raw)}")
    return raw


def decode_poly(poly: Sequence[str], staging: Optional[native.PinnedBuffer] = None):
    """List[str] -> n x 32 bytes in one native call (csrc/wire_py.cpp walks the list through the C API and
    decodes on host threads; no Python-level join).  With a `staging` buffer the bytes land in page-locked
    memory, ready for an asynchronous upload.  Raises ValueError on a malformed element."""
    if len(poly) == 0:
        return b""
    return native.wire_decode_list(poly, staging)


def encode_poly(raw: bytes) -> List[str]:
    return native.wire_encode_list(raw)


def _bitrev(i: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def _truthy(v) -> bool:
    # the reference declares --uncompressed with type=bool, so `--uncompressed true` arrives as a non-empty string
    # (utils/config.py:131-136); accept both
    if isinstance(v, str):
        return v.strip().lower() not in ("", "0", "false", "no", "off")
    return bool(v)


class _Helper:
    """A persistent helper thread: submit(fn) runs fn() there and returns an Event that is set when it has finished
    (fn reports through its own closure).  Daemon, so a client that is never stop()ped does not keep the process alive."""

    def __init__(self):
        self._q: "queue.SimpleQueue" = queue.SimpleQueue()
        self._t = threading.Thread(target=self._loop, name="zkp-b200-helper", daemon=True)
        self._t.start()

    def _loop(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            fn, done = item
            try:
                fn()
            except BaseException:  # fn reports through its own closure; the helper must outlive a failing job
                pass
            finally:
                done.set()

    def submit(self, fn) -> threading.Event:
        done = threading.Event()
        self._q.put((fn, done))
        return done

    def close(self):
        self._q.put(None)


class _Slot:
    """One pooled context with its page-locked staging buffers and what it remembers about the resident polynomial."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.staging: Optional[native.PinnedBuffer] = None
        self.staging_alt: Optional[native.PinnedBuffer] = None  # second buffer of the speculative worker_open
        self.resident_n = 0      # > 0: self.staging holds the n x 32 bytes of the polynomial this slot uploaded last
        self.resident_gen = -1   # ... and this is the library's generation of that upload
        self._helper: Optional[_Helper] = None

    def helper(self) -> "_Helper":
        if self._helper is None:
            self._helper = _Helper()
        return self._helper

    def close(self, close_ctx: bool = True):
        self.resident_n = 0
        if self._helper is not None:
            self._helper.close()
            self._helper = None
        for name in ("staging", "staging_alt"):
            if getattr(self, name) is not None:
                getattr(self, name).close()
                setattr(self, name, None)
        if close_ctx and self.ctx is not None:
            self.ctx.close()
        self.ctx = None


class Client:
    """`Client(port=..., bin=..., uncompressed=..., setup_path=..., precompute_path=...)` (reference base/miner.py:75-81).

    `port` and `bin` are accepted for signature compatibility and ignored (there is no prover process).
    `setup_path` / `precompute_path` / `uncompressed` name the SRS files exactly as the reference's flags do
    (Makefile:64-74): see zkp_subnet_b200/srsfile.py for the formats read.  A missing file is an ERROR unless the
    throw-away test SRS has been asked for explicitly (`test_srs=True` or ZKP_B200_TEST_SRS=1).

    Extensions (keyword-only in spirit; the reference never passes them):
      device / devices   GPU(s) to use; with several devices every one holds the SRS
      contexts           pooled contexts per device (concurrent requests overlap on the GPU)
      precompute         "eager": build the fixed-base tables in start(); "lazy": inside the first request of a row
      multi_gpu          "requests": each request runs on one GPU of the pool; "split": every request is split by
                         point range over all GPUs (zkp_mgpu_commit_open)
      poly_form          "evals" (default) or "coeffs": what the `poly` of worker_commit / worker_open holds
                         (reference neurons/validator.py:67 says "coefficients", its own flow
                         implies evaluations)
      row_order          "natural" (default) or "bitrev": which SRS row worker index i uses, R_i or R_bitrev(i)
    """

    def __init__(self, port: int = 1337, bin: Optional[str] = None, uncompressed: bool = False,
                 setup_path: Optional[str] = None, precompute_path: Optional[str] = None, device: int = 0,
                 seed: Optional[int] = None, devices: Optional[Sequence[int]] = None, contexts: int = 2,
                 precompute: str = "eager", multi_gpu: str = "requests", poly_form: str = "evals",
                 row_order: str = "natural", test_srs: Optional[bool] = None, staged_upload: Optional[bool] = None):
        if precompute not in ("eager", "lazy"):
            raise ValueError("precompute must be 'eager' or 'lazy'")
        if multi_gpu not in ("requests", "split"):
            raise ValueError("multi_gpu must be 'requests' or 'split'")
        if poly_form not in ("evals", "coeffs"):
            raise ValueError("poly_form must be 'evals' or 'coeffs'")
        if row_order not in ("natural", "bitrev"):
            raise ValueError("row_order must be 'natural' or 'bitrev'")
        self.port = port
        self.bin = bin
        self.uncompressed = _truthy(uncompressed)
        self.setup_path = setup_path
        self.precompute_path = precompute_path
        self.devices = list(devices) if devices else [device]
        self.device = self.devices[0]
        self.contexts = max(1, int(contexts))
        self.precompute = precompute
        self.multi_gpu = multi_gpu
        self.poly_form = poly_form
        self.row_order = row_order
        self.test_srs = test_srs
        if staged_upload is True:
            self.STAGE_MIN = 1 << 17
        elif staged_upload is False:
            self.STAGE_MIN = None
        self.scale = None
        self.machines_scale = None
        self.srs_source = None     # "file:<format>" or "test-trapdoor" once started
        self._roots: List[native.Context] = []   # one per device: owns the SRS
        self._slots: List[_Slot] = []
        self._free: "queue.LifoQueue[_Slot]" = queue.LifoQueue()
        self._mg: Optional[native.MultiContext] = None
        self._mg_lock = threading.Lock()
        self._attached = False
        self._pinned = True        # polynomials are decoded into page-locked staging buffers (zkp_host_alloc)
        self._seed = seed if seed is not None else secrets.randbits(63)
        self._counter = 0
        self._counter_lock = threading.Lock()

    # ---- lifecycle (reference base/miner.py:82-84,155,181)
    def start(self, scale: int = 18, machines_scale: int = 8) -> None:
        if machines_scale > scale:
            raise ValueError("machines_scale must not exceed scale")
        self.scale, self.machines_scale = int(scale), int(machines_scale)
        log_n = self.scale - self.machines_scale
        source = srsfile.find_source(self.setup_path, self.precompute_path, self.uncompressed, self.scale, self.machines_scale)
        if source is None:
            opted = self.test_srs if self.test_srs is not None else os.environ.get("ZKP_B200_TEST_SRS", "") not in ("", "0")
            if not opted:
                raise native.ZkpError(native.ZKP_ERR_IO,
                                      f"SRS file {self.setup_path!r} (precompute {self.precompute_path!r}) not found.  Generate the "
                                      f"files with `python -m zkp_subnet_b200.setup` or download the ceremony files; a throw-away "
                                      f"SRS from a PUBLIC trapdoor is only generated when asked for explicitly "
                                      f"(Client(test_srs=True) or ZKP_B200_TEST_SRS=1)")
            print("zkp_b200: WARNING: using a TEST SRS generated from a PUBLIC trapdoor -- openings can be forged by anyone; "
                  "never use this on a live network (nothing is written to setup_path)", file=sys.stderr, flush=True)
        if self.multi_gpu == "split" and len(self.devices) > 1:
            self._mg = native.MultiContext(self.devices)
            if source is None:
                self._mg.srs_generate(TEST_TAU_X, TEST_TAU_Y, log_n, self.machines_scale, native.LAYOUT_POINT_RANGE)
            else:
                srsfile.load_into_multi(self._mg, source, log_n, self.machines_scale, native.LAYOUT_POINT_RANGE)
            if self.precompute == "eager":
                self._mg.prebuild_tables()
            roots = [self._mg.ctx(0)]  # verify / fft / eval / rng run on device 0
        else:
            roots = []
            for dev in self.devices:
                ctx = self._make_root(dev)
                roots.append(ctx)
                if source is None:
                    ctx.srs_generate(TEST_TAU_X, TEST_TAU_Y, log_n, self.machines_scale)
                else:
                    srsfile.load_into(ctx, source)
                    if ctx.srs_shape() != (log_n, self.machines_scale):
                        raise native.ZkpError(native.ZKP_ERR_STATE, f"SRS files hold shape {ctx.srs_shape()}, "
                                              f"expected {(log_n, self.machines_scale)}")
                if self.precompute == "eager":
                    ctx.prebuild_tables()
        self.srs_source = "test-trapdoor" if source is None else f"file:{source.kind}"
        self._roots = roots
        for ctx in roots:
            ctx.set_poly_form(self.poly_form == "coeffs")
        for ctx in roots:
            self._add_slot(ctx)
            for _ in range(self.contexts - 1):
                self._add_slot(ctx.fork())
        if self.precompute == "eager":
            self._warm()

    def _warm(self) -> None:
        """One throw-away request per pooled context (an all-zero polynomial): every device workspace, page-locked
        staging buffer and stream is allocated HERE, so that the first real request costs what every later one costs
        (measured at 2^24: 1.08 s for the first request of a cold context against 0.17 s warm)."""
        if not self._pinned:
            return
        n = 1 << (self.scale - self.machines_scale)
        x = (2).to_bytes(32, "big")
        for slot in self._slots:
            if slot.staging is None:
                slot.staging = native.PinnedBuffer(32 * n)
            slot.staging.write(bytes(32 * n))
            if self._mg is not None:
                self._mg.commit_open(0, slot.staging, x)
                break
            slot.ctx.worker_commit_open(0, slot.staging, x)
            slot.resident_n = 0

    def _make_root(self, device: int):
        """the context that owns the SRS of one device (tests substitute a CPU double here)"""
        return native.Context(device)

    def _add_slot(self, ctx) -> None:
        s = _Slot(ctx)
        self._slots.append(s)
        self._free.put(s)

    def attach(self, ctx: native.Context, scale: int, machines_scale: int) -> "Client":
        """Use an existing context (its SRS already resident) instead of start(); for benchmarks and tests that
        share one GPU context between the raw C-ABI calls and this wire-level shim."""
        self.scale, self.machines_scale = int(scale), int(machines_scale)
        self._attached = True
        self._roots = [ctx]
        self._add_slot(ctx)
        return self

    def stop(self) -> None:
        slots, self._slots = self._slots, []
        self._free = queue.LifoQueue()
        roots = set(id(r) for r in self._roots)
        # forks first, then the contexts that own the SRS
        for s in slots:
            if id(s.ctx) not in roots:
                s.close()
        for s in slots:
            if s.ctx is not None:
                s.close(close_ctx=not self._attached)
        self._roots = []
        if self._mg is not None:
            self._mg.close()
            self._mg = None

    # ---- pool
    class _Lease:
        def __init__(self, client: "Client"):
            self.client = client
            self.slot: Optional[_Slot] = None

        def __enter__(self) -> _Slot:
            c = self.client
            if not c._slots:
                raise native.ZkpError(native.ZKP_ERR_STATE, "Client.start() has not been called")
            self.slot = c._free.get()
            return self.slot

        def __exit__(self, *exc) -> bool:
            self.client._free.put(self.slot)
            return False

    def _lease(self) -> "Client._Lease":
        return Client._Lease(self)

    def _need(self) -> native.Context:
        if not self._roots:
            raise native.ZkpError(native.ZKP_ERR_STATE, "Client.start() has not been called")
        return self._roots[0]

    def _row(self, i: int) -> int:
        i = int(i)
        if self.row_order == "bitrev" and self.machines_scale:
            if not 0 <= i < (1 << self.machines_scale):
                raise ValueError("worker index out of range")
            return _bitrev(i, self.machines_scale)
        return i

    def _decode(self, slot: _Slot, poly: Sequence[str]):
        """Decode a wire polynomial into the slot's page-locked staging buffer (grown on demand)."""
        need = 32 * len(poly)
        slot.resident_n = 0  # the staging buffer is about to change
        if need == 0:
            return b""
        if not self._pinned:
            return decode_poly(poly)
        if slot.staging is None or slot.staging.capacity < need:
            if slot.staging is not None:
                slot.staging.close()
            slot.staging = native.PinnedBuffer(max(need, 32 << (self.scale - self.machines_scale)))
        return decode_poly(poly, slot.staging)

    # From this many elements on, the list is decoded and uploaded in chunks (native.Context.stage_list), each chunk's copy
    # running beside the decode of the next.  Measured at 2^20 on the 16-core GPU host (tools/pool_ab.py,
    # profiles/r2_pool_ab.txt), chunks of 2^18: worker_commit_and_open 13.90 -> 13.5 ms, the reference's two calls
    # 15.3 -> 14.9 ms.  (Before the codec kept its worker threads between calls every chunk cost a spawn/join of the
    # decoder threads, ~0.13 ms, and the chunked form was no faster: 14.27 vs 14.13 ms.)  Client(staged_upload=False)
    # turns it off; Client(staged_upload=True) lowers the threshold to 2^17 elements.
    STAGE_MIN: Optional[int] = 1 << 19

    def _stage(self, slot: _Slot, poly: Sequence[str]) -> Optional[int]:
        """Large polynomials: decode into the slot's staging buffer chunk by chunk, each chunk's host-to-device copy
        running while the next chunk is decoded.  Returns the generation of the upload (the polynomial is then resident
        on the device), or None when the list is small / the client has no page-locked staging (plain decode applies)."""
        n = len(poly)
        if self.STAGE_MIN is None or not self._pinned or n < self.STAGE_MIN or self._mg is not None:
            return None
        slot.resident_n = 0
        if slot.staging is None or slot.staging.capacity < 32 * n:
            if slot.staging is not None:
                slot.staging.close()
            slot.staging = native.PinnedBuffer(max(32 * n, 32 << (self.scale - self.machines_scale)))
        gen = slot.ctx.stage_list(poly, slot.staging)
        slot.resident_n, slot.resident_gen = n, gen
        return gen

    @staticmethod
    def _mark_resident(slot: _Slot, n: int) -> None:
        try:
            gen, rn = slot.ctx.resident_generation()
        except native.ZkpError:
            gen, rn = -1, 0
        slot.resident_n, slot.resident_gen = (n, gen) if rn == n else (0, -1)

    @staticmethod
    def _fail(e: Exception) -> Response:
        code = 400 if isinstance(e, (ValueError, native.ZkpError)) and getattr(e, "code", native.ZKP_ERR_ARG) in (
            native.ZKP_ERR_ARG, native.ZKP_ERR_ENCODING) else 500
        return Response(code, {"error": str(e)})

    def _next_seed(self) -> int:
        with self._counter_lock:
            self._counter += 1
            return (self._seed + 0x9E3779B97F4A7C15 * self._counter) & 0xFFFFFFFFFFFFFFFF

    # ---- prover calls
    def worker_commit(self, i: int, poly: Sequence[str]) -> Response:
        try:
            row = self._row(i)
            with self._lease() as slot:
                gen = self._stage(slot, poly)
                if gen is not None:
                    com = slot.ctx.worker_commit_resident(row, len(poly), gen)
                else:
                    buf = self._decode(slot, poly)
                    if self._mg is not None:
                        with self._mg_lock:
                            com = self._mg.msm_g1(row, buf)
                    else:
                        com = slot.ctx.worker_commit(row, buf)
                        self._mark_resident(slot, len(poly))
            return Response(200, {"commitment": _b64_point(com)})
        except (ValueError, native.ZkpError) as e:
            return self._fail(e)

    def worker_open(self, i: int, poly: Sequence[str], x: str) -> Response:
        try:
            row, xb = self._row(i), _decode_any(x, 32)
            with self._lease() as slot:
                if self._mg is not None:
                    buf = self._decode(slot, poly)
                    with self._mg_lock:
                        _, y, proof = self._mg.commit_open(row, buf, xb)
                else:
                    y, proof = self._open_on(slot, row, poly, xb)
            return Response(200, {"eval": _b64_fr(y), "proof": _b64_point(proof)})
        except (ValueError, TypeError, native.ZkpError) as e:
            return self._fail(e)

    def _open_on(self, slot: _Slot, i: int, poly: Sequence[str], xb: bytes):
        """worker_open.  The reference miner calls worker_commit(i, poly) and then worker_open(i, poly, x) with the
        same list (neurons/miner.py:56-61), i.e. it ships the polynomial twice.  When the polynomial THIS slot uploaded
        last is still resident on the GPU, its opening starts at once while a helper thread decodes the list that
        was actually given into the second staging buffer and compares it, element by element, with the bytes of the
        resident one.  Equal (the reference flow): the result is already on its way and neither the decode nor a
        second upload is on the critical path.  Different: the speculative result is dropped and the regular path
        runs on the freshly decoded bytes.  The speculative call names the upload by its generation
        (zkp_worker_open_resident_gen): if anything else -- another client of a shared context, a raw call -- has
        rewritten the staged polynomial since, the library refuses and the regular path runs."""
        ctx = slot.ctx
        n = len(poly)
        if not (n and slot.resident_n == n and slot.staging is not None):
            gen = self._stage(slot, poly)
            if gen is not None:
                return ctx.worker_open_resident_gen(i, n, gen, xb)
            y, proof = ctx.worker_open(i, self._decode(slot, poly), xb)
            self._mark_resident(slot, n)
            return y, proof
        if slot.staging_alt is None or slot.staging_alt.capacity < 32 * n:
            if slot.staging_alt is not None:
                slot.staging_alt.close()
            slot.staging_alt = native.PinnedBuffer(max(32 * n, slot.staging.capacity))
        # The list is decoded on the slot's helper thread and the GPU call is made from THIS thread: the decoder keeps
        # the GIL for its whole call (ctypes.PyDLL), the prover call releases it (ctypes.CDLL), so the helper gets
        # going the moment this thread is inside the library.  The job is handed over right before the prover call:
        # started any earlier the helper would take the GIL first and the decode would run BEFORE the GPU work
        # instead of beside it (measured at 2^20: 9.5 ms per call instead of 7.1).  The helper is a persistent thread
        # (one per slot, created on first use): creating a threading.Thread per call costs 60-100 us, a twentieth of a
        # whole request at the mainnet row size.
        dec = {}

        def run():
            try:
                dec["ok"] = native.wire_decode_list(poly, slot.staging_alt, same_as=slot.staging)
            except Exception as e:
                dec["err"] = e

        spec = err = None
        done = slot.helper().submit(run)   # handed over right before the prover call (see above)
        try:
            spec = ctx.worker_open_resident_gen(i, n, slot.resident_gen, xb)
        except native.ZkpError as e:  # resident polynomial replaced by another call on a shared context, bad x, ...
            err = e
        finally:
            done.wait()
        if "err" in dec:
            raise dec["err"]
        buf, same = dec["ok"]
        if same and spec is not None:
            return spec
        if same and err is not None and err.code != native.ZKP_ERR_STATE:
            raise err
        # not the resident polynomial: the new bytes become the staged ones
        slot.resident_n = 0
        slot.staging, slot.staging_alt = slot.staging_alt, slot.staging
        y, proof = ctx.worker_open(i, buf, xb)
        self._mark_resident(slot, n)
        return y, proof

    def worker_commit_and_open(self, i: int, poly: Sequence[str], x: str) -> Response:
        """Fused form of the reference's rpc_commit_and_open (neurons/miner.py:56-61): one decode, one upload."""
        try:
            row, xb = self._row(i), _decode_any(x, 32)
            with self._lease() as slot:
                gen = self._stage(slot, poly)
                if gen is not None:
                    com, y, proof = slot.ctx.worker_commit_open_resident(row, len(poly), gen, xb)
                else:
                    buf = self._decode(slot, poly)
                    if self._mg is not None:
                        with self._mg_lock:
                            com, y, proof = self._mg.commit_open(row, buf, xb)
                    else:
                        com, y, proof = slot.ctx.worker_commit_open(row, buf, xb)
                        self._mark_resident(slot, len(poly))
            return Response(200, {"commitment": _b64_point(com), "eval": _b64_fr(y), "proof": _b64_point(proof)})
        except (ValueError, TypeError, native.ZkpError) as e:
            return self._fail(e)

    def worker_commit_and_open_batch(self, items: Sequence[Dict[str, Any]]) -> Response:
        """Several requests ({"i", "poly", "alpha"}) in ONE launch set (zkp_worker_commit_open_batch): what a miner
        serving several validators at once wants at the mainnet row size.  Answers {"results": [...]} with one
        {"commitment", "eval", "proof"} or {"error"} per item, in order.  Not part of the reference's Client."""
        try:
            if self._mg is not None:
                raise native.ZkpError(native.ZKP_ERR_STATE, "batches are not available with multi_gpu='split'")
            with self._lease() as slot:
                slot.resident_n = 0
                rows, bufs, xs, errs = [], [], [], {}
                n = 1 << (self.scale - self.machines_scale)
                stage = native.PinnedBuffer(32 * n * max(1, len(items)))
                try:
                    for k, it in enumerate(items):
                        try:
                            if len(it["poly"]) != n:
                                raise ValueError(f"polynomial must have {n} elements")
                            raw = native.wire_decode_list(it["poly"])
                            xb = _decode_any(it["alpha"], 32)
                            row = self._row(it["i"])
                        except (ValueError, TypeError, KeyError) as e:
                            errs[k] = str(e)
                            continue
                        stage.write(raw, 32 * n * len(rows))
                        rows.append(row)
                        xs.append(xb)
                        bufs.append(k)
                    out = []
                    if rows:
                        views = [(ctypes_view(stage, 32 * n * j, 32 * n)) for j in range(len(rows))]
                        out = slot.ctx.worker_commit_open_batch(rows, views, b"".join(xs))
                finally:
                    stage.close()
                results: List[Dict[str, Any]] = [None] * len(items)  # type: ignore
                for j, k in enumerate(bufs):
                    st, com, y, proof = out[j]
                    results[k] = ({"commitment": _b64_point(com), "eval": _b64_fr(y), "proof": _b64_point(proof)} if st == 0
                                  else {"error": f"zkp_b200 error {st}: malformed field element"})
                for k, msg in errs.items():
                    results[k] = {"error": msg}
            return Response(200, {"results": results})
        except (ValueError, TypeError, native.ZkpError) as e:
            return self._fail(e)

    def worker_verify(self, i: int, proof: str, alpha: str, eval: str, commitment: str) -> Response:
        # malformed encodings are a failed verification with status 200, never an HTTP-style error
        # (reference tests/test_validator.py:66,79-86,103-104 expect reward 0.0, not an exception)
        try:
            args = (_decode_any(proof, 48), _decode_any(alpha, 32), _decode_any(eval, 32), _decode_any(commitment, 48))
        except (ValueError, TypeError):
            return Response(200, {"valid": False})
        try:
            return Response(200, {"valid": self._need().worker_verify(self._row(i), *args)})
        except (ValueError, native.ZkpError) as e:
            return self._fail(e)

    def worker_verify_batch(self, items: Sequence[Dict[str, Any]], alpha: str) -> Response:
        """All responses of one challenge in one call: `items` are dicts with keys i, proof, eval, commitment (the
        arguments of worker_verify); answers {"valid": [bool, ...]} in the same order.  Not part of the reference's
        Client (it verifies one response per call, neurons/validator.py:168-170); malformed items are False."""
        try:
            zero48, zero32 = b"\xff" * 48, b"\xff" * 32  # placeholders that fail to decode -> valid = False
            idx, proofs, evals, coms = [], [], [], []
            for it in items:
                idx.append(self._row(it["i"]))
                try:
                    p, e, c = _decode_any(it["proof"], 48), _decode_any(it["eval"], 32), _decode_any(it["commitment"], 48)
                except (ValueError, TypeError, KeyError):
                    p, e, c = zero48, zero32, zero48
                proofs.append(p); evals.append(e); coms.append(c)
            try:
                a = _decode_any(alpha, 32)
            except (ValueError, TypeError):
                return Response(200, {"valid": [False] * len(idx)})
            if not idx:
                return Response(200, {"valid": []})
            return Response(200, {"valid": self._need().worker_verify_batch(idx, b"".join(proofs), a, b"".join(evals), b"".join(coms))})
        except (ValueError, native.ZkpError) as e:
            return self._fail(e)

    # ---- Pianist master node.  Not part of the reference's Client yet ("multi-miner proofs ... not yet
    #      implemented", reference neurons/validator.py:198; roadmap README.md:38); named after the worker_* calls.
    #      The points come from OTHER parties (the workers), so every one is subgroup-checked before it is added.
    def master_commit(self, commitments: Sequence[str]) -> Response:
        """com = sum_i com_i over the workers' commitments."""
        try:
            raw = b"".join(_decode_any(c, 48) for c in commitments)
            return Response(200, {"commitment": _b64_point(native.g1_sum_checked(raw))})
        except (ValueError, TypeError, native.ZkpError) as e:
            return self._fail(e)

    def _worker_order(self, vals: Sequence[str]) -> List[str]:
        """values indexed by worker -> indexed by SRS row (identity in natural order)"""
        if self.row_order != "bitrev":
            return list(vals)
        out: List[Any] = [None] * len(vals)
        for i, v in enumerate(vals):
            out[self._row(i)] = v
        return out

    def master_open(self, evals: Sequence[str], proofs: Sequence[str], beta: str) -> Response:
        """From the workers' (eval, proof) answers at a common alpha: pi_X = sum_i pi_i, z = f(alpha, beta) and the
        Y-direction proof pi_Y."""
        try:
            pix = native.g1_sum_checked(b"".join(_decode_any(p, 48) for p in proofs))
            z, piy = self._need().master_open_y(b"".join(_decode_any(e, 32) for e in self._worker_order(evals)), _decode_any(beta, 32))
            return Response(200, {"eval": _b64_fr(z), "proof_x": _b64_point(pix), "proof_y": _b64_point(piy)})
        except (ValueError, TypeError, native.ZkpError) as e:
            return self._fail(e)

    def master_verify(self, proof_x: str, proof_y: str, alpha: str, beta: str, eval: str, commitment: str) -> Response:
        try:
            args = (_decode_any(commitment, 48), _decode_any(proof_x, 48), _decode_any(proof_y, 48), _decode_any(alpha, 32),
                    _decode_any(beta, 32), _decode_any(eval, 32))
        except (ValueError, TypeError):
            return Response(200, {"valid": False})
        try:
            return Response(200, {"valid": self._need().master_verify(*args)})
        except native.ZkpError as e:
            return self._fail(e)

    def fft(self, poly: Sequence[str], left: bool = True, inverse: bool = False) -> Response:
        try:
            with self._lease() as slot:
                out = slot.ctx.fft(self._decode(slot, poly), bool(left), bool(inverse))
            return Response(200, {"poly": encode_poly(out)})
        except (ValueError, native.ZkpError) as e:
            return self._fail(e)

    def eval(self, poly: Sequence[str], x: str) -> Response:
        try:
            with self._lease() as slot:
                y = slot.ctx.eval(self._decode(slot, poly), _decode_any(x, 32))
            return Response(200, {"y": _b64_fr(y)})
        except (ValueError, TypeError, native.ZkpError) as e:
            return self._fail(e)

    def challenge_evals(self, polys: Sequence[Sequence[str]], x: str) -> Response:
        """The evaluations a validator needs for one challenge, f_i(x) for every row, in one call (the reference
        does an inverse fft and an eval per row: neurons/validator.py:106-120).  Not part of the reference's Client."""
        try:
            rows = len(polys)
            if rows == 0 or any(len(p) != len(polys[0]) for p in polys):
                raise ValueError("rows of equal, non-zero length expected")
            with self._lease() as slot:
                flat = [s for p in polys for s in p]
                raw = slot.ctx.challenge_evals(self._decode(slot, flat), rows, _decode_any(x, 32))
            return Response(200, {"evals": [_b64_fr(raw[32 * i:32 * i + 32]) for i in range(rows)]})
        except (ValueError, TypeError, native.ZkpError) as e:
            return self._fail(e)

    def random_poly(self) -> Response:
        """Random bivariate polynomial as 2^machines_scale rows of 2^(scale-machines_scale) evaluations
        (reference neurons/validator.py:67-75)."""
        try:
            rows, n = 1 << self.machines_scale, 1 << (self.scale - self.machines_scale)
            with self._lease() as slot:
                slot.resident_n = 0
                raw = slot.ctx.random_poly(self._next_seed(), rows * n)
            flat = encode_poly(raw)
            return Response(200, {"poly": [flat[r * n:(r + 1) * n] for r in range(rows)]})
        except native.ZkpError as e:
            return self._fail(e)

    def random_point(self) -> Response:
        try:
            with self._lease() as slot:
                slot.resident_n = 0
                return Response(200, {"point": _b64_fr(slot.ctx.random_point(self._next_seed()))})
        except native.ZkpError as e:
            return self._fail(e)


def ctypes_view(buf: native.PinnedBuffer, offset: int, nbytes: int):
    """a ctypes char array over a slice of a page-locked buffer (no copy)"""
    import ctypes
    return (ctypes.c_char * nbytes).from_address(ctypes.addressof(buf.buf) + offset)
