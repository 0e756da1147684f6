"""CPU: the C-ABI library builds, loads, exports every symbol include/zkp_b200.h declares, and refuses to
run without a GPU (no CPU fallback).  No device compute is attempted here."""
import ctypes
import os
import re

import pytest

from zkp_subnet_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "zkp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(zkp_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_the_reference_surface():
    syms = declared_symbols()
    for needed in ("zkp_ctx_create", "zkp_ctx_destroy", "zkp_worker_commit", "zkp_worker_open", "zkp_worker_verify",
                   "zkp_fft", "zkp_eval", "zkp_random_poly", "zkp_random_point", "zkp_worker_commit_open"):
        assert needed in syms


def test_library_exports_every_declared_symbol():
    if not os.path.exists(native.LIB_PATH):
        pytest.fail(f"{native.LIB_PATH} missing: run `make` or __graft_entry__.build()")
    handle = ctypes.CDLL(native.LIB_PATH)
    for sym in declared_symbols():
        assert hasattr(handle, sym), f"{sym} declared in include/zkp_b200.h but not exported"
    assert sorted(native.EXPORTED_SYMBOLS) == declared_symbols()


def test_no_cpu_fallback_without_gpu():
    lib = native.lib()
    if lib.zkp_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(native.ZkpError) as e:
        native.Context(0)
    assert e.value.code == native.ZKP_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_client_requires_start():
    from zkp_subnet_b200.client import Client
    c = Client(port=1337, bin="./prover", uncompressed=True, setup_path=None, precompute_path=None)
    with pytest.raises(native.ZkpError):
        c._need()
    c.stop()  # idempotent
