"""Minimal stand-in for the `bittensor` package (test infrastructure).

The reference's neurons import bittensor for the chain plumbing (wallets, subtensor, metagraph, axon / dendrite) and
for two things the hot path touches: the `bt.Synapse` base class of `base.protocol.Prove` (reference
base/protocol.py:24) and `bt.logging`.  bittensor is not installable here (no network), so the tests that drive the
UNMODIFIED reference classes against the drop-in `fourier.Client` put this directory on sys.path.  It provides exactly
what importing reference `neurons/`, `base/` and `utils/` needs at module level plus a pydantic `Synapse` with the
`dendrite.process_time` field the validator's scoring reads (reference neurons/validator.py:152-176).  No chain, no
network, no policy: constructing a full neuron (`Miner(config)`) is out of scope -- the tests create the neuron objects
with `__new__` and attach the client, which is all `forward` / `generate_challenge` / `reward` use.
"""
from __future__ import annotations

import logging as _pylogging
from typing import Any, Optional

from pydantic import BaseModel, ConfigDict, Field

from . import errors  # noqa: F401


class TerminalInfo(BaseModel):
    model_config = ConfigDict(validate_assignment=False, extra="allow")
    status_code: Optional[int] = None
    status_message: Optional[str] = None
    process_time: Optional[Any] = None
    ip: Optional[str] = None
    port: Optional[int] = None
    hotkey: Optional[str] = None


class Synapse(BaseModel):
    model_config = ConfigDict(validate_assignment=True, extra="allow")
    name: Optional[str] = None
    timeout: Optional[float] = 12.0
    dendrite: Optional[TerminalInfo] = Field(default_factory=TerminalInfo)
    axon: Optional[TerminalInfo] = Field(default_factory=TerminalInfo)

    def deserialize(self):
        return self


class _Logging:
    """bt.logging: callable (bt.logging(config=..., logging_dir=...)) with the level methods the reference uses"""

    def __init__(self):
        self._log = _pylogging.getLogger("bittensor-stub")

    def __call__(self, *a, **k):
        return self

    def _emit(self, level, *parts):
        self._log.log(level, " ".join(str(p) for p in parts))

    def info(self, *p, **k): self._emit(_pylogging.INFO, *p)
    def debug(self, *p, **k): self._emit(_pylogging.DEBUG, *p)
    def trace(self, *p, **k): self._emit(_pylogging.DEBUG, *p)
    def warning(self, *p, **k): self._emit(_pylogging.WARNING, *p)
    def error(self, *p, **k): self._emit(_pylogging.ERROR, *p)
    def success(self, *p, **k): self._emit(_pylogging.INFO, *p)
    def on(self): pass
    def off(self): pass
    def set_trace(self, *a): pass
    def set_debug(self, *a): pass

    @staticmethod
    def add_args(parser):
        pass


logging = _Logging()


def debug(*a, **k):
    pass


def trace(*a, **k):
    pass


def turn_console_on():
    pass


class _Plumbing:
    """base class of the chain objects the reference subclasses or annotates with; never instantiated by the tests"""

    def __init__(self, *a, **k):
        raise NotImplementedError("the bittensor stub has no chain: create neurons with __new__ and attach a client")

    @classmethod
    def add_args(cls, parser):
        pass


class MockSubtensor(_Plumbing): pass
class MockWallet(_Plumbing): pass
class metagraph(_Plumbing): pass
class dendrite(_Plumbing): pass
class axon(_Plumbing): pass
class wallet(_Plumbing): pass
class subtensor(_Plumbing): pass
class Config(dict): pass


def config(parser=None):
    raise NotImplementedError("the bittensor stub does not parse neuron configurations")
