"""Host-side mirror of the miner's proof path (reference neurons/miner.py:38-61,106-135): same method
names, argument meaning and error behaviour, minus the bittensor axon/blacklist/priority plumbing (out
of scope: chain policy, no math).  `forward` times the proof with perf_counter, returns a NEW Prove with
poly=[] and alpha=None, and on any exception returns the input synapse unfilled."""
from __future__ import annotations

import logging
import time
import typing

from .client import Client
from .protocol import Prove

log = logging.getLogger("zkp_subnet_b200.miner")


class Miner:
    def __init__(self, client: Client, fused: bool = True):
        self.client = client
        self.fused = fused  # one decode + one upload for commit and open
        self.last_elapsed: typing.Optional[float] = None

    def rpc_commit(self, i: int, poly: typing.List[str]) -> str:
        with self.client.worker_commit(i, poly) as response:
            if response.status_code != 200:
                log.error("RPC request failed with status: %s", response.status_code)
                raise Exception("Failed to commit to the polynomial.")
            return response.json().get("commitment")

    def rpc_open(self, i: int, poly: typing.List[str], x: str) -> typing.Tuple[str, str]:
        with self.client.worker_open(i, poly, x) as response:
            if response.status_code != 200:
                log.error("RPC request failed with status: %s", response.status_code)
                raise Exception("Failed to verify the proof.")
            return response.json().get("eval"), response.json().get("proof")

    def rpc_commit_and_open(self, i: int, poly: typing.List[str], alpha: str) -> typing.Tuple[str, str, str]:
        if self.fused and hasattr(self.client, "worker_commit_and_open"):
            with self.client.worker_commit_and_open(i, poly, alpha) as response:
                if response.status_code != 200:
                    log.error("RPC request failed with status: %s", response.status_code)
                    raise Exception("Failed to commit to / open the polynomial.")
                j = response.json()
                return j.get("commitment"), j.get("eval"), j.get("proof")
        commitment = self.rpc_commit(i, poly)
        eval, proof = self.rpc_open(i, poly, alpha)
        return commitment, eval, proof

    def forward(self, synapse: Prove) -> Prove:
        try:
            before = time.perf_counter()
            commitment, eval, proof = self.rpc_commit_and_open(synapse.index, synapse.poly, synapse.alpha)
            self.last_elapsed = time.perf_counter() - before
            log.info("Proof generation completed in %s seconds", self.last_elapsed)
            return Prove(index=int(synapse.index), poly=[], alpha=None, eval=eval, commitment=commitment, proof=proof)
        except Exception as e:  # noqa: BLE001 -- the reference swallows everything (neurons/miner.py:133-135)
            log.error("Failed to forward synapse: %s", e)
            return synapse
