"""Host-side mirror of the validator's challenge + verify path (reference neurons/validator.py:35-42,
58-120,135-192): `Challenge`, the rpc_* wrappers, `generate_challenge`, `reward`, `get_rewards`.  The
dendrite query loop, EMA scoring and weight setting are chain policy and out of scope."""
from __future__ import annotations

import logging
from typing import List, Optional

from .client import Client
from .protocol import Prove

log = logging.getLogger("zkp_subnet_b200.validator")


class Challenge:
    def __init__(self, polys: List[List[str]], alpha: str, evals: List[str]):
        self.polys = polys
        self.alpha = alpha
        self.evals = evals

    def to_synapse(self, i: int) -> Prove:
        return Prove(index=i, poly=self.polys[i], eval=self.evals[i], alpha=self.alpha)


class Validator:
    def __init__(self, client: Client, batched: bool = True):
        self.client = client
        self.batched = batched  # use the one-call challenge / verification entries when the client offers them

    def _call(self, response, key: str, what: str):
        with response as r:
            if r.status_code != 200:
                log.error("RPC request failed with status: %s", r.status_code)
                raise Exception(what)
            return r.json().get(key)

    def rpc_fft(self, poly: List[str], left: bool, inverse: bool) -> List[str]:
        return self._call(self.client.fft(poly, left, inverse), "poly", "Failed to commit to the polynomial.")

    def rpc_random_poly(self) -> List[List[str]]:
        return self._call(self.client.random_poly(), "poly", "Failed to generate a random polynomial.")

    def rpc_worker_verify(self, i: int, proof: str, alpha: str, eval: str, commitment: str) -> bool:
        return self._call(self.client.worker_verify(i, proof, alpha, eval, commitment), "valid", "Failed to verify the proof.")

    def rpc_random_x(self) -> str:
        return self._call(self.client.random_point(), "point", "Failed to generate a random x.")

    def rpc_eval(self, poly: List[str], x: str) -> str:
        return self._call(self.client.eval(poly, x), "y", "Failed to evaluate the polynomial.")

    def generate_challenge(self, machines_count: int) -> Challenge:
        poly = self.rpc_random_poly()
        alpha = self.rpc_random_x()
        if self.batched and hasattr(self.client, "challenge_evals"):
            # same numbers as the per-row inverse fft + Horner below, in one call (barycentric form on the GPU)
            evals = self._call(self.client.challenge_evals(poly[:machines_count], alpha), "evals", "Failed to evaluate the challenge.")
            return Challenge(polys=poly, alpha=alpha, evals=evals)
        evals = []
        for i in range(machines_count):
            fft_coeffs = self.rpc_fft(poly[i], left=True, inverse=True)
            evals.append(self.rpc_eval(fft_coeffs, alpha))
        return Challenge(polys=poly, alpha=alpha, evals=evals)

    def reward(self, challenge: Challenge, response: Prove, process_time: Optional[float], timeout: float) -> float:
        """reference neurons/validator.py:135-176: 0 for missing fields, late or invalid answers, else
        1 - process_time/timeout.  The eval checked is the validator's own, never the miner's."""
        if response.commitment is None or response.proof is None:
            return 0.0
        if process_time is None or process_time > timeout:
            return 0.0
        valid = self.rpc_worker_verify(response.index, response.proof, challenge.alpha, challenge.evals[response.index],
                                       response.commitment)
        if not valid:
            return 0.0
        return 1.0 - process_time / timeout

    def get_rewards(self, challenge: Challenge, responses: List[Prove], process_times: List[Optional[float]],
                    timeout: float) -> List[float]:
        """reference neurons/validator.py:178-192.  Same rewards as calling `reward` per response; when the client
        offers `worker_verify_batch` the responses that are complete and on time are verified in one call."""
        if not self.batched or not hasattr(self.client, "worker_verify_batch"):
            return [self.reward(challenge, r, t, timeout) for r, t in zip(responses, process_times)]
        rewards = [0.0] * len(responses)
        live = [k for k, (r, t) in enumerate(zip(responses, process_times))
                if r.commitment is not None and r.proof is not None and t is not None and t <= timeout]
        if live:
            items = [{"i": responses[k].index, "proof": responses[k].proof, "eval": challenge.evals[responses[k].index],
                      "commitment": responses[k].commitment} for k in live]
            valid = self._call(self.client.worker_verify_batch(items, challenge.alpha), "valid", "Failed to verify the proofs.")
            for k, ok in zip(live, valid):
                if ok:
                    rewards[k] = 1.0 - process_times[k] / timeout
        return rewards
