"""The `Prove` wire type of the hot path, field for field as in reference base/protocol.py:24-63, without
the bittensor dependency (bt.Synapse is a pydantic model; the reference's miner only reads
`index`, `poly`, `alpha` and writes `eval`, `commitment`, `proof`)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional


@dataclass
class Prove:
    index: int
    poly: List[str] = field(default_factory=list)
    alpha: Optional[str] = None
    eval: Optional[str] = None
    commitment: Optional[str] = None
    proof: Optional[str] = None

    def deserialize(self) -> "Prove":
        return self
