"""`from fourier import Client` -- the import the reference's neurons use (base/miner.py:26,
base/validator.py:28).  Putting this repository on PYTHONPATH ahead of the Rust `fourier` package makes
the B200 backend the prover behind an unmodified miner / validator."""
from zkp_subnet_b200.client import Client, Response  # noqa: F401

__all__ = ["Client", "Response"]
