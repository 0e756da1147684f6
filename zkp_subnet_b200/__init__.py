"""zkp_subnet_b200 -- B200-native KZG prover backend behind the apollozkp/zkp-subnet `fourier.Client` surface."""
