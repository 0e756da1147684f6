#!/usr/bin/env python3
"""Generates tests/golden/vectors.json from oracle/bls12_381.py (plain big-int arithmetic).

The reference's own prover (`fourier`, Rust, un-vendored) cannot run here, so these are NOT outputs of
the reference: they are (a) the reference's de-facto known answer TEST_POLY/TEST_POINT/TEST_EVAL copied
from reference tests/test_miner.py:33-55, (b) the standard ZCash G1 encodings, and (c) vectors derived
by the Python oracle (cross-checked against SURVEY.md section 8c's independently derived values).
Run:  python tests/golden/make_vectors.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import bls12_381 as o  # noqa: E402

TEST_POLY = [
    "aUXcXE/02sinJ4ybjw1GEzIM+H/5R/Iayb9CMn7BlEg", "aOQMCI2Ce8zgLO80vcjBK7Al++oEe8bADAyMXJJbf68",
    "ZygfrBZOk0i4BpO6MNXU4xHeWHjrPSDjSlhQe0hLJDw", "X3w3fa5rnZq6113BXk//n+dSDR+FIkyV9IX0SXgVTFo",
    "LYXDdqRAtuJcP3wRVZtqJ2hAI/NsPXoKzX59AZ3jmcc", "Sm+5XwJBs1g3ceeZEgyHquPIQ+zbUKOCVKkuGYloki8",
    "EAUHn5bsQSpxn+Lp+mfUIdmPtN7EGBRZ5ZQw9dUCvSo", "ZJYLhpIGLcsBwP+6xWlHiomtiA7Tyd9xC+1c519IRpM",
    "A8KIIVWkR2Qr0h+xzyVT+AlVcT8Ju7vZck4sv9ixnUE", "CrB/7LWe40NfYSn81gLLUZ5W17QmlBYz43o7Z2okgw8",
    "EvpYYUWe/7rmVIJ9mL/f6lVF3fi7lihXlGPaIfF0YrU", "amKWoDdtgHUw2wnci7Bp/97D11QUl7gscioZnWt8WwY",
    "FT0sgbVNfhw+g+phx/Zv2IFV8XE+5YHivoQ4yp/uGgI", "IWvMxK6X/j4dSyHDdcRhQPoVPnhoIBpDSAiJBHrNDC0",
    "OBvU/pJOsQ4I8qIn09sgg6oOWh9mHNPHAsS4qTheeDk", "cjp2QP1+ZUcxMVY6tVFJFqyGHCaVzmUT5QYeWX5eGoE",
]
TEST_POINT = "RWAG//VkEtMp1SeQHQKHelgaic+md8qWPrnWgHZiNMw"
TEST_EVAL = "KXMqHg4HSrBe5qnld5TFrRlluYtsjG7N6WrHduoG/1s"
TAU_X = o.TEST_SECRET
TAU_Y = 0x1234567890ABCDEF1234567890ABCDEF

poly = [o.fr_from_b64(s) for s in TEST_POLY]
x = o.fr_from_b64(TEST_POINT)
mono, lag = o.srs_monomial(16, TAU_X), o.srs_lagrange(16, TAU_X)
yA, pA = o.kzg_open_coeffs(poly, x, mono)
yB, pB = o.kzg_open_evals(poly, x, lag)
w16 = o.root_of_unity(16)
xd = pow(w16, 5, o.R)
yD, pD = o.kzg_open_evals(poly, xd, lag)

# Pianist rows: scale 6, machines_scale 2 -> 4 rows of 16 (the reference's unit-test shape, tests/conftest.py:26-27)
Rs = o.lagrange_at(4, TAU_Y)
pianist = []
for i in range(4):
    row = o.srs_lagrange(16, TAU_X, scale=Rs[i])
    com = o.kzg_commit(poly, row)
    y, proof = o.kzg_open_evals(poly, x, row)
    pianist.append({"row": i, "scale_point": o.g1_compress(o.g1_mul(o.G1_GEN, Rs[i])).hex(),
                    "commitment": o.g1_compress(com).hex(), "eval": o.fr_to_b64(y), "proof": o.g1_compress(proof).hex()})

# Pianist master (config 5 in miniature): 4 different sub-polynomials (rotations of TEST_POLY), one per row;
# aggregated commitment, X-opening at alpha = TEST_POINT, Y-opening at beta = TEST_EVAL (and at a domain point)
beta = o.fr_from_b64(TEST_EVAL)
m_rows = []
for i in range(4):
    row = o.srs_lagrange(16, TAU_X, scale=Rs[i])
    fi = poly[i:] + poly[:i]
    yi, pi_i = o.kzg_open_evals(fi, x, row)
    m_rows.append((o.kzg_commit(fi, row), yi, pi_i))
m_com = o.master_aggregate([r[0] for r in m_rows])
m_pix = o.master_aggregate([r[2] for r in m_rows])
m_z, m_piy = o.master_open_y([r[1] for r in m_rows], beta, TAU_Y)
g2x, g2y = o.g2_mul(o.G2_GEN, TAU_X), o.g2_mul(o.G2_GEN, TAU_Y)
assert o.master_verify(m_com, m_pix, m_piy, x, beta, m_z, g2x, g2y)
assert not o.master_verify(m_com, m_pix, m_piy, x, beta, (m_z + 1) % o.R, g2x, g2y)
beta_d = pow(o.root_of_unity(4), 3, o.R)
m_zd, m_piyd = o.master_open_y([r[1] for r in m_rows], beta_d, TAU_Y)
assert m_zd == m_rows[3][1] and o.master_verify(m_com, m_pix, m_piyd, x, beta_d, m_zd, g2x, g2y)
master = {
    "rows": [{"commitment": o.g1_compress(c).hex(), "eval": o.fr_to_b64(yy), "proof": o.g1_compress(pp).hex()}
             for c, yy, pp in m_rows],
    "commitment": o.g1_compress(m_com).hex(), "proof_x": o.g1_compress(m_pix).hex(),
    "beta": o.fr_to_b64(beta), "z": o.fr_to_b64(m_z), "proof_y": o.g1_compress(m_piy).hex(),
    "beta_in_domain": o.fr_to_b64(beta_d), "z_in_domain": o.fr_to_b64(m_zd), "proof_y_in_domain": o.g1_compress(m_piyd).hex(),
}

vec = {
    "source": "oracle/bls12_381.py (big-int); TEST_* copied from reference tests/test_miner.py:33-55",
    "tau_x": str(TAU_X), "tau_y": str(TAU_Y),
    "test_poly": TEST_POLY, "test_point": TEST_POINT, "test_eval": TEST_EVAL,
    "g1_encodings": {
        "G": o.g1_compress(o.G1_GEN).hex(), "negG": o.g1_compress(o.g1_neg(o.G1_GEN)).hex(),
        "2G": o.g1_compress(o.g1_mul(o.G1_GEN, 2)).hex(), "inf": o.g1_compress(None).hex(),
    },
    "roots_of_unity": {str(k): hex(o.root_of_unity(1 << k)) for k in (4, 6, 12, 16, 20, 24, 32)},
    "A_coeff_form": {"eval": o.fr_to_b64(yA), "commitment": o.g1_compress(o.kzg_commit(poly, mono)).hex(),
                     "proof": o.g1_compress(pA).hex()},
    "B_eval_form": {"eval": o.fr_to_b64(yB), "commitment": o.g1_compress(o.kzg_commit(poly, lag)).hex(),
                    "proof": o.g1_compress(pB).hex()},
    "B_in_domain": {"x": o.fr_to_b64(xd), "eval": o.fr_to_b64(yD), "proof": o.g1_compress(pD).hex()},
    "ntt16": [o.fr_to_b64(v) for v in o.ntt(poly)],
    "intt16": [o.fr_to_b64(v) for v in o.ntt(poly, inverse=True)],
    "pianist_4x16": pianist,
    "pianist_master_4x16": master,
    "g2_tau_y": [hex(c) for c in (lambda p: (p[0][0], p[0][1], p[1][0], p[1][1]))(g2y)],
    "g2_tau_x": [hex(c) for c in (lambda p: (p[0][0], p[0][1], p[1][0], p[1][1]))(o.g2_mul(o.G2_GEN, TAU_X))],
}
out = os.path.join(os.path.dirname(__file__), "vectors.json")
json.dump(vec, open(out, "w"), indent=1)
print("wrote", out)
