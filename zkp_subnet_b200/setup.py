"""`python -m zkp_subnet_b200.setup` -- the counterpart of the reference prover's `setup` subcommand
(reference tests/conftest.py:50-65):

    ./prover setup --setup-path P --precompute-path Q --scale S --machines-scale M
                   --generate-setup --generate-precompute --overwrite

Same flags, same meaning; the work runs on the GPU (there is no CPU fallback).

  --generate-setup        write the monomial SRS [tau_x^j tau_y^i]_1 + [tau_x]_2, [tau_y]_2 to --setup-path.  The
                          trapdoor is drawn from the OS entropy source, used once and never printed or stored
                          (a single-party setup: fine for tests and local networks, not a ceremony).
                          --test-trapdoor uses the PUBLIC test values instead (reproducible files for tests).
  --generate-precompute   derive the workers' Lagrange rows U[i][j] = [R_i(tau_y) L_j(tau_x)]_1 and the row scale
                          points from the setup file by inverse group FFTs -- no trapdoor needed, so this is also how
                          the rows are obtained from a downloaded ceremony setup file -- and write them to
                          --precompute-path.
  --uncompressed [bool]   96-byte G1 points (default: 48-byte compressed), as in the reference's *.uncompressed /
                          *.compressed file names (Makefile:30-48, tests/conftest.py:28-29)
File layout: zkp_subnet_b200/srsfile.py.
"""
from __future__ import annotations

import argparse
import os
import secrets
import sys

from . import native, srsfile
from .client import TEST_TAU_X, TEST_TAU_Y, _truthy

FR_MODULUS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m zkp_subnet_b200.setup", description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--setup-path", required=True)
    ap.add_argument("--precompute-path", required=True)
    ap.add_argument("--scale", type=int, required=True)
    ap.add_argument("--machines-scale", type=int, required=True)
    ap.add_argument("--generate-setup", action="store_true")
    ap.add_argument("--generate-precompute", action="store_true")
    ap.add_argument("--overwrite", action="store_true")
    ap.add_argument("--uncompressed", nargs="?", const="true", default="false")
    ap.add_argument("--test-trapdoor", action="store_true", help="use the PUBLIC test trapdoor (forgeable; tests only)")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    if args.machines_scale > args.scale:
        ap.error("--machines-scale must not exceed --scale")
    uncompressed = _truthy(args.uncompressed)
    log_m, log_n = args.machines_scale, args.scale - args.machines_scale
    for flag, path in ((args.generate_setup, args.setup_path), (args.generate_precompute, args.precompute_path)):
        if flag and os.path.exists(path) and not args.overwrite:
            print(f"{path} exists (pass --overwrite to replace it)", file=sys.stderr)
            return 1
    if not (args.generate_setup or args.generate_precompute):
        print("nothing to do: pass --generate-setup and/or --generate-precompute", file=sys.stderr)
        return 1
    with native.Context(args.device) as ctx:
        if args.generate_setup:
            if args.test_trapdoor:
                tx, ty = TEST_TAU_X, TEST_TAU_Y
                print("zkp_b200 setup: WARNING: PUBLIC test trapdoor -- anyone can forge openings against this SRS", file=sys.stderr)
            else:
                tx, ty = 2 + secrets.randbelow(FR_MODULUS - 2), 2 + secrets.randbelow(FR_MODULUS - 2)
            ctx.srs_generate_monomial2(tx, ty, log_n, log_m)
            del tx, ty
            srsfile.write_setup(ctx, args.setup_path, uncompressed)
            print(f"wrote {args.setup_path} ({os.path.getsize(args.setup_path)} bytes, 2^{log_m} x 2^{log_n} points)")
        if args.generate_precompute:
            if not args.generate_setup:
                src = srsfile.find_source(args.setup_path, None, uncompressed, args.scale, args.machines_scale)
                if src is None or src.kind != "raw":
                    print(f"{args.setup_path}: no monomial setup file to derive the precompute file from", file=sys.stderr)
                    return 1
                n = 1 << log_n
                ctx.srs_set_shape(log_n, log_m)
                with open(src.setup_path, "rb") as f:
                    for i in range(1 << log_m):
                        srsfile._import_row(ctx, i, f.read(src.point_bytes * n), src.point_bytes, None)
            ctx.srs_monomial_to_lagrange()
            srsfile.write_precompute(ctx, args.precompute_path, uncompressed)
            print(f"wrote {args.precompute_path} ({os.path.getsize(args.precompute_path)} bytes)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
