"""SRS files behind `--setup_path` / `--precompute_path` / `--uncompressed` (reference utils/config.py:131-150,
Makefile:30-48,64-74, tests/conftest.py:50-65).

The reference's prover (`fourier`, external and un-vendored) writes and reads these files; its on-disk layout is not
documented anywhere in the reference tree, so this module states ONE layout, implements both directions, and keeps the
parsing in this single place so that a different layout is a local change (SURVEY.md section 8c "open conventions").

  setup file        ("the SRS": what a ceremony publishes)
      rows x n G1 points  [tau_x^j tau_y^i]_1, row i (machine index) major, column j = power of X
      then 2 x 192 bytes  [tau_x]_2, [tau_y]_2   (ZCash uncompressed G2: x.c1, x.c0, y.c1, y.c0)
  precompute file   (what the workers multiply by; derivable from the setup file alone -- group inverse FFTs, no trapdoor)
      rows x n G1 points  U[i][j] = [R_i(tau_y) L_j(tau_x)]_1  (Lagrange bases over the natural-order domains)
      then rows x 48 bytes  [R_i(tau_y)]_1, ZCash compressed  (the per-row scale points of worker_verify)
  G1 points are ZCash uncompressed (96 bytes, `--uncompressed true`, files named *.uncompressed) or compressed
  (48 bytes, files named *.compressed); the encoding is recognised from the file size, the flag only breaks ties.
  rows = 2^machines_scale, n = 2^(scale - machines_scale).

Besides this "raw" layout the loader accepts the self-describing container written by zkp_srs_save (magic
"ZKPB200S": Lagrange rows, scale points and G2 points in one file) under `setup_path`.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

from . import native

MAGIC = b"ZKPB200S"
G2_TRAILER = 2 * 192


@dataclass
class Source:
    kind: str                      # "native" | "raw"
    setup_path: str
    precompute_path: Optional[str]
    point_bytes: int = 96          # raw: 96 (uncompressed) or 48 (compressed) G1 points in the setup file
    pre_point_bytes: int = 96      # ... and in the precompute file
    log_n: int = 0
    log_m: int = 0


def _point_size(path: str, count: int, trailer: int, prefer: int) -> int:
    size = os.path.getsize(path)
    fits = [p for p in (96, 48) if size == p * count + trailer]
    if not fits:
        raise native.ZkpError(native.ZKP_ERR_IO,
                              f"{path}: {size} bytes is neither {96 * count + trailer} (uncompressed) nor {48 * count + trailer} "
                              f"(compressed) bytes -- not an SRS file for this scale / machines_scale (layout: zkp_subnet_b200/srsfile.py)")
    return prefer if prefer in fits else fits[0]


def find_source(setup_path: Optional[str], precompute_path: Optional[str], uncompressed: bool, scale: int,
                machines_scale: int) -> Optional[Source]:
    """What Client.start() should load, or None when there is no setup file."""
    if not setup_path or not os.path.exists(setup_path):
        return None
    with open(setup_path, "rb") as f:
        head = f.read(8)
    if head == MAGIC:
        return Source("native", setup_path, None)
    log_m, log_n = machines_scale, scale - machines_scale
    count = 1 << scale
    prefer = 96 if uncompressed else 48
    src = Source("raw", setup_path, None, _point_size(setup_path, count, G2_TRAILER, prefer), 96, log_n, log_m)
    if precompute_path and os.path.exists(precompute_path):
        src.precompute_path = precompute_path
        src.pre_point_bytes = _point_size(precompute_path, count, 48 << log_m, prefer)
    return src


def _import_row(ctx, row: int, raw: bytes, point_bytes: int, scale_point: Optional[bytes]) -> None:
    if point_bytes == 96:
        ctx.srs_import_row(row, raw, scale_point)
    else:
        ctx.srs_import_row_compressed(row, raw, scale_point)


def _read_g2(setup_path: str):
    with open(setup_path, "rb") as f:
        f.seek(-G2_TRAILER, os.SEEK_END)
        t = f.read(G2_TRAILER)
    return t[:192], t[192:]


def load_into(ctx, src: Source) -> None:
    """Make the Lagrange rows of `src` resident in `ctx` (one whole SRS on one device)."""
    if src.kind == "native":
        ctx.srs_load(src.setup_path)
        return
    n, rows = 1 << src.log_n, 1 << src.log_m
    ctx.srs_set_shape(src.log_n, src.log_m)
    gx, gy = _read_g2(src.setup_path)
    if src.precompute_path:
        pb = src.pre_point_bytes
        with open(src.precompute_path, "rb") as f:
            f.seek(pb * n * rows)
            scale_points = f.read(48 * rows)
            f.seek(0)
            for i in range(rows):
                _import_row(ctx, i, f.read(pb * n), pb, scale_points[48 * i:48 * i + 48])
    else:
        # no precompute file: derive the worker rows from the monomial setup on the GPU (no trapdoor involved)
        pb = src.point_bytes
        with open(src.setup_path, "rb") as f:
            for i in range(rows):
                _import_row(ctx, i, f.read(pb * n), pb, None)
        ctx.srs_monomial_to_lagrange()
    ctx.srs_import_g2(0, gx)
    ctx.srs_import_g2(1, gy)


def load_into_multi(mg, src: Source, log_n: int, log_m: int, layout: int) -> None:
    """Fill the contexts of a MultiContext: whole SRS on every device (LAYOUT_ROWS) or point-range shards."""
    ndev = len(mg.devices)
    if layout == native.LAYOUT_ROWS:
        for k in range(ndev):
            load_into(mg.ctx(k), src)
        mg.set_layout(layout, log_n, log_m)
        return
    log_shards = max(0, min(ndev.bit_length() - 1, log_n))
    shards, n, rows = 1 << log_shards, 1 << log_n, 1 << log_m
    nl = n >> log_shards
    full = native.Context(mg.devices[0])
    try:
        load_into(full, src)
        if full.srs_shape() != (log_n, log_m):
            raise native.ZkpError(native.ZKP_ERR_STATE, f"SRS files hold shape {full.srs_shape()}, expected {(log_n, log_m)}")
        g2 = [full.srs_export_g2(0)]
        try:
            g2.append(full.srs_export_g2(1))
        except native.ZkpError:
            pass
        ctxs = [mg.ctx(k) for k in range(shards)]
        for c in ctxs:
            c.srs_set_shape(log_n - log_shards, log_m)
        for i in range(rows):
            pts = full.srs_export_row(i, n)
            sp = full.srs_export_scale_point(i)
            for k, c in enumerate(ctxs):
                c.srs_import_row(i, pts[96 * nl * k:96 * nl * (k + 1)], sp)
        for k, c in enumerate(ctxs):
            c.srs_set_shard(log_n, k)
            for which, g in enumerate(g2):
                c.srs_import_g2(which, g)
    finally:
        full.close()
    mg.set_layout(layout, log_n, log_m)


def write_setup(ctx, path: str, uncompressed: bool) -> None:
    """The resident rows (monomial form) + the two G2 points -> setup file."""
    log_n, log_m = ctx.srs_shape()
    n = 1 << log_n
    with open(path, "wb") as f:
        for i in range(1 << log_m):
            f.write(ctx.srs_export_row(i, n) if uncompressed else ctx.srs_export_row_compressed(i, n))
        f.write(ctx.srs_export_g2(0))
        f.write(ctx.srs_export_g2(1))


def write_precompute(ctx, path: str, uncompressed: bool) -> None:
    """The resident rows (Lagrange form) + their scale points -> precompute file."""
    log_n, log_m = ctx.srs_shape()
    n = 1 << log_n
    with open(path, "wb") as f:
        for i in range(1 << log_m):
            f.write(ctx.srs_export_row(i, n) if uncompressed else ctx.srs_export_row_compressed(i, n))
        for i in range(1 << log_m):
            f.write(ctx.srs_export_scale_point(i))
