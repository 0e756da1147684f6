#!/usr/bin/env python3
"""Benchmark of the KZG hot path (BASELINE.json metric: "KZG commit+open/sec @2^20 BLS12-381; G1 MSM Mpts/s").

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a, through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle's restatement of the reference
                                                           # prover at the SAME size (2^20), all host threads

A step = one commit + open of one random degree-2^20 polynomial given in evaluation form (BASELINE.json
configs[2]; the largest single-GPU configuration on which the metric is quoted).  For N > 1 the job is the
Pianist split of north_star / configs[4]: ONE global vector of N x 2^20 evaluations, sub-polynomial (SRS row) r on
GPU r, no collective on the inner loop; EVERY timed step ends with the cross-GPU combine (each rank contributes its
two partial points as Jacobian coordinates, 288 bytes, rank 0 adds them and compresses once) -- measured inside the loop, not added afterwards -- and
after the loop rank 0 checks the aggregate against the oracle and with the master node's pairing check.  Weak scaling.
Launch for N > 1:  python -m torch.distributed.run --nproc-per-node N bench.py ...   (torch is used only for the
process group: barrier, max-over-ranks; the 288-byte exchange itself goes through host shared memory).

The same jobs are then run through the IN-LIBRARY multi-GPU entries (zkp_mgpu_*: one process, one host thread per
device, no torch) by rank 0 over all N devices while the other ranks wait, and must give the same bytes
(`mgpu_in_library`).

Prints ONE JSON line on rank 0 (contract in the task statement): value = device-timed throughput with the
polynomial resident in HBM; e2e = the same call through the C ABI with host buffers (H2D + D2H inside);
roofline = the dominant kernel (bucket accumulation) against the IMAD.WIDE issue peak measured in this
process; cpu_baseline = the oracle on the host cores (a reported figure, not the target).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TAU_X = 1927409816240961209460912649124
TAU_Y = 0x1234567890ABCDEF1234567890ABCDEF
FR_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
METRIC = "KZG commit+open/sec @2^20 BLS12-381"
UNIT = "commit+open/s"
FQ_MUL_MACS = 300  # 2*12^2 + 12 wide multiply-accumulates per Fq Montgomery product
SEED_POLY = 0xB200 + 3   # seed 0xB200 + config#: ONE global stream, rank r owns elements [r n, (r+1) n)
SEED_MSM24 = 0xB200 + 4


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin: float = 0.0, t_end: float = float("inf")) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        window = [l for t, l in self.lines if t_begin <= t <= t_end] or [l for _, l in self.lines[-3:]]
        for line in window:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def workload_config(log_n: int, world: int) -> dict:
    """`config` of the JSON line: ONE description of the workload, identical for our arm and for --impl reference"""
    return {"workload": f"KZG commit+open of a random degree-2^{log_n} polynomial in evaluation form over BLS12-381 "
                        f"(BASELINE configs[2]); Lagrange SRS from the public test trapdoor; "
                        f"N>1 = Pianist split of ONE global vector, one sub-polynomial per GPU, 288-byte exchange + sum "
                        f"inside every timed step",
            "log_n": log_n, "rows": 1 << (world - 1).bit_length(),
            "l2": "GPU arm: flushed (256 MiB memset) before every timed iteration of `value`; e2e working set "
                  "(fixed-base tables 3.25 GiB gathered at random + sorted entry pairs 104 MiB + buckets 96 MiB) exceeds the 126 MB L2",
            "seed": "0xB200+3 (one global stream; rank r owns elements [r n, (r+1) n))"}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def lagrange_row_scalars(ref, log_m: int):
    """R_i(tau_y), i < 2^log_m (the per-row factors of the Pianist SRS), as integers"""
    if log_m == 0:
        return [1]
    return ref.split32(ref.lagrange_scalars(1 << log_m, TAU_Y))


def cpu_reference_sample(srs96: bytes, poly: bytes, x: bytes, threads: int):
    """One commit + open of the oracle (C restatement) on `threads` host threads; returns seconds."""
    from oracle import ref
    t0 = time.perf_counter()
    com = ref.msm(srs96, poly, threads)
    y, proof = ref.open_evals(poly, x, srs96, threads)
    return time.perf_counter() - t0, com, y, proof


def run_reference(args):
    """--impl reference: the reference's CPU prover cannot be built here (external Rust crate `fourier`, no
    cargo / network), so this arm times the oracle port (cpu_baseline.kind = "port") with every host thread, on
    the SAME workload as our arm: each step is one real commit + open of a 2^20-evaluation polynomial (about 3.5 s
    on 16 threads), nothing extrapolated.  The SRS (one Lagrange row from the public test trapdoor) is generated on
    the CPU before the timed region."""
    rank, world, local = dist_env()
    if rank != 0:
        return 0
    from oracle import ref
    threads = os.cpu_count() or 1
    log_n = args.ref_log_n
    n = 1 << log_n
    t0 = time.perf_counter()
    srs = ref.srs(n, TAU_X, "lagrange", threads=threads)
    t_srs = time.perf_counter() - t0
    times = []
    check = None
    for step in range(args.warmup + args.steps):
        poly = ref.random_scalars_ctr(SEED_POLY, 0, n) if step == 0 else ref.random_scalars(SEED_POLY + step, n)
        x = ref.random_scalars(77 + step, 1)
        dt, com, y, proof = cpu_reference_sample(srs, poly, x, threads)
        if step == 0:
            check = com.hex()
        if step >= args.warmup:
            times.append(dt)
    sec_per_step = statistics.mean(times)
    value = 1.0 / sec_per_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(log_n, max(1, args.gpus)),
        "arm": "CPU restatement of the reference prover (the Rust `fourier` binary cannot be built offline); every step is a "
               "commit+open at the full size on all host threads; at N > 1 rank 0 alone runs it (one sub-polynomial)",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} real commit+open steps at n=2^{log_n} on {threads} host threads "
                                   f"({sec_per_step:.2f} s each, min {min(times):.2f} max {max(times):.2f}); "
                                   f"SRS generated on the CPU in {t_srs:.1f} s before the timed region; restatement, not blst"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "first_commitment": check,
    }
    print(json.dumps(line), flush=True)
    return 0


class Gatherer:
    """all_gather of a fixed-size byte string over NCCL with preallocated tensors (the N partial points of a step)."""

    def __init__(self, dist, local: int, nbytes: int):
        import torch
        self.torch, self.dist, self.n = torch, dist, nbytes
        self.world = dist.get_world_size()
        self.src_host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        self.src = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{local}")
        self.dst = torch.empty(self.world * nbytes, dtype=torch.uint8, device=f"cuda:{local}")
        self.dst_host = torch.empty(self.world * nbytes, dtype=torch.uint8).pin_memory()

    def __call__(self, mine: bytes):
        t = self.torch
        self.src_host.copy_(t.frombuffer(bytearray(mine), dtype=t.uint8))
        self.src.copy_(self.src_host, non_blocking=True)
        self.dist.all_gather_into_tensor(self.dst, self.src)
        self.dst_host.copy_(self.dst, non_blocking=True)
        t.cuda.current_stream().synchronize()
        raw = self.dst_host.numpy().tobytes()
        return [raw[r * self.n:(r + 1) * self.n] for r in range(self.world)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--ref-log-n", type=int, default=20, help="size of the CPU arm (default: the real workload)")
    ap.add_argument("--cpu-baseline-log-n", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config2", action="store_true", help="skip the 2^16 / 2^12 / pipelined side measurements")
    ap.add_argument("--no-msm24", action="store_true", help="skip the sharded 2^24 MSM (BASELINE configs[3])")
    ap.add_argument("--no-mgpu", action="store_true", help="skip the in-library multi-GPU leg")
    ap.add_argument("--no-2p24-open", action="store_true", help="skip the 2^24 commit+open line")
    ap.add_argument("--msm-log-n", type=int, default=24)
    ap.add_argument("--combine", default="host", choices=["host", "nccl"],
                    help="how the ranks' partial points reach rank 0: host shared memory (default) or an NCCL all_gather")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference(args)

    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    dist = host_group = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_group = dist.new_group(backend="gloo")  # host-side barrier while rank 0 drives every GPU itself

    from zkp_subnet_b200 import native
    ctx = native.Context(local)  # raises without a GPU: no CPU fallback
    log_n = args.log_n
    n = 1 << log_n
    log_m = (world - 1).bit_length()
    row = rank  # Pianist: sub-polynomial `rank` on GPU `rank`
    ctx.srs_generate(TAU_X, TAU_Y, log_n, log_m)
    t0 = time.perf_counter()
    ctx.prebuild_tables(row, 1)  # what Client.start(precompute="eager") does: no table build inside a request
    t_tables = time.perf_counter() - t0
    poly = native.PinnedBuffer(32 * n).write(ctx.random_poly_range(SEED_POLY, rank * n, n))
    x = ctx.random_point(0xA1FA)
    c, W, fq_muls = ctx.msm_info(n)

    def barrier():
        if dist is not None:
            dist.barrier()

    gather = hx = None
    if dist is not None:
        if args.combine == "nccl":
            g_nccl = Gatherer(dist, local, 288)
            gather = lambda step, mine: g_nccl(mine)  # noqa: E731
        else:
            from zkp_subnet_b200 import sharding
            hx = sharding.HostExchange(rank, world, 288, os.environ.get("MASTER_PORT", "0"))
            dist.barrier()
            hx.attach()
            gather = hx.gather
    agg = [None]
    step_no = [0]

    def combine():
        """the cross-GPU step of the Pianist job: 288 bytes per rank (host-resident results, exchanged host to host),
        rank 0 adds 2 x N Jacobian points and compresses the two sums (two field inversions per step in the whole job)"""
        step_no[0] += 1
        parts = gather(step_no[0], ctx.last_points_jacobian())  # 2 x 144 bytes per rank, no inversion on the ranks
        if rank == 0:
            cat = b"".join(parts)
            agg[0] = (native.g1_sum_jacobian(cat, world, 288), native.g1_sum_jacobian(cat[144:], world, 288))

    # ---- device-timed value: polynomial resident in HBM, L2 flushed between iterations; at N > 1 every step
    #      includes its combine (wall clock of gather + sum, added to the device time of the same step)
    sampler = ClockSampler(local)
    sampler.start()  # nvidia-smi needs ~1 s to start reporting; only samples inside the timed window are used
    # Pre-warm: on a box that has been idle the first ~0.1 s of sustained load runs the accumulation kernel about 4%
    # slower (same binary, measured), so the device gets 25 untimed steps before the W warm-up steps
    ctx.bench_commit_open(row, poly, x, 25, False)
    for _ in range(args.warmup):
        ctx.bench_commit_open(row, poly, x, 1, True)
        if gather:
            combine()
    barrier()
    t_begin = time.perf_counter()
    t_dev = t_comb = 0.0
    kernel_ms = []
    for _ in range(args.steps):
        ms_iter, ms_k, launches, com, y, proof = ctx.bench_commit_open(row, poly, x, 1, True)
        t_dev += ms_iter
        kernel_ms.append(ms_k)
        if gather:
            t0 = time.perf_counter()
            combine()
            t_comb += (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop(t_begin, time.perf_counter())
    t_rank = t_dev + t_comb
    ms_kernel = statistics.mean(kernel_ms)

    # ---- e2e: the C-ABI call with host buffers (H2D of the polynomial from page-locked host memory + D2H of the
    #      results inside the timed region, every step), plus the combine at N > 1
    for _ in range(2):
        ctx.worker_commit_open(row, poly, x)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e_com, e_y, e_proof = ctx.worker_commit_open(row, poly, x)
        if gather:
            combine()
    e2e_rank = (time.perf_counter() - t0) * 1e3
    assert (e_com, e_y, e_proof) == (com, y, proof), "e2e and device-resident paths disagree"
    poly_bytes = poly.tobytes()
    # the same call from ordinary pageable memory (what a caller gets without zkp_host_alloc)
    def median_ms(fn, reps=7):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append((time.perf_counter() - t0) * 1e3)
        return sorted(ts)[len(ts) // 2]
    e2e_pageable_ms = median_ms(lambda: ctx.worker_commit_open(row, poly_bytes, x))

    if dist is not None:
        import torch
        t = torch.tensor([t_rank, e2e_rank, ms_kernel, t_comb], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_job, e2e_job, ms_kernel_max, t_comb_max = t.tolist()
    else:
        t_job, e2e_job, ms_kernel_max, t_comb_max = t_rank, e2e_rank, ms_kernel, 0.0

    # ---- correctness of what was just timed, on hardware (outside the timed region)
    ok = ctx.worker_verify(row, proof, x, y, com)
    checks = {"worker_verify": bool(ok)}
    all_parts = None
    if dist is not None:
        from zkp_subnet_b200 import sharding
        all_parts = sharding.gather_bytes(dist, com + y + proof, f"cuda:{local}")
    if rank == 0:
        from oracle import ref
        # the polynomial each rank used is the oracle's restatement of the device stream, and the (aggregated)
        # commitment equals [sum_i R_i(tau_y) sum_j f_ij L_j(tau_x)] G computed from those inputs by the oracle
        assert poly_bytes == ref.random_scalars_ctr(SEED_POLY, 0, n), "device RNG differs from its oracle restatement"
        Rs = lagrange_row_scalars(ref, log_m)
        L = ref.lagrange_scalars(n, TAU_X)
        acc = 0
        for r in range(world):
            pr = poly_bytes if r == 0 else ref.random_scalars_ctr(SEED_POLY, r * n, n)
            acc = (acc + Rs[r] * int.from_bytes(ref.fr_dot(pr, L), "big")) % FR_MOD
        expect = ref.g1_mul_gen(ref.fr_be(acc))
        if world == 1:
            assert com == expect, "commitment differs from the oracle's trapdoor value"
            checks["commitment_equals_oracle"] = True
        else:
            coms = b"".join(p[:48] for p in all_parts)
            ys = b"".join(p[48:80] for p in all_parts)
            proofs = b"".join(p[80:128] for p in all_parts)
            assert agg[0] == (native.g1_sum(coms), native.g1_sum(proofs)), "in-loop aggregate differs from the sum of the ranks' answers"
            assert agg[0][0] == expect, "aggregated commitment differs from the oracle's trapdoor value"
            beta = ctx.random_point(0xBE7A)
            if (1 << log_m) == world:  # one row per GPU: the master node's Y-direction opening + bivariate pairing check
                z, pi_y = ctx.master_open_y(ys, beta)
                assert ctx.master_verify(agg[0][0], agg[0][1], pi_y, x, beta, z), "master_verify rejects the aggregate"
                checks["master_verify_aggregate"] = True
            checks["aggregate_equals_oracle"] = True
            checks["aggregated_commitment"] = agg[0][0].hex()

    # ---- through the reference-facing shim: fourier.Client.worker_commit_and_open(i, List[str], str) -- base64
    #      decode of 2^20 strings on the host + the call above + base64 of the results
    client_ms = client_two_calls_ms = None
    if rank == 0:
        import base64
        from zkp_subnet_b200.client import Client, encode_poly
        cl = Client().attach(ctx, log_n + log_m, log_m)
        strs = encode_poly(poly_bytes)
        xs = base64.b64encode(x).decode().rstrip("=")
        resp = cl.worker_commit_and_open(row, strs, xs)
        client_ms = median_ms(lambda: cl.worker_commit_and_open(row, strs, xs))  # median of 7 calls
        assert resp.status_code == 200 and base64.b64decode(resp.json()["commitment"]) == com
        # the UNMODIFIED reference miner makes two calls and ships the polynomial twice
        # (neurons/miner.py:56-61: rpc_commit, then rpc_open)
        cl.worker_commit(row, strs)
        cl.worker_open(row, strs, xs)  # warm-up: allocates the second page-locked staging buffer (one-off)
        r1 = cl.worker_commit(row, strs)
        r2 = cl.worker_open(row, strs, xs)
        client_two_calls_ms = median_ms(lambda: (cl.worker_commit(row, strs), cl.worker_open(row, strs, xs)))
        assert r1.status_code == 200 and base64.b64decode(r1.json()["commitment"]) == com
        assert r2.status_code == 200 and base64.b64decode(r2.json()["proof"]) == proof
        cl.stop()
        del strs

    t0 = time.perf_counter()
    for _ in range(5):
        ctx.worker_verify(row, proof, x, y, com)
    verify_ms = (time.perf_counter() - t0) * 1e3 / 5  # host arithmetic (pairing): the validator's cost per response

    # ---- BASELINE configs[3]: one G1 MSM of 2^24 points (SRS row 1.5 GiB), point-range sharded over the N GPUs --
    #      rank g holds points [g n/N, (g+1) n/N) and elements [g n/N, (g+1) n/N) of ONE global scalar vector (the same
    #      vector at every N), runs the whole Pippenger locally and contributes one partial point, combined INSIDE every
    #      timed repetition (strong scaling of a single MSM; the second half of the metric string)
    msm24 = None
    if not args.no_msm24 and world & (world - 1) == 0:
        lg24, log_shards = args.msm_log_n, world.bit_length() - 1
        n_local = (1 << lg24) >> log_shards
        ctx24 = native.Context(local)
        t0 = time.perf_counter()
        ctx24.srs_generate_shard(TAU_X, TAU_Y, lg24, 0, rank, log_shards)
        t_srs = time.perf_counter() - t0
        sc24 = native.PinnedBuffer(32 * n_local).write(ctx24.random_poly_range(SEED_MSM24, rank * n_local, n_local))
        ctx24.bench_msm(0, sc24, 1, True)  # builds the fixed-base tables of the shard
        g24 = gather
        full24 = [None]

        def combine24():
            step_no[0] += 1
            parts = g24(step_no[0], ctx24.last_points_jacobian())
            if rank == 0:
                full24[0] = native.g1_sum_jacobian(b"".join(parts), world, 288)
        if g24:
            combine24()
        barrier()
        reps24, t24, t24c, k24s = 3, 0.0, 0.0, []
        for _ in range(reps24):
            ms, part24 = ctx24.bench_msm(0, sc24, 1, True)
            t24 += ms
            k24s.append(ctx24.bench_last_kernel_ms())
            if g24:
                t0 = time.perf_counter()
                combine24()
                t24c += (time.perf_counter() - t0) * 1e3
        ms24, comb24 = (t24 + t24c) / reps24, t24c / reps24
        c24, W24, muls24 = ctx24.msm_info(n_local)
        if dist is not None:
            import torch
            tt = torch.tensor([ms24, comb24], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms24_job, comb24_job = tt.tolist()
        else:
            ms24_job, comb24_job, full24[0] = ms24, 0.0, part24
        expect_ok = None
        if rank == 0 and not args.no_cpu_baseline:
            from oracle import ref
            nn = 1 << lg24
            exp24 = ref.g1_mul_gen(ref.fr_dot(ref.random_scalars_ctr(SEED_MSM24, 0, nn), ref.lagrange_scalars(nn, TAU_X)))
            assert full24[0] == exp24, "sharded 2^24 commitment differs from the oracle's trapdoor value"
            expect_ok = True
        msm24 = {"log_n": lg24, "points_per_gpu": n_local, "ms": ms24_job, "combine_ms_inside": comb24_job,
                 "mpts_per_s": (1 << lg24) / (ms24_job * 1e-3) / 1e6,
                 "window_bits": c24, "windows": W24, "fq_mul_per_s_per_gpu": muls24 / ((ms24 - comb24) * 1e-3),
                 "accumulate_kernel_ms": statistics.mean(k24s), "srs_shard_generation_s": t_srs,
                 "commitment": full24[0].hex() if full24[0] else None, "commitment_equals_oracle": expect_ok,
                 "note": "ONE global scalar vector (seed 0xB200+4), the same at every N, sliced by rank; scalars resident in "
                         "HBM, L2 flushed; per repetition: device time of the local MSM + wall time of the gather and sum of "
                         "the N uncompressed partials; max over ranks"}
        sc24.close()
        ctx24.close()

    # ---- the same jobs through the IN-LIBRARY multi-GPU entries: rank 0 alone drives all N devices (one host thread
    #      per device inside libzkp_b200.so, no torch on the path); the other ranks wait on a host barrier
    mgpu = None
    if not args.no_mgpu:
        if dist is not None:
            dist.barrier(group=host_group)
        if rank == 0:
            mgpu = mgpu_leg(native, args, world, log_n, log_m, x, all_parts, (com, y, proof), agg[0], msm24)
        if dist is not None:
            dist.barrier(group=host_group)

    if hx is not None:
        dist.barrier()
        hx.close()
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0 if ok else 1

    side = {}
    if not args.no_config2 and world == 1:  # side figures are per-GPU: measured in the N = 1 run only
        side = side_measurements(native, ctx, local, log_n, log_m, row, poly, x, (com, y, proof), world, args)

    imad_peak, fq_chain_peak = ctx.bench_peaks()
    ms_msm, _ = ctx.bench_msm(row, poly, 5, True)
    ms_kernel_alone = ctx.bench_last_kernel_ms()  # the dominant kernel with nothing else on the device
    ms_ntt = ctx.bench_ntt(n, 3, False)
    value = world * args.steps / (t_job * 1e-3)
    # device -> host per step: the bit-plane partial sums of both MSMs (192-byte XYZZ records, folded by ~40 host
    # point operations), the evaluation y and two status words
    d2h_bytes = 2 * (c * 192 + 4) + 32 + 4
    e2e_value = world * args.steps / (e2e_job * 1e-3)
    # dominant kernel: level-0 bucket accumulation, 10 Fq products per mixed addition, n*W additions
    acc_fq_muls = 10.0 * n * W
    achieved = acc_fq_muls / (ms_kernel_alone * 1e-3) / 1e9
    # ceiling = the better of the two live measurements: raw IMAD.WIDE issue rate / 300, or a dependent chain of
    # Fq products at full occupancy (the latter schedules the same instruction mix slightly better)
    peak = max(imad_peak / FQ_MUL_MACS, fq_chain_peak) / 1e9
    peak_source = ("fq_mul_chain (k_peak_fq_mul)" if fq_chain_peak >= imad_peak / FQ_MUL_MACS else "imad_wide / 300 (k_peak_imad_wide)")
    peaks_file = {}
    try:
        peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = peaks_file.get("hbm_gbs", 6650.0)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "accumulate_traffic.json"))).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_job / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": workload_config(log_n, world),
        "msm_plan": {"window_bits": c, "windows": W, "fixed_base_tables": True},
        "gpu_launches": int(launches) * args.steps,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 32 + 32,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_job / args.steps,
                "api": "zkp_worker_commit_open (C ABI; polynomial in page-locked host memory from zkp_host_alloc -> results on host)",
                "ms_per_step_pageable_input": e2e_pageable_ms,
                "ms_per_call_via_fourier_Client_list_of_base64_str": client_ms,
                "ms_via_fourier_Client_worker_commit_then_worker_open": client_two_calls_ms},
        "roofline": {"bound": "imad", "kernel": "k_accumulate<level0>", "achieved": achieved, "peak": peak,
                     "unit": "G Fq-mul/s", "frac": achieved / peak, "frac_executed": achieved / peak * 2712.0 / 3000.0,
                     "peak_source": peak_source + ", measured in this process on this device; MEASURED_PEAKS.json holds no integer peak",
                     "traffic": traffic, "traffic_source": "profiles/accumulate_traffic.json (ncu --set full capture, not re-measured per run)",
                     "note": "bound is INT32 multiply issue (IMAD.WIDE.U32, fmaheavy pipe), neither HBM nor tensor: "
                             "algorithmic work = 10 Fq products x 300 wide MACs per bucket addition (SURVEY 8d); the kernel "
                             "executes 2712 of those 3000 MACs (dedicated squaring, one fused two-product reduction), which is "
                             "what frac_executed counts; peak = IMAD.WIDE rate measured in this process / 300",
                     "executed_wide_macs_per_addition": 2712, "algorithmic_wide_macs_per_addition": 3000,
                     "imad_wide_per_s_measured": imad_peak, "fq_mul_chain_per_s_measured": fq_chain_peak,
                     "kernel_ms": ms_kernel_alone, "kernel_ms_in_step": ms_kernel_max,
                     "kernel_share_of_step": 2 * ms_kernel_alone / (t_job / args.steps),
                     "note2": "kernel_ms is the mean CUDA-event duration of the kernel inside single MSMs (device otherwise "
                              "idle); inside a step the commit and open MSMs overlap on two streams, so per-launch durations "
                              "there (kernel_ms_in_step) are stretched by sharing the SMs"},
        "msm": {"mpts_per_s": n / (ms_msm * 1e-3) / 1e6, "ms": ms_msm, "fq_muls": fq_muls,
                "fq_mul_per_s": fq_muls / (ms_msm * 1e-3), "frac_of_imad_peak": fq_muls / (ms_msm * 1e-3) / 1e9 / peak},
        "ntt": {"ms": ms_ntt, "achieved_gbs": 64.0 * n / (ms_ntt * 1e-3) / 1e9, "hbm_peak_gbs": hbm_peak,
                "frac_hbm": 64.0 * n / (ms_ntt * 1e-3) / 1e9 / hbm_peak,
                "fr_mul_frac_of_imad_peak": (n / 2 * log_n) * 136 / (ms_ntt * 1e-3) / imad_peak},
        "msm_sharded": msm24,
        "mgpu_in_library": mgpu,
        "combine_ms_per_step": t_comb_max / args.steps,
        "combine_note": "measured INSIDE the timed loop of `value` (and of e2e) at N > 1: every rank publishes its two partial "
                        "points as Jacobian coordinates (288 bytes, no inversion, no square root); rank 0 waits for all ranks, adds "
                        "2 x N points, compresses the two sums.  transport = " + args.combine +
                        " (host: POSIX shared memory -- the partials are host-resident, the window fold runs on the host; "
                        "nccl: H2D + all_gather + D2H of the same bytes)",
        "table_prebuild_s": t_tables,
        "checks": checks,
        "verified": bool(ok), "worker_verify_ms_per_call_host": verify_ms,
    }
    line.update(side)
    if msm24 is not None:
        # per-MSM accounting of the shard against the same ceiling
        msm24["frac_of_imad_peak_per_gpu"] = msm24["fq_mul_per_s_per_gpu"] / 1e9 / peak
    # ---- CPU baseline (oracle port) at the REAL size, rank 0 at N = 1 only
    if world == 1 and not args.no_cpu_baseline:
        from oracle import ref
        threads = os.cpu_count() or 1
        lg = args.cpu_baseline_log_n
        nn = 1 << lg
        if lg != log_n:
            ctx.srs_generate(TAU_X, TAU_Y, lg, 0)
        srs = ctx.srs_export_row(0, nn)  # the same SRS row as the GPU arm (exported, not recomputed on the CPU)
        bpoly = poly_bytes if lg == log_n else ref.random_scalars_ctr(SEED_POLY, 0, nn)
        reps, total = 0, 0.0
        while reps < 2 or (total < 9.0 and reps < 4):
            dt, ccom, cy, cproof = cpu_reference_sample(srs, bpoly, x, threads)
            total += dt
            reps += 1
        gres = (com, y, proof) if lg == log_n else ctx.worker_commit_open(0, bpoly, x)
        line["cpu_baseline"] = {
            "value": 1.0 / (total / reps), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"oracle/kzg_ref.c commit+open at n=2^{lg} (the workload itself, nothing extrapolated), {reps} reps on "
                      f"{threads} threads ({total / reps:.3f} s each); restatement, not blst",
            "matches_gpu": (ccom, cy, cproof) == tuple(gres)}
        assert line["cpu_baseline"]["matches_gpu"], "CPU restatement and GPU disagree"
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def mgpu_leg(native, args, world, log_n, log_m, x, all_parts, single, agg, msm24):
    """zkp_mgpu_*: the whole N-GPU job from ONE process (rank 0), timed by wall clock around the library call."""
    n = 1 << log_n
    out = {"devices": world, "api": "zkp_mgpu_pianist_commit_open / zkp_mgpu_msm_g1 / zkp_mgpu_commit_open (one process, one host "
                                    "thread per device inside the library, no torch / NCCL)"}
    K = max(5, min(args.steps, 20))
    with native.MultiContext(list(range(world))) as mg:
        mg.srs_generate(TAU_X, TAU_Y, log_n, log_m, native.LAYOUT_ROWS)
        mg.prebuild_tables()
        polys = native.PinnedBuffer(32 * n * world).write(mg.ctx(0).random_poly_range(SEED_POLY, 0, n * world))
        rows = list(range(world))
        res = mg.pianist_commit_open(rows, polys, x)  # uploads; warm
        ctxs = [mg.ctx(k) for k in range(world)]

        def flush_all():  # the L2 rule of the main loop: every device's L2 flushed between timed calls (outside the timing)
            for c in ctxs:
                c.bench_flush_l2()
        for _ in range(3):
            mg.pianist_commit_open(rows, polys, x, native.MGPU_RESIDENT)
        dt = 0.0
        for _ in range(K):
            flush_all()
            t0 = time.perf_counter()
            res_r = mg.pianist_commit_open(rows, polys, x, native.MGPU_RESIDENT)
            dt += time.perf_counter() - t0
        dt /= K
        dte = 0.0
        for _ in range(K):
            flush_all()
            t0 = time.perf_counter()
            res_e = mg.pianist_commit_open(rows, polys, x)
            dte += time.perf_counter() - t0
        dte /= K
        assert res == res_r == res_e
        if all_parts is not None:
            assert [c + y + p for c, y, p in zip(res[0], res[1], res[2])] == all_parts, "in-library and per-rank answers differ"
            assert (res[3], res[4]) == agg, "in-library aggregate differs from the torchrun aggregate"
        else:
            assert (res[0][0], res[1][0], res[2][0]) == single and (res[3], res[4]) == (single[0], single[2])
        out["pianist"] = {"commit_open_per_s_resident": world / dt, "ms_per_step_resident": dt * 1e3,
                          "commit_open_per_s_e2e_pinned": world / dte, "ms_per_step_e2e": dte * 1e3, "steps": K,
                          "matches_per_rank_run": True,
                          "note": "wall clock around each library call (L2 of every device flushed between calls, outside the "
                                  "timing), host-side fold + sum + compression included; "
                                  "resident = ZKP_MGPU_RESIDENT (no upload); e2e = N x 32 MiB from page-locked memory every step"}
        polys.close()
    # one polynomial split by point range over the N GPUs: 2^20 commit+open latency and the 2^24 MSM
    if world & (world - 1) == 0:
        with native.MultiContext(list(range(world))) as mg:
            mg.srs_generate(TAU_X, TAU_Y, log_n, 0, native.LAYOUT_POINT_RANGE)
            mg.prebuild_tables()
            p1 = native.PinnedBuffer(32 * n).write(mg.ctx(0).random_poly_range(SEED_POLY, 0, n))
            r0 = mg.commit_open(0, p1, x)
            ctxs = [mg.ctx(k) for k in range(world)]
            for _ in range(3):
                mg.commit_open(0, p1, x, native.MGPU_RESIDENT)
            dt = dte = 0.0
            for _ in range(K):
                for c in ctxs:
                    c.bench_flush_l2()
                t0 = time.perf_counter()
                r1 = mg.commit_open(0, p1, x, native.MGPU_RESIDENT)
                dt += time.perf_counter() - t0
            for _ in range(K):
                for c in ctxs:
                    c.bench_flush_l2()
                t0 = time.perf_counter()
                r2 = mg.commit_open(0, p1, x)
                dte += time.perf_counter() - t0
            dt, dte = dt / K, dte / K
            assert r0 == r1 == r2
            if world == 1 or log_m == 0:
                assert r0 == single, "point-range split differs from the single-GPU answer"
            out["commit_open_2p20_split"] = {"ms_resident": dt * 1e3, "ms_e2e_pinned": dte * 1e3, "commitment": r0[0].hex(),
                                             "note": "ONE 2^20 polynomial over N GPUs (latency); same bytes as one GPU"}
            p1.close()
        if msm24 is not None and not args.no_msm24:
            lg24 = args.msm_log_n
            with native.MultiContext(list(range(world))) as mg:
                mg.srs_generate(TAU_X, TAU_Y, lg24, 0, native.LAYOUT_POINT_RANGE)
                sc = native.PinnedBuffer(32 << lg24).write(mg.ctx(0).random_poly_range(SEED_MSM24, 0, 1 << lg24))
                c0 = mg.msm_g1(0, sc)  # uploads, builds the tables
                mg.msm_g1(0, sc, native.MGPU_RESIDENT)
                dt = 0.0
                for _ in range(3):
                    for k in range(world):
                        mg.ctx(k).bench_flush_l2()
                    t0 = time.perf_counter()
                    c1 = mg.msm_g1(0, sc, native.MGPU_RESIDENT)
                    dt += time.perf_counter() - t0
                dt /= 3
                assert c0 == c1 and c0.hex() == msm24["commitment"], "in-library sharded MSM differs from the per-rank run"
                out["msm_2p24"] = {"ms_resident": dt * 1e3, "mpts_per_s": (1 << lg24) / dt / 1e6, "matches_per_rank_run": True,
                                   "note": "wall clock around zkp_mgpu_msm_g1 (host fold, sum of N Jacobian partials and one "
                                           "compression included), scalars resident, L2 flushed between calls"}
                sc.close()
    return out


def side_measurements(native, ctx, local, log_n, log_m, row, poly, x, single, world, args):
    """rank 0: the live row size (2^16), the validator flow (2^12), pipelined requests, 2^24 commit+open"""
    import concurrent.futures
    out = {}
    # ---- BASELINE configs[1]: degree 2^16 (the mainnet row size: scale 24, machines_scale 8)
    lg2, batch = 16, 32
    n2 = 1 << lg2
    c16 = native.Context(local)
    c16.srs_generate(TAU_X, TAU_Y, lg2, 2)
    c16.prebuild_tables()
    pins = [native.PinnedBuffer(32 * n2).write(c16.random_poly_range(0xB200 + 2, k * n2, n2)) for k in range(batch)]
    xs16 = [c16.random_point(100 + k) for k in range(batch)]
    rows16 = [k % 4 for k in range(batch)]
    lat = {}
    for mode, name in ((1, "fused_one_launch_set"), (0, "two_lanes")):
        c16.set_fuse(mode)
        r16 = c16.worker_commit_open(0, pins[0], xs16[0])
        t0 = time.perf_counter()
        for _ in range(batch):
            r16b = c16.worker_commit_open(0, pins[0], xs16[0])
        lat[name] = (time.perf_counter() - t0) * 1e3 / batch
        assert r16 == r16b
    c16.set_fuse(-1)
    ms_msm16, _ = c16.bench_msm(0, pins[0], 10, True)
    # one launch set for the whole batch (zkp_worker_commit_open_batch), host buffers in, results on host
    xs_cat = b"".join(xs16)
    res_b = c16.worker_commit_open_batch(rows16, pins, xs_cat)
    t0 = time.perf_counter()
    reps = 4
    for _ in range(reps):
        res_b2 = c16.worker_commit_open_batch(rows16, pins, xs_cat)
    thr_batch = batch * reps / (time.perf_counter() - t0)
    assert res_b == res_b2 and all(st == 0 for st, *_ in res_b)
    assert tuple(res_b[0][1:]) == r16
    ok16 = all(c16.worker_verify(rows16[k], res_b[k][3], xs16[k], res_b[k][2], res_b[k][1]) for k in (0, 5, 31))
    # two forked contexts alternating batches: the upload and the host-side folds of one batch beside the kernels of the other
    f2 = [c16, c16.fork()]

    def work_b(k):
        o = None
        for _ in range(reps):
            o = f2[k].worker_commit_open_batch(rows16, pins, xs_cat)
        return o
    with concurrent.futures.ThreadPoolExecutor(2) as ex:
        list(ex.map(work_b, range(2)))
        t0 = time.perf_counter()
        outs = list(ex.map(work_b, range(2)))
        thr_batch2 = 2 * batch * reps / (time.perf_counter() - t0)
    assert outs[0] == res_b and outs[1] == res_b
    # a pool of 4 contexts serving single requests (what fourier.Client does for concurrent forwards)
    pool = [c16, f2[1], c16.fork(), c16.fork()]

    def work_s(k):
        o = None
        for j in range(batch // 4):
            o = pool[k].worker_commit_open(rows16[k], pins[k], xs16[k])
        return o
    with concurrent.futures.ThreadPoolExecutor(4) as ex:
        list(ex.map(work_s, range(4)))
        t0 = time.perf_counter()
        list(ex.map(work_s, range(4)))
        thr_pool = batch / (time.perf_counter() - t0)
    _, _, muls16 = c16.msm_info(n2)
    imad_peak, fq_chain = ctx.bench_peaks()
    peak = max(imad_peak / FQ_MUL_MACS, fq_chain)
    out["config_2p16"] = {
        "log_n": lg2, "latency_ms_per_commit_open": min(lat.values()), "latency_ms": lat, "msm_ms": ms_msm16,
        "commit_open_per_s_1_context": 1e3 / min(lat.values()),
        "commit_open_per_s_batch32_one_launch_set": thr_batch,
        "commit_open_per_s_batch32_two_contexts": thr_batch2,
        "commit_open_per_s_pool_of_4_contexts_single_requests": thr_pool,
        "batch_frac_of_imad_peak": 2 * muls16 * max(thr_batch, thr_batch2) / peak,
        "ceiling_commit_open_per_s_at_100pct": peak / (2 * muls16),
        "verified": bool(ok16),
        "note": "host buffers (pinned) in, results on host; rows 0..3 of a 4-row SRS; per-GPU figures of rank 0; "
                "frac = 2 MSMs x canonical Fq-muls x requests/s / measured multiply-issue ceiling"}
    for f in pool[1:]:
        f.close()
    for p in pins:
        p.close()
    c16.close()

    # ---- BASELINE configs[0] on the GPU: the validator flow at degree 2^12 (testnet row: scale 20, machines_scale 8),
    #      generate_challenge -> worker_commit -> worker_open -> worker_verify through fourier.Client (List[str] wire format)
    from zkp_subnet_b200.client import Client
    from zkp_subnet_b200.validator import Validator
    cl = Client(test_srs=True, device=local, seed=1234)
    t0 = time.perf_counter()
    cl.start(scale=14, machines_scale=2)
    t_start = time.perf_counter() - t0
    v = Validator(cl)
    flow = {}
    for it in range(3):
        t0 = time.perf_counter()
        ch = v.generate_challenge(4)
        t1 = time.perf_counter()
        resp = [cl.worker_commit_and_open(i, ch.polys[i], ch.alpha).json() for i in range(4)]
        t2 = time.perf_counter()
        valid = [cl.worker_verify(i, resp[i]["proof"], ch.alpha, ch.evals[i], resp[i]["commitment"]).json()["valid"] for i in range(4)]
        t3 = time.perf_counter()
        flow = {"generate_challenge_ms": (t1 - t0) * 1e3, "commit_open_ms_per_row": (t2 - t1) * 1e3 / 4,
                "verify_ms_per_row": (t3 - t2) * 1e3 / 4, "all_valid": all(valid)}
    bad = resp[0]["proof"][:-2] + ("A" if resp[0]["proof"][-2] != "A" else "B") + resp[0]["proof"][-1]
    flow["tampered_rejected"] = not cl.worker_verify(0, bad, ch.alpha, ch.evals[0], resp[0]["commitment"]).json()["valid"]
    flow["client_start_s"] = t_start
    flow["note"] = "4 rows x 2^12 through fourier.Client (base64 lists in and out), last of 3 iterations"
    out["config_2p12_validator_flow"] = flow
    cl.stop()

    # ---- the same 2^20 request stream through 3 contexts on 3 host threads (requests in flight overlap the
    #      reduction tail and the upload of one with the accumulation of another); host buffers in, results out
    nctx, per = 3, 6
    ctxs = [ctx] + [ctx.fork() for _ in range(nctx - 1)]

    def work3(k):
        o = None
        for _ in range(per):
            o = ctxs[k].worker_commit_open(row, poly, x)
        return o
    with concurrent.futures.ThreadPoolExecutor(nctx) as ex:
        list(ex.map(work3, range(nctx)))
        t0 = time.perf_counter()
        outs3 = list(ex.map(work3, range(nctx)))
        thr3 = nctx * per / (time.perf_counter() - t0)
    out["pipelined_2p20"] = {"contexts": nctx, "commit_open_per_s": thr3, "matches": all(o == single for o in outs3),
                             "note": "e2e (pinned host buffers in, results on host), 3 forked contexts sharing one SRS and one "
                                     "table arena; compare with e2e.value of this rank"}
    for f in ctxs[1:]:
        f.close()

    # ---- north_star: commit+open at degree 2^24 (SRS row 1.5 GiB, tables 51.5 GB), warm, pinned and pageable input
    if world == 1 and not args.no_2p24_open:
        lg = args.msm_log_n
        nn = 1 << lg
        c24 = native.Context(local)
        c24.srs_generate(TAU_X, TAU_Y, lg, 0)
        t0 = time.perf_counter()
        c24.prebuild_tables()
        t_tab = time.perf_counter() - t0
        big = native.PinnedBuffer(32 * nn)
        t0 = time.perf_counter()
        big.write(bytes(32 * nn))
        c24.worker_commit_open(0, big, x)  # what Client.start(precompute="eager") does: workspaces allocated by a zero polynomial
        t_warm = time.perf_counter() - t0
        big.write(c24.random_poly_range(SEED_MSM24, 0, nn))
        first_t0 = time.perf_counter()
        r_first = c24.worker_commit_open(0, big, x)
        first_ms = (time.perf_counter() - first_t0) * 1e3
        t0 = time.perf_counter()
        for _ in range(3):
            r_w = c24.worker_commit_open(0, big, x)
        warm_ms = (time.perf_counter() - t0) * 1e3 / 3
        pageable = big.tobytes()
        c24.worker_commit_open(0, pageable, x)
        t0 = time.perf_counter()
        for _ in range(2):
            r_p = c24.worker_commit_open(0, pageable, x)
        page_ms = (time.perf_counter() - t0) * 1e3 / 2
        ms_res, _, _, *r_res = c24.bench_commit_open(0, big, x, 2, True)
        ms_m24, _ = c24.bench_msm(0, big, 2, True)
        assert r_first == r_w == r_p == tuple(r_res) and c24.worker_verify(0, r_w[2], x, r_w[1], r_w[0])
        out["commit_open_2p24"] = {"log_n": lg, "ms_resident": ms_res, "ms_e2e_pinned_warm": warm_ms, "ms_e2e_pageable_warm": page_ms,
                                   "ms_first_request_after_eager_start": first_ms, "msm_ms": ms_m24, "table_prebuild_s": t_tab,
                                   "workspace_warm_s": t_warm,
                                   "commitment": r_w[0].hex(), "verified": True,
                                   "note": "one GPU; 512 MiB of evaluations per request; target <= 2 x MSM + 25 ms"}
        del pageable
        big.close()
        c24.close()
    return out


if __name__ == "__main__":
    sys.exit(main())
