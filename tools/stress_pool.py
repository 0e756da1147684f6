"""Concurrency stress of the client pool and the shared table arena: 12 threads over 4 pooled contexts, 8 SRS rows but room
for only 3 fixed-base tables (so tables are evicted, rebuilt and sometimes unavailable WHILE other requests hold theirs),
every answer compared with the single-threaded one.  python tools/stress_pool.py [seconds]"""
import base64, os, random, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fourier import Client
from zkp_subnet_b200.client import encode_poly

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 30
scale, ms = 15, 3
n, rows = 1 << (scale - ms), 1 << ms
c = Client(test_srs=True, contexts=4, precompute="lazy")
c.start(scale=scale, machines_scale=ms)
root = c._need()
cw, W, _ = root.msm_info(n)
root.set_table_budget(3 * 2 * W * n * 128 + 4096)
polys = [encode_poly(root.random_poly(100 + k, n)) for k in range(6)]
xs = [base64.b64encode(root.random_point(k)).decode().rstrip("=") for k in range(4)]
expect = {}
for r in range(rows):
    for k in range(6):
        for j in range(4):
            if (r + k + j) % 3 == 0:
                expect[(r, k, j)] = c.worker_commit_and_open(r, polys[k], xs[j]).json()
keys = list(expect)
errors, counts = [], [0] * 12
stop = time.time() + secs

def work(t):
    rng = random.Random(t)
    while time.time() < stop and not errors:
        r, k, j = rng.choice(keys)
        want = expect[(r, k, j)]
        mode = rng.randrange(4)
        try:
            if mode == 0:
                got = c.worker_commit_and_open(r, polys[k], xs[j]).json()
                ok = got == want
            elif mode == 1:
                com = c.worker_commit(r, polys[k]).json()["commitment"]
                o = c.worker_open(r, polys[k], xs[j]).json()
                ok = (com, o["eval"], o["proof"]) == (want["commitment"], want["eval"], want["proof"])
            elif mode == 2:
                ok = c.worker_verify(r, want["proof"], xs[j], want["eval"], want["commitment"]).json()["valid"] is True
            else:
                items = [{"i": r, "poly": polys[k], "alpha": xs[j]}]
                r2, k2, j2 = rng.choice(keys)
                items.append({"i": r2, "poly": polys[k2], "alpha": xs[j2]})
                out = c.worker_commit_and_open_batch(items).json()["results"]
                ok = out[0] == want and out[1] == expect[(r2, k2, j2)]
            if not ok:
                errors.append((t, mode, r, k, j))
        except Exception as e:  # noqa: BLE001
            errors.append((t, mode, repr(e)))
        counts[t] += 1

ths = [threading.Thread(target=work, args=(t,)) for t in range(12)]
[t.start() for t in ths]
[t.join() for t in ths]
st = root.table_stats()
print(f"stress: {sum(counts)} operations from 12 threads in {secs:.0f} s over 4 contexts, arena {st}; errors: {errors[:3] if errors else 0}")
c.stop()
sys.exit(1 if errors else 0)
