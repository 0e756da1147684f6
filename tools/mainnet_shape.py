"""The reference's MAINNET command line end to end (Makefile:64-74: --scale 24 --machines_scale 8 --uncompressed true,
setup_24_8.uncompressed / precompute_24_8.uncompressed): write the two files with the setup CLI, start fourier.Client
on them, serve requests for 20 different worker indices (utils/config.py:237-242 samples 20 miners), verify every answer.
python tools/mainnet_shape.py [scale machines_scale] > profiles/r2_mainnet_shape.txt"""
import base64, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fourier import Client
from zkp_subnet_b200.client import encode_poly

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ms = int(sys.argv[2]) if len(sys.argv) > 2 else 8
d = tempfile.mkdtemp()
setup, pre = os.path.join(d, f"setup_{scale}_{ms}.uncompressed"), os.path.join(d, f"precompute_{scale}_{ms}.uncompressed")
t0 = time.perf_counter()
subprocess.check_call([sys.executable, "-m", "zkp_subnet_b200.setup", "--setup-path", setup, "--precompute-path", pre, "--scale", str(scale),
                       "--machines-scale", str(ms), "--generate-setup", "--generate-precompute", "--overwrite", "--uncompressed", "true"], cwd=ROOT)
print(f"setup CLI (generate-setup + generate-precompute, no trapdoor kept): {time.perf_counter() - t0:.1f} s; "
      f"files {os.path.getsize(setup)} + {os.path.getsize(pre)} bytes")
t0 = time.perf_counter()
c = Client(port=1337, bin="./prover", uncompressed="true", setup_path=setup, precompute_path=pre)
c.start(scale=scale, machines_scale=ms)
print(f"Client.start (load {1 << ms} rows, eager tables, warm-up): {time.perf_counter() - t0:.1f} s; source {c.srs_source}; "
      f"tables {c._need().table_stats()}")
n = 1 << (scale - ms)
ctx = c._need()
lat = []
for k in range(20):
    i = (k * 37) % (1 << ms)
    poly = encode_poly(ctx.random_poly(1000 + k, n))
    x = base64.b64encode(ctx.random_point(k)).decode().rstrip("=")
    t0 = time.perf_counter()
    com = c.worker_commit(i, poly).json()["commitment"]       # the reference miner's two calls (neurons/miner.py:56-61)
    o = c.worker_open(i, poly, x).json()
    lat.append((time.perf_counter() - t0) * 1e3)
    assert c.worker_verify(i, o["proof"], x, o["eval"], com).json()["valid"], i
    bad = o["proof"][:-2] + ("A" if o["proof"][-2] != "A" else "B") + o["proof"][-1]
    assert not c.worker_verify(i, bad, x, o["eval"], com).json()["valid"]
print(f"20 requests (worker_commit + worker_open through List[str], 20 different rows, all verified, tampered proofs rejected): "
      f"median {sorted(lat)[10]:.2f} ms, max {max(lat):.2f} ms per request")
c.stop()
os.remove(setup); os.remove(pre); os.rmdir(d)
