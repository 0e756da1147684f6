"""Host-side cost of the List[str] wire format at 2^LOG_N: decode into page-locked memory, encode, and the upload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << lg
ctx = native.Context(0)
raw = ctx.random_poly(1, n)
strs = native.wire_encode_list(raw)
pin = native.PinnedBuffer(32 * n)
def best(f, reps=7):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), sorted(ts)[len(ts) // 2]
print("decode into pinned  (min, median ms):", best(lambda: native.wire_decode_list(strs, pin)))
print("decode of 1/16 of the list          :", best(lambda: native.wire_decode_list(strs[: n // 16], pin)))
print("list slice [:n/16]                  :", best(lambda: strs[: n // 16]))
print("encode                              :", best(lambda: native.wire_encode_list(raw)))
print("cores:", os.cpu_count())
