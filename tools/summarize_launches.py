"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total and mean device
time, share of the total.  usage: python tools/summarize_launches.py launches.csv"""
import csv, re, sys, collections
for f in sys.argv[1:]:
    lines = [l for l in open(f) if not l.startswith("==")]
    tot = collections.defaultdict(float); cnt = collections.Counter()
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        if "k_accumulate" in short:
            short = "zkp::k_accumulate<level0>" if re.search(r"k_accumulate<(\(bool\))?1\b", name) else "zkp::k_accumulate<slots>"
        elif "k_affine_round" in short:
            short = "zkp::k_affine_round<first>" if re.search(r"k_affine_round<(\(bool\))?1\b", name) else "zkp::k_affine_round<next>"
        else:
            short = re.sub(r"<.*", "", short)
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        tot[short] += v; cnt[short] += 1
    total = sum(tot.values())
    print(f"{f}: {sum(cnt.values())} launches, {total / 1e3:.2f} ms of device time (cold-cache, serialised: compare SHARES)")
    print(f"  {'kernel':44s} {'launches':>8s} {'total ms':>10s} {'mean us':>10s} {'share':>7s}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"  {k[:44]:44s} {cnt[k]:8d} {v / 1e3:10.3f} {v / cnt[k]:10.1f} {100 * v / total:6.1f}%")
