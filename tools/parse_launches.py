"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches of the last MSM."""
import csv, re, sys
for f in sys.argv[1:]:
    with open(f) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            name = re.sub(r"\(.*", "", r["Kernel Name"])
            name = re.sub(r"<.*", "", name) + ("<L0>" if re.search(r"k_accumulate<(\(bool\))?1\b", r["Kernel Name"]) else "")
            val = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"]
            val = val / 1e3 if unit == "ns" else (val * 1e3 if unit == "ms" else val)
            rows.append((name[:48], val))
    idx = [i for i, r in enumerate(rows) if "k_decompose" in r[0] and (i == 0 or "k_sort_scan" not in rows[i - 1][0])]
    last = rows[idx[-1]:] if idx else rows
    tot = sum(v for _, v in last)
    print(f"{f}: {len(rows)} launches; last MSM = {len(last)} launches, {tot:.1f} us")
    for name, v in last:
        print(f"  {name:48s} {v:10.1f} us  {100*v/tot:5.1f}%")
