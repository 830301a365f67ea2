"""per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list"""
import collections, csv, re, sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
    agg[re.sub(r"\(.*", "", row["Kernel Name"])[:60]].append(v)
tot = sum(sum(v) for v in agg.values())
print(f"total {tot:.1f} us over {sum(len(v) for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print(f"{k:60s} n={len(v):5d} sum={sum(v):9.1f} us ({100 * sum(v) / tot:4.1f}%) mean={sum(v) / len(v):8.2f} max={max(v):8.2f}")
