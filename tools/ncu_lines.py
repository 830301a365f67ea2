"""per-source-line warp-stall samples of an ncu report taken with --import-source on (tuning aid):
python tools/ncu_lines.py REPORT [top]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
H = rows[hdr]
iS = H.index("# Samples"); iL = 0; iSrc = 1; iEx = H.index("Instructions Executed"); iBar = H.index("stall_barrier")
cur = None; src = {}
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in rows[hdr + 1:]:
    if len(r) != len(H):
        continue
    if r[iL]:
        if not r[iL].isdigit():
            continue
        cur = int(r[iL]); src[cur] = r[iSrc]
    if r[iS] == "" or cur is None:
        continue
    try:
        agg[cur][0] += int(r[iS]); agg[cur][1] += int(r[iEx] or 0); agg[cur][2] += int(r[iBar] or 0)
    except ValueError:
        pass
tot = sum(v[0] for v in agg.values())
print("total samples", tot)
key = (lambda kv: -(kv[1][0] - kv[1][2])) if "--work" in sys.argv else (lambda kv: -kv[1][0])
work = sum(v[0] - v[2] for v in agg.values())
print("samples not waiting at a barrier", work)
for line, (s, ex, bar) in sorted(agg.items(), key=key)[:top]:
    print(f"{line:5d} {s/tot:6.1%} work {(s-bar)/max(work,1):6.1%} barrier {bar/max(s,1):4.0%} inst {ex:>11d}  {src.get(line, '').strip()[:110]}")
