import csv, subprocess, sys, io, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
H = rows[hdr]
iS = H.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(H) if h.startswith("stall_") and "not_issued" not in h]
cur=None; src={}
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in rows[hdr+1:]:
    if len(r) != len(H): continue
    if r[0]:
        if not r[0].isdigit(): continue
        cur = int(r[0]); src[cur] = r[1]
    if cur is None or r[iS] == "": continue
    for i,h in stall_cols:
        try: v=int(r[i] or 0)
        except ValueError: v=0
        agg[cur][h]+=v; tot[h]+=v
T=sum(tot.values())
print("all:", {k[6:]: round(100*v/T,1) for k,v in tot.most_common(12)})
lines = sorted(agg.items(), key=lambda kv:-sum(kv[1].values()))[:int(sys.argv[2]) if len(sys.argv)>2 else 14]
for line,c in lines:
    s=sum(c.values())
    print(f"{line:4d} {100*s/T:5.1f}% ", {k[6:]: round(100*v/T,1) for k,v in c.most_common(5)}, src[line].strip()[:70])
