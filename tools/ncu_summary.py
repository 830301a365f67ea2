"""summarise an ncu report: key raw metrics + SASS stall samples by code region (tuning aid)"""
import csv, subprocess, sys, io
rep = sys.argv[1]; W = int(sys.argv[2]) if len(sys.argv) > 2 else 150
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); H = rows[0]; r = rows[2]
for k in ['gpu__time_duration.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']:
    if k in H: print(f"{k:75s} {r[H.index(k)]}")
for i,h in enumerate(H):
    if 'tensor' in h and 'pct' in h and r[i] not in ('','0'): print(f"{h:75s} {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); H = rows[1]
blocks = [i for i, x in enumerate(rows) if x and x[0] == 'Kernel Name']
end = blocks[1] if len(blocks) > 1 else None
data = [x for x in rows[2:end] if len(x) == len(H)]
iS = H.index('# Samples'); iSrc = H.index('Source'); iEx = H.index('Instructions Executed')
cols = {k: H.index(k) for k in H if k.startswith('stall_') and 'Not Issued' not in k}
tot = sum(int(x[iS]) for x in data); print('total samples', tot, 'n instr', len(data))
def op(s):
    t = s.strip().split(); return t[1] if t[0].startswith('@') else t[0]
for a in range(0, len(data), W):
    ch = data[a:a + W]; s = sum(int(x[iS]) for x in ch)
    if s < tot * 0.02: continue
    ops = {}
    for x in ch:
        o = op(x[iSrc]); ops[o] = ops.get(o, 0) + int(x[iEx])
    st = {k: sum(int(x[c]) for x in ch) for k, c in cols.items()}
    print(f"[{a:5d}] {s/tot:6.1%}", {k[6:]: v for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4]}, sorted(ops.items(), key=lambda kv: -kv[1])[:5])
