"""Top stalled SASS instructions of an ncu report's source page: python tools/ncu_stalls.py report.sass.csv [N] [kernel index]
(csv from: ncu -i report.ncu-rep --page source --csv --print-source sass; one block per profiled launch)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
rows = rows[starts[which]:starts[which + 1]]
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2]
val = lambda r, h: int(float(r[ix[h]] or 0)) if ix[h] < len(r) else 0
tot = sum(val(r, "# Samples") for r in data)
print("kernel", rows[0][1][:120])
print("total samples", tot, "instructions", len(data))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(val(r, h) for r in data) for h in stall_cols}
print("by reason:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
order = sorted(range(len(data)), key=lambda i: -val(data[i], "# Samples"))[:n]
for i in order:
    r = data[i]
    st = {h[6:]: val(r, h) for h in stall_cols if val(r, h) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{val(r, '# Samples'):6d} #{i:<5d} {r[ix['Source']][:72]:72s} {st}")
