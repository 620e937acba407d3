"""Top stall sites of an `ncu --page source --csv` dump (SASS view): python tools/ncu_top_stalls.py dump.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
total = sum(int(r[ci["# Samples"]] or 0) for r in body)
print(f"{rows[0][1]}: {total} samples, {len(body)} SASS instructions")
order = sorted(range(len(body)), key=lambda i: -int(body[i][ci["# Samples"]] or 0))[:top]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for i in sorted(order):
    r = body[i]
    n = int(r[ci["# Samples"]] or 0)
    why = sorted(((int(r[ci[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {100.0 * n / total:5.1f}%  {r[ci['Source']].strip()[:90]:90s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
