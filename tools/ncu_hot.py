"""Top stall locations of an ncu source-page CSV (ncu -i x.ncu-rep --page source --csv > x.csv).
    python tools/ncu_hot.py x.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
ranked = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:n]
for i in sorted(ranked):
    r = body[i]
    s = int(r[ix["# Samples"]] or 0)
    top = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {100*s/tot:5.1f}% ex={r[ix['Instructions Executed']]:>9s} {r[ix['Source']].strip()[:90]:90s} {' '.join(f'{c}:{v}' for v,c in top if v)}")
