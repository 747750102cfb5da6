"""Top stall sites of one kernel from `ncu --page source --csv` output (run here, no GPU):
   ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<pat> > src.csv ; python scripts/ncu_hot.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and (r[ix["# Samples"]] or "0").isdigit()]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
print("total samples", tot)
agg = {h: 0 for h in stall_cols}
for r in data:
  for h in stall_cols:
    agg[h] += int(r[ix[h]] or 0)
print("by reason:", ", ".join(f"{h[6:]} {100*v/tot:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:N]
for r in sorted(top, key=lambda r: int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]] or 0)):
  s = int(r[ix["# Samples"]] or 0)
  reasons = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
  print(f"{r[ix['Address']]:>8} {100*s/tot:5.1f}%  {r[ix['Source']][:70]:70s} " + " ".join(f"{n}:{c}" for c, n in reasons if c))
