"""Summarise an `ncu --page source --csv` dump: stall reasons overall and the hottest SASS instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0 for h in stall_cols}
recs = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: ns = int(r[idx["# Samples"]])
    except ValueError: continue
    st = {}
    for h in stall_cols:
        try: v = int(r[idx[h]])
        except ValueError: v = 0
        tot[h] += v; st[h] = v
    recs.append((ns, r[idx["Address"]], r[idx["Source"]], st, r[idx["Instructions Executed"]]))
allsamp = sum(tot.values())
print("stall totals:")
for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v: print(f"  {h:28s} {v:8d} {100*v/allsamp:5.1f}%")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(f"top {n} instructions by samples:")
for ns, addr, src, st, ie in sorted(recs, key=lambda x: -x[0])[:n]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"  {ns:6d} {addr[-6:]} exec={ie:>9s} {src[:90]:90s} {top[0][0][6:]}={top[0][1]} {top[1][0][6:]}={top[1][1]}")
