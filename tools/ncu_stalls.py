"""Summarise an `ncu --page source --csv` export: top stall sites (SASS) with their dominant stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
iS = hdr.index('# Samples'); isrc = hdr.index('Source')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []; tot = 0
for ln, r in enumerate(rows[2:]):
    try: s = int(r[iS])
    except Exception: continue
    tot += s; data.append((s, ln, r))
print(rows[0][1][:100]); print('total samples', tot)
for s, ln, r in sorted(data, key=lambda x: -x[0])[:top_n]:
    st = {hdr[i][6:]: int(r[i]) for i in stalls if r[i] not in ('', '0')}
    top = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(f"{s:6d} {100*s/tot:5.1f}% L{ln:5d} {r[isrc].strip()[:64]:64s} {top}")
