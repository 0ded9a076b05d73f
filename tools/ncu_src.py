"""Summarise an `ncu --page source --csv` dump: top SASS lines by stall samples.
usage: python tools/ncu_src.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
gi = lambda r, k: int(float(r[ix[k]] or 0))
tot = sum(gi(r, 'Instructions Executed') for r in data)
tots = sum(gi(r, '# Samples') for r in data)
print('total warp-inst', tot, 'samples', tots)
stalls = [h for h in hdr if h.startswith('stall_') and '(' not in h]
agg = {s: sum(gi(r, s) for r in data) for s in stalls}
print('stall mix:', {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for r in sorted(data, key=lambda r: -gi(r, '# Samples'))[:N]:
    st = {s[6:]: gi(r, s) for s in stalls if gi(r, s)}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(str(gi(r, '# Samples')).rjust(7), str(gi(r, 'Instructions Executed')).rjust(10), r[ix['Source']][:100].ljust(100), top)
