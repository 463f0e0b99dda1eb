#!/usr/bin/env python
"""Top instructions by stall samples from an `ncu --page source --csv` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
ins = []
for k, r in enumerate(rows[2:]):
    try:
        ins.append((int(r[ix['# Samples']]), k, r[ix['Source']].strip(), int(r[ix['Instructions Executed']]),
                    {h: int(r[ix[h]]) for h in hdr if h.startswith('stall_') and '(' not in h and r[ix[h]] not in ('', '0')}))
    except Exception:
        pass
tot = sum(x[0] for x in ins)
for smp, k, src, ex, st in sorted(ins, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{100*smp/tot:5.1f}% @{k:5d} ex={ex:9d} {src[:64]:64s} {dict(sorted(st.items(), key=lambda x: -x[1])[:3])}")
