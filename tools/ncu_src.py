#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: hot regions (runs of instructions with the same execution count),
opcode mix and stall samples.  usage: ncu_src.py file.csv [min_exec_fraction]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: ex = int(r[ix["Instructions Executed"]])
    except: continue
    ins.append((r[ix["Source"]].strip(), ex, int(r[ix["# Samples"]] or 0), int(r[ix["L1 Wavefronts Shared"]] or 0), int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)))
tot = sum(e for _, e, _, _, _ in ins); tots = sum(s for _, _, s, _, _ in ins)
print(f"static instrs {len(ins)}  executed warp-instr {tot:,}  samples {tots:,}")
# regions: consecutive instrs whose exec count within 2x of each other
regions = []; cur = []
for it in ins:
    if cur and not (0.7 < (it[1] + 1) / (cur[-1][1] + 1) < 1.43):
        regions.append(cur); cur = []
    cur.append(it)
if cur: regions.append(cur)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02
pos = 0
for reg in regions:
    e = sum(x[1] for x in reg); s = sum(x[2] for x in reg)
    if e / tot >= thr:
        ops = collections.Counter(x[0].split()[0] if not x[0].startswith('@') else x[0].split()[1] for x in reg)
        wv = sum(x[3] for x in reg); wi = sum(x[4] for x in reg)
        print(f"--- region @instr {pos}: {len(reg)} instrs, exec/instr ~{reg[0][1]:,}, {100*e/tot:.1f}% of executed, {100*s/max(1,tots):.1f}% of samples, smem wavefronts {wv:,} (ideal {wi:,})")
        print("    " + ", ".join(f"{k}:{v}" for k, v in ops.most_common(14)))
    pos += len(reg)
if len(sys.argv) > 3:
    a, b = map(int, sys.argv[3].split(':'))
    for k in range(a, b):
        print(k, ins[k][1], ins[k][2], ins[k][0])
