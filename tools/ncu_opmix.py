#!/usr/bin/env python3
"""Opcode histogram and stall summary of one kernel from `ncu -i X.ncu-rep --page source --csv`.
usage: ncu_opmix.py source.csv units   (units = attempts or points per launch, for per-unit counts)"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ix = {n: i for i, n in enumerate(hdr)}
ops, thr, smp = collections.Counter(), collections.Counter(), collections.Counter()
stalls = collections.Counter()
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
tot = tott = 0
for r in rows[h + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    if not m:
        continue
    op = m.group(2)
    n, t = int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]])
    ops[op] += n
    thr[op] += t
    smp[op] += int(r[ix["# Samples"]])
    tot += n
    tott += t
    for c in stall_cols:
        stalls[hdr[c]] += int(r[c])
print("warp instructions %d, thread instructions %d, lanes/instr %.2f, thread instr per unit %.1f"
      % (tot, tott, tott / tot, tott / units))
for op, n in thr.most_common(30):
    print("%-10s thread-instr/unit %7.1f   warp-instr %12d   samples %d" % (op, n / units, ops[op], smp[op]))
s = sum(stalls.values())
print("stall samples:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / s) for k, v in stalls.most_common(8)))
