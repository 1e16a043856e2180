"""Summarise an `ncu --page source --csv` dump: top SASS instructions by executed count / stall samples,
and totals per opcode.  Usage: python tools/ncu_src_summary.py file.csv [topN]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
def f(r, n):
    try:
        return float(r[col[n]])
    except Exception:
        return 0.0
tot_inst = sum(f(r, "Instructions Executed") for r in data)
tot_samp = sum(f(r, "# Samples") for r in data)
print(f"SASS lines {len(data)}  warp-instructions {tot_inst:.0f}  samples {tot_samp:.0f}")
ops = Counter(); samp = Counter()
for r in data:
    op = r[col["Source"]].split()
    name = op[0] if op and not op[0].startswith("@") else (op[1] if len(op) > 1 else "?")
    name = name.split(".")[0]
    ops[name] += f(r, "Instructions Executed"); samp[name] += f(r, "# Samples")
print("\nby opcode (share of instructions | share of stall samples)")
for name, c in ops.most_common(28):
    print(f"  {name:12s} {100*c/tot_inst:6.2f}%   {100*samp[name]/max(tot_samp,1):6.2f}%")
stall_cols = [n for n in hdr if n.startswith("stall_")]
tot = Counter()
for r in data:
    for n in stall_cols:
        tot[n] += f(r, n)
s = sum(tot.values())
print("\nstall reasons (all samples)")
for n, c in tot.most_common(12):
    print(f"  {n:28s} {100*c/max(s,1):6.2f}%")
print(f"\ntop {top} lines by samples")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top]:
    st = (sorted(((f(r, n), n) for n in stall_cols), reverse=True) + [(0.0, "stall_-"), (0.0, "stall_-")])[:2]
    print(f"  {f(r,'# Samples'):7.0f}  {f(r,'Instructions Executed'):10.0f}  {r[col['Source']][:70]:70s} {st[0][1][6:]}:{st[0][0]:.0f} {st[1][1][6:]}:{st[1][0]:.0f}")
