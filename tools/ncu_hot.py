"""Aggregate an `ncu --page source --csv` dump: top SASS instructions by stall samples and executed counts."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot_s = sum(int(r[ix["# Samples"]] or 0) for r in data)
tot_i = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print(f"total samples {tot_s}, warp instructions executed {tot_i}")
top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for r in top:
    s = int(r[ix["# Samples"]]); 
    stalls = {k[6:]: int(r[ix[k]] or 0) for k in ix if k.startswith("stall_") and "(" not in k}
    main = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    print(f"{s:6d} {100*s/tot_s:5.1f}%  exec {int(r[ix['Instructions Executed']]):8d}  {r[ix['Source']].strip()[:70]:70s} {main}")
by_op = collections.Counter(); by_op_s = collections.Counter()
for r in data:
    op = r[ix["Source"]].strip().split()
    op = [o for o in op if not o.startswith("@")][0].split(".")[0] if op else "?"
    by_op[op] += int(r[ix["Instructions Executed"]] or 0); by_op_s[op] += int(r[ix["# Samples"]] or 0)
print("by opcode: executed share / sample share")
for op, c in by_op.most_common(25):
    print(f"  {op:10s} {100*c/tot_i:5.1f}%  {100*by_op_s[op]/tot_s:5.1f}%")
