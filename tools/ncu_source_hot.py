"""Aggregate an ncu `--page source --csv` dump: warp instructions and mean active threads per contiguous SASS region / per opcode."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
# find header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isrc, iex, ith = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
ist = hdr.index("# Samples")
tot_i = tot_t = 0
per_op = collections.defaultdict(lambda: [0, 0, 0])
insts = []
for r in rows[hi + 1:]:
    if len(r) <= ith or not r[ia].startswith("0x"):
        continue
    try:
        ex, th, st = int(r[iex]), int(r[ith]), int(r[ist])
    except ValueError:
        continue
    op = r[isrc].strip().split()[0] if r[isrc].strip() else "?"
    if op.startswith("@"):
        op = r[isrc].strip().split()[1]
    op = op.split(".")[0]
    per_op[op][0] += ex; per_op[op][1] += th; per_op[op][2] += st
    tot_i += ex; tot_t += th
    insts.append((r[ia], r[isrc].strip(), ex, th, st))
print(f"total warp-inst {tot_i:,}  thread-inst {tot_t:,}  avg active {tot_t/max(tot_i,1):.2f}")
print("by opcode (top 25 by warp instructions):")
for op, (ex, th, st) in sorted(per_op.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"  {op:12s} {ex:>14,} ({100*ex/tot_i:5.1f}%)  active {th/max(ex,1):5.2f}  samples {st}")
# regions: split the instruction stream into chunks of 24 and show the heavy ones
print("hot regions (48-instruction windows):")
W = 48
wins = []
for k in range(0, len(insts), W):
    chunk = insts[k:k + W]
    ex = sum(c[2] for c in chunk); th = sum(c[3] for c in chunk); st = sum(c[4] for c in chunk)
    wins.append((k, ex, th, st, chunk))
for k, ex, th, st, chunk in wins:
    ops = collections.Counter(c[1].split()[0 if not c[1].startswith('@') else 1].split('.')[0] for c in chunk)
    print(f"  inst {k:5d}-{k+len(chunk):5d}: {ex:>13,} warp-inst ({100*ex/tot_i:5.1f}%) active {th/max(ex,1):5.2f} samples {st:6d}  {dict(ops.most_common(5))}")
