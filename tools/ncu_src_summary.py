"""Summarise an `ncu --page source --csv --print-source sass` dump: executed instructions by opcode,
stall samples by reason, and the hottest code regions.  Usage: python tools/ncu_src_summary.py dump.csv [frames]"""
import csv, sys, collections, re
path = sys.argv[1]; frames = float(sys.argv[2]) if len(sys.argv) > 2 else None
rows = list(csv.reader(open(path)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); stall = collections.Counter(); total = 0; samples = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
per = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[idx["Source"]].strip(); ex = int(r[idx["Instructions Executed"]] or 0); smp = int(r[idx["# Samples"]] or 0)
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
    op = m.group(2) if m else "?"
    ops[op] += ex; total += ex; samples += smp
    for c in stall_cols: stall[c] += int(r[idx[c]] or 0)
    per.append((ex, smp, src))
div = frames or 1.0
print(f"total warp-instr executed {total}  per-frame {total/div:.1f}   samples {samples}")
for op, c in ops.most_common(28): print(f"  {op:12s} {c:12d}  {c/div:9.1f}  {100*c/total:5.1f}%")
print("stall samples:")
for k, c in stall.most_common(12): print(f"  {k:28s} {c:8d} {100*c/max(1,samples):5.1f}%")
# hottest windows of 40 instructions by samples
W = 40; best = []
for i in range(0, len(per), W):
    blk = per[i:i+W]; best.append((sum(b[1] for b in blk), sum(b[0] for b in blk), i))
print("hottest 40-instruction windows (samples, executed, start index, first instr):")
for s, e, i in sorted(best, reverse=True)[:12]: print(f"  {s:7d} {e:10d} @{i:5d}  {per[i][2][:70]}")
