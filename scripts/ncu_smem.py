"""Shared-memory wavefronts by instruction type and by source region from an `ncu --page source --csv` dump.
usage: ncu_smem.py <src.csv> <n_clips>"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1]))); nclips = float(sys.argv[2])
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for r in body:
    op = [o for o in r[ci['Source']].split() if not o.startswith('@')][0]
    if op.startswith(('LDS', 'STS', 'LD.', 'ST.', 'ATOMS')):
        ex = float(r[ci['Instructions Executed']] or 0); wf = float(r[ci['L1 Wavefronts Shared']] or 0)
        ideal = float(r[ci['L1 Wavefronts Shared Ideal']] or 0)
        a = agg[op]; a[0] += ex; a[1] += wf; a[2] += ideal; a[3] += 1
tot = 0
for op, (ex, wf, ideal, n) in sorted(agg.items()):
    if ex == 0: continue
    tot += wf
    print(f"{op:10s} sass={n:4d} exec/clip={ex/nclips:8.1f} wf/instr={wf/ex:5.2f} ideal={ideal/ex:5.2f} wf/clip={wf/nclips:8.0f}")
print(f"total wavefronts / clip {tot/nclips:.0f}")
