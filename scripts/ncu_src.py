"""Summarise an `ncu --page source --csv` dump: instructions by opcode, stalls by reason, hot SASS."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]; body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    try: return float(r[ci[k]])
    except Exception: return 0.0
tot = sum(num(r, "Instructions Executed") for r in body)
ops = collections.Counter(); samp = collections.Counter()
for r in body:
    op = r[ci["Source"]].split()
    op = [o for o in op if not o.startswith("@")]
    name = op[0].split(".")[0] if op else "?"
    full = ".".join(op[0].split(".")[:2]) if op else "?"
    ops[full if name in ("LDS", "STS", "LDG", "STG", "SHFL", "BAR") else name] += num(r, "Instructions Executed")
    samp[name] += num(r, "# Samples")
print(f"total warp-instr {tot:.0f}  ({len(body)} SASS lines)")
for k, v in ops.most_common(28): print(f"  {k:14s} {v:12.0f} {100*v/tot:5.1f}%")
stalls = [h for h in hdr if h.startswith("stall_")]
st = {s: sum(num(r, s) for r in body) for s in stalls}
ts = sum(st.values()) or 1
print("stall samples:", ", ".join(f"{k[6:]}={100*v/ts:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
print("shared wavefronts", sum(num(r, "L1 Wavefronts Shared") for r in body), "ideal", sum(num(r, "L1 Wavefronts Shared Ideal") for r in body))
if len(sys.argv) > 2:
    n = int(sys.argv[2])
    print("hottest SASS by samples:")
    for r in sorted(body, key=lambda r: -num(r, "# Samples"))[:n]:
        print(f"  {num(r,'# Samples'):7.0f} {num(r,'Instructions Executed'):10.0f}  {r[ci['Source']][:90]}")
