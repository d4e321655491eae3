"""SASS opcode histogram per kernel of libsrfe.so (cuobjdump -sass) -> profiles/ evidence of what the kernels are made of.
usage: python scripts/sass_histogram.py [libsrfe.so] > profiles/rN_sass_histogram.txt"""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "speechrecognitionproject_b200/libsrfe.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
cur, hist = None, collections.OrderedDict()
for ln in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = demangle(m.group(1)); hist[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", ln)
    if m and cur:
        hist[cur][m.group(1)] += 1
KEY = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "SYNCS", "FFMA2", "FADD2", "FMUL2", "LDS", "STS", "LDG", "STG", "SHFL", "MUFU", "BAR", "ATOMS")
print("# static SASS instruction counts per kernel (prefix match on the mnemonic); Blackwell evidence: UTC*MMA = tcgen05.mma,")
print("# LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, FFMA2/FADD2/FMUL2 = f32x2")
for fn, h in hist.items():
    tot = sum(h.values())
    keyed = {k: sum(v for op, v in h.items() if op.startswith(k)) for k in KEY}
    print(f"\n{fn}\n  total {tot}  " + "  ".join(f"{k}={v}" for k, v in keyed.items() if v))
    print("  top: " + ", ".join(f"{op} {v}" for op, v in h.most_common(12)))
