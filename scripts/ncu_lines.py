"""Attribute an `ncu --page source --csv` SASS dump to CUDA source lines using nvdisasm -g line info.
usage: ncu_lines.py <src.csv> <libsrfe.so> [topN]"""
import csv, sys, re, subprocess, collections, os, tempfile, glob
src_csv, so = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(src_csv)))
kname = rows[0][1]
import itertools
def mangle(kname):
    # srfe_kernel<(int)512, (int)2, ..., (unsigned int)41984, (int)5, short> -> Itanium template-args fragment
    mk = re.search(r"(srfe_\w*kernel)<(.*)>\(", kname)
    args = mk.group(2).split(", ")
    out = ""
    for a in args:
        m = re.match(r"\((unsigned int|int|bool)\)(\d+)", a)
        if m: out += "L" + {"int": "i", "unsigned int": "j", "bool": "b"}[m.group(1)] + m.group(2) + "E"
        else: out += {"float": "f", "short": "s"}[a]
    return mk.group(1) + "I" + out + "E"
mangled = mangle(kname)
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]; body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
cub = [f for f in glob.glob(d + "/*.cubin") if "tables" not in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.splitlines()
lines, cur, infn = [], ("?", 0), False
for ln in dis:
    if ln.startswith(".text."):
        infn = mangled in ln
        continue
    if not infn: continue
    mm = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
    if mm: cur = (os.path.basename(mm.group(1)), int(mm.group(2))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln): lines.append(cur)
assert len(lines) >= len(body), (len(lines), len(body))
agg = collections.defaultdict(lambda: [0.0, 0.0])
tot_i = tot_s = 0.0
for r, loc in zip(body, lines):
    i = float(r[ci["Instructions Executed"]] or 0); s = float(r[ci["# Samples"]] or 0)
    agg[loc][0] += i; agg[loc][1] += s; tot_i += i; tot_s += s
srcs = {}
def text(loc):
    f, n = loc
    for base in ("speechrecognitionproject_b200/csrc/", ""):
        p = base + f
        if os.path.exists(p):
            if p not in srcs: srcs[p] = open(p).read().splitlines()
            return srcs[p][n - 1].strip()[:90] if n - 1 < len(srcs[p]) else ""
    return ""
print(f"{kname}\ninstr {tot_i:.0f} samples {tot_s:.0f}")
for loc, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100*i/tot_i:5.1f}%i {100*s/max(tot_s,1):5.1f}%s  {loc[0]}:{loc[1]:<4d} {text(loc)}")
if os.environ.get("REGIONS"):
    # address-ordered view: each SASS instruction is charged to the most recent srfe_kernels.cuh line (inlined helpers
    # inherit their caller); consecutive instructions in the same 10-line bucket are merged
    print("\naddress-ordered regions (sticky srfe_kernels.cuh line, merged per 10-line bucket; >= 0.3 % of samples)")
    sticky, runs = 0, []
    for r, loc in zip(body, lines):
        if loc[0] == os.environ.get("REGION_FILE", "srfe_kernels.cuh"): sticky = loc[1]
        b = sticky // 10
        i = float(r[ci["Instructions Executed"]] or 0); s = float(r[ci["# Samples"]] or 0)
        if runs and runs[-1][0] == b: runs[-1][1] += i; runs[-1][2] += s; runs[-1][3] += 1
        else: runs.append([b, i, s, 1])
    for b, i, s, n in runs:
        if s / tot_s >= 0.003: print(f"  lines {10*b:4d}-{10*b+9:<4d} sass {n:5d}  {100*i/tot_i:5.1f}%i {100*s/tot_s:5.1f}%s")
