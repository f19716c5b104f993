"""Aggregate an ncu source-page CSV of encoder_kernel by source function (uses nvdisasm line info of the built object).
usage: python tools/ncu_regions.py report.ncu-rep"""
import csv, re, collections, subprocess, sys, os, tempfile
rep = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "adafortitran_b200/build/tc_encoder.o")], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
infn = False; cur = None; a2l = {}
for l in sass:
    if ".text." in l and (".section" in l or l.startswith(".text.")): infn = "encoder_kernel" in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2)))
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if infn and m: a2l[int(m.group(1), 16)] = cur
# function table of tc_encoder.cu: start line -> name
src = open(os.path.join(root, "adafortitran_b200/csrc/tc_encoder.cu")).read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:template.*\n)?__device__ __forceinline__ \S+(?: \S+)? (\w+)\(", l) or re.match(r"^__global__ .* (\w+)\(", l)
    if m: funcs.append((i, m.group(1)))
def fn_of(line):
    name = "?"
    for st, n in funcs:
        if st <= line: name = n
        else: break
    return name
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = rows[1]
ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = None; agg = collections.Counter(); inst = collections.Counter(); tot = 0; lines = collections.Counter()
kern_lines = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr): continue
    a = int(r[ia], 16)
    if base is None: base = a
    ln = a2l.get(a - base)
    s = int(r[isamp] or 0); n = int(r[iinst] or 0)
    tot += s
    if ln is None: key = "(none)"
    elif ln[0] == "tc_encoder.cu":
        key = fn_of(ln[1])
        if key == "encoder_kernel": kern_lines[ln[1]] += s
    else: key = ln[0] + ":" + str(ln[1])
    agg[key] += s; inst[key] += n
print("total samples", tot)
for k, v in agg.most_common(40): print(f"{v / tot:6.2%}  inst {inst[k]:>12}  {k}")
print("--- encoder_kernel body lines")
for k, v in kern_lines.most_common(25): print(f"{v / tot:6.2%}  line {k}: {src[k - 1].strip()[:110]}")
