"""Maps the per-instruction samples of an ncu report (--page source --csv) onto CUDA source lines
using nvdisasm -g line info of the matching cubin.
usage: python tools/ncu_lines.py report.ncu-rep lib.so kernel_substring [top_n]"""
import csv, io, re, subprocess, sys, tempfile, os, glob
rep, so, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if "# Samples" in r)
isamp, isrc = hdr.index("# Samples"), hdr.index("Source")
# the CSV holds one section per captured launch ("Kernel Name", name): take the first match
want = kname.replace("ILi", "<").split("<")[0].split("_ZN3bgp")[-1].lstrip("0123456789")
inst, take = [], False
for r in rows:
    if r and r[0] == "Kernel Name":
        if take and inst:
            break
        take = want in r[1]
        continue
    if take and len(r) > isamp and r[isamp].isdigit():
        inst.append((r[isrc].strip(), int(r[isamp])))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
lines = None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout
    if kname in txt:
        # take the .text section of the kernel
        m = re.search(r"\.section\s+\.text\.[^\n]*" + re.escape(kname) + r"[^\n]*\n(.*?)(?=\n\s*\.section|\Z)", txt, re.S)
        if m:
            lines = m.group(1).splitlines()
            break
cur, seq = ("?", 0), []
for ln in lines:
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
        seq.append(cur)
print(f"ncu instructions {len(inst)}, disasm instructions {len(seq)}")
agg = {}
n = min(len(inst), len(seq))
for i in range(n):
    agg[seq[i]] = agg.get(seq[i], 0) + inst[i][1]
tot = sum(s for _, s in inst)
src_cache = {}
def src(f, l):
    for root in ("bayes-skopt_b200/csrc", "."):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][l - 1].strip() if l - 1 < len(src_cache[p]) else ""
    return ""
for (f, l), s in sorted(agg.items(), key=lambda kv: -kv[1])[:topn]:
    print(f"{s:7d} {100*s/tot:5.1f}%  {f}:{l:<4d} {src(f,l)[:95]}")
