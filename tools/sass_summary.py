"""SASS opcode summary of libbgp.so per kernel (developer tooling; runs without a GPU):
what proves which hardware paths a kernel uses -- DMMA (FP64 tensor), LDGSTS (cp.async), UBLKCP (TMA bulk
copy), SYNCS (mbarrier), cluster barriers (UCGABAR / BAR ... CLUSTER), MUFU, spills (STL/LDL).
usage: python tools/sass_summary.py [lib.so] > profiles/sass_opcodes_rNN.txt"""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "bayes-skopt_b200/libbgp.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
watch = ["DMMA", "DFMA", "DMUL", "DADD", "MUFU", "LDGSTS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UCGABAR", "BAR",
         "LDS", "STS", "LDG", "STG", "LDL", "STL", "SHFL", "ATOMS", "ATOMG", "RED", "CS2R", "ERRBAR", "MEMBAR", "FENCE"]
kern, counts, total = None, {}, {}
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        counts[kern], total[kern] = collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        base = op.split(".")[0]
        if base in watch:
            counts[kern][base] += 1
        if "CLUSTER" in op or base == "UCGABAR":
            counts[kern]["cluster-barrier"] += 1
print(f"SASS opcode summary of {so} (cuobjdump -sass; sm_100a)\n")
allc = collections.Counter()
for k in sorted(counts, key=lambda k: -total[k]):
    allc.update(counts[k])
    body = "  ".join(f"{o}={c}" for o, c in sorted(counts[k].items(), key=lambda kv: -kv[1]))
    print(f"{k[:90]:90s} {total[k]:6d} instr\n    {body}")
print("\nwhole library: " + "  ".join(f"{o}={c}" for o, c in sorted(allc.items(), key=lambda kv: -kv[1])))
