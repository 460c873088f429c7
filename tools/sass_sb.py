"""Decode the scheduling control bits (stall, scoreboards, wait mask) of a kernel's SASS.

usage: python tools/sass_sb.py <object> <function-substring> [pattern]
Prints every instruction matching `pattern` (default LDG|LDS|DMMA) with its write/read scoreboard
and wait mask; used to check that prefetch loads really stay in flight (developer tooling)."""
import re, subprocess, sys
obj, fn = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3] if len(sys.argv) > 3 else r"LDG|LDS|DMMA|LDGSTS|BAR|DEPBAR")
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.split("\n")
on = False
ins = []
i = 0
while i < len(txt):
    l = txt[i]
    if "Function :" in l:
        on = fn in l
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", l)
    if on and m and i + 1 < len(txt):
        m2 = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", txt[i + 1])
        if m2:
            c = (int(m2.group(1), 16) >> 41) & ((1 << 23) - 1)
            ins.append((m.group(1), m.group(2), c & 0xF, (c >> 4) & 1, (c >> 5) & 7, (c >> 8) & 7, (c >> 11) & 0x3F))
            i += 2
            continue
    i += 1
for a, t, st, y, wb, rb, w in ins:
    if pat.search(t) or w:
        print(f"{a} st{st:<2d} wb{wb} rb{rb} wait{w:06b}  {t[:100]}")
