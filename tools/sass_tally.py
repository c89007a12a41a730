"""tools/sass_tally.py -- opcode histogram of an address range of a cuobjdump -sass listing.

    cuobjdump -sass -fun <mangled> lib.so > k.sass
    python tools/sass_tally.py k.sass 0x580 0x34e0 [elements_per_iteration]
"""
import collections
import re
import sys

path, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
per = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
pat = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)")
hist = collections.Counter()
for line in open(path):
    m = pat.match(line)
    if not m:
        continue
    addr = int(m.group(1), 16)
    if lo <= addr < hi:
        op = m.group(2)
        mods = m.group(3)
        if op in ("LDG", "STG", "LDS", "STS", "MUFU", "IMAD", "I2F", "F2I", "F2F", "I2FP", "F2FP"):
            op += mods.split(".")[1] and "." + mods.split(".")[1] if mods else ""
        hist[op] += 1
total = sum(hist.values())
for op, c in hist.most_common():
    print(f"{c:6d} {c / per:7.2f}  {op}")
print(f"{total:6d} {total / per:7.2f}  TOTAL")
