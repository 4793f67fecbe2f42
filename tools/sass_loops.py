#!/usr/bin/env python3
"""List the loops (backward branches) of one kernel in `cuobjdump -sass` output with their instruction mix.
usage: cuobjdump -sass -fun <mangled> file.o | python tools/sass_loops.py [min_len]"""
import re
import sys
from collections import Counter

ins = []
for line in sys.stdin:
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
minlen = int(sys.argv[1]) if len(sys.argv) > 1 else 50
addr2i = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
    if not m:
        continue
    tgt = int(m.group(1), 16)
    if tgt < a and tgt in addr2i:
        j = addr2i[tgt]
        body = ins[j:i + 1]
        if len(body) < minlen:
            continue
        ops = Counter()
        for _, x in body:
            x = re.sub(r"^@!?U?P\d+\s+", "", x)
            ops[x.split()[0].split(".")[0]] += 1
        print(f"loop {tgt:#x}..{a:#x}: {len(body)} instr")
        print("   " + ", ".join(f"{k} {v}" for k, v in ops.most_common()))
