#!/usr/bin/env python
"""Counts the Blackwell-only SASS mnemonics per kernel of libtacult_b200.so (evidence that the tensor-core kernels are
tcgen05 / TMEM / TMA code, not mma.sync or library calls):  UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA load,
UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc, STTM = tcgen05.st.   python tools/sass_ops.py > profiles/sass_ops.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
LIB = os.path.join(ROOT, "tacult_b200", "lib", "libtacult_b200.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "UTCATOMSWS", "SYNCS", "REDUX", "HMMA", "DFMA", "DMUL", "MUFU.RSQ64H"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = demangle(m.group(1))
            cur = re.sub(r"\((int|bool)\)", "", cur)
            cur = re.sub(r"\(.*", "", cur).replace("void ", "")
            order.append(cur)
            continue
        if cur is None:
            continue
        for op in OPS:
            if re.search(r"\b" + re.escape(op) + r"\b", line) or (op.endswith("H") and op in line):
                counts[cur][op] += 1
    arch = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout.strip().splitlines()
    print("# library:", os.path.relpath(LIB, ROOT))
    print("# cubins :", ", ".join(a.split()[-1] for a in arch))
    print("# kernel".ljust(44) + "".join(op.rjust(12) for op in OPS))
    for k in order:
        print(k[:43].ljust(44) + "".join(str(counts[k][op]).rjust(12) for op in OPS))
    total = collections.Counter()
    for k in order:
        total.update(counts[k])
    print("TOTAL".ljust(44) + "".join(str(total[op]).rjust(12) for op in OPS))


if __name__ == "__main__":
    sys.exit(main())
