#!/usr/bin/env python
"""Builds a VARIANT of libtacult_b200.so with extra nvcc defines into tacult_b200/lib/variants/<name>/ for A/B runs:
    python tools/build_variant.py noprefilter -DTC_PUCT_PREFILTER=0
    TACULT_B200_LIB=tacult_b200/lib/variants/noprefilter/libtacult_b200.so python bench.py ..."""
import os
import subprocess
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
from tacult_b200 import build as b  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    out = os.path.join(b.LIBDIR, "variants", name)
    os.makedirs(out, exist_ok=True)
    objs, procs = [], []
    for src in b.sources():
        obj = os.path.join(out, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen(["nvcc", *[f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")], *flags, "-c", src, "-o", obj]))
    assert all(p.wait() == 0 for p in procs)
    lib = os.path.join(out, "libtacult_b200.so")
    subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs, "-cudart", "static"])
    for o in objs:
        os.remove(o)
    print(lib)


if __name__ == "__main__":
    main()
