#!/usr/bin/env python
"""A second build of the library with extra -D flags on some sources, for A/B timing on one GPU box:
  python tools/build_variant.py NAME front.cu:-DLFB_FRONT_CTAS=3 [dp_fused.cu:-DX=1 ...]
writes lofreq_b200/lib/var_NAME.so (selected at run time with LFB200_LIB=...)."""
import os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from lofreq_b200 import build as B

name = sys.argv[1]
over = {}
for a in sys.argv[2:]:
    src, flag = a.split(":", 1)
    over.setdefault(src, []).append(flag)
B.build()
objs = []
for s in B.SOURCES:
    if s in over:
        o = os.path.join(B.OBJDIR, "var_%s_%s.o" % (name, os.path.splitext(s)[0]))
        subprocess.check_call([B._nvcc()] + B.NVCC_FLAGS + over[s] + ["-c", os.path.join(B.CSRC, s), "-o", o])
        objs.append(o)
    else:
        objs.append(B._obj(s))
out = os.path.join(B.LIBDIR, "var_%s.so" % name)
subprocess.check_call([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-ldl", "-lrt"])
print(out)
