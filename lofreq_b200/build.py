"""Builds lofreq_b200/lib/liblofreq_b200.so in-tree with nvcc for sm_100a.
nvcc cross-compiles without a GPU; the built library travels with the repo.
Every source is compiled to its own object (in parallel, only when stale), then linked."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "liblofreq_b200.so")
SOURCES = ["snv_kernels.cu", "front.cu", "dp_fused.cu", "xl.cu", "poissbin.cu", "mailbox.cu", "binom.cu", "fisher.cu", "baq.cu", "synth.cu", "host_api.cpp", "shard_comm.cpp"]
HEADERS = ["internal.h", "dev_common.cuh", "screen_common.cuh", "baq_core.cuh", "synth_tables.h", os.path.join("..", "..", "include", "lofreq_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _obj(src):
    return os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    return any(_stale(LIB, [os.path.join(CSRC, s)] + hdrs) for s in SOURCES)


def _compile(src, verbose):
    cmd = [_nvcc()] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", _obj(src)]
    if verbose:
        print(" ".join(cmd))
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, out.stdout, out.stderr))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    todo = [s for s in SOURCES if force or _stale(_obj(s), [os.path.join(CSRC, s)] + hdrs)]
    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        list(ex.map(lambda s: _compile(s, verbose), todo))
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [_obj(s) for s in SOURCES] + ["-ldl", "-lrt"]
    if verbose:
        print(" ".join(cmd))
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("link failed:\n" + out.stdout + out.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
