"""examples/call_adapter.c — plain C99 host code over the C ABI (include/lofreq_b200.h), the shape of LoFreq's
main_call / call_vars around a fake pileup.  CPU: the header is valid C, the program links against the library and
refuses to run without a GPU.  GPU: it calls the planted variants and reports the test count."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    from lofreq_b200 import build as b
    libdir = os.path.dirname(b.LIB if os.path.exists(b.LIB) else b.build())      # never rebuild a library that is there
    exe = str(tmp_path / "call_adapter")
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "call_adapter.c"), "-L" + libdir, "-llofreq_b200", "-Wl,-rpath," + libdir, "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_c_adapter_builds_and_has_no_cpu_path(tmp_path):
    import torch
    exe = build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    out = subprocess.run([exe, "100", "50"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 1 and "no CPU path" in out.stderr


@pytest.mark.gpu
def test_c_adapter_calls_the_planted_variants(tmp_path):
    exe = build(tmp_path)
    n_cols, depth = 20000, 400
    out = subprocess.run([exe, str(n_cols), str(depth)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    m = re.search(r"Number of substitution tests performed: (\d+)", out.stderr)
    assert m and int(m.group(1)) > 0 and int(m.group(1)) % 3 == 0
    called = {}
    for ln in out.stdout.splitlines():
        f = ln.split("\t")
        assert f[0] == "chr1" and len(f) == 8
        called.setdefault(int(f[1]) - 1, []).append((f[3], f[4], int(f[5])))
    planted = [c for c in range(n_cols) if c % 97 == 5]
    # every planted variant (2-30 % of 400 reads at ~Q30) is called, with the planted alt allele and in position order
    for c in planted:
        ref, alt = "ACGT"[c & 3], "ACGT"[((c & 3) + 1 + c % 3) & 3]
        assert c in called and any(r == ref and a == alt and q > 20 for r, a, q in called[c]), c
    assert list(called) == sorted(called)
    assert len(called) <= len(planted) + 20         # and hardly anything else (sequencing errors stay below the threshold)
