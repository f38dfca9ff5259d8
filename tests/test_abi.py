"""The C ABI without a GPU: the library builds, loads, exports every symbol include/lofreq_b200.h
declares, the ctypes mirrors have the C struct layouts, and compute entry points fail loudly instead of
falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lofreq_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lfb200_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from lofreq_b200 import capi
    lib = capi.load()
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(capi.SYMBOLS) == syms


def test_struct_layouts_match_the_header(tmp_path):
    """compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirrors"""
    from lofreq_b200 import capi
    prog = tmp_path / "layout.c"
    prog.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "lofreq_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\\n", sizeof(lfb200_conf_t), sizeof(lfb200_batch_t), sizeof(lfb200_site_t),
         sizeof(lfb200_dense_out_t), sizeof(lfb200_summary_t));
  printf("%zu %zu %zu %zu %zu %zu\\n", offsetof(lfb200_site_t, lnp), offsetof(lfb200_site_t, pvalue),
         offsetof(lfb200_site_t, alt_count), offsetof(lfb200_site_t, qual), offsetof(lfb200_site_t, status),
         offsetof(lfb200_site_t, called));
  printf("%zu %zu\\n", offsetof(lfb200_conf_t, sig), offsetof(lfb200_conf_t, bonf_subst));
  return 0; }
''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(prog)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    got = [int(x) for x in out]
    S = capi.Site
    want = [C.sizeof(capi.Conf), C.sizeof(capi.Batch), C.sizeof(S), C.sizeof(capi.DenseOut), C.sizeof(capi.Summary),
            S.lnp.offset, S.pvalue.offset, S.alt_count.offset, S.qual.offset, S.status.offset, S.called.offset,
            capi.Conf.sig.offset, capi.Conf.bonf_subst.offset]
    assert got == want
    assert capi._site_dtype().itemsize == C.sizeof(S)


def test_conf_defaults_are_init_varcall_conf():
    import lofreq_b200
    c = lofreq_b200.varcall_conf()
    # snpcaller.c:626-651 / defaults.h
    assert (c.min_bq, c.min_alt_bq, c.def_alt_bq, c.min_jq, c.min_alt_jq, c.def_alt_jq) == (6, 6, 0, 0, 0, 0)
    assert (c.min_cov, c.bonf_dynamic, c.bonf_subst, c.num_snv_tests) == (1, 1, 1, 0)
    assert c.flag == 1 | 2 | 8 and abs(c.sig - 0.01) < 1e-9
    from oracle.pyoracle import default_conf
    d = default_conf()
    for k, v in d.items():
        assert abs(getattr(c, k) - v) < 1e-9, k


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import lofreq_b200
    from lofreq_b200.capi import Lfb200Error
    with pytest.raises(Lfb200Error, match="no CUDA device"):
        lofreq_b200.Caller(0)
    lib = lofreq_b200.capi.load()
    pv = np.zeros(3, np.longdouble)
    ep = np.full(10, 0.001)
    cn = np.array([1, 0, 0], np.int32)
    rc = lib.lfb200_snpcaller(pv.ctypes.data_as(C.c_void_p), ep.ctypes.data_as(C.c_void_p), 10,
                              cn.ctypes.data_as(C.c_void_p), 1, 1.0, -1)
    assert rc != 0          # the link-compatible snpcaller() refuses too
