"""The callback boundary and the reporting tail (8f #4), at the level the reference's own users see:
VCF text.  Identical plp_col_t objects (the fake pileup of oracle/call_harness.c) go

  (a) through the reference's REAL call_vars() -> call_snvs() -> report_var() -> vcf_write_var()
      (lofreq_call.c, vcf.c, fet.c, snpcaller.c compiled unmodified: oracle/_ref/libcallref.so, committed as
      tests/golden/vcf_boundary.npz by tests/golden/make_golden_vcf.py), and
  (b) through the product's drop-in callback  void lfb200_call_vars(const plp_col_t *, void *)  + lfb200_flush()
      (lofreq_b200/adapter/lofreq_adapter.c compiled against the reference's own plp.h / snpcaller.h / vcf.h),

and the VCF records (CHROM POS REF ALT QUAL, INFO DP / AF / SB / DP4 / HQA), conf->bonf_subst and num_snv_tests must be
byte-identical."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import synth_np
from oracle.pyoracle import CallOracle, default_conf, have_call_oracle
from test_oracle import GOLD

Z = np.load(os.path.join(GOLD, "vcf_boundary.npz"))


def _golden_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_vcf", os.path.join(GOLD, "make_golden_vcf.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.CASES


CASES = _golden_cases()


def test_golden_is_what_the_reference_writes_here():
    """where the callback oracle is built (this container; it also travels to the GPU box) the committed text is reproduced"""
    if not have_call_oracle():
        pytest.skip("oracle/_ref/libcallref.so not built (reference tree absent)")
    o = CallOracle()
    name, wl, c0, n, baq, over = CASES[0]
    b = synth_np.generate(wl, c0, n, with_baq=baq, with_strand=True)
    txt, bonf, tests = o.call_vars_vcf(b, b["strand8"], default_conf(**over), pos=np.arange(c0, c0 + n) % 100000)
    assert txt.encode() == Z["vcf_" + name].tobytes()
    assert [bonf, tests] == Z["counters_" + name].tolist()


def test_record_and_info_formatting_match_vcf_c():
    """lfb200_format_snv_info / lfb200_format_snv_record (host, no GPU) reproduce every golden record byte for byte from
    its parsed fields: vcf_var_sprintf_info (vcf.c:608-629) and vcf_write_var (vcf.c:469-495)"""
    from lofreq_b200 import capi
    lib = capi.load()
    n = 0
    for name, *_ in CASES:
        for line in Z["vcf_" + name].tobytes().decode().splitlines():
            chrom, pos, _id, ref, alt, qual, _flt, info = line.split("\t")
            kv = dict(x.split("=") for x in info.split(";"))
            dp4 = capi.Dp4(*[int(x) for x in kv["DP4"].split(",")])
            buf = C.create_string_buffer(512)
            k = lib.lfb200_format_snv_info(buf, 512, int(kv["DP"]), float(kv["AF"]), int(kv["SB"]), C.byref(dp4), int(kv["HQA"]))
            assert k > 0 and buf.value.decode() == info
            rec = C.create_string_buffer(1024)
            k = lib.lfb200_format_snv_record(rec, 1024, chrom.encode(), int(pos) - 1, ref.encode(), alt.encode(), int(qual), buf.value)
            assert k > 0 and rec.value.decode() == line + "\n"
            n += 1
    assert n > 100
    small = C.create_string_buffer(8)
    assert lib.lfb200_format_snv_info(small, 8, 1, 0.5, 0, C.byref(capi.Dp4(1, 1, 1, 1)), 1) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_adapter_writes_the_reference_vcf(case):
    if not have_call_oracle(adapter=True):
        pytest.skip("oracle/_ref/libcallb200.so not built (needs the reference headers)")
    name, wl, c0, n, baq, over = case
    b = synth_np.generate(wl, c0, n, with_baq=baq, with_strand=True)
    os.environ["LFB200_BATCH_COLS"] = "1000"        # several flushes inside the run + the final lfb200_flush()
    try:
        got, bonf, tests = CallOracle(adapter=True).call_vars_vcf(b, b["strand8"], default_conf(**over),
                                                                   pos=np.arange(c0, c0 + n) % 100000)
    finally:
        os.environ.pop("LFB200_BATCH_COLS", None)
    want = Z["vcf_" + name].tobytes().decode()
    assert got.splitlines() == want.splitlines()
    assert got == want
    assert [bonf, tests] == Z["counters_" + name].tolist()


@pytest.mark.gpu
def test_adapter_gates_and_live_reference():
    """columns call_vars() itself refuses (ref N, consensus indel, num_bases * 2 < coverage: lofreq_call.c:892, 928-932),
    several targets, against the live reference"""
    if not (have_call_oracle(adapter=True) and have_call_oracle()):
        pytest.skip("callback oracle not built")
    b = synth_np.generate("C4", 31000, 2500, with_baq=True, with_strand=True)
    n = 2500
    rng = np.random.default_rng(4)
    ref = b["ref_base"].copy()
    ref[rng.integers(0, n, 40)] = ord("N")
    cons0 = ref.copy()
    cons0[rng.integers(0, n, 60)] = ord("+")
    cons0[rng.integers(0, n, 60)] = ord("-")
    cov = b["nt_cnt"].sum(axis=1).astype(np.int32)
    cov[rng.integers(0, n, 50)] *= 3                 # num_bases * 2 < coverage_plp
    bb = dict(b, ref_base=ref, coverage=cov)
    want = CallOracle().call_vars_vcf(bb, b["strand8"], default_conf(), cons0=cons0, target="chrA")
    got = CallOracle(adapter=True).call_vars_vcf(bb, b["strand8"], default_conf(), cons0=cons0, target="chrA")
    assert got == want and len(want[0].splitlines()) > 5


@pytest.mark.gpu
@pytest.mark.parametrize("flag_over", [{}, dict(flag=2), dict(bonf_dynamic=0)])
def test_adapter_call_indels_bookkeeping(flag_over):
    """--call-indels through the adapter: the bookkeeping of call_indels() (lofreq_call.c:618-726: min_cov
    gate, poly-AT filter, one test per event in hash order, running bonf_indel, num_indel_tests), the tests batched
    through lfb200_indel_tests, INDEL records interleaved with the substitution records in the reference's order —
    VCF text, bonf_indel and num_indel_tests identical to the reference's call_vars()"""
    if not (have_call_oracle(adapter=True) and have_call_oracle()):
        pytest.skip("callback oracle not built")
    from oracle.pyoracle import synth_indels
    n = 3000
    b = synth_np.generate("C4", 88000, n, with_baq=True, with_strand=True)
    ind = synth_indels(n, b["nt_cnt"].sum(axis=1), seed=11, frac=0.08)
    conf = default_conf(**flag_over)
    want = CallOracle().call_vars_vcf(b, b["strand8"], dict(conf), indels=ind, target="chr7")
    os.environ["LFB200_BATCH_COLS"] = "700"
    try:
        got = CallOracle(adapter=True).call_vars_vcf(b, b["strand8"], dict(conf), indels=ind, target="chr7")
    finally:
        os.environ.pop("LFB200_BATCH_COLS", None)
    assert got[0].splitlines() == want[0].splitlines()
    assert got[1:] == want[1:]
    n_indel = sum("INDEL" in ln for ln in want[0].splitlines())
    assert n_indel > 50 and want[4] > 150 and want[3] == (want[4] + 1 if conf["bonf_dynamic"] else 1)


@pytest.mark.gpu
def test_link_level_swap_of_snpcaller_symbols():
    """lofreq_b200/adapter/snpcaller_shim.c: the reference's UNMODIFIED call_vars() / call_snvs() (lofreq_call.c) linked
    against the un-prefixed snpcaller() / plp_to_errprobs() / poissbin() of the shim (every column two GPU round trips) writes
    the VCF text of the reference"""
    if not have_call_oracle(swap=True):
        pytest.skip("oracle/_ref/libcallswap.so not built (needs the reference tree)")
    name, wl, c0, n, baq, over = CASES[0]
    n = 1500
    b = synth_np.generate(wl, c0, n, with_baq=baq, with_strand=True)
    got, bonf, tests = CallOracle(swap=True).call_vars_vcf(b, b["strand8"], default_conf(**over), pos=np.arange(c0, c0 + n) % 100000,
                                                           adapter=False)
    want = [ln for ln in Z["vcf_" + name].tobytes().decode().splitlines() if int(ln.split("\t")[1]) <= n]
    assert got.splitlines() == want and len(want) >= 10
    assert tests == 3 * int(bonf // 3) and bonf > 1000


@pytest.mark.gpu
def test_sb_qual_kernel_matches_report_var():
    """lfb200_sb_qual_batch (k_sb_qual: Fisher's exact test per DP4 table on the device) == PROB_TO_PHREDQUAL_SAFE of
    kt_fisher_exact's two-tailed p as report_var() computes it (lofreq_call.c:108-125, fet.c:62-101), INT_MAX case included"""
    import lofreq_b200
    from lofreq_b200 import capi
    c = lofreq_b200.Caller(0)
    t = np.ascontiguousarray(Z["sb_tables"], np.int32)
    sb = np.zeros(len(t), np.int32)
    capi.check(c.lib.lfb200_sb_qual_batch(c._ctx, len(t), t.ctypes.data_as(C.c_void_p), sb.ctypes.data_as(C.c_void_p)))
    bad = np.flatnonzero(sb.astype(np.int64) != Z["sb_qual"])
    assert len(bad) == 0, [(t[i].tolist(), int(sb[i]), int(Z["sb_qual"][i])) for i in bad[:5]]
    c.close()
