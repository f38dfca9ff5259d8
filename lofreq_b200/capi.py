"""ctypes binding of include/lofreq_b200.h — the same C ABI a LoFreq maintainer
would bind from C.  Loading fails loudly when the library has not been built;
creating a context fails loudly when there is no CUDA device (no CPU path)."""
import ctypes as C
import os

from . import build as _build

NUM_NONCONS = 3
USE_BAQ, USE_MQ, USE_SQ, USE_IDAQ = 1, 2, 4, 8
ST_VALUE, ST_LDBLMAX, ST_LDBLMIN = 0, 1, 2


class Conf(C.Structure):
    _fields_ = [("min_bq", C.c_int), ("min_alt_bq", C.c_int), ("def_alt_bq", C.c_int),
                ("min_jq", C.c_int), ("min_alt_jq", C.c_int), ("def_alt_jq", C.c_int),
                ("min_cov", C.c_int), ("bonf_dynamic", C.c_int), ("flag", C.c_int),
                ("sig", C.c_float), ("bonf_subst", C.c_longlong), ("num_snv_tests", C.c_longlong)]


class Batch(C.Structure):
    _fields_ = [("n_cols", C.c_longlong), ("col_off", C.c_void_p), ("nt_cnt", C.c_void_p),
                ("ref_base", C.c_void_p), ("coverage", C.c_void_p),
                ("bq", C.c_void_p), ("mq", C.c_void_p), ("baq", C.c_void_p), ("sq", C.c_void_p),
                ("num_bases", C.c_void_p)]


class Site(C.Structure):
    _fields_ = [("col", C.c_longlong), ("bonf", C.c_longlong), ("lnp", C.c_double * 3), ("ln_floor", C.c_double),
                ("pvalue", C.c_longdouble * 3), ("alt_count", C.c_int * 3), ("alt_raw_count", C.c_int * 3),
                ("qual", C.c_int * 3), ("status", C.c_ubyte * 3), ("called", C.c_ubyte * 3),
                ("flags", C.c_ubyte), ("reserved", C.c_ubyte)]


def _site_dtype():
    import numpy as np
    names, formats, offsets = [], [], []
    for name, np_fmt in (("col", "<i8"), ("bonf", "<i8"), ("lnp", ("<f8", (3,))), ("pvalue", (np.longdouble, (3,))),
                         ("alt_count", ("<i4", (3,))), ("alt_raw_count", ("<i4", (3,))), ("qual", ("<i4", (3,))),
                         ("status", ("u1", (3,))), ("called", ("u1", (3,))), ("flags", "u1"), ("ln_floor", "<f8")):
        names.append(name)
        formats.append(np_fmt)
        offsets.append(getattr(Site, name).offset)
    return np.dtype(dict(names=names, formats=formats, offsets=offsets, itemsize=C.sizeof(Site)))


def sites_to_numpy(sites, n):
    """structured numpy copy of the first n entries of a (Site * m) ctypes array"""
    import numpy as np
    dt = _site_dtype()
    if n <= 0:
        return np.zeros(0, dtype=dt)
    return np.frombuffer(sites, dtype=dt, count=n).copy()


class DenseOut(C.Structure):
    _fields_ = [("alt_counts", C.c_void_p), ("alt_raw_counts", C.c_void_p), ("tested", C.c_void_p),
                ("bonf_used", C.c_void_p), ("lnp", C.c_void_p), ("status", C.c_void_p),
                ("pvalues", C.c_void_p), ("called", C.c_void_p), ("qual", C.c_void_p)]


class Summary(C.Structure):
    _fields_ = [("n_cols", C.c_longlong), ("n_tested", C.c_longlong), ("n_sites", C.c_longlong),
                ("n_heavy", C.c_longlong), ("bonf_subst_final", C.c_longlong), ("num_snv_tests", C.c_longlong),
                ("n_unsupported", C.c_longlong)]


class PlpCol(C.Structure):
    """lfb200_plp_col_t: what plp_to_errprobs reads of a plp_col_t (plp.h:73-145)"""
    _fields_ = [("ref_base", C.c_char), ("coverage_plp", C.c_int), ("n", C.c_int * 4),
                ("base_quals", C.c_void_p * 4), ("map_quals", C.c_void_p * 4), ("baq_quals", C.c_void_p * 4),
                ("source_quals", C.c_void_p * 4)]


SITE_FN = C.CFUNCTYPE(None, C.POINTER(Site), C.c_longlong, C.c_char, C.c_int, C.c_void_p)


class Dp4(C.Structure):
    _fields_ = [("ref_fw", C.c_int), ("ref_rv", C.c_int), ("alt_fw", C.c_int), ("alt_rv", C.c_int)]


class Variant(C.Structure):
    _fields_ = [("tag", C.c_longlong), ("bonf", C.c_longlong), ("lnp", C.c_double), ("af", C.c_float), ("qual", C.c_int),
                ("dp", C.c_int), ("sb", C.c_int), ("hqa", C.c_int), ("dp4", Dp4), ("ref_base", C.c_char), ("alt_base", C.c_char)]


VARIANT_FN = C.CFUNCTYPE(None, C.POINTER(Variant), C.c_void_p)

# every symbol include/lofreq_b200.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = ["lfb200_create", "lfb200_destroy", "lfb200_last_error", "lfb200_init_conf", "lfb200_call_columns", "lfb200_set_host_planes",
           "lfb200_screen_device", "lfb200_ntested_device", "lfb200_test_device", "lfb200_ntested_copy_device", "lfb200_test_device_from", "lfb200_bonf_start_device", "lfb200_comm_unique_id", "lfb200_comm_init", "lfb200_comm_exchange", "lfb200_comm_gathered",
           "lfb200_sites_device", "lfb200_sites_view", "lfb200_sites_buffer", "lfb200_sites_begin", "lfb200_sites_end", "lfb200_set_site_pvalues", "lfb200_site_fill_pvalues",
           "lfb200_device_results", "lfb200_set_profiling", "lfb200_get_profile", "lfb200_dfma_peak", "lfb200_last_job_counts", "lfb200_graph_replays", "lfb200_kpa_glocal_batch", "lfb200_copy_counts_device", "lfb200_builder_create", "lfb200_builder_add_column", "lfb200_builder_flush", "lfb200_builder_destroy", "lfb200_builder_pending", "lfb200_sb_qual_batch", "lfb200_format_snv_info", "lfb200_format_indel_info", "lfb200_format_snv_record", "lfb200_builder_add_column_strands", "lfb200_builder_on_variant",
           "lfb200_snpcaller", "lfb200_snpcaller_batch", "lfb200_poissbin", "lfb200_poissbin_batch", "lfb200_batch_errprobs", "lfb200_plp_to_errprobs", "lfb200_indel_tests", "lfb200_binom", "lfb200_binom_batch", "lfb200_synth_depths",
           "lfb200_synth_columns"]

_lib = None


def lib_path():
    return _build.LIB


def load():
    """dlopen liblofreq_b200.so and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_build.LIB):
        raise RuntimeError("liblofreq_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`"
                           " (there is no CPU fallback)")
    # LFB200_LIB: another build of the same library (A/B timing of two builds on one GPU box)
    lib = C.CDLL(os.environ.get("LFB200_LIB") or _build.LIB)
    vp, ll = C.c_void_p, C.c_longlong
    lib.lfb200_create.restype = C.c_int
    lib.lfb200_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.lfb200_destroy.restype = None
    lib.lfb200_destroy.argtypes = [vp]
    lib.lfb200_last_error.restype = C.c_char_p
    lib.lfb200_last_error.argtypes = []
    lib.lfb200_init_conf.restype = None
    lib.lfb200_init_conf.argtypes = [C.POINTER(Conf)]
    lib.lfb200_call_columns.restype = C.c_int
    lib.lfb200_call_columns.argtypes = [vp, C.POINTER(Conf), C.POINTER(Batch), C.POINTER(DenseOut),
                                        C.POINTER(Site), ll, C.POINTER(Summary)]
    lib.lfb200_screen_device.restype = C.c_int
    lib.lfb200_screen_device.argtypes = [vp, C.POINTER(Conf), C.POINTER(Batch), vp]
    lib.lfb200_ntested_device.restype = C.c_int
    lib.lfb200_ntested_device.argtypes = [vp, vp, C.POINTER(ll)]
    lib.lfb200_test_device.restype = C.c_int
    lib.lfb200_test_device.argtypes = [vp, C.POINTER(Conf), vp]
    lib.lfb200_ntested_copy_device.restype = C.c_int
    lib.lfb200_ntested_copy_device.argtypes = [vp, vp, vp]
    lib.lfb200_test_device_from.restype = C.c_int
    lib.lfb200_test_device_from.argtypes = [vp, C.POINTER(Conf), vp, vp]
    lib.lfb200_bonf_start_device.restype = C.c_int
    lib.lfb200_bonf_start_device.argtypes = [vp, vp, C.c_int, ll, vp]
    lib.lfb200_sites_device.restype = C.c_int
    lib.lfb200_sites_device.argtypes = [vp, C.POINTER(Conf), vp, C.POINTER(Site), ll, C.POINTER(Summary)]
    lib.lfb200_sites_view.restype = C.c_int
    lib.lfb200_sites_view.argtypes = [vp, C.POINTER(Conf), vp, C.POINTER(C.POINTER(Site)), C.POINTER(Summary)]
    lib.lfb200_sites_buffer.restype = C.c_int
    lib.lfb200_sites_buffer.argtypes = [vp, C.POINTER(C.POINTER(Site))]
    lib.lfb200_set_site_pvalues.restype = C.c_int
    lib.lfb200_set_site_pvalues.argtypes = [vp, C.c_int]
    lib.lfb200_site_fill_pvalues.restype = None
    lib.lfb200_site_fill_pvalues.argtypes = [C.POINTER(Site), ll]
    lib.lfb200_comm_unique_id.restype = C.c_int
    lib.lfb200_comm_unique_id.argtypes = [vp]
    lib.lfb200_comm_init.restype = C.c_int
    lib.lfb200_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.lfb200_comm_exchange.restype = C.c_int
    lib.lfb200_comm_exchange.argtypes = [vp, vp, ll, ll, C.POINTER(vp)]
    lib.lfb200_comm_gathered.restype = C.c_int
    lib.lfb200_comm_gathered.argtypes = [vp, vp, vp, vp]
    lib.lfb200_set_host_planes.restype = C.c_int
    lib.lfb200_set_host_planes.argtypes = [vp, C.c_int]
    lib.lfb200_binom.restype = C.c_int
    lib.lfb200_binom.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double]
    lib.lfb200_binom_batch.restype = C.c_int
    lib.lfb200_binom_batch.argtypes = [vp, ll, vp, vp, vp, vp, vp, vp]
    lib.lfb200_sites_begin.restype = C.c_int
    lib.lfb200_sites_begin.argtypes = [vp, C.POINTER(Conf), vp, C.POINTER(Site), ll]
    lib.lfb200_sites_end.restype = C.c_int
    lib.lfb200_sites_end.argtypes = [vp, C.POINTER(Summary)]
    lib.lfb200_device_results.restype = C.c_int
    lib.lfb200_device_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.lfb200_set_profiling.restype = C.c_int
    lib.lfb200_set_profiling.argtypes = [vp, C.c_int]
    lib.lfb200_get_profile.restype = C.c_int
    lib.lfb200_get_profile.argtypes = [vp, C.POINTER(C.c_float)]
    lib.lfb200_builder_create.restype = C.c_int
    lib.lfb200_builder_create.argtypes = [C.POINTER(vp), vp, C.POINTER(Conf), ll, SITE_FN, vp]
    lib.lfb200_builder_add_column.restype = C.c_int
    lib.lfb200_builder_add_column.argtypes = [vp, ll, C.c_char, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    lib.lfb200_builder_flush.restype = C.c_int
    lib.lfb200_builder_flush.argtypes = [vp]
    lib.lfb200_builder_pending.restype = ll
    lib.lfb200_builder_pending.argtypes = [vp]
    lib.lfb200_builder_add_column_strands.restype = C.c_int
    lib.lfb200_builder_add_column_strands.argtypes = [vp, ll, C.c_char, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    lib.lfb200_builder_on_variant.restype = C.c_int
    lib.lfb200_builder_on_variant.argtypes = [vp, VARIANT_FN, vp]
    lib.lfb200_sb_qual_batch.restype = C.c_int
    lib.lfb200_sb_qual_batch.argtypes = [vp, ll, vp, vp]
    lib.lfb200_format_snv_info.restype = C.c_int
    lib.lfb200_format_snv_info.argtypes = [C.c_char_p, C.c_ulong, C.c_int, C.c_float, C.c_int, C.POINTER(Dp4), C.c_int]
    lib.lfb200_format_indel_info.restype = C.c_int
    lib.lfb200_format_indel_info.argtypes = [C.c_char_p, C.c_ulong, C.c_int, C.c_float, C.c_int, C.POINTER(Dp4), C.c_int]
    lib.lfb200_format_snv_record.restype = C.c_int
    lib.lfb200_format_snv_record.argtypes = [C.c_char_p, C.c_ulong, C.c_char_p, C.c_long, C.c_char, C.c_char, C.c_int, C.c_char_p]
    lib.lfb200_builder_destroy.restype = None
    lib.lfb200_builder_destroy.argtypes = [vp]
    lib.lfb200_copy_counts_device.restype = C.c_int
    lib.lfb200_copy_counts_device.argtypes = [vp, vp, vp]
    lib.lfb200_kpa_glocal_batch.restype = C.c_int
    lib.lfb200_kpa_glocal_batch.argtypes = [vp, C.c_longlong, vp, vp, vp, vp, vp, C.c_float, C.c_float, C.c_int, vp, vp]
    lib.lfb200_graph_replays.restype = C.c_longlong
    lib.lfb200_graph_replays.argtypes = [vp]
    lib.lfb200_last_job_counts.restype = C.c_int
    lib.lfb200_last_job_counts.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.lfb200_dfma_peak.restype = C.c_double
    lib.lfb200_dfma_peak.argtypes = [vp, vp]
    lib.lfb200_snpcaller.restype = C.c_int
    lib.lfb200_snpcaller.argtypes = [vp, vp, C.c_int, vp, ll, C.c_double, C.c_int]
    lib.lfb200_snpcaller_batch.restype = C.c_int
    lib.lfb200_snpcaller_batch.argtypes = [vp, ll, vp, vp, vp, vp, C.c_double, vp, vp, vp]
    lib.lfb200_poissbin.restype = vp
    lib.lfb200_poissbin.argtypes = [vp, vp, C.c_int, C.c_int, ll, C.c_double]
    lib.lfb200_poissbin_batch.restype = C.c_int
    lib.lfb200_poissbin_batch.argtypes = [vp, ll, vp, vp, vp, vp, C.c_double, vp, vp, vp, vp]
    lib.lfb200_batch_errprobs.restype = C.c_int
    lib.lfb200_batch_errprobs.argtypes = [vp, C.POINTER(Conf), C.POINTER(Batch), vp, vp, vp, vp, vp]
    lib.lfb200_plp_to_errprobs.restype = None
    lib.lfb200_plp_to_errprobs.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), vp, vp, vp, C.POINTER(PlpCol), C.POINTER(Conf)]
    lib.lfb200_indel_tests.restype = C.c_int
    lib.lfb200_indel_tests.argtypes = [vp, C.POINTER(Conf), ll, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.lfb200_synth_depths.restype = C.c_int
    lib.lfb200_synth_depths.argtypes = [C.c_int, ll, ll, vp, vp]
    lib.lfb200_synth_columns.restype = C.c_int
    lib.lfb200_synth_columns.argtypes = [C.c_int, ll, ll, vp, vp, vp, vp, vp, vp, vp]
    _lib = lib
    return lib


class Lfb200Error(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise Lfb200Error(load().lfb200_last_error().decode() or "lofreq_b200 call failed (rc=%d)" % rc)
