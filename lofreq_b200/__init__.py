"""lofreq_b200 — B200-native (sm_100a CUDA) implementation of LoFreq's per-pileup-column
SNV test behind the reference's own function surface.  The compute lives in
lib/liblofreq_b200.so (C ABI: include/lofreq_b200.h); this package is the thin host-side
mirror of the reference interface used by the tests and the benchmark."""
from .snpcaller import (Caller, varcall_conf, LDBL_MAX, LDBL_MIN, ST_VALUE, ST_LDBLMAX, ST_LDBLMIN)  # noqa: F401

__all__ = ["Caller", "varcall_conf", "LDBL_MAX", "LDBL_MIN", "ST_VALUE", "ST_LDBLMAX", "ST_LDBLMIN"]
