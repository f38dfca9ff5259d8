"""Synthetic pileup columns generated on the GPU; torch is used only to own the
device memory.  Bit-identical to oracle/synth_np.py (checked in tests/test_parity_gpu.py)."""
import ctypes as C

from . import capi

WORKLOAD_ID = {"C2": 2, "C3": 3, "C4": 4, "C5": 5}


def generate_device(workload, c0, n_cols, with_baq=False, device="cuda:0", pad=16):
    import torch
    lib = capi.load()
    wid = WORKLOAD_ID[workload]
    dev = torch.device(device)
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        depth = torch.empty(n_cols, dtype=torch.int32, device=dev)
        capi.check(lib.lfb200_synth_depths(wid, int(c0), int(n_cols), C.c_void_p(depth.data_ptr()), st))
        pitch = (depth.to(torch.int64) + (pad - 1)) // pad * pad
        col_off = torch.zeros(n_cols + 1, dtype=torch.int64, device=dev)
        torch.cumsum(pitch, 0, out=col_off[1:])
        total = int(col_off[-1].item())
        t = dict(col_off=col_off, depths=depth,
                 nt_cnt=torch.empty((n_cols, 4), dtype=torch.int32, device=dev),
                 ref_base=torch.empty(n_cols, dtype=torch.uint8, device=dev),
                 bq=torch.empty(total + 32, dtype=torch.uint8, device=dev),
                 mq=torch.empty(total + 32, dtype=torch.uint8, device=dev),
                 baq=torch.empty(total + 32, dtype=torch.uint8, device=dev) if with_baq else None,
                 sq=None, coverage=None, total_bytes=total)
        capi.check(lib.lfb200_synth_columns(
            wid, int(c0), int(n_cols), C.c_void_p(col_off.data_ptr()), C.c_void_p(t["nt_cnt"].data_ptr()),
            C.c_void_p(t["ref_base"].data_ptr()), C.c_void_p(t["bq"].data_ptr()), C.c_void_p(t["mq"].data_ptr()),
            C.c_void_p(t["baq"].data_ptr()) if with_baq else None, st))
    return t
