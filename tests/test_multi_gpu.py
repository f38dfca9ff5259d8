"""The multi-GPU data plane on hardware: region shards on 2 GPUs of one box, one process per GPU under
torch.distributed.run, through screen -> lfb200_comm_exchange (k_mail_exchange, mailbox.cu) -> test_device_from -> sites
for 72 batches (past MAIL_DEPTH = 64: slot reuse and acknowledgements), every site compared with the single-GPU run of
the same columns (tests/mgpu_worker.py).  Skipped when fewer than 2 GPUs are visible."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("exchange", ["mailbox", "nccl"])
def test_two_ranks_equal_single_gpu(exchange):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    env.pop("LFB200_EXCHANGE_NCCL", None)
    if exchange == "nccl":
        env["LFB200_EXCHANGE_NCCL"] = "1"
        env["MGPU_BATCHES"] = "8"
    port = 29500 + os.getpid() % 400
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "multi-GPU parity ok" in r.stdout
    print(r.stdout.strip().splitlines()[-1])
