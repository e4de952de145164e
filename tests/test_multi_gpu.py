"""The library's own collective on real GPUs (needs two or more devices on the box; skipped otherwise): every rank solves its
shard, qmb200_allgather_policy gathers the packed policy over NCCL on the context's communication stream, and the result equals the
single-process solution of the whole batch bit for bit (tools/multi_gpu_check.py, also run by hand under gpurun --gpus 2 / 8:
profiles/bench_r02_*gpu*.json carry the matching bench lines)."""
import os
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


def test_nccl_allgather_equals_single_process_solution():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU on this box")
    world = 2
    port = 29600 + (os.getpid() % 300)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tools", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mismatching problems: 0" in r.stdout
