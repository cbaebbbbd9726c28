"""The multi-GPU path on real GPUs (skipped on a one-GPU box): BASELINE config C4 through bench.py — the file lives on
rank 0's GPU, members are sharded by shard.partition, slabs travel by NCCL point-to-point, every member's CRC-32 / ISIZE is
checked on the device and every gathered member again on the host after the run."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif("_ngpu() < 2")
def test_c4_two_gpus_scatter_inflate_gather():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--workload", "c4",
                        "--c4-mib", "256", "--steps", "2", "--warmup", "3"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong" and line["value"] > 0
    assert line["config"]["inflated_bytes"] == 256 << 20


def test_c4_one_gpu_same_code_path():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "c4", "--c4-mib", "128", "--steps", "2",
                        "--warmup", "3"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 1 and line["value"] > 0 and line["config"]["inflated_bytes"] == 128 << 20
