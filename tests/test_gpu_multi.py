"""GPU, world size 2: the N > 1 path of the training step on real hardware (NCCL over NVLink) against the single-process CPU
oracle.  Skipped on a single-GPU box (the CPU gloo tests of tests/test_distributed_cpu.py cover the host logic there); run
with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_two_rank_train_step_matches_single_process_oracle(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = tmp_path / "multi.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=540)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(out.read_text())
    assert all(x["same_params"] for x in res), res                       # replicas agree bit for bit after the step
    r0 = res[0]
    print(r0)
    assert abs(r0["loss_mean"] - r0["oracle_loss"]) <= 1e-5 * abs(r0["oracle_loss"]), r0
    assert r0["worst_grad_rel"] <= 1e-5, r0                              # averaged gradient == oracle gradient of the global batch
    assert abs(r0["grad_norm"] - r0["oracle_norm"]) <= 1e-5 * r0["oracle_norm"], r0
    assert r0["worst_param_rel"] <= 1e-5, r0                             # parameters after clip + Adam (firm entries)
