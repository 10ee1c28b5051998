"""Multi-GPU parity of the data-parallel path on the CUDA kernels (tools/dp_selftest.py): needs >= 2 GPUs on the box, so the
driver's single-GPU `-m gpu` run skips it; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu` runs it
(outcome recorded in profiles/r02_dp_selftest_2gpu.json)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_all_reduced_gradients_equal_mean_of_shard_gradients_and_replicas_stay_identical():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "dp_selftest.py"), "--rays", "512"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-4000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["dp_selftest"] == "ok" and d["world_size"] == 2
    assert d["grad_rel_err_vs_mean_of_shards"] < 1e-5 and d["worst_per_parameter_rel_err"] < 1e-4
