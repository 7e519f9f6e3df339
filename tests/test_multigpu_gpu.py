"""GPU, two or more devices: the multi-GPU exchange inside the library (hc_comm_init, hc_fb_reduce, the gather of a tile-partitioned
ray-casting pass) through scripts/gpu_comm_check.py under torch.distributed.run, one process per GPU.  Skipped on a single-GPU box
(the host logic of the partition is covered on CPU over gloo by tests/test_multigpu_cpu.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_framebuffer_and_raycast_exchange_on_two_gpus(built):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ)
    env.pop("HC_LIB", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29577",
                        os.path.join(ROOT, "scripts", "gpu_comm_check.py")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["ok"] and out["raycast_split_equals_single_gpu"] and out["tiles_equal_after_3_passes_and_repeated_reduce"] and out["qmc_repeat_equal"]
