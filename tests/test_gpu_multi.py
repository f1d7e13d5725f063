"""Multi-GPU paths on real GPUs (skipped on a single-GPU box): one process per GPU under torchrun, NCCL.

The host-side logic of the same paths (sharded L-BFGS scalars, session sharding) runs on CPU over gloo in
tests/test_distributed_cpu.py; here the device-driven variants and the NCCL collectives are exercised."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs on one box")]


def _torchrun(script, nproc=2, env=None, port=29611):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", script)]
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, cwd=ROOT, env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("device_driven", ["1", "0"])
def test_joint_shared_v_model_on_two_gpus_equals_one_gpu(device_driven):
    """src/model/rrr.py:37-49: ONE model over 4 sessions with a shared V, sessions sharded over 2 ranks ([dV, loss]
    all-reduced, L-BFGS scalars all-gathered) == the same joint model on one GPU; V replicas bit-identical.  Both the
    device-driven optimiser (no host read inside the step) and the host-driven one."""
    out = _torchrun("joint_2gpu_check.py", env={"VS_LBFGS_DEVICE": device_driven}, port=29611 + int(device_driven))
    assert "bit-identical across ranks: True" in out


@pytest.mark.parametrize("mode", ["classic", "exact"])
def test_one_session_trial_sharded_on_two_gpus_equals_one_gpu(mode):
    """the trials of ONE session over 2 ranks (global z-score statistics, gradient all-reduce), in the
    classic 3-plane layout and in the default exact-operand layout."""
    _torchrun("trial_shard_check.py", env={"MODE": mode}, port=29621 + (mode == "exact"))
