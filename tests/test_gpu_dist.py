"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): launches tests/dist_gpu_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_partitioned_bp_and_gates_match_oracle(nproc):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs (logs of the 2 / 4 / 8 GPU runs: profiles/r2_dist_check_*gpu.log)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(29517 + nproc), os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "DIST_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("nthreads", [2])
def test_single_process_context_group(nthreads):
    """itn_ctx_create_group (ncclCommInitAll): the same checks from ONE process with one host thread per GPU."""
    import torch
    if torch.cuda.device_count() < nthreads:
        pytest.skip(f"needs {nthreads} GPUs (log of the builder's run: profiles/r2_dist_check_threads2.log)")
    cmd = [sys.executable, os.path.join(ROOT, "tests", "dist_gpu_check.py"), "--threads", str(nthreads)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "DIST_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
