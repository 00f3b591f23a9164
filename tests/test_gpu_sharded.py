"""Row-sharded potential solve + presence all-reduce over the library's NCCL communicator on 2 GPUs
(skipped on a single-GPU box; CPU coverage of the same solver code: tests/test_solver_sharded_cpu.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("halo", ["", "peer"])
def test_sharded_solve_two_gpus(halo):
    """Both halo transports of the NCCL communicator: grouped ncclSend/ncclRecv (default) and peer-memory stores over
    NVLink (SSRS_COMM_HALO=peer; falls back to NCCL when the ranks cannot map each other's staging blocks)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29500 + (os.getpid() + (13 if halo else 0)) % 400
    env = dict(os.environ)
    env.pop("SSRS_COMM_HALO", None)
    if halo:
        env["SSRS_COMM_HALO"] = halo
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "sharded_solve_run.py"), "1500", "1800", "30", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["halo_mode"] == "nccl" if not halo else out["halo_mode"] in ("peer", "nccl")
    assert out["sharded"]["converged"] in (1, 2)
    assert out["max_abs_diff_vs_single"] <= 2 * float(np.spacing(np.float32(1000.0)))     # float32-rounding level
    assert out["identical_on_all_ranks"] and out["presence_allreduce_ok"]


@pytest.mark.gpu
def test_seasonal_sharded_vs_case_parallel_two_gpus():
    """Seasonal mode on 2 GPUs, both ways of using them: every case sharded over the ranks vs the cases
    distributed over the ranks.  The potentials agree to float32 rounding (different hierarchies), the tracks are
    therefore different realisations and the smoothed summary maps agree statistically (normalised L1)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29500 + (os.getpid() + 7) % 400
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "seasonal_run.py"), "400", "480", "100", "4", "2000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["potential_max_diff_ulp"] <= 4.0
    assert out["summary_l1"] <= 0.2, out            # 4 cases x 2000 tracks, smoothed: sampling noise is ~0.1
