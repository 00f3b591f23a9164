"""Multi-rank host logic on CPU (gloo, world_size 2): block partition by global track id, presence all-reduce,
track gathering.  Each rank steps its shard with the C oracle (standing in for its GPU); the summed presence
must be bit-identical to a single-process run because the RNG is keyed by the global track id."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from ssrs_b200 import dist as D
from oracle import oracle_c as OC
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=int(sys.argv[4]))
rng = np.random.RandomState(0)
rows, cols, n = 60, 80, 101
U = (rng.rand(rows, cols) * (rng.rand(rows, cols) > 0.4)).astype(np.float32)
P = np.linspace(1000, 0, rows, dtype=np.float32)[:, None] + rng.rand(rows, cols).astype(np.float32)
starts = np.stack([rng.randint(2, 10, n), rng.randint(2, cols - 2, n)], 1).astype(np.int32)
lo, hi = D.shard_range(n, D.rank(), D.world_size())
out = OC.step_tracks(U, P, (rows, cols), starts[lo:hi], 0.0, 1, 1.0, seed=5, track_id0=lo, traj_cap=400)
pres = D.allreduce_sum(torch.from_numpy(out["presence"].copy()))
steps = D.allreduce_sum(torch.tensor([out["total_steps"]], dtype=torch.int64))
tracks = D.gather_tracks([out["traj"][t, :out["traj_len"][t]] for t in range(hi - lo)])
D.barrier()
if D.rank() == 0:
    whole = OC.step_tracks(U, P, (rows, cols), starts, 0.0, 1, 1.0, seed=5, track_id0=0, traj_cap=400)
    assert np.array_equal(pres.numpy(), whole["presence"])
    assert int(steps[0]) == whole["total_steps"]
    assert len(tracks) == n and all(np.array_equal(tracks[t], whole["traj"][t, :whole["traj_len"][t]]) for t in range(n))
    print("DIST_OK")
dist.destroy_process_group()
'''


def test_shard_range_partition():
    from ssrs_b200 import dist as D
    for n in (0, 1, 7, 100, 1_000_003):
        for w in (1, 2, 3, 4, 8):
            blocks = [D.shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert D.rank() == 0 and D.world_size() == 1


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_presence_allreduce(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = 29600 + world + (os.getpid() % 200)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), str(world)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "DIST_OK" in outs[0]
