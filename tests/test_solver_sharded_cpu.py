"""Row-sharded potential solve on CPU: ssrs_b200/csrc/potential.cu compiled with -DSSRS_HOST_EMU, one process per
rank, `ssrs_comm` callbacks over gloo (tests/hostemu.py).  The sharded solve must reproduce the reference's
golden potential and the single-rank solve to float32-rounding level, on every rank (each returns the full raster), and
the distributed setup must build exactly the hierarchy of the redundant one (bit-identical potential, same iterations)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
root, port, rank, world, rep_rows = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
os.environ["SSRS_X_REPROWS"] = rep_rows          # small grids: keep some coarse levels distributed
import hostemu
from oracle import oracle_np as O
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
ULP = float(np.spacing(np.float32(1000.0)))
g = np.load(os.path.join(root, "tests", "golden", "potential.npz"))
comm, calls = hostemu.gloo_comm()
for key, th in (("dem2", 0), ("rand", 90), ("rand", 45), ("dem", 270)):
    K = g[f"{key}_K"]
    if K.shape[0] < 4 * world:
        continue
    bn, bv = O.boundary_nodes(th, *K.shape)
    rc, phi, st, err = hostemu.solve_sharded(K, bn, bv, comm)
    assert rc == 0, err
    ref = g[f"{key}_phi_{th}"].astype(np.float64)
    e = np.abs(phi.astype(np.float64) - ref).max()
    assert e <= 1.5 * ULP, (key, th, e)
    rc1, phi1, st1, _ = hostemu.solve(K, bn, bv)
    assert np.abs(phi.astype(np.float64) - phi1).max() <= 1.0 * ULP
    # every rank holds the same full raster
    t = torch.from_numpy(phi.copy()); dist.broadcast(t, src=0)
    assert np.array_equal(t.numpy(), phi)
    print(f"rank {rank} {key} {th}: iterations {st.iterations} (single {st1.iterations}) levels {st.levels} err {e / ULP:.2f} ulp", flush=True)
# a larger synthetic case with island structure, several distributed levels
rng = np.random.RandomState(3)
K = (rng.rand(192, 160) * (rng.rand(192, 160) > 0.5)).astype(np.float32)
K[60:130, 30:120] = 0.8
bn, bv = O.boundary_nodes(0.0, *K.shape)
n0 = dict(calls)
rc, phi, st, err = hostemu.solve_sharded(K, bn, bv, comm)
assert rc == 0, err
rc1, phi1, st1, _ = hostemu.solve(K, bn, bv)
assert np.abs(phi.astype(np.float64) - phi1).max() <= 2.0 * ULP
assert phi.min() >= 0.0 and phi.max() <= 1000.0
assert calls["exchange"] > n0["exchange"] and calls["allreduce"] > n0["allreduce"] and calls["allgather"] > n0["allgather"]
# the distributed setup (every rank builds its own rows of the distributed levels) against round 1's redundant setup
# (every rank builds everything): the same hierarchy, hence bit-identical iterates
os.environ["SSRS_X_REDUNDANT_SETUP"] = "1"
rc2, phi2, st2, err2 = hostemu.solve_sharded(K, bn, bv, comm)
del os.environ["SSRS_X_REDUNDANT_SETUP"]
assert rc2 == 0, err2
assert st2.iterations == st.iterations and st2.levels == st.levels and list(st2.level_rows[:st.levels]) == list(st.level_rows[:st.levels])
assert np.array_equal(phi2, phi), float(np.abs(phi2.astype(np.float64) - phi).max())
dist.barrier()
if rank == 0:
    print("SHARDED_OK", calls, flush=True)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world,rep_rows", [(2, 200), (3, 200), (8, 100), (2, 65536)])
def test_sharded_solve_gloo(tmp_path, world, rep_rows):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = 29800 + world + (os.getpid() % 150)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), str(world), str(rep_rows)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "SHARDED_OK" in outs[0]
