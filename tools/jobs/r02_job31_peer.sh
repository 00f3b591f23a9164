TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29561 tools/sharded_solve_run.py 1500 1800 30 1 2>gpurun_out/r02_peer_small.err | grep "^{" | cut -c1-900; tail -5 gpurun_out/r02_peer_small.err
timeout 300 $TR --master-port 29562 tools/sharded_solve_run.py 5000 6000 10 3 2>gpurun_out/r02_peer_n2.err | grep "^{" > gpurun_out/r02_sharded_peer_n2.json; cut -c1-900 gpurun_out/r02_sharded_peer_n2.json; tail -3 gpurun_out/r02_peer_n2.err
SSRS_COMM_HALO=nccl timeout 300 $TR --master-port 29563 tools/sharded_solve_run.py 5000 6000 10 3 2>/dev/null | grep "^{" > gpurun_out/r02_sharded_ncclhalo_n2.json; cut -c1-900 gpurun_out/r02_sharded_ncclhalo_n2.json
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -3
