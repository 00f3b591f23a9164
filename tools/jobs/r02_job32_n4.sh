TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
SSRS_COMM_HALO=peer timeout 300 $TR --master-port 29571 tools/sharded_solve_run.py 5000 6000 10 3 2>gpurun_out/r02_peer_n4.err | grep "^{" > gpurun_out/r02_sharded_peer_n4.json; cut -c1-900 gpurun_out/r02_sharded_peer_n4.json; tail -3 gpurun_out/r02_peer_n4.err
timeout 300 $TR --master-port 29572 tools/sharded_solve_run.py 5000 6000 10 3 2>/dev/null | grep "^{" > gpurun_out/r02_sharded_ncclhalo_n4.json; cut -c1-900 gpurun_out/r02_sharded_ncclhalo_n4.json
