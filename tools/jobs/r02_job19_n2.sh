set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_potential.py tests/test_gpu_sharded.py tests/test_config1_parity.py -x -q -m gpu > gpurun_out/r02_pytest_graph.log 2>&1; tail -4 gpurun_out/r02_pytest_graph.log
python tools/solver_run.py 2>&1 | tail -1 | cut -c1-220
SSRS_X_NOGRAPH=1 python tools/solver_run.py 2>&1 | tail -1 | cut -c1-220
$TR --master-port 29531 tools/sharded_solve_run.py 5000 6000 10 3 2>/dev/null | grep "^{" > gpurun_out/r02_sharded_graph_n2.json; cat gpurun_out/r02_sharded_graph_n2.json | cut -c1-700
SSRS_X_NOGRAPH=1 $TR --master-port 29532 tools/sharded_solve_run.py 5000 6000 10 3 2>/dev/null | grep "^{" > gpurun_out/r02_sharded_nograph_n2.json; cat gpurun_out/r02_sharded_nograph_n2.json | cut -c1-700
