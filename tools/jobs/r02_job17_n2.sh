set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/r02_pytest_sharded_n2c.log 2>&1; tail -4 gpurun_out/r02_pytest_sharded_n2c.log
$TR --master-port 29521 tools/sharded_solve_run.py 5000 6000 10 3 > gpurun_out/r02_sharded_dist_n2.json 2> gpurun_out/r02_sharded_dist_n2.err || tail -20 gpurun_out/r02_sharded_dist_n2.err
cat gpurun_out/r02_sharded_dist_n2.json
SSRS_X_REDUNDANT_SETUP=1 $TR --master-port 29522 tools/sharded_solve_run.py 5000 6000 10 3 > gpurun_out/r02_sharded_redundant_n2.json 2> gpurun_out/r02_sharded_redundant_n2.err
cat gpurun_out/r02_sharded_redundant_n2.json
$TR --master-port 29523 tools/sharded_solve_run.py 10000 12000 10 2 > gpurun_out/r02_sharded_dist_c5_n2.json 2> gpurun_out/r02_sharded_dist_c5_n2.err || tail -20 gpurun_out/r02_sharded_dist_c5_n2.err
cat gpurun_out/r02_sharded_dist_c5_n2.json
