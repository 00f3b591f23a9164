set -x
python -m pytest tests/test_gpu_sharded.py tests/test_trackio.py -x -q -m gpu > gpurun_out/r02_pytest_sharded_n2b.log 2>&1; tail -5 gpurun_out/r02_pytest_sharded_n2b.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_d_c3_n2.json 2> gpurun_out/r02_bench_d_c3_n2.err || tail -30 gpurun_out/r02_bench_d_c3_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_d_c3_n2.json')); print('config3 n2', d['value'], d['e2e']['value'], d['ms_per_step'], d['fields'].get('sharded_vs_single_max_ulp'))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --workload config2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_d_c2_n2.json 2> gpurun_out/r02_bench_d_c2_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_d_c2_n2.json')); print('config2 n2', d['value'], d['e2e']['value'], d['ms_per_step'])"
