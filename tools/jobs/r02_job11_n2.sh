set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/r02_pytest_sharded_n2.log 2>&1; tail -5 gpurun_out/r02_pytest_sharded_n2.log
$TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_c3_n2.json 2> gpurun_out/r02_bench_c3_n2.err || tail -30 gpurun_out/r02_bench_c3_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c3_n2.json')); print('config3 n2', d['value'], d['e2e']['value'], d['ms_per_step'], d['fields'].get('sharded_vs_single_max_ulp'), d['fields']['potential_ms'], d['config']['workload'])"
$TR bench.py --gpus 2 --workload config2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c2_n2.json 2> gpurun_out/r02_bench_c2_n2.err || tail -30 gpurun_out/r02_bench_c2_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c2_n2.json')); print('config2 n2', d['value'], d['e2e']['value'], d['ms_per_step'])"
$TR bench.py --gpus 2 --workload config4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c4_n2.json 2> gpurun_out/r02_bench_c4_n2.err || tail -30 gpurun_out/r02_bench_c4_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c4_n2.json')); print('config4 n2', d['value'], d['e2e']['value'], d['ms_per_step'], {k:v for k,v in d['fields'].items() if 'ms' in k})"
$TR bench.py --gpus 2 --workload config5 --steps 3 --warmup 1 > gpurun_out/r02_bench_c5_n2.json 2> gpurun_out/r02_bench_c5_n2.err || tail -30 gpurun_out/r02_bench_c5_n2.err
cut -c1-1200 gpurun_out/r02_bench_c5_n2.json
