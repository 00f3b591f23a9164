set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
$TR bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c3_n8.json 2> gpurun_out/r02_bench_c3_n8.err || tail -30 gpurun_out/r02_bench_c3_n8.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c3_n8.json')); print('config3 n8', d['value'], d['e2e']['value'], d['ms_per_step'], d['fields'].get('sharded_vs_single_max_ulp'), d['fields']['potential_ms'], d['fields']['potential_first_ms'])"
$TR bench.py --gpus 8 --workload config5 --steps 4 --warmup 1 > gpurun_out/r02_bench_c5_n8_sharded.json 2> gpurun_out/r02_bench_c5_n8_sharded.err || tail -30 gpurun_out/r02_bench_c5_n8_sharded.err
cut -c1-300 gpurun_out/r02_bench_c5_n8_sharded.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c5_n8_sharded.json')); print(d['value'], d['ms_per_step'], d['fields'])"
$TR bench.py --gpus 8 --workload config5 --steps 16 --warmup 8 --seasonal-mode case_parallel > gpurun_out/r02_bench_c5_n8_caseparallel.json 2> gpurun_out/r02_bench_c5_n8_caseparallel.err || tail -30 gpurun_out/r02_bench_c5_n8_caseparallel.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c5_n8_caseparallel.json')); print(d['value'], d['ms_per_step'], d['fields'])"
