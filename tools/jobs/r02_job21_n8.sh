set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29541 tools/sharded_solve_run.py 5000 6000 10 3 2>gpurun_out/r02_sharded_dist_n8.err | grep "^{" > gpurun_out/r02_sharded_dist_n8.json; cut -c1-800 gpurun_out/r02_sharded_dist_n8.json; tail -5 gpurun_out/r02_sharded_dist_n8.err
$TR --master-port 29542 tools/sharded_solve_run.py 10000 12000 10 2 2>/dev/null | grep "^{" > gpurun_out/r02_sharded_dist_c5_n8.json; cut -c1-800 gpurun_out/r02_sharded_dist_c5_n8.json
$TR --master-port 29543 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_e_c3_n8.json 2> gpurun_out/r02_bench_e_c3_n8.err || tail -30 gpurun_out/r02_bench_e_c3_n8.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_e_c3_n8.json')); print('config3 n8', d['value'], d['e2e']['value'], d['ms_per_step'], d['fields'].get('sharded_vs_single_max_ulp'), d['fields']['potential_ms'], d['fields']['potential_first_ms'], d['fields']['potential_stats']['setup_ms'], d['fields']['potential_stats']['solve_ms'])"
