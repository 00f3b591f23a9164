TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -3
$TR --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_final_c3_n2.json 2> gpurun_out/r02_bench_final_c3_n2.err || tail -30 gpurun_out/r02_bench_final_c3_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_final_c3_n2.json')); print('config3 n2', d['value'], d['e2e']['value'], d['ms_per_step'], d['fields'].get('sharded_vs_single_max_ulp'), d['fields']['potential_ms'])"
$TR --master-port 29552 tools/sharded_solve_run.py 10000 12000 10 2 2>/dev/null | grep "^{" > gpurun_out/r02_sharded_final_c5_n2.json; cut -c1-900 gpurun_out/r02_sharded_final_c5_n2.json
