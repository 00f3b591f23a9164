TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -3
$TR --master-port 29581 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_final2_c3_n2.json 2> gpurun_out/r02_bench_final2_c3_n2.err || tail -30 gpurun_out/r02_bench_final2_c3_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_final2_c3_n2.json')); print('config3 n2', d['value'], d['e2e']['value'], d['ms_per_step'], d['fields'].get('sharded_vs_single_max_ulp'), d['fields']['potential_ms'], d['fields'].get('potential_halo_mode'))"
