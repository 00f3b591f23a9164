set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_d.log 2>&1; tail -6 gpurun_out/r02_pytest_gpu_d.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_d_n1.json 2> gpurun_out/r02_bench_d_n1.err || tail -30 gpurun_out/r02_bench_d_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_d_n1.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['fields']['updraft_ms'], d['fields']['potential_ms'], d['roofline']['frac'], d['roofline_l2'], d['gpu_launches'], d['warmup'], d['cpu_baseline']['value'])"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_d_n1_k5.json 2> gpurun_out/r02_bench_d_n1_k5.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_d_n1_k5.json')); print('K=5', d['value'], d['e2e']['value'], d['ms_per_step'])"
