set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err || tail -20 gpurun_out/r02_bench_final_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_final_n1.json')); print('FINAL', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline_l2']['frac'], d['cpu_baseline'], d['fields']['potential_ms'], d['fields']['updraft_ms'], d['clocks'])"
python bench.py --impl reference > gpurun_out/r02_bench_final_reference.json 2> gpurun_out/r02_bench_final_reference.err || tail -20 gpurun_out/r02_bench_final_reference.err
cut -c1-600 gpurun_out/r02_bench_final_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_final.log 2>&1; wc -l gpurun_out/r02_launches_bench_final.csv
