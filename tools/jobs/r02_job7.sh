set -x
python tools/l2_peak.py > gpurun_out/r02_l2_peak.txt 2>&1; cat gpurun_out/r02_l2_peak.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_a_n1.json 2> gpurun_out/r02_bench_a_n1.err || tail -30 gpurun_out/r02_bench_a_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_a_n1.json')); print(d['value'], d['e2e'], d['ms_per_step'], d['roofline']['launch_ms_alone'], d['roofline']['launch_ms_in_flight'], d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline'].get('port',{}).get('value'))"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_a_n1_k5.json 2> gpurun_out/r02_bench_a_n1_k5.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_a_n1_k5.json')); print('K=5', d['value'], d['e2e']['value'], d['ms_per_step'])"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_a_ref.json 2> gpurun_out/r02_bench_a_ref.err || tail -30 gpurun_out/r02_bench_a_ref.err
cut -c1-600 gpurun_out/r02_bench_a_ref.json
