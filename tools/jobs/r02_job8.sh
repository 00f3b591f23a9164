set -x
(cd tools/ubench && ./l2bw) > gpurun_out/r02_l2_peak.txt 2>&1; cat gpurun_out/r02_l2_peak.txt
python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu_a.log 2>&1; tail -6 gpurun_out/r02_pytest_gpu_a.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_b_n1.json 2> gpurun_out/r02_bench_b_n1.err || tail -30 gpurun_out/r02_bench_b_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_b_n1.json')); print(d['value'], d['e2e'], d['ms_per_step'], d['roofline']['launch_ms_alone'], d['roofline']['launch_ms_in_flight'])"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_b.csv -k regex:"step_tracks|interleave|updraft" python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:step_tracks --launch-skip 60 --launch-count 6 -o gpurun_out/r02_prof_step_phased python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
