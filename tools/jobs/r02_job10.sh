set -x
python -m pytest tests/test_gpu_potential.py tests/test_gpu_updraft.py tests/test_gpu_simulator.py -x -q -m gpu -s > gpurun_out/r02_pytest_gpu_c.log 2>&1; grep -E "10 m truth|5000x6000: float64|passed|failed" gpurun_out/r02_pytest_gpu_c.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams 1 > gpurun_out/plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:step_tracks --launch-skip 136 --launch-count 4 -o gpurun_out/r02_prof_step_phased python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams 1 > gpurun_out/ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none --csv --log-file gpurun_out/r02_step_phased_traffic.csv -k regex:step_tracks --launch-skip 136 --launch-count 34 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams 1 > gpurun_out/ncu_traffic.log 2>&1
tail -2 gpurun_out/ncu_traffic.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv -k regex:"step_tracks|interleave|updraft" python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c_n1.json 2> gpurun_out/r02_bench_c_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c_n1.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['fields']['potential_ms'], d['cpu_baseline']['value'])"
