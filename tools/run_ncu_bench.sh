# bench.py first (plain), then its launch list and one --set full capture of the stepping kernel
python bench.py --steps 3 --warmup 3 > gpurun_out/bench11.json 2> gpurun_out/bench11.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_r01.csv -k regex:"step_tracks|interleave|updraft" python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench_list.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:step_tracks --launch-skip 3 --launch-count 1 -o gpurun_out/prof_tracks_r01_final python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench_full.log 2>&1
tail -2 gpurun_out/ncu_bench_full.log
python -c "
import json; d=json.load(open('gpurun_out/bench11.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['fields']['potential_ms'], d['roofline'])"
