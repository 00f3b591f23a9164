# bench.py first (plain), then its launch list and one --set full capture of the stepping kernel
TAG=${1:-v8}
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_$TAG.csv -k regex:"step_tracks|interleave|updraft" python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench_list.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:step_tracks --launch-skip 3 --launch-count 1 -o gpurun_out/prof_tracks_$TAG python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench_full.log 2>&1
tail -2 gpurun_out/ncu_bench_full.log
python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['fields']['potential_ms'], d['roofline'])"
