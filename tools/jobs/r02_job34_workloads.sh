for wl in config3 config4; do
  python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_final_${wl}_n1.json 2> gpurun_out/r02_bench_final_${wl}_n1.err || tail -20 gpurun_out/r02_bench_final_${wl}_n1.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_final_${wl}_n1.json')); print('$wl n1', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['fields']['potential_ms'])"
done
python bench.py --workload config5 --steps 3 --warmup 1 > gpurun_out/r02_bench_final_config5_n1.json 2> gpurun_out/r02_bench_final_config5_n1.err || tail -20 gpurun_out/r02_bench_final_config5_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_final_config5_n1.json')); print('config5 n1', d['value'], d['ms_per_step'], d['fields'])"
