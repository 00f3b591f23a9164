timeout 300 python -m pytest tests/test_gpu_tracks.py -x -q 2>&1 | tail -2
for v in default h1; do
  if [ $v = h1 ]; then export SSRS_B200_LIB=$PWD/variants/libssrs_$v.so; else unset SSRS_B200_LIB; fi
  python bench.py --steps 20 --warmup 5 --tracks-per-gpu 125000 --no-cpu-baseline > gpurun_out/r02_var125_$v.json 2> gpurun_out/r02_var125_$v.err || tail -5 gpurun_out/r02_var125_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_var125_$v.json')); print('VAR125 $v', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_ms_alone'])"
done
