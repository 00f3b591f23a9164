timeout 300 python -m pytest tests/test_gpu_tracks.py tests/test_gpu_simulator.py -x -q 2>&1 | tail -2
for v in default nohybrid default2; do
  if [ $v = nohybrid ]; then export SSRS_B200_LIB=$PWD/variants/libssrs_$v.so; else unset SSRS_B200_LIB; fi
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_var_$v.json 2> gpurun_out/r02_var_$v.err || tail -5 gpurun_out/r02_var_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_var_$v.json')); print('VAR $v', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_ms_alone'])"
done
