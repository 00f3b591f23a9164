for v in "$@"; do
  SSRS_B200_LIB=$PWD/variants/libssrs_$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_var_$v.json 2> gpurun_out/r02_var_$v.err || tail -5 gpurun_out/r02_var_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_var_$v.json')); print('VAR $v', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_ms_alone'])"
done
