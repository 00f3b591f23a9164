set -x
for v in 0 1 2; do
  SSRS_B200_LIB=$PWD/variants/libssrs_lut$v.so timeout 600 python -m pytest tests/test_gpu_tracks.py -x -q 2>&1 | tail -2
  SSRS_B200_LIB=$PWD/variants/libssrs_lut$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_lut$v.json 2> gpurun_out/r02_lut$v.err || tail -5 gpurun_out/r02_lut$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_lut$v.json')); print('LUT$v', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_ms_alone'])"
done
for v in 1 2 0; do
  SSRS_B200_LIB=$PWD/variants/libssrs_lut$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_lut${v}b.json 2> gpurun_out/r02_lut${v}b.err || tail -5 gpurun_out/r02_lut${v}b.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_lut${v}b.json')); print('LUT${v}b', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_ms_alone'])"
done
timeout 400 python tools/solver_truth_large.py > gpurun_out/r02_solver_truth_large.txt 2>&1; cat gpurun_out/r02_solver_truth_large.txt
