for s in 20 24 16; do
  python bench.py --steps 20 --warmup 5 --streams $s --no-cpu-baseline > gpurun_out/r02_streams_$s.json 2> gpurun_out/r02_streams_$s.err || tail -5 gpurun_out/r02_streams_$s.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_streams_$s.json')); print('STREAMS $s', d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['warmup'])"
done
