# round 2, GPU call 1: baselines, table-walk ceiling microbenchmark, ncu of the bulk stepping regime and of the shipped stencil
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
(cd tools/ubench && ./walk) > gpurun_out/r02_walk_ubench.log 2>&1
tail -40 gpurun_out/r02_walk_ubench.log
python tools/step_latency.py > gpurun_out/r02_step_latency_base.log 2>&1
tail -8 gpurun_out/r02_step_latency_base.log
python tools/step_bulk.py 1000000 > gpurun_out/r02_bulk_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:step_tracks --launch-skip 1 --launch-count 1 -o gpurun_out/r02_prof_tracks_bulk python tools/step_bulk.py 1000000 > gpurun_out/r02_ncu_bulk.log 2>&1
tail -3 gpurun_out/r02_bulk_plain.log gpurun_out/r02_ncu_bulk.log
python tools/stencil_run.py > gpurun_out/r02_stencil_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:updraft --launch-skip 4 --launch-count 3 -o gpurun_out/r02_prof_updraft python tools/stencil_run.py > gpurun_out/r02_ncu_stencil.log 2>&1
tail -5 gpurun_out/r02_stencil_plain.log
