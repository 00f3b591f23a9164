set -x
python -m pytest tests/test_gpu_walk.py -x -q -m gpu > gpurun_out/r02_pytest_walk2.log 2>&1
tail -5 gpurun_out/r02_pytest_walk2.log
python tools/walk_bulk.py 1000000 > gpurun_out/r02_walk_bulk_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:walk_kernel --launch-skip 0 --launch-count 3 -o gpurun_out/r02_prof_walk_bulk python tools/walk_bulk.py 1000000 > gpurun_out/r02_ncu_walk_bulk.log 2>&1
tail -2 gpurun_out/r02_walk_bulk_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_walk_bulk.csv -k regex:"walk_kernel|transition" python tools/walk_bulk.py 1000000 > /dev/null 2>&1
head -50 gpurun_out/r02_launches_walk_bulk.csv | cut -c1-200
