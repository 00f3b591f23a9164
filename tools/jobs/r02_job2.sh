set -x
python -m pytest tests/test_gpu_walk.py tests/test_config1_parity.py tests/test_gpu_tracks.py -x -q -m gpu > gpurun_out/r02_pytest_walk.log 2>&1
tail -15 gpurun_out/r02_pytest_walk.log
python tools/walk_bench.py 10 > gpurun_out/r02_walk_bench.log 2>&1
tail -30 gpurun_out/r02_walk_bench.log
