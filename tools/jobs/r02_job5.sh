set -x
python -m pytest tests/test_gpu_tracks.py tests/test_gpu_walk.py -x -q -m gpu > gpurun_out/r02_pytest_phased.log 2>&1
tail -5 gpurun_out/r02_pytest_phased.log
python tools/step_modes.py 12 > gpurun_out/r02_step_modes.log 2>&1
tail -60 gpurun_out/r02_step_modes.log
