set -x
python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_e.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu_e.log
python tools/solver_run.py > gpurun_out/r02_solver_run.log 2>&1; tail -1 gpurun_out/r02_solver_run.log | cut -c1-400
