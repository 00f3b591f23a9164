set -x
SSRS_NO_WARMUP=1 python tools/solver_run.py > gpurun_out/plain_solver.log 2>&1 &&
SSRS_NO_WARMUP=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --kernel-name-base demangled --log-file gpurun_out/r02_solver_launches_all.csv python tools/solver_run.py > gpurun_out/ncu_solver_all.log 2>&1
tail -3 gpurun_out/plain_solver.log; wc -l gpurun_out/r02_solver_launches_all.csv
