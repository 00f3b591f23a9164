# ncu --set full of the solver's dominant kernels at 5000 x 6000
python tools/solver_sweep.py SSRS_X_FUSEC=0 SSRS_X_FUSEC=0 SSRS_X_FUSEC=1 2>&1 | tee gpurun_out/sweep5.log
SSRS_NO_WARMUP=1 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"fine_jacobi32|fine_first_residual32|fine_apply_dots" --launch-skip 8 --launch-count 3 -o gpurun_out/prof_solver_fine_r01 python tools/solver_run.py > gpurun_out/ncu_solver5.log 2>&1
SSRS_NO_WARMUP=1 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"ell_first32|ell_residual32|ell_jacobi32|restrict32" --launch-skip 0 --launch-count 3 -o gpurun_out/prof_solver_ell_r01 python tools/solver_run.py > gpurun_out/ncu_solver6.log 2>&1
tail -2 gpurun_out/ncu_solver5.log gpurun_out/ncu_solver6.log
