set -x
python tools/step_pipe.py 20 6,8,12,16 > gpurun_out/r02_pipe_a.log 2>&1; tail -7 gpurun_out/r02_pipe_a.log
SSRS_B200_LIB=$PWD/ssrs_b200/libssrs_b200_minb8.so python tools/step_pipe.py 20 8,12 > gpurun_out/r02_pipe_b.log 2>&1; tail -5 gpurun_out/r02_pipe_b.log
SSRS_X_PHASE_PCT=50 python tools/step_pipe.py 20 8,12 > gpurun_out/r02_pipe_c.log 2>&1; tail -5 gpurun_out/r02_pipe_c.log
SSRS_X_PHASE_PCT=12 python tools/step_pipe.py 20 8,12 > gpurun_out/r02_pipe_d.log 2>&1; tail -5 gpurun_out/r02_pipe_d.log
