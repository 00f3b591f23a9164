set -x
python -m pytest tests/test_gpu_updraft.py tests/test_gpu_potential.py::test_golden_potentials tests/test_gpu_simulator.py -x -q -m gpu > gpurun_out/r02_pytest_updraft.log 2>&1; tail -5 gpurun_out/r02_pytest_updraft.log
python tools/stencil_run.py > gpurun_out/r02_stencil_packed.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:updraft --launch-skip 4 --launch-count 2 -o gpurun_out/r02_prof_updraft_packed python tools/stencil_run.py > gpurun_out/r02_ncu_stencil2.log 2>&1
cat gpurun_out/r02_stencil_packed.log
