set -x
python tools/solver_truth_sweep.py > gpurun_out/r02_solver_truth_sweep.log 2>&1; cat gpurun_out/r02_solver_truth_sweep.log | tail -9
python -m pytest tests -q -m gpu --deselect tests/test_gpu_potential.py::test_refined_truth_10m > gpurun_out/r02_pytest_gpu_b.log 2>&1; tail -8 gpurun_out/r02_pytest_gpu_b.log
python tools/track_length_report.py > gpurun_out/r02_track_lengths.jsonl 2>&1; cat gpurun_out/r02_track_lengths.jsonl
