#!/bin/bash
# full GPU suite + smoke after the 'cubic' interpolation and the in-tree scan (bounded: the round's GPU budget is nearly spent)
timeout 150 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_pytest_gpu_final3.log
timeout 30 python __graft_entry__.py smoke > gpurun_out/r02_smoke_final3.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu_final3.log; tail -2 gpurun_out/r02_smoke_final3.log
