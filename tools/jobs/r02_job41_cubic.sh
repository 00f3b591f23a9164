#!/bin/bash
timeout 30 python tools/cubic_run.py > gpurun_out/r02_cubic_run.txt 2> gpurun_out/r02_cubic_run.err || { tail -5 gpurun_out/r02_cubic_run.err; exit 1; }
timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cubic.csv python tools/cubic_run.py > gpurun_out/r02_ncu_cubic.log 2>&1
cat gpurun_out/r02_cubic_run.txt; tail -3 gpurun_out/r02_ncu_cubic.log
