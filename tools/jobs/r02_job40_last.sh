#!/bin/bash
# last GPU call of round 2: the new cache-key test, the snapshot test with 'cubic', the full-size interpolation test
# (prints the cubic timing), then BASELINE configs[0] reference-vs-GPU on the same box
timeout 50 python -m pytest tests/test_gpu_simulator.py tests/test_wind_thermals.py -m gpu -q -s \
  -k "cache or wind_sites or full_size" 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r02_pytest_gpu_last.log
timeout 55 python tools/config1_compare.py > gpurun_out/r02_config1_compare.json 2> gpurun_out/r02_config1_compare.err
tail -4 gpurun_out/r02_pytest_gpu_last.log; cut -c1-600 gpurun_out/r02_config1_compare.json
