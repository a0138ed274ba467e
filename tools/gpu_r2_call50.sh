#!/bin/bash
# round 2, call 50: transpose parity with the new geometries and carried windows; C3 timing of the default; fuzz
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "transpose" > gpurun_out/pytest_gpu50.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu50.log
timeout -k 10 600 python tools/transpose_carry_probe.py --configs "1:0,1:0:256x2048" > gpurun_out/transpose_default_probe.jsonl 2> gpurun_out/transpose_default_probe.err
echo "probe rc=$?"; cat gpurun_out/transpose_default_probe.jsonl
timeout -k 10 600 python tools/transpose_fuzz.py > gpurun_out/transpose_fuzz3.log 2>&1
echo "fuzz rc=$?"; tail -2 gpurun_out/transpose_fuzz3.log | cut -c1-300
