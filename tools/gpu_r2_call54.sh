#!/bin/bash
# round 2, call 54: bitmap-rank kernel with the columns per thread as a template argument: parity + C3 timing
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "transpose" > gpurun_out/pytest_gpu54.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu54.log
timeout -k 10 600 python tools/transpose_carry_probe.py --configs "1:0,1:0:256x2048,1:400:256x1024" > gpurun_out/transpose_kc_probe.jsonl 2> gpurun_out/transpose_kc_probe.err
echo "probe rc=$?"; cat gpurun_out/transpose_kc_probe.jsonl
