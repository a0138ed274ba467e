#!/bin/bash
# round 2, call 51: values loaded beside the rows (registers) in the 256x1024 / 512x2048 geometries: parity + C3 timing
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "transpose" > gpurun_out/pytest_gpu51.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu51.log
timeout -k 10 600 python tools/transpose_carry_probe.py --configs "1:0,1:0:256x2048,1:0:512x2048,1:400:256x1024" > gpurun_out/transpose_earlyx_probe.jsonl 2> gpurun_out/transpose_earlyx_probe.err
echo "probe rc=$?"; cat gpurun_out/transpose_earlyx_probe.jsonl
