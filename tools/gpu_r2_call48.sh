#!/bin/bash
# round 2, call 48: chunk-sort transpose with carry-over (full rounds): parity tests, fuzz, C3 timings per band count
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "transpose" > gpurun_out/pytest_gpu48.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu48.log
timeout -k 10 600 python tools/transpose_carry_probe.py > gpurun_out/transpose_carry_probe.jsonl 2> gpurun_out/transpose_carry_probe.err
echo "probe rc=$?"; cat gpurun_out/transpose_carry_probe.jsonl; tail -3 gpurun_out/transpose_carry_probe.err
