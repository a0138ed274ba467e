#!/bin/bash
# round 2, call 52: 256x512 geometry (6 CTAs per SM, 40 registers) at C3
mkdir -p gpurun_out
timeout -k 10 600 python tools/transpose_carry_probe.py --configs "1:0,1:0:256x512,1:400:256x512,1:200:256x512,0:0:256x512" > gpurun_out/transpose_512_probe.jsonl 2> gpurun_out/transpose_512_probe.err
echo "probe rc=$?"; cat gpurun_out/transpose_512_probe.jsonl; tail -2 gpurun_out/transpose_512_probe.err
