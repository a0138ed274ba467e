#!/bin/bash
# round 2, call 49: chunk-sort transpose geometries with more warps per SM (256x1024: 4 CTAs, 512x2048: 2 CTAs of 16 warps)
mkdir -p gpurun_out
timeout -k 10 600 python tools/transpose_carry_probe.py --configs "1:0,1:0:256x1024,1:200:256x1024,0:0:256x1024,1:0:512x2048,1:148:512x2048,0:0:512x2048" > gpurun_out/transpose_geom_probe.jsonl 2> gpurun_out/transpose_geom_probe.err
echo "probe rc=$?"; cat gpurun_out/transpose_geom_probe.jsonl; tail -3 gpurun_out/transpose_geom_probe.err
