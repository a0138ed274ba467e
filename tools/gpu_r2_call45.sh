#!/bin/bash
# round 2, call 45: pinned-chunk size of the pageable staging (512 KB .. 8 MB), one-shot C2 transpose
mkdir -p gpurun_out
for kb in 512 1024 2048 4096 8192; do
  SB200_COPY_CHUNK_KB=$kb timeout -k 10 300 python tools/e2e_transpose_probe.py --reps 6 > gpurun_out/copy_chunk_$kb.log 2>&1
  echo "chunk $kb KB:"; grep "^rep" gpurun_out/copy_chunk_$kb.log | tail -3
done
