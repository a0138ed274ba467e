#!/bin/bash
# evidence for the bench line: the bench itself, the launch list of the same command, and one full capture of
# the dominant kernels (the sweeps; band_scatter_kernel is what the row sums run on before the row-ordered copy)
set +e
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'band_scatter_kernel|sweep_kernel' -s 14 -c 6 -o gpurun_out/prof_bench_final2 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
