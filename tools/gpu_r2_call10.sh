#!/bin/bash
# round 2, call 10: full GPU suite (bitmap-rank transpose, N4 ops), transpose timings, bench N=1, ncu launch list + traffic captures
mkdir -p gpurun_out
export SB200_EXCHANGE_TIMEOUT_S=60
timeout -k 10 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu10.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu10.log
OUT=gpurun_out/opbench10.jsonl; : > $OUT
trun() { local tag=$1; shift; local wl=$1; shift
  env SB200_TRACE=1 "$@" timeout -k 10 300 python tools/opbench.py --workload $wl --ops transpose --reps 6 --tag $tag >> $OUT 2>> gpurun_out/opbench10_trace.err; }
trun bitmap C3
trun bitmap_k2 C3 SB200_TRANSPOSE_KCOLS=2
trun bitmap_4096 C3 SB200_TRANSPOSE_CFG=256x4096
trun bitmap_b444 C3 SB200_TRANSPOSE_BANDS=444
trun bitmap_b200s8 C3 SB200_TRANSPOSE_BANDS=200 SB200_TRANSPOSE_SPLITS=8
trun bitmap_b120s12 C3 SB200_TRANSPOSE_BANDS=120 SB200_TRANSPOSE_SPLITS=12
trun match C3 SB200_TRANSPOSE_RANK=match
trun bitmap C1
trun place_forced C2 SB200_TRANSPOSE_PATH=place
grep -v build $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], d['workload'][:2], d['op'], d['ms_median'], d['frac_measured'])"
timeout -k 10 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1_c.json 2> gpurun_out/bench_n1_c.err
echo "bench rc=$?"
# ncu: launch list of the bench command (same command line first without ncu), then full captures of the two new dominant kernels
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/plain_bench_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_bench_r02.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python tools/opbench.py --workload C4 --ops spmv_t,spmv --reps 2 --bmc 1 > gpurun_out/plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bandsweep -s 2 -c 2 -o gpurun_out/prof_bandsweep_c4 \
  python tools/opbench.py --workload C4 --ops spmv_t,spmv --reps 2 --bmc 1 > gpurun_out/ncu_c4.log 2>&1
echo "ncu c4 rc=$?"
python tools/opbench.py --workload C3 --ops transpose --reps 2 > gpurun_out/plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:transpose_bitrank -s 1 -c 1 -o gpurun_out/prof_bitrank_c3 \
  python tools/opbench.py --workload C3 --ops transpose --reps 2 > gpurun_out/ncu_c3.log 2>&1
echo "ncu c3 rc=$?"
