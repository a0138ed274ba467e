#!/bin/bash
# row-indexed ops: parity on both paths, then banded-vs-L2 timing on the three big configs
set +e
mkdir -p gpurun_out
echo "== pytest gpu (row paths)"; timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider -k "row_indexed or golden_reductions or golden_spmv or synth_reductions or cross_identities or linearity" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for wl in C2 C3 C4; do
  for plan in 1 0; do
    echo "== $wl ROW_PLAN=$plan"; SB200_ROW_PLAN=$plan timeout 900 python tools/opbench.py --workload $wl --ops rowSums,spmv --reps 5 --warmup 2 --tag plan$plan 2>&1 | tail -2 | cut -c1-40,100-260
  done
done
for sp in 1 2 8; do echo "== C2 banded splits=$sp"; SB200_ROW_PLAN=1 SB200_SCATTER_SPLITS=$sp timeout 600 python tools/opbench.py --workload C2 --ops rowSums --reps 5 2>&1 | tail -1 | cut -c100-260; done
echo "== C2 banded bands=296 splits=2"; SB200_ROW_PLAN=1 SB200_BANDS=296 SB200_SCATTER_SPLITS=2 timeout 600 python tools/opbench.py --workload C2 --ops rowSums --reps 5 2>&1 | tail -1 | cut -c100-260
