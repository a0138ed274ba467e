#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (row paths)"; timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider -k "row_indexed or golden_reductions or golden_spmv or synth_reductions or cross_identities or linearity" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for ls in 0 1 2 4 8; do
  echo "== C2 lockstep=$ls"; SB200_LOCKSTEP=$ls SB200_ROW_PLAN=1 timeout 300 python tools/opbench.py --workload C2 --ops rowSums,spmv --reps 5 --warmup 2 2>&1 | tail -2 | cut -c100-260
done
for sp in 2 8; do echo "== C2 lockstep=4 splits=$sp"; SB200_SCATTER_SPLITS=$sp SB200_ROW_PLAN=1 timeout 300 python tools/opbench.py --workload C2 --ops rowSums --reps 5 2>&1 | tail -1 | cut -c100-260; done
for wl in C3 C4; do for ls in 0 4; do echo "== $wl lockstep=$ls"; SB200_LOCKSTEP=$ls SB200_ROW_PLAN=1 timeout 600 python tools/opbench.py --workload $wl --ops rowSums,spmv --reps 3 --warmup 1 2>&1 | tail -2 | cut -c100-260; done; done
