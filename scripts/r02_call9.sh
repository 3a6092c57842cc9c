#!/bin/bash
# round 2, GPU call 9 (1 GPU): panel write-back off the chain, split look-ahead on, two-step trsm_rows loop; potf2 phase timestamps
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call9_summary.txt; : > $S
say() { echo "$@" | tee -a $S; }
say "== kernel timings and potf2 phases =="
AB_PRECS=d timeout 300 python scripts/potf2_ab.py 2>&1 | grep -v latency | tee -a $S | tail -40
say "== GPU tests (lapack, multigpu, robustness) =="
timeout 1500 python -m pytest tests/test_gpu_lapack.py tests/test_gpu_multigpu.py tests/test_gpu_robustness.py -m gpu -q > $OUT/r02c9_gpu_tests.log 2>&1; say "gpu tests rc=$?"; tail -4 $OUT/r02c9_gpu_tests.log | tee -a $S
say "== configs =="
for SP in 1 0; do CHOL_LOOKAHEAD_SPLIT=$SP CFG_OUT=r02c9_configs_split$SP.json timeout 900 python scripts/configs_bench.py cfg0 cfg1 2>&1 | grep -E "potr" | sed "s/^/SPLIT=$SP /" | tee -a $S; done
CFG2_PRECS=z CFG_OUT=r02c9_configs_z_potri.json timeout 900 python scripts/configs_bench.py cfg2 cfg3 2>&1 | grep -E "potr" | tee -a $S
say "== launch list N=8192 =="
ONE_N=8192 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/r02c9_launches_dpotrf_n8192.csv \
    python scripts/one_potrf.py > $OUT/r02c9_ncu_n8192.log 2>&1; say "ncu N=8192 rc=$?"
say "== done =="
