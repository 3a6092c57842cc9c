#!/bin/bash
# round 2, GPU call 3 (1 GPU): GPU suite on the current tree, BASELINE configs at their stated sizes with raw JSON,
# launch lists of one cuDpotrf at N = 8192 / 16384 (the serial panel chain), headline bench.
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call3_summary.txt; : > $S
say() { echo "$@" | tee -a $S; }
say "== GPU suite =="
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r02c3_gpu_tests.log 2>&1; say "gpu tests rc=$?"; tail -4 $OUT/r02c3_gpu_tests.log | tee -a $S
say "== configs =="
CFG_OUT=r02c3_configs.json timeout 900 python scripts/configs_bench.py cfg0 cfg1 cfg2 cfg3 2>&1 | grep -E "potr" | tee -a $S
say "== launch lists (ncu, serialised, cold) =="
for N in 8192 16384; do
  ONE_N=$N timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/r02c3_launches_dpotrf_n$N.csv \
    python scripts/one_potrf.py > $OUT/r02c3_ncu_n$N.log 2>&1; say "ncu N=$N rc=$?"
done
say "== bench N=65536 =="
timeout 900 python bench.py --steps 2 --warmup 3 > $OUT/r02c3_bench_n65536.json 2> $OUT/r02c3_bench_n65536.err; say "bench rc=$?"
tail -c 3000 $OUT/r02c3_bench_n65536.json | tee -a $S
say "== done =="
