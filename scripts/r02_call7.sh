#!/bin/bash
# round 2, GPU call 7 (1 GPU): potf2 with the in-register diagonal block, single-batch staging, k-chunk 32 in the small GEMM
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call7_summary.txt; : > $S
say() { echo "$@" | tee -a $S; }
say "== kernel timings =="
AB_PRECS=dzsc timeout 300 python scripts/potf2_ab.py 2>&1 | grep -v latency | tee -a $S | tail -40
say "== GPU tests =="
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r02c7_gpu_tests.log 2>&1; say "gpu tests rc=$?"; tail -8 $OUT/r02c7_gpu_tests.log | tee -a $S
say "== configs =="
CFG_OUT=r02c7_configs.json timeout 900 python scripts/configs_bench.py cfg0 cfg1 2>&1 | grep -E "potr" | tee -a $S
say "== launch list N=8192 =="
ONE_N=8192 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/r02c7_launches_dpotrf_n8192.csv \
    python scripts/one_potrf.py > $OUT/r02c7_ncu_n8192.log 2>&1; say "ncu N=8192 rc=$?"
say "== ncu --set full, source level, dense sampling: potf2 (N=1024 potrf) =="
ONE_N=1024 ONE_REPS=1 timeout 600 ncu --set full --import-source on --clock-control none --warp-sampling-interval 0 -k regex:"potf2_kernel" -s 2 -c 1 -o $OUT/r02c7_potf2 \
    python scripts/one_potrf.py > $OUT/r02c7_ncu_potf2.log 2>&1; say "ncu potf2 rc=$?"
say "== done =="
