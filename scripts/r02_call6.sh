#!/bin/bash
# round 2, GPU call 6 (1 GPU): warp-panel potf2, rolled trsm_rows, c-precision tensor path (3xTF32 real embedding)
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call6_summary.txt; : > $S
say() { echo "$@" | tee -a $S; }
say "== kernel timings =="
AB_PRECS=dzsc timeout 300 python scripts/potf2_ab.py 2>&1 | grep -v latency | tee -a $S | tail -40
say "== GPU tests =="
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r02c6_gpu_tests.log 2>&1; say "gpu tests rc=$?"; tail -25 $OUT/r02c6_gpu_tests.log | tee -a $S
say "== configs =="
CFG2_PRECS=sc CFG_OUT=r02c6_configs.json timeout 900 python scripts/configs_bench.py cfg0 cfg1 cfg2 2>&1 | grep -E "potr" | tee -a $S
say "== launch list N=8192 =="
ONE_N=8192 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/r02c6_launches_dpotrf_n8192.csv \
    python scripts/one_potrf.py > $OUT/r02c6_ncu_n8192.log 2>&1; say "ncu N=8192 rc=$?"
say "== ncu --set full, source level, dense sampling: potf2, trsm_rows (N=1024 potrf) =="
ONE_N=1024 ONE_REPS=1 timeout 600 ncu --set full --import-source on --clock-control none --warp-sampling-interval 0 -k regex:"potf2_kernel" -s 2 -c 1 -o $OUT/r02c6_potf2 \
    python scripts/one_potrf.py > $OUT/r02c6_ncu_potf2.log 2>&1; say "ncu potf2 rc=$?"
ONE_N=2048 ONE_REPS=1 timeout 600 ncu --set full --import-source on --clock-control none --warp-sampling-interval 0 -k regex:"trsm_rows" -s 4 -c 2 -o $OUT/r02c6_trsm_rows \
    python scripts/one_potrf.py > $OUT/r02c6_ncu_trsm.log 2>&1; say "ncu trsm_rows rc=$?"
say "== done =="
