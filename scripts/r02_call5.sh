#!/bin/bash
# round 2, GPU call 5 (1 GPU): staging fixes in the diagonal-block kernels, FP64 latency probes, robustness tests
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call5_summary.txt; : > $S
say() { echo "$@" | tee -a $S; }
say "== latencies and kernel timings =="
AB_PRECS=dz timeout 300 python scripts/potf2_ab.py 2>&1 | tee -a $S | tail -40
say "== GPU tests: lapack, multigpu, robustness =="
timeout 1500 python -m pytest tests/test_gpu_lapack.py tests/test_gpu_multigpu.py tests/test_gpu_robustness.py tests/test_dropin_c.py -m gpu -q > $OUT/r02c5_gpu_tests.log 2>&1; say "gpu tests rc=$?"; tail -25 $OUT/r02c5_gpu_tests.log | tee -a $S
say "== configs =="
CFG_OUT=r02c5_configs.json timeout 900 python scripts/configs_bench.py cfg0 cfg1 2>&1 | grep -E "potr" | tee -a $S
say "== launch list N=8192 =="
ONE_N=8192 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/r02c5_launches_dpotrf_n8192.csv \
    python scripts/one_potrf.py > $OUT/r02c5_ncu_n8192.log 2>&1; say "ncu N=8192 rc=$?"
say "== ncu --set full, source level, dense sampling: potf2 (N=1024 potrf) =="
ONE_N=1024 ONE_REPS=1 timeout 600 ncu --set full --import-source on --clock-control none --sampling-interval 0 -k regex:"potf2_kernel" -s 2 -c 1 -o $OUT/r02c5_potf2 \
    python scripts/one_potrf.py > $OUT/r02c5_ncu_potf2.log 2>&1; say "ncu potf2 rc=$?"
ONE_N=1024 ONE_REPS=1 timeout 600 ncu --set full --import-source on --clock-control none --sampling-interval 0 -k regex:"trsm_rows" -s 2 -c 1 -o $OUT/r02c5_trsm_rows \
    python scripts/one_potrf.py > $OUT/r02c5_ncu_trsm.log 2>&1; say "ncu trsm_rows rc=$?"
say "== done =="
