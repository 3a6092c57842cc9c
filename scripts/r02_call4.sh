#!/bin/bash
# round 2, GPU call 4 (1 GPU): substitution-based panel chain (new potf2 with in-kernel look-ahead, trsm_rows; no inverse
# in POTRF): kernel timings, parity suite, configs, ncu source-level capture of potf2, headline bench with the new bench.py
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call4_summary.txt; : > $S
say() { echo "$@" | tee -a $S; }
say "== kernel timings =="
timeout 300 python scripts/potf2_ab.py 2>&1 | tee -a $S | tail -40
say "== GPU suite =="
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r02c4_gpu_tests.log 2>&1; say "gpu tests rc=$?"; tail -6 $OUT/r02c4_gpu_tests.log | tee -a $S
say "== configs =="
CFG_OUT=r02c4_configs.json timeout 900 python scripts/configs_bench.py cfg0 cfg1 cfg2 cfg3 2>&1 | grep -E "potr" | tee -a $S
say "== launch list N=8192 =="
ONE_N=8192 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/r02c4_launches_dpotrf_n8192.csv \
    python scripts/one_potrf.py > $OUT/r02c4_ncu_n8192.log 2>&1; say "ncu N=8192 rc=$?"
say "== ncu --set full, source level: potf2 and trsm_rows (N=2048 potrf) =="
ONE_N=2048 ONE_REPS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"potf2_kernel|trsm_rows" -s 2 -c 4 -o $OUT/r02c4_panel_kernels \
    python scripts/one_potrf.py > $OUT/r02c4_ncu_panel.log 2>&1; say "ncu panel rc=$?"
say "== bench N=65536 =="
timeout 900 python bench.py --steps 2 --warmup 3 > $OUT/r02c4_bench_n65536.json 2> $OUT/r02c4_bench_n65536.err; say "bench rc=$?"
tail -c 4000 $OUT/r02c4_bench_n65536.json | tee -a $S; tail -5 $OUT/r02c4_bench_n65536.err | tee -a $S
say "== done =="
