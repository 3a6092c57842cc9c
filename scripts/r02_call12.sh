#!/bin/bash
# round 2, GPU call 12 (1 GPU): micro-optimised chain kernels, full GPU suite, configs, ncu --set full of the DMMA update
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call12_summary.txt; : > $S
say() { echo "$@" | tee -a $S; }
say "== kernel timings =="
AB_PHASES=0 AB_PRECS=d timeout 300 python scripts/potf2_ab.py 2>&1 | grep -v latency | tee -a $S | tail -10
say "== GPU suite =="
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r02c12_gpu_tests.log 2>&1; say "gpu tests rc=$?"; tail -5 $OUT/r02c12_gpu_tests.log | tee -a $S
say "== configs =="
CFG_OUT=r02c12_configs.json timeout 900 python scripts/configs_bench.py cfg0 cfg1 cfg2 cfg3 2>&1 | grep -E "potr" | tee -a $S
say "== launch list N=8192 =="
ONE_N=8192 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/r02c12_launches_dpotrf_n8192.csv \
    python scripts/one_potrf.py > $OUT/r02c12_ncu_n8192.log 2>&1; say "ncu N=8192 rc=$?"
say "== ncu --set full: the shipped DMMA trailing-update kernel at the bench's own shape (lower SYRK n=32768, k=1024) =="
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gemm_dmma -s 1 -c 1 -o $OUT/r02c12_dmma_update \
    python scripts/one_syrk.py > $OUT/r02c12_ncu_dmma.log 2>&1; say "ncu dmma rc=$?"
say "== done =="
