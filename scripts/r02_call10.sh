#!/bin/bash
# round 2, GPU call 10 (1 GPU): block-size sweeps with the substitution-based panel chain
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call10_summary.txt; : > $S
say() { echo "$@" | tee -a $S; }
say "== nb sweeps (potrf lower, best of 2) =="
timeout 200 python scripts/nb_sweep.py d 4096 128 256 512 2>&1 | grep potrf | tee -a $S
timeout 200 python scripts/nb_sweep.py d 8192 128 256 384 512 2>&1 | grep potrf | tee -a $S
timeout 200 python scripts/nb_sweep.py d 16384 256 384 512 768 1024 2>&1 | grep potrf | tee -a $S
timeout 300 python scripts/nb_sweep.py d 32768 512 768 1024 2>&1 | grep potrf | tee -a $S
timeout 300 python scripts/nb_sweep.py z 16384 128 256 512 2>&1 | grep potrf | tee -a $S
timeout 300 python scripts/nb_sweep.py s 16384 256 512 1024 2>&1 | grep potrf | tee -a $S
timeout 300 python scripts/nb_sweep.py s 32768 512 1024 2>&1 | grep potrf | tee -a $S
timeout 300 python scripts/nb_sweep.py c 16384 256 512 1024 2>&1 | grep potrf | tee -a $S
say "== done =="
