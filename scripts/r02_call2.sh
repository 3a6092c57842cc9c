#!/bin/bash
# round 2, GPU call 2: tcgen05 3xTF32 bring-up, GPU suite with the promoted defaults, remaining A/Bs
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call2_summary.txt; : > $S
say() { echo "$@" | tee -a $S; }
say "== tf32x3 bring-up =="
timeout 180 python scripts/tf32x3_check.py --time > $OUT/tf32x3_check.log 2>&1; say "tf32x3_check rc=$?"; tail -22 $OUT/tf32x3_check.log | tee -a $S
say "== GPU suite (new defaults, nothing staged) =="
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r02_gpu_tests.log 2>&1; say "gpu tests rc=$?"; tail -4 $OUT/r02_gpu_tests.log | tee -a $S
say "== dpotrf N=16384: look-ahead split off/on under the new defaults =="
for SP in 0 1; do CHOL_LOOKAHEAD_SPLIT=$SP CFG_OUT=cfg1_split$SP.json timeout 300 python scripts/configs_bench.py cfg1 2>&1 | grep -E "^dpotrf" | sed "s/^/LOOKAHEAD_SPLIT=$SP /" | tee -a $S; done
say "== spotrf N=32768: FFMA vs tcgen05 3xTF32 updates =="
for T in 0 1; do CHOL_S_TENSOR=$T CFG2_PRECS=s CFG_OUT=cfg2_s_tensor$T.json timeout 300 python scripts/configs_bench.py cfg2 2>&1 | grep -E "potrf" | sed "s/^/S_TENSOR=$T /" | tee -a $S; done
say "== done =="
