#!/bin/bash
# round 2, GPU call 11 (2 GPUs): distributed block-cyclic TRTRI / LAUUM / POTRI (NCCL), new block sizes
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call11_summary.txt; : > $S
say() { echo "$@" | tee -a $S; }
nvidia-smi -L | tee -a $S
say "== multi-GPU tests (all visible devices in ONE process; inverse = distributed sweeps by default) =="
timeout 1200 python -m pytest tests/test_gpu_multigpu.py -m gpu -q -rs > $OUT/r02c11_mgpu_tests.log 2>&1; say "rc=$?"; tail -12 $OUT/r02c11_mgpu_tests.log | tee -a $S
say "== host-pointer POTRI: staged / distributed / slab =="
for N in 16384 32768; do AB_N=$N timeout 600 python scripts/mgpu_inverse_ab.py 2>&1 | grep potri | tee -a $S; done
say "== configs on GPU 0 with the new default block sizes =="
CFG2_PRECS=c CFG_OUT=r02c11_configs.json timeout 900 python scripts/configs_bench.py cfg0 cfg1 cfg2 cfg3 2>&1 | grep -E "potr" | tee -a $S
say "== done =="
