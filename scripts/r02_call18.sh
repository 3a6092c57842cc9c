#!/bin/bash
# round 2, GPU call 18 (1 GPU, ~1 min): the A/B switch CHOL_PANEL_PIPELINE=0 (panel solve on the panel stream) on hardware
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call18_summary.txt; : > $S
CHOL_PANEL_PIPELINE=0 timeout 120 python -m pytest tests/test_gpu_lapack.py -m gpu -q -k "potrf_vs_oracle or mid_size" > $OUT/r02c18_tests.log 2>&1; echo "CHOL_PANEL_PIPELINE=0 parity rc=$?" | tee -a $S; tail -2 $OUT/r02c18_tests.log | tee -a $S
