#!/bin/bash
# First GPU call after a CPU-only stretch: verify on hardware everything that was written without it
# (tests marked `staged`, library paths behind environment switches - DESIGN.md §4.1) and A/B it.
#   gpurun --timeout 3000 -- 'bash scripts/verify_staged.sh'            # everything, ~45 min
#   gpurun --timeout 1200 -- 'STAGES="tests chain" bash scripts/verify_staged.sh'
# Stages, in order of importance: tests chain engine e2e fp32 inverse memcheck (and `scale`, only on request, under gpurun --gpus N).  Every command runs under its
# own `timeout`; results land in gpurun_out/staged_* with a one-page gpurun_out/staged_summary.txt.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
STAGES=${STAGES:-"tests chain engine e2e fp32 inverse memcheck"}
SUM=$OUT/staged_summary.txt
: > $SUM
say() { echo "$@" | tee -a $SUM; }
has() { [[ " $STAGES " == *" $1 "* ]]; }

bench_line() {  # file N label
  python - "$1" "$2" "$3" <<'PY' | tee -a $SUM
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("N=%s %s: value %.0f  e2e %.0f GFLOP/s  info %s" % (sys.argv[2], sys.argv[3], d["value"], d["e2e"]["value"], d["e2e"]["info"]))
except Exception as e:
    print("N=%s %s: FAILED %r" % (sys.argv[2], sys.argv[3], e))
PY
}

if has tests; then
  say "== regular GPU tests (two default kernels and the host runtime changed since the last hardware run) =="
  timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/staged_gpu_tests.log 2>&1
  say "gpu tests rc=$?"; tail -3 $OUT/staged_gpu_tests.log | tee -a $SUM
  say "== staged tests (new entry points, reference test inputs, C unit program, switched paths) =="
  CHOL_TEST_STAGED=1 timeout 900 python -m pytest tests -m gpu -k staged -q -x > $OUT/staged_tests.log 2>&1
  say "staged tests rc=$?"; tail -5 $OUT/staged_tests.log | tee -a $SUM
fi

if has chain; then
  say "== panel chain: diagonal-block kernel V2 (CHOL_POTF2_V2), small-tile GEMM (CHOL_SMALL_GEMM), blocked panel solve (CHOL_PANEL_BLOCKED), split look-ahead (CHOL_LOOKAHEAD_SPLIT) =="
  CHOL_POTF2_V2=1 CHOL_SMALL_GEMM=1 timeout 1200 python -m pytest tests/test_gpu_blas.py tests/test_gpu_lapack.py tests/test_gpu_multigpu.py -m gpu -q -x \
    > $OUT/staged_chain_tests.log 2>&1
  say "parity with both switches on rc=$?"; tail -2 $OUT/staged_chain_tests.log | tee -a $SUM
  for V in 0 1; do
    CHOL_POTF2_V2=$V timeout 200 python scripts/potf2_ab.py 2>&1 | grep potf2_inv | tee -a $SUM
  done
  CHOL_PANEL_BLOCKED=1 timeout 900 python -m pytest tests/test_gpu_lapack.py tests/test_gpu_multigpu.py -m gpu -q -x > $OUT/staged_blocked_tests.log 2>&1
  say "parity with the blocked panel solve (CHOL_PANEL_BLOCKED) rc=$?"; tail -2 $OUT/staged_blocked_tests.log | tee -a $SUM
  CHOL_LOOKAHEAD_SPLIT=1 timeout 900 python -m pytest tests/test_gpu_lapack.py tests/test_gpu_multigpu.py -m gpu -q -x > $OUT/staged_split_tests.log 2>&1
  say "parity with the split look-ahead (CHOL_LOOKAHEAD_SPLIT) rc=$?"; tail -2 $OUT/staged_split_tests.log | tee -a $SUM
  for CFG in "0 0 0 0" "1 0 0 0" "0 1 0 0" "0 0 1 0" "0 0 0 1" "1 1 1 1"; do
    set -- $CFG
    CHOL_POTF2_V2=$1 CHOL_SMALL_GEMM=$2 CHOL_PANEL_BLOCKED=$3 CHOL_LOOKAHEAD_SPLIT=$4 timeout 300 python scripts/configs_bench.py cfg1 2>&1 \
      | grep -E "^dpotrf" | sed "s/^/POTF2_V2=$1 SMALL_GEMM=$2 PANEL_BLOCKED=$3 LOOKAHEAD_SPLIT=$4 /" | tee -a $SUM
  done
fi

if has chain; then
  say "== TRTRI look-ahead (CHOL_TRTRI_LOOKAHEAD): parity, then dpotrf + dpotri N=32768 =="
  CHOL_TRTRI_LOOKAHEAD=1 timeout 600 python -m pytest tests/test_gpu_lapack.py -m gpu -q -x -k "trtri or potri" > $OUT/staged_trtri_la_tests.log 2>&1
  say "parity rc=$?"; tail -2 $OUT/staged_trtri_la_tests.log | tee -a $SUM
  for LA in 0 1; do
    CHOL_TRTRI_LOOKAHEAD=$LA timeout 400 python scripts/configs_bench.py cfg3 2>&1 | grep -iE "potri|trtri" | sed "s/^/TRTRI_LOOKAHEAD=$LA /" | tee -a $SUM
  done
fi

if has engine; then
  say "== tile engine: staggered kernel with the C tile staged through the ring (CHOL_GEMM_PP=2) =="
  CHOL_GEMM_PP=2 timeout 1200 python -m pytest tests/test_gpu_blas.py tests/test_gpu_lapack.py tests/test_gpu_multigpu.py -m gpu -q -x \
    > $OUT/staged_ppc_tests.log 2>&1
  say "parity with CHOL_GEMM_PP=2 rc=$?"; tail -2 $OUT/staged_ppc_tests.log | tee -a $SUM
  for PP in 0 1 2; do
    CHOL_GEMM_PP=$PP timeout 400 python bench.py --order 65536 --steps 2 --warmup 3 --e2e-steps 1 --skip-cpu-baseline \
      > $OUT/staged_pp${PP}_n65536.json 2> $OUT/staged_pp${PP}_n65536.err
    bench_line $OUT/staged_pp${PP}_n65536.json 65536 "CHOL_GEMM_PP=$PP (dpotrf)"
  done
  for PP in 1 2; do
    CHOL_GEMM_PP=$PP timeout 400 python scripts/configs_bench.py cfg2 2>&1 | grep -E "^zpotrf" | sed "s/^/GEMM_PP=$PP /" | tee -a $SUM
  done
fi

if has e2e; then
  say "== e2e: upload overlapped with the factorisation (CHOL_MGPU_H2D_OVERLAP) =="
  for N in 32768 65536; do
    for OV in 0 1; do
      CHOL_MGPU_H2D_OVERLAP=$OV timeout 400 python bench.py --order $N --steps 1 --warmup 3 --e2e-steps 2 --skip-cpu-baseline \
        > $OUT/staged_e2e_n${N}_ov${OV}.json 2> $OUT/staged_e2e_n${N}_ov${OV}.err
      bench_line $OUT/staged_e2e_n${N}_ov${OV}.json $N "overlap=$OV"
    done
  done
fi

if has fp32; then
  say "== FP32 engine with the deeper k-tile (CHOL_FFMA_BK=16): parity, then s/c potrf N=32768 =="
  CHOL_FFMA_BK=16 timeout 900 python -m pytest tests/test_gpu_blas.py tests/test_gpu_lapack.py -m gpu -q -x > $OUT/staged_ffma_bk16_tests.log 2>&1
  say "BK=16 parity rc=$?"; tail -2 $OUT/staged_ffma_bk16_tests.log | tee -a $SUM
  for BK in 8 16; do
    CHOL_FFMA_BK=$BK timeout 400 python scripts/configs_bench.py cfg2 2>&1 | grep -E "^[sc]potrf" | sed "s/^/FFMA_BK=$BK /" | tee -a $SUM
  done
fi

if has inverse; then
  say "== host-pointer POTRI: staged on one GPU vs slab-distributed (CHOL_MGPU_INVERSE=slab) =="
  timeout 600 python scripts/mgpu_inverse_ab.py 2>&1 | grep -E "^potri" | tee -a $SUM
fi

if has memcheck; then
  say "== memcheck of a small slice (after the same command has passed plain) =="
  SLICE='tests/test_gpu_lapack.py -q -x -k "test_potrf_vs_oracle and (257 or 300) or test_trtri_vs_oracle and 300 or test_lauum_potri_logdet_vs_oracle and 257"'
  if eval timeout 300 python -m pytest $SLICE > $OUT/staged_slice_plain.log 2>&1; then
    eval timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $SLICE > $OUT/staged_memcheck.log 2>&1
    say "memcheck rc=$?"
    grep -E "ERROR SUMMARY|passed|failed" $OUT/staged_memcheck.log | tail -3 | tee -a $SUM
  else
    say "slice failed without the sanitizer; memcheck skipped"
  fi
fi
if has scale; then  # not in the default list: run under `gpurun --gpus N`, e.g. STAGES=scale bash scripts/verify_staged.sh
  NG=$(python -c "import torch; print(torch.cuda.device_count())")
  say "== multi-GPU ($NG GPUs): default vs the shorter panel chain (POTF2_V2 + SMALL_GEMM + PANEL_BLOCKED + LOOKAHEAD_SPLIT) =="
  for MODE in default chain; do
    if [ $MODE = chain ]; then export CHOL_POTF2_V2=1 CHOL_SMALL_GEMM=1 CHOL_PANEL_BLOCKED=1 CHOL_LOOKAHEAD_SPLIT=1; fi
    timeout 900 python bench.py --gpus $NG --order 65536 --steps 2 --warmup 3 --e2e-steps 1 --skip-cpu-baseline \
      > $OUT/staged_scale_${NG}gpu_$MODE.json 2> $OUT/staged_scale_${NG}gpu_$MODE.err
    bench_line $OUT/staged_scale_${NG}gpu_$MODE.json 65536 "$NG GPUs, $MODE"
  done
  unset CHOL_POTF2_V2 CHOL_SMALL_GEMM CHOL_PANEL_BLOCKED CHOL_LOOKAHEAD_SPLIT
fi
say "== done =="
