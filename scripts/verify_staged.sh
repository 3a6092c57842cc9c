#!/bin/bash
# First GPU call after a CPU-only stretch: verify everything that was written without hardware
# (tests marked `staged`, library paths behind environment switches) and A/B it, in ONE gpurun call.
#   gpurun --timeout 1500 -- 'bash scripts/verify_staged.sh'
# Every step runs under its own `timeout`; results land in gpurun_out/staged_*.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT

echo "== staged tests ==" | tee $OUT/staged_summary.txt
CHOL_TEST_STAGED=1 timeout 600 python -m pytest tests -m gpu -k staged -q -x > $OUT/staged_tests.log 2>&1
echo "staged tests rc=$?" | tee -a $OUT/staged_summary.txt
tail -5 $OUT/staged_tests.log | tee -a $OUT/staged_summary.txt

echo "== regular gpu tests (default paths were refactored on the host side) ==" | tee -a $OUT/staged_summary.txt
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/staged_gpu_tests.log 2>&1
echo "gpu tests rc=$?" | tee -a $OUT/staged_summary.txt
tail -3 $OUT/staged_gpu_tests.log | tee -a $OUT/staged_summary.txt

echo "== e2e A/B: upload overlapped with the factorisation ==" | tee -a $OUT/staged_summary.txt
for N in 32768 65536; do
  for OV in 0 1; do
    CHOL_MGPU_H2D_OVERLAP=$OV timeout 400 python bench.py --order $N --steps 1 --warmup 3 --e2e-steps 2 --skip-cpu-baseline \
      > $OUT/staged_e2e_n${N}_ov${OV}.json 2> $OUT/staged_e2e_n${N}_ov${OV}.err
    python - "$OUT/staged_e2e_n${N}_ov${OV}.json" $N $OV <<'PY' | tee -a $OUT/staged_summary.txt
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("N=%s overlap=%s value=%.0f e2e=%.0f GFLOP/s info=%s" % (sys.argv[2], sys.argv[3], d["value"], d["e2e"]["value"], d["e2e"]["info"]))
except Exception as e:
    print("N=%s overlap=%s FAILED %r" % (sys.argv[2], sys.argv[3], e))
PY
  done
done

echo "== host-pointer POTRI: staged on one GPU vs slab-distributed ==" | tee -a $OUT/staged_summary.txt
timeout 600 python scripts/mgpu_inverse_ab.py 2>&1 | tee -a $OUT/staged_summary.txt

echo "== memcheck of a small slice (after the same command has passed plain) ==" | tee -a $OUT/staged_summary.txt
SLICE='tests/test_gpu_lapack.py -q -x -k "test_potrf_vs_oracle and (257 or 300) or test_trtri_vs_oracle and 300 or test_lauum_potri_logdet_vs_oracle and 257"'
if eval timeout 300 python -m pytest $SLICE > $OUT/staged_slice_plain.log 2>&1; then
  eval timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $SLICE > $OUT/staged_memcheck.log 2>&1
  echo "memcheck rc=$?" | tee -a $OUT/staged_summary.txt
  grep -E "ERROR SUMMARY|passed|failed" $OUT/staged_memcheck.log | tail -3 | tee -a $OUT/staged_summary.txt
else
  echo "slice failed without the sanitizer; memcheck skipped" | tee -a $OUT/staged_summary.txt
fi

echo "== FP32 engine with the deeper k-tile (CHOL_FFMA_BK=16): parity, then s/c/z potrf N=32768 A/B ==" | tee -a $OUT/staged_summary.txt
CHOL_FFMA_BK=16 timeout 900 python -m pytest tests/test_gpu_blas.py tests/test_gpu_lapack.py -m gpu -q -x > $OUT/staged_ffma_bk16_tests.log 2>&1
echo "BK=16 parity rc=$?" | tee -a $OUT/staged_summary.txt
tail -2 $OUT/staged_ffma_bk16_tests.log | tee -a $OUT/staged_summary.txt
for BK in 8 16; do
  echo "CHOL_FFMA_BK=$BK" | tee -a $OUT/staged_summary.txt
  CHOL_FFMA_BK=$BK timeout 400 python scripts/configs_bench.py cfg2 2>&1 | grep -E "^[sc]potrf" | tee -a $OUT/staged_summary.txt
done

echo "== diagonal-block kernel V2 (CHOL_POTF2_V2=1): parity, kernel time alone, dpotrf N=16384 ==" | tee -a $OUT/staged_summary.txt
CHOL_POTF2_V2=1 timeout 900 python -m pytest tests/test_gpu_lapack.py tests/test_gpu_multigpu.py -m gpu -q -x > $OUT/staged_potf2_v2_tests.log 2>&1
echo "V2 parity rc=$?" | tee -a $OUT/staged_summary.txt
tail -2 $OUT/staged_potf2_v2_tests.log | tee -a $OUT/staged_summary.txt
for V in 0 1; do
  CHOL_POTF2_V2=$V timeout 200 python scripts/potf2_ab.py 2>&1 | grep potf2_inv | tee -a $OUT/staged_summary.txt
  CHOL_POTF2_V2=$V timeout 300 python scripts/configs_bench.py cfg1 2>&1 | grep -E "^dpotrf" | sed "s/^/V2=$V /" | tee -a $OUT/staged_summary.txt
done

echo "== small-tile FP64 GEMM for the panel chain (CHOL_SMALL_GEMM=1): parity, then dpotrf N=16384 alone and with POTF2 V2 ==" | tee -a $OUT/staged_summary.txt
CHOL_SMALL_GEMM=1 timeout 900 python -m pytest tests/test_gpu_blas.py tests/test_gpu_lapack.py -m gpu -q -x > $OUT/staged_small_gemm_tests.log 2>&1
echo "small GEMM parity rc=$?" | tee -a $OUT/staged_summary.txt
tail -2 $OUT/staged_small_gemm_tests.log | tee -a $OUT/staged_summary.txt
CHOL_SMALL_GEMM=1 timeout 300 python scripts/configs_bench.py cfg1 2>&1 | grep -E "^dpotrf" | sed "s/^/SMALL_GEMM=1 /" | tee -a $OUT/staged_summary.txt
CHOL_SMALL_GEMM=1 CHOL_POTF2_V2=1 timeout 300 python scripts/configs_bench.py cfg1 2>&1 | grep -E "^dpotrf" | sed "s/^/SMALL_GEMM=1 POTF2_V2=1 /" | tee -a $OUT/staged_summary.txt
