#!/bin/bash
# The host drivers on the mock back end (tests/mock) under AddressSanitizer: "device" memory is heap memory
# there, so any launch whose pointer / leading-dimension arithmetic leaves its workspace or the caller's matrix is
# reported.  No GPU needed.   bash scripts/mock_asan.sh [groups...]   (default: lapack blas multi1 multi3 dist3)
set -eu
cd "$(dirname "$0")/.."
D=$(mktemp -d)
for f in cuda-cholesky_b200/host/*.c tests/mock/mock_cuda.c; do
  gcc -std=c99 -O1 -g -fsanitize=address -fno-omit-frame-pointer -fPIC -D_POSIX_C_SOURCE=200809L -Iinclude -I/usr/local/cuda/include -c $f -o $D/$(basename $f .c).o
done
g++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -fPIC -c tests/mock/mock_kernels.cpp -o $D/mock_kernels.o
g++ -shared -fsanitize=address -o $D/libcholesky_b200_mock.so $D/*.o -lpthread -ldl -lm
gcc -std=c99 -O1 -g -fsanitize=address -fPIC -shared tests/mock/mock_nccl.c -o $D/libnccl.so.2 -lpthread
ASAN=$(gcc -print-file-name=libasan.so)
for g in ${@:-lapack blas multi1 multi3 dist3}; do
  for sw in "" "CHOL_PANEL_PIPELINE=0 CHOL_LOOKAHEAD_SPLIT=0 CHOL_TRTRI_LOOKAHEAD=0"; do
    echo "== $g ${sw:-(defaults)}"
    env $sw ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 LD_PRELOAD=$ASAN LD_LIBRARY_PATH=$D OMP_NUM_THREADS=2 \
      python tests/mock/run_host_driver_checks.py $D/libcholesky_b200_mock.so $g 2>&1 | grep -E "AddressSanitizer|MOCK PASSED|FAILED|checks" | head -5
  done
done
rm -rf $D
