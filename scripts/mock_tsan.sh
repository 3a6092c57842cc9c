#!/bin/bash
# The multi-threaded host runtime (one worker thread per GPU, NCCL rendezvous, the stream schedule) on the mock back end
# (tests/mock) under ThreadSanitizer, driven by the plain-C callers of tests/c so that no Python runtime is in the
# process.  The mock's streams are real threads synchronised with pthread primitives, so a host-side access to a
# buffer that a stream has not finished with, or two workers touching shared library state unsynchronised, is
# reported.  No GPU needed.   bash scripts/mock_tsan.sh [devices...]   (default: 1 3 5)
set -eu
cd "$(dirname "$0")/.."
D=$(mktemp -d)
for f in cuda-cholesky_b200/host/*.c tests/mock/mock_cuda.c; do
  gcc -std=c99 -O1 -g -fsanitize=thread -fPIC -D_POSIX_C_SOURCE=200809L -Iinclude -I/usr/local/cuda/include -c $f -o $D/$(basename $f .c).o
done
g++ -std=c++17 -O1 -g -fsanitize=thread -fPIC -c tests/mock/mock_kernels.cpp -o $D/mock_kernels.o
g++ -shared -fsanitize=thread -o $D/libcholesky_b200_mock.so $D/*.o -lpthread -ldl -lm
gcc -std=c99 -O1 -g -fsanitize=thread -fPIC -shared tests/mock/mock_nccl.c -o $D/libnccl.so.2 -lpthread
for prog in cumultigpu_unit mgpu_potri_chain; do
  gcc -std=c99 -O1 -g -fsanitize=thread -Iinclude -I/usr/local/cuda/include tests/c/$prog.c -o $D/$prog -L$D -lcholesky_b200_mock -lm -Wl,-rpath,$D
done
for P in ${@:-1 3 5}; do
  for prog in cumultigpu_unit mgpu_potri_chain; do
    for jit in 0 1; do
      echo "== $prog, $P mock device(s), stream jitter $jit"
      MOCK_CUDA_JITTER=$jit MOCK_CUDA_DEVICES=$P LD_LIBRARY_PATH=$D TSAN_OPTIONS="halt_on_error=0" $D/$prog 2>&1 |
        grep -E "ThreadSanitizer|SUMMARY|PASSED|FAILED" | sort | uniq -c | head -8
    done
  done
done
rm -rf $D
