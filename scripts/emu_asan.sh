#!/bin/bash
# The kernel emulation harnesses (tests/emu) under AddressSanitizer: global operands are heap buffers and shared
# memory is ordinary memory there, so a kernel that reads or writes past a matrix edge, a workspace or its shared
# tiles is reported.  No GPU needed; a few minutes.   bash scripts/emu_asan.sh [harness...]
set -eu
cd "$(dirname "$0")/.."
D=$(mktemp -d)
for h in ${@:-emu_small_gemm emu_panel emu_ffma emu_dmma}; do
  g++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -Itests/emu -I/usr/local/cuda/include -Icuda-cholesky_b200/csrc -Iinclude \
    tests/emu/$h.cpp -o $D/$h -lpthread
  echo "== $h"
  ASAN_OPTIONS=detect_leaks=0 EMU_WATCHDOG=2300 $D/$h 2>&1 | grep -E "AddressSanitizer|EMU PASSED|EMU FAILED|MISMATCH" | head -5
done
rm -rf $D
