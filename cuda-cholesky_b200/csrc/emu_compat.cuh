// emu_compat.cuh — the one place where kernel sources differ between nvcc and the CPU emulation of
// tests/emu (tests/test_kernel_emulation.py).  Under nvcc every macro expands to exactly the tokens
// the kernels used to spell out, so the generated SASS is unchanged (scripts/sass_compare.py).
#pragma once
#ifndef CHOL_CPU_EMU
#define CHOL_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#else
#define CHOL_DYN_SMEM(name) unsigned char *name = emu::dyn_smem()
#endif
