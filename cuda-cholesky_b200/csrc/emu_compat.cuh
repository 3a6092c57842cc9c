// emu_compat.cuh — the one place where kernel sources differ between nvcc and the CPU emulation of
// tests/emu (tests/test_kernel_emulation.py).  Under nvcc every macro expands to exactly the tokens
// the kernels used to spell out, so the generated SASS is unchanged (scripts/sass_compare.py).
#pragma once
#ifndef CHOL_CPU_EMU
#define CHOL_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define CHOL_CP_ASYNC_COMMIT() asm volatile("cp.async.commit_group;" ::: "memory")
#define CHOL_CP_ASYNC_WAIT(n) asm volatile("cp.async.wait_group %0;" ::"n"(n) : "memory")
#define CHOL_DYN_SMEM128(name) extern __shared__ __align__(128) unsigned char name[]
#define CHOL_SETMAXNREG_DEC(n) asm volatile("setmaxnreg.dec.sync.aligned.u32 " #n ";")
#define CHOL_SETMAXNREG_INC(n) asm volatile("setmaxnreg.inc.sync.aligned.u32 " #n ";")
#else
#define CHOL_DYN_SMEM(name) unsigned char *name = emu::dyn_smem()
#define CHOL_CP_ASYNC_COMMIT() emu::cp_async_commit()   /* the emulation completes a group as LATE as wait_group allows */
#define CHOL_CP_ASYNC_WAIT(n) emu::cp_async_wait(n)
#define CHOL_DYN_SMEM128(name) unsigned char *name = emu::dyn_smem()
#define CHOL_SETMAXNREG_DEC(n) ((void)0)
#define CHOL_SETMAXNREG_INC(n) ((void)0)
#endif
