/*
 * cumultigpu.h — task objects and the multi-GPU context.
 * Same signatures as reference include/cumultigpu.h:11-96 (CUtask: create/destroy/execute;
 * CUmultiGPU: create/destroy/count/runTask/synchronize).  The implementation behind them is new:
 * the reference's pthread-per-GPU ring-buffer scheduler that stages every tile over PCIe
 * (multigpu/multigpu.c:13-538) is replaced by one worker thread per GPU that keeps the matrix
 * RESIDENT in HBM in a 1-D block-cyclic layout and exchanges factored panels GPU-to-GPU over
 * NVLink (cuda-cholesky_b200/host/multigpu.c).
 */
#ifndef CHOL_B200_CUMULTIGPU_H
#define CHOL_B200_CUMULTIGPU_H

#include <stddef.h>
#include <cuda.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct __cutask_st *CUtask;

/* CUDA_ERROR_INVALID_VALUE if function is NULL or (args == NULL && size > 0);
 * CUDA_ERROR_OUT_OF_MEMORY if the argument copy cannot be allocated (multigpu.c:321-353). */
CUresult cuTaskCreate(CUtask *, CUresult (*)(const void *), const void *, size_t);
/* Blocks until the task has run, returns its result through *result (may be NULL) and frees it
 * (multigpu.c:355-375). */
CUresult cuTaskDestroy(CUtask, CUresult *);
/* Runs the task on the calling thread (multigpu.c:377-400). */
CUresult cuTaskExecute(CUtask);

typedef struct __cumultigpu_st *CUmultiGPU;

/* One context (the device's primary context) and one worker thread per listed device. */
CUresult cuMultiGPUCreate(CUmultiGPU *, CUdevice *, int);
CUresult cuMultiGPUDestroy(CUmultiGPU);
int cuMultiGPUGetContextCount(CUmultiGPU);
/* Queues the task on worker i with that GPU's context current;
 * CUDA_ERROR_INVALID_VALUE if i is out of range (multigpu.c:488-496). */
CUresult cuMultiGPURunTask(CUmultiGPU, int, CUtask);
/* Waits until every worker has drained its queue and its context is idle. */
CUresult cuMultiGPUSynchronize(CUmultiGPU);

#ifdef __cplusplus
}
#endif
#endif
