/*
 * rng.h — device pseudo-random number generator: the CUrng64 part of the reference's include/rng.h:326-413
 * (same type and function names, same argument meaning), which the reference declares but never implemented
 * (rng/curng64.c is empty).  SURVEY.md §8(f).4: the input generator of the BASELINE configs on the device.
 *
 * One generator type is offered:
 *
 *   cu_dsfmt_19937_t   dSFMT-19937 whose output is, bit for bit, the stream the reference's CPU generator
 *                      `dsfmt_19937_t` (rng/dsfmt_19937.c + rng/dsfmt.c) yields for the same seed and the same
 *                      sequence of calls - produced in parallel on the device by jumping ahead in the linear
 *                      recurrence (csrc/rng_dsfmt.cu), so a matrix filled on the GPU equals the one the CPU
 *                      harness fills (tests compare them exactly).
 *
 * Not offered: `mtgp64_11213_t` / `mtgp32_11213_t` (include/rng.h:324,413) - their 11 600-line parameter tables
 * (rng/mtgp64_11213.c) are third-party data that defines the generator, and the reference never ran them; the CPU
 * generators (rng32 / rng64) are the oracle's business, not the product's.
 *
 * Semantics kept from the CPU generator the stream is defined by: values are taken in pairs from one 128-bit word
 * of the state, and a request for an odd count drops the high half of its last word (rng/dsfmt.c:245-251), so
 * n1 values followed by n2 values equal n1 + n2 values in one call only when n1 is even.
 * The four floating-point conversions are the reference's arithmetic (rng/dsfmt.c:148-285) literally - including
 * getCloseClose's shift by 10 (its comment says 12), which makes that one return values in [0.4995, 0.5].
 * cuRng64Get returns the raw 64-bit halves (bit patterns of doubles in [1, 2)), all n of them in order (the reference's
 * CPU `get`, rng/dsfmt.c:127-146, stores whole words while more than two values remain and then writes the low half of
 * one more word to the wrong index; that slip is not reproduced - the values it does store in place are the same).
 *
 * Calls are asynchronous on the generator's stream (default: the legacy default stream 0); a generator is not
 * thread-safe; different generators may be used from different threads.
 */
#ifndef CHOL_B200_RNG_H
#define CHOL_B200_RNG_H

#include <stddef.h>
#include <stdint.h>
#include <cuda.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference include/rng.h:330,335 */
typedef struct __curng64_type_st *CUrng64_t;
typedef struct __curng64_st *CUrng64;

/* reference include/rng.h:346.  Adopts the calling thread's current context (device 0's primary context if
 * there is none), like cuBLASCreate.  CUDA_ERROR_INVALID_VALUE for an unknown type. */
CUresult cuRng64Create(CUrng64 *, const CUrng64_t);
/* reference include/rng.h:353 */
CUresult cuRng64Destroy(CUrng64);
/* reference include/rng.h:361: seed and rewind.  Builds the generator's jump table on the device (61 MB). */
CUresult cuRng64Set(const CUrng64, uint64_t seed);
/* reference include/rng.h:370: n raw 64-bit values to device memory (8-byte aligned) */
CUresult cuRng64Get(const CUrng64, CUdeviceptr, size_t n);
/* reference include/rng.h:380-410: n doubles in (0,1), (0,1], [0,1) and "[0,1]" (see above) to device memory */
CUresult cuRng64GetOpenOpen(const CUrng64, CUdeviceptr, size_t n);
CUresult cuRng64GetOpenClose(const CUrng64, CUdeviceptr, size_t n);
CUresult cuRng64GetCloseOpen(const CUrng64, CUdeviceptr, size_t n);
CUresult cuRng64GetCloseClose(const CUrng64, CUdeviceptr, size_t n);

extern const CUrng64_t cu_dsfmt_19937_t;

/* ---- additions (not in the reference) ---- */
/* later calls run on `stream`; waits for the work already queued on the previous one */
CUresult cuRng64SetStream(CUrng64, CUstream stream);
/* advance the stream as a Get of n values would, without producing them (a host counter: any rank of a multi-GPU
 * job can produce its own window of one global stream) */
CUresult cuRng64Discard(CUrng64, size_t n);
/* 128-bit words of the stream consumed since cuRng64Set */
unsigned long long cuRng64Position(CUrng64);
/* tuning / measurement hooks (like cuLAPACKSetBlockSize / cuLAPACKProfile): a request is cut into segments of 2^level
 * refills (382 values each); bias (-8 .. 8) shifts the level the library would choose - the values produced do not
 * depend on it.  Profile(1) times the two kernels of every later Get with events on the generator's stream; ProfileRead
 * waits for the last Get and returns its jump and generation times, segment count and segment length (any pointer
 * may be NULL); CUDA_ERROR_NOT_READY if no Get has run since Profile(1). */
CUresult cuRng64SetSegmentBias(CUrng64, int bias);
CUresult cuRng64Profile(CUrng64, int enable);
CUresult cuRng64ProfileRead(CUrng64, float *jump_ms, float *generate_ms, int *segments, int *segment_refills);

#ifdef __cplusplus
}
#endif
#endif
