/*
 * rng_internal.h — private interface between host/rng.c (the cuRng64* entry points) and host/rng_poly.c (GF(2)
 * polynomial arithmetic for the jump-ahead of the device dSFMT-19937).  No CUDA types here.
 */
#ifndef CHOL_RNG_INTERNAL_H
#define CHOL_RNG_INTERNAL_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#define CHOL_DSFMT_N 191                      /* 128-bit words of the state: (19937 - 128) / 104 + 1, rng/dsfmt.c:4 */
#define CHOL_DSFMT_WORDS (CHOL_DSFMT_N + 1)   /* ... plus the lung */
#define CHOL_DSFMT_STATE_BYTES (CHOL_DSFMT_WORDS * 16)

/* the state the reference's `set` leaves (rng/dsfmt.c:37-90); word i = st[i][0] (low 64 bits), st[i][1] (high) */
void chol_dsfmt_seed_state(uint64_t seed, uint64_t st[CHOL_DSFMT_WORDS][2]);
/* one refill (191 steps, rng/dsfmt.c:92-125) on the host - used to find and verify the polynomial, never for output */
void chol_dsfmt_refill_host(uint64_t st[CHOL_DSFMT_WORDS][2]);

/* make sure the process-wide polynomial p annihilates s0 (p(T) s0 = 0, checked exactly); computes or extends it if
 * needed.  Returns 0 and D = deg p, generation = a counter that changes whenever p does; -1 on failure (out of
 * memory, or no polynomial within the dimension bound - which would mean the generator is not the one assumed). */
int chol_dsfmt_poly_ensure(const uint64_t s0[CHOL_DSFMT_WORDS][2], int *D, unsigned *generation);
/* rows x^((j0 + i) * 2^level) mod p, i < count, as little-endian 32-bit words (stride_u32 >= 2 * ceil(D / 64)),
 * padded with zeros.  level >= 3.  -1 if `generation` is stale or on allocation failure. */
int chol_dsfmt_poly_rows(unsigned generation, int level, uint64_t j0, size_t count, uint32_t *dst, size_t stride_u32);
void chol_dsfmt_poly_reset(void);

#endif
