/*
 * dsfmt_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * Portable (no SSE) restatement of the dSFMT-19937 generator that BASELINE.json names as the
 * seed of the synthetic SPD inputs ("A = B*B^T + N*I, dSFMT seed").
 * Follows reference rng/dsfmt.c:37-90 (`set`: LCG seeding with 1812433253, mask into [1,2),
 * period certification), rng/dsfmt.c:92-125 (`generate`: the lung recursion with SL1=19,
 * SR=12, POS1=117) and the constants of rng/dsfmt_19937.c:1-15.  getCloseOpen (value - 1.0,
 * rng/dsfmt.c:219-252) gives doubles in [0,1).
 * Pinned by tests/golden/dsfmt19937_seed20121018.json, produced by oracle/pin_dsfmt.c which
 * compiles the reference's own rng/dsfmt_19937.c (tests/golden/make_golden.py).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define DS_MEXP 19937
#define DS_N ((DS_MEXP - 128) / 104 + 1) /* 191 128-bit words + 1 lung */
#define DS_POS1 117
#define DS_SL1 19
#define DS_SR 12
#define DS_MSK1 UINT64_C(0x000ffafffffffb3f)
#define DS_MSK2 UINT64_C(0x000ffdfffc90fffd)
#define DS_FIX1 UINT64_C(0x90014964b32f4329)
#define DS_FIX2 UINT64_C(0x3b8d12ac548a7c7a)
#define DS_PCV1 UINT64_C(0x3d84e1ac0dc82880)
#define DS_PCV2 UINT64_C(0x0000000000000001)
#define DS_LOW UINT64_C(0x000FFFFFFFFFFFFF)
#define DS_HIGH UINT64_C(0x3FF0000000000000)

typedef struct {
  uint64_t s[DS_N + 1][2];
  size_t idx; /* next 128-bit word to hand out, DS_N = regenerate */
} oracle_dsfmt_t;

size_t oracle_dsfmt_state_size(void) { return sizeof(oracle_dsfmt_t); }

void oracle_dsfmt_seed(oracle_dsfmt_t *g, uint64_t seed) {
  uint32_t w[(DS_N + 1) * 4];
  w[0] = (uint32_t)seed;
  for (size_t i = 1; i < (DS_N + 1) * 4; i++)
    w[i] = UINT32_C(1812433253) * (w[i - 1] ^ (w[i - 1] >> 30)) + (uint32_t)i;
  memcpy(g->s, w, sizeof w); /* little-endian layout, as the reference's union */
  for (size_t i = 0; i < DS_N; i++)
    for (int h = 0; h < 2; h++) g->s[i][h] = (g->s[i][h] & DS_LOW) | DS_HIGH;
  uint64_t inner = ((g->s[DS_N][0] ^ DS_FIX1) & DS_PCV1) ^ ((g->s[DS_N][1] ^ DS_FIX2) & DS_PCV2);
  for (int sh = 32; sh > 0; sh >>= 1) inner ^= inner >> sh;
  if ((inner & 1) == 0) g->s[DS_N][1] ^= 1; /* PCV2 is odd: flip the lowest bit of the lung */
  g->idx = DS_N;
}

static void ds_step(uint64_t r[2], const uint64_t b[2], uint64_t lung[2]) {
  const uint64_t t0 = r[0], t1 = r[1], L0 = lung[0], L1 = lung[1];
  lung[0] = (t0 << DS_SL1) ^ (L1 >> 32) ^ (L1 << 32) ^ b[0];
  lung[1] = (t1 << DS_SL1) ^ (L0 >> 32) ^ (L0 << 32) ^ b[1];
  r[0] = (lung[0] >> DS_SR) ^ (lung[0] & DS_MSK1) ^ t0;
  r[1] = (lung[1] >> DS_SR) ^ (lung[1] & DS_MSK2) ^ t1;
}

static void ds_refill(oracle_dsfmt_t *g) {
  uint64_t lung[2] = {g->s[DS_N][0], g->s[DS_N][1]};
  for (size_t i = 0; i < DS_N; i++) ds_step(g->s[i], g->s[(i + DS_POS1) % DS_N], lung);
  g->s[DS_N][0] = lung[0];
  g->s[DS_N][1] = lung[1];
  g->idx = 0;
}

/* n doubles in [0,1): pairs are taken from one 128-bit word; an odd tail uses the low half and
 * discards the high half, like the reference's getCloseOpen. */
void oracle_dsfmt_fill_close_open(oracle_dsfmt_t *g, double *out, size_t n) {
  size_t i = 0;
  while (i < n) {
    if (g->idx >= DS_N) ds_refill(g);
    const uint64_t *wrd = g->s[g->idx++];
    double d[2];
    memcpy(d, wrd, sizeof d);
    out[i++] = d[0] - 1.0;
    if (i < n) out[i++] = d[1] - 1.0;
  }
}

/* The reference's five output forms (rng/dsfmt.c), as raw 64-bit patterns so that they can be compared exactly:
 *   kind 0  the 64-bit halves themselves (`get`, rng/dsfmt.c:127-146 - there the last pair of an even request is
 *           lost to a `while (n > 2)`; the restatement hands out all n)
 *   kind 1  getCloseOpen  :231  d - 1.0
 *   kind 2  getOpenClose  :196  2.0 - d
 *   kind 3  getOpenOpen   :160-161  (bits | 1) - 1.0
 *   kind 4  getCloseClose :266  (double)(bits >> 10) * (1.0 / 9007199254740991.0)   (shift by 10 as written there)
 * Pairs come from one 128-bit word; an odd tail uses the low half and discards the high half. */
static uint64_t ds_convert(uint64_t w, int kind) {
  double d, r;
  memcpy(&d, &w, sizeof d);
  switch (kind) {
    case 0: return w;
    case 1: r = d - 1.0; break;
    case 2: r = 2.0 - d; break;
    case 3: {
      uint64_t v = w | 1u;
      memcpy(&d, &v, sizeof d);
      r = d - 1.0;
      break;
    }
    default: {
      volatile double p = (double)(w >> 10) * (1.0 / 9007199254740991.0);
      r = p;
    }
  }
  uint64_t o;
  memcpy(&o, &r, sizeof o);
  return o;
}

void oracle_dsfmt_fill(oracle_dsfmt_t *g, uint64_t *out, size_t n, int kind) {
  size_t i = 0;
  while (i < n) {
    if (g->idx >= DS_N) ds_refill(g);
    const uint64_t *wrd = g->s[g->idx++];
    out[i++] = ds_convert(wrd[0], kind);
    if (i < n) out[i++] = ds_convert(wrd[1], kind);
  }
}
