/*
 * rng_poly.c — the GF(2)-polynomial side of the device dSFMT-19937 generator (host/rng.c, csrc/rng_dsfmt.cu).
 * Plain C, no CUDA: seeding (reference rng/dsfmt.c:37-90), one refill T of the state (rng/dsfmt.c:92-125), the
 * minimal polynomial p of T on a seed state (Berlekamp-Massey), and the residues a_r = x^r mod p on a grid of refill
 * counts r, which is what the jump kernel needs: T^r s0 = sum_i a_r[i] T^i s0.
 *
 * Nothing here is on the data path: the refill below runs only to find / verify p (once per process and, as a
 * check, once per seed); every output value is produced by the device kernels.
 *
 * p is kept per process.  Every seed is checked against it exactly (p(T) s0 must be the zero state) and p is multiplied
 * by the missing factor if it is not - this happens: the state of seed 20121018 has a minimal polynomial of degree 19991,
 * that of seed 12345 one of degree 19993 = 191 * 104 + 128 + 1, the dimension bound (the top 12 bits of the 382 doubles
 * never change), which then serves both.  `generation` tells the callers that p changed (their Krylov blocks are stale).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "rng_internal.h"

#define DS_POS1 117
#define DS_SL1 19
#define DS_SR 12
#define DS_MSK1 UINT64_C(0x000ffafffffffb3f) /* rng/dsfmt_19937.c:5-6 */
#define DS_MSK2 UINT64_C(0x000ffdfffc90fffd)
#define DS_FIX1 UINT64_C(0x90014964b32f4329) /* rng/dsfmt_19937.c:11-14 */
#define DS_FIX2 UINT64_C(0x3b8d12ac548a7c7a)
#define DS_PCV1 UINT64_C(0x3d84e1ac0dc82880)
#define DS_PCV2 UINT64_C(0x0000000000000001)
#define DS_LOW_MASK UINT64_C(0x000FFFFFFFFFFFFF) /* rng/dsfmt.c:5-6 */
#define DS_HIGH_CONST UINT64_C(0x3FF0000000000000)

/* the states T^k s0 span at most 191 * 104 + 128 + 1 dimensions (the top 12 bits of the 382 doubles never change) */
#define DS_DEG_BOUND 20000
#define CK_STRIDE 65536u /* refills between two cached checkpoints a_(m * CK_STRIDE) */
#define MAX_LEVEL 40

/* rng/dsfmt.c:37-90 (`set`): 32-bit LCG fill, mask into [1,2), period certification */
void chol_dsfmt_seed_state(uint64_t seed, uint64_t st[CHOL_DSFMT_WORDS][2]) {
  uint32_t w[CHOL_DSFMT_WORDS * 4];
  w[0] = (uint32_t)seed;
  for (size_t i = 1; i < CHOL_DSFMT_WORDS * 4; i++) w[i] = UINT32_C(1812433253) * (w[i - 1] ^ (w[i - 1] >> 30)) + (uint32_t)i;
  for (size_t i = 0; i < CHOL_DSFMT_WORDS; i++) { /* little-endian 32-bit pairs, as the reference's union lays them out */
    st[i][0] = (uint64_t)w[4 * i] | ((uint64_t)w[4 * i + 1] << 32);
    st[i][1] = (uint64_t)w[4 * i + 2] | ((uint64_t)w[4 * i + 3] << 32);
  }
  for (size_t i = 0; i < CHOL_DSFMT_N; i++)
    for (int h = 0; h < 2; h++) st[i][h] = (st[i][h] & DS_LOW_MASK) | DS_HIGH_CONST;
  uint64_t inner = ((st[CHOL_DSFMT_N][0] ^ DS_FIX1) & DS_PCV1) ^ ((st[CHOL_DSFMT_N][1] ^ DS_FIX2) & DS_PCV2);
  for (int sh = 32; sh > 0; sh >>= 1) inner ^= inner >> sh;
  if ((inner & 1) == 0) st[CHOL_DSFMT_N][1] ^= 1; /* PCV2 is odd */
}

/* one refill: the 191 steps of rng/dsfmt.c:92-125 */
void chol_dsfmt_refill_host(uint64_t st[CHOL_DSFMT_WORDS][2]) {
  uint64_t l0 = st[CHOL_DSFMT_N][0], l1 = st[CHOL_DSFMT_N][1];
  for (int i = 0; i < CHOL_DSFMT_N; i++) {
    const int b = i + DS_POS1 < CHOL_DSFMT_N ? i + DS_POS1 : i + DS_POS1 - CHOL_DSFMT_N;
    const uint64_t t0 = st[i][0], t1 = st[i][1];
    const uint64_t n0 = (t0 << DS_SL1) ^ (l1 >> 32) ^ (l1 << 32) ^ st[b][0];
    const uint64_t n1 = (t1 << DS_SL1) ^ (l0 >> 32) ^ (l0 << 32) ^ st[b][1];
    l0 = n0;
    l1 = n1;
    st[i][0] = (l0 >> DS_SR) ^ (l0 & DS_MSK1) ^ t0;
    st[i][1] = (l1 >> DS_SR) ^ (l1 & DS_MSK2) ^ t1;
  }
  st[CHOL_DSFMT_N][0] = l0;
  st[CHOL_DSFMT_N][1] = l1;
}

/* ---------------------------------------------------------------------------------------------------------------
 * bit vectors
 * ------------------------------------------------------------------------------------------------------------- */
static inline size_t chol_min_sz(size_t a, size_t b) { return a < b ? a : b; }
static inline int bit_get(const uint64_t *v, size_t i) { return (int)((v[i >> 6] >> (i & 63)) & 1); }
static inline void bit_set(uint64_t *v, size_t i) { v[i >> 6] |= UINT64_C(1) << (i & 63); }

/* dst ^= src << sh over nw words (bits shifted beyond nw words are dropped) */
static void xor_shifted(uint64_t *dst, const uint64_t *src, size_t sh, size_t nw) {
  const size_t ws = sh >> 6, bs = sh & 63;
  if (ws >= nw) return;
  if (bs == 0) {
    for (size_t w = ws; w < nw; w++) dst[w] ^= src[w - ws];
  } else {
    dst[ws] ^= src[0] << bs;
    for (size_t w = ws + 1; w < nw; w++) dst[w] ^= (src[w - ws] << bs) | (src[w - ws - 1] >> (64 - bs));
  }
}

static int poly_degree(const uint64_t *p, size_t nw) {
  for (size_t w = nw; w-- > 0;)
    if (p[w]) return (int)(w * 64 + 63 - (size_t)__builtin_clzll(p[w]));
  return -1;
}

/* Berlekamp-Massey over GF(2): the shortest linear recurrence of seq[0..n).  Returns its length L and the
 * characteristic polynomial (degree L, sum_j p_j s_(k+j) = 0 for all k) in p (nw words, nw * 64 > n / 2 + 64). */
static int berlekamp_massey(const uint64_t *seq, size_t n, uint64_t *p, size_t nw) {
  uint64_t *C = calloc(3 * nw + 1, sizeof *C), *B, *Tm, *W;
  uint64_t *Wbuf = calloc(nw, sizeof *Wbuf);
  if (C == NULL || Wbuf == NULL) {
    free(C);
    free(Wbuf);
    return -1;
  }
  B = C + nw;
  Tm = B + nw;
  W = Wbuf; /* bit i = s_(k-i): the sequence read backwards from the current position */
  C[0] = B[0] = 1;
  size_t L = 0, m = 1;
  for (size_t k = 0; k < n; k++) {
    /* W <<= 1, insert s_k */
    const size_t wl = chol_min_sz(nw, (k >> 6) + 1);
    for (size_t w = wl; w-- > 1;) W[w] = (W[w] << 1) | (W[w - 1] >> 63);
    W[0] = (W[0] << 1) | (uint64_t)bit_get(seq, k);
    uint64_t acc = 0;
    const size_t lw = (L >> 6) + 1;
    for (size_t w = 0; w < lw; w++) acc ^= C[w] & W[w];
    if (__builtin_parityll(acc)) {
      if (2 * L <= k) {
        memcpy(Tm, C, nw * sizeof *C);
        xor_shifted(C, B, m, nw);
        memcpy(B, Tm, nw * sizeof *C);
        L = k + 1 - L;
        m = 1;
      } else {
        xor_shifted(C, B, m, nw);
        m++;
      }
    } else {
      m++;
    }
  }
  /* connection polynomial C (sum_i C_i s_(k-i) = 0) -> characteristic polynomial p_j = C_(L-j) */
  memset(p, 0, nw * sizeof *p);
  for (size_t j = 0; j <= L; j++)
    if (bit_get(C, L - j)) bit_set(p, j);
  free(C);
  free(Wbuf);
  return (int)L;
}

/* ---------------------------------------------------------------------------------------------------------------
 * the per-process polynomial context
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct level_cache {
  uint64_t j0;
  size_t count;
  uint64_t *rows;
} level_cache;

static struct {
  pthread_mutex_t lock;
  int D;            /* deg p, 0 = not computed yet */
  size_t pw;        /* 64-bit words of a residue */
  uint64_t *p;      /* D + 1 bits in pw + 1 words */
  uint64_t *tab;    /* 256 x pw: (v(x) * x^D) mod p for the 8-bit polynomials v */
  unsigned generation;
  uint64_t *ck;     /* checkpoints a_(m * CK_STRIDE), m < ck_count */
  size_t ck_count, ck_cap;
  level_cache level[MAX_LEVEL];
} G = {.lock = PTHREAD_MUTEX_INITIALIZER};

static void state_xor(uint64_t a[CHOL_DSFMT_WORDS][2], uint64_t b[CHOL_DSFMT_WORDS][2]) {
  for (int i = 0; i < CHOL_DSFMT_WORDS; i++) {
    a[i][0] ^= b[i][0];
    a[i][1] ^= b[i][1];
  }
}
static bool state_is_zero(uint64_t a[CHOL_DSFMT_WORDS][2]) {
  uint64_t o = 0;
  for (int i = 0; i < CHOL_DSFMT_WORDS; i++) o |= a[i][0] | a[i][1];
  return o == 0;
}

/* v <- q(T) v */
static void apply_poly(const uint64_t *q, int deg, uint64_t v[CHOL_DSFMT_WORDS][2]) {
  uint64_t acc[CHOL_DSFMT_WORDS][2];
  memset(acc, 0, sizeof acc);
  for (int j = 0; j <= deg; j++) {
    if (bit_get(q, (size_t)j)) state_xor(acc, v);
    if (j < deg) chol_dsfmt_refill_host(v);
  }
  memcpy(v, acc, sizeof acc);
}

/* a fixed pseudo-random choice of 64 state bits per functional (splitmix64) */
static uint64_t splitmix(uint64_t *x) {
  uint64_t z = (*x += UINT64_C(0x9e3779b97f4a7c15));
  z = (z ^ (z >> 30)) * UINT64_C(0xbf58476d1ce4e5b9);
  z = (z ^ (z >> 27)) * UINT64_C(0x94d049bb133111eb);
  return z ^ (z >> 31);
}

static void free_caches(void) {
  for (int l = 0; l < MAX_LEVEL; l++) {
    free(G.level[l].rows);
    memset(&G.level[l], 0, sizeof G.level[l]);
  }
  free(G.ck);
  G.ck = NULL;
  G.ck_count = G.ck_cap = 0;
}

static int build_table(void) {
  const size_t pw = G.pw;
  const int D = G.D;
  free(G.tab);
  G.tab = calloc(256 * pw, sizeof *G.tab);
  uint64_t *xp = calloc(8 * pw, sizeof *xp);
  if (G.tab == NULL || xp == NULL) {
    free(xp);
    return -1;
  }
  /* xp[b] = x^(D+b) mod p */
  for (size_t w = 0; w < pw; w++) xp[w] = G.p[w];
  if (D & 63) xp[(size_t)D >> 6] &= ~(UINT64_C(1) << (D & 63)); /* drop the leading term (D % 64 == 0: it lives in word pw, not copied) */
  for (int b = 1; b < 8; b++) {
    uint64_t *cur = xp + (size_t)b * pw, *prev = xp + (size_t)(b - 1) * pw;
    const int carry = bit_get(prev, (size_t)D - 1);
    for (size_t w = pw; w-- > 1;) cur[w] = (prev[w] << 1) | (prev[w - 1] >> 63);
    cur[0] = prev[0] << 1;
    if (D & 63) cur[pw - 1] &= (UINT64_C(1) << (D & 63)) - 1;
    if (carry)
      for (size_t w = 0; w < pw; w++) cur[w] ^= xp[w];
  }
  for (int v = 1; v < 256; v++) {
    uint64_t *t = G.tab + (size_t)v * pw;
    for (int b = 0; b < 8; b++)
      if ((v >> b) & 1)
        for (size_t w = 0; w < pw; w++) t[w] ^= xp[(size_t)b * pw + w];
  }
  free(xp);
  return 0;
}

/* a <- a * x^8 mod p */
static void step8(uint64_t *a) {
  const size_t pw = G.pw;
  const int D = G.D;
  unsigned v = 0;
  for (int b = 0; b < 8; b++) v |= (unsigned)bit_get(a, (size_t)(D - 8 + b)) << b;
  for (size_t w = pw; w-- > 1;) a[w] = (a[w] << 8) | (a[w - 1] >> 56);
  a[0] <<= 8;
  if (D & 63) a[pw - 1] &= (UINT64_C(1) << (D & 63)) - 1;
  if (v) {
    const uint64_t *t = G.tab + (size_t)v * pw;
    for (size_t w = 0; w < pw; w++) a[w] ^= t[w];
  }
}

/* make p annihilate s0: p <- p * (minimal polynomial of p(T) s0), repeated until p(T) s0 = 0 */
static int refine_poly(uint64_t s0[CHOL_DSFMT_WORDS][2]) {
  const size_t nw = (DS_DEG_BOUND + 64) / 64 + 2;
  uint64_t v[CHOL_DSFMT_WORDS][2], u[CHOL_DSFMT_WORDS][2];
  uint64_t *p = calloc(nw, sizeof *p), *q = calloc(nw, sizeof *q), *r = calloc(nw, sizeof *r);
  uint64_t *seq = calloc((2 * DS_DEG_BOUND + 256) / 64 + 2, sizeof *seq);
  int rc = -1;
  if (p == NULL || q == NULL || r == NULL || seq == NULL) goto done;
  int dp = 0;
  if (G.D > 0) {
    memcpy(p, G.p, (G.pw + 1) * sizeof *p); /* G.pw + 1 <= nw because G.D <= DS_DEG_BOUND */
    dp = G.D;
  } else {
    p[0] = 1;
  }
  memcpy(v, s0, sizeof v);
  apply_poly(p, dp, v);
  uint64_t mix = UINT64_C(20121018);
  for (int round = 0; round < 64 && !state_is_zero(v); round++) {
    const int rem = DS_DEG_BOUND - dp;
    if (rem <= 0) goto done;
    const size_t n = 2 * (size_t)rem + 64;
    /* functional: parity of 64 pseudo-randomly chosen bits of the state (round 0: bit 0 alone also works; any works) */
    size_t pos[64];
    for (int k = 0; k < 64; k++) pos[k] = (size_t)(splitmix(&mix) % (CHOL_DSFMT_WORDS * 128));
    memset(seq, 0, ((n + 63) / 64 + 1) * sizeof *seq);
    memcpy(u, v, sizeof u);
    const uint64_t *flat = &u[0][0];
    for (size_t k = 0; k < n; k++) {
      int par = 0;
      for (int b = 0; b < 64; b++) par ^= bit_get(flat, pos[b]);
      if (par) bit_set(seq, k);
      chol_dsfmt_refill_host(u);
    }
    const int dq = berlekamp_massey(seq, n, q, nw);
    if (dq < 0) goto done;
    if (dq == 0) continue; /* this functional does not see what is left of v */
    if (dp + dq > DS_DEG_BOUND) goto done;
    /* p <- p * q, v <- q(T) v */
    memset(r, 0, nw * sizeof *r);
    const uint64_t *small = dq < dp ? q : p, *big = dq < dp ? p : q;
    const int ds = dq < dp ? dq : dp;
    for (int j = 0; j <= ds; j++)
      if (bit_get(small, (size_t)j)) xor_shifted(r, big, (size_t)j, nw);
    memcpy(p, r, nw * sizeof *p);
    dp += dq;
    apply_poly(q, dq, v);
  }
  if (!state_is_zero(v) || poly_degree(p, nw) != dp || dp < 16) goto done;
  /* install */
  free(G.p);
  G.D = dp;
  G.pw = ((size_t)dp + 63) / 64;
  G.p = calloc(G.pw + 1, sizeof *G.p);
  if (G.p == NULL) {
    G.D = 0;
    goto done;
  }
  memcpy(G.p, p, (G.pw + 1) * sizeof *G.p);
  free_caches();
  G.generation++;
  if (build_table() != 0) {
    G.D = 0;
    goto done;
  }
  rc = 0;
done:
  free(p);
  free(q);
  free(r);
  free(seq);
  return rc;
}

int chol_dsfmt_poly_ensure(const uint64_t s0[CHOL_DSFMT_WORDS][2], int *D, unsigned *generation) {
  uint64_t v[CHOL_DSFMT_WORDS][2];
  int rc = 0;
  pthread_mutex_lock(&G.lock);
  memcpy(v, s0, sizeof v);
  if (G.D > 0) {
    apply_poly(G.p, G.D, v);
    if (!state_is_zero(v)) {
      memcpy(v, s0, sizeof v);
      rc = refine_poly(v);
    }
  } else {
    rc = refine_poly(v);
  }
  if (rc == 0) {
    *D = G.D;
    *generation = G.generation;
  }
  pthread_mutex_unlock(&G.lock);
  return rc;
}

/* rows a_((j0 + i) * 2^level), i < count, as 32-bit little-endian words, stride_u32 words apart (>= 2 * pw) */
int chol_dsfmt_poly_rows(unsigned generation, int level, uint64_t j0, size_t count, uint32_t *dst, size_t stride_u32) {
  int rc = -1;
  pthread_mutex_lock(&G.lock);
  if (G.D == 0 || generation != G.generation || level < 3 || level >= MAX_LEVEL || count == 0 || stride_u32 < 2 * G.pw) goto out;
  const size_t pw = G.pw;
  level_cache *lc = &G.level[level];
  const uint64_t Lr = UINT64_C(1) << level; /* refills between rows */
  if (!(lc->rows != NULL && j0 >= lc->j0 && j0 + count <= lc->j0 + lc->count)) {
    uint64_t *rows = malloc(count * pw * sizeof *rows);
    uint64_t *cur = calloc(pw, sizeof *cur);
    if (rows == NULL || cur == NULL) {
      free(rows);
      free(cur);
      goto out;
    }
    if (G.ck_count == 0) { /* checkpoint 0 = the polynomial 1 */
      G.ck = calloc(16 * pw, sizeof *G.ck);
      if (G.ck == NULL) {
        free(rows);
        free(cur);
        goto out;
      }
      G.ck_cap = 16;
      G.ck[0] = 1;
      G.ck_count = 1;
    }
    const uint64_t target = j0 * Lr;
    uint64_t m = target / CK_STRIDE;
    if (m >= G.ck_count) m = G.ck_count - 1;
    uint64_t r = m * CK_STRIDE;
    memcpy(cur, G.ck + m * pw, pw * sizeof *cur);
    if (lc->rows != NULL && lc->j0 <= j0) { /* a cached row of this level at or before the target, later than the checkpoint */
      uint64_t jc = j0 < lc->j0 + lc->count ? j0 : lc->j0 + lc->count - 1;
      if (jc * Lr > r) {
        r = jc * Lr;
        memcpy(cur, lc->rows + (jc - lc->j0) * pw, pw * sizeof *cur);
      }
    }
    bool oom = false;
    for (size_t i = 0; i < count && !oom; i++) {
      const uint64_t want = (j0 + i) * Lr;
      while (r < want) {
        step8(cur);
        r += 8;
        if (r % CK_STRIDE == 0 && r / CK_STRIDE == G.ck_count) {
          if (G.ck_count == G.ck_cap) {
            uint64_t *n = realloc(G.ck, 2 * G.ck_cap * pw * sizeof *n);
            if (n == NULL) {
              oom = true;
              break;
            }
            G.ck = n;
            G.ck_cap *= 2;
          }
          memcpy(G.ck + G.ck_count * pw, cur, pw * sizeof *cur);
          G.ck_count++;
        }
      }
      memcpy(rows + i * pw, cur, pw * sizeof *cur);
    }
    free(cur);
    if (oom) {
      free(rows);
      goto out;
    }
    free(lc->rows);
    lc->rows = rows;
    lc->j0 = j0;
    lc->count = count;
  }
  for (size_t i = 0; i < count; i++) {
    const uint64_t *src = lc->rows + (j0 - lc->j0 + i) * pw;
    uint32_t *d = dst + i * stride_u32;
    for (size_t w = 0; w < pw; w++) {
      d[2 * w] = (uint32_t)src[w];
      d[2 * w + 1] = (uint32_t)(src[w] >> 32);
    }
    for (size_t w = 2 * pw; w < stride_u32; w++) d[w] = 0;
  }
  rc = 0;
out:
  pthread_mutex_unlock(&G.lock);
  return rc;
}

/* test hook: forget the polynomial and every cache (the next seed recomputes it) */
void chol_dsfmt_poly_reset(void) {
  pthread_mutex_lock(&G.lock);
  free_caches();
  free(G.p);
  free(G.tab);
  G.p = G.tab = NULL;
  G.D = 0;
  G.pw = 0;
  G.generation++;
  pthread_mutex_unlock(&G.lock);
}
