// emu_rng.cpp — csrc/rng_dsfmt.cu (the device dSFMT-19937: warp-parallel refill, windowed conversion / store, Krylov
// snapshots, the GF(2) jump kernel) executed on the CPU (tests/emu/cuda_emu.h) and compared bit for bit with the
// oracle's sequential generator (oracle/dsfmt_oracle.c, pinned to the reference's rng/dsfmt_19937.c).  The start states
// and polynomial rows come from the product's host code (host/rng_poly.c), linked in as C.
// Prints "EMU PASSED" / exit code 1.  TEST INFRASTRUCTURE ONLY.
#define CUDA_EMU_IMPLEMENTATION
#include "cuda_emu.h"

#include "../../cuda-cholesky_b200/csrc/rng_dsfmt.cu"

#include <random>
#include <string>

extern "C" {
#include "../../cuda-cholesky_b200/host/rng_internal.h"
// oracle/dsfmt_oracle.c
size_t oracle_dsfmt_state_size(void);
void oracle_dsfmt_seed(void *g, uint64_t seed);
void oracle_dsfmt_fill(void *g, uint64_t *out, size_t n, int kind);
}

using namespace chol;

static int failures = 0, cases = 0;
typedef uint64_t state_t[CHOL_DSFMT_WORDS][2];

static void expect(bool ok, const char *what) {
  ++cases;
  if (!ok) {
    ++failures;
    std::printf("MISMATCH: %s\n", what);
  }
}

// the oracle's stream: words [w_begin, w_end) of seed's stream as n_out converted values (an odd n_out drops the last high half)
static std::vector<uint64_t> oracle_window(uint64_t seed, unsigned long long w_begin, unsigned long long n_out, int kind) {
  std::vector<unsigned char> g(oracle_dsfmt_state_size());
  oracle_dsfmt_seed(g.data(), seed);
  std::vector<uint64_t> skip(2 * w_begin + 2);
  if (w_begin) oracle_dsfmt_fill(g.data(), skip.data(), 2 * w_begin, kind);  // an even count: no half is dropped
  std::vector<uint64_t> out(n_out + 1, 0);
  oracle_dsfmt_fill(g.data(), out.data(), n_out, kind);
  out.resize(n_out);
  return out;
}

static void run_generate(int kind, const W128 *states, unsigned long long refill0, unsigned seg, int nseg, unsigned long long w0,
                         unsigned long long w1, unsigned long long *out, unsigned long long n_out, W128 *snap) {
  const dim3 grid((nseg + GEN_WARPS - 1) / GEN_WARPS), block(GEN_WARPS * 32);
  switch (kind) {
    case 0: emu::launch(grid, block, 0, [&] { dsfmt_generate_kernel<0>(states, refill0, seg, nseg, w0, w1, out, n_out, snap); }); break;
    case 1: emu::launch(grid, block, 0, [&] { dsfmt_generate_kernel<1>(states, refill0, seg, nseg, w0, w1, out, n_out, snap); }); break;
    case 2: emu::launch(grid, block, 0, [&] { dsfmt_generate_kernel<2>(states, refill0, seg, nseg, w0, w1, out, n_out, snap); }); break;
    case 3: emu::launch(grid, block, 0, [&] { dsfmt_generate_kernel<3>(states, refill0, seg, nseg, w0, w1, out, n_out, snap); }); break;
    default: emu::launch(grid, block, 0, [&] { dsfmt_generate_kernel<4>(states, refill0, seg, nseg, w0, w1, out, n_out, snap); }); break;
  }
}

// ./emu_rng fuzz SEED COUNT: random segmentations (refills per segment, segment count, first refill) and random windows
// inside them (start word, value count odd / even, 8- or 16-byte aligned pointer, output kind) against the oracle's stream
static int fuzz(unsigned fseed, int count) {
  const uint64_t seed = 20121018u;
  state_t s0;
  chol_dsfmt_seed_state(seed, s0);
  std::mt19937 g(fseed);
  auto ri = [&](unsigned long long lo, unsigned long long hi) { return lo + (unsigned long long)(g() % (unsigned)(hi - lo + 1)); };
  for (int it = 0; it < count; ++it) {
    const unsigned seg = (unsigned)ri(1, 5);
    const int nseg = (int)ri(1, 9);
    const unsigned long long refill0 = ri(0, 6);
    std::vector<W128> states((size_t)nseg * DS_STATE_WORDS);
    state_t cur;
    std::memcpy(cur, s0, sizeof cur);
    for (unsigned long long r = 0, sg = 0; sg < (unsigned long long)nseg; ++r) {
      if (r == refill0 + sg * seg) {
        std::memcpy(&states[sg * DS_STATE_WORDS], cur, sizeof cur);
        ++sg;
      }
      chol_dsfmt_refill_host(cur);
    }
    const unsigned long long lo = refill0 * 191, hi = (refill0 + (unsigned long long)nseg * seg) * 191;  // covered stream words
    const unsigned long long w0 = ri(lo, hi - 1);
    unsigned long long words = ri(0, 3) == 0 ? ri(1, 3) : ri(1, hi - w0);
    if (words > hi - w0) words = hi - w0;
    const unsigned long long n = 2 * words - (ri(0, 1) ? 1 : 0);  // an odd count drops the last high half
    const int kind = (int)ri(0, 4), misalign = (int)ri(0, 1);
    std::vector<unsigned long long> buf(n + 8, 0x3333333333333333ull);
    unsigned long long *out = buf.data() + 2;
    while (((uintptr_t)out & 15) != (misalign ? 8u : 0u)) ++out;
    run_generate(kind, states.data(), refill0, seg, nseg, w0, w0 + words, out, n, nullptr);
    const auto want = oracle_window(seed, w0, n, kind);
    const bool ok = std::memcmp(out, want.data(), n * 8) == 0 && out[n] == 0x3333333333333333ull && out[-1] == 0x3333333333333333ull;
    char what[160];
    std::snprintf(what, sizeof what, "fuzz: %d segments of %u refills from refill %llu, window at word %llu, %llu values, kind %d, %s pointer", nseg, seg,
                  refill0, w0, n, kind, misalign ? "8-byte" : "16-byte");
    expect(ok, what);
  }
  std::printf("fuzz: %d cases, %d mismatches\n", cases, failures);
  if (failures) return 1;
  std::printf("EMU PASSED\n");
  return 0;
}

int main(int argc, char **argv) {
  if (argc > 3 && std::string(argv[1]) == "fuzz") return fuzz((unsigned)atoi(argv[2]), atoi(argv[3]));
  const uint64_t seed = 20121018u;
  state_t s0;
  chol_dsfmt_seed_state(seed, s0);

  // ---- 1. one segment from the seed state: the warp-parallel refill against the oracle, every output kind ----
  for (int kind = 0; kind < 5; ++kind) {
    const unsigned refills = 4;
    const unsigned long long n = 2ull * 191 * refills;
    std::vector<unsigned long long> out(n + 2, 0x1111111111111111ull);
    run_generate(kind, (const W128 *)s0, 0, refills, 1, 0, 191ull * refills, out.data(), n, nullptr);
    const auto want = oracle_window(seed, 0, n, kind);
    bool ok = std::memcmp(out.data(), want.data(), n * 8) == 0 && out[n] == 0x1111111111111111ull;
    char what[96];
    std::snprintf(what, sizeof what, "kind %d: %u refills of one segment vs the oracle", kind, refills);
    expect(ok, what);
    std::printf("kind %d: first 4 refills identical to the oracle: %s\n", kind, ok ? "yes" : "NO");
  }

  // ---- 2. several segments (two CTAs, the second one partly empty), windows that start and end inside a refill, odd
  //         counts, output pointers that are only 8-byte aligned ----
  {
    const unsigned seg = 3;
    const int nseg = 6;
    const unsigned long long refill0 = 2;  // the segments start after refills 2, 5, 8, ...
    std::vector<W128> states((size_t)nseg * DS_STATE_WORDS);
    state_t cur;
    std::memcpy(cur, s0, sizeof cur);
    for (unsigned long long r = 0, g = 0; g < (unsigned long long)nseg; ++r) {
      if (r == refill0 + g * seg) {
        std::memcpy(&states[g * DS_STATE_WORDS], cur, sizeof cur);
        ++g;
      }
      chol_dsfmt_refill_host(cur);
    }
    struct { unsigned long long w0, n; int misalign; } win[] = {
        {2 * 191, 2ull * 191 * 18, 0},       // exactly the segments' refills
        {2 * 191 + 5, 2000, 0},              // starts inside refill 2
        {3 * 191 + 190, 3333, 1},            // odd count, unaligned pointer, starts at the last word of a refill
        {7 * 191 + 17, 1, 0},                // a single value
        {19 * 191 + 100, 181, 1},            // the tail of the last segment's last refill, odd
    };
    for (auto &w : win)
      for (int kind = 0; kind < 5; kind += (w.n > 3000 ? 1 : 2)) {
        const unsigned long long words = w.n / 2 + (w.n & 1);
        std::vector<unsigned long long> buf(w.n + 8, 0x2222222222222222ull);
        unsigned long long *out = buf.data() + 2;
        while (((uintptr_t)out & 15) != (w.misalign ? 8u : 0u)) ++out;
        run_generate(kind, states.data(), refill0, seg, nseg, w.w0, w.w0 + words, out, w.n, nullptr);
        const auto want = oracle_window(seed, w.w0, w.n, kind);
        bool ok = std::memcmp(out, want.data(), w.n * 8) == 0 && out[w.n] == 0x2222222222222222ull && out[-1] == 0x2222222222222222ull;
        char what[128];
        std::snprintf(what, sizeof what, "window at word %llu, %llu values, kind %d, %s pointer", w.w0, w.n, kind, w.misalign ? "8-byte" : "16-byte");
        expect(ok, what);
      }
    std::printf("windowed multi-segment generation: %d cases so far, %d mismatches\n", cases, failures);
  }

  // ---- 3. snapshot mode: the Krylov rows ----
  {
    const int rows = 7;
    std::vector<W128> K((size_t)rows * DS_STATE_WORDS);
    std::memcpy(K.data(), s0, sizeof s0);
    run_generate(0, K.data(), 0, rows - 1, 1, 0, 0, nullptr, 0, K.data());
    state_t cur;
    std::memcpy(cur, s0, sizeof cur);
    bool ok = true;
    for (int i = 0; i < rows; ++i) {
      ok = ok && std::memcmp(&K[(size_t)i * DS_STATE_WORDS], cur, sizeof cur) == 0;
      chol_dsfmt_refill_host(cur);
    }
    expect(ok, "snapshot rows K_i = T^i s0");
    std::printf("snapshot mode (Krylov rows): %s\n", ok ? "ok" : "MISMATCH");
  }

  // ---- 4. the jump kernel: start states as GF(2) combinations of the Krylov block, with the product's own polynomial rows ----
  {
    int D = 0;
    unsigned gen = 0;
    if (chol_dsfmt_poly_ensure((const uint64_t(*)[2])s0, &D, &gen) != 0) {
      std::printf("MISMATCH: no polynomial\n");
      return 1;
    }
    std::printf("minimal polynomial: degree %d\n", D);
    std::vector<uint32_t> K((size_t)D * DS_STATE_U32);
    state_t cur;
    std::memcpy(cur, s0, sizeof cur);
    for (int i = 0; i < D; ++i) {
      std::memcpy(&K[(size_t)i * DS_STATE_U32], cur, sizeof cur);
      chol_dsfmt_refill_host(cur);
    }
    const int poly_u32 = 2 * ((D + 63) / 64);
    struct { int level; uint64_t j0; int nseg; unsigned gy; } jc[] = {{4, 3, 11, 3}, {6, 0, 8, 1}, {5, 700, 3, 20}, {4, 100, 37, 7}, {7, 1, 16, 14}};
    for (auto &c : jc) {
      std::vector<uint32_t> polys((size_t)c.nseg * poly_u32);
      if (chol_dsfmt_poly_rows(gen, c.level, c.j0, (size_t)c.nseg, polys.data(), (size_t)poly_u32) != 0) {
        std::printf("MISMATCH: polynomial rows\n");
        return 1;
      }
      std::vector<uint32_t> states((size_t)c.nseg * DS_STATE_U32, c.gy > 1 ? 0u : 0xdeadbeefu);
      const dim3 grid((c.nseg + JUMP_GT - 1) / JUMP_GT, c.gy);
      emu::launch(grid, dim3(JUMP_THREADS), 0, [&] { dsfmt_jump_kernel((const uint4 *)K.data(), polys.data(), poly_u32, D, states.data(), c.nseg); });
      // direct: iterate the host refill to every segment start
      bool ok = true;
      std::memcpy(cur, s0, sizeof cur);
      unsigned long long r = 0;
      for (int g = 0; g < c.nseg; ++g) {
        const unsigned long long want_r = (c.j0 + g) << c.level;
        for (; r < want_r; ++r) chol_dsfmt_refill_host(cur);
        ok = ok && std::memcmp(&states[(size_t)g * DS_STATE_U32], cur, sizeof cur) == 0;
      }
      char what[128];
      std::snprintf(what, sizeof what, "jump to refills %llu + g * %u, %d segments, %u slices", (unsigned long long)(c.j0 << c.level), 1u << c.level, c.nseg, c.gy);
      expect(ok, what);
      std::printf("%s: %s\n", what, ok ? "ok" : "MISMATCH");
      // ... and the generation kernel continues from the jumped states exactly where the oracle's stream is
      if (ok && c.level == 4) {
        const unsigned long long w0 = (c.j0 << c.level) * 191 + 77, n = 5001;
        std::vector<unsigned long long> out(n + 2, 0);
        run_generate(1, (const W128 *)states.data(), c.j0 << c.level, 1u << c.level, c.nseg, w0, w0 + (n + 1) / 2, out.data(), n, nullptr);
        const auto want = oracle_window(seed, w0, n, 1);
        const bool ok2 = std::memcmp(out.data(), want.data(), n * 8) == 0;
        expect(ok2, "jump + generate vs the oracle's stream");
        std::printf("jump + generate, 5001 values from word %llu: %s\n", w0, ok2 ? "identical to the oracle" : "MISMATCH");
      }
    }
  }
  std::printf("%d cases, %d mismatches\n", cases, failures);
  if (failures) return 1;
  std::printf("EMU PASSED\n");
  return 0;
}
