// ptx_emu.h — CPU stand-ins for the inline-PTX primitives of csrc/ptx_sm100.cuh (included from there
// when CHOL_CPU_EMU is defined): mbarrier (phase + pending-arrival + transaction-byte accounting,
// blocking waits), TMA bulk copies (deferred: the data lands when a thread waits on the copy's barrier), cp.async with mbarrier arrival, and the
// m8n8k4 FP64 MMA as a warp-collective exchange.  TEST INFRASTRUCTURE ONLY.
//
// A wait on a phase that can never complete blocks forever here exactly as it would spin forever on the
// GPU - which is the point: protocol deadlocks show up on the CPU (the harness runs under a timeout).
#pragma once
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <vector>

namespace chol {

namespace emu_mbar {
struct State {
  long long count = 0, pending = 0, tx = 0;
  unsigned phase = 0;  // number of completed phases
};
inline std::mutex &mu() { static std::mutex m; return m; }
inline std::condition_variable &cv() { static std::condition_variable c; return c; }
inline std::map<const void *, State> &table() { static std::map<const void *, State> t; return t; }
// who is blocked where: written around every blocking wait, dumped by the watchdog (EMU_WATCHDOG=seconds)
struct Waiting { const void *bar = nullptr; unsigned parity = 0; };
inline Waiting *waiting() { static Waiting w[1024]; return w; }
inline void dump(const char *why) {
  std::lock_guard<std::mutex> l(mu());
  std::fprintf(stderr, "emu watchdog: %s\n", why);
  for (auto &kv : table())
    std::fprintf(stderr, "  mbarrier %p: count %lld pending %lld tx %lld completed phases %u\n", kv.first, kv.second.count, kv.second.pending,
                 kv.second.tx, kv.second.phase);
  for (int t = 0; t < 1024; ++t)
    if (waiting()[t].bar) std::fprintf(stderr, "  thread %d (warp %d) waits on %p for parity %u\n", t, t / 32, waiting()[t].bar, waiting()[t].parity);
}
inline void settle(State &s) {  // phase completion: all arrivals in and all announced bytes landed
  if (s.pending == 0 && s.tx == 0) {
    ++s.phase;
    s.pending = s.count;
    cv().notify_all();
  }
}
inline void arrive(const void *bar, long long n, long long expect_bytes) {
  std::lock_guard<std::mutex> l(mu());
  State &s = table()[bar];
  s.tx += expect_bytes;
  s.pending -= n;
  if (s.pending < 0) { std::fprintf(stderr, "emu: mbarrier %p over-arrived\n", bar); std::abort(); }
  settle(s);
}
inline void complete_tx(const void *bar, long long bytes) {
  std::lock_guard<std::mutex> l(mu());
  State &s = table()[bar];
  s.tx -= bytes;
  settle(s);
}
// Asynchronous copies that report to a barrier (TMA): the data lands as LATE as the program allows - when a thread
// actually waits on that barrier (or at once if somebody already does).  A consumer that reads the stage without
// waiting for its barrier therefore reads stale data here; with an immediate copy that bug would be invisible.
// EMU_TMA_EAGER=1 restores the immediate copy.
struct Pending { std::function<void()> copy; long long bytes; };
inline std::map<const void *, std::vector<Pending>> &pending() { static std::map<const void *, std::vector<Pending>> p; return p; }
inline std::map<const void *, int> &waiters() { static std::map<const void *, int> w; return w; }
inline bool eager() { static const bool e = [] { const char *v = std::getenv("EMU_TMA_EAGER"); return v && std::atoi(v); }(); return e; }
inline void land_locked(const void *bar) {  // mu() held
  auto it = pending().find(bar);
  if (it == pending().end()) return;
  State &s = table()[bar];
  for (Pending &p : it->second) {
    p.copy();
    s.tx -= p.bytes;
  }
  pending().erase(it);
  settle(s);
}
inline void async_copy(const void *bar, long long bytes, std::function<void()> copy) {
  std::lock_guard<std::mutex> l(mu());
  if (eager() || waiters()[bar] > 0) {
    copy();
    State &s = table()[bar];
    s.tx -= bytes;
    settle(s);
  } else {
    pending()[bar].push_back(Pending{std::move(copy), bytes});
  }
}
}  // namespace emu_mbar

inline void mbar_init(uint64_t *bar, uint32_t count) {
  std::lock_guard<std::mutex> l(emu_mbar::mu());
  emu_mbar::State s;
  s.count = s.pending = count;
  emu_mbar::table()[bar] = s;
  emu_mbar::pending().erase(bar);  // (a previous block's shared memory lived at the same address)
  emu_mbar::waiters()[bar] = 0;
}
inline void mbar_fence_init() {}
inline void mbar_arrive(uint64_t *bar) { ::emu::jitter(); emu_mbar::arrive(bar, 1, 0); }
inline void mbar_arrive_n(uint64_t *bar, uint32_t n) { emu_mbar::arrive(bar, n, 0); }
inline void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) { emu_mbar::arrive(bar, 1, bytes); }
// returns once the phase with this parity has completed (phase k has parity k & 1; a waiter is never
// two phases behind in a correct protocol - the 1-bit parity cannot tell, neither here nor on the GPU)
inline void mbar_wait(uint64_t *bar, uint32_t parity) {
  ::emu::jitter();
  std::unique_lock<std::mutex> l(emu_mbar::mu());
  emu_mbar::State &s = emu_mbar::table()[bar];
  emu_mbar::waiting()[threadIdx.x & 1023] = emu_mbar::Waiting{bar, parity};
  emu_mbar::land_locked(bar);  // deferred TMA copies reporting to this barrier land now, not before
  ++emu_mbar::waiters()[bar];
  emu_mbar::cv().wait(l, [&] { return (s.phase & 1u) != (parity & 1u); });
  --emu_mbar::waiters()[bar];
  emu_mbar::waiting()[threadIdx.x & 1023] = emu_mbar::Waiting{};
}

inline void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  if ((bytes & 15u) || ((uintptr_t)dst_smem & 15u) || ((uintptr_t)src_gmem & 15u)) {
    std::fprintf(stderr, "emu: cp.async.bulk needs 16-byte aligned size and addresses (%u bytes, dst %p, src %p)\n", bytes, dst_smem, src_gmem);
    std::abort();
  }
  emu_mbar::async_copy(bar, bytes, [=] { std::memcpy(dst_smem, src_gmem, bytes); });
}
inline void l2_prefetch_bulk(const void *, uint32_t) {}

// cp.async (LDGSTS) of one element; cp_async_mbar_arrive_noinc makes the barrier see the thread's
// copies: here the copies land right before that arrival (the latest moment the program allows)
template <int BYTES>
inline void cp_async_elem(void *dst_smem, const void *src_gmem) { ::emu::cp_async(dst_smem, src_gmem, BYTES); }
inline void cp_async_mbar_arrive_noinc(uint64_t *bar) {
  ::emu::cp_async_commit();
  ::emu::cp_async_wait(0);
  emu_mbar::arrive(bar, 1, 0);
}

// D(8x8) += A(8x4) * B(4x8).  lane = 4*g + t:  a = A[g][t], b = B[t][g], d0/d1 = D[g][2t], D[g][2t+1]
inline void dmma_884(double &d0, double &d1, double a, double b) {
  ::emu::Warp &w = ::emu::current()->warps[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  double ab[2] = {a, b};
  std::memcpy(w.slot[lane], ab, sizeof ab);
  w.bar.wait();
  for (int k = 0; k < 4; ++k) {
    double ak, b0, b1;
    std::memcpy(&ak, w.slot[4 * g + k], 8);                 // A[g][k]
    std::memcpy(&b0, w.slot[4 * (2 * t) + k] + 8, 8);       // B[k][2t]
    std::memcpy(&b1, w.slot[4 * (2 * t + 1) + k] + 8, 8);   // B[k][2t+1]
    d0 = std::fma(ak, b0, d0);
    d1 = std::fma(ak, b1, d1);
  }
  w.bar.wait();
}

}  // namespace chol
