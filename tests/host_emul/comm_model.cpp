// Host MODEL of the in-kernel all-reduce protocol of dair_pll_b200/csrc/cn_comm.cuh (test infrastructure only):
// the same buffers (CommBuf: monotone flags + two alternating data rows per rank), the same four steps per epoch
// (store rows everywhere, release-store the epoch into every flag, acquire-spin on the own flags, sum in rank
// order), with one std::thread per rank and randomised delays.  It checks what the device code relies on: with
// two data buffers and "flag >= epoch" waits no rank ever reads a row of the wrong epoch, however far ranks drift.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <random>
#include <thread>
#include <vector>

namespace {
constexpr int MAXW = 16, MAXE = 32;
struct Buf {
  std::atomic<unsigned long long> flag[MAXW];
  std::atomic<double> data[2][MAXW][MAXE];
};
}  // namespace

extern "C" int comm_model_run(int world, int epochs, int n, unsigned seed) {
  std::vector<Buf> bufs(world);
  for (auto& b : bufs) {
    for (int r = 0; r < MAXW; ++r) b.flag[r].store(0);
    for (int p = 0; p < 2; ++p)
      for (int r = 0; r < MAXW; ++r)
        for (int i = 0; i < MAXE; ++i) b.data[p][r][i].store(0.0);
  }
  std::atomic<int> bad{0};
  auto worker = [&](int rank) {
    std::mt19937 rng(seed * 977u + rank);
    unsigned long long epoch = 0;
    for (int k = 0; k < epochs; ++k) {
      if (rng() % 7 == 0) std::this_thread::sleep_for(std::chrono::microseconds(rng() % 50));
      const unsigned long long e = epoch + 1;
      const int p = (int)(e & 1ull);
      for (int r = 0; r < world; ++r)
        for (int i = 0; i < n; ++i)
          bufs[r].data[p][rank][i].store((double)(e * 1000 + rank * 10 + i), std::memory_order_relaxed);
      std::atomic_thread_fence(std::memory_order_seq_cst);
      for (int r = 0; r < world; ++r) bufs[r].flag[rank].store(e, std::memory_order_release);
      for (int r = 0; r < world; ++r)
        while (bufs[rank].flag[r].load(std::memory_order_acquire) < e) std::this_thread::yield();
      for (int i = 0; i < n; ++i) {
        double s = 0.0, want = 0.0;
        for (int r = 0; r < world; ++r) {
          s += bufs[rank].data[p][r][i].load(std::memory_order_relaxed);
          want += (double)(e * 1000 + r * 10 + i);
        }
        if (s != want) bad.fetch_add(1);
      }
      epoch = e;
    }
  };
  std::vector<std::thread> threads;
  for (int r = 0; r < world; ++r) threads.emplace_back(worker, r);
  for (auto& t : threads) t.join();
  return bad.load();
}
