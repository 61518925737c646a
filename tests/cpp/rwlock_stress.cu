// Host-only stress test of ehb::RwLock (csrc/index_impl.h), the writer-preferring reader/writer lock under every
// ehb_index: readers never overlap a writer, writers are exclusive, and a writer gets in while readers keep
// arriving back to back (the property glibc's reader-preferring rwlock does not give).  No GPU needed.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <shared_mutex>
#include <thread>
#include <vector>

#include "../../embeddinghub_b200/csrc/index_impl.h"

int main() {
  ehb::RwLock rw;
  std::atomic<int> readers{0}, writers{0}, violations{0}, writes_done{0};
  std::atomic<bool> stop{false};
  std::atomic<long> reads_done{0};
  std::vector<std::thread> th;
  for (int t = 0; t < 8; ++t)
    th.emplace_back([&] {
      while (!stop.load()) {
        std::shared_lock<ehb::RwLock> g(rw);
        readers++;
        if (writers.load() != 0) violations++;
        for (volatile int i = 0; i < 200; ++i) {
        }
        readers--;
        reads_done++;
      }
    });
  // readers now overlap continuously; every writer must still get through, and quickly
  double worst_ms = 0;
  for (int w = 0; w < 200; ++w) {
    auto t0 = std::chrono::steady_clock::now();
    {
      std::unique_lock<ehb::RwLock> g(rw);
      double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (ms > worst_ms) worst_ms = ms;
      if (writers.fetch_add(1) != 0) violations++;
      if (readers.load() != 0) violations++;
      for (volatile int i = 0; i < 2000; ++i) {
      }
      writers--;
      writes_done++;
    }
    std::this_thread::sleep_for(std::chrono::microseconds(200));
  }
  stop = true;
  for (auto& t : th) t.join();
  std::printf("writes %d reads %ld violations %d worst writer wait %.2f ms\n", writes_done.load(), reads_done.load(),
              violations.load(), worst_ms);
  bool ok = violations.load() == 0 && writes_done.load() == 200 && reads_done.load() > 1000 && worst_ms < 2000.0;
  std::printf(ok ? "OK\n" : "FAILED\n");
  return ok ? 0 : 1;
}
