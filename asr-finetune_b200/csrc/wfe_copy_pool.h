// Staging copy pool of the host entry (wfe_extract_host*): plain C++ (no CUDA), so that tests/host/copy_pool_host.cpp can
// build it with g++ -- also under ThreadSanitizer -- and stress it on the CPU (tests/test_host_cpu.py).
#pragma once
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

namespace wfe_host {

// Host-side staging copies (pageable caller memory <-> the pinned ring) on a few persistent threads: one memcpy stream
// moves ~8 GB/s, a PCIe 5 link 50+, so a single-threaded staging loop was what bound the drop-in call on the arrays the
// reference's loader yields (one separately allocated pageable array per clip): 62 ms for 256 clips against 11 ms of PCIe.
// (Non-temporal stores instead of memcpy were measured too: no gain, 15.5 vs 15.4 ms.)
class CopyPool {
 public:
  struct Job {
    char* dst;
    const char* src;
    size_t n;
  };
  explicit CopyPool(int n_threads) {
    for (int i = 0; i < n_threads; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  // copies every job (cut into pieces of at most 256 KB); the calling thread takes part; returns when all are done
  void run(const std::vector<Job>& jobs) {
    start(jobs);
    finish();
  }
  // The same in two halves: start() hands the pieces to the worker threads and returns, finish() joins in and waits
  // (one batch at a time: finish() before the next start()).  Pieces are claimed under the mutex -- 50 ns against the
  // 30 us a piece takes -- so that a worker that wakes up late, or is still leaving the previous batch, either gets a
  // piece of the CURRENT batch or nothing (the lock-free claim of the first version could pair an index of the old batch
  // with the new piece table).
  void start(const std::vector<Job>& jobs) {
    std::lock_guard<std::mutex> lk(mu_);
    pieces_.clear();
    for (const Job& j : jobs)
      for (size_t o = 0; o < j.n; o += kPiece) pieces_.push_back({j.dst + o, j.src + o, std::min(kPiece, j.n - o)});
    next_ = 0;
    left_ = pieces_.size();
    if (!pieces_.empty() && !workers_.empty() && pieces_.size() >= 4) {
      ++gen_;
      cv_.notify_all();
    }
  }
  void finish() {
    work();
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [this] { return left_ == 0; });
  }

 private:
  static constexpr size_t kPiece = 256 * 1024;
  void work() {
    for (;;) {
      Job q;
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (next_ >= pieces_.size()) return;
        q = pieces_[next_++];
      }
      memcpy(q.dst, q.src, q.n);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--left_ == 0) done_cv_.notify_all();
      }
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
      }
      work();
    }
  }
  std::vector<std::thread> workers_;
  std::vector<Job> pieces_;
  size_t next_ = 0, left_ = 0;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

}  // namespace wfe_host
