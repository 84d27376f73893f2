// CPU stress of the host entry's staging copy pool (asr-finetune_b200/csrc/wfe_copy_pool.h), the way wfe_extract_host
// drives it: one batch at a time, job lists of every size (empty, below the four-piece threshold that wakes the workers,
// thousands of pieces), back to back so that workers still leaving one batch meet the next.  Built by
// tests/test_host_cpu.py with g++ (and with -fsanitize=thread as a stand-alone program).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../asr-finetune_b200/csrc/wfe_copy_pool.h"

namespace {
uint64_t next(uint64_t& s) {
  s ^= s << 13;
  s ^= s >> 7;
  s ^= s << 17;
  return s;
}
}  // namespace

extern "C" {

// returns the number of batches in which a destination byte was wrong (0 = pass)
int copy_pool_stress(int threads, int batches, uint64_t seed, int use_start_finish) {
  wfe_host::CopyPool pool(threads);
  uint64_t s = seed ? seed : 1;
  const size_t kArena = (size_t)24 << 20;
  std::vector<char> src(kArena), dst(kArena);
  for (size_t i = 0; i < kArena; ++i) src[i] = (char)(next(s) >> 24);
  int bad = 0;
  for (int b = 0; b < batches; ++b) {
    // a batch: 0..40 clips of 0 .. 1.5 MB (1 to 6 pieces of 256 KB), packed like the ring packs them
    const int n = (int)(next(s) % 41);
    std::vector<wfe_host::CopyPool::Job> jobs;
    size_t pos = 0;
    const char fill = (char)(b * 37 + 1);
    for (int i = 0; i < n; ++i) {
      size_t len = (size_t)(next(s) % (3 << 19));
      if (next(s) % 5 == 0) len = next(s) % 64;  // tiny clips
      if (pos + len > kArena) break;
      jobs.push_back({dst.data() + pos, src.data() + pos, len});
      pos = (pos + len + 31) & ~(size_t)31;
    }
    for (const auto& j : jobs) memset(j.dst, fill, j.n);  // whatever the previous batch left there must be replaced
    if (use_start_finish) {
      pool.start(jobs);
      pool.finish();
    } else {
      pool.run(jobs);
    }
    bool ok = true;
    for (const auto& j : jobs) ok = ok && memcmp(j.dst, j.src, j.n) == 0;
    if (!ok) ++bad;
  }
  return bad;
}

}  // extern "C"

#ifdef COPY_POOL_MAIN
int main(int argc, char** argv) {
  const int threads = argc > 1 ? atoi(argv[1]) : 4, batches = argc > 2 ? atoi(argv[2]) : 200;
  const int bad = copy_pool_stress(threads, batches, 12345, 0) + copy_pool_stress(threads, batches, 999, 1) +
                  copy_pool_stress(0, 20, 5, 0);
  printf("copy_pool_stress: %d bad batches\n", bad);
  return bad != 0;
}
#endif
