// ThreadSanitizer harness for the host copy pool (csrc/hostcopy.hpp): many gathers / scatters / comparisons of random
// size and thread count, interleaved with wake-up calls, every result checked.  Build + run (no GPU):
//   g++ -std=c++17 -O1 -g -fsanitize=thread -pthread -I parapint_b200/csrc tools/tsan_copy_pool.cpp -o /tmp/tsan_copy_pool \
//       && /tmp/tsan_copy_pool 3000
#include <cstdio>
#include <cstdlib>
#include <random>

#include "hostcopy.hpp"

int main(int argc, char **argv) {
  const int rounds = argc > 1 ? std::atoi(argv[1]) : 2000;
  std::mt19937_64 rng(argc > 2 ? (uint64_t)std::atoll(argv[2]) : 7);
  auto &pool = ppb::CopyPool::instance();
  std::vector<std::vector<double>> src, back;
  std::vector<double> staging;
  for (int r = 0; r < rounds; ++r) {
    const int nseg = 1 + (int)(rng() % 40);
    const int threads = 1 + (int)(rng() % 8);
    src.assign(nseg, {});
    back.assign(nseg, {});
    std::vector<void *> ptr(nseg), bptr(nseg);
    std::vector<int64_t> off(nseg), len(nseg);
    int64_t total = 0;
    for (int k = 0; k < nseg; ++k) {
      const size_t n = (rng() % 4 == 0) ? (size_t)(rng() % 200000) : (size_t)(rng() % 3000);
      src[k].resize(n);
      back[k].assign(n, -1.0);
      for (size_t i = 0; i < n; ++i) src[k][i] = (double)(rng() % 1000003);
      ptr[k] = src[k].data();
      bptr[k] = back[k].data();
      off[k] = total * 8;
      len[k] = (int64_t)n * 8;
      total += (int64_t)n;
    }
    staging.assign((size_t)total + 1, -2.0);
    if (rng() % 3 == 0) pool.wake(threads, (int64_t)(rng() % 200000));
    pool.run(nseg, ptr.data(), off.data(), len.data(), reinterpret_cast<char *>(staging.data()), true, threads);
    for (int k = 0; k < nseg; ++k)
      if (len[k] && std::memcmp(reinterpret_cast<char *>(staging.data()) + off[k], src[k].data(), (size_t)len[k]) != 0) {
        std::fprintf(stderr, "gather mismatch in round %d segment %d\n", r, k);
        return 1;
      }
    if (rng() % 4 == 0) pool.wake(threads, (int64_t)(rng() % 50000));
    pool.run(nseg, bptr.data(), off.data(), len.data(), reinterpret_cast<char *>(staging.data()), false, threads);
    if (!pool.equal(nseg, ptr.data(), bptr.data(), len.data(), threads)) {
      std::fprintf(stderr, "scatter mismatch in round %d\n", r);
      return 1;
    }
    if (total > 0) {   // a single changed byte must be seen
      int k = (int)(rng() % nseg);
      while (len[k] == 0) k = (k + 1) % nseg;
      back[k][rng() % back[k].size()] += 1.0;
      if (pool.equal(nseg, ptr.data(), bptr.data(), len.data(), threads)) {
        std::fprintf(stderr, "difference missed in round %d\n", r);
        return 1;
      }
    }
    if (r % 500 == 0) std::printf("round %d: %d segments, %lld doubles, %d threads\n", r, nseg, (long long)total, threads);
  }
  std::printf("ok: %d rounds\n", rounds);
  return 0;
}
