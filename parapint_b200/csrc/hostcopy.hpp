// Host-side gather / scatter of many small arrays into / out of one staging buffer with a few worker threads.
// The plugin's inputs are ~2 leaves per block (COO value arrays of K_i and A_i, right-hand-side blocks) scattered
// over the Python heap; copying them one by one from Python costs more than the GPU spends factorising them.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include <cstdlib>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace ppb {

// Copy into the pinned staging buffer with non-temporal stores: the next reader is the GPU's DMA engine, and
// lines left dirty in several cores' caches make that read slower than the whole factorisation step.
inline void stream_copy(char *dst, const char *src, size_t n) {
#if defined(__x86_64__)
  static const bool nt = [] {
    const char *e = std::getenv("PARAPINT_B200_NT_STORES");
    return !(e && e[0] == '0');
  }();
  if (nt && n >= 256 && ((reinterpret_cast<uintptr_t>(dst) ^ reinterpret_cast<uintptr_t>(src)) & 7) == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
    while ((reinterpret_cast<uintptr_t>(dst) & 15) && n >= 8) {
      std::memcpy(dst, src, 8);
      dst += 8; src += 8; n -= 8;
    }
    size_t blocks = n / 64;
    for (size_t b = 0; b < blocks; ++b) {
      const __m128i a0 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src));
      const __m128i a1 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 16));
      const __m128i a2 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 32));
      const __m128i a3 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 48));
      _mm_stream_si128(reinterpret_cast<__m128i *>(dst), a0);
      _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 16), a1);
      _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 32), a2);
      _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 48), a3);
      src += 64; dst += 64;
    }
    n -= blocks * 64;
    if (n) std::memcpy(dst, src, n);
    _mm_sfence();
    return;
  }
#endif
  std::memcpy(dst, src, n);
}

class CopyPool {
 public:
  static CopyPool &instance() {
    static CopyPool p;
    return p;
  }
  // segs[k]: `len[k]` bytes at `ptr[k]` <-> staging offset `off[k]`; to_staging selects the direction
  void run(int64_t nseg, void *const *ptr, const int64_t *off, const int64_t *len, char *staging, bool to_staging,
           int threads) {
    dispatch({nseg, ptr, off, len, staging, to_staging, 0, 0, nullptr}, threads);
  }
  // true when the bytes of a[k] and b[k] (len[k] each) agree for every k: the index arrays of a freshly built KKT
  // matrix against the analysed pattern, compared by a few threads instead of one numpy call per leaf
  bool equal(int64_t nseg, void *const *a, void *const *b, const int64_t *len, int threads) {
    mismatch_.store(0, std::memory_order_relaxed);
    dispatch({nseg, a, nullptr, len, nullptr, false, 0, 0, b}, threads);
    return mismatch_.load(std::memory_order_relaxed) == 0;
  }
  // A job follows shortly (the caller is about to wait for the GPU and will scatter the result, or is walking the
  // matrix whose values it will gather): wake the workers now and let them spin for it for at most `spin_ns`, so the
  // job does not pay a futex wake-up (tens of microseconds -- as much as copying a megabyte).  Bounded: without a job
  // the workers go back to sleep when the window closes.
  void wake(int threads, int64_t spin_ns) {
    const int T = std::min(std::max(threads, 1), kMax);
    if (T <= 1 || spin_ns <= 0) return;
    std::lock_guard<std::mutex> serial(api_);
    ensure_workers(T - 1);
    {
      std::lock_guard<std::mutex> lk(mu_);
      ++ping_;
      ping_spin_ns_ = std::min(spin_ns, kMaxWakeNs);
    }
    cv_.notify_all();
  }

 private:
  static constexpr int kMax = 8;
  struct Job {
    int64_t nseg;
    void *const *ptr;
    const int64_t *off, *len;
    char *staging;
    bool to_staging;
    int64_t total;
    int T;
    void *const *other;   // non-null: compare ptr[k] with other[k] instead of copying
  };
  void dispatch(Job j, int threads) {
    int64_t total = 0;
    for (int64_t k = 0; k < j.nseg; ++k) total += j.len[k];
    int T = (int)std::min<int64_t>(std::max(threads, 1), 1 + total / (256 << 10));  // >= 256 KB per thread
    T = std::min(T, kMax);
    j.total = total;
    j.T = T;
    if (T <= 1) {
      work(j, 0, total);
      return;
    }
    std::lock_guard<std::mutex> serial(api_);
    ensure_workers(T - 1);
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = j;
      pending_ = T - 1;
      left_.store(T - 1, std::memory_order_relaxed);
      ++gen_;
      gen_pub_.store(gen_, std::memory_order_release);
    }
    cv_.notify_all();
    work(j, 0, total / T);
    // the shares are equal, so the workers finish within microseconds of this thread: spin before sleeping
    for (const auto t0 = now_ns(); left_.load(std::memory_order_acquire) != 0 && now_ns() - t0 < kSpinNs;) cpu_relax();
    if (left_.load(std::memory_order_acquire) != 0) {
      std::unique_lock<std::mutex> lk(mu_);
      done_.wait(lk, [&] { return pending_ == 0; });
    }
    while (left_.load(std::memory_order_acquire) != 0) cpu_relax();   // every worker is out of its bookkeeping
  }
  static void cpu_relax() {
#if defined(__x86_64__)
    _mm_pause();
#endif
  }
  void work(const Job &j, int64_t b0, int64_t b1) {
    if (!j.other) {
      copy_range(j.nseg, j.ptr, j.off, j.len, j.staging, j.to_staging, b0, b1);
      return;
    }
    int64_t pos = 0;
    for (int64_t k = 0; k < j.nseg && pos < b1; ++k) {
      const int64_t lo = std::max(pos, b0), hi = std::min(pos + j.len[k], b1);
      if (lo < hi && std::memcmp(static_cast<const char *>(j.ptr[k]) + (lo - pos),
                                 static_cast<const char *>(j.other[k]) + (lo - pos), (size_t)(hi - lo)) != 0) {
        mismatch_.store(1, std::memory_order_relaxed);
        return;
      }
      pos += j.len[k];
    }
  }
  // copy the part of the concatenated byte stream [b0, b1) (segments in order)
  static void copy_range(int64_t nseg, void *const *ptr, const int64_t *off, const int64_t *len, char *staging,
                         bool to_staging, int64_t b0, int64_t b1) {
    int64_t pos = 0;
    for (int64_t k = 0; k < nseg && pos < b1; ++k) {
      const int64_t lo = std::max(pos, b0), hi = std::min(pos + len[k], b1);
      if (lo < hi) {
        char *user = static_cast<char *>(ptr[k]) + (lo - pos);
        char *st = staging + off[k] + (lo - pos);
        if (to_staging) stream_copy(st, user, (size_t)(hi - lo));
        else std::memcpy(user, st, (size_t)(hi - lo));
      }
      pos += len[k];
    }
  }
  void ensure_workers(int n) {
    while ((int)workers_.size() < n) {
      const int id = (int)workers_.size() + 1;
      workers_.emplace_back([this, id] { loop(id); });
    }
  }
  void loop(int id) {
    uint64_t seen = 0, seen_ping = 0;
    int64_t window = kSpinNs;
    for (;;) {
      Job j;
      // a gather comes in a few shares a few microseconds apart (pp_stage_values): stay awake for a moment after a job
      // instead of paying a futex wake-up per share; after a wake-up call (wake) for as long as the caller asked
      for (const auto t0 = now_ns(); gen_pub_.load(std::memory_order_acquire) == seen && now_ns() - t0 < window;) cpu_relax();
      window = kSpinNs;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || gen_ != seen || ping_ != seen_ping; });
        if (stop_) return;
        if (gen_ == seen) {   // a wake-up call, no job yet: spin at the top of the loop for the job that follows
          seen_ping = ping_;
          window = ping_spin_ns_;
          continue;
        }
        seen = gen_;
        seen_ping = ping_;
        j = job_;
      }
      if (id >= j.T) continue;   // this job uses fewer threads
      work(j, j.total * id / j.T, j.total * (id + 1) / j.T);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
      left_.fetch_sub(1, std::memory_order_acq_rel);   // last: the dispatcher may start the next job once this is 0
    }
  }
  CopyPool() = default;
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
  }
  static constexpr int64_t kSpinNs = 30000;   // spin for at most 30 us (by the clock: a `pause` is 10-140 cycles)
  static constexpr int64_t kMaxWakeNs = 2000000;   // a wake-up call keeps the workers spinning for at most 2 ms
  static int64_t now_ns() {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }
  std::atomic<int> mismatch_{0};
  std::atomic<int> left_{0};
  std::atomic<uint64_t> gen_pub_{0};
  std::mutex api_, mu_;
  std::condition_variable cv_, done_;
  std::vector<std::thread> workers_;
  Job job_{};
  uint64_t gen_ = 0, ping_ = 0;
  int64_t ping_spin_ns_ = 0;
  int pending_ = 0;
  bool stop_ = false;
};

}  // namespace ppb
