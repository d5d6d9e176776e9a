// See host_pack.h.  The inner loops have an AVX2 form (32 source elements per step), selected at run time, and a
// portable form; both produce the same words.
#include "host_pack.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#ifdef __linux__
#include <sched.h>
#endif

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define QLDPC_X86 1
#endif

namespace qldpc {

int default_host_threads() {
  if (const char* env = std::getenv("QLDPC_HOST_THREADS")) return std::max(0, std::min(64, std::atoi(env)));
  unsigned hw = std::thread::hardware_concurrency();
#ifdef __linux__
  cpu_set_t set;  // the cores this process may actually run on (containers, taskset)
  if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0) hw = (unsigned)CPU_COUNT(&set);
#endif
  if (hw == 0) hw = 1;
  // one process per GPU: the launcher says how many processes share this host's cores
  for (const char* name : {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "MV2_COMM_WORLD_LOCAL_SIZE"})
    if (const char* env = std::getenv(name)) {
      const int local = std::atoi(env);
      if (local > 1) hw = std::max(1u, hw / (unsigned)local);
      break;
    }
  // packing needs ~8 threads to outrun the raw copy over the link (profiles/r1_bench_host_pack.jsonl); with fewer
  // cores to itself a process copies the raw rows instead
  return hw < 6 ? 0 : (int)std::min(16u, hw);
}

// ----------------------------------------------------------------------------------------------- row kernels

template <typename T>
static inline uint32_t word_portable(const T* p, int count) {
  uint32_t w = 0;
  for (int b = 0; b < count; ++b) w |= (uint32_t)(p[b] != 0) << b;
  return w;
}

#ifdef QLDPC_X86
__attribute__((target("avx2"))) static void pack_rows_avx2_i32(const int32_t* src, int64_t r0, int64_t r1, int cols,
                                                               int words, uint32_t* dst) {
  const __m256i zero = _mm256_setzero_si256();
  const int full = cols / 32;
  for (int64_t r = r0; r < r1; ++r) {
    const int32_t* p = src + r * cols;
    uint32_t* d = dst + r * words;
    for (int w = 0; w < full; ++w, p += 32) {
      uint32_t word = 0;
      for (int k = 0; k < 4; ++k) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + 8 * k));
        const int eq = _mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpeq_epi32(v, zero)));
        word |= (uint32_t)(~eq & 0xFF) << (8 * k);
      }
      d[w] = word;
    }
    for (int w = full; w < words; ++w, p += 32) d[w] = word_portable(p, std::max(0, std::min(32, cols - 32 * w)));
  }
}

__attribute__((target("avx2"))) static void pack_rows_avx2_u8(const uint8_t* src, int64_t r0, int64_t r1, int cols,
                                                              int words, uint32_t* dst) {
  const __m256i zero = _mm256_setzero_si256();
  const int full = cols / 32;
  for (int64_t r = r0; r < r1; ++r) {
    const uint8_t* p = src + r * cols;
    uint32_t* d = dst + r * words;
    for (int w = 0; w < full; ++w, p += 32) {
      const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
      d[w] = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, zero));
    }
    for (int w = full; w < words; ++w, p += 32) d[w] = word_portable(p, std::max(0, std::min(32, cols - 32 * w)));
  }
}

// one byte per bit from packed words: 8 bits -> 8 bytes through a 64-bit multiply-free spread
__attribute__((target("avx2"))) static void unpack_rows_avx2(const uint32_t* src, int64_t r0, int64_t r1, int cols,
                                                             int words, uint8_t* dst) {
  const __m256i sel = _mm256_setr_epi8(0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3,
                                       3, 3, 3, 3);
  const __m256i bits = _mm256_set1_epi64x((long long)0x8040201008040201ull);
  const __m256i one = _mm256_set1_epi8(1);
  const int full = cols / 32;
  for (int64_t r = r0; r < r1; ++r) {
    const uint32_t* s = src + r * words;
    uint8_t* d = dst + r * cols;
    for (int w = 0; w < full; ++w) {
      const __m256i v = _mm256_shuffle_epi8(_mm256_set1_epi32((int)s[w]), sel);  // byte k of the word in lanes 8k..8k+7
      const __m256i hit = _mm256_cmpeq_epi8(_mm256_and_si256(v, bits), bits);
      _mm256_storeu_si256(reinterpret_cast<__m256i*>(d + 32 * w), _mm256_and_si256(hit, one));
    }
    for (int c = 32 * full; c < cols; ++c) d[c] = (uint8_t)((s[c >> 5] >> (c & 31)) & 1u);
  }
}
#endif

template <typename T>
static void pack_rows_portable(const T* src, int64_t r0, int64_t r1, int cols, int words, uint32_t* dst) {
  for (int64_t r = r0; r < r1; ++r) {
    const T* p = src + r * cols;
    uint32_t* d = dst + r * words;
    for (int w = 0; w < words; ++w) d[w] = word_portable(p + 32 * w, std::max(0, std::min(32, cols - 32 * w)));
  }
}

static void unpack_rows_portable(const uint32_t* src, int64_t r0, int64_t r1, int cols, int words, uint8_t* dst) {
  for (int64_t r = r0; r < r1; ++r) {
    const uint32_t* s = src + r * words;
    uint8_t* d = dst + r * cols;
    for (int c = 0; c < cols; ++c) d[c] = (uint8_t)((s[c >> 5] >> (c & 31)) & 1u);
  }
}

static bool have_avx2() {
#ifdef QLDPC_X86
  static const bool ok = __builtin_cpu_supports("avx2");
  return ok;
#else
  return false;
#endif
}

// ----------------------------------------------------------------------------------------------- thread pool

HostPacker::HostPacker(int threads) : nthreads_(std::max(1, threads)) {
  for (int i = 1; i < nthreads_; ++i) pool_.emplace_back(&HostPacker::worker, this, i);
}

HostPacker::~HostPacker() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_start_.notify_all();
  for (auto& t : pool_) t.join();
}

void HostPacker::worker(int id) {
  uint64_t seen = 0;
  for (;;) {
    const std::function<void(int)>* job = nullptr;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_start_.wait(lk, [&] { return stop_ || generation_ != seen; });
      if (stop_) return;
      seen = generation_;
      job = job_;
    }
    (*job)(id);
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) cv_done_.notify_one();
    }
  }
}

void HostPacker::run(const std::function<void(int)>& job) {
  if (nthreads_ > 1) {
    std::lock_guard<std::mutex> lk(mu_);
    job_ = &job;
    pending_ = nthreads_ - 1;
    ++generation_;
  }
  cv_start_.notify_all();
  job(0);
  if (nthreads_ > 1) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return pending_ == 0; });
  }
}

void HostPacker::pack(const void* src, int elem, int64_t rows, int cols, int words, uint32_t* dst) {
  if (rows <= 0) return;
  const bool avx2 = have_avx2();
  const int T = nthreads_;
  run([&](int id) {
    const int64_t r0 = rows * id / T, r1 = rows * (id + 1) / T;
#ifdef QLDPC_X86
    if (avx2) {
      if (elem == 4) pack_rows_avx2_i32((const int32_t*)src, r0, r1, cols, words, dst);
      else pack_rows_avx2_u8((const uint8_t*)src, r0, r1, cols, words, dst);
      return;
    }
#endif
    if (elem == 4) pack_rows_portable((const int32_t*)src, r0, r1, cols, words, dst);
    else pack_rows_portable((const uint8_t*)src, r0, r1, cols, words, dst);
  });
}

void HostPacker::unpack(const uint32_t* src, int64_t rows, int cols, int words, uint8_t* dst) {
  if (rows <= 0) return;
  const bool avx2 = have_avx2();
  const int T = nthreads_;
  run([&](int id) {
    const int64_t r0 = rows * id / T, r1 = rows * (id + 1) / T;
#ifdef QLDPC_X86
    if (avx2) {
      unpack_rows_avx2(src, r0, r1, cols, words, dst);
      return;
    }
#endif
    unpack_rows_portable(src, r0, r1, cols, words, dst);
  });
}

}  // namespace qldpc
