// See host_pack.h.  The inner loops have an AVX-512 form (vptestmd / vptestmb straight into mask registers), an AVX2 form (32 source elements per step) and a portable form, selected at run time;
// all produce the same words.
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
  // Always pack on the host: the packed rows are 1/32 (int) or 1/8 (byte) of the bytes, and the source buffers then
  // cross the host memory bus once instead of once for the DMA engine; with few cores per process (8 ranks on a
  // 32-core box) the call is bound by host memory bandwidth either way (bench.py reports that roofline).
  return (int)std::max(1u, std::min(16u, hw));
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

// AVX-512: 16 ints (or 64 bytes) per test instruction, result lands in a mask register.  The loads are ordinary
// (write-back memory: a non-temporal load is only a hint there) and there is NO software prefetch: measured on the
// B200 hosts, a prefetchnta ahead of the stream costs a third of the aggregate rate as soon as many threads stream at
// once (32 threads: 190 GB/s without, 128 GB/s with; AVX2: 172 GB/s -- tools/bench_host_pack_ranks.py).
__attribute__((target("avx512f,avx512bw"))) static void pack_rows_avx512_i32(const int32_t* src, int64_t r0, int64_t r1,
                                                                             int cols, int words, uint32_t* dst) {
  const int full = cols / 32;
  for (int64_t r = r0; r < r1; ++r) {
    const int32_t* p = src + r * cols;
    uint32_t* d = dst + r * words;
    for (int w = 0; w < full; ++w, p += 32) {
      const __m512i a = _mm512_loadu_si512(p), b = _mm512_loadu_si512(p + 16);
      d[w] = (uint32_t)_mm512_test_epi32_mask(a, a) | ((uint32_t)_mm512_test_epi32_mask(b, b) << 16);
    }
    for (int w = full; w < words; ++w, p += 32) d[w] = word_portable(p, std::max(0, std::min(32, cols - 32 * w)));
  }
}

__attribute__((target("avx512f,avx512bw"))) static void pack_rows_avx512_u8(const uint8_t* src, int64_t r0, int64_t r1,
                                                                            int cols, int words, uint32_t* dst) {
  const int full = cols / 64;  // two words per step
  for (int64_t r = r0; r < r1; ++r) {
    const uint8_t* p = src + r * cols;
    uint32_t* d = dst + r * words;
    int w = 0;
    for (int k = 0; k < full; ++k, p += 64, w += 2) {
      const __m512i a = _mm512_loadu_si512(p);
      const uint64_t m = (uint64_t)_mm512_test_epi8_mask(a, a);
      d[w] = (uint32_t)m;
      d[w + 1] = (uint32_t)(m >> 32);
    }
    for (; w < words; ++w, p += 32) d[w] = word_portable(p, std::max(0, std::min(32, cols - 32 * w)));
  }
}

// streaming read of `bytes` bytes (the packers' access pattern without their arithmetic): the host-memory roofline probe
__attribute__((target("avx512f,avx512bw"))) static uint64_t read_avx512(const uint8_t* p, size_t bytes) {
  __m512i acc = _mm512_setzero_si512();
  size_t i = 0;
  for (; i + 64 <= bytes; i += 64) {
    acc = _mm512_or_si512(acc, _mm512_loadu_si512(p + i));
  }
  return (uint64_t)_mm512_reduce_or_epi64(acc);
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

// QLDPC_HOST_ISA=avx2|portable in the environment caps the instruction set of the packers (measurement knob)
static int isa_cap() {
  static const int cap = [] {
    const char* e = std::getenv("QLDPC_HOST_ISA");
    if (!e) return 2;
    return !std::strcmp(e, "portable") ? 0 : !std::strcmp(e, "avx2") ? 1 : 2;
  }();
  return cap;
}

static bool have_avx512() {
#ifdef QLDPC_X86
  static const bool ok = isa_cap() >= 2 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
  return ok;
#else
  return false;
#endif
}

static bool have_avx2() {
#ifdef QLDPC_X86
  static const bool ok = isa_cap() >= 1 && __builtin_cpu_supports("avx2");
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
  const bool avx2 = have_avx2(), avx512 = have_avx512();
  const int T = nthreads_;
  run([&](int id) {
    const int64_t r0 = rows * id / T, r1 = rows * (id + 1) / T;
#ifdef QLDPC_X86
    if (avx512) {
      if (elem == 4) pack_rows_avx512_i32((const int32_t*)src, r0, r1, cols, words, dst);
      else pack_rows_avx512_u8((const uint8_t*)src, r0, r1, cols, words, dst);
      return;
    }
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

uint64_t HostPacker::read_all(const void* src, size_t bytes) {
  std::vector<uint64_t> part((size_t)nthreads_, 0);
  const int T = nthreads_;
  const bool avx512 = have_avx512();
  run([&](int id) {
    const size_t b0 = bytes / 64 * id / T * 64, b1 = id == T - 1 ? bytes : bytes / 64 * (id + 1) / T * 64;
    const uint8_t* p = (const uint8_t*)src + b0;
    uint64_t acc = 0;
#ifdef QLDPC_X86
    if (avx512) {
      part[(size_t)id] = read_avx512(p, b1 - b0);
      return;
    }
#endif
    for (size_t i = 0; i + 8 <= b1 - b0; i += 8) {
      uint64_t v;
      std::memcpy(&v, p + i, 8);
      acc |= v;
    }
    part[(size_t)id] = acc;
  });
  uint64_t acc = 0;
  for (uint64_t v : part) acc |= v;
  return acc;
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

// ----------------------------------------------------------------------------------------------- weight-W stream

static uint64_t msvc_accept_limit(uint32_t R) {
  // MSVC's uniform_int_distribution<int>(0, R-1) over a 32-bit engine keeps u iff
  //   u / R < 0xFFFFFFFF / R  ||  0xFFFFFFFF % R == R - 1        and returns u % R
  const uint32_t q = 0xFFFFFFFFu / R, r = 0xFFFFFFFFu % R;
  return r == R - 1 ? (uint64_t)1 << 32 : (uint64_t)q * R;
}

WeightWGenerator::WeightWGenerator(uint32_t seed, int n, int weight) : n_(n), weight_(weight) {
  state_[0] = seed;  // std::mt19937 seeding (ISO C++ [rand.eng.mers])
  for (int i = 1; i < 624; ++i) state_[i] = 1812433253u * (state_[i - 1] ^ (state_[i - 1] >> 30)) + (uint32_t)i;
  limit_n_ = msvc_accept_limit((uint32_t)n);
  limit_3_ = msvc_accept_limit(3u);
  magic_n_ = ~(uint64_t)0 / (uint64_t)n + 1;
}

// One MT19937 block: the twist of the 624-word state, then the tempered outputs written to `dst`.  Returns whether
// any output is >= `limit` (the caller's cheap test for "nothing in this block can be rejected").
#ifdef QLDPC_X86
__attribute__((target_clones("avx2", "default")))  // the loops below vectorise; 8 lanes where the CPU has them
#endif
bool WeightWGenerator::block(uint32_t* __restrict dst, uint32_t limit) {
  uint32_t* s = state_;
  auto mix = [](uint32_t hi, uint32_t lo, uint32_t far) {
    const uint32_t y = (hi & 0x80000000u) | (lo & 0x7FFFFFFFu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
  };
  for (int i = 0; i < 227; ++i) s[i] = mix(s[i], s[i + 1], s[i + 397]);
  for (int i = 227; i < 623; ++i) s[i] = mix(s[i], s[i + 1], s[i - 227]);
  s[623] = mix(s[623], s[0], s[396]);
  uint32_t over = 0;
  for (int i = 0; i < 624; ++i) {
    uint32_t y = s[i];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    y ^= y >> 18;
    dst[i] = y;
    over |= (uint32_t)(y >= limit);
  }
  return over != 0;
}

void WeightWGenerator::refill() {
  block(out_, 0xFFFFFFFFu);
  pos_ = 0;
}

// serial part: the next `need` accepted draws of the stream, in order (index, type, index, type, ...)
void WeightWGenerator::produce(uint32_t* __restrict dst, size_t need) {
  // locals: the stores into the draw buffer must not force reloads of the generator's members (dst never aliases
  // them; out_ itself is rewritten by refill() / memcpy below, so it is read through a plain pointer)
  const uint32_t* src = out_;
  const uint64_t lim[2] = {limit_n_, limit_3_};
  size_t have = 0;
  int pos = pos_;
  // both acceptance limits are within n of 2^32; below the smaller one a word is accepted whatever its role
  const uint32_t sure = (uint32_t)std::min<uint64_t>(std::min(lim[0], lim[1]), 0xFFFFFFFFull);
  while (have < need) {
    if (pos == 624) {
      if (need - have >= 624) {  // a whole block is wanted: temper it straight into the draw buffer
        if (!block(dst + have, sure)) {
          have += 624;
          continue;
        }
        std::memcpy(out_, dst + have, sizeof out_);  // some word may be rejected: walk the block below
      } else {
        refill();
      }
      pos = 0;
    }
    // Rejections are rare (n / 2^32 per index draw, 2^-32 per type draw): test the rest of the block at once and
    // copy it when nothing is rejected; otherwise walk it word by word.
    const int take = (int)std::min<size_t>((size_t)(624 - pos), need - have);
    const uint32_t* blk = src + pos;
    const int first_index = (int)(have & 1);  // offset of the first word that is an index draw
    bool rejected = false;
    for (int k = first_index; k < take; k += 2) rejected |= (uint64_t)blk[k] >= lim[0];
    for (int k = first_index ^ 1; k < take; k += 2) rejected |= (uint64_t)blk[k] >= lim[1];
    if (!rejected) {
      std::memcpy(dst + have, blk, (size_t)take * sizeof(uint32_t));
      have += (size_t)take;
      pos += take;
      continue;
    }
    for (int k = 0; k < take && have < need; ++k, ++pos) {
      const uint32_t u = blk[k];
      dst[have] = u;
      have += (uint64_t)u < lim[have & 1];  // a rejected word is overwritten by the next one
    }
  }
  pos_ = pos;
}

// parallel part: frames [f0, f1) of a block whose accepted draws start at `draws`
void WeightWGenerator::map(const uint32_t* draws, int64_t f0, int64_t f1, int words, uint32_t* hx, uint32_t* hz) const {
  const int n = n_, W = weight_;
  const uint64_t magic = magic_n_;
  std::memset(hx + (size_t)f0 * words, 0, (size_t)(f1 - f0) * words * sizeof(uint32_t));
  std::memset(hz + (size_t)f0 * words, 0, (size_t)(f1 - f0) * words * sizeof(uint32_t));
  for (int64_t f = f0; f < f1; ++f) {
    const uint32_t* d = draws + (size_t)f * 2 * W;
    uint32_t* x = hx + (size_t)f * words;
    uint32_t* z = hz + (size_t)f * words;
    for (int i = 0; i < W; ++i) {  // DecoderCPU.h:449-458: type 0 -> X, 1 -> X and Z, 2 -> Z; collisions allowed
      const uint64_t low = magic * d[2 * i];
      const uint32_t index = (uint32_t)(((unsigned __int128)low * (uint32_t)n) >> 64);  // == d[2i] % n
      const uint32_t type = d[2 * i + 1] % 3u;
      const uint32_t bit = 1u << (index & 31);
      if (type != 2) x[index >> 5] |= bit;
      if (type != 0) z[index >> 5] |= bit;
    }
  }
}

void WeightWGenerator::next(int64_t frames, int words, uint32_t* hx, uint32_t* hz, HostPacker* pool) {
  const size_t per_frame = 2 * (size_t)weight_;
  const int T = pool ? pool->threads() : 1;
  if (T < 3 || frames < 4096) {  // too small to split: produce, then map on this thread
    if (draws_.size() < (size_t)frames * per_frame) draws_.resize((size_t)frames * per_frame);
    produce(draws_.data(), (size_t)frames * per_frame);
    map(draws_.data(), 0, frames, words, hx, hz);
    return;
  }
  // Blocks of frames go through a two-stage pipeline: the calling thread produces the draws of block k+1 while the
  // pool's other threads map block k (two draw buffers).
  const int64_t block = 8192;
  const size_t cap = (size_t)block * per_frame;
  if (draws_.size() < 2 * cap) draws_.resize(2 * cap);
  const int64_t nblocks = (frames + block - 1) / block;
  for (int64_t k = 0; k <= nblocks; ++k) {
    const int64_t pf0 = k * block, pf1 = std::min(frames, pf0 + block);              // block produced in this step
    const int64_t mf0 = (k - 1) * block, mf1 = std::min(frames, mf0 + block);        // block mapped in this step
    uint32_t* pbuf = draws_.data() + (size_t)(k & 1) * cap;
    const uint32_t* mbuf = draws_.data() + (size_t)((k - 1) & 1) * cap;
    pool->parallel([&](int id) {
      if (id == 0) {
        if (k < nblocks) produce(pbuf, (size_t)(pf1 - pf0) * per_frame);
      } else if (k >= 1) {
        const int64_t cnt = mf1 - mf0, w = id - 1, W1 = T - 1;
        const int64_t a0 = cnt * w / W1, a1 = cnt * (w + 1) / W1;
        // frame indices are relative to the block: its rows start at frame mf0, its draws at mbuf
        map(mbuf, a0, a1, words, hx + (size_t)mf0 * words, hz + (size_t)mf0 * words);
      }
    });
  }
}

}  // namespace qldpc
