// Host-side marshalling for the host-buffer entry points: the reference's layouts spend one int (or one byte) per bit
// (DecoderGPU.h:136-137,193), i.e. 32x (8x) the information they carry, and the host-to-device link is the
// bottleneck of those calls.  HostPacker turns rows of one-element-per-bit into LSB-first 32-bit words on the host,
// with a small pool of worker threads, so that only the packed rows cross the link.  This is data marshalling only:
// no part of the decode runs on the host.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace qldpc {

class HostPacker {
 public:
  explicit HostPacker(int threads);
  ~HostPacker();
  int threads() const { return nthreads_; }
  // dst[r * words + w] bit b = (src[r * cols + 32 w + b] != 0); elem = bytes per source element (1 or 4)
  void pack(const void* src, int elem, int64_t rows, int cols, int words, uint32_t* dst);
  // dst[r * cols + c] = bit c of src[r * words ..] as one byte
  void unpack(const uint32_t* src, int64_t rows, int cols, int words, uint8_t* dst);
  // streaming read of a host buffer by all threads (OR of its words): the memory-bandwidth probe behind bench.py's
  // host_mem_roofline
  uint64_t read_all(const void* src, size_t bytes);
  // runs job(id) for id = 0 .. threads()-1, id 0 on the calling thread; returns when all are done
  void parallel(const std::function<void(int)>& job) { run(job); }

 private:
  void run(const std::function<void(int)>& job);
  void worker(int id);
  int nthreads_;
  std::vector<std::thread> pool_;
  std::mutex mu_;
  std::condition_variable cv_start_, cv_done_;
  const std::function<void(int)>* job_ = nullptr;
  uint64_t generation_ = 0;
  int pending_ = 0;
  bool stop_ = false;
};

// The reference's fixed-weight error generator (DecoderCPU.h:394-396,446-459): one std::mt19937 stream shared by all
// frames, W x (qubit index uniform in [0,n), Pauli type uniform in {0,1,2}) per frame through MSVC's
// uniform_int_distribution mapping (the one the published results files were made with).  The stream is inherently
// serial; it is split into a serial producer (raw MT19937 words, acceptance test) and a parallel consumer (modular
// reduction, bit setting) so that the serial part is ~2 ns per draw.
class WeightWGenerator {
 public:
  WeightWGenerator(uint32_t seed, int n, int weight);
  // next `frames` patterns, bit-packed rows of `words` 32-bit words: x rows into hx, z rows into hz
  void next(int64_t frames, int words, uint32_t* hx, uint32_t* hz, HostPacker* pool);

 private:
  bool block(uint32_t* dst, uint32_t limit);
  void refill();
  void produce(uint32_t* dst, size_t need);
  void map(const uint32_t* draws, int64_t f0, int64_t f1, int words, uint32_t* hx, uint32_t* hz) const;
  uint32_t state_[624];
  uint32_t out_[624];
  int pos_ = 624;
  int n_, weight_;
  uint64_t limit_n_, limit_3_;  // a raw word is accepted for a range R iff it is below limit_R (MSVC's rejection rule)
  uint64_t magic_n_;            // floor(2^64 / n) + 1: exact u % n by two multiplications
  std::vector<uint32_t> draws_;
};

// default worker count: QLDPC_HOST_THREADS if set, else min(16, hardware threads / local ranks), at least 1
int default_host_threads();

}  // namespace qldpc
