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

// default worker count: QLDPC_HOST_THREADS if set, else min(16, hardware threads / local ranks), 0 if that is below 6
int default_host_threads();

}  // namespace qldpc
