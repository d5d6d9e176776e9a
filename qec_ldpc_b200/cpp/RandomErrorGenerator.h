// RandomErrorGenerator (QEC_LDPC/RandomErrorGenerator.h:5-45): host-side fixed-weight error patterns, W draws of
// (index, type) with type 0 -> X, 1 -> X and Z, 2 -> Z; duplicates allowed.  Unlike the reference (:24, which records
// mt19937::default_seed) `seed` holds the seed actually used.  The draws use the rejection+modulo mapping of the
// toolchain the reference's results were produced with, so a generator seeded with a results file's "Rand Seed"
// replays that file's patterns.  The Monte-Carlo hot path does not use this class: it generates depolarizing noise on
// the device (qldpc_get_statistics_depolarizing) or replays this stream inside qldpc_get_statistics_weightw.
#pragma once
#include <cstdint>
#include <random>
#include <vector>

class RandomErrorGenerator {
  std::mt19937 _engine;
  uint32_t _numVars;
  uint32_t draw(uint32_t range) {
    for (;;) {
      const uint32_t u = (uint32_t)_engine();
      if (u / range < 0xFFFFFFFFu / range || 0xFFFFFFFFu % range == range - 1) return u % range;
    }
  }

 public:
  unsigned int seed;
  explicit RandomErrorGenerator(int numVars) : _numVars((uint32_t)numVars) {
    std::random_device rd;
    seed = rd();
    _engine.seed(seed);
  }
  RandomErrorGenerator(int numVars, unsigned int seed_) : _engine(seed_), _numVars((uint32_t)numVars), seed(seed_) {}

  void GenerateError(std::vector<int>& xErrors, std::vector<int>& zErrors, int errorWeight) {
    for (int i = 0; i < errorWeight; ++i) {
      const uint32_t index = draw(_numVars);
      const uint32_t error = draw(3u);
      if (error == 0 || error == 1) xErrors[index] = 1;
      if (error == 2 || error == 1) zErrors[index] = 1;
    }
  }
};
