// qec_ldpc -- command-line driver compatible with the reference's run files.
//
// Reference behaviour kept (QEC_LDPC/main.cu:43-118): one argument, an init file holding
//     codeFile  w  W  COUNT  MAX_ITERATIONS  p
// (whitespace separated, QEC_LDPC/init.txt); for every weight w..W one GetStatistics(w, COUNT, p, MAX_ITERATIONS) run is
// appended, in the CodeStatistics text format, to results/<code>_W_<w>_MAX_<it>_p_<p>.txt; progress goes to
// output_log.txt.  Deliberate differences: the exit status is 0 on success and 1 on failure (the reference returns 1 on
// success and exits 0 on errors, main.cu:59,117), results/ is created if missing, and the decoder runs on the GPU.
//
// Extensions (options after the init file):
//   --seed S          fixed seed instead of std::random_device (reproduces a results file from its "Rand Seed")
//   --depolarizing    one run of COUNT frames of depolarizing(p) noise generated on the device (w, W ignored)
//   --gpus N          shard the frames of a --depolarizing run over N devices (global frame ids: same counters for any N)
//   --device D        CUDA device for single-device runs
//   --sweep p0:p1:k   FER-vs-p sweep (what the reference's results/ tree was made for, main.cu:91-104): k depolarizing
//                     points from p0 to p1 (the init file's p is ignored), COUNT frames per batch; every point appends
//                     its record to results/<code>_depolarizing_MAX_<it>_p_<p>.txt in the reference's format and one
//                     line to results/<code>_sweep_MAX_<it>.txt (p, frames, frame errors, FER, Wilson 95% interval, ...)
//   --target-errors E with --sweep: keep adding batches of COUNT frames to a point (one continued global frame stream)
//                     until it has seen E frame errors ...
//   --max-frames M    ... or M frames (default 100 x COUNT).  Batch boundaries do not depend on the device count, so
//                     neither do the stopping point and the counters.
//   codeFile may be "qc:J,K,L,P,sigma,tau" to build the code from its parameters instead of reading a file.
#include <sys/stat.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iostream>
#include <sstream>
#include <thread>

#include "CodeStatistics.h"
#include "DecoderGPU.h"
#include "QC_LDPC_CSS.h"

namespace {

struct Options {
  std::string initFile;
  bool depolarizing = false, haveSeed = false, sweep = false;
  unsigned long long seed = 0;
  int gpus = 1, device = -1;
  double p0 = 0, p1 = 0;
  int points = 0;
  long long targetErrors = 0, maxFrames = 0;
};

// Wilson score interval (95%) of a binomial proportion.
void wilson95(long long bad, long long n, double& lo, double& hi) {
  if (n <= 0) { lo = 0; hi = 1; return; }
  const double z = 1.959963984540054, ph = (double)bad / (double)n, den = 1.0 + z * z / n;
  const double ctr = (ph + z * z / (2.0 * n)) / den;
  const double half = z * std::sqrt(ph * (1.0 - ph) / n + z * z / (4.0 * (double)n * n)) / den;
  lo = std::max(0.0, ctr - half);
  hi = std::min(1.0, ctr + half);
}

Quantum_LDPC_Code loadCode(const std::string& spec) {
  if (spec.compare(0, 3, "qc:") == 0) {
    int v[6];
    if (sscanf(spec.c_str() + 3, "%d,%d,%d,%d,%d,%d", v, v + 1, v + 2, v + 3, v + 4, v + 5) != 6)
      throw std::string("bad code specification " + spec + " (want qc:J,K,L,P,sigma,tau)");
    return QC_LDPC_CSS(v[0], v[1], v[2], v[3], v[4], v[5]);
  }
  return Quantum_LDPC_Code::createFromFile(spec);
}

std::string resultsName(const Quantum_LDPC_Code& code, const std::string& tag, int maxIterations, float p) {
  std::stringstream name;  // main.cu:93-94
  name << "results/" << code << tag << "_MAX_" << maxIterations << "_p_" << p << ".txt";
  return name.str();
}

void append(const std::string& file, const CodeStatistics& stats) {
  std::ofstream out(file, std::ios_base::app);
  if (!out.is_open()) throw std::string("Unable to open results file " + file);
  out << stats << std::endl << std::endl;  // main.cu:102
}

CodeStatistics runDepolarizing(const Quantum_LDPC_Code& code, const Options& o, long long count, float p, int maxIterations,
                               unsigned long long seed, unsigned long long firstFrame = 0, uint64_t* total = nullptr) {
  auto t0 = std::chrono::high_resolution_clock::now();
  const int G = o.gpus;
  std::vector<std::vector<uint64_t>> part(G, std::vector<uint64_t>(QLDPC_NUM_COUNTERS, 0));
  std::vector<std::string> errors(G);
  std::vector<std::thread> workers;
  for (int g = 0; g < G; ++g)
    workers.emplace_back([&, g] {
      try {  // contiguous global frame-id range per device
        const long long lo = count * g / G, hi = count * (g + 1) / G;
        DecoderGPU dec(code, G == 1 ? o.device : g, (int)std::min<long long>(hi - lo > 0 ? hi - lo : 1, 1 << 20));
        dec.GetStatisticsDepolarizing(hi - lo, p, maxIterations, seed, firstFrame + (unsigned long long)lo, part[g].data());
      } catch (std::string& s) {
        errors[g] = s;
      }
    });
  for (auto& w : workers) w.join();
  uint64_t k[QLDPC_NUM_COUNTERS] = {0};
  for (int g = 0; g < G; ++g) {
    if (!errors[g].empty()) throw errors[g];
    for (int i = 0; i < QLDPC_NUM_COUNTERS; ++i) k[i] += part[g][i];
  }
  const long long us =
      std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::high_resolution_clock::now() - t0).count();
  CodeStatistics s = {code, (unsigned)seed, (unsigned)k[QLDPC_C_FRAMES], (unsigned)k[QLDPC_C_XTESTED],
                      (unsigned)k[QLDPC_C_ZTESTED], 0u, (unsigned)k[QLDPC_C_CORRECTED], (unsigned)k[QLDPC_C_SYNX],
                      (unsigned)k[QLDPC_C_SYNZ], (unsigned)k[QLDPC_C_LOGICAL], (unsigned)k[QLDPC_C_CVX],
                      (unsigned)k[QLDPC_C_CVZ], us};
  s.iterationsX = k[QLDPC_C_ITERSX];
  s.iterationsZ = k[QLDPC_C_ITERSZ];
  if (total)
    for (int i = 0; i < QLDPC_NUM_COUNTERS; ++i) total[i] += k[i];
  return s;
}

// One sweep point: batches of `count` frames of one continued frame stream (frame ids from firstFrame) until the
// stopping rule fires.  Returns the summed record (its duration is the sum of the batches).
CodeStatistics runSweepPoint(const Quantum_LDPC_Code& code, const Options& o, long long count, float p, int maxIterations,
                             unsigned long long seed, unsigned long long firstFrame, uint64_t* k) {
  const long long cap = std::min<long long>(o.maxFrames > 0 ? o.maxFrames : (o.targetErrors > 0 ? 100 * count : count),
                                            2000000000ll);  // the record's counters are 32 bits wide
  long long micros = 0;
  for (int i = 0; i < QLDPC_NUM_COUNTERS; ++i) k[i] = 0;
  do {
    const long long done = (long long)k[QLDPC_C_FRAMES], batch = std::min(count, cap - done);
    CodeStatistics b = runDepolarizing(code, o, batch, p, maxIterations, seed, firstFrame + (unsigned long long)done, k);
    micros += b.durationMicroSeconds;
  } while ((long long)k[QLDPC_C_FRAMES] < cap && o.targetErrors > 0 &&
           (long long)(k[QLDPC_C_FRAMES] - k[QLDPC_C_CORRECTED]) < o.targetErrors);
  CodeStatistics s = {code, (unsigned)seed, (unsigned)k[QLDPC_C_FRAMES], (unsigned)k[QLDPC_C_XTESTED],
                      (unsigned)k[QLDPC_C_ZTESTED], 0u, (unsigned)k[QLDPC_C_CORRECTED], (unsigned)k[QLDPC_C_SYNX],
                      (unsigned)k[QLDPC_C_SYNZ], (unsigned)k[QLDPC_C_LOGICAL], (unsigned)k[QLDPC_C_CVX],
                      (unsigned)k[QLDPC_C_CVZ], micros};
  s.iterationsX = k[QLDPC_C_ITERSX];
  s.iterationsZ = k[QLDPC_C_ITERSZ];
  return s;
}

}  // namespace

int main(int argc, char** argv) {
  std::ofstream log("output_log.txt", std::ios::app);
  if (!log.is_open()) {
    std::cerr << "Unable to open output log file" << std::endl;
    return 1;
  }
  std::time_t ts = std::chrono::system_clock::to_time_t(std::chrono::system_clock::now());
  log << std::endl << std::ctime(&ts);

  Options o;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--depolarizing") o.depolarizing = true;
    else if (a == "--seed" && i + 1 < argc) { o.seed = strtoull(argv[++i], nullptr, 10); o.haveSeed = true; }
    else if (a == "--gpus" && i + 1 < argc) o.gpus = std::max(1, atoi(argv[++i]));
    else if (a == "--device" && i + 1 < argc) o.device = atoi(argv[++i]);
    else if (a == "--sweep" && i + 1 < argc) {
      if (sscanf(argv[++i], "%lf:%lf:%d", &o.p0, &o.p1, &o.points) != 3 || o.points < 1 || !(o.p0 >= 0) || !(o.p1 >= o.p0)) {
        log << "--sweep wants p0:p1:k with 0 <= p0 <= p1 and k >= 1" << std::endl;
        return 1;
      }
      o.sweep = o.depolarizing = true;
    }
    else if (a == "--target-errors" && i + 1 < argc) o.targetErrors = atoll(argv[++i]);
    else if (a == "--max-frames" && i + 1 < argc) o.maxFrames = atoll(argv[++i]);
    else if (o.initFile.empty() && a.compare(0, 2, "--") != 0) o.initFile = a;
    else {
      log << "Unknown argument " << a << std::endl;
      return 1;
    }
  }
  if (o.initFile.empty()) {
    log << "Must provide initialization file." << std::endl;  // main.cu:56-60
    return 1;
  }
  std::ifstream init(o.initFile);
  if (!init.is_open()) {
    log << "Unable to open init file \"" << o.initFile << "\". Please make sure the file exists in the current directory."
        << std::endl;
    return 1;
  }
  log << "Initializing run from file " << o.initFile << std::endl;

  try {
    std::string codeFile;
    int w = 0, W = 0, COUNT = 0, MAX_ITERATIONS = 0;
    float p = 0.f;
    init >> codeFile >> w >> W >> COUNT >> MAX_ITERATIONS >> p;  // main.cu:74-89
    if (!init) throw std::string("init file must hold: codeFile w W COUNT MAX_ITERATIONS p");
    Quantum_LDPC_Code code = loadCode(codeFile);
    ::mkdir("results", 0777);

    // The record's "Rand Seed" field is 32 bits wide (CodeStatistics.h:9), and a run must be replayable from it:
    // generated seeds are 32-bit, a given one must fit.
    if (o.haveSeed && o.seed > 0xFFFFFFFFull) throw std::string("--seed must fit in 32 bits (the results file records it as such)");
    if (COUNT < 0) throw std::string("COUNT must not be negative");
    if (o.sweep) {
      std::random_device rd;
      const unsigned long long seed = o.haveSeed ? o.seed : (unsigned long long)rd();
      std::stringstream sname;
      sname << "results/" << code << "_sweep_MAX_" << MAX_ITERATIONS << ".txt";
      std::ofstream summary(sname.str(), std::ios_base::app);
      if (!summary.is_open()) throw std::string("Unable to open sweep summary file " + sname.str());
      summary << "# seed " << seed << " frames-per-batch " << COUNT << " target-errors " << o.targetErrors << " gpus " << o.gpus
              << "\n# p frames frame_errors FER wilson95_lo wilson95_hi logical syndromeX syndromeZ mean_itX mean_itZ" << std::endl;
      log << "sweep seed " << seed << std::endl;
      for (int i = 0; i < o.points; ++i) {
        const float pi = (float)(o.points == 1 ? o.p0 : o.p0 + (o.p1 - o.p0) * i / (o.points - 1));
        uint64_t k[QLDPC_NUM_COUNTERS];
        // every point has its own range of global frame ids (2^40 apart), so points never share a Philox stream
        CodeStatistics stats = runSweepPoint(code, o, COUNT, pi, MAX_ITERATIONS, seed, (unsigned long long)i << 40, k);
        const std::string file = resultsName(code, "_depolarizing", MAX_ITERATIONS, pi);
        std::cout << file << std::endl;
        append(file, stats);
        const long long frames = (long long)k[QLDPC_C_FRAMES], bad = frames - (long long)k[QLDPC_C_CORRECTED];
        double lo, hi;
        wilson95(bad, frames, lo, hi);
        const double f = frames ? (double)frames : 1.0;
        summary << pi << " " << frames << " " << bad << " " << bad / f << " " << lo << " " << hi << " " << k[QLDPC_C_LOGICAL]
                << " " << k[QLDPC_C_SYNX] << " " << k[QLDPC_C_SYNZ] << " " << k[QLDPC_C_ITERSX] / f << " "
                << k[QLDPC_C_ITERSZ] / f << std::endl;
        log << "sweep p=" << pi << ": " << frames << " frames, " << bad << " frame errors, FER " << bad / f << " ["
            << lo << ", " << hi << "]" << std::endl;
      }
    } else if (o.depolarizing) {
      std::random_device rd;
      const unsigned long long seed = o.haveSeed ? o.seed : (unsigned long long)rd();
      log << "depolarizing seed " << seed << std::endl;
      const std::string file = resultsName(code, "_depolarizing", MAX_ITERATIONS, p);
      std::cout << file << std::endl;
      CodeStatistics stats = runDepolarizing(code, o, COUNT, p, MAX_ITERATIONS, seed);
      append(file, stats);
      const double frames = stats.numErrorsTested ? (double)stats.numErrorsTested : 1.0;
      log << "depolarizing p=" << p << ": " << stats.numErrorsTested << " frames on " << o.gpus << " GPU(s), "
          << stats.durationMicroSeconds << " us, FER " << 1.0 - stats.corrected / frames << ", mean iterations X "
          << stats.iterationsX / frames << " Z " << stats.iterationsZ / frames << std::endl;
    } else {
      DecoderGPU decoder(code, o.device);
      for (; w <= W; ++w) {  // main.cu:91-104
        std::stringstream tag;
        tag << "_W_" << w;
        const std::string file = resultsName(code, tag.str(), MAX_ITERATIONS, p);
        std::cout << file << std::endl;
        CodeStatistics stats = o.haveSeed ? decoder.GetStatistics(w, COUNT, p, MAX_ITERATIONS, (unsigned)o.seed)
                                          : decoder.GetStatistics(w, COUNT, p, MAX_ITERATIONS);
        append(file, stats);
      }
    }
  } catch (std::string& s) {  // main.cu:106-112
    log << s << std::endl;
    std::cerr << s << std::endl;
    return 1;
  }
  log << "Run complete." << std::endl;
  return 0;
}
