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
//   codeFile may be "qc:J,K,L,P,sigma,tau" to build the code from its parameters instead of reading a file.
#include <sys/stat.h>

#include <chrono>
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
  bool depolarizing = false, haveSeed = false;
  unsigned long long seed = 0;
  int gpus = 1, device = -1;
};

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
                               unsigned long long seed) {
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
        dec.GetStatisticsDepolarizing(hi - lo, p, maxIterations, seed, (unsigned long long)lo, part[g].data());
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

    if (o.depolarizing) {
      std::random_device rd;
      const unsigned long long seed = o.haveSeed ? o.seed : ((unsigned long long)rd() << 32 | rd());
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
