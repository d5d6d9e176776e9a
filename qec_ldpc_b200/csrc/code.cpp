// Host-side code construction: quasi-cyclic CSS expansion, edge tables, code-file IO, logical-check matrix.
// See code.h for the reference citations.
#include "code.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>

namespace qldpc {

// ---------------------------------------------------------------------------------------------------
// GF(2) helpers
// ---------------------------------------------------------------------------------------------------

// Gauss-Jordan elimination; keeps only the non-zero (basis) rows, in order of their pivot column.  Matrices of
// big codes (n ~ 10^4: 10^10..10^11 word operations) are eliminated by a few threads, each owning a block of rows;
// the result does not depend on the thread count (rows are independent once the pivot row is fixed).
namespace {
struct SpinBarrier {
  explicit SpinBarrier(int n) : count(n) {}
  void wait() {
    const int gen = generation.load(std::memory_order_acquire);
    if (arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == count) {
      arrived.store(0, std::memory_order_relaxed);
      generation.store(gen + 1, std::memory_order_release);
    } else {
      int spins = 0;
      while (generation.load(std::memory_order_acquire) == gen)
        if (++spins > 4096) std::this_thread::yield();
    }
  }
  const int count;
  std::atomic<int> arrived{0}, generation{0};
};
}  // namespace

static int rref(BitMatrix& a, std::vector<int>* pivots) {
  const int W = a.words;
  const double work = (double)a.rows * a.rows * W;
  int nthreads = 1;
  if (work > 2e10) nthreads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if (const char* env = std::getenv("QLDPC_HOST_THREADS")) nthreads = std::max(1, std::min(64, std::atoi(env)));
  nthreads = std::min(nthreads, std::max(1, a.rows));
  int rank = 0, col = 0, k0 = 0, wi = 0;
  uint32_t bit = 0;
  bool done = false;
  // picks the next pivot (serial part): leaves col / wi / bit / k0 describing it, or done = true
  auto next_pivot = [&]() {
    for (; col < a.cols && rank < a.rows; ++col) {
      wi = col >> 5;
      bit = 1u << (col & 31);
      int piv = -1;
      for (int r = rank; r < a.rows; ++r)
        if (a.w[(size_t)r * W + wi] & bit) { piv = r; break; }
      if (piv < 0) continue;
      if (piv != rank) std::swap_ranges(a.row(piv), a.row(piv) + W, a.row(rank));
      const uint32_t* p = a.row(rank);
      k0 = 0;  // the pivot row is zero in every earlier pivot column: usually its leading words are all zero
      while (k0 < wi && p[k0] == 0) ++k0;
      if (pivots) pivots->push_back(col);
      return;
    }
    done = true;
  };
  auto eliminate = [&](int r0, int r1) {
    // locals, so that the inner loop vectorises (the captured variables could alias the words being written)
    const int pivot_row = rank, first = k0, word = wi, len = W - k0;
    const uint32_t mask = bit;
    const uint32_t* __restrict p = a.row(pivot_row) + first;
    uint32_t* base = a.w.data();
    for (int r = r0; r < r1; ++r)
      if (r != pivot_row && (base[(size_t)r * W + word] & mask)) {
        uint32_t* __restrict q = base + (size_t)r * W + first;
        for (int k = 0; k < len; ++k) q[k] ^= p[k];
      }
  };
  if (nthreads == 1) {
    for (next_pivot(); !done; next_pivot()) {
      eliminate(0, a.rows);
      ++rank;
      ++col;
    }
  } else {
    SpinBarrier barrier(nthreads);
    auto worker = [&](int t) {
      const int r0 = (int)((long long)a.rows * t / nthreads), r1 = (int)((long long)a.rows * (t + 1) / nthreads);
      for (;;) {
        if (t == 0) next_pivot();
        barrier.wait();  // pivot published
        if (done) return;
        eliminate(r0, r1);
        barrier.wait();  // all rows updated before the next pivot search reads them
        if (t == 0) { ++rank; ++col; }
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto& th : pool) th.join();
  }
  a.rows = rank;
  a.w.resize((size_t)rank * W);
  return rank;
}

int row_reduce(BitMatrix& a) { return rref(a, nullptr); }

BitMatrix null_space(const BitMatrix& a_in) {
  BitMatrix a = a_in;
  std::vector<int> piv;
  const int rank = rref(a, &piv);
  std::vector<char> is_piv(a.cols, 0);
  for (int c : piv) is_piv[c] = 1;
  BitMatrix ns(a.cols - rank, a.cols);
  int k = 0;
  for (int f = 0; f < a.cols; ++f) {
    if (is_piv[f]) continue;
    ns.set(k, f);
    for (int r = 0; r < rank; ++r)
      if (a.get(r, f)) ns.set(k, piv[r]);
    ++k;
  }
  return ns;
}

// ---------------------------------------------------------------------------------------------------
// Exponent matrices (Hagiwara-Imai, quant-ph/0701020): QEC_LDPC_CSS.cu:37-39, :43-65, :67-90
// ---------------------------------------------------------------------------------------------------

static int64_t mod_pow(int64_t b, int64_t e, int64_t P) {
  int64_t r = 1 % P;
  b %= P;
  while (e > 0) {
    if (e & 1) r = r * b % P;
    b = b * b % P;
    e >>= 1;
  }
  return r;
}

void qc_exponents(int J, int K, int L, int P, int sigma, int tau, std::vector<int32_t>& hHC, std::vector<int32_t>& hHD) {
  if (J < 1 || K < 1 || L < 2 || P < 2) throw std::string("QC parameters out of range");
  int inv = 0;
  for (int i = 1; i < P; ++i)
    if ((int64_t)i * sigma % P == 1) { inv = i; break; }
  if (!inv) throw std::string("sigma has no inverse modulo P");
  auto spow = [&](int e) -> int64_t { return e < 0 ? mod_pow(inv, -e, P) : mod_pow(sigma, e, P); };
  hHC.assign((size_t)J * L, 0);
  hHD.assign((size_t)K * L, 0);
  const int half = L / 2;
  for (int j = 0; j < J; ++j)
    for (int l = 0; l < L; ++l)
      hHC[(size_t)j * L + l] = l < half ? (int32_t)spow(l - j) : (int32_t)(P - (tau * spow(j - 1 + l)) % P);
  for (int k = 0; k < K; ++k)
    for (int l = 0; l < L; ++l)
      hHD[(size_t)k * L + l] = l < half ? (int32_t)((tau * spow(l - k - 1)) % P) : (int32_t)(P - spow(k + l));
}

// Closed-form circulant index generation (QEC_LDPC_CSS.cu:99-131 expands the same blocks densely):
// check (b, r) of a side touches variable l*P + (h[b][l] + r) mod P in block column l; variable (l, x) touches
// check b*P + (x - h[b][l]) mod P in block row b.  Ascending l / b is ascending variable / check index, which is
// the order DecoderCPU::InitIndexArrays (DecoderCPU.h:51-63) produces by scanning the dense matrix.
static void tables_from_exponents(SideTables& t, const std::vector<int32_t>& h, int B, int L, int P) {
  const int n = L * P;
  t.m = B * P;
  t.dc = L;
  t.dv = B;
  t.E = t.m * t.dc;
  t.hexp = h;
  t.chk_var.resize(t.E);
  t.var_chk.resize(t.E);
  t.var_edge.resize(t.E);
  for (int b = 0; b < B; ++b)
    for (int r = 0; r < P; ++r)
      for (int l = 0; l < L; ++l) t.chk_var[(size_t)(b * P + r) * L + l] = l * P + (h[(size_t)b * L + l] + r) % P;
  for (int l = 0; l < L; ++l)
    for (int x = 0; x < P; ++x)
      for (int b = 0; b < B; ++b) {
        const int r = ((x - h[(size_t)b * L + l]) % P + P) % P;
        const int e = b * P + r, v = l * P + x;
        t.var_chk[(size_t)v * B + b] = e;
        t.var_edge[(size_t)v * B + b] = e * L + l;
      }
  (void)n;
}

// DecoderCPU::InitIndexArrays, DecoderCPU.h:41-84: ascending scan; degrees taken from the first row / column
// and required to be regular (the reference silently assumes it, :69,:78).
static void tables_from_dense(SideTables& t, const int32_t* pcm, int m, int n) {
  t.m = m;
  t.dc = t.dv = 0;
  for (int v = 0; v < n; ++v) t.dc += pcm[v] != 0;
  for (int e = 0; e < m; ++e) t.dv += pcm[(size_t)e * n] != 0;
  if (t.dc == 0 || t.dv == 0 || (int64_t)m * t.dc != (int64_t)n * t.dv)
    throw std::string("parity-check matrix is not (dc,dv)-regular");
  t.E = m * t.dc;
  t.chk_var.assign(t.E, 0);
  t.var_chk.assign(t.E, 0);
  t.var_edge.assign(t.E, 0);
  t.hexp.clear();
  std::vector<int> cfill(m, 0), vfill(n, 0);
  for (int e = 0; e < m; ++e)
    for (int v = 0; v < n; ++v)
      if (pcm[(size_t)e * n + v]) {
        if (cfill[e] >= t.dc || vfill[v] >= t.dv) throw std::string("parity-check matrix is not (dc,dv)-regular");
        t.chk_var[(size_t)e * t.dc + cfill[e]] = v;
        t.var_chk[(size_t)v * t.dv + vfill[v]] = e;
        t.var_edge[(size_t)v * t.dv + vfill[v]] = e * t.dc + cfill[e];
        ++cfill[e];
        ++vfill[v];
      }
  for (int e = 0; e < m; ++e)
    if (cfill[e] != t.dc) throw std::string("parity-check matrix is not (dc,dv)-regular");
  for (int v = 0; v < n; ++v)
    if (vfill[v] != t.dv) throw std::string("parity-check matrix is not (dc,dv)-regular");
}

static BitMatrix pcm_bits(const SideTables& t, int n) {
  BitMatrix h(t.m, n);
  for (int e = 0; e < t.m; ++e)
    for (int i = 0; i < t.dc; ++i) h.set(e, t.chk_var[(size_t)e * t.dc + i]);
  return h;
}

// Order the (row-reduced) logical rows: x-only, z-only, mixed.
static void classify_logical(Code& c, BitMatrix& rows) {
  const int n = c.n, W = rows.words;
  std::vector<int> cls(rows.rows);
  for (int r = 0; r < rows.rows; ++r) {
    bool hx = false, hz = false;
    for (int col = 0; col < 2 * n; ++col)
      if (rows.get(r, col)) (col < n ? hx : hz) = true;
    cls[r] = hx && hz ? 2 : hz ? 1 : 0;
  }
  BitMatrix out(rows.rows, 2 * n);
  int k = 0;
  c.lx = c.lz = c.lm = 0;
  for (int pass = 0; pass < 3; ++pass)
    for (int r = 0; r < rows.rows; ++r)
      if (cls[r] == pass) {
        std::copy(rows.row(r), rows.row(r) + W, out.row(k++));
        (pass == 0 ? c.lx : pass == 1 ? c.lz : c.lm)++;
      }
  c.logical = out;
}

// Generated logical check: x-part must lie in rowspace(pcmX), z-part in rowspace(pcmZ), i.e. be orthogonal to
// ker(pcmX) resp. ker(pcmZ) -- the same kernel as the file-supplied iMinusP.
static void generate_logical(Code& c) {
  BitMatrix nx = null_space(pcm_bits(c.side[0], c.n));
  BitMatrix nz = null_space(pcm_bits(c.side[1], c.n));
  // The stacked matrix [nx 0; 0 nz] is block diagonal, so its reduced form is the two reduced blocks stacked
  // (x pivots come first); reducing them separately works on rows of half the width.
  row_reduce(nx);
  row_reduce(nz);
  BitMatrix rows(nx.rows + nz.rows, 2 * c.n);
  for (int r = 0; r < nx.rows; ++r)
    for (int col = 0; col < c.n; ++col)
      if (nx.get(r, col)) rows.set(r, col);
  for (int r = 0; r < nz.rows; ++r)
    for (int col = 0; col < c.n; ++col)
      if (nz.get(r, col)) rows.set(nx.rows + r, c.n + col);
  classify_logical(c, rows);
  c.logical_from_file = false;
}

static void logical_from_dense(Code& c, const int32_t* imp) {
  const int w = 2 * c.n;
  BitMatrix raw(w, w);
  for (int r = 0; r < w; ++r)
    for (int col = 0; col < w; ++col)
      if (imp[(size_t)r * w + col] & 1) raw.set(r, col);
  c.iminusp_raw = raw;
  BitMatrix rows = raw;
  row_reduce(rows);
  classify_logical(c, rows);
  c.logical_from_file = true;
}

static void detect_qc(Code& c) {
  c.is_qc = false;
  std::vector<int32_t> hc, hd;
  try {
    qc_exponents(c.J, c.K, c.L, c.P, c.sigma, c.tau, hc, hd);
  } catch (std::string&) {
    return;
  }
  SideTables x, z;
  tables_from_exponents(x, hc, c.J, c.L, c.P);
  tables_from_exponents(z, hd, c.K, c.L, c.P);
  if (x.m == c.side[0].m && z.m == c.side[1].m && x.chk_var == c.side[0].chk_var && z.chk_var == c.side[1].chk_var) {
    c.is_qc = true;
    c.side[0].hexp = hc;
    c.side[1].hexp = hd;
  }
}

Code* code_from_qc(int J, int K, int L, int P, int sigma, int tau) {
  std::vector<int32_t> hc, hd;
  qc_exponents(J, K, L, P, sigma, tau, hc, hd);
  Code* c = new Code;
  c->J = J; c->K = K; c->L = L; c->P = P; c->sigma = sigma; c->tau = tau;
  c->n = L * P;
  tables_from_exponents(c->side[0], hc, J, L, P);
  tables_from_exponents(c->side[1], hd, K, L, P);
  c->is_qc = true;
  generate_logical(*c);
  return c;
}

Code* code_from_dense(int J, int K, int L, int P, int sigma, int tau, const int32_t* pcmX, const int32_t* pcmZ,
                      const int32_t* iMinusP) {
  if (J < 1 || K < 1 || L < 1 || P < 1 || !pcmX || !pcmZ) throw std::string("bad code parameters");
  Code* c = new Code;
  try {
    c->J = J; c->K = K; c->L = L; c->P = P; c->sigma = sigma; c->tau = tau;
    c->n = L * P;  // Quantum_LDPC_Code.h:84
    tables_from_dense(c->side[0], pcmX, J * P, c->n);
    tables_from_dense(c->side[1], pcmZ, K * P, c->n);
    detect_qc(*c);
    if (iMinusP) logical_from_dense(*c, iMinusP);
    else generate_logical(*c);
  } catch (...) {
    delete c;
    throw;
  }
  return c;
}

// Quantum_LDPC_Code::createFromFile, Quantum_LDPC_Code.h:26-80.  Like the reference's getArrayFromString
// (:28-41) each matrix line is read as whitespace-separated ints, at most rows*cols of them, missing entries 0
// (so a file without line 4 yields an all-zero iMinusP, i.e. no frame is ever counted as a logical error).
static void parse_ints(const char* s, const char* end, std::vector<int32_t>& out, size_t want) {
  out.assign(want, 0);
  size_t k = 0;
  while (s < end && k < want) {
    while (s < end && (*s == ' ' || *s == '\t' || *s == '\r')) ++s;
    if (s >= end) break;
    char* q;
    long v = strtol(s, &q, 10);
    if (q == s) break;
    out[k++] = (int32_t)v;
    s = q;
  }
}

Code* code_from_file(const std::string& path) {
  std::ifstream ifs(path.c_str(), std::ios::binary);
  if (!ifs.is_open()) throw std::string("Unable to find code file " + path);  // Quantum_LDPC_Code.h:78
  std::stringstream buf;
  buf << ifs.rdbuf();
  const std::string data = buf.str();
  const char* lines[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  const char* ends[4] = {nullptr, nullptr, nullptr, nullptr};
  const char* p = data.data();
  const char* fin = p + data.size();
  for (int i = 0; i < 4; ++i) {
    lines[i] = p;
    const char* nl = (const char*)memchr(p, '\n', (size_t)(fin - p));
    ends[i] = nl ? nl : fin;
    p = nl ? nl + 1 : fin;
  }
  std::vector<int32_t> prm;
  parse_ints(lines[0], ends[0], prm, 6);
  const int J = prm[0], K = prm[1], L = prm[2], P = prm[3], s = prm[4], t = prm[5];
  if (J < 1 || K < 1 || L < 1 || P < 1) throw std::string("malformed parameter line in code file " + path);
  const size_t n = (size_t)L * P;
  std::vector<int32_t> X, Z, I;
  parse_ints(lines[1], ends[1], X, (size_t)J * P * n);
  parse_ints(lines[2], ends[2], Z, (size_t)K * P * n);
  parse_ints(lines[3], ends[3], I, 4 * n * n);
  return code_from_dense(J, K, L, P, s, t, X.data(), Z.data(), I.data());
}

void code_write_file(const Code& c, const std::string& path) {
  std::string out;
  out.reserve((size_t)8 * c.n * c.n + 1024);
  char tmp[128];
  snprintf(tmp, sizeof tmp, "%d\t%d\t%d\t%d\t%d\t%d\n", c.J, c.K, c.L, c.P, c.sigma, c.tau);
  out += tmp;
  auto emit = [&](const std::vector<int32_t>& m, bool last) {
    for (size_t i = 0; i < m.size(); ++i) {
      out += (char)('0' + (m[i] & 1));
      if (i + 1 < m.size()) out += '\t';
    }
    if (!last) out += '\n';
  };
  for (int s = 0; s < 2; ++s) {
    std::vector<int32_t> d((size_t)c.side[s].m * c.n);
    c.dense_pcm(s, d.data());
    emit(d, false);
  }
  const int w = 2 * c.n;
  std::vector<int32_t> imp((size_t)w * w, 0);
  const BitMatrix& src = c.iminusp_raw.rows ? c.iminusp_raw : c.logical;
  for (int r = 0; r < src.rows && r < w; ++r)
    for (int col = 0; col < w; ++col) imp[(size_t)r * w + col] = src.get(r, col);
  emit(imp, true);
  std::ofstream ofs(path.c_str(), std::ios::binary);
  if (!ofs.is_open()) throw std::string("Unable to write code file " + path);
  ofs.write(out.data(), (std::streamsize)out.size());
}

// ---------------------------------------------------------------------------------------------------
// Code methods
// ---------------------------------------------------------------------------------------------------

std::string Code::name() const {
  // Quantum_LDPC_Code.h:145-150; k is printed as numEqsZ - numEqsX there (not the true dimension) and the
  // reference's results file names depend on that string (main.cu:94), so it is kept.
  char b[160];
  snprintf(b, sizeof b, "[J=%d,K=%d,L=%d,P=%d,s=%d,t=%d][[n=%d,k=%d]]", J, K, L, P, sigma, tau, n,
           side[1].m - side[0].m);
  return b;
}

void Code::dense_pcm(int s, int32_t* out) const {
  const SideTables& t = side[s];
  std::fill(out, out + (size_t)t.m * n, 0);
  for (int e = 0; e < t.m; ++e)
    for (int i = 0; i < t.dc; ++i) out[(size_t)e * n + t.chk_var[(size_t)e * t.dc + i]] = 1;
}

bool Code::is_css() const {
  BitMatrix hz = pcm_bits(side[1], n);
  const SideTables& x = side[0];
  for (int e = 0; e < x.m; ++e)
    for (int f = 0; f < hz.rows; ++f) {
      int acc = 0;
      for (int i = 0; i < x.dc; ++i) acc ^= (int)hz.get(f, x.chk_var[(size_t)e * x.dc + i]);
      if (acc) return false;
    }
  return true;
}

// Quantum_LDPC_Code::GetSyndromeX/Z, Quantum_LDPC_Code.h:94-124 (sum of products mod 2, sparse here).
void Code::syndrome(int s, const int32_t* err, int32_t* syn) const {
  const SideTables& t = side[s];
  for (int e = 0; e < t.m; ++e) {
    int x = 0;
    for (int i = 0; i < t.dc; ++i) x += err[t.chk_var[(size_t)e * t.dc + i]];
    syn[e] = x % 2;
  }
}

// Quantum_LDPC_Code::CheckLogicalError, Quantum_LDPC_Code.h:126-142.
bool Code::check_logical(const int32_t* e) const {
  for (int r = 0; r < logical.rows; ++r) {
    int sum = 0;
    for (int col = 0; col < 2 * n; ++col)
      if (logical.get(r, col)) sum += e[col];
    if (sum % 2 != 0) return true;
  }
  return false;
}

}  // namespace qldpc
