/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference decode path of cantwellc/QEC_LDPC.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / `--impl reference` legs may load it.  The product (libqldpc_b200.so) has no CPU path and
 * fails loudly without a GPU.
 *
 * Parity is PINNED: the restatement is validated (tests/test_oracle.py) against
 *   (a) the unmodified reference compiled here from /root/reference into oracle/_ref/ (oracle/ref_harness.cpp):
 *       per-iteration messages, per-frame decisions/flags on depolarizing and weight-W patterns, and
 *   (b) the reference's published results files K1..K5, bit-exact counters, and
 *   (c) the committed golden fixtures under tests/golden/ (made by tests/golden/make_golden.py).
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference/QEC_LDPC/).
 * The restatement works on sparse edge arrays instead of the reference's dense numVars x numEqs float
 * arrays; floating-point operations and their ORDER are exactly the reference's, so results are bit-identical.
 * Built with -ffp-contract=off (oracle/Makefile).
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------
 * Code construction
 * ---------------------------------------------------------------------------------------------- */

static int powmod_signed(int base, int invbase, int e, int P) {
  /* repeated multiplication, negative powers through the inverse: QEC_LDPC_CSS.cu:51-52,58-59 */
  long t = 1;
  int i;
  if (e < 0) for (i = 0; i < -e; ++i) t = (t * invbase) % P;
  else for (i = 0; i < e; ++i) t = (t * base) % P;
  return (int)t;
}

/* Hagiwara-Imai exponent matrices.  QEC_LDPC_CSS.cu:37-39 (invSigma), :43-65 (hHC), :67-90 (hHD). */
void oracle_qc_exponents(int J, int K, int L, int P, int sigma, int tau, int* hHC, int* hHD) {
  int inv = 1, j, k, l;
  while ((inv * sigma) % P != 1) ++inv; /* first element of Z_P^* with inv*sigma == 1, :38 */
  for (j = 0; j < J; ++j)
    for (l = 0; l < L; ++l) {
      int t;
      if (l < L / 2) t = powmod_signed(sigma, inv, -j + l, P);                                   /* :48-53 */
      else t = P - (int)(((long)tau * powmod_signed(sigma, inv, j - 1 + l, P)) % P);            /* :56-61 */
      hHC[j * L + l] = t;
    }
  for (k = 0; k < K; ++k)
    for (l = 0; l < L; ++l) {
      int t;
      if (l < L / 2) t = (int)(((long)tau * powmod_signed(sigma, inv, -k - 1 + l, P)) % P);      /* :72-78 */
      else t = P - powmod_signed(sigma, inv, k + l, P);                                          /* :81-86 */
      hHD[k * L + l] = t;
    }
}

/* Adjacency tables by an ascending scan of the dense matrix: DecoderCPU.h:41-84 (InitIndexArrays).
 * Like the reference (:69,:78) the degrees are taken from row 0 / column 0 and assumed regular. */
static int tables_from_dense(oracle_code* c, int side, const int* pcm) {
  const int m = c->m[side], n = c->n;
  int e, v, dc = 0, dv = 0;
  for (v = 0; v < n; ++v) dc += pcm[v] != 0;
  for (e = 0; e < m; ++e) dv += pcm[(size_t)e * n] != 0;
  c->dc[side] = dc;
  c->dv[side] = dv;
  c->E[side] = m * dc;
  if (m * dc != n * dv) return -1;
  c->chk_var[side] = (int*)malloc(sizeof(int) * (size_t)m * dc);
  c->var_chk[side] = (int*)malloc(sizeof(int) * (size_t)n * dv);
  c->var_edge[side] = (int*)malloc(sizeof(int) * (size_t)n * dv);
  int* cfill = (int*)calloc((size_t)m, sizeof(int));
  int* vfill = (int*)calloc((size_t)n, sizeof(int));
  int ok = 0;
  for (e = 0; e < m; ++e)
    for (v = 0; v < n; ++v)
      if (pcm[(size_t)e * n + v]) {
        if (cfill[e] >= dc || vfill[v] >= dv) { ok = -1; continue; }
        c->chk_var[side][e * dc + cfill[e]] = v;
        c->var_chk[side][v * dv + vfill[v]] = e;
        c->var_edge[side][v * dv + vfill[v]] = e * dc + cfill[e];
        cfill[e]++;
        vfill[v]++;
      }
  for (e = 0; e < m; ++e) if (cfill[e] != dc) ok = -1;
  for (v = 0; v < n; ++v) if (vfill[v] != dv) ok = -1;
  free(cfill);
  free(vfill);
  return ok;
}

/* Dense parity-check matrix of one side from the adjacency tables. */
void oracle_dense_pcm(const oracle_code* c, int side, int* out) {
  const int m = c->m[side], n = c->n, dc = c->dc[side];
  memset(out, 0, sizeof(int) * (size_t)m * n);
  for (int e = 0; e < m; ++e)
    for (int i = 0; i < dc; ++i) out[(size_t)e * n + c->chk_var[side][e * dc + i]] = 1;
}

/* Quantum_LDPC_Code ctor, Quantum_LDPC_Code.h:82-88: n = L*P, numEqsX = J*P, numEqsZ = K*P. */
oracle_code* oracle_code_dense(int J, int K, int L, int P, int sigma, int tau, const int* pcmX, const int* pcmZ) {
  oracle_code* c = (oracle_code*)calloc(1, sizeof *c);
  c->J = J; c->K = K; c->L = L; c->P = P; c->sigma = sigma; c->tau = tau;
  c->n = L * P;
  c->m[0] = J * P;
  c->m[1] = K * P;
  if (tables_from_dense(c, 0, pcmX) || tables_from_dense(c, 1, pcmZ)) { oracle_code_free(c); return NULL; }
  return c;
}

/* Circulant expansion, QEC_LDPC_CSS.cu:99-131: row `row` of H has a one in column
 * (h[row/P][cl] + row%P) % P + cl*P for every block column cl. */
oracle_code* oracle_code_qc(int J, int K, int L, int P, int sigma, int tau) {
  const int n = L * P;
  int* hHC = (int*)malloc(sizeof(int) * J * L);
  int* hHD = (int*)malloc(sizeof(int) * K * L);
  oracle_qc_exponents(J, K, L, P, sigma, tau, hHC, hHD);
  int* X = (int*)calloc((size_t)J * P * n, sizeof(int));
  int* Z = (int*)calloc((size_t)K * P * n, sizeof(int));
  for (int row = 0; row < J * P; ++row)
    for (int cl = 0; cl < L; ++cl) X[(size_t)row * n + (hHC[(row / P) * L + cl] + row % P) % P + cl * P] = 1;
  for (int row = 0; row < K * P; ++row)
    for (int cl = 0; cl < L; ++cl) Z[(size_t)row * n + (hHD[(row / P) * L + cl] + row % P) % P + cl * P] = 1;
  oracle_code* c = oracle_code_dense(J, K, L, P, sigma, tau, X, Z);
  free(hHC); free(hHD); free(X); free(Z);
  return c;
}

void oracle_code_free(oracle_code* c) {
  if (!c) return;
  for (int s = 0; s < 2; ++s) { free(c->chk_var[s]); free(c->var_chk[s]); free(c->var_edge[s]); }
  free(c->lmat);
  free(c);
}

void oracle_code_info(const oracle_code* c, int out[16]) {
  int v[16] = {c->J, c->K, c->L, c->P, c->sigma, c->tau, c->n, c->m[0], c->m[1], c->dc[0], c->dc[1],
               c->dv[0], c->dv[1], c->E[0], c->E[1], c->lrows};
  memcpy(out, v, sizeof v);
}

void oracle_code_tables(const oracle_code* c, int side, int* chk_var, int* var_chk, int* var_edge) {
  if (chk_var) memcpy(chk_var, c->chk_var[side], sizeof(int) * (size_t)c->E[side]);
  if (var_chk) memcpy(var_chk, c->var_chk[side], sizeof(int) * (size_t)c->E[side]);
  if (var_edge) memcpy(var_edge, c->var_edge[side], sizeof(int) * (size_t)c->E[side]);
}

/* iMinusP (Quantum_LDPC_Code.h:21, file line 4) or any matrix with the same kernel: rows x 2n, 0/1. */
void oracle_set_logical(oracle_code* c, const int* mat, int rows) {
  free(c->lmat);
  c->lrows = rows;
  c->lmat = (uint8_t*)malloc((size_t)rows * 2 * c->n);
  for (size_t i = 0; i < (size_t)rows * 2 * c->n; ++i) c->lmat[i] = (uint8_t)(mat[i] & 1);
}

/* ------------------------------------------------------------------------------------------------
 * Depolarizing error generator (north-star replacement for the reference's weight-W generator):
 * counter-based Philox4x32-10 (Salmon et al., SC'11; the published constants), integer thresholds only,
 * so the CUDA kernel and this restatement agree bit for bit.
 *   key     = (seed lo, seed hi)
 *   counter = (frame lo, frame hi, q >> 2, 0x51454331)          one call serves qubits 4b..4b+3
 *   draw r  = out[q & 3];   T = floor(p * 2^32), t1 = T/3, t2 = 2T/3
 *   r < t1 -> X   (type 0),  t1 <= r < t2 -> Y (type 1: X and Z),  t2 <= r < T -> Z (type 2)
 * Type -> bit mapping as DecoderCPU.h:456-457.
 * ---------------------------------------------------------------------------------------------- */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; ++round) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void oracle_depolarizing_thresholds(float p, uint32_t t[3]) {
  double s = (double)p * 4294967296.0;
  uint64_t T = s <= 0 ? 0 : s >= 4294967295.0 ? 4294967295ull : (uint64_t)s;
  t[0] = (uint32_t)(T / 3);
  t[1] = (uint32_t)(2 * T / 3);
  t[2] = (uint32_t)T;
}

void oracle_depolarizing(const oracle_code* c, uint64_t seed, uint64_t frame, float p, uint8_t* xerr, uint8_t* zerr) {
  uint32_t t[3], key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  oracle_depolarizing_thresholds(p, t);
  for (int b = 0; 4 * b < c->n; ++b) {
    uint32_t ctr[4] = {(uint32_t)frame, (uint32_t)(frame >> 32), (uint32_t)b, 0x51454331u}, out[4];
    oracle_philox4x32_10(ctr, key, out);
    for (int w = 0; w < 4 && 4 * b + w < c->n; ++w) {
      uint32_t r = out[w];
      int type = r < t[0] ? 0 : r < t[1] ? 1 : r < t[2] ? 2 : -1;
      xerr[4 * b + w] = (uint8_t)(type == 0 || type == 1);
      zerr[4 * b + w] = (uint8_t)(type == 2 || type == 1);
    }
  }
}

/* [nframes x n] byte patterns for frames first_frame .. first_frame+nframes-1 (OpenMP over frames). */
void oracle_depolarizing_bulk(const oracle_code* c, uint64_t seed, uint64_t first_frame, int nframes, float p,
                              uint8_t* xerr, uint8_t* zerr) {
#pragma omp parallel for schedule(static)
  for (int f = 0; f < nframes; ++f)
    oracle_depolarizing(c, seed, first_frame + (uint64_t)f, p, xerr + (size_t)f * c->n, zerr + (size_t)f * c->n);
}

/* ------------------------------------------------------------------------------------------------
 * Weight-W compatibility generator: std::mt19937 + MSVC's uniform_int_distribution mapping, drawn in the
 * order index, type, W times per frame, frames in sequence.  DecoderCPU.h:394-396, :449-458.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint32_t s[624]; int i; } mt19937_t;
static void mt_seed(mt19937_t* g, uint32_t seed) {
  g->s[0] = seed;
  for (int i = 1; i < 624; ++i) g->s[i] = 1812433253u * (g->s[i - 1] ^ (g->s[i - 1] >> 30)) + (uint32_t)i;
  g->i = 624;
}
static uint32_t mt_next(mt19937_t* g) {
  if (g->i >= 624) {
    for (int k = 0; k < 624; ++k) {
      uint32_t y = (g->s[k] & 0x80000000u) | (g->s[(k + 1) % 624] & 0x7FFFFFFFu);
      g->s[k] = g->s[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
    }
    g->i = 0;
  }
  uint32_t y = g->s[g->i++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9D2C5680u;
  y ^= (y << 15) & 0xEFC60000u;
  y ^= y >> 18;
  return y;
}
static uint32_t msvc_uniform(mt19937_t* g, uint32_t R) { /* rejection + modulo */
  for (;;) {
    uint32_t u = mt_next(g);
    if (u / R < 0xFFFFFFFFu / R || 0xFFFFFFFFu % R == R - 1) return u % R;
  }
}
void oracle_weightw_stream(uint32_t seed, int W, int n, int nframes, uint8_t* xerr, uint8_t* zerr) {
  mt19937_t g;
  mt_seed(&g, seed);
  memset(xerr, 0, (size_t)nframes * n);
  memset(zerr, 0, (size_t)nframes * n);
  for (int f = 0; f < nframes; ++f)
    for (int i = 0; i < W; ++i) {
      uint32_t index = msvc_uniform(&g, (uint32_t)n); /* :451 */
      uint32_t error = msvc_uniform(&g, 3u);          /* :453 */
      if (error == 0 || error == 1) xerr[(size_t)f * n + index] = 1; /* :456 */
      if (error == 2 || error == 1) zerr[(size_t)f * n + index] = 1; /* :457 */
    }
}

/* ------------------------------------------------------------------------------------------------
 * Syndrome, BP, decode tail, logical check
 * ---------------------------------------------------------------------------------------------- */

/* s = H e mod 2.  Quantum_LDPC_Code.h:94-108 (X), :110-124 (Z): dense MACs there, the dc ones of each row here. */
void oracle_syndrome(const oracle_code* c, int side, const uint8_t* err, uint8_t* syn) {
  const int m = c->m[side], dc = c->dc[side];
  for (int e = 0; e < m; ++e) {
    int x = 0;
    for (int i = 0; i < dc; ++i) x += err[c->chk_var[side][e * dc + i]];
    syn[e] = (uint8_t)(x % 2);
  }
}

/* DecoderCPU.h:231-246.  Non-edge entries of the reference's dense array are 0 and skipped there (:238). */
static int check_convergence(const float* q, int E, float high, float low) {
  for (int i = 0; i < E; ++i)
    if (q[i] != 0.0f) {
      if (q[i] > low && q[i] < high) return 0;
    }
  return 1;
}

/* One side of BeliefPropogation, DecoderCPU.h:249-292, with EqNodeUpdate :150-186 and VarNodeUpdate :188-229.
 * q[edge] = variable->check message (reference: varNodes[var*numEqs+eq]),
 * r[edge] = check->variable message (reference: eqNodes[eq*numVars+var]); edge = e*dc + i, check-major.
 * Optional traces receive q and r after every iteration (first trace_cap iterations).
 * Returns the number of iterations executed. */
int oracle_bp(const oracle_code* c, int side, const uint8_t* syn, float errorProbability, int maxIterations, float* q,
              float* r, float* q_trace, float* r_trace, int trace_cap) {
  const int m = c->m[side], n = c->n, dc = c->dc[side], dv = c->dv[side], E = c->E[side];
  const int* var_edge = c->var_edge[side];
  float p = 2.0f / 3.0f * errorProbability; /* :259 */
  float high = 0.99f, low = 0.01f;          /* :260-261 */
  for (int i = 0; i < E; ++i) q[i] = p;     /* :265-267 (InitVarNodes :135-148) */
  int N = maxIterations, converge = 0, it = 0;
  for (int nn = 0; nn < N; ++nn) { /* :280 */
    if (converge) break;           /* :282 */
    /* EqNodeUpdate :160-185 */
    for (int e = 0; e < m; ++e)
      for (int i = 0; i < dc; ++i) {
        float product = 1.0f;
        for (int k = 0; k < dc; ++k) {
          if (k == i) continue;
          float value = q[e * dc + k];
          product *= (1.0f - 2.0f * value); /* :175 */
        }
        if (syn[e]) r[e * dc + i] = 0.5 * (1.0f + product); /* :179, double literal as in the reference */
        else r[e * dc + i] = 0.5f * (1.0f - product);       /* :182 */
      }
    /* VarNodeUpdate :195-228 */
    int last = nn == N - 1; /* :284 */
    for (int v = 0; v < n; ++v)
      for (int j = 0; j < dv; ++j) {
        float prodP = p;               /* :209 */
        float prodOneMinusP = 1.0f - p; /* :210 */
        for (int k = 0; k < dv; ++k) {
          if (j == k && !last) continue; /* :215 */
          float pk = r[var_edge[v * dv + k]];
          prodOneMinusP *= (1.0f - pk); /* :220 */
          prodP *= pk;                  /* :221 */
        }
        q[var_edge[v * dv + j]] = prodP / (prodOneMinusP + prodP); /* :223 */
      }
    if (it < trace_cap) {
      if (q_trace) memcpy(q_trace + (size_t)it * E, q, sizeof(float) * E);
      if (r_trace) memcpy(r_trace + (size_t)it * E, r, sizeof(float) * E);
    }
    ++it;
    if (nn % 10 == 0) converge = check_convergence(q, E, high, low); /* :287-290 */
  }
  return it;
}

/* Decode, DecoderCPU.h:317-390.  Returns the ErrorCode bitmask of Decoder.h:14-23; bit 6 (64) is an extra
 * diagnostic = final state holds a NaN (not part of the reference's code). */
int oracle_decode(const oracle_code* c, const uint8_t* synX, const uint8_t* synZ, float errorProbability,
                  int maxIterations, uint8_t* outX, uint8_t* outZ, int iters[2]) {
  int code = 0;
  const uint8_t* syn[2] = {synX, synZ};
  uint8_t* out[2] = {outX, outZ};
  for (int side = 0; side < 2; ++side) {
    const int n = c->n, m = c->m[side], dv = c->dv[side], E = c->E[side];
    float* q = (float*)malloc(sizeof(float) * E);
    float* r = (float*)malloc(sizeof(float) * E);
    uint8_t* s2 = (uint8_t*)malloc((size_t)m);
    int it = oracle_bp(c, side, syn[side], errorProbability, maxIterations, q, r, NULL, NULL, 0);
    if (iters) iters[side] = it;
    for (int v = 0; v < n; ++v) { /* :354-373: any edge slot >= 0.5f */
      uint8_t bit = 0;
      for (int k = 0; k < dv; ++k)
        if (q[c->var_edge[side][v * dv + k]] >= 0.5f) { bit = 1; break; }
      out[side][v] = bit;
    }
    if (!check_convergence(q, E, 0.99f, 0.01f)) code |= side == 0 ? 4 : 8; /* :375-378 */
    oracle_syndrome(c, side, out[side], s2);                                 /* :380-384 */
    if (memcmp(s2, syn[side], (size_t)m) != 0) code |= side == 0 ? 1 : 2;
    for (int i = 0; i < E; ++i) if (q[i] != q[i]) { code |= 64; break; }
    free(q); free(r); free(s2);
  }
  return code;
}

/* CheckLogicalError, Quantum_LDPC_Code.h:126-142: any odd row of lmat * e. */
int oracle_check_logical(const oracle_code* c, const uint8_t* err2n) {
  const int w = 2 * c->n;
  for (int i = 0; i < c->lrows; ++i) {
    int sum = 0;
    const uint8_t* row = c->lmat + (size_t)i * w;
    for (int j = 0; j < w; ++j) sum += row[j] & err2n[j];
    if (sum % 2 != 0) return 1;
  }
  return 0;
}

/* Per-frame bookkeeping of GetStatistics, DecoderCPU.h:461-521, on supplied error patterns
 * ([nframes x n] bytes).  flags[f] = ErrorCode | 16 logical | 32 corrected | 64 NaN; iters[2f], iters[2f+1]. */
static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void frame_stats(const oracle_code* c, const uint8_t* xe, const uint8_t* ze, float ep, int maxit,
                        uint64_t k[12], uint8_t* flag, uint8_t* it2, uint8_t* ox, uint8_t* oz, uint8_t* scratch) {
  const int n = c->n;
  uint8_t* sx = scratch;
  uint8_t* sz = sx + c->m[0];
  uint8_t* xd = sz + c->m[1];
  uint8_t* zd = xd + n;
  uint8_t* res = zd + n;
  int anyx = 0, anyz = 0, iters[2];
  for (int i = 0; i < n; ++i) { anyx |= xe[i]; anyz |= ze[i]; }
  oracle_syndrome(c, 0, xe, sx); /* :461 */
  oracle_syndrome(c, 1, ze, sz); /* :462 */
  k[OC_FRAMES]++;
  k[OC_XTESTED] += anyx != 0; /* :464-473 */
  k[OC_ZTESTED] += anyz != 0;
  int ec = oracle_decode(c, sx, sz, ep, maxit, xd, zd, iters); /* :477 */
  int fl = ec;
  int dEX = ec & 1, dEZ = ec & 2;
  k[OC_SYNX] += dEX != 0; /* :480-489 */
  k[OC_SYNZ] += dEZ != 0;
  if (!(dEX || dEZ)) { /* :492-509 */
    for (int i = 0; i < n; ++i) {
      res[i] = (uint8_t)((xe[i] + xd[i]) % 2);
      res[n + i] = (uint8_t)((ze[i] + zd[i]) % 2);
    }
    if (oracle_check_logical(c, res)) { k[OC_LOGICAL]++; fl |= 16; } else { k[OC_CORRECTED]++; fl |= 32; }
  }
  if (ec & 4) k[OC_CVX]++; /* :514-521 */
  if (ec & 8) k[OC_CVZ]++;
  k[OC_ITERSX] += (uint64_t)iters[0];
  k[OC_ITERSZ] += (uint64_t)iters[1];
  if (ec & 64) k[OC_NANFRAMES]++;
  if (flag) *flag = (uint8_t)fl;
  if (it2) { it2[0] = (uint8_t)(iters[0] > 255 ? 255 : iters[0]); it2[1] = (uint8_t)(iters[1] > 255 ? 255 : iters[1]); }
  if (ox) memcpy(ox, xd, (size_t)n);
  if (oz) memcpy(oz, zd, (size_t)n);
}

double oracle_run_frames(const oracle_code* c, const uint8_t* xerr, const uint8_t* zerr, int nframes,
                         float errorProbability, int maxIterations, int nthreads, uint64_t counters[12],
                         uint8_t* flags, uint8_t* iters, uint8_t* outX, uint8_t* outZ) {
  const int n = c->n;
  uint64_t tot[12] = {0};
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
  double t0 = now_s();
#pragma omp parallel
  {
    uint64_t k[12] = {0};
    uint8_t* scratch = (uint8_t*)malloc((size_t)c->m[0] + c->m[1] + 4 * (size_t)n);
#pragma omp for schedule(dynamic, 16)
    for (int f = 0; f < nframes; ++f)
      frame_stats(c, xerr + (size_t)f * n, zerr + (size_t)f * n, errorProbability, maxIterations, k,
                  flags ? flags + f : NULL, iters ? iters + 2 * (size_t)f : NULL, outX ? outX + (size_t)f * n : NULL,
                  outZ ? outZ + (size_t)f * n : NULL, scratch);
    free(scratch);
#pragma omp critical
    for (int i = 0; i < 12; ++i) tot[i] += k[i];
  }
  double t1 = now_s();
  memcpy(counters, tot, sizeof tot);
  return t1 - t0;
}

double oracle_run_depolarizing(const oracle_code* c, uint64_t seed, uint64_t first_frame, int nframes, float p,
                               int maxIterations, int nthreads, uint64_t counters[12], uint8_t* flags, uint8_t* iters) {
  const int n = c->n;
  uint64_t tot[12] = {0};
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
  double t0 = now_s();
#pragma omp parallel
  {
    uint64_t k[12] = {0};
    uint8_t* scratch = (uint8_t*)malloc((size_t)c->m[0] + c->m[1] + 6 * (size_t)n);
    uint8_t* xe = scratch + c->m[0] + c->m[1] + 4 * (size_t)n;
    uint8_t* ze = xe + n;
#pragma omp for schedule(dynamic, 16)
    for (int f = 0; f < nframes; ++f) {
      oracle_depolarizing(c, seed, first_frame + (uint64_t)f, p, xe, ze);
      frame_stats(c, xe, ze, p, maxIterations, k, flags ? flags + f : NULL, iters ? iters + 2 * (size_t)f : NULL, NULL,
                  NULL, scratch);
    }
    free(scratch);
#pragma omp critical
    for (int i = 0; i < 12; ++i) tot[i] += k[i];
  }
  double t1 = now_s();
  memcpy(counters, tot, sizeof tot);
  return t1 - t0;
}

/* GetStatistics(W, COUNT, p, MAXIT, seed), DecoderCPU.h:392-530: count = COUNT / nThreads frames per thread
 * (:426), i.e. (COUNT / nThreads) * nThreads frames in total, drawn from one sequential stream (:448-459). */
void oracle_get_statistics_weightw(const oracle_code* c, int W, int count, float errorProbability, int maxIterations,
                                   uint32_t seed, int nthreads, uint64_t counters[12]) {
  if (nthreads <= 0) nthreads = oracle_max_threads();
  int total = (count / nthreads) * nthreads;
  uint8_t* xe = (uint8_t*)malloc((size_t)total * c->n + 1);
  uint8_t* ze = (uint8_t*)malloc((size_t)total * c->n + 1);
  oracle_weightw_stream(seed, W, c->n, total, xe, ze);
  oracle_run_frames(c, xe, ze, total, errorProbability, maxIterations, nthreads, counters, NULL, NULL, NULL, NULL);
  free(xe);
  free(ze);
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
