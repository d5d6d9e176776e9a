/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference decode path (see oracle.c).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may use this. */
#ifndef QLDPC_ORACLE_H
#define QLDPC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_code {
  int J, K, L, P, sigma, tau, n;
  int m[2];        /* number of checks: X side = J*P, Z side = K*P                       */
  int dc[2];       /* check degree (= L for the QC construction)                          */
  int dv[2];       /* variable degree (J / K)                                             */
  int E[2];        /* edges = m*dc = n*dv                                                 */
  int* chk_var[2]; /* [m*dc] neighbours of each check, ascending variable index           */
  int* var_chk[2]; /* [n*dv] neighbours of each variable, ascending check index           */
  int* var_edge[2];/* [n*dv] check-major edge id (e*dc+i) of the k-th edge of a variable  */
  int lrows;       /* logical-check matrix (iMinusP or an equivalent): lrows x 2n, 0/1    */
  uint8_t* lmat;
} oracle_code;

oracle_code* oracle_code_qc(int J, int K, int L, int P, int sigma, int tau);
oracle_code* oracle_code_dense(int J, int K, int L, int P, int sigma, int tau, const int* pcmX, const int* pcmZ);
void oracle_code_free(oracle_code* c);
void oracle_code_info(const oracle_code* c, int out[16]);
void oracle_code_tables(const oracle_code* c, int side, int* chk_var, int* var_chk, int* var_edge);
void oracle_qc_exponents(int J, int K, int L, int P, int sigma, int tau, int* hHC, int* hHD);
void oracle_dense_pcm(const oracle_code* c, int side, int* out);
void oracle_set_logical(oracle_code* c, const int* mat, int rows);

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void oracle_depolarizing_thresholds(float p, uint32_t t[3]);
void oracle_depolarizing(const oracle_code* c, uint64_t seed, uint64_t frame, float p, uint8_t* xerr, uint8_t* zerr);

void oracle_depolarizing_bulk(const oracle_code* c, uint64_t seed, uint64_t first_frame, int nframes, float p,
                              uint8_t* xerr, uint8_t* zerr);

void oracle_syndrome(const oracle_code* c, int side, const uint8_t* err, uint8_t* syn);
int oracle_bp(const oracle_code* c, int side, const uint8_t* syn, float errorProbability, int maxIterations, float* q,
              float* r, float* q_trace, float* r_trace, int trace_cap);
int oracle_decode(const oracle_code* c, const uint8_t* synX, const uint8_t* synZ, float errorProbability,
                  int maxIterations, uint8_t* outX, uint8_t* outZ, int iters[2]);
int oracle_check_logical(const oracle_code* c, const uint8_t* err2n);

void oracle_weightw_stream(uint32_t seed, int W, int n, int nframes, uint8_t* xerr, uint8_t* zerr);
double oracle_run_frames(const oracle_code* c, const uint8_t* xerr, const uint8_t* zerr, int nframes,
                         float errorProbability, int maxIterations, int nthreads, uint64_t counters[12],
                         uint8_t* flags, uint8_t* iters, uint8_t* outX, uint8_t* outZ);
double oracle_run_depolarizing(const oracle_code* c, uint64_t seed, uint64_t first_frame, int nframes, float p,
                               int maxIterations, int nthreads, uint64_t counters[12], uint8_t* flags, uint8_t* iters);
void oracle_get_statistics_weightw(const oracle_code* c, int W, int count, float errorProbability, int maxIterations,
                                   uint32_t seed, int nthreads, uint64_t counters[12]);
int oracle_max_threads(void);

/* counters[12] layout shared with the product's C ABI (include/qldpc_b200.h) */
enum { OC_FRAMES = 0, OC_XTESTED, OC_ZTESTED, OC_CORRECTED, OC_SYNX, OC_SYNZ, OC_LOGICAL, OC_CVX, OC_CVZ, OC_ITERSX,
       OC_ITERSZ, OC_NANFRAMES };

#ifdef __cplusplus
}
#endif
#endif
