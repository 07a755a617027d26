// mfit_chain.h -- device-side chain state shared by chain_kernels.cu and api.cu
#ifndef MFIT_CHAIN_H
#define MFIT_CHAIN_H
#include <stdint.h>

#define MFIT_MAX_P 96
#define MFIT_MODE_ADAPTIVE 0      /* walker-adaptive-steps-full :862-942 */
#define MFIT_MODE_MANY 1          /* walker-many-steps :849-853 */
#define MFIT_MODE_SAMPLE_REGION 2 /* walker-sample-region :949-969 */

// one dataset as the chain kernels see it
struct MfitDsDev {
  const double* partial;  // [n_seg][stride] per-segment sums written by the eval kernel
  int n_seg;
  int ll_kind;
  double cst;             // theta-independent part of the log-likelihood
  const double* cst_pc;   // per-chain constants instead (walker sets with per-walker data), or null
  int owned;              // this rank evaluates the dataset
  int prior_kind;         // 0 flat, 1 bounds, 2 user kernel (lp written by an NVRTC kernel)
  int nb;
  const int* b_idx;
  const double* b_lo;
  const double* b_hi;
  int xrank, xslot;       // exchange location when datasets are sharded over ranks
};

// struct-of-arrays chain state; every per-chain array has `stride` (W rounded up to 128) entries
struct MfitChains {
  int W, P, D;
  long long stride;
  int max_seg;          // largest n_seg over this rank's data sets (picks the shape of the segment reduction)
  long long w0, w1;     // chain range a per-step launch (k_post, k_reest) covers: [w0, w1), the whole set by default
  double* cur_theta;    // [P][stride]   (walker-last-step params)
  double* cur_prob;     // [stride]
  double* prop_theta;   // [P][stride]   proposal under evaluation
  double* best_theta;   // [P][stride]   (walker-most-likely-step)
  double* best_prob;
  double* L;            // [P*P][stride] per-chain l-matrix, row-major entries
  long long* loop_i;    // reference loop variable i
  unsigned char* shutting;
  unsigned char* done;  // also the eval kernel's skip mask
  unsigned char* need_reest;
  int* reest_list;      // [stride] chains flagged for a re-estimate, the launch range [w0, w1) fills its own stretch from w0 on
  int* reest_count;     // [4 + 4] per chain group: entries in its stretch; finished CTAs of its k_reest launch
  int group;            // chain group of this launch (0 when the chains advance as one)
  double* temperature;
  long long* length;    // walker-length
  long long* age;       // walker-age
  long long* rows;      // rows ever pushed to the trace ring (slot = row % cap)
  long long* reset_index;
  long long* steps_run;
  int* l_updates;
  double* trace_theta;  // [cap][P][stride]
  double* trace_prob;   // [cap][stride]
  long long cap;
  double* lld;          // [D][stride] per-dataset log-likelihood of the proposal
  double* lpd;          // [D][stride] per-dataset log-prior of the proposal
  const double* xch;    // exchange buffer [R][slots][2][stride] when datasets are sharded (else null)
  int xslots;
  // Peer exchange (data sets sharded over GPUs, no collective call on the step path): the buffer is doubled -- xch +
  // xpar * xpar_stride is the half in use for the exchange in progress -- and every rank's k_finalize stores its own rows
  // straight into the buffers of all peers over NVLink (xpeer[q], mapped peer memory), then raises xflag_peer[q][xrank] to
  // xepoch + 1; the kernels that read the terms first wait until xflag[r] has got there for every r.
  int xpar;                         // 0 / 1
  long long xpar_stride;            // elements between the halves (0: not doubled, host-side collective)
  int xrank, xranks;                // this GPU's rank among the owners; 0 ranks = no peer exchange
  unsigned long long xepoch;        // exchanges completed before this one
  double* xpeer[8];                 // every rank's exchange buffer as seen from this GPU ([xrank] = xch)
  unsigned long long* xflag_peer[8];  // every rank's flag array [xranks] as seen from this GPU
  volatile unsigned long long* xflag;  // this GPU's flags
  int* xdone;                       // blocks of k_finalize that have pushed their rows (reset by the last one)
  int* xerr;                        // set when a wait timed out (a peer died): the host turns it into MFIT_ERR_NCCL
  const MfitDsDev* ds;  // [D] device array
};

struct MfitRunCfg {
  int mode;
  long long n, Tn, S, tail, mwl;
  int auto_mode, sampling_opt, greedy;
  double temperature0;
  const double* temps;  // [Tn]
  const double* z;      // [W][nstream][P] or null
  const double* u;      // [W][nstream] or null
  long long nstream;
  long long t;          // take-step index inside this run
  unsigned long long seed;
  long long chain_base;
  double sr_scale;      // sample-region initial scale
  int l_per_chain;      // the l-matrix handed to the run is W x P x P (one per chain) instead of one shared P x P
};

// ---- small problems: K Metropolis-Hastings steps per launch, one CTA per chain (chain_kernels.cu k_run_small)
#define MFIT_SMALL_MAX_DS 8
#define MFIT_SMALL_THREADS 128
struct MfitSmallDs {       // one shared-data set with a built-in family
  const double *x, *y, *yw, *w, *c;
  long long n;
  int family, degree, ll_kind;
  int idx[16];             // family slot -> shared parameter index
};
struct MfitSmallArgs {
  int n_ds;
  MfitSmallDs d[MFIT_SMALL_MAX_DS];
};

#ifdef __CUDACC__
void mfit_launch_run_small(const MfitChains& c, const MfitRunCfg& r, const MfitSmallArgs& a, long long steps, cudaStream_t st);
void mfit_launch_init_run(const MfitChains& c, const MfitRunCfg& r, const double* l_shared, cudaStream_t st);
void mfit_launch_reest(const MfitChains& c, const MfitRunCfg& r, cudaStream_t st);
void mfit_launch_pre(const MfitChains& c, const MfitRunCfg& r, cudaStream_t st);
void mfit_launch_post(const MfitChains& c, const MfitRunCfg& r, bool fuse_pre, cudaStream_t st);
void mfit_launch_finalize(const MfitChains& c, cudaStream_t st);
void mfit_launch_peer_wait(const MfitChains& c, cudaStream_t st);
void mfit_launch_first_step(const MfitChains& c, cudaStream_t st);
void mfit_launch_force_step(const MfitChains& c, cudaStream_t st);
void mfit_launch_priors(const MfitChains& c, const double* theta, cudaStream_t st);
void mfit_launch_count_active(const MfitChains& c, int* out, cudaStream_t st);
void mfit_launch_keep_walks(const MfitChains& c, long long keep, cudaStream_t st);
void mfit_launch_modify(const MfitChains& c, int mode, long long number, cudaStream_t st);
void mfit_launch_stats(const MfitChains& c, void* stats_dev, cudaStream_t st);
void mfit_launch_trace_gather(const MfitChains& c, long long take, double* theta, double* prob, long long* nvalid,
                              cudaStream_t st);
void mfit_launch_trace_restore(const MfitChains& c, long long rows, const double* theta, const double* prob,
                               const long long* n_valid, const long long* age, const double* best_theta,
                               const double* best_prob, cudaStream_t st);
void mfit_launch_philox_stream(unsigned long long seed, long long chain_base, long long W, long long first,
                               long long nsteps, int P, double* z, double* u, cudaStream_t st);
void mfit_launch_transpose_in(const double* src_rowmajor, double* dst, long long W, int P, long long stride,
                              cudaStream_t st);
void mfit_launch_transpose_out(const double* src, double* dst_rowmajor, long long W, int P, long long stride,
                               cudaStream_t st);
void mfit_launch_eval_finalize(const MfitChains& c, double* prob, double* ll, double* lp, cudaStream_t st);
void mfit_launch_l_read(const MfitChains& c, double* out, cudaStream_t st);
int mfit_measure_fp64_dev(cudaStream_t st, double* dfma_flops, double* cost_exp, double* cost_log, double* cost_sin,
                          double* cost_div);
void mfit_launch_flush(void* buf, size_t bytes, cudaStream_t st);
void mfit_launch_analytics(const MfitChains& c, long long take, int* sel, double* ml_theta, double* ml_prob,
                           double* acceptance, double* cov, double* l_out, cudaStream_t st);
void mfit_launch_median(const MfitChains& c, long long take, double* out, cudaStream_t st);
void mfit_launch_envelope_select(const MfitChains& c, long long take, long long w_first, long long n_chains, long long cols,
                                 const double* fit, double* out_theta, int* n_sel, int* sel, cudaStream_t st);
void mfit_launch_trace_regrow(const MfitChains& c, double* new_theta, double* new_prob, long long new_cap,
                              cudaStream_t st);
#endif
#endif
