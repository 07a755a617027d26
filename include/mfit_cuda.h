/* mfit_cuda.h -- C ABI of libmfit_cuda.so, the B200 (sm_100a) backend for the
 * Metropolis-Hastings hot path of afranson/Lisp-MCMC.
 *
 * Every entry point names the reference interface it stands in for; citations are
 * into /root/reference/mcmc-fitting.lisp unless another file is named.  The host
 * (SBCL through CFFI -- lisp_mcmc_b200/lisp/mfit-cuda.lisp -- or the Python mirror in
 * lisp_mcmc_b200/walker.py through ctypes) owns every host pointer only for the
 * duration of a call; the library copies.  The library owns device memory, its
 * stream, the NVRTC cache and the NCCL communicator.  No callbacks into the host.
 *
 * All functions return 0 on success or a negative mfit_status; the message is
 * available from mfit_last_error().  NaN/Inf in likelihoods are returned, never
 * raised.  A context is single-owner: one host thread at a time.
 *
 * There is no CPU fallback: without a CUDA device mfit_ctx_create fails with
 * MFIT_ERR_CUDA.
 */
#ifndef MFIT_CUDA_H
#define MFIT_CUDA_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MFIT_ABI_VERSION 1

typedef enum {
  MFIT_OK = 0,
  MFIT_ERR_ARG = -1,      /* bad argument (reference: Lisp `error` in clean-data :819,:825) */
  MFIT_ERR_CUDA = -2,     /* CUDA runtime / driver failure, or no device */
  MFIT_ERR_SEXPR = -3,    /* s-expression could not be parsed or lowered */
  MFIT_ERR_NVRTC = -4,    /* NVRTC unavailable or compilation failed (log in mfit_last_error) */
  MFIT_ERR_STATE = -5,    /* call made in the wrong order (e.g. run before chains_init) */
  MFIT_ERR_STREAM = -6,   /* injected z/u stream shorter than the steps requested */
  MFIT_ERR_NCCL = -7,
  MFIT_ERR_NOMEM = -8
} mfit_status;

/* Model families (the reference ships none -- README.md:11 tells users to write them; SURVEY.md 8(d)
 * defines these).  theta order:
 *   POLY (degree D): c0..cD            f = c0 + x(c1 + x(c2 + ...))
 *   EXPO:      a b c                   f = a exp(b x) + c
 *   SINUS:     a omega phi c           f = a sin(omega x + phi) + c
 *   LORENTZ:   a x0 gamma c            f = a gamma^2 / ((x-x0)^2 + gamma^2) + c
 *   GAUSS_LIN: a mu s b0 b1            f = a exp(-((x-mu)/s)^2 / 2) + b0 + b1 x
 *   NV_MULTI (K dips, `degree` = K): scale_1..K mu_1..K sigma bg0  (names: nv-specific.lisp:25-30)
 *                                      f = bg0 - sum_k scale_k sigma^2 / ((x-mu_k)^2 + sigma^2)
 *   SEXPR: user lambda lowered to CUDA C and compiled by NVRTC for sm_100a */
enum { MFIT_FAM_POLY = 0, MFIT_FAM_EXPO = 1, MFIT_FAM_SINUS = 2, MFIT_FAM_LORENTZ = 3,
       MFIT_FAM_GAUSS_LIN = 4, MFIT_FAM_NV_MULTI = 5, MFIT_FAM_SEXPR = 100 };

/* :log-liklihood slot values.  NORMAL = log-liklihood-normal :393-400, NORMAL_CUTOFF =
 * log-liklihood-normal-cutoff :419-427, POISSON = (create-log-liklihood-function
 * (lambda (y model error) (log-poisson model y))) :402-417 with log-poisson :382-383. */
enum { MFIT_LL_NORMAL = 0, MFIT_LL_NORMAL_CUTOFF = 1, MFIT_LL_POISSON = 2,
       MFIT_LL_USER = 3 /* set by mfit_likelihood_sexpr only */ };

enum { MFIT_AUTO_NONE = 0, MFIT_AUTO_PROB_SETTLE = 1 };      /* :auto, :862 */
enum { MFIT_OPT_COVARIANCE = 0, MFIT_OPT_BEST_VALUE = 1 };   /* :sampling-optimization, :862 */

typedef struct mfit_ctx mfit_ctx;

/* keyword arguments of walker-adaptive-steps-full, :862 */
typedef struct {
  int64_t n;                    /* :n (default 100000); walker-adaptive-steps :946 passes 30000 */
  double temperature;           /* :temperature (default 1d3; walker-adaptive-steps passes 10) */
  int32_t auto_mode;            /* :auto  MFIT_AUTO_* (:slope-settle is not supported) */
  int32_t sampling_optimization;/* :sampling-optimization MFIT_OPT_* */
  int64_t max_walker_length;    /* :max-walker-length, 0 = nil */
  int32_t greedy;               /* 1 = accept only uphill (walker-pretend-take-step :1117) */
  int32_t reserved;
} mfit_adapt_cfg;

/* per-chain summary, filled by mfit_stats_read (struct walker :467-479 scalar slots) */
typedef struct {
  int64_t length;          /* walker-length */
  int64_t age;             /* walker-age */
  int64_t steps_this_run;  /* take-step calls made by the last run */
  double last_prob;        /* (walker-step-prob (walker-last-step w)) */
  double best_prob;        /* (walker-step-prob (walker-most-likely-step w)) */
  double temperature;      /* temperature at exit of the last run */
  double accept_200;       /* walker-get :acceptance :take 200, :506-508 */
  int32_t shutting_down;   /* shutting-down-p at exit */
  int32_t l_updates;       /* number of l-matrix updates (rescale or re-estimate) in the last run */
} mfit_chain_stats;

/* ---- context ------------------------------------------------------------------------------ */
int mfit_abi_version(void);
/* `stream` is a cudaStream_t (NULL = the library creates its own non-blocking stream). */
int mfit_ctx_create(int device, void* stream, mfit_ctx** out);
void mfit_ctx_destroy(mfit_ctx* ctx);
const char* mfit_last_error(mfit_ctx* ctx);     /* ctx may be NULL: last create error */
int mfit_synchronize(mfit_ctx* ctx);
int64_t mfit_launch_count(mfit_ctx* ctx);       /* kernels launched by this ctx so far */

/* ---- walker-create :1132-1163: problem definition -------------------------------------------
 * D datasets (length of the :function list), P shared parameters (length of the :params plist). */
int mfit_problem_define(mfit_ctx* ctx, int n_datasets, int n_params);
/* :data (list x y) and :data-error of dataset d.  sigma==NULL means data-error 1 (:1144).
 * Stands in for clean-data / clean-data-error / to-double-floats :774-835 (the host flattens). */
int mfit_dataset_upload(mfit_ctx* ctx, int d, const double* x, const double* y, const double* sigma, int64_t n);
/* Multi-dimensional inputs (walker-create docstring :1137: "x may itself be a sequence"): every x of the :data is a
 * sequence of x_dims numbers, flattened by the host component-major -- component k of point i at x[k * n + i] -- and
 * the :function (a lowered lambda, mfit_model_sexpr) reads them with elt / aref / svref / nth / first ... fourth.
 * x_dims = 0 is mfit_dataset_upload.  A model that uses x as a number on such data, or indexes beyond x_dims, is the
 * reference's type error: MFIT_ERR_ARG when the chains are initialised. */
#define MFIT_MAX_XDIMS 8
int mfit_dataset_upload_xd(mfit_ctx* ctx, int d, const double* x, int x_dims, const double* y, const double* sigma,
                           int64_t n);
/* Walker set with per-walker data -- the reference's own many-walker workload (nv-specific.lisp:50-66: one walker per
 * measured spectrum, same :function / priors, `(mapc #'walker-adaptive-steps walkers)`): chain w of W is fitted to its
 * own row of y (W x n row-major) and of x / sigma unless x_shared / sigma_shared say that one n-vector serves all.
 * mfit_chains_init / mfit_eval_logp must then be called with exactly W chains.  One warp per chain on the device. */
int mfit_dataset_upload_per_walker(mfit_ctx* ctx, int d, const double* x, const double* y, const double* sigma,
                                   int64_t n, int64_t W, int x_shared, int sigma_shared);
/* :function slot of dataset d = built-in family; param_index maps family theta slots to the
 * shared parameter vector (global fits share keys across datasets, README.md "Global Parameter Fitting"). */
int mfit_model_builtin(mfit_ctx* ctx, int d, int family, int degree, const int32_t* param_index, int n_idx);
/* :function slot = user lambda text, e.g. "(lambda (x &key a b &allow-other-keys) (+ b (* a x)))";
 * key_names are the P keys of the :params plist in order.  Lowered to CUDA C, NVRTC -> sm_100a. */
int mfit_model_sexpr(mfit_ctx* ctx, int d, const char* lambda_text, const char* const* key_names, int n_keys);
int mfit_likelihood_set(mfit_ctx* ctx, int d, int kind);
/* :log-liklihood = (create-log-liklihood-function <lambda>) :402-417 with any per-point (lambda (y model error) ...):
 * the text is lowered to CUDA C (csrc/lower.cpp, kind 2) and compiled by NVRTC into the same likelihood kernel, together
 * with the data set's :function (built-in family or lowered lambda).  The kernel sums g(y_i, f(x_i; theta), error) in
 * data order; nothing is hoisted.  As in the reference (:416) `error` is the data set's WHOLE stddev list, not the
 * point's: (elt error k) / (first error) ... with a literal k < 16 read it (a per-data-set constant).  The reference's
 * own densities may be called: (log-normal x mu sigma), (log-poisson lambda y), (log-factorial y) -- the latter two
 * with the lambda's y only (a table built at upload, single-float accumulation :379-380).
 * (lambda (y model error) (log-poisson model y)) is recognised and becomes MFIT_LL_POISSON itself. */
int mfit_likelihood_sexpr(mfit_ctx* ctx, int d, const char* lambda_text);
/* :log-prior slot.  flat = log-prior-flat :340-343; bounds = (prior-bounds-let ((k lo hi)...) bounds-total)
 * :346-369 with constant bounds (idx<0 means the key is absent: getf default 0d0); sexpr = any
 * (lambda (params data) ...) body, prior-bounds-let included. */
int mfit_prior_flat(mfit_ctx* ctx, int d);
int mfit_prior_bounds(mfit_ctx* ctx, int d, int n, const int32_t* idx, const double* lo, const double* hi);
int mfit_prior_sexpr(mfit_ctx* ctx, int d, const char* lambda_text, const char* const* key_names, int n_keys);

/* s-expression -> CUDA C lowering, exposed so it can be inspected and tested without a GPU.
 * kind 0 = model lambda, 1 = prior lambda, 2 = per-point likelihood lambda (key_names unused).
 * Returns bytes needed (incl. NUL) or a negative status. */
int mfit_lower_sexpr(const char* lambda_text, const char* const* key_names, int n_keys, int kind,
                     char* out, int out_len, char* err, int err_len);

/* lower + NVRTC-compile (sm_100a cubin) without touching a GPU; returns the cubin size or a negative status */
int mfit_compile_sexpr(const char* lambda_text, const char* const* key_names, int n_keys, int kind,
                       char* err, int err_len);
/* points per reduction segment = 128 * tiles (0 = default: a function of the point count only, so that
 * log-likelihood bits do not depend on chain count or GPU count) */
int mfit_set_segment_tiles(mfit_ctx* ctx, int tiles_per_segment);
/* the geometry in use for data set d once chains exist or a batch has been evaluated: 128-point tiles per segment and
 * segments (= partial sums per chain, added in segment order); tests pin parity at the benchmarked geometry with it */
int mfit_segment_geometry(mfit_ctx* ctx, int d, int32_t* tiles_per_segment, int32_t* n_segments);

/* ---- walker-make-step :1067-1070 over a batch of parameter vectors ------------------------------
 * theta: W x P row-major.  ll, lp: W x D row-major per-dataset terms (either may be NULL);
 * prob: W totals = (sum_d ll) + (sum_d lp) in dataset order. */
int mfit_eval_logp(mfit_ctx* ctx, const double* theta, int64_t W, double* prob, double* ll, double* lp);

/* ---- chains: W independent walkers sharing the problem ------------------------------------------
 * theta0: W x P.  Evaluates the first step like walker-create :1148-1162 (length = age = 1).
 * chain_id_base: global id of chain 0 (Philox streams are keyed by global chain id, so a sharded run
 * draws the same numbers as a single-GPU run).  trace_capacity: rows of history kept per chain
 * (0 = as many as the next run needs, bounded by free memory; always >= max(1000, 10*max(50,P)) + 1). */
int mfit_chains_init(mfit_ctx* ctx, const double* theta0, int64_t W, uint64_t seed, int64_t chain_id_base,
                     int64_t trace_capacity);
/* walker-adaptive-steps-full :862-942.  l_matrix: NULL, or P x P row-major shared by all chains
 * (:l-matrix).  z: W x nstream x P normals and u: W x nstream uniforms, indexed by the count of
 * take-step calls in this run (NULL = on-device Philox4x32-10).  steps_launched (may be NULL)
 * receives the number of MH steps enqueued (the longest chain). */
int mfit_run_adaptive(mfit_ctx* ctx, const mfit_adapt_cfg* cfg, const double* l_matrix,
                      const double* z, const double* u, int64_t nstream, int64_t* steps_launched);
/* The same run in two halves, so that the host can poll mfit-walker-estop (:860, tested every iteration at :904),
 * report progress or time it: begin = the set-up of :866-901, advance = up to max_steps more iterations of the
 * `do` loop :902-942 (-1 = to the end).  *finished is set when every chain has left the loop. */
int mfit_run_begin(mfit_ctx* ctx, const mfit_adapt_cfg* cfg, const double* l_matrix, const double* z, const double* u,
                   int64_t nstream);
int mfit_run_advance(mfit_ctx* ctx, int64_t max_steps, int64_t* steps_done, int* finished);
/* walker-many-steps :849-853: n steps, constant l-matrix (required), temperature 1 */
int mfit_run_many_steps(mfit_ctx* ctx, int64_t n, const double* l_matrix, const double* z, const double* u,
                        int64_t nstream);
/* the same with one l-matrix per chain (W x P x P): walker-many-steps' default, (diagonal-covariance (scale-plist 1e-2
 * median-params)) :851, is made per walker, and a walker set holds W of them */
int mfit_run_many_steps_per_chain(mfit_ctx* ctx, int64_t n, const double* l_matrices, const double* z, const double* u,
                                  int64_t nstream);
/* walker-sample-region :949-969 */
int mfit_run_sample_region(mfit_ctx* ctx, double initial_scale, const double* z, int64_t nstream);
/* walker-force-take-step :1124-1129 */
int mfit_force_take_step(mfit_ctx* ctx);
/* mfit-walker-estop :860: asynchronous stop request, polled between launches */
int mfit_request_stop(mfit_ctx* ctx);

/* walker-get :steps / :log-liklihoods :take, newest first.  theta: W x take x P, prob: W x take,
 * n_valid: W (rows filled per chain; the rest is NaN).  Any output may be NULL. */
int mfit_trace_read(mfit_ctx* ctx, int64_t take, double* theta, double* prob, int64_t* n_valid);
int mfit_stats_read(mfit_ctx* ctx, mfit_chain_stats* stats /* W */);
/* walker-load :989-1001 revived (checkpoint / resume): after mfit_chains_init with W chains, replace every chain's walk
 * by a saved one -- the inverse of mfit_trace_read.  theta W x rows x P and prob W x rows, newest first; chain w holds
 * n_valid[w] <= rows steps.  The last step becomes the newest row; walker-length = steps held; walker-age = age[w]
 * (NULL: the steps held), which also keys the Philox stream, so a restored chain continues exactly as the saved one
 * would have; the most likely step is the saved one (best_theta W x P, best_prob W) unless the walk holds a better
 * one, or, with NULL, the first strictly largest prob from the oldest row up (walker-modify :add-walks :558-563). */
int mfit_chains_restore(mfit_ctx* ctx, int64_t rows, const double* theta, const double* prob, const int64_t* n_valid,
                        const int64_t* age, const double* best_theta, const double* best_prob);
/* most-likely-step :474 and last-step :475.  theta outputs W x P. */
int mfit_best_read(mfit_ctx* ctx, double* theta, double* prob);
int mfit_last_read(mfit_ctx* ctx, double* theta, double* prob);
/* l-matrix in use at the end of the last run, W x P x P */
int mfit_l_matrix_read(mfit_ctx* ctx, double* l);
/* walker-get analytics over the newest `take` rows of every chain, computed on the device.  Any output
 * may be NULL.
 *   most_likely_theta W x P, most_likely_prob W : :most-likely-step :503-505 (largest prob; of equal ones the oldest,
 *                                                 as the reference's `reduce` leaves it)
 *   median W x P        : :median-params :516-523 = nth-percentile 50 :1493-1504 (mean of the two middle order
 *                         statistics when the count is even)
 *   acceptance W        : :acceptance :506-508
 *   covariance W x P x P: :covariance-matrix = lplist-covariance :614-643 of the :unique-steps :491-495
 *   l_matrix W x P x P  : :l-matrix :543 = cholesky-decomp of the covariance of the differences of the :forward-steps;
 *                         its diagonal is :stddev-params :526-539.  NaN where the reference would signal an error
 *                         (fewer than two forward steps). */
int mfit_trace_analytics(mfit_ctx* ctx, int64_t take, double* most_likely_theta, double* most_likely_prob,
                         double* median, double* acceptance, double* covariance, double* l_matrix);
/* walker-get-data-and-fit :1231-1255 for every chain at once, on data set d (:fn-number): the fit curve and the
 * max / min envelope the reference plots (walker-plot-data-and-fit :1257).
 *   take            :take, clamped per chain to its walker length (:1233-1235)
 *   which_solution  MFIT_FIT_MOST_LIKELY: (walker-get :most-likely-step :take take); MFIT_FIT_MEDIAN: :median-params
 *   x_fit[n_fit]    abscissae, made by the caller: (linspace (reduce #'min x-data) (reduce #'max x-data) :len 1000)
 *   y_fit   W x n_fit : (apply fn x params) :1246
 *   y_max, y_min W x n_fit : over the ceil(0.66 take) most probable entries of the chain's whole walk (:1248-1252;
 *                   0.66 is a single-float and so is the product), the largest / smallest model value at every x;
 *                   NaN for a chain without steps
 *   params  W x P   : the parameters of the fit curve
 *   n_distinct W    : distinct steps among the selected entries (a rejected proposal repeats its step in the walk)
 * x-shift / y-shift of the reference are plain additions left to the caller.  Any output may be NULL. */
#define MFIT_FIT_MOST_LIKELY 0
#define MFIT_FIT_MEDIAN 1
int mfit_fit_envelope(mfit_ctx* ctx, int d, int64_t take, int which_solution, const double* x_fit, int n_fit,
                      double* y_fit, double* y_max, double* y_min, double* params, int64_t* n_distinct);
/* walker-modify :keep-walks :568-569 applied to every chain */
int mfit_keep_walks(mfit_ctx* ctx, int64_t keep);
/* the other list surgery of walker-modify :546-580, applied to every chain: :burn-walks (the walk loses its OLDEST
 * `number` steps, :566-567), :keep-walks (= mfit_keep_walks), :reset (the walk shrinks to its OLDEST step, which becomes
 * the last step -- (last walk) on a newest-first list, :570-573), :reset-to-most-likely (:574-578).  walker-age and the
 * most likely step are left alone, as in the reference.  (:add-step is the MH step itself, :add-walks is
 * mfit_chains_restore, :delete is mfit_ctx_destroy.) */
#define MFIT_MODIFY_BURN_WALKS 1
#define MFIT_MODIFY_KEEP_WALKS 2
#define MFIT_MODIFY_RESET 3
#define MFIT_MODIFY_RESET_TO_MOST_LIKELY 4
int mfit_walker_modify(mfit_ctx* ctx, int mode, int64_t number);

/* normals / uniforms the on-device Philox stream yields for (seed, global chain id, step index):
 * lets a test replay a Philox run through the injected-stream path of another implementation. */
int mfit_philox_stream(mfit_ctx* ctx, uint64_t seed, int64_t chain_id_base, int64_t W, int64_t first_step,
                       int64_t nsteps, int P, double* z /* W x nsteps x P */, double* u /* W x nsteps */);

/* ---- multi-GPU (one process per GPU) --------------------------------------------------------------
 * Chain-sharded runs need no collective on the data path.  Dataset-sharded global fits
 * (walker-make-step's sum over datasets, :1068-1069, split across ranks) exchange per-dataset
 * ll/lp every step with ncclAllGather and sum them in dataset order on every rank. */
int mfit_nccl_unique_id(void* id_out /* 128 bytes */);
int mfit_nccl_attach(mfit_ctx* ctx, const void* unique_id, int rank, int n_ranks);
/* Peer exchange: the same sharding without a collective call on the step path.  Every rank reserves an exchange block
 * (mfit_peer_setup: plain device memory, exported as a CUDA IPC handle), the 128-byte handles of all ranks are passed to
 * every rank in rank order (mfit_peer_attach maps them: IPC between processes, direct peer access inside one), and from
 * then on the kernel that reduces a rank's data sets stores the per-chain terms straight into every peer's block over
 * NVLink and raises a flag there; the accept kernel of the other GPUs waits for the flags.  Needs all ranks on one box
 * with peer access (NVSwitch).  max_chains / max_datasets_per_rank bound the chains of a walker set and the data sets a
 * rank may own.  With a communicator attached as well (mfit_nccl_attach) only the rare stop test uses it. */
#define MFIT_PEER_HANDLE_BYTES 128
int mfit_peer_setup(mfit_ctx* ctx, int rank, int n_ranks, int64_t max_chains, int max_datasets_per_rank, void* handle_out);
int mfit_peer_attach(mfit_ctx* ctx, const void* handles /* n_ranks x MFIT_PEER_HANDLE_BYTES */);
/* owner[d] = rank that evaluates dataset d (others skip it). Requires mfit_nccl_attach or mfit_peer_attach. */
int mfit_dataset_owners(mfit_ctx* ctx, const int32_t* owner);
/* gather chain-sharded traces on every rank: local (W_local x take x (P+1)) -> out (n_ranks x ...) */
int mfit_nccl_allgather_f64(mfit_ctx* ctx, const double* local_host, int64_t count, double* out_host);

/* ---- several GPUs behind ONE handle (one host thread per GPU inside the library; no Python, no MPI) -----------------
 * What the Lisp host's :devices uses.  The walker set is sharded over the listed devices of this box either by CHAINS
 * (MFIT_SHARD_CHAINS: W chains split contiguously, every data set on every GPU, Philox streams keyed by global chain id,
 * nothing crosses GPUs during a run) or by DATA SETS (MFIT_SHARD_DATASETS: data set d lives on GPU d mod n, every GPU holds
 * all chains, the per-data-set terms travel through the peer exchange above).  Either way results are bit-identical to one
 * GPU.  Calls mirror the single-GPU ones; per-chain outputs come back in global chain order. */
typedef struct mfit_multi mfit_multi;
enum { MFIT_SHARD_CHAINS = 0, MFIT_SHARD_DATASETS = 1 };
int mfit_multi_create(int n_gpus, const int* dev_ids, int shard_mode, mfit_multi** out);
void mfit_multi_destroy(mfit_multi* m);
const char* mfit_multi_last_error(mfit_multi* m);
int mfit_multi_size(mfit_multi* m);
mfit_ctx* mfit_multi_ctx(mfit_multi* m, int rank);          /* the per-GPU context (read-only uses: launch counts, stats) */
int64_t mfit_multi_launch_count(mfit_multi* m);             /* kernels launched by all of its contexts */
/* MFIT_SHARD_DATASETS: chains the exchange blocks are sized for (default 131072; 32 bytes x chains x data sets per GPU x
 * GPUs on every GPU).  Before mfit_multi_problem_define. */
int mfit_multi_reserve(mfit_multi* m, int64_t max_chains);
int mfit_multi_problem_define(mfit_multi* m, int n_datasets, int n_params);
int mfit_multi_dataset_upload(mfit_multi* m, int d, const double* x, int x_dims, const double* y, const double* sigma, int64_t n);
int mfit_multi_model_builtin(mfit_multi* m, int d, int family, int degree, const int32_t* param_index, int n_idx);
int mfit_multi_model_sexpr(mfit_multi* m, int d, const char* lambda_text, const char* const* key_names, int n_keys);
int mfit_multi_likelihood_set(mfit_multi* m, int d, int kind);
int mfit_multi_likelihood_sexpr(mfit_multi* m, int d, const char* lambda_text);
int mfit_multi_prior_flat(mfit_multi* m, int d);
int mfit_multi_prior_bounds(mfit_multi* m, int d, int n, const int32_t* idx, const double* lo, const double* hi);
int mfit_multi_prior_sexpr(mfit_multi* m, int d, const char* lambda_text, const char* const* key_names, int n_keys);
int mfit_multi_eval_logp(mfit_multi* m, const double* theta, int64_t W, double* prob, double* ll, double* lp);
int mfit_multi_chains_init(mfit_multi* m, const double* theta0, int64_t W, uint64_t seed, int64_t trace_capacity);
int mfit_multi_run_adaptive(mfit_multi* m, const mfit_adapt_cfg* cfg, const double* l_matrix, int64_t* steps_launched);
int mfit_multi_run_begin(mfit_multi* m, const mfit_adapt_cfg* cfg, const double* l_matrix);
int mfit_multi_run_advance(mfit_multi* m, int64_t max_steps, int64_t* steps_done, int* finished);
int mfit_multi_run_many_steps(mfit_multi* m, int64_t n, const double* l_matrix);
int mfit_multi_request_stop(mfit_multi* m);
int mfit_multi_trace_read(mfit_multi* m, int64_t take, double* theta, double* prob, int64_t* n_valid);
int mfit_multi_stats_read(mfit_multi* m, mfit_chain_stats* stats /* W */);
int mfit_multi_best_read(mfit_multi* m, double* theta, double* prob);
int mfit_multi_last_read(mfit_multi* m, double* theta, double* prob);
int mfit_multi_l_matrix_read(mfit_multi* m, double* l /* W x P x P */);
/* the trace consumers and editors (mfit_trace_analytics, mfit_fit_envelope, mfit_walker_modify, mfit_keep_walks,
 * mfit_chains_restore) on a sharded walker set: same arguments, global chain order */
int mfit_multi_trace_analytics(mfit_multi* m, int64_t take, double* most_likely_theta, double* most_likely_prob, double* median,
                               double* acceptance, double* covariance, double* l_matrix);
int mfit_multi_fit_envelope(mfit_multi* m, int d, int64_t take, int which_solution, const double* x_fit, int n_fit, double* y_fit,
                            double* y_max, double* y_min, double* params, int64_t* n_distinct);
int mfit_multi_walker_modify(mfit_multi* m, int mode, int64_t number);
int mfit_multi_keep_walks(mfit_multi* m, int64_t keep);
int mfit_multi_chains_restore(mfit_multi* m, int64_t rows, const double* theta, const double* prob, const int64_t* n_valid,
                              const int64_t* age, const double* best_theta, const double* best_prob);

/* ---- measurement helpers ----------------------------------------------------------------------------
 * FP64 DFMA issue peak of the device this ctx is bound to, measured live (FLOP/s, 2 per DFMA), and the
 * DFMA-equivalent cost of the FP64 transcendentals used by the model families. */
int mfit_measure_fp64(mfit_ctx* ctx, double* dfma_flops, double* cost_exp, double* cost_log, double* cost_sin,
                      double* cost_div);
/* record CUDA events around the likelihood kernels of subsequent runs (off by default) */
int mfit_set_timing(mfit_ctx* ctx, int on);
/* device time of the eval kernels of the last run / eval call, seconds (CUDA events on the ctx stream),
 * and how many eval launches it covered */
int mfit_eval_kernel_time(mfit_ctx* ctx, double* seconds, int64_t* launches);
/* 128-point tiles the likelihood kernels evaluated, per loop variant, since the last call while timing was on:
 * [0] guard-free, [1] in-place clamps only, [2] recording guard, [3] guard-free with fused argument arithmetic
 * (csrc/mfit_eval.cuh MFIT_MODE_*).  bench.py weights the FP64 instruction count of the roofline with it. */
int mfit_eval_mode_tiles(mfit_ctx* ctx, int64_t* tiles4);
/* A run advances the chains as up to four groups on separate streams when there are enough of them, so that one group's
 * per-chain kernels overlap the others' likelihood kernels (results do not depend on it).  0 = automatic (3, 2 or 1 by
 * chain count; 2 at most for walker sets with per-walker data), 1 = one group (also what mfit_set_timing(1) implies: an overlapped kernel cannot be timed with events),
 * 2..4 = that many whenever possible. */
int mfit_set_groups(mfit_ctx* ctx, int n_groups);
/* Small problems (one walker or a few, up to a few thousand points, built-in families and priors -- the reference's own
 * mcmc-fit use, test.lisp:14-24) advance many steps per kernel launch, one CTA per chain, instead of three launches per
 * step.  0 = automatic (by size), 1 = never, 2 = whenever the problem qualifies (any size).  Results agree with the
 * three-launch path to rounding (the likelihood sum is formed in another fixed order), not bit for bit. */
int mfit_set_small_path(mfit_ctx* ctx, int mode);
/* write > L2 bytes on the ctx stream (bench hygiene between timed steps) */
int mfit_flush_l2(mfit_ctx* ctx);
/* the cudaStream_t the context enqueues on */
int mfit_stream_handle(mfit_ctx* ctx, void** stream_out);

#ifdef __cplusplus
}
#endif
#endif
