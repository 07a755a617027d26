/* mfit_oracle.h -- C face of the CPU ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the hot path of
 * afranson/Lisp-MCMC (mcmc-fitting.lisp).  It is the *checker* for the CUDA
 * product in lisp_mcmc_b200/: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  Nothing in
 * lisp_mcmc_b200/ links, imports or executes anything in this directory.
 *
 * Parity pinning: the reference is Common Lisp and no Lisp exists in this
 * image, so the reference cannot be executed.  The oracle is pinned against
 * the only known-answer vectors the reference carries (example-covariance /
 * example-l-matrix, mcmc-fitting.lisp:729-751, the docstring examples) and
 * against closed forms computed with mpmath; for log-likelihood values and
 * accept/reject sequences the reference holds no golden vector, so parity for
 * those is "parity unpinned beyond the restatement" (see DESIGN.md).
 *
 * All file:line citations are into /root/reference/mcmc-fitting.lisp unless
 * another file is named.
 */
#ifndef MFIT_ORACLE_H
#define MFIT_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* model families (defined by this build); ids shared with include/mfit_cuda.h */
enum { ORC_FAM_POLY = 0, ORC_FAM_EXPO = 1, ORC_FAM_SINUS = 2, ORC_FAM_LORENTZ = 3,
       ORC_FAM_GAUSS_LIN = 4, ORC_FAM_NV_MULTI = 5, ORC_FAM_SEXPR = 100 };
enum { ORC_LL_NORMAL = 0, ORC_LL_NORMAL_CUTOFF = 1, ORC_LL_POISSON = 2, ORC_LL_USER = 3 };
enum { ORC_AUTO_NONE = 0, ORC_AUTO_PROB_SETTLE = 1 };
enum { ORC_OPT_COVARIANCE = 0, ORC_OPT_BEST_VALUE = 1 };
enum { ORC_SUM_REFERENCE = 0, ORC_SUM_COMPENSATED = 1 };

/* a1-a3: per-point densities, :372-383 */
double orc_log_normal(double x, double mu, double sigma);
double orc_log_factorial(double n);
double orc_log_poisson(double lambda, double k);

/* a13: :614-643, :583-598, :605-611 ; a10: :679-700 */
void orc_lplist_covariance(const double* rows, int n, int P, double* cov);
void orc_cholesky_decomp(const double* cov, int P, double* L);
void orc_scale_array(double s, double* m, int count);
void orc_get_covariant_sample(const double* means, const double* L, int P, const double* z, double* out);
/* rows: n x P newest-first steps (params), returns chol(cov(diff rows)) like walker-get :l-matrix minus selection */
void orc_lplist_to_l_matrix(const double* rows, int n, int P, double* L);

/* built-in family formulas, natural (reference-style) operation order */
double orc_model_builtin(int family, int degree, const double* th, double x);

/* s-expression functions: (lambda (x &key a b &allow-other-keys) body) and
 * (lambda (params data) (prior-bounds-let (...) body)) */
typedef struct orc_fn orc_fn;
orc_fn* orc_fn_parse(const char* text, const char* const* keys, int P, char* err, int errlen);
void orc_fn_free(orc_fn* f);
double orc_fn_call_model(orc_fn* f, double x, const double* theta);
double orc_fn_call_model_v(orc_fn* f, const double* x, int xdim, const double* theta);
double orc_fn_call_prior(orc_fn* f, const double* theta);

/* temperature schedule, :875-878 */
void orc_temperature_schedule(long n, int P, double temperature, double* temps /* max(n,10S) */);
long orc_temp_steps(long n, int P);

/* walker (a9-a17) */
typedef struct orc_walker orc_walker;
orc_walker* orc_walker_new(int D, int P);
void orc_walker_free(orc_walker* w);
int orc_walker_dataset(orc_walker* w, int d, const double* x, const double* y, const double* sigma, long n);
/* multi-dimensional inputs (:1137): every x is a sequence of xd numbers, x[i*xd + k]; s-expression models only */
int orc_walker_dataset_xd(orc_walker* w, int d, const double* x, int xd, const double* y, const double* sigma, long n);
int orc_walker_model_builtin(orc_walker* w, int d, int family, int degree, const int* idx, int n_idx);
int orc_walker_model_sexpr(orc_walker* w, int d, const char* text, const char* const* keys);
int orc_walker_likelihood(orc_walker* w, int d, int kind);
/* :log-liklihood = (create-log-liklihood-function (lambda (y model error) ...)), mcmc-fitting.lisp:402-417; `error` is the
 * data set's whole stddev list (:416): (elt error k) and friends with a literal index */
int orc_walker_likelihood_sexpr(orc_walker* w, int d, const char* text);
int orc_walker_prior_flat(orc_walker* w, int d);
int orc_walker_prior_bounds(orc_walker* w, int d, int n, const int* idx, const double* lo, const double* hi);
int orc_walker_prior_sexpr(orc_walker* w, int d, const char* text, const char* const* keys);
const char* orc_walker_error(orc_walker* w);
/* walker-make-step :1067-1070 ; ll_d / lp_d may be NULL */
int orc_walker_logp(orc_walker* w, const double* theta, int sum_mode, double* prob, double* ll_d, double* lp_d);
/* walker-create's first step :1148-1162 */
int orc_walker_init(orc_walker* w, const double* theta0);

typedef struct {
  long n;                  /* :n */
  double temperature;      /* :temperature */
  int auto_mode;           /* :auto */
  int sampling_optimization;
  long max_walker_length;  /* 0 = nil */
  int greedy;              /* 0 */
} orc_adapt_cfg;

/* z: nstream x P normals, u: nstream uniforms, both indexed by the count of
 * take-step calls inside this call.  NULL => internal PRNG (seeded). */
int orc_walker_adaptive_steps_full(orc_walker* w, const orc_adapt_cfg* cfg, const double* l_matrix_or_null,
                                   const double* z, const double* u, long nstream, long* steps_taken);
/* the same loop interrupted after max_steps take-step calls (mfit-walker-estop, :860/:904); -1 = no limit */
int orc_walker_adaptive_steps_limited(orc_walker* w, const orc_adapt_cfg* cfg, const double* l_matrix_or_null,
                                      const double* z, const double* u, long nstream, long max_steps,
                                      long* steps_taken);
int orc_walker_many_steps(orc_walker* w, long n, const double* l_matrix_or_null,
                          const double* z, const double* u, long nstream);
int orc_walker_sample_region(orc_walker* w, double initial_scale, const double* z, long nstream);
int orc_walker_force_take_step(orc_walker* w);
void orc_walker_seed(orc_walker* w, unsigned long long seed);

long orc_walker_length(orc_walker* w);
long orc_walker_age(orc_walker* w);
/* newest-first, like walker-get :steps :take */
long orc_walker_trace(orc_walker* w, long take, double* theta, double* prob);
void orc_walker_most_likely(orc_walker* w, double* theta, double* prob);
void orc_walker_last_l_matrix(orc_walker* w, double* L);
/* walker-get :acceptance :take -> count, len (exact rational count/len) */
void orc_walker_acceptance(orc_walker* w, long take, long* count, long* len);
int orc_walker_keep_walks(orc_walker* w, long keep);
/* walker-modify :566-578: mode 1 :burn-walks, 2 :keep-walks, 3 :reset, 4 :reset-to-most-likely */
int orc_walker_modify(orc_walker* w, int mode, long number);
/* test support: replace the walk (theta n x P, prob n; newest first) */
int orc_walker_load_walk(orc_walker* w, const double* theta, const double* prob, long n);
/* walker-get-data-and-fit :1231-1255 on dataset d: fit curve at the most likely (0) / median (1) parameters of the
 * newest `take` steps and the max / min envelope over the ceil(0.66 take) most probable steps of the whole walk */
int orc_walker_data_and_fit(orc_walker* w, long take, int d, int which_solution, const double* x_fit, long n_fit,
                            double* y_fit, double* max_ys, double* min_ys, double* params);

/* multi-threaded CPU baseline: W independent walkers sharing one configuration
 * (a template walker is cloned), internal PRNG, returns wall seconds */
double orc_bench_many(orc_walker* templ, const double* theta0 /* W x P */, long W, const orc_adapt_cfg* cfg,
                      int n_threads, unsigned long long seed, double* last_prob /* W */, long max_steps);

#ifdef __cplusplus
}
#endif
#endif
