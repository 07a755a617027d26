// orc_walker.cpp -- ORACLE (test infrastructure only; see mfit_oracle.h).
//
// Restatement of the Metropolis-Hastings hot path of mcmc-fitting.lisp with the
// reference's operation order kept: per-point log-normal / log-poisson,
// sequential left-to-right sums, theta' = theta + L.z with a row sum started at
// 0d0, accept iff p1>p0 or (p1-p0)/T > ln u, the acceptance-window L-matrix
// rescale and forward-step covariance re-estimate, the cosine temperature
// schedule and the :prob-settle auto stop.  Compile with -ffp-contract=off: SBCL
// on x86-64 never fuses a multiply with an add.
#include "orc_internal.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <random>
#include <stdexcept>
#include <thread>

// --------------------------------------------------------------- a1-a3
// log-normal, :372-377:  (+ (* -1/2 (log (* 2 pi))) (* -1 (log sigma)) (* -1/2 (expt (/ (- x mu) sigma) 2d0)))
extern "C" double orc_log_normal(double x, double mu, double sigma) {
  const double pi = 3.141592653589793;
  double a = -0.5 * std::log(2.0 * pi);
  double b = -1.0 * std::log(sigma);
  double z = (x - mu) / sigma;
  double c = -0.5 * std::pow(z, 2.0);  // (expt z 2d0): double exponent -> libm pow
  return (a + b) + c;
}

// log-factorial, :379-380 with up-to :256-261.  (reduce (lambda (x y) (+ x (log y))) '(0 1 ... n)):
// the first element 0 is the initial accumulator; (log <integer>) is a SINGLE-float, so the
// whole sum runs in single precision.  n = 0 returns the integer 0.
extern "C" double orc_log_factorial(double n) {
  long long m = (long long)n;
  float acc = 0.0f;
  for (long long j = 1; j <= m; j++) acc = acc + (float)std::log((double)j);
  return (double)acc;
}

// log-poisson, :382-383:  (- (* k (log lambda)) lambda (log-factorial k))
extern "C" double orc_log_poisson(double lambda, double k) {
  return ((k * std::log(lambda)) - lambda) - orc_log_factorial(k);
}

// --------------------------------------------------------------- a13 / a10
// lplist-covariance, :614-643.  rows[k*P + i] = value of parameter i in plist k.
extern "C" void orc_lplist_covariance(const double* rows, int n, int P, double* cov) {
  std::vector<double> avg(P), an((size_t)P * n);
  for (int i = 0; i < P; i++) {
    double s = rows[i];  // (reduce #'+ x): starts from the first element
    for (int k = 1; k < n; k++) s = s + rows[(size_t)k * P + i];
    avg[i] = s / (double)n;  // :626
  }
  for (int i = 0; i < P; i++)
    for (int k = 0; k < n; k++) an[(size_t)i * n + k] = rows[(size_t)k * P + i] - avg[i];  // :635
  for (int i = 0; i < P; i++)
    for (int j = 0; j < P; j++) {
      double s = 0.0;
      for (int k = 0; k < n; k++) s = s + (an[(size_t)i * n + k] * an[(size_t)j * n + k]) / (double)n;  // :643
      cov[i * P + j] = s;
    }
}

// cholesky-decomp, :583-598 (Cholesky-Banachiewicz, sqrt(max 0 .) on the diagonal, unguarded divide)
extern "C" void orc_cholesky_decomp(const double* cov, int P, double* L) {
  for (int i = 0; i < P * P; i++) L[i] = 0.0;
  for (int i = 0; i < P; i++)
    for (int k = 0; k <= i; k++) {
      double tmp = 0.0;
      for (int j = 0; j < k; j++) tmp = tmp + L[i * P + j] * L[k * P + j];
      if (i == k) {
        double v = cov[i * P + k] - tmp;
        L[i * P + k] = std::sqrt(v > 0.0 ? v : (v != v ? v : 0.0));  // (max 0d0 v); NaN stays NaN
      } else
        L[i * P + k] = (cov[i * P + k] - tmp) / L[k * P + k];
    }
}

extern "C" void orc_scale_array(double s, double* m, int count) {  // :605-611
  for (int i = 0; i < count; i++) m[i] = s * m[i];
}

// get-covariant-sample, :679-700: r_i = (sum_j L_ij z_j, from 0d0, j ascending, full row) + theta_i
extern "C" void orc_get_covariant_sample(const double* means, const double* L, int P, const double* z, double* out) {
  for (int i = 0; i < P; i++) {
    double s = 0.0;
    for (int j = 0; j < P; j++) s = s + L[i * P + j] * z[j];
    out[i] = s + means[i];
  }
}

// diff-lplist + lplist-to-l-matrix, :277-280, :263-275, :675-677
extern "C" void orc_lplist_to_l_matrix(const double* rows, int n, int P, double* L) {
  std::vector<double> diffs((size_t)(n > 1 ? n - 1 : 0) * P);
  for (int k = 0; k + 1 < n; k++)
    for (int i = 0; i < P; i++) diffs[(size_t)k * P + i] = rows[(size_t)(k + 1) * P + i] - rows[(size_t)k * P + i];
  std::vector<double> cov((size_t)P * P);
  orc_lplist_covariance(diffs.data(), n - 1, P, cov.data());
  orc_cholesky_decomp(cov.data(), P, L);
}

// --------------------------------------------------------------- built-in families
// Natural-form definitions (the Lisp a user of the reference would write); canonical
// s-expression texts live in lisp_mcmc_b200/families.py and tests check both agree bit for bit.
extern "C" double orc_model_builtin(int family, int degree, const double* th, double x) {
  switch (family) {
    case ORC_FAM_POLY: {  // (+ c0 (* x (+ c1 (* x (+ c2 ...)))))
      double r = th[degree];
      for (int k = degree - 1; k >= 0; k--) r = th[k] + x * r;
      return r;
    }
    case ORC_FAM_EXPO:  // (+ (* a (exp (* b x))) c)
      return th[0] * std::exp(th[1] * x) + th[2];
    case ORC_FAM_SINUS:  // (+ (* a (sin (+ (* omega x) phi))) c)
      return th[0] * std::sin(th[1] * x + th[2]) + th[3];
    case ORC_FAM_LORENTZ: {  // (+ (/ (* a (* gamma gamma)) (+ (* (- x x0) (- x x0)) (* gamma gamma))) c)
      double g2 = th[2] * th[2], d = x - th[1];
      return (th[0] * g2) / (d * d + g2) + th[3];
    }
    case ORC_FAM_GAUSS_LIN: {  // (+ (* a (exp (* -1/2 (expt (/ (- x mu) s) 2)))) b0 (* b1 x))
      double z = (x - th[1]) / th[2];
      double e = std::exp(-0.5 * (z * z));
      return (th[0] * e + th[3]) + th[4] * x;
    }
    case ORC_FAM_NV_MULTI: {  // (- bg0 (/ (* s1 (* g g)) (+ (* (- x mu1) (- x mu1)) (* g g))) (/ (* s2 ...) ...) ...)
      int K = degree;  // theta = scale_1..K, mu_1..K, sigma, bg0
      double g2 = th[2 * K] * th[2 * K];
      double r = th[2 * K + 1];
      for (int k = 0; k < K; k++) {
        double d = x - th[K + k];
        r = r - (th[k] * g2) / (d * d + g2);
      }
      return r;
    }
  }
  return NAN;
}

// --------------------------------------------------------------- temperature schedule
extern "C" long orc_temp_steps(long n, int P) {
  long S = 10L * (P > 50 ? P : 50);
  return n > 10 * S ? n : 10 * S;
}
// :875-878: (max 1 (* (cos (* x pi (+ 1 (* 2 (floor temp-steps 5000))) (/ (* 2 temp-steps)))) temperature))
extern "C" void orc_temperature_schedule(long n, int P, double temperature, double* temps) {
  const double pi = 3.141592653589793;
  long Tn = orc_temp_steps(n, P);
  double k = (double)(1 + 2 * (Tn / 5000));
  double inv = 1.0 / (double)(2 * Tn);  // the ratio 1/(2Tn) coerced to double
  for (long x = 0; x < Tn; x++) {
    double a = (((double)x * pi) * k) * inv;
    double t = std::cos(a) * temperature;
    temps[x] = t > 1.0 ? t : 1.0;
  }
}

// --------------------------------------------------------------- walker
namespace {

struct Dataset {
  std::vector<double> x, y, sigma;
  int xd = 1;                    // components of every x (multi-dimensional inputs, walker-create docstring :1137)
  std::vector<double> xv;        // xd > 1: [n][xd], the components of point i; `x` then holds the first component
  int family = ORC_FAM_POLY, degree = 1;
  std::vector<int> idx;
  orc::Fn* fn = nullptr;        // model s-expression
  int ll_kind = ORC_LL_NORMAL;
  orc::Fn* ll_fn = nullptr;     // ORC_LL_USER: the per-point (lambda (y model error) ...) of create-log-liklihood-function
  int prior_kind = 0;           // 0 flat, 1 bounds, 2 s-expression
  std::vector<int> b_idx;
  std::vector<double> b_lo, b_hi;
  orc::Fn* prior_fn = nullptr;
};

struct Step {
  double prob;
  std::vector<double> params;
};

}  // namespace

struct orc_walker {
  int D = 0, P = 0;
  std::vector<Dataset> ds;
  std::vector<Step> walk;  // oldest first (the reference list is newest first)
  long length = 0, age = 0;
  Step most_likely, last;
  std::vector<double> lastL;
  std::string err;
  std::mt19937_64 rng{12345};
  std::vector<std::vector<std::string>> keep_keys;
  ~orc_walker() {
    for (auto& d : ds) { if (d.fn) orc::fn_free(d.fn); if (d.prior_fn) orc::fn_free(d.prior_fn); if (d.ll_fn) orc::fn_free(d.ll_fn); }
  }
};

namespace {

double model_eval(const Dataset& d, const double* theta, double x) {
  if (d.family == ORC_FAM_SEXPR) return orc::fn_call_model(d.fn, x, theta);
  double th[64];
  for (size_t k = 0; k < d.idx.size(); k++) th[k] = theta[d.idx[k]];
  return orc_model_builtin(d.family, d.degree, th, x);
}

struct Summer {  // reference order, or Neumaier-compensated for the "how far is the reference from exact" column
  int mode;
  bool first = true;
  double s = 0.0, c = 0.0;
  explicit Summer(int m) : mode(m) {}
  void add(double t) {
    if (first) { s = t; first = false; return; }  // (reduce #'+ list) starts from the first element
    if (mode == ORC_SUM_REFERENCE) { s = s + t; return; }
    double u = s + t;
    if (std::fabs(s) >= std::fabs(t)) c += (s - u) + t; else c += (t - u) + s;
    s = u;
  }
  double get() const { return mode == ORC_SUM_REFERENCE ? s : s + c; }
};

// a4 log-liklihood-normal :393-400, a6 -cutoff :419-427, a5 create-log-liklihood-function :402-417 with
// (lambda (y model error) (log-poisson model y))
double dataset_ll(const Dataset& d, const double* theta, int mode) {
  Summer sum(mode);
  size_t n = d.x.size();
  if (n == 0) return 0.0;  // (reduce #'+ nil) = 0
  for (size_t i = 0; i < n; i++) {
    double f = d.xd > 1 ? (d.family == ORC_FAM_SEXPR ? orc::fn_call_model_v(d.fn, &d.xv[i * d.xd], d.xd, theta) : NAN)
                        : model_eval(d, theta, d.x[i]);
    double t;
    switch (d.ll_kind) {
      case ORC_LL_NORMAL: t = orc_log_normal(d.y[i], f, d.sigma[i]); break;
      case ORC_LL_NORMAL_CUTOFF: {
        double v = orc_log_normal(d.y[i], f, d.sigma[i]);
        t = (v > -5000.0) ? v : (v != v ? v : -5000.0);  // (max -5000d0 v)
        break;
      }
      case ORC_LL_USER: t = orc::fn_call_ll(d.ll_fn, d.y[i], f, d.sigma.data(), (int)d.sigma.size()); break;  // :416
      default: t = orc_log_poisson(f, d.y[i]); break;
    }
    sum.add(t);
  }
  return sum.get();
}

// a7 prior-bounds-let with constant bounds and body = bounds-total, :346-369
double bounds_prior(const Dataset& d, const double* theta) {
  double total = 0.0;
  bool first = true;
  for (size_t k = 0; k < d.b_idx.size(); k++) {
    double p = d.b_idx[k] >= 0 ? theta[d.b_idx[k]] : 0.0;  // (getf params key 0d0)
    double lo = d.b_lo[k], hi = d.b_hi[k], b;
    if (lo < p && p < hi)
      b = 0.0;
    else {
      double a1 = std::fabs(p - hi), a2 = std::fabs(p - lo);
      double m = a1;               // (min a1 a2): returns a2 only when a2 < a1
      if (a2 < a1) m = a2;
      b = -1e10 * (std::exp(m * 1e-5) - 1.0);
    }
    if (first) { total = b; first = false; } else total = total + b;
  }
  return total;
}

double dataset_lp(const Dataset& d, const double* theta) {
  switch (d.prior_kind) {
    case 0: return 0.0;  // log-prior-flat :340-343
    case 1: return bounds_prior(d, theta);
    default: return orc::fn_call_prior(d.prior_fn, theta);
  }
}

// walker-make-step :1067-1070
double make_step(orc_walker* w, const double* theta, int mode, double* ll_d, double* lp_d) {
  double ll = 0.0, lp = 0.0;
  for (int d = 0; d < w->D; d++) {
    double v = dataset_ll(w->ds[d], theta, mode);
    if (ll_d) ll_d[d] = v;
    ll = d == 0 ? v : ll + v;
  }
  for (int d = 0; d < w->D; d++) {
    double v = dataset_lp(w->ds[d], theta);
    if (lp_d) lp_d[d] = v;
    lp = d == 0 ? v : lp + v;
  }
  return ll + lp;
}

// walker-modify :add-step :549-555
void add_step(orc_walker* w, const Step& s) {
  w->walk.push_back(s);
  w->last = s;
  w->length += 1;
  w->age += 1;
  if (s.prob > w->most_likely.prob) w->most_likely = s;
}

// walker-get :steps :take -> window [lo, size) of the oldest-first vector, newest = back
size_t window_lo(orc_walker* w, long take) {
  size_t n = w->walk.size();
  if (take <= 0) return 0;
  long m = take < w->length ? take : w->length;  // (min (walker-length walker) take)
  if ((size_t)m > n) m = (long)n;
  return n - (size_t)m;
}

// walker-get :acceptance :506-508 with remove-consecutive-duplicates :220-223 (eql on the probs)
void acceptance(orc_walker* w, long take, long* count, long* len) {
  size_t lo = window_lo(w, take), n = w->walk.size();
  long c = 0;
  for (size_t j = n; j-- > lo;) {  // newest first; element kept unless eql to the next older one
    if (j == lo) { c++; break; }
    uint64_t a, b;
    memcpy(&a, &w->walk[j].prob, 8);
    memcpy(&b, &w->walk[j - 1].prob, 8);
    if (a != b) c++;
  }
  *count = c;
  *len = (long)(n - lo);
}

// exact (< count/len thr) with thr a float literal such as 0.2 (single-float 0.2 = 0.20000000298...)
bool acc_lt(long count, long len, float thr) { return (double)count < (double)thr * (double)len; }
bool acc_gt(long count, long len, float thr) { return (double)count > (double)thr * (double)len; }

struct Stream {
  const double* z;
  const double* u;
  long n, t = 0;
  int P;
};

void polar_normal(orc_walker* w, double* out) {  // alexandria:gaussian-random (Marsaglia polar), first value only
  std::uniform_real_distribution<double> U(0.0, 2.0);
  for (;;) {
    double x1 = U(w->rng) - 1.0, x2 = U(w->rng) - 1.0;
    double s = x1 * x1 + x2 * x2;
    if (s < 1.0 && s > 0.0) {
      *out = x1 * std::sqrt(-2.0 * std::log(s) / s);
      return;
    }
  }
}

// walker-take-step :1072-1095 (greedy = walker-pretend-take-step :1097-1122)
bool take_step(orc_walker* w, const double* L, double temperature, Stream& st, bool greedy) {
  int P = w->P;
  double zbuf[64];
  std::vector<double> zbig;
  double* z = zbuf;
  if (P > 64) { zbig.resize(P); z = zbig.data(); }
  double u;
  if (st.z) {
    if (st.t >= st.n) { w->err = "injected stream exhausted"; return false; }
    for (int j = 0; j < P; j++) z[j] = st.z[(size_t)st.t * P + j];
    u = st.u ? st.u[st.t] : 0.5;
  } else {
    for (int j = 0; j < P; j++) polar_normal(w, &z[j]);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    u = U(w->rng);
    if (u <= 0.0) u = std::numeric_limits<double>::min();
  }
  st.t++;
  const Step& prev = w->last;
  double p0 = prev.prob;
  Step next;
  next.params.resize(P);
  orc_get_covariant_sample(prev.params.data(), L, P, z, next.params.data());
  next.prob = make_step(w, next.params.data(), ORC_SUM_REFERENCE, nullptr, nullptr);
  double p1 = next.prob;
  bool accept;
  if (greedy)
    accept = p1 > p0;  // :1117
  else
    accept = (p1 > p0) || ((p1 - p0) / temperature > std::log(u));  // :1091-1092
  if (accept) add_step(w, next);
  else { Step keep = prev; add_step(w, keep); }
  return true;
}

void diag_of(const std::vector<double>& v, int P, std::vector<double>& L) {  // diagonal-covariance :710-727
  L.assign((size_t)P * P, 0.0);
  for (int i = 0; i < P; i++) L[(size_t)i * P + i] = v[i];
}

// get-optimal-mcmc-l-matrix :888-895; returns false if dimension check :936 fails (keep l-matrix)
bool get_optimal(orc_walker* w, long take, int sampling_opt, std::vector<double>& L, std::vector<double>& out) {
  int P = w->P;
  if (sampling_opt == ORC_OPT_BEST_VALUE) {  // :895
    diag_of(w->most_likely.params, P, out);
    orc_scale_array(1e-5, out.data(), P * P);
    return true;
  }
  double scale = (2.38 * 2.38) / (double)P;  // (/ (expt 2.38d0 2) num-params)
  // walker-get :forward-steps :497-502, newest first
  size_t lo = window_lo(w, take), n = w->walk.size();
  std::vector<double> rows;
  int m = 0;
  for (size_t j = n; j-- > lo + 1;) {
    if (w->walk[j].prob <= w->walk[j - 1].prob) continue;
    if (!(w->walk[j].prob > w->walk[j - 1].prob)) continue;  // NaN: (<= a b) false in CL would trap; treat as not forward
    rows.insert(rows.end(), w->walk[j].params.begin(), w->walk[j].params.end());
    m++;
  }
  if (m == 0) {  // (elt nil 0) -> type-error -> handler returns l-matrix, which scale-array mutates
    orc_scale_array(scale, L.data(), P * P);
    out = L;
    return true;
  }
  if (m == 1) return false;  // 0x0 result fails (= num-params (array-dimension tmp 0)) :936
  std::vector<double> Lnew((size_t)P * P);
  orc_lplist_to_l_matrix(rows.data(), m, P, Lnew.data());
  bool ok = true;
  for (double v : Lnew)
    if (!std::isfinite(v)) ok = false;
  if (!ok) {  // FPO / DB0 handlers :892-893 (0/0 would be an unhandled invalid-operation in SBCL; same fallback here)
    orc_scale_array(scale, L.data(), P * P);
    out = L;
    return true;
  }
  orc_scale_array(scale, Lnew.data(), P * P);
  out = Lnew;
  return true;
}

// stable-probs-p :880-885 on (walker-get :log-liklihoods :take S), newest first
bool stable_probs(orc_walker* w, long S) {
  size_t lo = window_lo(w, S), n = w->walk.size();
  size_t cnt = n - lo;
  if (cnt < 200) return false;
  double early = w->walk[n - 1].prob;  // (subseq probs 0 200): the 200 newest
  for (size_t j = n - 200; j < n; j++) early = std::fmax(early, w->walk[j].prob);
  double late = w->walk[lo].prob;      // the 200 oldest of the window
  for (size_t j = lo; j < lo + 200; j++) late = std::fmax(late, w->walk[j].prob);
  double mn = w->walk[lo].prob;
  for (size_t j = lo; j < n; j++) mn = std::fmin(mn, w->walk[j].prob);
  double spread = early - mn;
  return (std::fabs(early - late) < 0.5) && (4.0 < spread && spread < 9.0);
}

}  // namespace

extern "C" {

orc_walker* orc_walker_new(int D, int P) {
  orc_walker* w = new orc_walker();
  w->D = D;
  w->P = P;
  w->ds.resize(D);
  w->most_likely.prob = -std::numeric_limits<double>::max();  // most-negative-double-float :463
  w->last.prob = -std::numeric_limits<double>::max();
  return w;
}
void orc_walker_free(orc_walker* w) { delete w; }
const char* orc_walker_error(orc_walker* w) { return w->err.c_str(); }
void orc_walker_seed(orc_walker* w, unsigned long long seed) { w->rng.seed(seed); }

int orc_walker_dataset(orc_walker* w, int d, const double* x, const double* y, const double* sigma, long n) {
  if (d < 0 || d >= w->D) return -1;
  Dataset& s = w->ds[d];
  s.xd = 1;
  s.xv.clear();
  s.x.assign(x, x + n);
  s.y.assign(y, y + n);
  if (sigma) s.sigma.assign(sigma, sigma + n);
  else s.sigma.assign(n, 1.0);  // (if data-error data-error 1) :1144
  return 0;
}
// :data whose x elements are themselves sequences of xd numbers: x[i*xd + k] = component k of point i
int orc_walker_dataset_xd(orc_walker* w, int d, const double* x, int xd, const double* y, const double* sigma, long n) {
  if (d < 0 || d >= w->D || xd < 1) return -1;
  Dataset& s = w->ds[d];
  s.xd = xd;
  s.xv.assign(x, x + (size_t)n * xd);
  s.x.resize(n);
  for (long i = 0; i < n; i++) s.x[i] = x[(size_t)i * xd];
  s.y.assign(y, y + n);
  if (sigma) s.sigma.assign(sigma, sigma + n);
  else s.sigma.assign(n, 1.0);
  return 0;
}
int orc_walker_model_builtin(orc_walker* w, int d, int family, int degree, const int* idx, int n_idx) {
  if (d < 0 || d >= w->D) return -1;
  w->ds[d].family = family;
  w->ds[d].degree = degree;
  w->ds[d].idx.assign(idx, idx + n_idx);
  return 0;
}
static int parse_into(orc_walker* w, orc::Fn** slot, const char* text, const char* const* keys) {
  try {
    std::vector<std::string> ks;
    for (int i = 0; i < w->P; i++) ks.push_back(keys[i]);
    orc::Fn* f = orc::fn_parse(text, ks);
    if (*slot) orc::fn_free(*slot);
    *slot = f;
    return 0;
  } catch (std::exception& e) {
    w->err = e.what();
    return -2;
  }
}
int orc_walker_model_sexpr(orc_walker* w, int d, const char* text, const char* const* keys) {
  if (d < 0 || d >= w->D) return -1;
  int rc = parse_into(w, &w->ds[d].fn, text, keys);
  if (rc == 0) w->ds[d].family = ORC_FAM_SEXPR;
  return rc;
}
int orc_walker_likelihood(orc_walker* w, int d, int kind) {
  if (d < 0 || d >= w->D) return -1;
  w->ds[d].ll_kind = kind;
  return 0;
}
// :log-liklihood = (create-log-liklihood-function (lambda (y model error) ...)) :402-417
int orc_walker_likelihood_sexpr(orc_walker* w, int d, const char* text) {
  if (d < 0 || d >= w->D) return -1;
  try {
    orc::Fn* f = orc::fn_parse(text, std::vector<std::string>());
    if (w->ds[d].ll_fn) orc::fn_free(w->ds[d].ll_fn);
    w->ds[d].ll_fn = f;
    w->ds[d].ll_kind = ORC_LL_USER;
    return 0;
  } catch (std::exception& e) {
    w->err = e.what();
    return -2;
  }
}
int orc_walker_prior_flat(orc_walker* w, int d) { w->ds[d].prior_kind = 0; return 0; }
int orc_walker_prior_bounds(orc_walker* w, int d, int n, const int* idx, const double* lo, const double* hi) {
  Dataset& s = w->ds[d];
  s.prior_kind = 1;
  s.b_idx.assign(idx, idx + n);
  s.b_lo.assign(lo, lo + n);
  s.b_hi.assign(hi, hi + n);
  return 0;
}
int orc_walker_prior_sexpr(orc_walker* w, int d, const char* text, const char* const* keys) {
  int rc = parse_into(w, &w->ds[d].prior_fn, text, keys);
  if (rc == 0) w->ds[d].prior_kind = 2;
  return rc;
}

int orc_walker_logp(orc_walker* w, const double* theta, int sum_mode, double* prob, double* ll_d, double* lp_d) {
  try {
    *prob = make_step(w, theta, sum_mode, ll_d, lp_d);
    return 0;
  } catch (std::exception& e) { w->err = e.what(); return -3; }
}

int orc_walker_init(orc_walker* w, const double* theta0) {
  try {
    Step s;
    s.params.assign(theta0, theta0 + w->P);
    s.prob = make_step(w, theta0, ORC_SUM_REFERENCE, nullptr, nullptr);
    w->walk.clear();
    w->walk.push_back(s);
    w->last = s;
    w->most_likely = s;
    w->length = 1;  // :472-473 defaults
    w->age = 1;
    return 0;
  } catch (std::exception& e) { w->err = e.what(); return -3; }
}

// walker-adaptive-steps-full :862-942.  max_steps >= 0 interrupts the loop after that many take-step calls,
// like setting mfit-walker-estop (:860, :904) from another thread; -1 = run to the end.
static int adaptive_impl(orc_walker* w, const orc_adapt_cfg* cfg, const double* l_matrix_or_null, const double* z,
                         const double* u, long nstream, long max_steps, long* steps_taken) {
  try {
    int P = w->P;
    long n = cfg->n;
    long reset_index = 10000;
    long mwl = cfg->max_walker_length > 0 ? cfg->max_walker_length / 2 : 0;  // (floor max-walker-length 2)
    long S = 10L * (P > 50 ? P : 50);
    bool shutting = false;
    long Tn = n > 10 * S ? n : 10 * S;
    std::vector<double> temps(Tn);
    orc_temperature_schedule(n, P, cfg->temperature, temps.data());
    double temperature = cfg->temperature;
    long tail = S > 2000 ? S : 2000;
    std::vector<double> L;
    Stream st{z, u, nstream, 0, P};
    if (l_matrix_or_null) L.assign(l_matrix_or_null, l_matrix_or_null + (size_t)P * P);
    else {  // :896-901
      long c, len;
      acceptance(w, 100, &c, &len);
      diag_of(w->most_likely.params, P, L);
      if (!(w->length < S || acc_lt(c, len, 0.1f))) {
        std::vector<double> out;
        if (get_optimal(w, S, cfg->sampling_optimization, L, out)) L = out;
      }
    }
    for (long i = 1; i < n; i++) {
      if (max_steps >= 0 && st.t >= max_steps) break;  // estop
      bool trig = (!shutting && (n - i) < tail);
      if (!trig && cfg->auto_mode != ORC_AUTO_NONE && !shutting && (i % 1000 == 0) && i > 2 * S) {
        long c, len;
        acceptance(w, 1000, &c, &len);
        if (acc_gt(c, len, 0.2f) && (double)c < 0.5 * (double)len) trig = stable_probs(w, S);
      }
      if (trig) {  // :915-917
        temperature = 1.0;
        shutting = true;
        i = n - tail;
      }
      if (!take_step(w, L.data(), temperature, st, cfg->greedy != 0)) return -4;
      if (!shutting && i < Tn) temperature = temps[i];  // :920-921
      if (mwl && i == reset_index) {                    // :923-927
        if (w->length > mwl) {
          orc_walker_keep_walks(w, mwl);
          reset_index += 1 + mwl;
        } else
          reset_index += 1 + (mwl - w->length);
      }
      if (i > 0) {  // :929-942
        bool enter = false;
        long c = 0, len = 1;
        if (i % 200 == 0) {
          acceptance(w, 200, &c, &len);
          if (acc_lt(c, len, 0.2f) || acc_gt(c, len, 0.4f)) enter = true;
        }
        if (!enter && !shutting && (i % (2 * S) == 0)) enter = true;
        if (enter) {
          acceptance(w, 200, &c, &len);
          if (acc_gt(c, len, 0.2f) && acc_lt(c, len, 0.4f)) {
            std::vector<double> out;
            if (get_optimal(w, S, cfg->sampling_optimization, L, out)) L = out;
          } else if (acc_lt(c, len, 0.2f))
            orc_scale_array(0.1, L.data(), P * P);
          else if (acc_gt(c, len, 0.4f))
            orc_scale_array(1.9, L.data(), P * P);
        }
      }
    }
    w->lastL = L;
    if (steps_taken) *steps_taken = st.t;
    return 0;
  } catch (std::exception& e) { w->err = e.what(); return -3; }
}

int orc_walker_adaptive_steps_full(orc_walker* w, const orc_adapt_cfg* cfg, const double* l_matrix_or_null,
                                   const double* z, const double* u, long nstream, long* steps_taken) {
  return adaptive_impl(w, cfg, l_matrix_or_null, z, u, nstream, -1, steps_taken);
}
int orc_walker_adaptive_steps_limited(orc_walker* w, const orc_adapt_cfg* cfg, const double* l_matrix_or_null,
                                      const double* z, const double* u, long nstream, long max_steps,
                                      long* steps_taken) {
  return adaptive_impl(w, cfg, l_matrix_or_null, z, u, nstream, max_steps, steps_taken);
}

// walker-many-steps :849-853 (constant l-matrix, temperature 1)
int orc_walker_many_steps(orc_walker* w, long n, const double* l_matrix_or_null, const double* z, const double* u,
                          long nstream) {
  try {
    int P = w->P;
    std::vector<double> L;
    if (l_matrix_or_null) L.assign(l_matrix_or_null, l_matrix_or_null + (size_t)P * P);
    else {
      w->err = "walker-many-steps without l-matrix needs :median-params (post-processing, out of scope)";
      return -5;
    }
    Stream st{z, u, nstream, 0, P};
    for (long i = 0; i < n; i++)
      if (!take_step(w, L.data(), 1.0, st, false)) return -4;
    w->lastL = L;
    return 0;
  } catch (std::exception& e) { w->err = e.what(); return -3; }
}

// walker-sample-region :949-969
int orc_walker_sample_region(orc_walker* w, double initial_scale, const double* z, long nstream) {
  try {
    int P = w->P;
    std::vector<double> L;
    diag_of(w->most_likely.params, P, L);
    orc_scale_array(initial_scale, L.data(), P * P);
    Stream st{z, nullptr, nstream, 0, P};
    for (long i = 1; i < 3000; i++) {
      if (!take_step(w, L.data(), 1.0, st, true)) return -4;
      if (i > 0 && i % 20 == 0) {
        long c, len;
        acceptance(w, 50, &c, &len);
        if (c * 50 == len) orc_scale_array(0.25, L.data(), P * P);       // (= acc 1/50)
        else if (c * 50 > 4 * len) orc_scale_array(1.7, L.data(), P * P);  // (> acc 4/50)
      }
    }
    w->lastL = L;
    return 0;
  } catch (std::exception& e) { w->err = e.what(); return -3; }
}

// walker-force-take-step :1124-1129
int orc_walker_force_take_step(orc_walker* w) {
  try {
    Step s;
    s.params = w->last.params;
    s.prob = make_step(w, s.params.data(), ORC_SUM_REFERENCE, nullptr, nullptr);
    add_step(w, s);
    return 0;
  } catch (std::exception& e) { w->err = e.what(); return -3; }
}

long orc_walker_length(orc_walker* w) { return w->length; }
long orc_walker_age(orc_walker* w) { return w->age; }

long orc_walker_trace(orc_walker* w, long take, double* theta, double* prob) {
  size_t lo = window_lo(w, take), n = w->walk.size();
  long r = 0;
  for (size_t j = n; j-- > lo; r++) {
    if (theta) memcpy(theta + (size_t)r * w->P, w->walk[j].params.data(), sizeof(double) * w->P);
    if (prob) prob[r] = w->walk[j].prob;
  }
  return r;
}
void orc_walker_most_likely(orc_walker* w, double* theta, double* prob) {
  if (theta) memcpy(theta, w->most_likely.params.data(), sizeof(double) * w->P);
  if (prob) *prob = w->most_likely.prob;
}
void orc_walker_last_l_matrix(orc_walker* w, double* L) {
  if (w->lastL.size() == (size_t)w->P * w->P) memcpy(L, w->lastL.data(), sizeof(double) * w->P * w->P);
}
void orc_walker_acceptance(orc_walker* w, long take, long* count, long* len) { acceptance(w, take, count, len); }

// test support: replace the walk by a given one (newest first, like walker-get :steps), so that the trace analytics
// can be checked on exactly the walk another implementation produced
extern "C" int orc_walker_load_walk(orc_walker* w, const double* theta, const double* prob, long n) {
  w->walk.clear();
  for (long r = n; r-- > 0;) {
    Step s;
    s.prob = prob[r];
    s.params.assign(theta + (size_t)r * w->P, theta + (size_t)(r + 1) * w->P);
    w->walk.push_back(s);
  }
  w->length = n;
  if (n > 0) w->last = w->walk.back();
  return 0;
}

// walker-get-data-and-fit :1231-1255.  x_fit is (linspace (reduce #'min x-data) (reduce #'max x-data) :len 1000),
// made by the caller (exact rationals, :235-247: oracle.py linspace).
//   take          clamped to the walker length (:1233-1235); <= 0 means nil
//   params        which_solution 0: (walker-get :most-likely-step :take take) = reduce keeping x only when strictly
//                 greater, i.e. of equal maxima the OLDEST (:503-505); 1: :median-params = nth-percentile 50 (:516-523,
//                 :1493-1504: position 50 (len-1)/100, the MEAN of the two neighbours when fractional)
//   y_fit         (apply fn x params) over x-fit (:1246)
//   sorted-walks  (subseq (sort (walker-get :steps) #'> :key prob) 0 (ceiling (* 0.66 take))) (:1248-1249): the WHOLE
//                 walk sorted, 0.66 a single-float, the product formed in single precision
//   max_ys/min_ys (reduce #'max y) / (reduce #'min y) over those steps at every x (:1250-1252)
// x-shift / y-shift are plain additions done by the caller.
extern "C" int orc_walker_data_and_fit(orc_walker* w, long take, int d, int which_solution, const double* x_fit,
                                       long n_fit, double* y_fit, double* max_ys, double* min_ys, double* params) {
  if (d < 0 || d >= w->D) return -1;
  const size_t n = w->walk.size();
  if (n == 0) return -2;
  long tk = (take <= 0 || take > w->length) ? w->length : take;
  if ((size_t)tk > n) tk = (long)n;
  const size_t lo = n - (size_t)tk;
  std::vector<double> par(w->P);
  if (which_solution == 0) {
    size_t best = n - 1;                       // x = newest; y walks towards older steps
    for (size_t j = n - 1; j-- > lo;)
      if (!(w->walk[best].prob > w->walk[j].prob)) best = j;
    par = w->walk[best].params;
  } else {
    std::vector<double> col((size_t)tk);
    for (int i = 0; i < w->P; i++) {
      for (size_t j = lo; j < n; j++) col[j - lo] = w->walk[j].params[i];
      std::sort(col.begin(), col.end());
      const long num = 50 * (tk - 1), pos = num / 100, rem = num % 100;
      par[i] = rem == 0 ? col[pos] : (col[pos] + col[pos + 1]) / 2.0;
    }
  }
  if (params) memcpy(params, par.data(), sizeof(double) * w->P);
  const Dataset& ds = w->ds[d];
  for (long i = 0; i < n_fit; i++) y_fit[i] = model_eval(ds, par.data(), x_fit[i]);
  const float prod = 0.66f * (float)tk;        // (* 0.66 take): single-float contagion
  const long M = (long)std::ceil(prod);
  if (M < 1 || (size_t)M > n) return -3;       // subseq would signal an error
  std::vector<size_t> order(n);
  for (size_t j = 0; j < n; j++) order[j] = n - 1 - j;  // list order: newest first
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return w->walk[a].prob > w->walk[b].prob; });
  for (long i = 0; i < n_fit; i++) {
    double mx = 0.0, mn = 0.0;
    for (long k = 0; k < M; k++) {
      const double v = model_eval(ds, w->walk[order[k]].params.data(), x_fit[i]);
      if (k == 0) { mx = v; mn = v; }
      else { if (v > mx) mx = v; if (v < mn) mn = v; }
    }
    max_ys[i] = mx;
    min_ys[i] = mn;
  }
  return 0;
}

// walker-modify :keep-walks :568-569 -- keeps the NEWEST keep steps
int orc_walker_keep_walks(orc_walker* w, long keep) {
  if ((size_t)keep < w->walk.size()) w->walk.erase(w->walk.begin(), w->walk.end() - keep);
  w->length = keep;
  return 0;
}

// walker-modify :burn-walks :566-567 (drops the OLDEST burn-number steps), :reset :570-573 ((last walk): the walk shrinks
// to its OLDEST step, which also becomes the last step), :reset-to-most-likely :574-578
int orc_walker_modify(orc_walker* w, int mode, long number) {
  switch (mode) {
    case 1: {  // :burn-walks
      const long keep = w->length - number;
      if (keep < 0) return -1;  // subseq would signal an error
      return orc_walker_keep_walks(w, keep);
    }
    case 2: return orc_walker_keep_walks(w, number);
    case 3: {
      if (w->walk.empty()) return -1;
      Step s = w->walk.front();
      w->walk.assign(1, s);
      w->last = s;
      w->length = 1;
      return 0;
    }
    case 4: {
      w->walk.assign(1, w->most_likely);
      w->last = w->most_likely;
      w->length = 1;
      return 0;
    }
  }
  return -1;
}

// CPU baseline: W independent walkers, chains split across threads
double orc_bench_many(orc_walker* templ, const double* theta0, long W, const orc_adapt_cfg* cfg, int n_threads,
                      unsigned long long seed, double* last_prob, long max_steps) {
  if (n_threads < 1) n_threads = 1;
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; t++) {
    pool.emplace_back([=]() {
      for (long c = t; c < W; c += n_threads) {
        orc_walker w;
        w.D = templ->D;
        w.P = templ->P;
        w.ds = templ->ds;  // shares the Fn pointers (read-only during evaluation)
        w.most_likely.prob = -std::numeric_limits<double>::max();
        w.rng.seed(seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(c + 1));
        orc_walker_init(&w, theta0 + (size_t)c * w.P);
        long taken = 0;
        adaptive_impl(&w, cfg, nullptr, nullptr, nullptr, 0, max_steps, &taken);
        if (last_prob) last_prob[c] = w.last.prob;
        for (auto& d : w.ds) { d.fn = nullptr; d.prior_fn = nullptr; }  // owned by the template
      }
    });
  }
  for (auto& th : pool) th.join();
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
