// Accuracy harness for lisp_mcmc_b200/csrc/mfit_math.cuh: compiles the device header for the host (the arithmetic
// is plain IEEE double + FMA) and measures the error in ulps against long double libm.
//   g++ -O2 -std=c++17 -mfma -ffp-contract=off -DMFIT_HOST_TEST -I lisp_mcmc_b200/csrc tools/check_math.cpp -o /tmp/check_math
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "mfit_math.cuh"
static MfitGuard G;
static double mfit_exp(double x){return mfit_exp_t<false>(x,G);}
static double mfit_log(double x){return mfit_log_t<false>(x,G);}
static double mfit_sin(double x){return mfit_sin_t<false>(x,G);}
static double mfit_cos(double x){return mfit_cos_t<false>(x,G);}

static double ulp_err(double got, long double ref) {
  if (std::isnan(got) && std::isnan((double)ref)) return 0;
  if (std::isinf(got) || std::isinf((double)ref)) return got == (double)ref ? 0 : 1e30;
  double r = (double)ref;
  if (r == 0.0) return std::fabs(got) < 1e-300 ? 0 : 1e30;
  int e;
  std::frexp(r, &e);
  long double u = std::ldexp(1.0L, e - 53);
  return (double)(fabsl((long double)got - ref) / u);
}

int main() {
  std::mt19937_64 g(1);
  auto U = [&](double a, double b) { return a + (b - a) * std::uniform_real_distribution<double>(0, 1)(g); };
  double worst;
  const int N = 4000000;
  worst = 0;
  for (int i = 0; i < N; i++) { double x = U(-708.3, 709.7); worst = std::fmax(worst, ulp_err(mfit_exp(x), expl((long double)x))); }
  for (int i = 0; i < N; i++) { double x = U(-2, 2); worst = std::fmax(worst, ulp_err(mfit_exp(x), expl((long double)x))); }
  printf("exp   max ulp %.3f\n", worst);
  worst = 0;
  for (int i = 0; i < N; i++) {
    double z = U(-37, 37), zs = std::fabs(z) * MFIT_GAUSS_ZSCALE, s, p;   // the routine takes |zs| (mfit_gauss_reduce)
    mfit_gauss_parts(zs, &s, &p, G);
    // reference for what the routine promises: exp(-zs^2 ln2/256) of the rounded zs
    long double ref = expl(-(long double)zs * zs * (logl(2.0L) / (1 << MFIT_EXP_BITS)));
    worst = std::fmax(worst, ulp_err(s * p, ref));
  }
  printf("gauss max ulp %.3f\n", worst);
  { double s, p; G.reset(); mfit_gauss_parts(37.0 * MFIT_GAUSS_ZSCALE, &s, &p, G); printf("guard gauss(37) tripped=%d ; ", (int)G.tripped());
    G.reset(); mfit_gauss_parts(37.7 * MFIT_GAUSS_ZSCALE, &s, &p, G); printf("gauss(37.7)=%g tripped=%d ; ", s * p, (int)G.tripped());
    G.reset(); mfit_gauss_parts(1e200, &s, &p, G); printf("gauss(1e200)=%g tripped=%d ; ", s * p, (int)G.tripped());
    G.reset(); mfit_gauss_parts(INFINITY, &s, &p, G); printf("gauss(inf)=%g tripped=%d ; ", s * p, (int)G.tripped());
    G.reset(); mfit_gauss_parts(NAN, &s, &p, G); printf("gauss(nan) tripped=%d\n", (int)G.tripped()); }
  worst = 0;
  for (int i = 0; i < N; i++) { double x = std::exp(U(-700, 700)); worst = std::fmax(worst, ulp_err(mfit_log(x), logl((long double)x))); }
  for (int i = 0; i < N; i++) { double x = U(0.5, 2.0); worst = std::fmax(worst, ulp_err(mfit_log(x), logl((long double)x))); }
  for (int i = 0; i < N; i++) { double x = 1.0 + U(-1e-3, 1e-3); worst = std::fmax(worst, ulp_err(mfit_log(x), logl((long double)x))); }
  for (int i = 0; i < N; i++) { double x = U(0.0, 200.0); worst = std::fmax(worst, ulp_err(mfit_log(x), logl((long double)x))); }
  printf("log   max ulp %.3f\n", worst);
  { double xs[] = {1.0, 0.0, -1.0, INFINITY, NAN, 1e-310, 2.3e-308, 1.7e308};
    for (double x : xs) { G.reset(); double v = mfit_log(x); printf("  log(%g)=%g tripped=%d", x, v, (int)G.tripped()); } printf("\n"); }
  double worst_abs = 0;
  worst = 0;
  for (int i = 0; i < N; i++) {
    double x = U(-100, 100);
    long double rs = sinl((long double)x), rc = cosl((long double)x);
    double es = std::fabs((double)((long double)mfit_sin(x) - rs)), ec = std::fabs((double)((long double)mfit_cos(x) - rc));
    worst_abs = std::fmax(worst_abs, std::fmax(es, ec));
    worst = std::fmax(worst, ulp_err(mfit_sin(x), rs));   // sin keeps relative accuracy through its zeros
  }
  printf("sincos |x|<100: max abs err %.3g (cos is absolute-accurate only: it is sin(pi/2 - |r|)), sin max ulp %.3f\n", worst_abs, worst);
  worst_abs = 0;
  for (int i = 0; i < N; i++) {
    double x = U(-9e4, 9e4);
    worst_abs = std::fmax(worst_abs, std::fabs((double)((long double)mfit_sin(x) - sinl((long double)x))));
    worst_abs = std::fmax(worst_abs, std::fabs((double)((long double)mfit_cos(x) - cosl((long double)x))));
  }
  printf("sincos |x|<9e4: max abs err %.3g\n", worst_abs);
  { double xs[] = {99999.0, 100001.0, INFINITY, NAN};
    for (double x : xs) { G.reset(); mfit_sin(x); printf("  sin(%g) tripped=%d", x, (int)G.tripped()); } printf("\n"); }
  { double xs[] = {707.9, 708.1, -707.9, -708.1, INFINITY, -INFINITY, NAN};
    for (double x : xs) { G.reset(); double v = mfit_exp(x); printf("  exp(%g)=%g tripped=%d", x, v, (int)G.tripped()); } printf("\n"); }
  return 0;
}
