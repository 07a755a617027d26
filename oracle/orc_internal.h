// orc_internal.h -- ORACLE internals (test infrastructure only; see mfit_oracle.h)
#ifndef ORC_INTERNAL_H
#define ORC_INTERNAL_H
#include <string>
#include <vector>
#include "mfit_oracle.h"

namespace orc {

// a Common Lisp number (or generalized boolean) as the evaluator sees it
struct Val {
  enum K { RAT, SGL, DBL, BOOL } k = BOOL;
  long long num = 0, den = 1;
  float f = 0.f;
  double d = 0.0;
  bool b = false;
  static Val rat(long long n, long long d);
  static Val sgl(float x);
  static Val dbl(double x);
  static Val boolean(bool x);
  double as_double() const;
};

struct Fn;
Fn* fn_parse(const char* text, const std::vector<std::string>& keys);
double fn_call_model(Fn* f, double x, const double* theta);
double fn_call_model_v(Fn* f, const double* x, int xdim, const double* theta);
double fn_call_prior(Fn* f, const double* theta);
double fn_call_ll(Fn* f, double y, double model, const double* err, int n_err);
void fn_free(Fn* f);

}  // namespace orc
#endif
