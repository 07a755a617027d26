// lower.cpp -- s-expression -> CUDA C lowering for the user plugin slots.
//
// The reference calls a user :function as (apply fn x params) once per data point
// (mcmc-fitting.lisp:400, :416, :426) and a user :log-prior as (funcall lp params data)
// once per step (:1069).  Here the printed form of those lambdas is translated ONCE into
// CUDA C, which NVRTC compiles for sm_100a (nvrtc_glue.cpp); nothing is interpreted at
// run time.  The Lisp macro in lisp/mfit-cuda.lisp only prints the lambda and hands the
// text to mfit_model_sexpr / mfit_prior_sexpr, so the lowering it triggers is exactly what
// mfit_lower_sexpr returns and can be inspected and tested without a GPU or a Lisp.
//
// Typing is static and follows Common Lisp contagion: exact rationals are folded at
// translation time, single-float literals stay `float`, everything touched by a parameter
// or x is `double` (walker-create coerces params and data with to-double-floats, :833-835).
// n-ary + - * / associate left to right, (expt b <integer>) is SBCL's binary powering,
// irrational functions of singles/rationals are evaluated in double and rounded to single.
// prior-bounds-let (:346-369) is expanded the way the macro expands it.
//
// Documented deviations (all <= 1 ulp of the affected term): NVRTC contracts a*b+c into an
// FMA; (expt b 2d0) style double exponents with small integer values are lowered to
// multiplications instead of pow; branches of an `if` with different float formats are
// widened to the wider one.
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "mfit_internal.h"

namespace {

typedef long long i64;

struct Sx {
  enum T { LIST, SYM, NUM } t = LIST;
  std::string sym;
  // number payload
  int nk = 0;  // 0 rational, 1 single, 2 double
  i64 num = 0, den = 1;
  double fv = 0.0;
  std::vector<Sx> kids;
};

struct Err : std::runtime_error {
  explicit Err(const std::string& s) : std::runtime_error(s) {}
};

i64 gcdll(i64 a, i64 b) {
  if (a < 0) a = -a;
  if (b < 0) b = -b;
  while (b) { i64 t = a % b; a = b; b = t; }
  return a ? a : 1;
}

bool read_number(const std::string& tok, Sx& out) {
  const size_t n = tok.size();
  size_t i = 0;
  if (n == 0) return false;
  if (tok[i] == '+' || tok[i] == '-') i++;
  size_t j = i, id = 0;
  while (j < n && isdigit((unsigned char)tok[j])) { j++; id++; }
  out.t = Sx::NUM;
  if (j == n) {
    if (!id) return false;
    out.nk = 0; out.num = strtoll(tok.c_str(), nullptr, 10); out.den = 1;
    return true;
  }
  if (tok[j] == '/') {
    size_t k = j + 1, dd = 0;
    while (k < n && isdigit((unsigned char)tok[k])) { k++; dd++; }
    if (!id || !dd || k != n) return false;
    i64 a = strtoll(tok.substr(0, j).c_str(), nullptr, 10), b = strtoll(tok.substr(j + 1).c_str(), nullptr, 10);
    if (b == 0) throw Err("ratio with zero denominator: " + tok);
    i64 g = gcdll(a, b);
    out.nk = 0; out.num = a / g; out.den = b / g;
    return true;
  }
  size_t fd = 0;
  bool point = false;
  if (tok[j] == '.') {
    point = true;
    j++;
    while (j < n && isdigit((unsigned char)tok[j])) { j++; fd++; }
  }
  if (id + fd == 0) return false;
  char marker = 0;
  size_t mpos = j;
  if (j < n) {
    char c = (char)tolower((unsigned char)tok[j]);
    if (c != 'e' && c != 's' && c != 'f' && c != 'd' && c != 'l') return false;
    marker = c;
    size_t k = j + 1, ed = 0;
    if (k < n && (tok[k] == '+' || tok[k] == '-')) k++;
    while (k < n && isdigit((unsigned char)tok[k])) { k++; ed++; }
    if (!ed || k != n) return false;
  }
  if (!marker && !point) return false;
  if (!marker && fd == 0) {  // "12." is an integer to the CL reader
    out.nk = 0; out.num = strtoll(tok.c_str(), nullptr, 10); out.den = 1;
    return true;
  }
  std::string s = tok;
  if (marker) s[mpos] = 'e';
  if (marker == 'd' || marker == 'l') { out.nk = 2; out.fv = strtod(s.c_str(), nullptr); }
  else { out.nk = 1; out.fv = (double)strtof(s.c_str(), nullptr); }  // *read-default-float-format* single-float
  return true;
}

struct Reader {
  const char* p;
  void ws() {
    for (;;) {
      while (*p && isspace((unsigned char)*p)) p++;
      if (*p == ';') { while (*p && *p != '\n') p++; continue; }
      if (p[0] == '#' && p[1] == '|') {
        p += 2;
        while (*p && !(p[0] == '|' && p[1] == '#')) p++;
        if (*p) p += 2;
        continue;
      }
      return;
    }
  }
  Sx read() {
    ws();
    if (!*p) throw Err("unexpected end of text");
    if (*p == '(') {
      p++;
      Sx l;
      for (;;) {
        ws();
        if (!*p) throw Err("missing )");
        if (*p == ')') { p++; return l; }
        l.kids.push_back(read());
      }
    }
    if (*p == ')') throw Err("unexpected )");
    if (*p == '\'') { p++; return read(); }
    if (p[0] == '#' && p[1] == '\'') { p += 2; return read(); }
    if (*p == '"') {
      p++;
      while (*p && *p != '"') { if (*p == '\\' && p[1]) p++; p++; }
      if (*p) p++;
      Sx s; s.t = Sx::SYM; s.sym = "\"";
      return s;
    }
    std::string tok;
    while (*p && !isspace((unsigned char)*p) && *p != '(' && *p != ')' && *p != ';') tok.push_back(*p++);
    Sx a;
    if (read_number(tok, a)) return a;
    a.t = Sx::SYM;
    for (auto& c : tok) c = (char)toupper((unsigned char)c);
    // strip a package prefix (mfit::x, cl:exp)
    size_t colon = tok.rfind(':');
    if (colon != std::string::npos && colon > 0) tok = tok.substr(colon + 1);
    a.sym = tok;
    return a;
  }
};

// ------------------------------------------------------------------ typed expressions
enum Ty { T_RAT, T_SGL, T_DBL, T_BOOL };

struct Ex {
  Ty ty = T_DBL;
  std::string code;     // C expression (parenthesised where needed)
  i64 num = 0, den = 1; // T_RAT
  bool bconst = false, bval = false;  // T_BOOL known at translation time
  // what is known about the dependence on the model's x, for the proof Model::safe() makes about the arguments of the
  // fast math routines: 0 = does not depend on x (value = B), 1 = affine in x (value = A x + B), 2 = unknown.
  // A and B are double-valued C expressions over parameters and constants only ("" = zero).
  int xk = 2;
  std::string A, B;
};

std::string fmt_double(double v) {
  char buf[64];
  if (std::isinf(v)) return v > 0 ? "(1.0/0.0)" : "(-1.0/0.0)";
  if (std::isnan(v)) return "(0.0/0.0)";
  snprintf(buf, sizeof buf, "%.17g", v);
  std::string s = buf;
  if (s.find_first_of(".eE") == std::string::npos) s += ".0";
  if (s[0] == '-') s = "(" + s + ")";
  return s;
}
std::string fmt_float(float v) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.9g", (double)v);
  std::string s = buf;
  if (s.find_first_of(".eE") == std::string::npos) s += ".0";
  s += "f";
  if (s[0] == '-') s = "(" + s + ")";
  return s;
}

Ex mk_rat(i64 n, i64 d) {
  if (d == 0) throw Err("division by zero in a constant expression");
  if (d < 0) { n = -n; d = -d; }
  i64 g = gcdll(n, d);
  Ex e; e.ty = T_RAT; e.num = n / g; e.den = d / g;
  return e;
}
Ex mk(Ty t, const std::string& c) { Ex e; e.ty = t; e.code = c; return e; }
Ex mk_bool(bool v) { Ex e; e.ty = T_BOOL; e.bconst = true; e.bval = v; e.code = v ? "true" : "false"; return e; }

double rat_value(const Ex& e) { return (double)e.num / (double)e.den; }

// code of e converted to type `to` (T_SGL or T_DBL)
std::string as_code(const Ex& e, Ty to) {
  if (e.ty == T_BOOL) throw Err("a boolean was used where a number is required");
  if (e.ty == T_RAT) return to == T_SGL ? fmt_float((float)rat_value(e)) : fmt_double(rat_value(e));
  if (e.ty == to) return e.code;
  if (to == T_DBL) return "((double)" + e.code + ")";
  return "((float)" + e.code + ")";
}
Ty wider(Ty a, Ty b) { return a > b ? a : b; }

// (A, B) of e if it is known to be affine in x
bool affine_parts(const Ex& e, std::string& A, std::string& B) {
  if (e.ty == T_BOOL) return false;
  if (e.ty == T_RAT) { A.clear(); B = fmt_double(rat_value(e)); return true; }
  if (e.xk == 0) { A.clear(); B = e.B; return true; }
  if (e.xk == 1) { A = e.A; B = e.B; return true; }
  return false;
}
std::string aff_join(const std::string& a, char op, const std::string& b) {
  if (a.empty() && b.empty()) return "";
  if (op == '+') return a.empty() ? b : b.empty() ? a : "(" + a + " + " + b + ")";
  if (op == '-') return b.empty() ? a : a.empty() ? "(-" + b + ")" : "(" + a + " - " + b + ")";
  if (a.empty() || b.empty()) return "";            // '*' with a zero factor
  return "(" + a + " * " + b + ")";
}
// r = a op b: carry the affine form along when it survives the operation
void affine_combine(Ex& r, char op, const Ex& a, const Ex& b) {
  r.xk = 2;
  if (r.ty == T_RAT || r.ty == T_BOOL) return;
  std::string Aa, Ba, Ab, Bb;
  if (!affine_parts(a, Aa, Ba) || !affine_parts(b, Ab, Bb)) return;
  if (op == '+' || op == '-') {
    r.A = aff_join(Aa, op, Ab);
    r.B = aff_join(Ba, op, Bb);
  } else if (op == '*') {
    if (!Aa.empty() && !Ab.empty()) return;         // quadratic in x
    r.A = Aa.empty() ? aff_join(Ba, '*', Ab) : aff_join(Aa, '*', Bb);
    r.B = aff_join(Ba, '*', Bb);
  } else {
    if (!Ab.empty() || Bb.empty()) return;          // x in the divisor (or a zero divisor)
    r.A = Aa.empty() ? "" : "(" + Aa + " / " + Bb + ")";
    r.B = Ba.empty() ? "" : "(" + Ba + " / " + Bb + ")";
  }
  r.xk = r.A.empty() ? 0 : 1;
}

struct Lower {
  std::vector<std::string> keys;         // upper-case, no ':'
  std::vector<int> used;                 // theta indices referenced, in first-use order
  std::map<std::string, std::vector<Ex>> scope;  // name -> stack of bindings
  std::vector<std::string>* block = nullptr;     // statements of the block being emitted
  int tmp = 0;
  bool fast_math = false;  // model lambdas: exp/log/sin/cos -> mfit_math.cuh (prior kernels keep libdevice)
  int uses = 0;            // MFIT_USES_* tables the lowered code needs
  bool guarded = false;    // some fast routine with a restricted domain is called
  struct Check { std::string fn, A, B; };
  std::vector<Check> checks;   // arguments of fast routines that are affine in x: what safe() has to bound
  bool provable = true;        // every such argument is
  std::string x_name;      // the model's positional argument
  int x_dims = 0;          // > 0: x is a sequence and components 0 .. x_dims - 1 are read (elt / aref / nth / first ...)
  bool x_scalar = false;   // x is used as a number
  // likelihood lambdas (kind 2): (lambda (y model error) ...) of create-log-liklihood-function :402-417
  bool is_ll = false;
  std::string y_name, err_name;
  bool uses_logfact = false;   // (log-factorial y) / (log-poisson . y): the per-point table value `lf`
  int err_max = -1;            // highest index of the stddev list read

  std::string fresh(const char* stem) { return std::string(stem) + std::to_string(tmp++); }

  Ex theta(int i) {
    bool seen = false;
    for (int u : used) seen = seen || u == i;
    if (!seen) used.push_back(i);
    Ex e = mk(T_DBL, "th" + std::to_string(i));
    e.xk = 0;
    e.B = e.code;
    return e;
  }
  int key_index(std::string k) const {
    if (!k.empty() && k[0] == ':') k = k.substr(1);
    for (size_t i = 0; i < keys.size(); i++)
      if (keys[i] == k) return (int)i;
    return -1;
  }
  // :single-item parameter style (mcmc-fitting.lisp:8-14, :1078-1088, :1153-1155): ONE key whose value is a list, a vector
  // or an n x 1 array of numbers -- (lambda (x &key params &allow-other-keys) (+ (elt params 0) (* (elt params 1) x))).
  // The hosts flatten it: the key names arrive as NAME[0] .. NAME[n-1], element k of the sequence is parameter NAME[k].
  std::set<std::string> seq_vars;        // &key variables bound to such a sequence
  int seq_key_count(const std::string& name) const {
    int n = 0;
    while (key_index(name + "[" + std::to_string(n) + "]") >= 0) n++;
    return n;
  }
  // the sequence-valued key a form denotes: the &key variable itself, or (getf params :NAME) in a prior lambda
  bool seq_key_of(const Sx& seq, std::string& name) const {
    if (seq.t == Sx::SYM) {
      if (!seq_vars.count(seq.sym)) return false;
      name = seq.sym;
      return true;
    }
    if (seq.t == Sx::LIST && seq.kids.size() >= 3 && seq.kids[0].t == Sx::SYM && seq.kids[0].sym == "GETF" && seq.kids[2].t == Sx::SYM) {
      std::string k = seq.kids[2].sym;
      if (!k.empty() && k[0] == ':') k = k.substr(1);
      if (seq_key_count(k) == 0) return false;
      name = k;
      return true;
    }
    return false;
  }
  void bind(const std::string& name, const Ex& e) { scope[name].push_back(e); }
  void unbind(const std::string& name) { scope[name].pop_back(); }

  // bind `name` to the value of e through a local variable (so it is evaluated once)
  Ex materialise(const Ex& e, const char* stem) {
    if (e.ty == T_RAT || (e.ty == T_BOOL && e.bconst)) return e;
    std::string v = fresh(stem);
    const char* ty = e.ty == T_DBL ? "double" : e.ty == T_SGL ? "float" : "bool";
    block->push_back(std::string("const ") + ty + " " + v + " = " + e.code + ";");
    Ex r = e;
    r.code = v;
    return r;
  }

  // ---------------------------------------------------------------- arithmetic
  Ex arith(char op, const Ex& a, const Ex& b) {
    if (a.ty == T_BOOL || b.ty == T_BOOL) throw Err("arithmetic on a boolean");
    if (a.ty == T_RAT && b.ty == T_RAT) {
      __int128 n, d;
      switch (op) {
        case '+': n = (__int128)a.num * b.den + (__int128)b.num * a.den; d = (__int128)a.den * b.den; break;
        case '-': n = (__int128)a.num * b.den - (__int128)b.num * a.den; d = (__int128)a.den * b.den; break;
        case '*': n = (__int128)a.num * b.num; d = (__int128)a.den * b.den; break;
        default:
          if (b.num == 0) throw Err("division by zero in a constant expression");
          n = (__int128)a.num * b.den; d = (__int128)a.den * b.num; break;
      }
      if (d < 0) { n = -n; d = -d; }
      __int128 x = n < 0 ? -n : n, y = d;
      while (y) { __int128 t = x % y; x = y; y = t; }
      if (x == 0) x = 1;
      n /= x; d /= x;
      const __int128 lim = (__int128)1 << 62;
      if (n > lim || n < -lim || d > lim) throw Err("rational constant overflow");
      return mk_rat((i64)n, (i64)d);
    }
    // (* 1 x) and (* x 1) with an exact 1 return x unchanged
    if (op == '*' && a.ty == T_RAT && a.num == 1 && a.den == 1) return b;
    if (op == '*' && b.ty == T_RAT && b.num == 1 && b.den == 1) return a;
    Ty t = wider(a.ty, b.ty);
    if (t == T_RAT) t = T_SGL;
    Ex r = mk(t, "(" + as_code(a, t) + " " + op + " " + as_code(b, t) + ")");
    affine_combine(r, op, a, b);
    return r;
  }

  // SBCL intexp: binary powering, base squared each round (code/numbers.lisp)
  Ex intexp(Ex base, i64 power) {
    if (power < 0) return arith('/', mk_rat(1, 1), intexp(base, -power));
    if (power > 4096) throw Err("integer exponent too large to unroll");
    base = materialise(base, "pb");
    i64 nextn = power >> 1;
    Ex total = (power & 1) ? base : mk_rat(1, 1);
    for (;;) {
      power = nextn;
      if (power == 0) return total;
      base = materialise(arith('*', base, base), "pb");
      nextn = power >> 1;
      if (power & 1) total = arith('*', base, total);
    }
  }

  // the expression mentions nothing but parameters (th<i>), literals and casts: it can be evaluated in safe()
  static bool only_parameters(const std::string& c) {
    for (size_t i = 0; i < c.size();) {
      const unsigned char ch = (unsigned char)c[i];
      if (isdigit(ch) || (ch == '.' && i + 1 < c.size() && isdigit((unsigned char)c[i + 1]))) {  // a numeric literal
        while (i < c.size() && (isalnum((unsigned char)c[i]) || c[i] == '.' ||
                                ((c[i] == '+' || c[i] == '-') && (c[i - 1] == 'e' || c[i - 1] == 'E')))) i++;
        continue;
      }
      if (isalpha(ch) || ch == '_') {
        size_t j = i;
        while (j < c.size() && (isalnum((unsigned char)c[j]) || c[j] == '_')) j++;
        const std::string id = c.substr(i, j - i);
        bool ok = id == "double" || id == "float";
        if (!ok && id.size() > 2 && id[0] == 't' && id[1] == 'h') {
          ok = true;
          for (size_t k = 2; k < id.size(); k++) ok = ok && isdigit((unsigned char)id[k]);
        }
        if (!ok) return false;
        i = j;
        continue;
      }
      i++;
    }
    return true;
  }

  Ex irrational(const char* fn0, const Ex& a) {
    std::string fn = fn0;
    if (a.ty == T_BOOL) throw Err(fn + " of a boolean");
    std::string tail = ")";
    if (fast_math) {
      const std::string plain = fn;
      if (fn == "exp") { fn = "mfit_exp_t<SLOW>"; uses |= 1; guarded = true; tail = ", gd)"; }
      else if (fn == "log") { fn = "mfit_log_t<SLOW>"; uses |= 2; guarded = true; tail = ", gd)"; }
      else if (fn == "sin") { fn = "mfit_sin_t<SLOW>"; guarded = true; tail = ", gd)"; }
      else if (fn == "cos") { fn = "mfit_cos_t<SLOW>"; guarded = true; tail = ", gd)"; }
      if (fn != plain) {   // a routine with a restricted fast domain: can safe() bound its argument over the x range?
        Check c;
        c.fn = plain;
        if (affine_parts(a, c.A, c.B) && only_parameters(c.A) && only_parameters(c.B)) checks.push_back(c);
        else provable = false;
      }
    }
    if (a.ty == T_DBL) return mk(T_DBL, fn + "(" + a.code + tail);
    // rational or single argument: computed in double, result is a single-float
    return mk(T_SGL, "((float)" + fn + "(" + as_code(a, T_DBL) + tail + ")");
  }

  // exact three-way knowledge about a rational constant c and its double rounding r: sign(r - c)
  static int rounding_side(const Ex& c) {
    double r = rat_value(c);
    if (std::isinf(r) || r == 0.0) return 0;
    int e;
    double m = frexp(r, &e);            // r = m 2^e, |m| in [0.5,1)
    i64 mi = (i64)ldexp(m, 53);          // 53-bit integer mantissa
    int sh = e - 53;                     // r = mi 2^sh
    // compare mi 2^sh * den  with  num
    __int128 lhs = (__int128)mi * c.den, rhs = c.num;
    if (sh >= 0) { if (sh > 60) return 0; lhs <<= sh; }
    else { if (-sh > 60) return 0; rhs <<= -sh; }
    return lhs < rhs ? -1 : lhs > rhs ? 1 : 0;
  }

  // one link of a comparison chain, CL semantics (exact across number formats)
  Ex compare2(const std::string& op, const Ex& a, const Ex& b) {
    if (a.ty == T_BOOL || b.ty == T_BOOL) throw Err("comparison of a boolean");
    if (a.ty == T_RAT && b.ty == T_RAT) {
      __int128 l = (__int128)a.num * b.den, r = (__int128)b.num * a.den;
      bool v = op == "<" ? l < r : op == ">" ? l > r : op == "<=" ? l <= r : op == ">=" ? l >= r : op == "==" ? l == r : l != r;
      return mk_bool(v);
    }
    // float formats convert to double exactly; a rational constant is compared through its
    // rounded double with the operator adjusted so that the outcome equals the exact one
    std::string o = op;
    if (a.ty == T_RAT || b.ty == T_RAT) {
      const bool const_right = b.ty == T_RAT;
      const Ex& c = const_right ? b : a;
      int side = rounding_side(c);  // sign(r - c)
      if (side != 0) {
        // x OP c  with r = rounded c.  For x a double: x < c <=> (r > c ? x < r : x <= r) etc.
        std::string oo = const_right ? o : (o == "<" ? ">" : o == ">" ? "<" : o == "<=" ? ">=" : o == ">=" ? "<=" : o);
        if (oo == "<") oo = side > 0 ? "<" : "<=";
        else if (oo == "<=") oo = side > 0 ? "<" : "<=";
        else if (oo == ">") oo = side > 0 ? ">=" : ">";
        else if (oo == ">=") oo = side > 0 ? ">=" : ">";
        else if (oo == "==") return mk_bool(false);  // c is not representable: never equal
        else if (oo == "!=") return mk_bool(true);
        const Ex& x = const_right ? a : b;
        return mk(T_BOOL, "(" + as_code(x, T_DBL) + " " + oo + " " + fmt_double(rat_value(c)) + ")");
      }
    }
    return mk(T_BOOL, "(" + as_code(a, T_DBL) + " " + o + " " + as_code(b, T_DBL) + ")");
  }

  Ex logical_and(const Ex& a, const Ex& b) {
    if (a.ty == T_BOOL && a.bconst) return a.bval ? b : a;
    if (b.ty == T_BOOL && b.bconst) return b.bval ? a : b;
    return mk(T_BOOL, "(" + a.code + " && " + b.code + ")");
  }
  Ex truth(const Ex& e) {  // generalized boolean: every number is true
    if (e.ty == T_BOOL) return e;
    return mk_bool(true);
  }

  // ---------------------------------------------------------------- forms
  std::vector<Ex> args(const Sx& s, size_t from = 1) {
    std::vector<Ex> v;
    for (size_t i = from; i < s.kids.size(); i++) v.push_back(lower(s.kids[i]));
    return v;
  }
  static bool is_declare(const Sx& s) {
    return s.t == Sx::LIST && !s.kids.empty() && s.kids[0].t == Sx::SYM && s.kids[0].sym == "DECLARE";
  }
  Ex body(const Sx& s, size_t from) {  // implicit progn
    Ex last = mk_bool(false);
    for (size_t i = from; i < s.kids.size(); i++) {
      if (is_declare(s.kids[i])) continue;
      last = lower(s.kids[i]);
    }
    return last;
  }

  Ex lower_if(const Sx& test, const Sx* then_form, size_t then_from, const Sx& then_holder, const Sx* else_form) {
    Ex c = truth(lower(test));
    auto run_then = [&]() { return then_form ? lower(*then_form) : body(then_holder, then_from); };
    auto run_else = [&]() { return else_form ? lower(*else_form) : mk_bool(false); };
    if (c.bconst) return c.bval ? run_then() : run_else();
    std::vector<std::string>* outer = block;
    std::vector<std::string> ba, bb;
    block = &ba;
    Ex a = run_then();
    block = &bb;
    Ex b = run_else();
    block = outer;
    Ty t;
    if (a.ty == T_BOOL || b.ty == T_BOOL) {
      if (a.ty != T_BOOL || b.ty != T_BOOL) {
        // (if c number nil): only meaningful as a generalized boolean
        a = truth(a);
        b = truth(b);
      }
      t = T_BOOL;
    } else {
      t = wider(a.ty, b.ty);
      if (t == T_RAT) {
        if (a.num == b.num && a.den == b.den) return a;
        // two different exact constants: the value is a float of the format the context demands;
        // CL would keep them exact, so represent as double (exact for every rational we can fold
        // that is a dyadic; otherwise the nearest double)
        t = T_DBL;
      }
    }
    const char* cty = t == T_DBL ? "double" : t == T_SGL ? "float" : "bool";
    std::string ca = t == T_BOOL ? a.code : as_code(a, t), cb = t == T_BOOL ? b.code : as_code(b, t);
    if (ba.empty() && bb.empty()) return mk(t, "(" + c.code + " ? " + ca + " : " + cb + ")");
    std::string r = fresh("r");
    block->push_back(std::string(cty) + " " + r + ";");
    std::string st = "if (" + c.code + ") {";
    for (auto& l : ba) st += " " + l;
    st += " " + r + " = " + ca + "; } else {";
    for (auto& l : bb) st += " " + l;
    st += " " + r + " = " + cb + "; }";
    block->push_back(st);
    return mk(t, r);
  }

  Ex lower_let(const Sx& s, bool star) {
    if (s.kids.size() < 2 || s.kids[1].t != Sx::LIST) throw Err("malformed let");
    std::vector<std::pair<std::string, Ex>> pend;
    std::vector<std::string> bound;
    for (auto& b : s.kids[1].kids) {
      std::string name;
      Ex init = mk_bool(false);
      if (b.t == Sx::SYM) name = b.sym;
      else {
        if (b.kids.empty() || b.kids[0].t != Sx::SYM) throw Err("malformed let binding");
        name = b.kids[0].sym;
        if (b.kids.size() > 1) init = materialise(lower(b.kids[1]), "v");
      }
      if (star) { bind(name, init); bound.push_back(name); }
      else pend.push_back({name, init});
    }
    for (auto& pb : pend) { bind(pb.first, pb.second); bound.push_back(pb.first); }
    Ex r = body(s, 2);
    for (auto& n : bound) unbind(n);
    return r;
  }

  // prior-bounds-let, mcmc-fitting.lisp:346-369
  Ex lower_prior_bounds_let(const Sx& s) {
    if (s.kids.size() < 2 || s.kids[1].t != Sx::LIST) throw Err("malformed prior-bounds-let");
    const Sx& khl = s.kids[1];
    std::vector<std::string> bound;
    auto name_of = [](const Sx& e) {
      if (e.t != Sx::LIST || e.kids.size() != 3 || e.kids[0].t != Sx::SYM) throw Err("prior-bounds-let wants (key low high)");
      std::string k = e.kids[0].sym;
      return k[0] == ':' ? k.substr(1) : k;
    };
    for (auto& e : khl.kids) {  // (param (the double-float (getf params key 0d0)))  :353
      std::string name = name_of(e);
      int i = key_index(name);
      bind(name, i >= 0 ? theta(i) : mk(T_DBL, "0.0"));
      bound.push_back(name);
    }
    std::vector<Ex> bs;
    for (auto& e : khl.kids) {  // :358-360
      std::string name = name_of(e);
      Ex p = scope[name].back();
      Ex lo = materialise(lower(e.kids[1]), "lo"), hi = materialise(lower(e.kids[2]), "hi");
      Ex test = logical_and(compare2("<", lo, p), compare2("<", p, hi));
      Ex dhi = mk(T_DBL, "fabs(" + as_code(arith('-', p, hi), T_DBL) + ")");
      Ex dlo = mk(T_DBL, "fabs(" + as_code(arith('-', p, lo), T_DBL) + ")");
      dhi = materialise(dhi, "dh");
      dlo = materialise(dlo, "dl");
      Ex m = mk(T_DBL, "((" + dlo.code + " < " + dhi.code + ") ? " + dlo.code + " : " + dhi.code + ")");  // (min dhi dlo)
      Ex pen = mk(T_DBL, "(-1e10 * (exp(" + m.code + " * 1e-5) - 1.0))");
      Ex b;
      if (test.bconst) b = test.bval ? mk(T_DBL, "0.0") : pen;
      else b = mk(T_DBL, "(" + test.code + " ? 0.0 : " + pen.code + ")");
      b = materialise(b, "bd");
      bind(name + "-BOUND", b);
      bound.push_back(name + "-BOUND");
      bs.push_back(b);
    }
    Ex total = bs.empty() ? mk_rat(0, 1) : bs[0];  // (+ b1 b2 ...)  :368
    for (size_t i = 1; i < bs.size(); i++) total = arith('+', total, bs[i]);
    total = materialise(total, "bt");
    bind("BOUNDS-TOTAL", total);
    bound.push_back("BOUNDS-TOTAL");
    Ex r = body(s, 2);
    for (auto& n : bound) unbind(n);
    return r;
  }

  Ex minmax(bool is_min, std::vector<Ex> a) {
    if (a.empty()) throw Err("min/max without arguments");
    Ex acc = a[0];
    for (size_t i = 1; i < a.size(); i++) {
      const Ex& b = a[i];
      if (acc.ty == T_RAT && b.ty == T_RAT) {
        __int128 l = (__int128)b.num * acc.den, r = (__int128)acc.num * b.den;
        if (is_min ? l < r : l > r) acc = b;
        continue;
      }
      Ty t = wider(acc.ty, b.ty);
      if (t == T_RAT) t = T_SGL;
      Ex x = materialise(mk(t, as_code(acc, t)), "m"), y = materialise(mk(t, as_code(b, t)), "m");
      // acc is replaced only when the new argument compares strictly below (above) it
      acc = mk(t, "((" + y.code + (is_min ? " < " : " > ") + x.code + ") ? " + y.code + " : " + x.code + ")");
    }
    return acc;
  }

  // the symbol names the model's positional argument and is not shadowed by a let binding
  bool is_x(const Sx& s) const {
    if (s.t != Sx::SYM || x_name.empty() || s.sym != x_name) return false;
    auto it = scope.find(s.sym);
    return it != scope.end() && it->second.size() == 1 && it->second.back().code == "x";
  }

  Ex lower(const Sx& s) {
    if (s.t == Sx::NUM) {
      if (s.nk == 0) return mk_rat(s.num, s.den);
      Ex lit = s.nk == 1 ? mk(T_SGL, fmt_float((float)s.fv)) : mk(T_DBL, fmt_double(s.fv));
      lit.xk = 0;
      lit.B = as_code(lit, T_DBL);
      return lit;
    }
    if (s.t == Sx::SYM) {
      if (s.sym == "PI") return mk(T_DBL, "3.141592653589793");
      if (s.sym == "T") return mk_bool(true);
      if (s.sym == "NIL") return mk_bool(false);
      auto it = scope.find(s.sym);
      if (it == scope.end() || it->second.empty()) throw Err("unbound variable " + s.sym);
      if (is_x(s)) {
        if (x_dims > 0) throw Err("x is used both as a number and as a sequence");
        x_scalar = true;
      }
      return it->second.back();
    }
    if (s.kids.empty()) return mk_bool(false);
    if (s.kids[0].t != Sx::SYM) throw Err("only symbols can be called");
    const std::string& f = s.kids[0].sym;
    const size_t na = s.kids.size() - 1;
    auto need = [&](size_t lo, size_t hi) {
      if (na < lo || na > hi) throw Err("wrong number of arguments to " + f);
    };
    if (f == "THE") { need(2, 2); return lower(s.kids[2]); }
    {
      // multi-dimensional inputs (walker-create docstring, mcmc-fitting.lisp:1137): x is itself a sequence and the model
      // reads its components -- each becomes one more shared-memory-staged data array (xv[k] in the lowered code)
      static const struct { const char* name; int fixed; } acc[] = {
          {"ELT", -1}, {"AREF", -1}, {"SVREF", -1}, {"NTH", -2}, {"FIRST", 0}, {"SECOND", 1}, {"THIRD", 2},
          {"FOURTH", 3}, {"CAR", 0}, {"CADR", 1}, {"CADDR", 2}};
      for (auto& a : acc)
        if (f == a.name) {
          const bool fixed = a.fixed >= 0;
          if (!is_ll && na >= 1 && (a.fixed != -2 || na >= 2)) {
            // element k of a sequence-valued parameter (:single-item style): (elt params k), (aref params k), (aref params k 0)
            std::string sk;
            if (seq_key_of(s.kids[a.fixed == -2 ? 2 : 1], sk)) {
              i64 idx = a.fixed;
              if (!fixed) {
                if (na < 2 || na > (f == "AREF" ? 3u : 2u)) throw Err("wrong number of arguments to " + f);
                const Sx& ix = s.kids[a.fixed == -2 ? 1 : 2];
                if (ix.t != Sx::NUM || ix.nk != 0 || ix.den != 1 || ix.num < 0) throw Err(f + ": the index must be a literal non-negative integer");
                idx = ix.num;
                if (na == 3) {  // an n x 1 array: the column index is 0
                  const Sx& jx = s.kids[3];
                  if (jx.t != Sx::NUM || jx.nk != 0 || jx.den != 1 || jx.num != 0) throw Err("aref: the parameter array is n x 1, its second index must be 0");
                }
              } else {
                need(1, 1);
              }
              const int ti = key_index(sk + "[" + std::to_string(idx) + "]");
              if (ti < 0) throw Err(f + ": index " + std::to_string(idx) + " is beyond the parameter sequence " + sk);
              return theta(ti);
            }
          }
          need(fixed ? 1 : 2, fixed ? 1 : 2);
          const Sx& seq = s.kids[a.fixed == -2 ? 2 : 1];
          if (is_ll) {
            // `error` is the data set's whole stddev LIST (:416), the lambda has no index of the point it is called for:
            // only a literal position can be read -- a per-data-set constant, spliced into the kernel as MFIT_ERR(k)
            if (seq.t != Sx::SYM || seq.sym != err_name) throw Err(f + ": only the error list can be indexed in a likelihood lambda");
            i64 idx = a.fixed;
            if (!fixed) {
              const Sx& ix = s.kids[a.fixed == -2 ? 1 : 2];
              if (ix.t != Sx::NUM || ix.nk != 0 || ix.den != 1 || ix.num < 0) throw Err(f + ": the index must be a literal non-negative integer");
              idx = ix.num;
            }
            if (idx >= 16) throw Err(f + ": only the first 16 elements of the error list can be read");
            if ((int)idx > err_max) err_max = (int)idx;
            return mk(T_DBL, "MFIT_ERR(" + std::to_string(idx) + ")");
          }
          if (!is_x(seq)) throw Err(f + ": only the model's positional argument can be indexed");
          i64 idx = a.fixed;
          if (!fixed) {
            const Sx& ix = s.kids[a.fixed == -2 ? 1 : 2];
            if (ix.t != Sx::NUM || ix.nk != 0 || ix.den != 1 || ix.num < 0) throw Err(f + ": the index must be a literal non-negative integer");
            idx = ix.num;
          }
          if (idx >= MFIT_MAX_XDIMS) throw Err(f + ": at most " + std::to_string(MFIT_MAX_XDIMS) + " components of x are supported");
          if (x_scalar) throw Err("x is used both as a number and as a sequence");
          if ((int)idx + 1 > x_dims) x_dims = (int)idx + 1;
          return mk(T_DBL, "xv[" + std::to_string(idx) + "]");
        }
    }
    if (f == "LOG-NORMAL") {  // :372-377
      need(3, 3);
      Ex a = lower(s.kids[1]), b = lower(s.kids[2]), c = lower(s.kids[3]);
      if (a.ty == T_BOOL || b.ty == T_BOOL || c.ty == T_BOOL) throw Err("log-normal of a non-number (the error LIST itself is not a number: use (elt error k))");
      return mk(T_DBL, "mfit_u_log_normal(" + as_code(a, T_DBL) + ", " + as_code(b, T_DBL) + ", " + as_code(c, T_DBL) + ")");
    }
    if (f == "LOG-FACTORIAL" || f == "LOG-POISSON") {  // :379-383; the factorial of the DATUM only (a table built at upload)
      const bool pois = f == "LOG-POISSON";
      need(pois ? 2 : 1, pois ? 2 : 1);
      const Sx& k = s.kids[pois ? 2 : 1];
      if (!is_ll || k.t != Sx::SYM || k.sym != y_name)
        throw Err(f + ": the count must be the likelihood lambda's y (log-factorial of anything else is not supported)");
      uses_logfact = true;
      if (!pois) return mk(T_SGL, "((float)lf)");   // (log-factorial n) is a single-float
      Ex lam = lower(s.kids[1]);
      if (lam.ty == T_BOOL) throw Err("log-poisson of a non-number");
      return mk(T_DBL, "mfit_u_log_poisson(" + as_code(lam, T_DBL) + ", y, lf)");
    }
    if (f == "DECLARE") return mk_bool(false);
    if (f == "PROGN") return body(s, 1);
    if (f == "LET") return lower_let(s, false);
    if (f == "LET*") return lower_let(s, true);
    if (f == "PRIOR-BOUNDS-LET") return lower_prior_bounds_let(s);
    if (f == "GETF") {
      need(2, 3);
      if (s.kids[2].t != Sx::SYM) throw Err("getf wants a literal key");
      int i = key_index(s.kids[2].sym);
      if (i >= 0) return theta(i);
      if (na == 3) return lower(s.kids[3]);
      throw Err("getf of a key that is not a parameter, without default: " + s.kids[2].sym);
    }
    if (f == "IF") { need(2, 3); return lower_if(s.kids[1], &s.kids[2], 0, s, na == 3 ? &s.kids[3] : nullptr); }
    if (f == "WHEN") { need(1, 1000); return lower_if(s.kids[1], nullptr, 2, s, nullptr); }
    if (f == "UNLESS") {
      need(1, 1000);
      Sx neg; neg.kids.resize(2); neg.kids[0].t = Sx::SYM; neg.kids[0].sym = "NOT"; neg.kids[1] = s.kids[1];
      return lower_if(neg, nullptr, 2, s, nullptr);
    }
    if (f == "COND") {
      // (cond (t1 b1...) (t2 b2...) ...) == nested if
      Sx chain; bool have = false;
      for (size_t i = s.kids.size() - 1; i >= 1; i--) {
        const Sx& cl = s.kids[i];
        if (cl.t != Sx::LIST || cl.kids.empty()) throw Err("malformed cond clause");
        Sx node; node.kids.resize(have ? 4 : 3);
        node.kids[0].t = Sx::SYM; node.kids[0].sym = "IF";
        node.kids[1] = cl.kids[0];
        Sx pr; pr.kids.resize(1); pr.kids[0].t = Sx::SYM; pr.kids[0].sym = "PROGN";
        for (size_t j = 1; j < cl.kids.size(); j++) pr.kids.push_back(cl.kids[j]);
        node.kids[2] = pr;
        if (have) node.kids[3] = chain;
        chain = node; have = true;
      }
      if (!have) return mk_bool(false);
      return lower(chain);
    }
    if (f == "AND" || f == "OR") {
      std::vector<Ex> a = args(s);
      if (a.empty()) return mk_bool(f == "AND");
      Ex acc = truth(a[0]);
      for (size_t i = 1; i < a.size(); i++) {
        Ex b = truth(a[i]);
        if (f == "AND") acc = logical_and(acc, b);
        else {
          if (acc.bconst) acc = acc.bval ? acc : b;
          else if (b.bconst) acc = b.bval ? b : acc;
          else acc = mk(T_BOOL, "(" + acc.code + " || " + b.code + ")");
        }
      }
      return acc;
    }
    if (f == "NOT" || f == "NULL") {
      need(1, 1);
      Ex a = truth(lower(s.kids[1]));
      if (a.bconst) return mk_bool(!a.bval);
      return mk(T_BOOL, "(!" + a.code + ")");
    }
    if (f == "+" || f == "*") {
      std::vector<Ex> a = args(s);
      if (a.empty()) return mk_rat(f == "+" ? 0 : 1, 1);
      Ex acc = a[0];
      for (size_t i = 1; i < a.size(); i++) acc = arith(f[0], acc, a[i]);
      return acc;
    }
    if (f == "-" || f == "/") {
      std::vector<Ex> a = args(s);
      need(1, 1000);
      if (a.size() == 1) {
        if (f == "-") {
          if (a[0].ty == T_RAT) return mk_rat(-a[0].num, a[0].den);
          if (a[0].ty == T_BOOL) throw Err("negation of a boolean");
          Ex neg = mk(a[0].ty, "(-" + a[0].code + ")");
          affine_combine(neg, '-', mk_rat(0, 1), a[0]);
          return neg;
        }
        return arith('/', mk_rat(1, 1), a[0]);
      }
      Ex acc = a[0];
      for (size_t i = 1; i < a.size(); i++) acc = arith(f[0], acc, a[i]);
      return acc;
    }
    if (f == "1+") { need(1, 1); return arith('+', lower(s.kids[1]), mk_rat(1, 1)); }
    if (f == "1-") { need(1, 1); return arith('-', lower(s.kids[1]), mk_rat(1, 1)); }
    static const struct { const char* lisp; const char* c; } unary[] = {
        {"EXP", "exp"}, {"SIN", "sin"}, {"COS", "cos"}, {"TAN", "tan"}, {"ASIN", "asin"}, {"ACOS", "acos"},
        {"SINH", "sinh"}, {"COSH", "cosh"}, {"TANH", "tanh"}, {"SQRT", "sqrt"}};
    for (auto& u : unary)
      if (f == u.lisp) { need(1, 1); return irrational(u.c, lower(s.kids[1])); }
    if (f == "LOG") {
      need(1, 2);
      Ex a = irrational("log", lower(s.kids[1]));
      if (na == 2) return arith('/', a, irrational("log", lower(s.kids[2])));  // (log x base)
      return a;
    }
    if (f == "ATAN") {
      need(1, 2);
      Ex a = lower(s.kids[1]);
      if (na == 1) return irrational("atan", a);
      Ex b = lower(s.kids[2]);
      std::string c = "atan2(" + as_code(a, T_DBL) + ", " + as_code(b, T_DBL) + ")";
      if (a.ty == T_DBL || b.ty == T_DBL) return mk(T_DBL, c);
      return mk(T_SGL, "((float)" + c + ")");
    }
    if (f == "EXPT") {
      need(2, 2);
      Ex b = lower(s.kids[1]), p = lower(s.kids[2]);
      if (p.ty == T_RAT && p.den == 1) return intexp(b, p.num);
      if (b.ty == T_BOOL || p.ty == T_BOOL) throw Err("expt of a boolean");
      // a float exponent that is a small whole number literal: multiplications instead of pow
      if (s.kids[2].t == Sx::NUM && s.kids[2].nk != 0) {
        double pv = s.kids[2].fv;
        if (pv == std::floor(pv) && std::fabs(pv) <= 8 && pv != 0) {
          Ty t = wider(b.ty, p.ty);
          Ex bb = mk(t, as_code(b, t));
          return intexp(bb, (i64)pv);
        }
      }
      std::string c = "pow(" + as_code(b, T_DBL) + ", " + as_code(p, T_DBL) + ")";
      if (b.ty == T_DBL || p.ty == T_DBL) return mk(T_DBL, c);
      return mk(T_SGL, "((float)" + c + ")");
    }
    if (f == "ABS") {
      need(1, 1);
      Ex a = lower(s.kids[1]);
      if (a.ty == T_RAT) return mk_rat(a.num < 0 ? -a.num : a.num, a.den);
      if (a.ty == T_BOOL) throw Err("abs of a boolean");
      return mk(a.ty, std::string(a.ty == T_DBL ? "fabs(" : "fabsf(") + a.code + ")");
    }
    if (f == "MIN") return minmax(true, args(s));
    if (f == "MAX") return minmax(false, args(s));
    if (f == "SIGNUM") {
      need(1, 1);
      Ex a = lower(s.kids[1]);
      if (a.ty == T_RAT) return mk_rat(a.num > 0 ? 1 : a.num < 0 ? -1 : 0, 1);
      if (a.ty == T_BOOL) throw Err("signum of a boolean");
      Ex v = materialise(a, "sg");
      const char* one = a.ty == T_DBL ? "1.0" : "1.0f";
      return mk(a.ty, "((" + v.code + " > 0) ? " + one + " : ((" + v.code + " < 0) ? -" + one + " : " + v.code + "))");
    }
    if (f == "FLOAT") {
      need(1, 2);
      Ex a = lower(s.kids[1]);
      if (na == 2) {
        Ex proto = lower(s.kids[2]);
        return mk(proto.ty == T_DBL ? T_DBL : T_SGL, as_code(a, proto.ty == T_DBL ? T_DBL : T_SGL));
      }
      if (a.ty == T_RAT) return mk(T_SGL, as_code(a, T_SGL));
      return a;
    }
    if (f == "COERCE") {
      need(2, 2);
      Ex a = lower(s.kids[1]);
      const Sx& ty = s.kids[2];
      std::string tn = ty.t == Sx::SYM ? ty.sym : "";
      if (tn == "DOUBLE-FLOAT" || tn == "LONG-FLOAT") return mk(T_DBL, as_code(a, T_DBL));
      if (tn == "SINGLE-FLOAT" || tn == "FLOAT" || tn == "SHORT-FLOAT") return a.ty == T_DBL && tn == "FLOAT" ? a : mk(T_SGL, as_code(a, T_SGL));
      throw Err("coerce to an unsupported type");
    }
    static const struct { const char* lisp; const char* c; } rel[] = {
        {"<", "<"}, {">", ">"}, {"<=", "<="}, {">=", ">="}, {"=", "=="}};
    for (auto& r : rel)
      if (f == r.lisp) {
        need(1, 1000);
        std::vector<Ex> a = args(s);
        for (auto& e : a) e = materialise(e, "c");
        Ex acc = mk_bool(true);
        for (size_t i = 0; i + 1 < a.size(); i++) acc = logical_and(acc, compare2(r.c, a[i], a[i + 1]));
        return acc;
      }
    if (f == "/=") {
      std::vector<Ex> a = args(s);
      for (auto& e : a) e = materialise(e, "c");
      Ex acc = mk_bool(true);
      for (size_t i = 0; i < a.size(); i++)
        for (size_t j = i + 1; j < a.size(); j++) acc = logical_and(acc, compare2("!=", a[i], a[j]));
      return acc;
    }
    if (f == "SQUARE") { need(1, 1); Ex a = materialise(lower(s.kids[1]), "sq"); return arith('*', a, a); }
    throw Err("unsupported operator " + f);
  }
};

struct Parsed {
  bool is_prior = false;
  std::string x_name;
  const Sx* lambda_list = nullptr;
  size_t body_from = 0;
};

std::string lower_text(const char* text, const char* const* key_names, int n_keys, int kind, int* x_dims_out = nullptr) {
  Reader rd{text};
  Sx top = rd.read();
  if (top.t != Sx::LIST || top.kids.size() < 3 || top.kids[0].t != Sx::SYM) throw Err("expected (lambda (args) body)");
  size_t li = 1;
  if (top.kids[0].sym == "DEFUN") li = 2;
  else if (top.kids[0].sym != "LAMBDA") throw Err("expected lambda or defun");
  if (top.kids.size() <= li + 1) throw Err("lambda without a body");
  const Sx& ll = top.kids[li];
  if (ll.t != Sx::LIST) throw Err("malformed lambda list");

  Lower L;
  for (int i = 0; i < n_keys; i++) {
    std::string k = key_names[i];
    for (auto& c : k) c = (char)toupper((unsigned char)c);
    if (!k.empty() && k[0] == ':') k = k.substr(1);
    L.keys.push_back(k);
  }
  std::vector<std::string> pre;  // statements
  L.block = &pre;
  L.fast_math = kind == 0;

  std::vector<std::string> positional;
  bool in_key = false;
  for (auto& a : ll.kids) {
    if (a.t == Sx::SYM && a.sym == "&KEY") { in_key = true; continue; }
    if (a.t == Sx::SYM && (a.sym == "&ALLOW-OTHER-KEYS")) continue;
    if (a.t == Sx::SYM && a.sym[0] == '&') throw Err("unsupported lambda-list keyword " + a.sym);
    if (!in_key) {
      if (a.t != Sx::SYM) throw Err("malformed positional argument");
      positional.push_back(a.sym);
      continue;
    }
    std::string name = a.t == Sx::SYM ? a.sym : (a.kids.empty() ? std::string() : a.kids[0].sym);
    if (name.empty()) throw Err("malformed &key argument");
    int i = L.key_index(name);
    if (i >= 0) L.bind(name, L.theta(i));
    else if (L.seq_key_count(name) > 0) {   // :single-item style: the key's value is a sequence, read with elt / aref ...
      L.seq_vars.insert(name);
      L.bind(name, mk_bool(true));          // a sequence is no number: using it as one is a type error in the reference too
    } else if (a.t == Sx::LIST && a.kids.size() > 1) L.bind(name, L.lower(a.kids[1]));  // default value form
    else L.bind(name, mk_bool(false));                                                 // nil
  }
  std::string out;
  if (kind == 0) {
    if (positional.size() != 1) throw Err("a model lambda takes exactly one positional argument (x)");
    L.x_name = positional[0];
    Ex xe = mk(T_DBL, "x");
    xe.xk = 1;
    xe.A = "1.0";
    L.bind(positional[0], xe);
  } else if (kind == 2) {
    if (positional.size() != 3 || in_key) throw Err("a likelihood lambda takes (y model error)");
    L.is_ll = true;
    L.y_name = positional[0];
    L.err_name = positional[2];
    L.bind(positional[0], mk(T_DBL, "y"));
    L.bind(positional[1], mk(T_DBL, "f"));
    L.bind(positional[2], mk_bool(true));   // a list is no number: using it as one is a type error in the reference too
  } else {
    if (positional.size() != 2) throw Err("a prior lambda takes (params data)");
  }
  Ex r = mk_bool(false);
  bool any = false;
  for (size_t i = li + 1; i < top.kids.size(); i++) {
    const Sx& b = top.kids[i];
    if (Lower::is_declare(b)) continue;
    if (b.t == Sx::SYM && b.sym == "\"" && i + 1 < top.kids.size()) continue;  // docstring
    r = L.lower(b);
    any = true;
  }
  if (!any || r.ty == T_BOOL) throw Err("the function does not return a number");
  std::string ret = as_code(r, T_DBL);

  std::string loads, decls;
  for (int u : L.used) {
    decls += "  double th" + std::to_string(u) + ";\n";
    loads += "    th" + std::to_string(u) + " = theta[(size_t)" + std::to_string(u) + " * stride + chain];\n";
  }
  std::string stmts;
  for (auto& sline : pre) stmts += "    " + sline + "\n";
  if (x_dims_out) *x_dims_out = kind == 0 ? L.x_dims : kind == 2 ? L.err_max + 1 : 0;
  if (kind == 2) {
    // the per-point term of create-log-liklihood-function (:410-417): acc + g(y, model, error), summed by the kernel in
    // data order.  NARR = 4 when the log-factorial table of y is read (it travels in the slot of the cutoff constants).
    out += "template <>\nstruct MfitTerm<MFITK_USER> {\n";
    out += std::string("  static constexpr int NARR = ") + (L.uses_logfact ? "4" : "2") + ";\n  static constexpr int USES = 0;\n";
    out += "  template <bool SLOW, class G>\n  __device__ __forceinline__ static double add(double acc, double f, double y, double, double lf, G&) {\n";
    out += "    (void)f; (void)y; (void)lf;\n" + stmts + "    return acc + (" + ret + ");\n  }\n};\n";
    return out;
  }
  if (kind == 0) {
    // a user model carries no proof about its arguments (safe() is false): its fast routines always run guarded
    out += "struct MfitUser {\n  static constexpr int USES = " + std::to_string(L.uses) + ";\n";
    out += std::string("  static constexpr bool GUARDED = ") + (L.guarded ? "true" : "false") + ";\n";
    out += "  static constexpr bool CLAMP_ONLY = false;\n  static constexpr bool FUSABLE = false;\n";
    out += "  __device__ __forceinline__ bool positive(const MfitEvalArgs&) const { return false; }\n";
    out += "  static constexpr int XDIMS = " + std::to_string(L.x_dims > 0 ? L.x_dims : 1) + ";\n";
    out += std::string("  static constexpr bool XVEC = ") + (L.x_dims > 0 ? "true" : "false") + ";\n" + decls;
    out += "  __device__ __forceinline__ void load(const MfitEvalArgs& a, long long chain) {\n";
    out += "    const double* theta = a.theta; const long long stride = a.stride;\n    (void)theta; (void)stride; (void)chain;\n" + loads + "  }\n";
    out += "  __device__ __forceinline__ void setup() {}\n";
    // safe(): every argument of a fast routine is affine in x with parameter-only coefficients, hence monotone over the
    // data set's x range: bounding it at both ends bounds it everywhere (NaN anywhere makes a comparison false)
    if (L.provable && !L.checks.empty() && L.x_dims == 0) {
      std::string cond;
      int ci = 0;
      std::string pre2;
      for (auto& c : L.checks) {
        const std::string A = c.A.empty() ? "0.0" : c.A, B = c.B.empty() ? "0.0" : c.B;
        const std::string v0 = "s" + std::to_string(ci) + "a", v1 = "s" + std::to_string(ci) + "b";
        // with no dependence on x the range of x must not enter (0 * inf)
        pre2 += "    const double " + v0 + " = " + (c.A.empty() ? B : "fma(" + A + ", a.xmin, " + B + ")") + ", " + v1 + " = " +
                (c.A.empty() ? B : "fma(" + A + ", a.xmax, " + B + ")") + ";\n";
        std::string t;
        // plain comparisons, not fmax / fmin: those drop a NaN operand
        auto both = [&](const std::string& pre_, const std::string& post_) {
          return pre_ + v0 + post_ + " && " + pre_ + v1 + post_;
        };
        if (c.fn == "exp") t = both("fabs(", ") < 700.0");
        else if (c.fn == "log") t = both("", " > 1e-290") + " && " + both("", " < 1e290");
        else t = both("fabs(", ") < 9.0e4");
        cond += (ci ? " && " : "") + t;
        ci++;
      }
      out += "  __device__ __forceinline__ bool safe(const MfitEvalArgs& a) const {\n" + pre2 + "    return " + cond + ";\n  }\n";
    } else {
      out += "  __device__ __forceinline__ bool safe(const MfitEvalArgs&) const { return false; }\n";
    }
    if (L.x_dims > 0)
      out += "  template <bool SLOW, class G>\n  __device__ __forceinline__ double evalv(const double (&xv)[XDIMS], G& gd) const {\n    (void)xv; (void)gd;\n" + stmts + "    return " + ret + ";\n  }\n};\n";
    else
      out += "  template <bool SLOW, class G>\n  __device__ __forceinline__ double eval(double x, G& gd) const {\n    (void)x; (void)gd;\n" + stmts + "    return " + ret + ";\n  }\n};\n";
  } else {
    out += "__device__ __forceinline__ double mfit_user_prior(const double* theta, long long stride, long long chain) {\n";
    out += "  (void)theta; (void)stride; (void)chain;\n";
    std::string d2;
    for (int u : L.used)
      d2 += "  const double th" + std::to_string(u) + " = theta[(size_t)" + std::to_string(u) + " * stride + chain];\n";
    out += d2;
    for (auto& sline : pre) out += "  " + sline + "\n";
    out += "  return " + ret + ";\n}\n";
  }
  return out;
}

}  // namespace

extern "C" int mfit_lower_sexpr(const char* lambda_text, const char* const* key_names, int n_keys, int kind, char* out,
                                int out_len, char* err, int err_len) {
  if (!lambda_text || n_keys < 0 || (n_keys > 0 && !key_names) || kind < 0 || kind > 2) {
    if (err && err_len > 0) snprintf(err, err_len, "bad argument");
    return MFIT_ERR_ARG;
  }
  try {
    std::string s = lower_text(lambda_text, key_names, n_keys, kind);
    if (out && out_len > 0) {
      snprintf(out, out_len, "%s", s.c_str());
    }
    return (int)s.size() + 1;
  } catch (std::exception& e) {
    if (err && err_len > 0) snprintf(err, err_len, "%s", e.what());
    return MFIT_ERR_SEXPR;
  }
}

// (lambda (y model error) (log-poisson model y)) -- the way the reference spells a Poisson likelihood (:402-417 with
// :382-383) -- is the built-in kind: same kernel, same hoisted constant, bit for bit.
int mfit_ll_canonical_kind(const char* text) {
  try {
    Reader rd{text};
    Sx top = rd.read();
    if (top.t != Sx::LIST || top.kids.size() < 3 || top.kids[0].t != Sx::SYM || top.kids[0].sym != "LAMBDA") return -1;
    const Sx& ll = top.kids[1];
    if (ll.t != Sx::LIST || ll.kids.size() != 3) return -1;
    for (auto& a : ll.kids) if (a.t != Sx::SYM || a.sym[0] == '&') return -1;
    const Sx* body = nullptr;
    for (size_t i = 2; i < top.kids.size(); i++) {
      if (Lower::is_declare(top.kids[i])) continue;
      if (body) return -1;
      body = &top.kids[i];
    }
    if (!body || body->t != Sx::LIST || body->kids.size() != 3 || body->kids[0].t != Sx::SYM) return -1;
    if (body->kids[0].sym == "LOG-POISSON" && body->kids[1].t == Sx::SYM && body->kids[1].sym == ll.kids[1].sym &&
        body->kids[2].t == Sx::SYM && body->kids[2].sym == ll.kids[0].sym)
      return MFIT_LL_POISSON;
  } catch (...) {
  }
  return -1;
}

// C++ face used by api.cu
int mfit_lower_to_string(const char* text, const char* const* keys, int n_keys, int kind, std::string& out,
                         std::string& err, int* x_dims) {
  try {
    out = lower_text(text, keys, n_keys, kind, x_dims);
    return 0;
  } catch (std::exception& e) {
    err = e.what();
    return MFIT_ERR_SEXPR;
  }
}
