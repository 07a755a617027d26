// orc_sexpr.cpp -- ORACLE (test infrastructure only; see mfit_oracle.h).
//
// A small Common-Lisp-flavoured evaluator for the user plugin slots of the
// reference: the :function lambda  (lambda (x &key a b &allow-other-keys) body)
// called as (apply fn x params) at mcmc-fitting.lisp:400, and the :log-prior
// lambda (lambda (params data) ...) called at :1069, including the
// prior-bounds-let macro (:346-369).  It follows CL numeric contagion
// (integer/ratio -> single-float -> double-float) dynamically, the way SBCL
// evaluates the same forms: n-ary + - * / fold left to right, (expt x <int>)
// is repeated squaring (SBCL intexp), (expt x 2d0) is libm pow, irrational
// functions of rationals/singles return singles.
#include "orc_internal.h"
#include <cctype>
#include <cmath>
#include <limits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <set>
#include <stdexcept>

extern "C" double orc_log_normal(double x, double mu, double sigma);
extern "C" double orc_log_factorial(double n);
extern "C" double orc_log_poisson(double lambda, double k);

namespace orc {

// ---------------------------------------------------------------- reader
struct Sx {  // s-expression cell
  enum T { LIST, SYM, NUM } t = LIST;
  std::string sym;  // upper-cased symbol (keywords keep the leading ':')
  Val num;
  std::vector<Sx> kids;
};

static bool parse_number(const std::string& tok, Val& out) {
  // integer | ratio | float with optional exponent marker e s f d l
  size_t i = 0, n = tok.size();
  if (n == 0) return false;
  if (tok[i] == '+' || tok[i] == '-') i++;
  if (i >= n) return false;
  size_t digits = 0, j = i;
  while (j < n && isdigit((unsigned char)tok[j])) { j++; digits++; }
  if (j == n && digits) { out = Val::rat(strtoll(tok.c_str(), nullptr, 10), 1); return true; }
  if (j < n && tok[j] == '/' && digits) {
    size_t k = j + 1, d2 = 0;
    while (k < n && isdigit((unsigned char)tok[k])) { k++; d2++; }
    if (k != n || !d2) return false;
    long long nu = strtoll(tok.substr(0, j).c_str(), nullptr, 10);
    long long de = strtoll(tok.substr(j + 1).c_str(), nullptr, 10);
    out = Val::rat(nu, de);
    return true;
  }
  // float
  bool has_point = false;
  size_t fd = 0;
  if (j < n && tok[j] == '.') {
    has_point = true;
    j++;
    while (j < n && isdigit((unsigned char)tok[j])) { j++; fd++; }
  }
  if (digits + fd == 0) return false;
  char marker = 0;
  if (j < n) {
    char c = (char)tolower((unsigned char)tok[j]);
    if (c == 'e' || c == 's' || c == 'f' || c == 'd' || c == 'l') {
      marker = c;
      size_t k = j + 1;
      if (k < n && (tok[k] == '+' || tok[k] == '-')) k++;
      size_t ed = 0;
      while (k < n && isdigit((unsigned char)tok[k])) { k++; ed++; }
      if (k != n || !ed) return false;
    } else
      return false;
  }
  if (!marker && !has_point) return false;
  if (!marker && fd == 0) {  // "12." reads as an integer in CL
    out = Val::rat(strtoll(tok.c_str(), nullptr, 10), 1);
    return true;
  }
  std::string s = tok;
  if (marker) s[j] = 'e';
  if (marker == 'd' || marker == 'l')
    out = Val::dbl(strtod(s.c_str(), nullptr));
  else
    out = Val::sgl(strtof(s.c_str(), nullptr));  // *read-default-float-format* = single-float
  return true;
}

struct Reader {
  const char* p;
  explicit Reader(const char* s) : p(s) {}
  void skip() {
    for (;;) {
      while (*p && isspace((unsigned char)*p)) p++;
      if (*p == ';') { while (*p && *p != '\n') p++; continue; }
      if (p[0] == '#' && p[1] == '|') {
        p += 2;
        while (*p && !(p[0] == '|' && p[1] == '#')) p++;
        if (*p) p += 2;
        continue;
      }
      break;
    }
  }
  Sx read() {
    skip();
    if (!*p) throw std::runtime_error("unexpected end of s-expression");
    if (*p == '(') {
      p++;
      Sx l;
      l.t = Sx::LIST;
      for (;;) {
        skip();
        if (!*p) throw std::runtime_error("missing ')'");
        if (*p == ')') { p++; break; }
        l.kids.push_back(read());
      }
      return l;
    }
    if (*p == ')') throw std::runtime_error("unexpected ')'");
    if (*p == '"') {  // string (docstring): kept as an inert symbol
      p++;
      while (*p && *p != '"') { if (*p == '\\' && p[1]) p++; p++; }
      if (*p) p++;
      Sx a;
      a.t = Sx::SYM;
      a.sym = "\"STR";
      return a;
    }
    if (*p == '\'') { p++; return read(); }  // quote: data only appears as literal lists we never evaluate
    if (p[0] == '#' && p[1] == '\'') { p += 2; return read(); }
    std::string tok;
    while (*p && !isspace((unsigned char)*p) && *p != '(' && *p != ')' && *p != ';') tok.push_back(*p++);
    Sx a;
    Val v;
    if (parse_number(tok, v)) { a.t = Sx::NUM; a.num = v; return a; }
    a.t = Sx::SYM;
    for (auto& c : tok) c = (char)toupper((unsigned char)c);
    a.sym = tok;
    return a;
  }
};

// ---------------------------------------------------------------- numerics
double Val::as_double() const {
  switch (k) {
    case RAT: return (double)num / (double)den;
    case SGL: return (double)f;
    case DBL: return d;
    default: return b ? 1.0 : 0.0;
  }
}
static float as_single(const Val& v) {
  if (v.k == Val::RAT) return (float)((double)v.num / (double)v.den);
  if (v.k == Val::SGL) return v.f;
  return (float)v.d;
}
static long long gcdll(long long a, long long b) {
  if (a < 0) a = -a;
  if (b < 0) b = -b;
  while (b) { long long t = a % b; a = b; b = t; }
  return a ? a : 1;
}
Val Val::rat(long long n, long long d) {
  Val v;
  v.k = RAT;
  if (d == 0) throw std::runtime_error("division by zero (rational)");
  if (d < 0) { n = -n; d = -d; }
  long long g = gcdll(n, d);
  v.num = n / g;
  v.den = d / g;
  return v;
}
Val Val::sgl(float x) { Val v; v.k = SGL; v.f = x; return v; }
Val Val::dbl(double x) { Val v; v.k = DBL; v.d = x; return v; }
Val Val::boolean(bool x) { Val v; v.k = BOOL; v.b = x; return v; }

static int rank_of(const Val& a) { return a.k == Val::RAT ? 0 : a.k == Val::SGL ? 1 : 2; }

static Val arith(char op, const Val& a, const Val& b) {
  if (a.k == Val::BOOL || b.k == Val::BOOL) throw std::runtime_error("arithmetic on a boolean");
  int r = rank_of(a) > rank_of(b) ? rank_of(a) : rank_of(b);
  if (r == 0) {
    long long n, d;
    switch (op) {
      case '+': n = a.num * b.den + b.num * a.den; d = a.den * b.den; break;
      case '-': n = a.num * b.den - b.num * a.den; d = a.den * b.den; break;
      case '*': n = a.num * b.num; d = a.den * b.den; break;
      default: n = a.num * b.den; d = a.den * b.num; break;
    }
    return Val::rat(n, d);
  }
  if (r == 1) {
    float x = as_single(a), y = as_single(b);
    switch (op) {
      case '+': return Val::sgl(x + y);
      case '-': return Val::sgl(x - y);
      case '*': return Val::sgl(x * y);
      default: return Val::sgl(x / y);
    }
  }
  double x = a.as_double(), y = b.as_double();
  switch (op) {
    case '+': return Val::dbl(x + y);
    case '-': return Val::dbl(x - y);
    case '*': return Val::dbl(x * y);
    default: return Val::dbl(x / y);
  }
}

// exact comparison across the numeric tower (floats convert to double exactly)
static int cmp(const Val& a, const Val& b) {
  if (a.k == Val::RAT && b.k == Val::RAT) {
    __int128 l = (__int128)a.num * b.den, r = (__int128)b.num * a.den;
    return l < r ? -1 : l > r ? 1 : 0;
  }
  // rational vs float: CL compares exactly; long double keeps small ratios exact enough
  long double x = a.k == Val::RAT ? (long double)a.num / (long double)a.den : (long double)a.as_double();
  long double y = b.k == Val::RAT ? (long double)b.num / (long double)b.den : (long double)b.as_double();
  if (x != x || y != y) return 2;  // unordered
  return x < y ? -1 : x > y ? 1 : 0;
}

// SBCL intexp (code/numbers.lisp): binary powering, base squared each round
static Val intexp(Val base, long long power) {
  if (power < 0) return arith('/', Val::rat(1, 1), intexp(base, -power));
  long long nextn = power >> 1;
  Val total = (power & 1) ? base : Val::rat(1, 1);
  for (;;) {
    power = nextn;
    if (power == 0) return total;
    base = arith('*', base, base);
    nextn = power >> 1;
    if (power & 1) total = arith('*', base, total);
  }
}

typedef double (*fn1)(double);
static Val irrational(fn1 f, const Val& a) {
  if (a.k == Val::DBL) return Val::dbl(f(a.d));
  // (exp 1) and (exp 1.0) are single-floats: computed in double, coerced (SBCL irrat.lisp)
  return Val::sgl((float)f(a.as_double()));
}

// ---------------------------------------------------------------- compiled tree
enum Op {
  O_CONST, O_VAR, O_ADD, O_SUB, O_MUL, O_DIV, O_INC, O_DEC, O_EXP, O_LOG, O_SIN, O_COS, O_TAN, O_ASIN,
  O_ACOS, O_ATAN, O_SINH, O_COSH, O_TANH, O_SQRT, O_EXPT, O_ABS, O_MIN, O_MAX, O_SIGNUM, O_IF, O_AND,
  O_OR, O_NOT, O_LT, O_GT, O_LE, O_GE, O_EQ, O_NE, O_LET, O_PROGN, O_FLOAT, O_XELT,
  O_LOGNORMAL, O_LOGPOISSON, O_LOGFACT   // the reference's own per-point densities, callable from a likelihood lambda (:372-383)
};

struct Node {
  Op op = O_CONST;
  Val c;
  int slot = -1;               // O_VAR: env slot; O_LET: first slot bound
  std::vector<Node> kids;      // O_LET: kids[0..nb-1] = init forms, kids[nb..] = body
  int nb = 0;
  std::vector<int> slots;      // O_LET: env slot of each binding
};

struct Fn {
  Node body;
  int nslots = 0;
  int P = 0;
  bool is_prior = false;
  bool is_ll = false;              // (lambda (y model error) ...) of create-log-liklihood-function :402-417
  int y_slot = -1, model_slot = -1; // its first two arguments; the third (the whole stddev LIST, :416) is the sequence x_slot
  int x_slot = -1;
  std::vector<int> theta_slot;     // per theta index -> slot (or -1)
  std::vector<std::pair<int, Val>> defaults;  // slot, default for keys not in the plist
};

struct Scope {
  std::vector<std::pair<std::string, int>> names;
  int find(const std::string& s) const {
    for (int i = (int)names.size() - 1; i >= 0; i--)
      if (names[i].first == s) return names[i].second;
    return -1;
  }
};

struct Compiler {
  Fn* fn;
  Scope sc;
  std::vector<std::string> keys;  // upper-case, without ':'
  std::set<std::string> seq_vars;  // &key variables whose value is a sequence (:single-item style; keys NAME[0] .. NAME[n-1])
  bool has_seq_key(const std::string& name) const {
    for (auto& k : keys)
      if (k == name + "[0]") return true;
    return false;
  }
  bool seq_key_of(const Sx& seq, std::string& name) const {
    if (seq.t == Sx::SYM) {
      if (!seq_vars.count(seq.sym)) return false;
      name = seq.sym;
      return true;
    }
    if (seq.t == Sx::LIST && seq.kids.size() >= 3 && seq.kids[0].t == Sx::SYM && seq.kids[0].sym == "GETF" && seq.kids[2].t == Sx::SYM) {
      std::string k = seq.kids[2].sym;
      if (!k.empty() && k[0] == ':') k = k.substr(1);
      if (!has_seq_key(k)) return false;
      name = k;
      return true;
    }
    return false;
  }
  int new_slot(const std::string& name) {
    int s = fn->nslots++;
    sc.names.push_back({name, s});
    return s;
  }
  Node konst(Val v) { Node n; n.op = O_CONST; n.c = v; return n; }
  Node var(int slot) { Node n; n.op = O_VAR; n.slot = slot; return n; }
  Node call(Op op, std::vector<Node> kids) { Node n; n.op = op; n.kids = std::move(kids); return n; }

  Node compile(const Sx& s) {
    if (s.t == Sx::NUM) return konst(s.num);
    if (s.t == Sx::SYM) {
      if (s.sym == "PI") return konst(Val::dbl(3.141592653589793));
      if (s.sym == "T") return konst(Val::boolean(true));
      if (s.sym == "NIL") return konst(Val::boolean(false));
      int slot = sc.find(s.sym);
      if (slot < 0) throw std::runtime_error("unbound variable " + s.sym);
      return var(slot);
    }
    if (s.kids.empty()) return konst(Val::boolean(false));
    const Sx& h = s.kids[0];
    if (h.t != Sx::SYM) throw std::runtime_error("cannot call a non-symbol");
    const std::string& f = h.sym;
    auto args = [&](size_t from = 1) {
      std::vector<Node> v;
      for (size_t i = from; i < s.kids.size(); i++) v.push_back(compile(s.kids[i]));
      return v;
    };
    if (f == "THE") return compile(s.kids.at(2));
    {
      // multi-dimensional inputs (walker-create docstring :1137): x is itself a sequence (x1 x2 ...) and the model
      // picks its components with elt / aref / svref / nth / first ... fourth / car / cadr / caddr
      static const struct { const char* name; int fixed; } acc[] = {
          {"ELT", -1}, {"AREF", -1}, {"SVREF", -1}, {"NTH", -2}, {"FIRST", 0}, {"SECOND", 1}, {"THIRD", 2},
          {"FOURTH", 3}, {"CAR", 0}, {"CADR", 1}, {"CADDR", 2}};
      for (auto& a : acc)
        if (f == a.name) {
          const bool fixed = a.fixed >= 0;
          if (s.kids.size() >= 2 && (a.fixed != -2 || s.kids.size() >= 3)) {
            // :single-item parameter style (:8-14, :1078-1088): the key's value is a list / vector / n x 1 array whose
            // element k is parameter NAME[k] of the flattened key list
            std::string sk;
            if (seq_key_of(s.kids[a.fixed == -2 ? 2 : 1], sk)) {
              long long idx = a.fixed;
              if (!fixed) {
                if (s.kids.size() < 3 || s.kids.size() > (f == "AREF" ? 4u : 3u)) throw std::runtime_error("wrong number of arguments to " + f);
                const Sx& ix = s.kids[a.fixed == -2 ? 1 : 2];
                if (ix.t != Sx::NUM || ix.num.k != Val::RAT || ix.num.den != 1 || ix.num.num < 0)
                  throw std::runtime_error(f + ": the index must be a literal non-negative integer");
                idx = ix.num.num;
                if (s.kids.size() == 4) {
                  const Sx& jx = s.kids[3];
                  if (jx.t != Sx::NUM || jx.num.k != Val::RAT || jx.num.den != 1 || jx.num.num != 0)
                    throw std::runtime_error("aref: the parameter array is n x 1, its second index must be 0");
                }
              } else if (s.kids.size() != 2) {
                throw std::runtime_error("wrong number of arguments to " + f);
              }
              const std::string want = sk + "[" + std::to_string(idx) + "]";
              for (size_t i = 0; i < keys.size(); i++)
                if (keys[i] == want) return var(theta_var((int)i));
              throw std::runtime_error(f + ": index beyond the parameter sequence " + sk);
            }
          }
          if (s.kids.size() != (fixed ? 2u : 3u)) throw std::runtime_error("wrong number of arguments to " + f);
          const Sx& seq = s.kids[a.fixed == -2 ? 2 : 1];
          if (seq.t != Sx::SYM || fn->x_slot < 0 || sc.find(seq.sym) != fn->x_slot)
            throw std::runtime_error(f + ": only the model's positional argument can be indexed");
          long long idx = a.fixed;
          if (!fixed) {
            const Sx& ix = s.kids[a.fixed == -2 ? 1 : 2];
            if (ix.t != Sx::NUM || ix.num.k != Val::RAT || ix.num.den != 1 || ix.num.num < 0)
              throw std::runtime_error(f + ": the index must be a literal non-negative integer");
            idx = ix.num.num;
          }
          Node n;
          n.op = O_XELT;
          n.slot = (int)idx;
          return n;
        }
    }
    if (f == "DECLARE") return konst(Val::boolean(false));
    if (f == "PROGN") return call(O_PROGN, args());
    if (f == "LET" || f == "LET*") {
      const Sx& binds = s.kids.at(1);
      Node n;
      n.op = O_LET;
      size_t mark = sc.names.size();
      std::vector<std::pair<std::string, int>> pending;
      bool star = f == "LET*";
      n.slot = fn->nslots;
      for (auto& b : binds.kids) {
        std::string name = b.t == Sx::SYM ? b.sym : b.kids.at(0).sym;
        Node init = (b.t == Sx::LIST && b.kids.size() > 1) ? compile(b.kids[1]) : konst(Val::boolean(false));
        int slot = fn->nslots++;
        if (star) sc.names.push_back({name, slot});
        else pending.push_back({name, slot});
        n.slots.push_back(slot);
        n.kids.push_back(std::move(init));
      }
      n.nb = (int)n.kids.size();
      for (auto& pnd : pending) sc.names.push_back(pnd);
      for (size_t i = 2; i < s.kids.size(); i++) {
        if (s.kids[i].t == Sx::LIST && !s.kids[i].kids.empty() && s.kids[i].kids[0].sym == "DECLARE") continue;
        n.kids.push_back(compile(s.kids[i]));
      }
      sc.names.resize(mark);
      return n;
    }
    if (f == "PRIOR-BOUNDS-LET") return compile_prior_bounds_let(s);
    if (f == "GETF") {  // (getf params :key default)
      const Sx& key = s.kids.at(2);
      std::string k = key.sym.substr(key.sym[0] == ':' ? 1 : 0);
      for (size_t i = 0; i < keys.size(); i++)
        if (keys[i] == k) return var(theta_var((int)i));
      if (s.kids.size() > 3) return compile(s.kids[3]);
      throw std::runtime_error("getf of missing key without default: " + k);
    }
    if (f == "IF") {
      std::vector<Node> a = args();
      if (a.size() == 2) a.push_back(konst(Val::boolean(false)));
      return call(O_IF, std::move(a));
    }
    if (f == "WHEN") {
      std::vector<Node> a;
      a.push_back(compile(s.kids.at(1)));
      a.push_back(call(O_PROGN, args(2)));
      a.push_back(konst(Val::boolean(false)));
      return call(O_IF, std::move(a));
    }
    if (f == "COND") {
      Node res = konst(Val::boolean(false));
      for (size_t i = s.kids.size() - 1; i >= 1; i--) {
        const Sx& cl = s.kids[i];
        std::vector<Node> a;
        a.push_back(compile(cl.kids.at(0)));
        std::vector<Node> body;
        for (size_t j = 1; j < cl.kids.size(); j++) body.push_back(compile(cl.kids[j]));
        a.push_back(call(O_PROGN, std::move(body)));
        a.push_back(std::move(res));
        res = call(O_IF, std::move(a));
      }
      return res;
    }
    static const struct { const char* name; Op op; } table[] = {
        {"+", O_ADD}, {"-", O_SUB}, {"*", O_MUL}, {"/", O_DIV}, {"1+", O_INC}, {"1-", O_DEC},
        {"EXP", O_EXP}, {"LOG", O_LOG}, {"SIN", O_SIN}, {"COS", O_COS}, {"TAN", O_TAN},
        {"ASIN", O_ASIN}, {"ACOS", O_ACOS}, {"ATAN", O_ATAN}, {"SINH", O_SINH}, {"COSH", O_COSH},
        {"TANH", O_TANH}, {"SQRT", O_SQRT}, {"EXPT", O_EXPT}, {"ABS", O_ABS}, {"MIN", O_MIN},
        {"MAX", O_MAX}, {"SIGNUM", O_SIGNUM}, {"AND", O_AND}, {"OR", O_OR}, {"NOT", O_NOT},
        {"<", O_LT}, {">", O_GT}, {"<=", O_LE}, {">=", O_GE}, {"=", O_EQ}, {"/=", O_NE},
        {"FLOAT", O_FLOAT}, {"LOG-NORMAL", O_LOGNORMAL}, {"LOG-POISSON", O_LOGPOISSON}, {"LOG-FACTORIAL", O_LOGFACT}};
    for (auto& e : table)
      if (f == e.name) return call(e.op, args());
    throw std::runtime_error("unsupported operator " + f);
  }

  int theta_var(int i) {
    if (fn->theta_slot[i] < 0) fn->theta_slot[i] = fn->nslots++;
    return fn->theta_slot[i];
  }

  // macro expansion of prior-bounds-let, :346-369
  Node compile_prior_bounds_let(const Sx& s) {
    const Sx& khl = s.kids.at(1);
    Node n;
    n.op = O_LET;
    size_t mark = sc.names.size();
    n.slot = fn->nslots;
    // (param (the double-float (getf params key 0d0)))  :353
    for (auto& e : khl.kids) {
      std::string key = e.kids.at(0).sym;
      std::string name = key.substr(key[0] == ':' ? 1 : 0);
      Node init = konst(Val::dbl(0.0));
      for (size_t i = 0; i < keys.size(); i++)
        if (keys[i] == name) init = var(theta_var((int)i));
      int slot = fn->nslots++;
      sc.names.push_back({name, slot});
      n.slots.push_back(slot);
      n.kids.push_back(std::move(init));
    }
    // (param-bound (if (< lo p hi) 0d0 (* -1d10 (- (exp (* (min (abs (- p hi)) (abs (- p lo))) 1d-5)) 1))))  :358-360
    std::vector<int> bound_slots;
    for (auto& e : khl.kids) {
      std::string key = e.kids.at(0).sym;
      std::string name = key.substr(key[0] == ':' ? 1 : 0);
      Node p = var(sc.find(name));
      Node test = call(O_LT, {compile(e.kids.at(1)), p, compile(e.kids.at(2))});
      Node d_hi = call(O_ABS, {call(O_SUB, {p, compile(e.kids.at(2))})});
      Node d_lo = call(O_ABS, {call(O_SUB, {p, compile(e.kids.at(1))})});
      Node m = call(O_MIN, {d_hi, d_lo});
      Node ex = call(O_EXP, {call(O_MUL, {m, konst(Val::dbl(1e-5))})});
      Node pen = call(O_MUL, {konst(Val::dbl(-1e10)), call(O_SUB, {ex, konst(Val::rat(1, 1))})});
      Node init = call(O_IF, {test, konst(Val::dbl(0.0)), pen});
      int slot = fn->nslots++;
      sc.names.push_back({name + "-BOUND", slot});
      bound_slots.push_back(slot);
      n.slots.push_back(slot);
      n.kids.push_back(std::move(init));
    }
    // (bounds-total (+ b1 b2 ...))  :368
    {
      std::vector<Node> bs;
      for (int sl : bound_slots) bs.push_back(var(sl));
      int slot = fn->nslots++;
      sc.names.push_back({"BOUNDS-TOTAL", slot});
      n.slots.push_back(slot);
      n.kids.push_back(call(O_ADD, std::move(bs)));
    }
    n.nb = (int)n.kids.size();
    for (size_t i = 2; i < s.kids.size(); i++) n.kids.push_back(compile(s.kids[i]));
    sc.names.resize(mark);
    return n;
  }
};

static bool truthy(const Val& v) { return !(v.k == Val::BOOL && !v.b); }

// the components of a multi-dimensional x for the call in progress (null: x is a number)
static thread_local const double* t_xvec = nullptr;
static thread_local int t_xdim = 0;

static Val eval(const Node& n, Val* env) {
  switch (n.op) {
    case O_CONST: return n.c;
    case O_XELT:
      if (!t_xvec) throw std::runtime_error("x is a number, not a sequence");
      if (n.slot >= t_xdim) throw std::runtime_error("index beyond the end of x");
      return Val::dbl(t_xvec[n.slot]);
    case O_VAR: return env[n.slot];
    case O_ADD: case O_MUL: {
      char op = n.op == O_ADD ? '+' : '*';
      if (n.kids.empty()) return Val::rat(n.op == O_ADD ? 0 : 1, 1);
      Val acc = eval(n.kids[0], env);
      for (size_t i = 1; i < n.kids.size(); i++) acc = arith(op, acc, eval(n.kids[i], env));
      return acc;
    }
    case O_SUB: case O_DIV: {
      char op = n.op == O_SUB ? '-' : '/';
      Val acc = eval(n.kids.at(0), env);
      if (n.kids.size() == 1) return arith(op, Val::rat(n.op == O_SUB ? 0 : 1, 1), acc);
      for (size_t i = 1; i < n.kids.size(); i++) acc = arith(op, acc, eval(n.kids[i], env));
      return acc;
    }
    case O_INC: return arith('+', eval(n.kids.at(0), env), Val::rat(1, 1));
    case O_DEC: return arith('-', eval(n.kids.at(0), env), Val::rat(1, 1));
    case O_EXP: return irrational(::exp, eval(n.kids.at(0), env));
    case O_LOG: {
      Val a = eval(n.kids.at(0), env);
      if (n.kids.size() == 2) {  // (log x base) = (/ (log x) (log base))
        Val b = eval(n.kids[1], env);
        return arith('/', irrational(::log, a), irrational(::log, b));
      }
      return irrational(::log, a);
    }
    case O_SIN: return irrational(::sin, eval(n.kids.at(0), env));
    case O_COS: return irrational(::cos, eval(n.kids.at(0), env));
    case O_TAN: return irrational(::tan, eval(n.kids.at(0), env));
    case O_ASIN: return irrational(::asin, eval(n.kids.at(0), env));
    case O_ACOS: return irrational(::acos, eval(n.kids.at(0), env));
    case O_ATAN: {
      Val a = eval(n.kids.at(0), env);
      if (n.kids.size() == 2) {
        Val b = eval(n.kids[1], env);
        if (a.k == Val::DBL || b.k == Val::DBL) return Val::dbl(::atan2(a.as_double(), b.as_double()));
        return Val::sgl((float)::atan2(a.as_double(), b.as_double()));
      }
      return irrational(::atan, a);
    }
    case O_SINH: return irrational(::sinh, eval(n.kids.at(0), env));
    case O_COSH: return irrational(::cosh, eval(n.kids.at(0), env));
    case O_TANH: return irrational(::tanh, eval(n.kids.at(0), env));
    case O_SQRT: return irrational(::sqrt, eval(n.kids.at(0), env));
    case O_EXPT: {
      Val b = eval(n.kids.at(0), env), p = eval(n.kids.at(1), env);
      if (p.k == Val::RAT && p.den == 1) return intexp(b, p.num);
      if (b.k == Val::DBL || p.k == Val::DBL) return Val::dbl(::pow(b.as_double(), p.as_double()));
      return Val::sgl((float)::pow(b.as_double(), p.as_double()));
    }
    case O_ABS: {
      Val a = eval(n.kids.at(0), env);
      if (a.k == Val::RAT) return Val::rat(a.num < 0 ? -a.num : a.num, a.den);
      if (a.k == Val::SGL) return Val::sgl(fabsf(a.f));
      return Val::dbl(fabs(a.d));
    }
    case O_MIN: case O_MAX: {
      Val acc = eval(n.kids.at(0), env);
      for (size_t i = 1; i < n.kids.size(); i++) {
        Val b = eval(n.kids[i], env);
        int c = cmp(b, acc);
        if ((n.op == O_MIN && c == -1) || (n.op == O_MAX && c == 1)) acc = b;
      }
      return acc;
    }
    case O_SIGNUM: {
      Val a = eval(n.kids.at(0), env);
      int c = cmp(a, Val::rat(0, 1));
      if (a.k == Val::RAT) return Val::rat(c, 1);
      if (a.k == Val::SGL) return Val::sgl(c == 0 ? a.f : (float)c);
      return Val::dbl(c == 0 ? a.d : (double)c);
    }
    case O_LOGNORMAL: {  // (log-normal x mu sigma) :372-377
      if (n.kids.size() != 3) throw std::runtime_error("log-normal takes (x mu sigma)");
      const Val a = eval(n.kids[0], env), b = eval(n.kids[1], env), c = eval(n.kids[2], env);
      if (a.k == Val::BOOL || b.k == Val::BOOL || c.k == Val::BOOL) throw std::runtime_error("log-normal of a non-number");
      return Val::dbl(orc_log_normal(a.as_double(), b.as_double(), c.as_double()));
    }
    case O_LOGPOISSON: {  // (log-poisson lambda k) :382-383
      if (n.kids.size() != 2) throw std::runtime_error("log-poisson takes (lambda k)");
      const Val a = eval(n.kids[0], env), b = eval(n.kids[1], env);
      if (a.k == Val::BOOL || b.k == Val::BOOL) throw std::runtime_error("log-poisson of a non-number");
      return Val::dbl(orc_log_poisson(a.as_double(), b.as_double()));
    }
    case O_LOGFACT: {  // (log-factorial n) :379-380: a single-float
      if (n.kids.size() != 1) throw std::runtime_error("log-factorial takes (n)");
      const Val a = eval(n.kids[0], env);
      if (a.k == Val::BOOL) throw std::runtime_error("log-factorial of a non-number");
      return Val::sgl((float)orc_log_factorial(a.as_double()));
    }
    case O_FLOAT: {
      Val a = eval(n.kids.at(0), env);
      if (n.kids.size() == 2) {
        Val proto = eval(n.kids[1], env);
        if (proto.k == Val::DBL) return Val::dbl(a.as_double());
        return Val::sgl(as_single(a));
      }
      if (a.k == Val::RAT) return Val::sgl(as_single(a));
      return a;
    }
    case O_IF: return truthy(eval(n.kids[0], env)) ? eval(n.kids[1], env) : eval(n.kids[2], env);
    case O_AND: {
      Val v = Val::boolean(true);
      for (auto& k : n.kids) { v = eval(k, env); if (!truthy(v)) return v; }
      return v;
    }
    case O_OR: {
      Val v = Val::boolean(false);
      for (auto& k : n.kids) { v = eval(k, env); if (truthy(v)) return v; }
      return v;
    }
    case O_NOT: return Val::boolean(!truthy(eval(n.kids.at(0), env)));
    case O_LT: case O_GT: case O_LE: case O_GE: case O_EQ: {
      Val prev = eval(n.kids.at(0), env);
      bool ok = true;
      for (size_t i = 1; i < n.kids.size(); i++) {
        Val cur = eval(n.kids[i], env);  // CL evaluates all arguments
        int c = cmp(prev, cur);
        bool t = n.op == O_LT ? c == -1 : n.op == O_GT ? c == 1 : n.op == O_LE ? (c == -1 || c == 0)
                 : n.op == O_GE ? (c == 1 || c == 0) : c == 0;
        ok = ok && t;
        prev = cur;
      }
      return Val::boolean(ok);
    }
    case O_NE: {
      std::vector<Val> vs;
      for (auto& k : n.kids) vs.push_back(eval(k, env));
      for (size_t i = 0; i < vs.size(); i++)
        for (size_t j = i + 1; j < vs.size(); j++)
          if (cmp(vs[i], vs[j]) == 0) return Val::boolean(false);
      return Val::boolean(true);
    }
    case O_LET: {
      for (int i = 0; i < n.nb; i++) env[n.slots[i]] = eval(n.kids[i], env);
      Val v = Val::boolean(false);
      for (size_t i = n.nb; i < n.kids.size(); i++) v = eval(n.kids[i], env);
      return v;
    }
    case O_PROGN: {
      Val v = Val::boolean(false);
      for (auto& k : n.kids) v = eval(k, env);
      return v;
    }
  }
  return Val::boolean(false);
}

Fn* fn_parse(const char* text, const std::vector<std::string>& keys_in) {
  Reader rd(text);
  Sx top = rd.read();
  if (top.t != Sx::LIST || top.kids.size() < 3) throw std::runtime_error("expected (lambda (args) body)");
  size_t li = 1;
  if (top.kids[0].sym == "DEFUN") li = 2;
  else if (top.kids[0].sym != "LAMBDA") throw std::runtime_error("expected lambda or defun");
  const Sx& ll = top.kids.at(li);
  std::unique_ptr<Fn> fn(new Fn());
  Compiler c;
  c.fn = fn.get();
  for (auto k : keys_in) {
    for (auto& ch : k) ch = (char)toupper((unsigned char)ch);
    if (!k.empty() && k[0] == ':') k = k.substr(1);
    c.keys.push_back(k);
  }
  fn->P = (int)c.keys.size();
  fn->theta_slot.assign(fn->P, -1);
  bool in_key = false, seen_key_marker = false;
  std::vector<std::string> positional;
  for (auto& a : ll.kids) {
    if (a.t == Sx::SYM && a.sym == "&KEY") { in_key = true; seen_key_marker = true; continue; }
    if (a.t == Sx::SYM && a.sym == "&ALLOW-OTHER-KEYS") continue;
    if (!in_key) { positional.push_back(a.sym); continue; }
    std::string name = a.t == Sx::SYM ? a.sym : a.kids.at(0).sym;
    int idx = -1;
    for (size_t i = 0; i < c.keys.size(); i++)
      if (c.keys[i] == name) idx = (int)i;
    if (idx >= 0) {
      int slot = c.theta_var(idx);
      c.sc.names.push_back({name, slot});
    } else if (c.has_seq_key(name)) {
      c.seq_vars.insert(name);
      int slot = c.new_slot(name);
      fn->defaults.push_back({slot, Val::boolean(true)});   // a sequence is no number
    } else {
      int slot = c.new_slot(name);
      Val dv = Val::boolean(false);
      if (a.t == Sx::LIST && a.kids.size() > 1) {
        if (a.kids[1].t != Sx::NUM) throw std::runtime_error("&key default must be a literal number");
        dv = a.kids[1].num;
      }
      fn->defaults.push_back({slot, dv});
    }
  }
  fn->is_prior = !seen_key_marker && positional.size() == 2;
  fn->is_ll = !seen_key_marker && positional.size() == 3;
  if (fn->is_ll) {
    fn->y_slot = c.new_slot(positional[0]);
    fn->model_slot = c.new_slot(positional[1]);
    fn->x_slot = c.new_slot(positional[2]);   // `error` is a sequence: (elt error k) etc. go through the same accessors as a sequence x
  } else if (!fn->is_prior) {
    if (positional.size() != 1) throw std::runtime_error("model lambda must take exactly one positional argument");
    fn->x_slot = c.new_slot(positional[0]);
  }
  std::vector<Node> body;
  for (size_t i = li + 1; i < top.kids.size(); i++) {
    const Sx& b = top.kids[i];
    if (b.t == Sx::LIST && !b.kids.empty() && b.kids[0].t == Sx::SYM && b.kids[0].sym == "DECLARE") continue;
    if (b.t == Sx::SYM && b.sym == "\"STR" && top.kids.size() > li + 2) continue;  // docstring
    body.push_back(c.compile(b));
  }
  fn->body = c.call(O_PROGN, std::move(body));
  return fn.release();
}

static double run(Fn* f, double x, const double* theta, bool x_is_sequence = false) {
  Val envbuf[256];
  std::vector<Val> big;
  Val* env = envbuf;
  if (f->nslots > 256) { big.resize(f->nslots); env = big.data(); }
  for (int i = 0; i < f->P; i++)
    if (f->theta_slot[i] >= 0) env[f->theta_slot[i]] = Val::dbl(theta[i]);
  for (auto& d : f->defaults) env[d.first] = d.second;
  if (f->x_slot >= 0) env[f->x_slot] = x_is_sequence ? Val::boolean(true) : Val::dbl(x);  // a sequence is no number
  Val v = eval(f->body, env);
  if (v.k == Val::BOOL) throw std::runtime_error("function returned a non-number");
  return v.as_double();
}

double fn_call_model(Fn* f, double x, const double* theta) { return run(f, x, theta); }
// (apply fn x params) with x a sequence of xdim numbers: using x itself as a number is a type error
double fn_call_model_v(Fn* f, const double* x, int xdim, const double* theta) {
  struct Bind {
    Bind(const double* x, int d) { t_xvec = x; t_xdim = d; }
    ~Bind() { t_xvec = nullptr; t_xdim = 0; }
  } bind(x, xdim);
  return run(f, std::numeric_limits<double>::quiet_NaN(), theta, true);
}
double fn_call_prior(Fn* f, const double* theta) { return run(f, 0.0, theta); }
// (funcall ll-fn y (apply fn x params) stddev) :416 -- stddev is the data set's whole error list
double fn_call_ll(Fn* f, double y, double model, const double* err, int n_err) {
  if (!f->is_ll) throw std::runtime_error("not a (lambda (y model error) ...)");
  struct Bind {
    Bind(const double* x, int d) { t_xvec = x; t_xdim = d; }
    ~Bind() { t_xvec = nullptr; t_xdim = 0; }
  } bind(err, n_err);
  Val envbuf[256];
  std::vector<Val> big;
  Val* env = envbuf;
  if (f->nslots > 256) { big.resize(f->nslots); env = big.data(); }
  env[f->y_slot] = Val::dbl(y);
  env[f->model_slot] = Val::dbl(model);
  env[f->x_slot] = Val::boolean(true);  // a list is no number
  Val v = eval(f->body, env);
  if (v.k == Val::BOOL) throw std::runtime_error("likelihood lambda returned a non-number");
  return v.as_double();
}
void fn_free(Fn* f) { delete f; }

}  // namespace orc

// ---------------------------------------------------------------- C face
struct orc_fn { orc::Fn* f; };

extern "C" orc_fn* orc_fn_parse(const char* text, const char* const* keys, int P, char* err, int errlen) {
  try {
    std::vector<std::string> ks;
    for (int i = 0; i < P; i++) ks.push_back(keys[i]);
    orc_fn* h = new orc_fn();
    h->f = orc::fn_parse(text, ks);
    return h;
  } catch (std::exception& e) {
    if (err && errlen > 0) snprintf(err, errlen, "%s", e.what());
    return nullptr;
  }
}
extern "C" void orc_fn_free(orc_fn* f) { if (f) { orc::fn_free(f->f); delete f; } }
extern "C" double orc_fn_call_model(orc_fn* f, double x, const double* theta) {
  try { return orc::fn_call_model(f->f, x, theta); }
  catch (std::exception& e) { if (getenv("ORC_DEBUG")) fprintf(stderr, "orc: %s\n", e.what()); return NAN; }
}
extern "C" double orc_fn_call_model_v(orc_fn* f, const double* x, int xdim, const double* theta) {
  try { return orc::fn_call_model_v(f->f, x, xdim, theta); }
  catch (std::exception& e) { if (getenv("ORC_DEBUG")) fprintf(stderr, "orc: %s\n", e.what()); return NAN; }
}
extern "C" double orc_fn_call_prior(orc_fn* f, const double* theta) {
  try { return orc::fn_call_prior(f->f, theta); }
  catch (std::exception& e) { if (getenv("ORC_DEBUG")) fprintf(stderr, "orc: %s\n", e.what()); return NAN; }
}
