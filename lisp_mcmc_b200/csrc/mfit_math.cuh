// mfit_math.cuh -- FP64 exp / log / sin / cos built for DFMA-pipe-bound kernels (freestanding, NVRTC-safe).
//
// B200 issues one warp-wide DFMA every 2 cycles per SM sub-partition; the likelihood kernels are bound by that
// pipe (ncu: profiles/), so what matters is the number of FP64 instructions per call, and that the other
// instructions (integer, LDS, select) stay below one per FP64 instruction so that they hide in the pipe's shadow.
// CUDA's libdevice spends 22 / 46 / 32 DFMA issue slots on exp / log / sin (measured, mfit_measure_fp64); these
// spend 9 / 12 / 15 FP64 instructions: table-driven range reduction (Tang) with the tables staged in shared
// memory, short polynomials with coefficients taken straight from constant memory, exponent arithmetic on the
// integer pipe.  Accuracy is <= 2.4 ulp (tools/check_math.cpp measures it against long double), far inside the
// 1e-12 relative tolerance on log-likelihoods.
//
// Special values cost nothing on the hot path: every routine has a FAST form that is only valid on its regular
// domain and records, in a sticky per-thread guard (one integer max per call), whether an argument left that
// domain; the kernel checks the guard once per data tile and, if it tripped (overflow, NaN, log of a non-positive
// number, |x| >= 1e5 for sin/cos ...), recomputes that tile with the SLOW forms, which are the libdevice routines.
// Fast forms are memory-safe for any input (table indices are masked).
#ifndef MFIT_MATH_CUH
#define MFIT_MATH_CUH

#ifdef MFIT_HOST_TEST  // tools/check_math.cpp compiles this header for the CPU
#include <cmath>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define __constant__ const
static inline int mfit_hi(double x) { long long b; memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int mfit_lo(double x) { long long b; memcpy(&b, &x, 8); return (int)b; }
static inline double mfit_mk(int hi, int lo) {
  long long b = ((long long)hi << 32) | (unsigned)lo;
  double x; memcpy(&x, &b, 8); return x;
}
static inline unsigned mfit_umax(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
#ifndef MFIT_EXP_BITS
#define MFIT_EXP_BITS 8
#endif
#ifndef MFIT_GAUSS_DEG
#define MFIT_GAUSS_DEG 3
#endif
#else
__device__ __forceinline__ int mfit_hi(double x) { return __double2hiint(x); }
__device__ __forceinline__ int mfit_lo(double x) { return __double2loint(x); }
__device__ __forceinline__ double mfit_mk(int hi, int lo) { return __hiloint2double(hi, lo); }
__device__ __forceinline__ unsigned mfit_umax(unsigned a, unsigned b) { return max(a, b); }
// Shared-memory copies, filled by mfit_math_stage() at kernel start.  Table indices differ from lane to lane, so a
// plain copy would serialise on bank conflicts (ncu, profiles/: 2.8 extra wavefronts per exp, the shared-memory
// crossbar ran at 62 % next to a 68 % FP64 pipe).  Each table is therefore replicated so that every lane of a
// crossbar phase reads its own banks: 16 copies of the 8-byte exp entries (a 64-bit load is served in two phases of
// 16 lanes), 8 copies of the 16-byte log entries (four phases of 8 lanes).  Entry j of copy c sits at [j][c].
#ifndef MFIT_EXP_BITS
#define MFIT_EXP_BITS 8     /* the exp table holds 2^(j / 2^B): 8 or 9 (mfit_math_tables.inc has both; 10 needs gen_math_tables.py to emit it) */
#endif
#ifndef MFIT_EXP_COPIES
#define MFIT_EXP_COPIES 16  /* replication in shared memory: 16 = conflict-free */
#endif
#ifndef MFIT_GAUSS_DEG
#define MFIT_GAUSS_DEG 3    /* degree of the Gaussian's polynomial in the reduced argument: 4 (< 1e-17) or 3.  3 with the 256-entry table
                               is the shipped setting: every evaluation is one FP64 instruction shorter (C2 step 0.4115 -> 0.3921 ms,
                               profiles/r02_ab_deg.txt) and the Gaussian factor -- a number in (0, 1] that is added to a background --
                               carries a relative error of at most 5.6e-14 whose MEAN over the reduced argument is zero
                               (tools/gen_math_tables.py: a biased fit adds up coherently over a peak).  On log-likelihoods that is
                               far inside the 1e-12 they are held to and of the size of the reference's own summation error (2e-14 at
                               C2's 100 000 points, profiles/r02_fullsize_parity.json).  Larger tables cost more in bank conflicts /
                               shared memory than the lower degree buys (same file). */
#endif
__shared__ __align__(16) unsigned long long mfit_s_exp2[(1 << MFIT_EXP_BITS) * MFIT_EXP_COPIES];
__shared__ __align__(16) double mfit_s_log[128 * 8 * 2];
#endif

#include "mfit_math_tables.inc"

#define MFIT_USES_EXP 1
#define MFIT_USES_LOG 2

#ifndef MFIT_HOST_TEST
// every thread of the CTA calls this before the first __syncthreads of the kernel
__device__ __forceinline__ void mfit_math_stage(int uses, int tid, int nthreads) {
  if (uses & MFIT_USES_EXP)
    for (int i = tid; i < (1 << MFIT_EXP_BITS) * MFIT_EXP_COPIES; i += nthreads) mfit_s_exp2[i] = mfit_c_exp2[i / MFIT_EXP_COPIES];
  if (uses & MFIT_USES_LOG)
    for (int i = tid; i < 128 * 8; i += nthreads) {
      mfit_s_log[2 * i] = mfit_c_log[2 * (i >> 3)];
      mfit_s_log[2 * i + 1] = mfit_c_log[2 * (i >> 3) + 1];
    }
}
#endif

// sticky record of the most extreme argument each fast routine has seen since reset()
struct MfitGuard {
  static constexpr bool FUSED = false;
  static constexpr bool ON = true;     // records arguments (see_*), checked once per tile
  static constexpr bool CLAMP = true;  // in-place clamps of the routines that have one (mfit_gauss_reduce)
  unsigned e, g, l, s, d, b;
  __device__ __forceinline__ void reset() { e = 0u; g = 0u; l = 0u; s = 0u; d = 0u; b = 0u; }
  __device__ __forceinline__ void see_exp(unsigned v) { e = mfit_umax(e, v); }
  __device__ __forceinline__ void see_gauss(unsigned v) { g = mfit_umax(g, v); }
  __device__ __forceinline__ void see_log(unsigned v) { l = mfit_umax(l, v); }
  __device__ __forceinline__ void see_sin(unsigned v) { s = mfit_umax(s, v); }
  __device__ __forceinline__ void see_div(unsigned v) { d = mfit_umax(d, v); }
  __device__ __forceinline__ void see_big(unsigned v) { b = mfit_umax(b, v); }
  __device__ __forceinline__ bool tripped() const {
    return e >= 0x40862000u      // exp: |x| >= 708 (overflow / underflow), inf, nan
        || g > 0x7ff00000u       // gauss: nan (underflow and +-inf are clamped in place, see mfit_gauss_parts)
        || l >= 0x7fe00000u      // log: x <= 0, subnormal, inf, nan   (l tracks hi(x) - 0x00100000 as unsigned)
        || s >= 0x40f86a00u      // sin/cos: |x| >= 1e5, inf, nan
        || b >= 0x7fe00000u      // an intermediate product at or beyond 2^1023, inf, nan (MfitNvMulti's paired numerator)
        || d >= 0x7fb00000u;     // rcp: |x| outside [2^-1021, 2^1022), zero, inf, nan   (d tracks |hi(x)| - 0x00200000)
  }
  // some log argument was a negative normal number (hi(x) in [0x80100000, 0xffefffff]): that term is NaN, hence so
  // is every sum it enters -- no need to redo anything (wild proposals of the reference's start-up phase drive
  // a log-poisson rate below zero all the time)
  __device__ __forceinline__ bool log_of_negative() const { return l >= 0x80000000u && l < 0xfff00000u; }
};
// The same interface with nothing behind it: used when a model has proved, from its parameters and the x range of
// the data set (Model::safe), that no argument can leave the fast domain -- the guard and clamp instructions
// (two to three integer-pipe slots per call) then disappear from the loop.
struct MfitNoGuard {
  static constexpr bool ON = false;
  static constexpr bool CLAMP = false;
  static constexpr bool FUSED = false;  // see MfitFusedNoGuard
  __device__ __forceinline__ void reset() {}
  __device__ __forceinline__ void see_exp(unsigned) {}
  __device__ __forceinline__ void see_gauss(unsigned) {}
  __device__ __forceinline__ void see_log(unsigned) {}
  __device__ __forceinline__ void see_sin(unsigned) {}
  __device__ __forceinline__ void see_div(unsigned) {}
  __device__ __forceinline__ void see_big(unsigned) {}
  __device__ __forceinline__ bool tripped() const { return false; }
  __device__ __forceinline__ bool log_of_negative() const { return false; }
};

// In between: the clamps stay, nothing is recorded.  For a model all of whose fast routines clamp in place (the
// Gaussian) and whose parameters are finite (Model::tame) no argument can produce anything but the right value, so
// there is nothing to check per tile -- one integer min per call instead of three integer instructions and a redo
// path.  This is what the wild proposals of the reference's start-up phase run on (its first l-matrix is diag(theta)).
struct MfitClampGuard : MfitNoGuard {
  static constexpr bool CLAMP = true;
};

// No guard, and the model may contract its argument arithmetic where Model::fusable() has shown that to be harmless
// (the Gaussian forms (x - mu)/s as one FMA when |mu| is not large next to s: the rounding of the hoisted product then
// acts like a change of mu in its last place or two).
struct MfitFusedNoGuard : MfitNoGuard {
  static constexpr bool FUSED = true;
};

/* 1.5 * 2^52 + 2^31: adding it rounds to an integer k held in the low mantissa word.  The 2^31 bias keeps the low
 * word free of borrows for |k| < 2^31 and vanishes from every use of k below (k & 63, k << 14, n & 3 are all taken
 * modulo a power of two that divides 2^31). */
#define MFIT_MAGIC 6755401588539392.0

// 2^(k/256) for the integer k held (biased) in the low word of tmp = q + MAGIC: table entry (k & 255), exponent
// k >> 8 added on the integer pipe (the table entries are pre-biased so that a plain k << 12 does both jobs).
// Two integer instructions and one conflict-free LDS.64.  256 entries (32 KB replicated) instead of 64 buy one
// polynomial degree, i.e. one FP64 instruction per call -- the pipe this code is bound by.
__device__ __forceinline__ double mfit_exp2_256th(double tmp) {
  const int k = mfit_lo(tmp);
#ifdef MFIT_HOST_TEST
  const unsigned long long t = mfit_c_exp2[k & ((1 << MFIT_EXP_BITS) - 1)];
#elif MFIT_EXP_BITS != 8 || MFIT_EXP_COPIES != 16
  // other table shapes (experiments): mask, then one multiply-add
  const unsigned base = (unsigned)__cvta_generic_to_shared(mfit_s_exp2) + ((threadIdx.x & (MFIT_EXP_COPIES - 1u)) << 3);
  unsigned long long t;
  const unsigned addr = ((unsigned)k & ((1u << MFIT_EXP_BITS) - 1u)) * (8u * MFIT_EXP_COPIES) + base;
  asm("ld.shared.u64 %0, [%1];" : "=l"(t) : "r"(addr));
#else
  // address = table + 8 (lane mod 16) + 128 (k & 255) in ONE instruction: an integer dot product of the bytes of k
  // with (128, 0, 0, 0) masks, scales and adds the base at once (IDP.4A).  Left to itself the compiler shifts,
  // masks-and-merges and adds (three instructions); a mask + multiply-add is two.  This loop is short of issue slots, not
  // of integer throughput: C2 step 0.4411 (three) -> 0.4308 (two) -> 0.4049 ms (one, with the other changes of round 2;
  // profiles/r02_kernel_shapes.md).
  const unsigned base = (unsigned)__cvta_generic_to_shared(mfit_s_exp2) + ((threadIdx.x & 15u) << 3);
  unsigned long long t;
  const unsigned addr = __dp4a((unsigned)k, 0x80u, base);
  asm("ld.shared.u64 %0, [%1];" : "=l"(t) : "r"(addr));
#endif
  return mfit_mk((int)(t >> 32) + (k << (20 - MFIT_EXP_BITS)), (int)t);
}

// exp(x): 9 FP64 instructions
template <bool SLOW, class G>
__device__ __forceinline__ double mfit_exp_t(double x, G& gd) {
  if (SLOW) return exp(x);
  if (G::ON) gd.see_exp((unsigned)mfit_hi(x) & 0x7fffffffu);
  const double tmp = fma(x, MFIT_K_L2E256 * (double)(1 << (MFIT_EXP_BITS - 8)), MFIT_MAGIC);  // x 2^B/ln2
  const double kd = tmp - MFIT_MAGIC;
  double r = fma(kd, -0x1.62e42fef00000p-9 / (double)(1 << (MFIT_EXP_BITS - 8)), x);  // ln2/2^B in two pieces (Cody-Waite); the first is a 32-bit constant
  r = fma(kd, -0x1.473de6af278edp-42 / (double)(1 << (MFIT_EXP_BITS - 8)), r);
  double p = fma(r, MFIT_K_EXN4, MFIT_K_EXN3);           // |r| <= ln2/512: degree 4 is good to 4e-17
  p = fma(r, p, 0.5);
  p = fma(r, p, 1.0);
  p = fma(r, p, 1.0);
  return mfit_exp2_256th(tmp) * p;
}

/* zs = z sqrt(2^(B-1)/ln2); |zs| beyond ZLIMIT (zs^2 = 1023 2^B: exp(-z^2/2) = 2^-1023) needs the clamp of the guarded form,
 * whose bound is the high word of a double just below it */
#if MFIT_EXP_BITS == 8
#define MFIT_GAUSS_ZLIMIT 511.75
#define MFIT_GAUSS_ZCLAMP_HI 0x407ffc00u
#define MFIT_GAUSS_ZSCALE 0x1.b2da4e9808a53p+3 /* sqrt(128/ln2) */
#elif MFIT_EXP_BITS == 9
#define MFIT_GAUSS_ZLIMIT 723.5
#define MFIT_GAUSS_ZCLAMP_HI 0x40869c00u
#define MFIT_GAUSS_ZSCALE (0x1.b2da4e9808a53p+3 * 0x1.6a09e667f3bcdp+0)
#else
#define MFIT_GAUSS_ZLIMIT 1023.375
#define MFIT_GAUSS_ZCLAMP_HI 0x408ffb00u
#define MFIT_GAUSS_ZSCALE 0x1.b2da4e9808a53p+4
#endif

// exp(-zs^2 ln2/256) = s * (1 + r q)  (zs = z sqrt(128/ln2) turns exp(-z^2/2) into this form): the square is both the
// argument and, through one FMA each, its rounding to 1/256 octaves and the exact remainder r.  Three FP64
// instructions here; the caller runs the polynomial (coefficients MFIT_K_EXS1..4, which it may pre-multiply by its
// amplitude) and folds s:  f = fma(s, a (1 + r q), rest).
// zs must not be negative (callers pass |d| |1/s|: the absolute values are operand modifiers of the multiply, free).
template <class G>
__device__ __forceinline__ void mfit_gauss_reduce(double zs, double* s, double* r, G& gd) {
  if (G::CLAMP) {
    // zs is clamped on the integer pipe at 511.75 (exp(-z^2/2) ~ 2^-1023, which the arithmetic below turns into
    // a subnormal-sized number, i.e. zero next to any amplitude): wild proposals underflow all the time during the
    // reference's start-up phase (its first l-matrix is diag(theta)), so that case must not leave the fast path.
    // One unsigned min on the high word: +inf clamps too, a NaN (0x7ff8...) would as well -- which is why the
    // recording guard looks at the word first and Model::tame() rules NaN out for the clamp-only guard.
    const unsigned ah = (unsigned)mfit_hi(zs);
    if (G::ON) gd.see_gauss(ah);
    zs = mfit_mk((int)min(ah, MFIT_GAUSS_ZCLAMP_HI), mfit_lo(zs));
  }
  const double tmp = fma(-zs, zs, MFIT_MAGIC);
  const double kd = tmp - MFIT_MAGIC;
  *r = fma(-zs, zs, -kd);
  *s = mfit_exp2_256th(tmp);
}
// the plain form (amplitude 1): exp(-zs^2 ln2/256) = s * p
template <class G>
__device__ __forceinline__ void mfit_gauss_parts(double zs, double* s, double* p, G& gd) {
  double r;
  mfit_gauss_reduce(zs, s, &r, gd);
#if MFIT_EXP_BITS == 8 && MFIT_GAUSS_DEG == 4
  double q = fma(r, MFIT_K_EXS4, MFIT_K_EXS3);
  q = fma(r, q, MFIT_K_EXS2);
  q = fma(r, q, MFIT_K_EXS1);
#elif MFIT_GAUSS_DEG == 4
  double q = fma(r, mfit_kg[MFIT_EXP_BITS - 8][1][3], mfit_kg[MFIT_EXP_BITS - 8][1][2]);
  q = fma(r, q, mfit_kg[MFIT_EXP_BITS - 8][1][1]);
  q = fma(r, q, mfit_kg[MFIT_EXP_BITS - 8][1][0]);
#else
  double q = fma(r, mfit_kg[MFIT_EXP_BITS - 8][0][2], mfit_kg[MFIT_EXP_BITS - 8][0][1]);
  q = fma(r, q, mfit_kg[MFIT_EXP_BITS - 8][0][0]);
#endif
  *p = fma(r, q, 1.0);
}


// log(x): 12 FP64 instructions.  x = 2^e m with m in [0.707, 1.414); r = m c_i - 1 (one FMA, exact before rounding),
// log x = e ln2 + L_i + log1p(r); the table intervals next to 1 have c = 1 (1/2) and L = 0 exactly, which keeps full
// relative accuracy around x = 1.
template <bool SLOW, class G>
__device__ __forceinline__ double mfit_log_t(double x, G& gd) {
  if (SLOW) return log(x);
  const int hx = mfit_hi(x);
  if (G::ON) gd.see_log((unsigned)(hx - 0x00100000));
  const int i = (hx >> 13) & 127;
  const int e = ((hx + 0x00096000) >> 20) - 1023;  // +1 exactly when i >= 53 (m >= 1.4140625 is halved)
  const double m = mfit_mk((hx & 0x000fffff) | 0x3ff00000, mfit_lo(x));
#ifdef MFIT_HOST_TEST
  const double c = mfit_c_log[2 * i], L = mfit_c_log[2 * i + 1];
#else
  const char* tab = reinterpret_cast<const char*>(mfit_s_log) + ((threadIdx.x & 7u) << 4);
  const double2 cl = *reinterpret_cast<const double2*>(tab + (i << 7));
  const double c = cl.x, L = cl.y;
#endif
  const double r = fma(m, c, -1.0);
  const double ed = mfit_mk(0x43300000, e ^ 0x80000000) - 0x1.0000080000000p+52;  // int -> double via the exponent trick
  const double r2 = r * r;
  double q = fma(r, MFIT_K_LG7, MFIT_K_LG6);             // |r| < 2^-7.9: the r^8 term of log1p is below 2e-18 of the result
  q = fma(r, q, MFIT_K_LG5);
  q = fma(r, q, -0.25);
  q = fma(r, q, MFIT_K_LG3);
  q = fma(r, q, -0.5);
  const double base = fma(ed, 0x1.62e42fef00000p-1, L);  // e ln2_hi is exact (32-bit constant)
  const double lp = fma(r2, q, r);                        // log1p(r)
  return fma(ed, 0x1.473de6af278edp-34, lp) + base;       // ln2_lo rides on the small term
}

// 1/x: the MUFU.RCP64H seed (about 22 bits) and two Newton steps -- 4 FP64 instructions and no branch, against the
// 8 FP64 instructions, the range test and the out-of-line slow path of the compiler's IEEE division (ncu: the
// Lorentzian loop spent more issue slots on that scaffolding than on FP64).  About 1 ulp, not correctly rounded.
template <bool SLOW, class G>
__device__ __forceinline__ double mfit_rcp_t(double x, G& gd) {
  if (SLOW) return 1.0 / x;
  if (G::ON) gd.see_div(((unsigned)mfit_hi(x) & 0x7fffffffu) - 0x00200000u);
  double y;
#ifdef MFIT_HOST_TEST
  y = (double)(float)(1.0 / x);
#else
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#endif
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// sin(x): x = n pi + r (Cody-Waite, three pieces: exact for |n| < 2^20), |r| <= pi/2, sin x = (-1)^n sin r with ONE odd
// polynomial of degree 17 for sin r -- 15 FP64 instructions and two integer ones for the sign.  (A quadrant scheme
// with separate sine and cosine polynomials saves one FP64 instruction but has to pick its coefficients per lane:
// 18 selects per call, which made the sinusoid kernels issue-bound instead of FP64-bound.)
__device__ __forceinline__ double mfit_sin_core(double x) {
  const double tmp = fma(x, MFIT_K_1OPI, MFIT_MAGIC);
  const double nd = tmp - MFIT_MAGIC;
  double r = fma(nd, -0x1.921fb54400000p+1, x);
  r = fma(nd, -0x1.0b4611a600000p-33, r);
  r = fma(nd, -0x1.3198a2e037073p-68, r);
  const double r2 = r * r;
  double q = fma(r2, MFIT_K_SP17, MFIT_K_SP15);
  q = fma(r2, q, MFIT_K_SP13);
  q = fma(r2, q, MFIT_K_SP11);
  q = fma(r2, q, MFIT_K_SP9);
  q = fma(r2, q, MFIT_K_SP7);
  q = fma(r2, q, MFIT_K_SP5);
  q = fma(r2, q, MFIT_K_SP3);
  const double res = fma(r2 * r, q, r);
  return mfit_mk(mfit_hi(res) ^ (mfit_lo(tmp) << 31), mfit_lo(res));  // odd n: negate
}
template <bool SLOW, class G>
__device__ __forceinline__ double mfit_sin_t(double x, G& gd) {
  if (SLOW) return sin(x);
  if (G::ON) gd.see_sin((unsigned)mfit_hi(x) & 0x7fffffffu);
  return mfit_sin_core(x);
}
// cos x = sin(x + pi/2); the shift is added in two pieces so that its rounding error stays below that of the reduction
template <bool SLOW, class G>
__device__ __forceinline__ double mfit_cos_t(double x, G& gd) {
  if (SLOW) return cos(x);
  if (G::ON) gd.see_sin((unsigned)mfit_hi(x) & 0x7fffffffu);
  // n = round(x/pi + 1/2) would need a half-integer magic constant; instead reduce |x| first and shift the remainder
  const double t2 = fma(fabs(x), MFIT_K_1OPI, MFIT_MAGIC);
  const double nd = t2 - MFIT_MAGIC;
  double r = fma(nd, -0x1.921fb54400000p+1, fabs(x));
  r = fma(nd, -0x1.0b4611a600000p-33, r);
  r = fma(nd, -0x1.3198a2e037073p-68, r);       // |x| = n pi + r, |r| <= pi/2;  cos x = (-1)^n cos r = (-1)^n sin(pi/2 - |r|)
  double rr = 0x1.921fb54442d18p+0 - fabs(r);
  rr = rr + 0x1.1a62633145c07p-54;
  const double r2 = rr * rr;
  double q = fma(r2, MFIT_K_SP17, MFIT_K_SP15);
  q = fma(r2, q, MFIT_K_SP13);
  q = fma(r2, q, MFIT_K_SP11);
  q = fma(r2, q, MFIT_K_SP9);
  q = fma(r2, q, MFIT_K_SP7);
  q = fma(r2, q, MFIT_K_SP5);
  q = fma(r2, q, MFIT_K_SP3);
  const double res = fma(r2 * rr, q, rr);
  return mfit_mk(mfit_hi(res) ^ (mfit_lo(t2) << 31), mfit_lo(res));
}

#endif  // MFIT_MATH_CUH
