// mfit_models.cuh -- hand-written FP64 model families for mfit_eval.cuh (freestanding, NVRTC-safe).
//
// The reference ships no model families (README.md:11 asks users to write them); SURVEY.md 8(d)
// defines these.  The canonical Lisp form of each family is in lisp_mcmc_b200/families.py and is what
// the oracle evaluates through the reference's generic (apply fn x params) path (mcmc-fitting.lisp:400).
// Here theta-only subexpressions (reciprocals, gamma^2, scale*gamma^2) are hoisted into load() so that
// the per-point code is FMA-dense; every deviation from the natural operation order is <= a few ulp of f.
#ifndef MFIT_MODELS_CUH
#define MFIT_MODELS_CUH

#define MFIT_TH(k) a.theta[(size_t)a.idx[k] * a.stride + chain]

// Every model struct provides:
//   load(a, chain)   read the chain's parameters (issued one work item ahead, so only raw values are kept live)
//   setup()          theta-only subexpressions, once per work item
//   safe(a)          true when no argument of a fast math routine can leave its domain for any x in [a.xmin, a.xmax]
//                    (NaN parameters make every comparison false -> not safe -> guarded path -> libdevice redo)
//   eval<SLOW>(x,gd) the model function
//   positive(a)      true when f(x) is a positive normal number for every x in [a.xmin, a.xmax] (called after setup, only
//                    together with safe): a likelihood term that takes log f -- log-poisson -- then needs no guard either
//   GUARDED          whether eval uses a fast routine with a restricted domain at all
//   FUSABLE          the model has a shorter evaluation for MfitFusedNoGuard, valid for chains with fusable(a) (and safe(a))
//   CLAMP_ONLY       every such routine clamps in place (MfitClampGuard): chains with tame() parameters need no per-tile
//                    check even where safe() cannot be proved; tame(a) is only called when CLAMP_ONLY
//   XDIMS, XVEC      components of x the model reads; XVEC: x is a sequence and the model provides evalv(xv[XDIMS], gd)
//                    instead of eval (lowered user models only, csrc/lower.cpp)

__device__ __forceinline__ double mfit_absmax(double u, double v) { return fmax(fabs(u), fabs(v)); }

// f = c0 + x (c1 + x (c2 + ... ))   -- Horner, D FMAs
template <int D>
struct MfitPoly {
  static constexpr int USES = 0;
  static constexpr int XDIMS = 1;
  static constexpr bool XVEC = false;
  static constexpr bool GUARDED = false;
  static constexpr bool CLAMP_ONLY = false;
  static constexpr bool FUSABLE = false;
  double c[D + 1];
  __device__ __forceinline__ void load(const MfitEvalArgs& a, long long chain) {
#pragma unroll
    for (int k = 0; k <= D; k++) c[k] = MFIT_TH(k);
  }
  __device__ __forceinline__ void setup() {}
  __device__ __forceinline__ bool safe(const MfitEvalArgs&) const { return true; }
  __device__ __forceinline__ bool positive(const MfitEvalArgs&) const { return D == 0 && c[0] > 1e-300 && c[0] < 1e300; }
  template <bool SLOW, class G>
  __device__ __forceinline__ double eval(double x, G& gd) const {
    double r = c[D];
#pragma unroll
    for (int k = D - 1; k >= 0; k--) r = fma(x, r, c[k]);
    return r;
  }
};

// f = a exp(b x) + c
struct MfitExpo {
  static constexpr int USES = MFIT_USES_EXP;
  static constexpr int XDIMS = 1;
  static constexpr bool XVEC = false;
  static constexpr bool GUARDED = true;
  static constexpr bool CLAMP_ONLY = false;
  static constexpr bool FUSABLE = false;
  double a_, b_, c_;
  __device__ __forceinline__ void load(const MfitEvalArgs& a, long long chain) {
    a_ = MFIT_TH(0); b_ = MFIT_TH(1); c_ = MFIT_TH(2);
  }
  __device__ __forceinline__ void setup() {}
  __device__ __forceinline__ bool safe(const MfitEvalArgs& a) const { return fabs(b_) * mfit_absmax(a.xmin, a.xmax) < 700.0; }
  // a >= 0, c > 0: f >= c; |b x| < 600 keeps a exp(b x) below 1e291
  __device__ __forceinline__ bool positive(const MfitEvalArgs& a) const {
    return a_ >= 0.0 && a_ < 1e30 && c_ > 1e-300 && c_ < 1e290 && fabs(b_) * mfit_absmax(a.xmin, a.xmax) < 600.0;
  }
  template <bool SLOW, class G>
  __device__ __forceinline__ double eval(double x, G& gd) const { return fma(a_, mfit_exp_t<SLOW>(b_ * x, gd), c_); }
};

// f = a sin(omega x + phi) + c
struct MfitSinus {
  static constexpr int USES = 0;
  static constexpr int XDIMS = 1;
  static constexpr bool XVEC = false;
  static constexpr bool GUARDED = true;
  static constexpr bool CLAMP_ONLY = false;
  static constexpr bool FUSABLE = false;
  double a_, om, ph, c_;
  __device__ __forceinline__ void load(const MfitEvalArgs& a, long long chain) {
    a_ = MFIT_TH(0); om = MFIT_TH(1); ph = MFIT_TH(2); c_ = MFIT_TH(3);
  }
  __device__ __forceinline__ void setup() {}
  __device__ __forceinline__ bool safe(const MfitEvalArgs& a) const {
    return fabs(om) * mfit_absmax(a.xmin, a.xmax) + fabs(ph) < 9.0e4;
  }
  // f lies in [c - |a|, c + |a|]
  __device__ __forceinline__ bool positive(const MfitEvalArgs&) const { return c_ - fabs(a_) > 1e-300 && c_ + fabs(a_) < 1e300; }
  template <bool SLOW, class G>
  __device__ __forceinline__ double eval(double x, G& gd) const { return fma(a_, mfit_sin_t<SLOW>(fma(om, x, ph), gd), c_); }
};

// f = a gamma^2 / ((x - x0)^2 + gamma^2) + c.  The quotient is a reciprocal (mfit_rcp_t) folded into the final FMA:
// 7 FP64 instructions per point, 9 with the log-normal term.
struct MfitLorentz {
  static constexpr int USES = 0;
  static constexpr int XDIMS = 1;
  static constexpr bool XVEC = false;
  static constexpr bool GUARDED = true;
  static constexpr bool CLAMP_ONLY = false;
  static constexpr bool FUSABLE = false;
  double ag2, x0, g2, c_;
  __device__ __forceinline__ void load(const MfitEvalArgs& a, long long chain) {
    ag2 = MFIT_TH(0); x0 = MFIT_TH(1); g2 = MFIT_TH(2); c_ = MFIT_TH(3);
  }
  __device__ __forceinline__ void setup() {
    g2 = g2 * g2;
    ag2 = ag2 * g2;
  }
  __device__ __forceinline__ bool safe(const MfitEvalArgs& a) const {
    const double dm = mfit_absmax(a.xmin - x0, a.xmax - x0);
    return g2 > 1e-290 && fma(dm, dm, g2) < 1e290;
  }
  // after setup: ag2 = a gamma^2 >= 0 and c > 0 give c <= f <= c + ag2 / g2
  __device__ __forceinline__ bool positive(const MfitEvalArgs&) const {
    return ag2 >= 0.0 && ag2 < 1e150 && g2 > 1e-140 && c_ > 1e-300 && c_ < 1e290;
  }
  template <bool SLOW, class G>
  __device__ __forceinline__ double eval(double x, G& gd) const {
    const double d = x - x0;
    if (SLOW) return ag2 / fma(d, d, g2) + c_;
    return fma(ag2, mfit_rcp_t<false>(fma(d, d, g2), gd), c_);
  }
};

// f = a exp(-((x - mu)/s)^2 / 2) + b0 + b1 x.  The Gaussian argument is formed directly in 1/256 octaves
// (zs = z sqrt(128/ln2), exp(-z^2/2) = 2^(-zs^2/256)), see mfit_gauss_reduce.  12 FP64 instructions per point (d, zs,
// lin, tmp, kd, r, four polynomial FMAs, a s, the final FMA), 14 with the log-normal term.  The polynomial takes its
// coefficients from the constant bank: folding the amplitude into per-chain coefficient registers saves the a s
// multiply but turns four FMAs into three-register-operand ones, which the FP64 pipe issues more slowly
// (tools/probe/horner_probe.cu: 56 against 63 DFMA/clk/SM; loop_probe: 9.98e11 against 1.03e12 point-evals/s).
struct MfitGaussLin {
  static constexpr int USES = MFIT_USES_EXP;
  static constexpr int XDIMS = 1;
  static constexpr bool XVEC = false;
  static constexpr bool GUARDED = true;
  static constexpr bool CLAMP_ONLY = true;
  static constexpr bool FUSABLE = true;
  double a_, mu, isq, b0, b1;  // load() leaves s in isq
  double nmi;                  // -mu isq (fused form)
  __device__ __forceinline__ void load(const MfitEvalArgs& a, long long chain) {
    a_ = MFIT_TH(0); mu = MFIT_TH(1); isq = MFIT_TH(2); b0 = MFIT_TH(3); b1 = MFIT_TH(4);
  }
  __device__ __forceinline__ void setup() {
    isq = fabs(1.0 / isq) * MFIT_GAUSS_ZSCALE;  // only its square matters
    nmi = -mu * isq;
  }
  __device__ __forceinline__ bool safe(const MfitEvalArgs& a) const {
    return mfit_absmax(a.xmin - mu, a.xmax - mu) * isq < MFIT_GAUSS_ZLIMIT && fabs(a_) < 1.0e300;
  }
  // zs = fma(x, isq, -mu isq) instead of (x - mu) isq saves one of 14 FP64 instructions per point.  The hoisted product is
  // rounded once, which shifts every argument by up to |mu isq| 2^-53 -- exactly what a change of mu by |mu| 2^-53 would
  // do, i.e. the result is the exact model of a parameter vector within half an ulp of the given one.  That is invisible
  // as long as the log-likelihood is not steeper in mu than ~1e3 per unit of relative change, guaranteed here by
  // |mu| <= 8 |s| (tests/test_gpu_parity.py::test_gauss_far_from_origin is the counter-example: |mu / s| up to 9500, 1.4e-11).
  __device__ __forceinline__ bool fusable(const MfitEvalArgs&) const { return fabs(nmi) <= 8.0 * MFIT_GAUSS_ZSCALE; }
  // a >= 0 and a background that is positive at both ends of the x range (it is linear in between)
  __device__ __forceinline__ bool positive(const MfitEvalArgs& a) const {
    const double l0 = fma(b1, a.xmin, b0), l1 = fma(b1, a.xmax, b0);
    return a_ >= 0.0 && a_ < 1e290 && l0 > 1e-300 && l1 > 1e-300 && l0 < 1e290 && l1 < 1e290;
  }
  // finite parameters and data: zs = |x - mu| |1/s| ZSCALE is then a number or +inf, never NaN, and the clamp of
  // mfit_gauss_reduce turns everything beyond 511.75 into an amplitude of a 2^-1023 (below 1e-28 for |a| < 1e280)
  __device__ __forceinline__ bool tame(const MfitEvalArgs& a) const {
    return fabs(a_) < 1.0e280 && fabs(mu) < 1.0e300 && isq < 1.0e300 && fabs(b0) < 1.0e300 && fabs(b1) < 1.0e300 &&
           a.xmin > -1.0e300 && a.xmax < 1.0e300;
  }
  template <bool SLOW, class G>
  __device__ __forceinline__ double eval(double x, G& gd) const {
    const double d = x - mu;
    const double lin = fma(b1, x, b0);
    if (SLOW) {
      const double z = d * (isq * (1.0 / MFIT_GAUSS_ZSCALE));
      return fma(a_, exp(-0.5 * (z * z)), lin);
    }
    double s, p;
    if (G::FUSED) mfit_gauss_parts(fma(x, isq, nmi), &s, &p, gd);   // sign-free: only MfitFusedNoGuard (no clamp) comes here
    else mfit_gauss_parts(fabs(d) * isq, &s, &p, gd);
    return fma(a_ * s, p, lin);
  }
};

// f = bg0 - sum_k scale_k sigma^2 / ((x - mu_k)^2 + sigma^2);  theta = scale_1..K mu_1..K sigma bg0
// Dips are taken in PAIRS over one reciprocal: s0/q0 + s1/q1 = (s0 q1 + s1 q0) / (q0 q1) -- 12 FP64 instructions and one
// MUFU per pair instead of 14 and two (the reference's double Lorentzian, nv-specific.lisp, is one pair; a last odd dip
// costs 7 as in MfitLorentz).  A few ulp of f either way.  The product needs twice the exponent range: safe() bounds
// q and the amplitudes by 1e140, and the recording guard also watches the numerator (see_big).
template <int K>
struct MfitNvMulti {
  static constexpr int USES = 0;
  static constexpr int XDIMS = 1;
  static constexpr bool XVEC = false;
  static constexpr bool GUARDED = true;
  static constexpr bool CLAMP_ONLY = false;
  static constexpr bool FUSABLE = false;
  double sg[K], mu[K], g2, bg0;
  __device__ __forceinline__ void load(const MfitEvalArgs& a, long long chain) {
    g2 = MFIT_TH(2 * K);
    bg0 = MFIT_TH(2 * K + 1);
#pragma unroll
    for (int k = 0; k < K; k++) { sg[k] = MFIT_TH(k); mu[k] = MFIT_TH(K + k); }
  }
  __device__ __forceinline__ void setup() {
    g2 = g2 * g2;
#pragma unroll
    for (int k = 0; k < K; k++) sg[k] = sg[k] * g2;
  }
  // sufficient and cheap (one compare per parameter; it runs once per chain in the walker-set kernels): centres and data
  // below 1e69 in magnitude keep (x - mu)^2 below 4e138, so every q = (x - mu)^2 + g2 lies in (1e-140, 1e140)
  __device__ __forceinline__ bool safe(const MfitEvalArgs& a) const {
    bool ok = g2 > 1e-140 && g2 < 1e139 && a.xmin > -1e69 && a.xmax < 1e69;
#pragma unroll
    for (int k = 0; k < K; k++) ok = ok && fabs(mu[k]) < 1e69 && fabs(sg[k]) < 1e140;
    return ok;
  }
  // after setup sg[k] = scale_k sigma^2: a dip is at most sg[k] / g2 deep (at its centre)
  __device__ __forceinline__ bool positive(const MfitEvalArgs&) const {
    double down = 0.0, up = 0.0;
#pragma unroll
    for (int k = 0; k < K; k++) { down += fmax(sg[k], 0.0); up += fmax(-sg[k], 0.0); }
    return g2 > 1e-140 && bg0 - down / g2 > 1e-300 && bg0 + up / g2 < 1e290;
  }
  template <bool SLOW, class G>
  __device__ __forceinline__ double eval(double x, G& gd) const {
    double r = bg0;
    if (SLOW) {
#pragma unroll
      for (int k = 0; k < K; k++) {
        const double d = x - mu[k];
        r -= sg[k] / fma(d, d, g2);
      }
      return r;
    }
    if constexpr (K >= 2) {
#pragma unroll
      for (int k = 0; k + 1 < K; k += 2) {
        const double d0 = x - mu[k], d1 = x - mu[k + 1];
        const double q0 = fma(d0, d0, g2), q1 = fma(d1, d1, g2);
        const double num = fma(sg[k], q1, sg[k + 1] * q0);
        if (G::ON) gd.see_big((unsigned)mfit_hi(num) & 0x7fffffffu);
        r = fma(-num, mfit_rcp_t<false>(q0 * q1, gd), r);
      }
    }
    if (K & 1) {
      const double d = x - mu[K - 1];
      r = fma(-sg[K - 1], mfit_rcp_t<false>(fma(d, d, g2), gd), r);
    }
    return r;
  }
};

#endif  // MFIT_MODELS_CUH
