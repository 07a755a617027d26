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

// f = c0 + x (c1 + x (c2 + ... ))   -- Horner, D FMAs
template <int D>
struct MfitPoly {
  static constexpr int USES = 0;
  double c[D + 1];
  __device__ __forceinline__ void load(const MfitEvalArgs& a, long long chain) {
#pragma unroll
    for (int k = 0; k <= D; k++) c[k] = MFIT_TH(k);
  }
  template <bool SLOW>
  __device__ __forceinline__ double eval(double x, MfitGuard& gd) const {
    double r = c[D];
#pragma unroll
    for (int k = D - 1; k >= 0; k--) r = fma(x, r, c[k]);
    return r;
  }
};

// f = a exp(b x) + c
struct MfitExpo {
  static constexpr int USES = MFIT_USES_EXP;
  double a_, b_, c_;
  __device__ __forceinline__ void load(const MfitEvalArgs& a, long long chain) {
    a_ = MFIT_TH(0); b_ = MFIT_TH(1); c_ = MFIT_TH(2);
  }
  template <bool SLOW>
  __device__ __forceinline__ double eval(double x, MfitGuard& gd) const { return fma(a_, mfit_exp_t<SLOW>(b_ * x, gd), c_); }
};

// f = a sin(omega x + phi) + c
struct MfitSinus {
  static constexpr int USES = 0;
  double a_, om, ph, c_;
  __device__ __forceinline__ void load(const MfitEvalArgs& a, long long chain) {
    a_ = MFIT_TH(0); om = MFIT_TH(1); ph = MFIT_TH(2); c_ = MFIT_TH(3);
  }
  template <bool SLOW>
  __device__ __forceinline__ double eval(double x, MfitGuard& gd) const { return fma(a_, mfit_sin_t<SLOW>(fma(om, x, ph), gd), c_); }
};

// f = a gamma^2 / ((x - x0)^2 + gamma^2) + c
struct MfitLorentz {
  static constexpr int USES = 0;
  double ag2, x0, g2, c_;
  __device__ __forceinline__ void load(const MfitEvalArgs& a, long long chain) {
    const double g = MFIT_TH(2);
    g2 = g * g;
    ag2 = MFIT_TH(0) * g2;
    x0 = MFIT_TH(1);
    c_ = MFIT_TH(3);
  }
  template <bool SLOW>
  __device__ __forceinline__ double eval(double x, MfitGuard& gd) const {
    const double d = x - x0;
    return ag2 / fma(d, d, g2) + c_;
  }
};

// f = a exp(-((x - mu)/s)^2 / 2) + b0 + b1 x.  The Gaussian argument is formed directly in 1/64 octaves
// (zs = z sqrt(32/ln2), exp(-z^2/2) = 2^(-zs^2/64)), see mfit_gauss_parts: 15 FP64 instructions per point in all.
struct MfitGaussLin {
  static constexpr int USES = MFIT_USES_EXP;
  double a_, mu, is_, isq, b0, b1;
  __device__ __forceinline__ void load(const MfitEvalArgs& a, long long chain) {
    a_ = MFIT_TH(0); mu = MFIT_TH(1); is_ = 1.0 / MFIT_TH(2); isq = is_ * MFIT_GAUSS_ZSCALE; b0 = MFIT_TH(3); b1 = MFIT_TH(4);
  }
  template <bool SLOW>
  __device__ __forceinline__ double eval(double x, MfitGuard& gd) const {
    const double d = x - mu;
    const double lin = fma(b1, x, b0);
    if (SLOW) {
      const double z = d * is_;
      return fma(a_, exp(-0.5 * (z * z)), lin);
    }
    double s, p;
    mfit_gauss_parts(d * isq, &s, &p, gd);
    return fma(a_ * s, p, lin);
  }
};

// f = bg0 - sum_k scale_k sigma^2 / ((x - mu_k)^2 + sigma^2);  theta = scale_1..K mu_1..K sigma bg0
template <int K>
struct MfitNvMulti {
  static constexpr int USES = 0;
  double sg[K], mu[K], g2, bg0;
  __device__ __forceinline__ void load(const MfitEvalArgs& a, long long chain) {
    const double g = MFIT_TH(2 * K);
    g2 = g * g;
    bg0 = MFIT_TH(2 * K + 1);
#pragma unroll
    for (int k = 0; k < K; k++) { sg[k] = MFIT_TH(k) * g2; mu[k] = MFIT_TH(K + k); }
  }
  template <bool SLOW>
  __device__ __forceinline__ double eval(double x, MfitGuard& gd) const {
    double r = bg0;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const double d = x - mu[k];
      r -= sg[k] / fma(d, d, g2);
    }
    return r;
  }
};

#endif  // MFIT_MODELS_CUH
