// mfit_eval.cuh -- the batched log-likelihood kernel (sm_100a).
//
// This is the GPU form of hot loop #2 of the reference: log-liklihood-normal
// (mcmc-fitting.lisp:393-400), log-liklihood-normal-cutoff (:419-427) and the
// create-log-liklihood-function closure (:402-417) over log-poisson (:382-383), i.e.
//     ll_d(theta) = sum_i  density(y_i, f(x_i; theta), sigma_i)
// evaluated for W parameter vectors (chains) at once against shared (x, y, sigma).
//
// Work decomposition: one work item = one chain-block (MFIT_CB chains, one chain per thread) x one
// segment of the data (tiles_per_seg tiles of MFIT_TILE points); persistent CTAs walk contiguous ranges of items.  The data tile is moved
// global -> shared by the TMA engine (cp.async.bulk, 1-D) through an MFIT_STAGES-deep
// mbarrier ring and is read back as warp-uniform (broadcast) 128-bit LDS; theta lives in
// registers, one chain per lane, so the FP64 pipe sees no cross-lane traffic at all and
// each fetched point is reused by 128 chains.  Per-chain sums are sequential inside a
// segment (two interleaved accumulators, fixed order) and the per-segment partials are
// added in segment order by the chain kernel, so results do not depend on grid size, SM
// count or GPU count.
//
// The file is freestanding (its only #include is mfit_math.cuh, which embed_source.py inlines) because the very
// same text is embedded in libmfit_cuda.so and handed to NVRTC together with a model struct generated from a
// user s-expression (csrc/lower.cpp).
#ifndef MFIT_EVAL_CUH
#define MFIT_EVAL_CUH
#include "mfit_math.cuh"

#define MFIT_TILE 512    /* points per shared-memory stage */
#define MFIT_STAGES 3    /* TMA ring depth */
#define MFIT_CB 128      /* chains per CTA = threads per CTA */
#define MFIT_MAXK 16     /* parameters a built-in family may gather */

#define MFITK_NORMAL 0
#define MFITK_CUTOFF 1
#define MFITK_POISSON 2

struct MfitEvalArgs {
  const double* x;      /* [n_pad] */
  const double* y;      /* [n_pad]                      (cutoff, poisson) */
  const double* yw;     /* [n_pad] y_i/sigma_i          (normal) */
  const double* w;      /* [n_pad] 1/sigma_i            (normal, cutoff) */
  const double* c;      /* [n_pad] -ln(2pi)/2 - ln sigma_i  (cutoff only) */
  long long n;          /* valid points */
  double xmin, xmax;    /* range of x over the data set (-inf, +inf when some x is not finite): Model::safe */
  int tiles_per_seg;
  int n_seg;
  int n_cb;             /* chain blocks */
  int W;                /* chains */
  long long stride;     /* chain stride of theta / partial (W rounded up to MFIT_CB) */
  const double* theta;  /* [P][stride] parameter vectors to evaluate */
  const unsigned char* skip; /* [stride] 1 = chain finished, nothing to do (may be null) */
  double* partial;      /* [n_seg][stride] per-segment sums */
  int idx[MFIT_MAXK];   /* built-in families: family slot -> shared parameter index */
  int degree;
};

// ---------------------------------------------------------------- PTX helpers (mbarrier + TMA bulk copy)
__device__ __forceinline__ unsigned mfit_smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mfit_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mfit_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mfit_fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mfit_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mfit_smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mfit_mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mfit_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mfit_mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MFIT_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MFIT_DONE;\n"
      "bra MFIT_WAIT;\n"
      "MFIT_DONE:\n"
      "}\n" ::"r"(mfit_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA: global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void mfit_tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   mfit_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(mfit_smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- per-point terms
template <int LL>
struct MfitTerm;

// log-normal :372-377 with the theta-independent part hoisted: the kernel accumulates
// sum r^2, r = (y - f)/sigma = y/sigma - f/sigma with y/sigma precomputed at upload (two FP64 instructions per
// point); the chain kernel forms  C_d - sum/2.
template <>
struct MfitTerm<MFITK_NORMAL> {
  static constexpr int NARR = 3;
  static constexpr int USES = 0;
  template <bool SLOW, class G>
  __device__ __forceinline__ static double add(double acc, double f, double yw, double w, double, G&) {
    double r = fma(-f, w, yw);
    return fma(r, r, acc);
  }
};
// log-liklihood-normal-cutoff :426: (max -5000d0 (log-normal y f sigma)) cannot be hoisted
template <>
struct MfitTerm<MFITK_CUTOFF> {
  static constexpr int NARR = 4;
  static constexpr int USES = 0;
  template <bool SLOW, class G>
  __device__ __forceinline__ static double add(double acc, double f, double y, double w, double c, G&) {
    double r = (y - f) * w;
    double t = fma(-0.5 * r, r, c);
    t = (t > -5000.0) ? t : ((t != t) ? t : -5000.0);
    return acc + t;
  }
};
// log-poisson :382-383 with sum_i log k_i! hoisted:  k ln(lambda) - lambda
template <>
struct MfitTerm<MFITK_POISSON> {
  static constexpr int NARR = 2;
  static constexpr int USES = MFIT_USES_LOG;
  template <bool SLOW, class G>
  __device__ __forceinline__ static double add(double acc, double f, double y, double, double, G& gd) {
    return acc + fma(y, mfit_log_t<SLOW>(f, gd), -f);
  }
};

// ---------------------------------------------------------------- one data tile, one chain per thread
// MFIT_UNROLL_PAIRS pairs of points in flight per thread (two independent accumulators, fixed order).  SLOW = libdevice
// math, used only to redo a tile in which a fast routine met an argument outside its domain (see mfit_math.cuh);
// G = MfitGuard records such arguments, MfitNoGuard is for chains whose Model::safe() has ruled them out.
#ifndef MFIT_UNROLL_PAIRS
#define MFIT_UNROLL_PAIRS 4
#endif
template <class Model, int LL, bool SLOW, class G>
__device__ __forceinline__ void mfit_tile_pass(const Model& m, const double* xs, const double* ys, const double* ws,
                                               const double* cs, int cnt, double& acc0, double& acc1, G& gd) {
  const int pairs = cnt >> 1;
  const double2* x2 = reinterpret_cast<const double2*>(xs);
  const double2* y2 = reinterpret_cast<const double2*>(ys);
  const double2* w2 = reinterpret_cast<const double2*>(ws);
  const double2* c2 = reinterpret_cast<const double2*>(cs);
  constexpr int UNROLL = MFIT_UNROLL_PAIRS;
#pragma unroll UNROLL
  for (int i = 0; i < pairs; i++) {
    const double2 xv = x2[i], yv = y2[i], wv = w2[i], cv = c2[i];
    const double f0 = m.template eval<SLOW>(xv.x, gd);
    const double f1 = m.template eval<SLOW>(xv.y, gd);
    acc0 = MfitTerm<LL>::template add<SLOW>(acc0, f0, yv.x, wv.x, cv.x, gd);
    acc1 = MfitTerm<LL>::template add<SLOW>(acc1, f1, yv.y, wv.y, cv.y, gd);
  }
  if (cnt & 1) {
    const int i = cnt - 1;
    acc0 = MfitTerm<LL>::template add<SLOW>(acc0, m.template eval<SLOW>(xs[i], gd), ys[i], ws[i], cs[i], gd);
  }
}

// ---------------------------------------------------------------- the kernel body
// Persistent CTAs: the work items (segment, chain-block) -- chain-block fastest -- are split into gridDim.x
// contiguous ranges, one per CTA, and the grid is sized to what is resident at once (occupancy x SM count), so there
// is no wave quantisation and the prologue (mbarriers, math tables) is paid once per CTA.  The TMA ring runs across
// item boundaries: while the chains of one item are being evaluated the tiles of the next are already in flight, and
// the next item's parameters are loaded into registers one item ahead.
template <class Model, int LL>
__device__ __forceinline__ void mfit_eval_body(const MfitEvalArgs& a) {
  constexpr int NARR = MfitTerm<LL>::NARR;
  constexpr unsigned STAGE_BYTES = NARR * MFIT_TILE * 8u;
  extern __shared__ __align__(128) unsigned char mfit_smem[];
  double* buf = reinterpret_cast<double*>(mfit_smem);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(mfit_smem + MFIT_STAGES * STAGE_BYTES);
  unsigned long long* empty = full + MFIT_STAGES;

  const int tid = threadIdx.x;
  const long long total = (long long)a.n_cb * a.n_seg;
  const long long item_lo = total * blockIdx.x / gridDim.x;
  const long long item_hi = total * (blockIdx.x + 1) / gridDim.x;
  if (item_lo >= item_hi) return;
  const long long n_tiles = (a.n + MFIT_TILE - 1) / MFIT_TILE;
  auto tiles_of = [&](long long seg) -> int {
    long long t0 = seg * a.tiles_per_seg, t1 = t0 + a.tiles_per_seg;
    if (t1 > n_tiles) t1 = n_tiles;
    return (int)(t1 - t0);
  };

  mfit_math_stage(Model::USES | MfitTerm<LL>::USES, tid, MFIT_CB);  // exp/log tables -> shared memory
  if (tid == 0) {
    for (int s = 0; s < MFIT_STAGES; s++) {
      mfit_mbar_init(&full[s], 1);
      mfit_mbar_init(&empty[s], MFIT_CB / 32);
    }
    mfit_fence_mbar_init();
  }
  __syncthreads();

  // producer cursor (thread 0 only): next tile to request = tile p_k of item p_item; `issued` counts requests
  long long p_item = item_lo;
  int p_k = 0;
  unsigned issued = 0;
  auto issue_next = [&]() {
    const long long seg = p_item / a.n_cb;
    const int s = (int)(issued % MFIT_STAGES);
    const long long off = (seg * a.tiles_per_seg + p_k) * (long long)MFIT_TILE;
    double* dst = buf + (size_t)s * NARR * MFIT_TILE;
    mfit_mbar_expect_tx(&full[s], STAGE_BYTES);
    mfit_tma_load_1d(dst, a.x + off, MFIT_TILE * 8u, &full[s]);
    mfit_tma_load_1d(dst + MFIT_TILE, (LL == MFITK_NORMAL ? a.yw : a.y) + off, MFIT_TILE * 8u, &full[s]);
    if (NARR > 2) mfit_tma_load_1d(dst + 2 * MFIT_TILE, a.w + off, MFIT_TILE * 8u, &full[s]);
    if (NARR > 3) mfit_tma_load_1d(dst + 3 * MFIT_TILE, a.c + off, MFIT_TILE * 8u, &full[s]);
    issued++;
    if (++p_k == tiles_of(seg)) { p_k = 0; p_item++; }
  };
  if (tid == 0)
    for (int s = 0; s < MFIT_STAGES && p_item < item_hi; s++) issue_next();

  auto chain_of = [&](long long item) -> long long { return (item % a.n_cb) * MFIT_CB + tid; };
  auto clamp_chain = [&](long long ch) -> long long { return ch < a.W ? ch : (long long)a.W - 1; };

  // GUARDED: some fast routine on the path has a restricted domain (transcendentals of the model, log of log-poisson)
  constexpr bool GUARDED = Model::GUARDED || MfitTerm<LL>::USES != 0;
  Model m_next;
  m_next.load(a, clamp_chain(chain_of(item_lo)));  // parameters of the first item, while its tiles are in flight
  unsigned consumed = 0;
  for (long long item = item_lo; item < item_hi; item++) {
    const long long seg = item / a.n_cb;
    const long long chain = chain_of(item);
    const bool valid = chain < a.W;
    const long long cl = clamp_chain(chain);
    const bool live = valid && !(a.skip != nullptr && a.skip[cl] != 0);
    const bool any_live = __syncthreads_or(live ? 1 : 0) != 0;  // a finished chain-block only drains its tiles
    Model m = m_next;
    if (item + 1 < item_hi) m_next.load(a, clamp_chain(chain_of(item + 1)));
    m.setup();  // theta-only subexpressions
    // warp-uniform: every chain of this warp stays inside the fast domain over the whole x range of the data set
    const bool unguarded = !GUARDED || (MfitTerm<LL>::USES == 0 && __all_sync(0xffffffffu, m.safe(a) ? 1 : 0) != 0);
    const int nt = tiles_of(seg);
    const long long t0 = seg * a.tiles_per_seg;
    double acc0 = 0.0, acc1 = 0.0;
    for (int k = 0; k < nt; k++, consumed++) {
      const int s = (int)(consumed % MFIT_STAGES);
      // refill the stage freed one tile ago (gives the other warps a tile of slack)
      if (tid == 0 && consumed >= 1 && p_item < item_hi) {
        const unsigned c1 = consumed - 1;
        mfit_mbar_wait(&empty[c1 % MFIT_STAGES], (c1 / MFIT_STAGES) & 1u);
        issue_next();
      }
      mfit_mbar_wait(&full[s], (consumed / MFIT_STAGES) & 1u);
      if (any_live) {
        const double* xs = buf + (size_t)s * NARR * MFIT_TILE;
        const double* ys = xs + MFIT_TILE;
        const double* ws = xs + (NARR > 2 ? 2 : 0) * MFIT_TILE;
        const double* cs = xs + (NARR > 3 ? 3 : 0) * MFIT_TILE;
        long long rem = a.n - (t0 + k) * (long long)MFIT_TILE;
        const int cnt = rem < MFIT_TILE ? (int)rem : MFIT_TILE;
        if (unguarded) {
          MfitNoGuard ng;
          mfit_tile_pass<Model, LL, false>(m, xs, ys, ws, cs, cnt, acc0, acc1, ng);
        } else {
          const double save0 = acc0, save1 = acc1;
          MfitGuard gd;
          gd.reset();
          mfit_tile_pass<Model, LL, false>(m, xs, ys, ws, cs, cnt, acc0, acc1, gd);
          if (gd.tripped()) {  // rare: overflow / NaN / log of a non-positive number ... -> redo the tile with libdevice math
            acc0 = save0;
            acc1 = save1;
            mfit_tile_pass<Model, LL, true>(m, xs, ys, ws, cs, cnt, acc0, acc1, gd);
          }
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) mfit_mbar_arrive(&empty[s]);
    }
    if (live) a.partial[(size_t)seg * a.stride + chain] = acc0 + acc1;
  }
}

#endif  // MFIT_EVAL_CUH
