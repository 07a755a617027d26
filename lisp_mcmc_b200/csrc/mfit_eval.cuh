// mfit_eval.cuh -- the batched log-likelihood kernel (sm_100a).
//
// This is the GPU form of hot loop #2 of the reference: log-liklihood-normal
// (mcmc-fitting.lisp:393-400), log-liklihood-normal-cutoff (:419-427) and the
// create-log-liklihood-function closure (:402-417) over log-poisson (:382-383), i.e.
//     ll_d(theta) = sum_i  density(y_i, f(x_i; theta), sigma_i)
// evaluated for W parameter vectors (chains) at once against shared (x, y, sigma).
//
// Work decomposition: one work item = one warp-block (MFIT_WB = 64 chains, two chains per lane) x one segment of
// the data (tiles_per_seg tiles of MFIT_TILE points).  Persistent CTAs; each of their warps takes items from a
// global counter on its own, chain-block fastest.  The data tile is moved global -> shared by the TMA engine
// (cp.async.bulk, 1-D) through a per-warp mbarrier ring that runs across item boundaries, and is read back
// as warp-uniform (broadcast) 128-bit LDS; theta lives in registers, so the FP64 pipe sees no cross-lane traffic.
//
// Why two chains per thread: a broadcast LDS.128 still returns 512 bytes to the register file and the shared-memory
// return path moves 128 B/clk/SM, i.e. 4 cycles per load, shared by the four schedulers.  With one chain per thread
// the 24 bytes of (x, y/sigma, 1/sigma) cost 6 of those cycles per warp-point and the kernel ran at the speed of
// that path, not of the FP64 pipe (tools/probe/loop_probe.cu: 5th-order polynomial at 62 % of the DFMA rate, ncu:
// profiles/); every fetched point now feeds two evaluations.
//
// Per-chain sums are sequential inside a segment (two interleaved accumulators, fixed order) and the per-segment
// partials are added in segment order by the chain kernel, so results do not depend on grid size, on which CTA took
// which item, on SM count or on GPU count.
//
// The file is freestanding (its only #include is mfit_math.cuh, which embed_source.py inlines) because the very
// same text is embedded in libmfit_cuda.so and handed to NVRTC together with a model struct generated from a
// user s-expression (csrc/lower.cpp).
#ifndef MFIT_EVAL_CUH
#define MFIT_EVAL_CUH
#include "mfit_math.cuh"

#ifndef MFIT_TILE
#define MFIT_TILE 128    /* points per shared-memory stage */
#endif
#ifndef MFIT_STAGES
#define MFIT_STAGES 3    /* per-warp TMA ring depth (the host lowers it when shared memory is short, MfitMultiArgs::stages) */
#endif
#define MFIT_MAX_SEG 1024 /* segments (= work items per chain-block, = partial sums per chain) a data set is cut into at most */
#ifndef MFIT_WARPS
#define MFIT_WARPS 4     /* warps per CTA.  Small CTAs: a CTA leaves the SM as soon as ITS warps find no more items, and a CTA
                            of the next launch (the next chain group's, on another stream) takes the slot -- with one
                            12-warp CTA per SM the hand-over waits for the slowest of twelve (0.4615 against 0.4411 ms per
                            C2 step, profiles/r02_kernel_shapes.md) */
#endif
#define MFIT_THREADS (32 * MFIT_WARPS) /* threads per CTA */
#define MFIT_WB 64       /* chains per work item: two per lane (lane l: chains l and l + 32 of the warp-block) */
#define MFIT_CB 256      /* granule of the chain dimension on the host: strides, chain groups (a multiple of MFIT_WB) */
#ifndef MFIT_MIN_CTAS
#define MFIT_MIN_CTAS 3   /* CTAs per SM the register allocation is bounded for: 3 x 128 threads x 168 registers */
#endif
#define MFIT_PC_THREADS 256 /* threads per CTA of the per-walker-data kernel (one warp per chain) */
#define MFIT_QN 8        /* per-warp ring of item ids handed from the producer side of the warp to its consumer side */
#define MFIT_MAXK 16     /* parameters a built-in family may gather */

#define MFITK_NORMAL 0
#define MFITK_CUTOFF 1
#define MFITK_POISSON 2
#define MFITK_USER 3     /* a lowered (lambda (y model error) ...): MfitTerm<MFITK_USER> is generated text (csrc/lower.cpp kind 2) */

struct MfitEvalArgs {
  const double* x;      /* [n_pad]; multi-dimensional inputs: component k at x + k * xd_stride */
  const double* y;      /* [n_pad]                      (cutoff, poisson) */
  const double* yw;     /* [n_pad] y_i/sigma_i          (normal) */
  const double* w;      /* [n_pad] 1/sigma_i            (normal, cutoff) */
  const double* c;      /* [n_pad] -ln(2pi)/2 - ln sigma_i  (cutoff only) */
  long long n;          /* valid points */
  double xmin, xmax;    /* range of x over the data set (-inf, +inf when some x is not finite): Model::safe */
  int tiles_per_seg;
  int n_seg;
  int n_cb;             /* chain blocks this launch covers */
  int cb0;              /* first of them (a run may split the chains into groups that advance on separate streams) */
  int W;                /* chains */
  long long stride;     /* chain stride of theta / partial (W rounded up to MFIT_CB) */
  const double* theta;  /* [P][stride] parameter vectors to evaluate */
  const unsigned char* skip; /* [stride] 1 = chain finished, nothing to do (may be null) */
  double* partial;      /* [n_seg][stride] per-segment sums */
  /* walker sets with per-walker data (mfit_eval_perchain_body): chain w reads x + w * x_stride (0: one x for all),
   * y / yw + w * pc_stride, w / c + w * w_stride; chains [cb0 * MFIT_CB, (cb0 + n_cb) * MFIT_CB) of this launch */
  long long x_stride, pc_stride, w_stride;  /* w_stride: rows of w / c (0: one sigma vector for all chains) */
  int ring_stages;      /* walker sets, staged kernel: > 0 = every warp keeps this many y rows in flight through its own
                           TMA ring in shared memory (mfit_perchain_ring); 0 = rows are read straight into registers */
  long long xd_stride;  /* distance between the component arrays of a multi-dimensional x (n_pad) */
  int idx[MFIT_MAXK];   /* built-in families: family slot -> shared parameter index */
  int degree;
  int diag;             /* measurement aid of the walker-set ring kernels (MFIT_PC_DIAG): 1 = rows are fetched, nothing is
                           evaluated; 2 = everything is evaluated, nothing is fetched (results are then meaningless) */
};

// One launch of the shared-data kernel evaluates up to MFIT_MAX_FUSED data sets that use the same model family and
// likelihood kind (the data sets of a global fit usually do): their work items share one queue, so a global fit costs
// one launch per step instead of one per data set and its last CTAs wait for the tail of one queue, not of eight.
#define MFIT_MAX_FUSED 8
struct MfitMultiArgs {
  int n_ds;
  int stages;                         /* depth of every warp's TMA ring (>= 1; 2 or more to overlap fetch and arithmetic) */
  int item_start[MFIT_MAX_FUSED + 1]; /* items [item_start[k], item_start[k+1]) belong to d[k]; chain-block fastest inside */
  int* sched;                         /* [2] zero between launches: next item to hand out, CTAs that have finished */
  unsigned long long* mode_tiles;     /* [4] or null: tiles evaluated per MFIT_MODE_* (measurement: which loop ran how often) */
  MfitEvalArgs d[MFIT_MAX_FUSED];
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

// ---------------------------------------------------------------- one data tile
// NC chains per thread (2, or 1 for a chain-block whose second half is empty), UNROLL pairs of points per iteration:
// 2 NC UNROLL evaluations in flight per thread, two accumulators per chain, fixed order.  SLOW = libdevice math, used
// only to redo a tile in which a fast routine met an argument outside its domain (see mfit_math.cuh); G = MfitGuard
// records such arguments, MfitNoGuard is for chains whose Model::safe() has ruled them out.
#ifndef MFIT_UNROLL_PAIRS
#define MFIT_UNROLL_PAIRS 4   /* 16 evaluations per iteration: same instruction count per point as 2, but the scheduler sees
                                 more independent work across the loop's back edge (C2 step 0.4201 -> 0.4116 ms) */
#endif
// f(x_i) for a model whose x is a number (eval) or a sequence of Model::XDIMS components (evalv; the component arrays
// of a stage lie MFIT_TILE apart)
template <class Model, bool SLOW, class G>
__device__ __forceinline__ double mfit_model_at(const Model& m, const double* xs, int i, G& gd) {
  if constexpr (Model::XVEC) {
    double xv[Model::XDIMS];
#pragma unroll
    for (int d = 0; d < Model::XDIMS; d++) xv[d] = xs[d * MFIT_TILE + i];
    return m.template evalv<SLOW>(xv, gd);
  } else {
    return m.template eval<SLOW>(xs[i], gd);
  }
}

// f(x) for the kernels that only take a number for x (walker sets with per-walker data, the fit envelope); the host
// refuses to launch them for a model that reads x as a sequence
template <class Model, bool SLOW, class G>
__device__ __forceinline__ double mfit_model_scalar(const Model& m, double x, G& gd) {
  if constexpr (Model::XVEC) return __longlong_as_double(0x7ff8000000000000LL);
  else return m.template eval<SLOW>(x, gd);
}

template <class Model, int LL, bool SLOW, int NC, class G>
__device__ __forceinline__ void mfit_tile_pass(const Model* m, const double* xs, const double* ys, const double* ws,
                                               const double* cs, int cnt, double* acc, G* gd) {
  const int pairs = cnt >> 1;
  const double2* x2 = reinterpret_cast<const double2*>(xs);
  const double2* y2 = reinterpret_cast<const double2*>(ys);
  const double2* w2 = reinterpret_cast<const double2*>(ws);
  const double2* c2 = reinterpret_cast<const double2*>(cs);
  constexpr int UNROLL = MFIT_UNROLL_PAIRS * (NC == 1 ? 2 : 1);
#pragma unroll UNROLL
  for (int i = 0; i < pairs; i++) {
    const double2 yv = y2[i], wv = w2[i], cv = c2[i];
    if constexpr (Model::XVEC) {
      double xa[Model::XDIMS], xb[Model::XDIMS];
#pragma unroll
      for (int d = 0; d < Model::XDIMS; d++) {
        const double2 xv = x2[d * (MFIT_TILE / 2) + i];
        xa[d] = xv.x;
        xb[d] = xv.y;
      }
#pragma unroll
      for (int c = 0; c < NC; c++) {
        const double f0 = m[c].template evalv<SLOW>(xa, gd[c]);
        const double f1 = m[c].template evalv<SLOW>(xb, gd[c]);
        acc[2 * c] = MfitTerm<LL>::template add<SLOW>(acc[2 * c], f0, yv.x, wv.x, cv.x, gd[c]);
        acc[2 * c + 1] = MfitTerm<LL>::template add<SLOW>(acc[2 * c + 1], f1, yv.y, wv.y, cv.y, gd[c]);
      }
    } else {
      const double2 xv = x2[i];
#pragma unroll
      for (int c = 0; c < NC; c++) {
        const double f0 = m[c].template eval<SLOW>(xv.x, gd[c]);
        const double f1 = m[c].template eval<SLOW>(xv.y, gd[c]);
        acc[2 * c] = MfitTerm<LL>::template add<SLOW>(acc[2 * c], f0, yv.x, wv.x, cv.x, gd[c]);
        acc[2 * c + 1] = MfitTerm<LL>::template add<SLOW>(acc[2 * c + 1], f1, yv.y, wv.y, cv.y, gd[c]);
      }
    }
  }
  if (cnt & 1) {
    const int i = cnt - 1;
#pragma unroll
    for (int c = 0; c < NC; c++)
      acc[2 * c] = MfitTerm<LL>::template add<SLOW>(acc[2 * c], mfit_model_at<Model, SLOW>(m[c], xs, i, gd[c]), ys[i], ws[i], cs[i], gd[c]);
  }
}

// fast pass with whatever protection the chains of this warp need (warp-uniform `mode`): MFIT_MODE_FREE = none
// (Model::safe holds for every chain), MFIT_MODE_CLAMP = in-place clamps only (Model::CLAMP_ONLY and Model::tame),
// MFIT_MODE_GUARD = recording guard; a tripped guard redoes the tile with libdevice
#define MFIT_MODE_FREE 0
#define MFIT_MODE_CLAMP 1
#define MFIT_MODE_GUARD 2
#define MFIT_MODE_FUSED 3   /* MFIT_MODE_FREE and Model::fusable for every chain (Model::FUSABLE) */
template <class Model, int LL, int NC>
__device__ __forceinline__ void mfit_tile(const Model* m, int mode, const double* xs, const double* ys,
                                          const double* ws, const double* cs, int cnt, double* acc) {
  constexpr bool GUARDED = Model::GUARDED || MfitTerm<LL>::USES != 0;
  if (Model::FUSABLE && mode == MFIT_MODE_FUSED) {
    MfitFusedNoGuard fg[NC];
    mfit_tile_pass<Model, LL, false, NC>(m, xs, ys, ws, cs, cnt, acc, fg);
  } else if (!GUARDED || mode == MFIT_MODE_FREE) {
    MfitNoGuard ng[NC];
    mfit_tile_pass<Model, LL, false, NC>(m, xs, ys, ws, cs, cnt, acc, ng);
  } else if (Model::CLAMP_ONLY && MfitTerm<LL>::USES == 0 && mode == MFIT_MODE_CLAMP) {
    MfitClampGuard cg[NC];
    mfit_tile_pass<Model, LL, false, NC>(m, xs, ys, ws, cs, cnt, acc, cg);
  } else {
    double save[2 * NC];
#pragma unroll
    for (int k = 0; k < 2 * NC; k++) save[k] = acc[k];
    MfitGuard gd[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) gd[c].reset();
    mfit_tile_pass<Model, LL, false, NC>(m, xs, ys, ws, cs, cnt, acc, gd);
#pragma unroll
    for (int c = 0; c < NC; c++) {
      if (!gd[c].tripped()) continue;
      if ((MfitTerm<LL>::USES & MFIT_USES_LOG) != 0 && (Model::USES & MFIT_USES_LOG) == 0 && gd[c].log_of_negative()) {
        // the only log on the path is log-poisson's own, (log lambda) :383, and its NaN reaches the sum whatever else happened
        acc[2 * c] = __longlong_as_double(0x7ff8000000000000LL);
      } else {  // rare: overflow / NaN / zero rate ... -> this chain's tile again with libdevice
        acc[2 * c] = save[2 * c];
        acc[2 * c + 1] = save[2 * c + 1];
        mfit_tile_pass<Model, LL, true, 1>(m + c, xs, ys, ws, cs, cnt, acc + 2 * c, gd + c);
      }
    }
  }
}

// ---------------------------------------------------------------- the kernel body
// Persistent CTAs (MFIT_MIN_CTAS per SM); every WARP is an independent worker.  A warp draws work items -- (segment, block of
// MFIT_WB = 64 chains), chain-block fastest -- from a.sched[0] until they run out, streams the item's data tiles
// through ITS OWN TMA ring (cp.async.bulk 1-D into per-warp shared memory, one mbarrier per stage, lane 0 issues) and
// keeps two chains per lane in registers.  Nothing couples the warps of a CTA after the prologue (math tables): no
// CTA-wide barrier, no shared producer.  (The first version of this kernel had one producer thread per 4-warp CTA: the
// producer's warp could not run ahead of the CTA's slowest warp, all three producers of an SM sat on the same
// scheduler, and the kernel ran at the pace of that scheduler -- 9 % below the bare loop, tools/probe/loop_probe.cu.)
// The data set is L2-resident and tiny next to the arithmetic (3 KB per 128 points x 64 chains x ~30 FP64 instructions),
// so fetching a tile once per warp instead of once per CTA costs nothing measurable.
// The ring runs across item boundaries: the next item is drawn (atomicAdd) one tile before its first tile is requested,
// its parameters are loaded one item ahead, so neither latency is exposed.
// MULTI = false: one data set, whose arguments are then compile-time-addressed kernel parameters (constant-bank
// operands); with MULTI the per-item data set index makes them indexed loads, which costs the one-data-set case 5 %.
template <class Model, int LL, bool MULTI>
__device__ __forceinline__ void mfit_eval_body_t(const MfitMultiArgs& ma) {
  constexpr int XD = Model::XDIMS;                  // component arrays of x (1 unless x is a sequence)
  constexpr int LNARR = MfitTerm<LL>::NARR;          // x, y, [w, [c]] as the likelihood term counts them
  constexpr int NARR = LNARR + XD - 1;               // arrays per stage: XD of x, then y, w, c
  constexpr unsigned STAGE_BYTES = NARR * MFIT_TILE * 8u;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char mfit_smem[];
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5, nwarp = blockDim.x >> 5;
  const int S = ma.stages;
  double* buf = reinterpret_cast<double*>(mfit_smem + (size_t)wib * S * STAGE_BYTES);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(mfit_smem + (size_t)nwarp * S * STAGE_BYTES) + wib * S;
  volatile int* q_items = reinterpret_cast<volatile int*>(mfit_smem + (size_t)nwarp * S * (STAGE_BYTES + 8)) + wib * MFIT_QN;

  const int total = ma.item_start[ma.n_ds];
  // item id -> data set and its (segment, warp-block) item
  auto ds_of = [&](int item) -> int {
    if (!MULTI) return 0;
    int k = 0;
    while (k + 1 < ma.n_ds && item >= ma.item_start[k + 1]) k++;
    return k;
  };
  auto start_of = [&](int k) -> int { return MULTI ? ma.item_start[k] : 0; };
  auto tiles_of = [&](const MfitEvalArgs& a, int seg) -> int {
    const long long n_tiles = (a.n + MFIT_TILE - 1) / MFIT_TILE;
    long long t0 = (long long)seg * a.tiles_per_seg, t1 = t0 + a.tiles_per_seg;
    if (t1 > n_tiles) t1 = n_tiles;
    return (int)(t1 - t0);
  };

  mfit_math_stage(Model::USES | MfitTerm<LL>::USES, tid, (int)blockDim.x);  // exp/log tables -> shared memory
  if (lane == 0) {
    for (int s = 0; s < S; s++) mfit_mbar_init(&full[s], 1);
    mfit_fence_mbar_init();
  }
  __syncthreads();  // tables staged (the only CTA-wide barrier before the epilogue)

  // ---- producer side of the warp (all values warp-uniform; lane 0 talks to the TMA engine) ----
  // cursor: next tile to request = tile p_k of item p_item (segment p_seg of data set p_ds); p_raw = the item id drawn
  // for after it (lane 0 only until it is broadcast), p_seq = items published so far
  int p_item = -1, p_k = 0, p_nt = 0, p_seq = 0, p_ds = 0, p_seg = 0, p_raw = 0;
  bool p_done = false;
  int p_stage = 0;          // ring slot the next request goes to
  long long p_off = 0;      // its first point (index into the data arrays)
  auto draw = [&]() {  // ask for the next item id; the answer is first used one tile later (next_item)
    if (lane == 0) p_raw = atomicAdd(ma.sched, 1);
  };
  auto next_item = [&]() {  // move the cursor to the item drawn last and publish its id to the consumer side
    p_item = __shfl_sync(FULL, p_raw, 0);
    if (p_item >= total) {
      p_item = -1;
      p_done = true;
    } else {
      p_ds = ds_of(p_item);
      const MfitEvalArgs& a = ma.d[MULTI ? p_ds : 0];
      p_seg = (p_item - start_of(p_ds)) / (a.n_cb * (MFIT_CB / MFIT_WB));
      p_k = 0;
      p_nt = tiles_of(a, p_seg);
      p_off = (long long)p_seg * a.tiles_per_seg * (long long)MFIT_TILE;
    }
    if (lane == 0) q_items[p_seq % MFIT_QN] = p_item;
    p_seq++;
    __syncwarp();
  };
  auto issue_next = [&]() {  // request one tile (if any is left), one stage ahead of the ring's oldest free slot
    if (p_done) return;
    if (p_k == p_nt) {
      next_item();
      if (p_done) return;
    }
    if (lane == 0) {
      const MfitEvalArgs& a = ma.d[MULTI ? p_ds : 0];
      const int s = p_stage;
      const long long off = p_off;
      double* dst = buf + (size_t)s * NARR * MFIT_TILE;
      mfit_mbar_expect_tx(&full[s], STAGE_BYTES);
#pragma unroll
      for (int d = 0; d < XD; d++) mfit_tma_load_1d(dst + d * MFIT_TILE, a.x + d * a.xd_stride + off, MFIT_TILE * 8u, &full[s]);
      mfit_tma_load_1d(dst + XD * MFIT_TILE, (LL == MFITK_NORMAL ? a.yw : a.y) + off, MFIT_TILE * 8u, &full[s]);
      if (LNARR > 2) mfit_tma_load_1d(dst + (XD + 1) * MFIT_TILE, a.w + off, MFIT_TILE * 8u, &full[s]);
      if (LNARR > 3) mfit_tma_load_1d(dst + (XD + 2) * MFIT_TILE, a.c + off, MFIT_TILE * 8u, &full[s]);
    }
    p_off += MFIT_TILE;
    if (++p_stage == S) p_stage = 0;
    if (++p_k == p_nt) draw();
  };
  draw();
  next_item();
  for (int s = 0; s + 1 < S; s++) issue_next();  // S - 1 tiles ahead of the one being evaluated

  // ---- consumer side ----
  Model m_next[2];
  bool live_next[2] = {false, false};
  auto load_item = [&](int it) {  // parameters and finished-flags of the warp's chains for item `it`
    const int k = ds_of(it);
    const MfitEvalArgs& a = ma.d[MULTI ? k : 0];
    const long long c0 = ((long long)a.cb0 * (MFIT_CB / MFIT_WB) + (it - start_of(k)) % (a.n_cb * (MFIT_CB / MFIT_WB))) * MFIT_WB + lane;
    const long long last = (long long)a.W - 1;
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const long long ch = c0 + 32 * c;
      const long long cl = ch < a.W ? ch : last;
      m_next[c].load(a, cl);
      live_next[c] = ch < a.W && !(a.skip != nullptr && a.skip[cl] != 0);
    }
  };
  bool have_next = false;
  int c_stage = 0;          // ring slot of the tile evaluated next, and the parity its mbarrier completes with
  unsigned c_phase = 0;
  for (int seq = 0;; seq++) {
    const int item = q_items[seq % MFIT_QN];  // published by this warp's own producer side before the item's first tile was requested
    if (item < 0) break;
    const int ds = ds_of(item);
    const MfitEvalArgs& a = ma.d[MULTI ? ds : 0];
    const int local = item - start_of(ds);
    const int n_wb = a.n_cb * (MFIT_CB / MFIT_WB);
    const int seg = local / n_wb;
    const long long cb0 = ((long long)a.cb0 * (MFIT_CB / MFIT_WB) + local % n_wb) * MFIT_WB;
    const long long chain[2] = {cb0 + lane, cb0 + 32 + lane};
    if (!have_next) load_item(item);
    Model m[2] = {m_next[0], m_next[1]};
    const bool live[2] = {live_next[0], live_next[1]};
    const bool any_live = __any_sync(FULL, (live[0] || live[1]) ? 1 : 0) != 0;  // a finished warp-block only drains its tiles
    const bool two = cb0 + 32 < a.W;                                           // warp-uniform: the second half holds chains
    m[0].setup();  // theta-only subexpressions
    m[1].setup();
    // warp-uniform: every chain of this warp stays inside the fast domain over the whole x range of the data set
    // (no guard at all), or at least has finite parameters where the model's fast routines clamp in place
    int mode = MFIT_MODE_GUARD;
    // a term that takes log f (log-poisson) is as safe as the model when the model also proves f positive and normal
    bool term_safe = MfitTerm<LL>::USES == 0;
    if constexpr (LL == MFITK_POISSON)
      term_safe = __all_sync(FULL, (m[0].positive(a) && (!two || m[1].positive(a))) ? 1 : 0) != 0;
    if (term_safe) {
      if (__all_sync(FULL, (m[0].safe(a) && (!two || m[1].safe(a))) ? 1 : 0) != 0) {
        mode = MFIT_MODE_FREE;
#ifndef MFIT_NO_FUSE   /* A/B aid: tools/build_variant.sh nofuse -DMFIT_NO_FUSE */
        if constexpr (Model::FUSABLE) {
          if (__all_sync(FULL, (m[0].fusable(a) && (!two || m[1].fusable(a))) ? 1 : 0) != 0) mode = MFIT_MODE_FUSED;
        }
#endif
      } else if constexpr (Model::CLAMP_ONLY) {
        if (__all_sync(FULL, (m[0].tame(a) && (!two || m[1].tame(a))) ? 1 : 0) != 0) mode = MFIT_MODE_CLAMP;
      }
    }
    const int nt = tiles_of(a, seg);
    if (ma.mode_tiles != nullptr && lane == 0 && any_live) atomicAdd(ma.mode_tiles + mode, (unsigned long long)nt);
    long long rem_ll = a.n - (long long)seg * a.tiles_per_seg * (long long)MFIT_TILE;  // points from the segment's first one on
    int rem = rem_ll < (long long)nt * MFIT_TILE ? (int)rem_ll : nt * MFIT_TILE;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    have_next = false;
    for (int k = 0; k < nt; k++) {
      const int s = c_stage;
      issue_next();  // into the stage evaluated one tile ago (every lane has passed the __syncwarp that ended it)
      if (!have_next && p_seq > seq + 1) {
        // the next item is known: its parameters travel while this item is evaluated
        const int nxt = q_items[(seq + 1) % MFIT_QN];
        if (nxt >= 0) { load_item(nxt); have_next = true; }
      }
      mfit_mbar_wait(&full[s], c_phase);
      if (++c_stage == S) { c_stage = 0; c_phase ^= 1u; }
      const int cnt = rem < MFIT_TILE ? rem : MFIT_TILE;
      rem -= MFIT_TILE;
      if (any_live) {
        const double* xs = buf + (size_t)s * NARR * MFIT_TILE;
        const double* ys = xs + XD * MFIT_TILE;
        const double* ws = xs + (LNARR > 2 ? XD + 1 : 0) * MFIT_TILE;
        const double* cs = xs + (LNARR > 3 ? XD + 2 : 0) * MFIT_TILE;
        if (two) mfit_tile<Model, LL, 2>(m, mode, xs, ys, ws, cs, cnt, acc);
        else mfit_tile<Model, LL, 1>(m, mode, xs, ys, ws, cs, cnt, acc);
      }
      __syncwarp();  // every lane is done with stage s
    }
    if (live[0]) a.partial[(size_t)seg * a.stride + chain[0]] = acc[0] + acc[1];
    if (live[1]) a.partial[(size_t)seg * a.stride + chain[1]] = acc[2] + acc[3];
  }
  __syncthreads();
  // the last CTA out re-arms the scheduler for the next launch on this stream
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(ma.sched + 1, 1) == (int)gridDim.x - 1) {
      ma.sched[0] = 0;
      ma.sched[1] = 0;
      __threadfence();
    }
  }
}

template <class Model, int LL>
__device__ __forceinline__ void mfit_eval_body(const MfitMultiArgs& ma) {
  if (ma.n_ds == 1) mfit_eval_body_t<Model, LL, false>(ma);
  else mfit_eval_body_t<Model, LL, true>(ma);
}

// ---------------------------------------------------------------- walker sets with per-walker data
// The reference's own many-walker workload (nv-specific.lisp:50-66): one walker per measured spectrum, a few hundred
// points each, same :function / priors, `(mapc #'walker-adaptive-steps walkers)`.  Nothing is shared between chains
// here, so the mapping flips: one WARP per chain, the lanes stride over that chain's points (coalesced 256-byte
// reads straight from global memory -- every byte is used once per step, there is nothing to stage), theta is
// warp-uniform, four independent accumulators per lane, and the 128 partial sums are combined in a fixed order
// (accumulators, then a shuffle tree).  This kernel is bound by L2 / HBM bandwidth (24 bytes per point-eval), not
// by the FP64 pipe.  Persistent warps stride over the chains of the launch.
#ifndef MFIT_PC_U
#define MFIT_PC_U 8 /* strides of 32 points whose loads are issued together (memory-level parallelism per warp) */
#endif
#ifndef MFIT_PC_STAGED_CTAS
#define MFIT_PC_STAGED_CTAS 2 /* min CTAs per SM the staged per-walker kernel is register-bounded for (3 or 4: 80 / 64 registers with spills, 20-35 % slower; 1: the compiler takes 200+ registers, one CTA per SM, 60 % slower) */
#endif
#define MFIT_PC_STAGE_MAX 2048 /* points up to which shared x / sigma vectors are staged in shared memory */
#define MFIT_PCR_THREADS 128   /* threads per CTA of the register form of the staged kernel (mfit_reg_row): four warps, one per scheduler */
#ifndef MFIT_PCR_CTAS
#define MFIT_PCR_CTAS 3        /* its CTAs per SM (register bound: 168).  4 (128 registers, a two-row ring) was measured: kernel
                                  0.125 against 0.118 ms, MH step of the set 0.155 against 0.158 ms -- occupancy is not what it lacks */
#endif
#define MFIT_PCR_BATCH 8       /* chains whose 32 per-lane sums wait in shared memory and are combined together (register form) */
// STAGED: x and 1/sigma (and the cutoff constants) are one vector for all chains and sit in shared memory (xs, ws, cs
// point there), so only the chain's own y values are streamed.
template <class Model, int LL, bool SLOW, bool STAGED, class G>
__device__ __forceinline__ double mfit_perchain_pass(const Model& m, const double* xs, const double* ys, const double* ws,
                                                     const double* cs, long long n, int lane, G& gd) {
  constexpr int NARR = MfitTerm<LL>::NARR;
  constexpr int U = MFIT_PC_U;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (long long i0 = 0; i0 < n; i0 += 32 * U) {
    // all global loads of this block first: a warp issues in order, so loads left inside the arithmetic would each
    // expose a full memory latency (the kernel ran at 1.9 TB/s that way)
    double xv[STAGED ? 1 : U], yv[U], wv[STAGED ? 1 : U], cv[STAGED ? 1 : U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long j = i0 + 32 * u + lane;
      const bool in = j < n;
      yv[u] = in ? ys[j] : 0.0;
      if (!STAGED) {
        xv[u] = in ? xs[j] : 0.0;
        wv[u] = (NARR > 2 && in) ? ws[j] : 0.0;
        cv[u] = (NARR > 3 && in) ? cs[j] : 0.0;
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long j = i0 + 32 * u + lane;
      if (j < n) {
        const double x = STAGED ? xs[j] : xv[u];
        const double w = STAGED ? (NARR > 2 ? ws[j] : 0.0) : wv[u];
        const double c = STAGED ? (NARR > 3 ? cs[j] : 0.0) : cv[u];
        const double f = mfit_model_scalar<Model, SLOW>(m, x, gd);
        acc[u & 3] = MfitTerm<LL>::template add<SLOW>(acc[u & 3], f, yv[u], w, c, gd);
      }
    }
  }
  double sum = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum = sum + __shfl_down_sync(0xffffffffu, sum, off);
  return sum;  // valid in lane 0
}

// one chain's row out of shared memory: lane sums acc[0..3].  Whole groups of four rows of 32 points run without a
// bounds test and with 32-bit indices (the register path spends 51 warp-instructions per point on 16 FP64 ones, ncu);
// row r adds into accumulator r mod 4, the order of the register path's blocks of eight.  One instance per guard type,
// so there is no per-point branch on it.  (A function template, not a generic lambda: NVRTC takes the latter's call
// operator for a host function.)
template <class Model, int LL, class G>
__device__ __forceinline__ void mfit_ring_row(const Model& m, const double* xs, const double* ws, const double* cs,
                                              const double* row, int n, int lane, double* acc, G& g) {
  constexpr int NARR = MfitTerm<LL>::NARR;
#define MFIT_RING_TERM(u, j)                                                                                      \
  acc[u] = MfitTerm<LL>::template add<false>(acc[u], mfit_model_scalar<Model, false>(m, xs[j], g), row[j],          \
                                             NARR > 2 ? ws[j] : 0.0, NARR > 3 ? cs[j] : 0.0, g)
  int j0 = lane;
#pragma unroll 2
  for (; j0 + 96 + (31 - lane) < n; j0 += 128) {  // warp-uniform: the group's last point exists
    MFIT_RING_TERM(0, j0);
    MFIT_RING_TERM(1, j0 + 32);
    MFIT_RING_TERM(2, j0 + 64);
    MFIT_RING_TERM(3, j0 + 96);
  }
  if (j0 < n) MFIT_RING_TERM(0, j0);
  if (j0 + 32 < n) MFIT_RING_TERM(1, j0 + 32);
  if (j0 + 64 < n) MFIT_RING_TERM(2, j0 + 64);
  if (j0 + 96 < n) MFIT_RING_TERM(3, j0 + 96);
#undef MFIT_RING_TERM
}

// The same row with the chain-invariant operands in REGISTERS: lane l of a warp always evaluates points l, l + 32, ...,
// l + 32 (NPT - 1) of whatever chain the warp is on, so their x are loaded once per warp (xr) instead of once per chain,
// 1/sigma is ONE number (the host takes this form only when the set's sigma is uniform -- the reference's NV use passes a
// number, nv-specific.lisp -- which spares sixteen registers), and only y comes out of the ring: one LDS per evaluation instead of three (the shared-memory version
// issued 46 warp-instructions per warp-point for 16 FP64 ones and ran at the pace of the issue ports, ncu
// profiles/r02_perchain_kernel_ncu_raw.csv).  Fully unrolled and branch-free: NPT independent evaluations per lane;
// slots beyond the row are evaluated at a real abscissa and dropped by a select.  Same
// accumulator per row and same arithmetic as mfit_ring_row: the sums are bit-identical.
// LASTONLY: the row is longer than 32 (NPT - 1) points, i.e. only the last slot can lie beyond it and only that one is
// selected (a compare and two selects per slot are 8 % of the per-chain instructions otherwise).
template <class Model, int LL, int NPT, bool LASTONLY, class G>
__device__ __forceinline__ void mfit_reg_row(const Model& m, const double* xr, double w0, int n_lane,
                                             const double* row, int lane, double* acc, G& g) {
#pragma unroll
  for (int u = 0; u < NPT; u++) {
    const double y = row[lane + 32 * u];
    const double f = mfit_model_scalar<Model, false>(m, xr[u], g);
    const double t = MfitTerm<LL>::template add<false>(acc[u & 3], f, y, w0, 0.0, g);
    if (LASTONLY && u < NPT - 1) acc[u & 3] = t;
    else acc[u & 3] = (32 * u < n_lane) ? t : acc[u & 3];   // n_lane = n - lane: point lane + 32 u exists (one compare with an immediate)
  }
}

// STAGED with a ring (a.ring_stages = S >= 2): the chain's own y row -- the only per-chain data here -- is fetched by the
// TMA engine (cp.async.bulk, one request per row) into a per-warp ring of S rows in shared memory, S - 1 rows ahead of the
// one being evaluated.  A warp that reads its rows straight into registers alternates between waiting for memory and
// computing, and 16 warps x 2 KB in flight per SM is about half of what HBM3e needs to stay busy (6.5 TB/s x ~0.8 us
// = 35 KB per SM *on average*): the kernel ran at 0.43 of the copy peak.  With the ring the bytes in flight no longer
// depend on what the warps are doing: 16 warps x (S - 1) rows of n_pad doubles per SM.
// Shared memory: [x][1/sigma][c] (3 n_pad doubles, as for the register path), then [warp][stage][n_pad] doubles, then
// [warp][stage] mbarriers.  Sums are formed exactly as in the register path (same four accumulators per lane, same
// shuffle tree), so the two are bit-identical.
// NPT > 0: the register form of the row (mfit_reg_row; rows of up to 32 NPT points, likelihoods without the cutoff
// constants); ring slots are then 32 NPT doubles apart so that the unrolled reads stay inside the ring.
template <class Model, int LL, int NPT = 0>
__device__ __forceinline__ void mfit_perchain_ring(const MfitEvalArgs& a, double* smem, long long n_pad) {
  constexpr bool GUARDED = Model::GUARDED || MfitTerm<LL>::USES != 0;
  const int S = a.ring_stages;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const long long slot_len = NPT > 0 ? 32 * NPT : n_pad;   // doubles per ring slot
  double* ring = smem + 3 * n_pad + (size_t)wib * S * slot_len;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + 3 * n_pad + (size_t)nwarp * S * slot_len) + wib * S;
  const double* xs = smem;
  const double* ws = smem + n_pad;
  const double* cs = smem + 2 * n_pad;
  // Register form: the 32 per-lane sums of a chain are not combined by a shuffle tree per chain (five dependent
  // shuffle + add levels, ~200 cycles in which the warp issues nothing else: 29 % of the kernel's samples, ncu) but parked
  // in shared memory, [chain of the batch][lane] with rows 33 apart; after MFIT_PCR_BATCH chains lane b adds up chain b's
  // row IN THE TREE'S ORDER (s_i + s_{i+16}, then + 8, 4, 2, 1), so the sums are the same bits.
  double* red = reinterpret_cast<double*>(reinterpret_cast<unsigned long long*>(smem + 3 * n_pad + (size_t)nwarp * S * slot_len) + (size_t)nwarp * S) +
                (size_t)wib * MFIT_PCR_BATCH * 33;
  int batch_n = 0;               // chains parked
  unsigned batch_live = 0u;      // bit b: chain b of the batch has a sum to write
  long long batch_first = 0;     // its first chain; chain b = batch_first + b * warps
  const long long warps = (long long)gridDim.x * nwarp;
  auto flush = [&]() {
    __syncwarp();
    if (lane < batch_n && ((batch_live >> lane) & 1u)) {
      const double* v = red + lane * 33;
      double t16[16];
#pragma unroll
      for (int i = 0; i < 16; i++) t16[i] = v[i] + v[i + 16];
#pragma unroll
      for (int i = 0; i < 8; i++) t16[i] = t16[i] + t16[i + 8];
#pragma unroll
      for (int i = 0; i < 4; i++) t16[i] = t16[i] + t16[i + 4];
      t16[0] = t16[0] + t16[2];
      t16[1] = t16[1] + t16[3];
      a.partial[batch_first + (long long)lane * warps] = t16[0] + t16[1];
    }
    __syncwarp();
    batch_n = 0;
    batch_live = 0u;
  };
  const long long w_lo = (long long)a.cb0 * MFIT_CB;
  long long w_hi = w_lo + (long long)a.n_cb * MFIT_CB;
  if (w_hi > a.W) w_hi = a.W;
  const long long first = w_lo + (long long)blockIdx.x * nwarp + wib;
  const double* ybase = LL == MFITK_NORMAL ? a.yw : a.y;
  const unsigned row_bytes = (unsigned)(n_pad * 8);
  if (lane == 0 && a.diag != 2)
    for (int s = 0; s < S; s++) {
      const long long ch = first + (long long)s * warps;
      if (ch < w_hi) {
        mfit_mbar_expect_tx(&bars[s], row_bytes);
        mfit_tma_load_1d(ring + (size_t)s * slot_len, ybase + ch * a.pc_stride, row_bytes, &bars[s]);
      }
    }
  const int n = (int)a.n;
  double xr[NPT > 0 ? NPT : 1];
  const double w0 = (NPT > 0 && MfitTerm<LL>::NARR > 2) ? ws[0] : 0.0;   // register form: sigma is uniform over the set
  const int n_lane = n - lane;
  if constexpr (NPT > 0) {
#pragma unroll
    for (int u = 0; u < NPT; u++) {
      const int j = lane + 32 * u;
      const bool ok = j < n;
      const int jj = ok ? j : lane % n;   // a real abscissa for the slots beyond the row (their terms are dropped)
      xr[u] = xs[jj];
    }
  }
  // The chain's parameters are theta[idx[k] * stride + chain]: with the products idx[k] * stride formed once (the host takes
  // this kernel only when they fit 31 bits) and the chain folded into the base pointer, a load costs one address instruction
  // per parameter instead of a 64-bit multiply-add each (40 of the ~500 instructions a warp issued per chain).
  MfitEvalArgs al = a;
#pragma unroll
  for (int k = 0; k < MFIT_MAXK; k++) al.idx[k] = (int)((long long)a.idx[k] * a.stride);
  al.stride = 1;
  Model m_next;
  bool skip_next = false;
  if (first < w_hi) {
    al.theta = a.theta + first;
    m_next.load(al, 0);
    skip_next = a.skip != nullptr && a.skip[first] != 0;
  }
  const long long row_step = warps * a.pc_stride;   // doubles between the rows of consecutive chains of this warp
  const double* row_next = ybase + (first + (long long)S * warps) * a.pc_stride;   // the row the next free slot takes
  int slot = 0;            // ring slot of this chain's row and the parity its mbarrier completes with (no division per chain)
  unsigned phase = 0u;
  for (long long chain = first; chain < w_hi; chain += warps) {
    Model m = m_next;
    const bool skip = skip_next;
    if (chain + warps < w_hi) {  // the next chain's parameters and its finished flag, one chain ahead
      al.theta = a.theta + (chain + warps);
      m_next.load(al, 0);
      skip_next = a.skip != nullptr && a.skip[chain + warps] != 0;
    }
    m.setup();
    const bool unguarded = !GUARDED || ((MfitTerm<LL>::USES == 0 || (LL == MFITK_POISSON && m.positive(a))) && m.safe(a));  // theta is warp-uniform
    if (a.diag != 2) mfit_mbar_wait(&bars[slot], phase);
    const double* row = ring + (size_t)slot * slot_len;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    double sum = 0.0;
    bool redo = false;
    if (!skip && a.diag != 1) {
      MfitGuard gd;
      gd.reset();
      MfitNoGuard ng;
      if constexpr (NPT > 0) {
        if (n > 32 * (NPT - 1)) {   // kernel-uniform
          if (unguarded) mfit_reg_row<Model, LL, NPT, true>(m, xr, w0, n_lane, row, lane, acc, ng);
          else mfit_reg_row<Model, LL, NPT, true>(m, xr, w0, n_lane, row, lane, acc, gd);
        } else {
          if (unguarded) mfit_reg_row<Model, LL, NPT, false>(m, xr, w0, n_lane, row, lane, acc, ng);
          else mfit_reg_row<Model, LL, NPT, false>(m, xr, w0, n_lane, row, lane, acc, gd);
        }
      } else {
        if (unguarded) mfit_ring_row<Model, LL>(m, xs, ws, cs, row, n, lane, acc, ng);
        else mfit_ring_row<Model, LL>(m, xs, ws, cs, row, n, lane, acc, gd);
      }
      redo = !unguarded && __any_sync(0xffffffffu, gd.tripped() ? 1 : 0) != 0;
      if (redo) sum = mfit_perchain_pass<Model, LL, true, true>(m, xs, row, ws, cs, a.n, lane, gd);  // rare: libdevice
    }
    __syncwarp();  // every lane is done with the row: its slot goes back to the TMA engine before the sums are combined
    if (lane == 0 && a.diag != 2) {
      const long long nxt = chain + (long long)S * warps;
      if (nxt < w_hi) {
        mfit_mbar_expect_tx(&bars[slot], row_bytes);
        mfit_tma_load_1d(ring + (size_t)slot * slot_len, row_next, row_bytes, &bars[slot]);
      }
    }
    row_next += row_step;
    if constexpr (NPT > 0) {
      if (batch_n == 0) batch_first = chain;
      if (!skip && !redo) {
        red[batch_n * 33 + lane] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        batch_live |= 1u << batch_n;
      } else if (!skip && lane == 0) {
        a.partial[chain] = sum;  // the libdevice pass has combined its lanes already
      }
      if (++batch_n == MFIT_PCR_BATCH) flush();
    } else if (!skip) {
      if (!redo) {
        sum = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum = sum + __shfl_down_sync(0xffffffffu, sum, off);
      }
      if (lane == 0) a.partial[chain] = sum;  // one segment
    }
    if (++slot == S) { slot = 0; phase ^= 1u; }
  }
  if constexpr (NPT > 0) {
    if (batch_n > 0) flush();
  }
}

template <class Model, int LL, bool STAGED, int NPT = 0>
__device__ __forceinline__ void mfit_eval_perchain_body_t(const MfitEvalArgs& a) {
  constexpr bool GUARDED = Model::GUARDED || MfitTerm<LL>::USES != 0;
  constexpr int NARR = MfitTerm<LL>::NARR;
  extern __shared__ __align__(16) double mfit_pc_smem[];
  mfit_math_stage(Model::USES | MfitTerm<LL>::USES, threadIdx.x, blockDim.x);
  const long long n_pad = (a.n + 31) & ~31LL;
  if (STAGED)
    for (long long i = threadIdx.x; i < a.n; i += blockDim.x) {
      mfit_pc_smem[i] = a.x[i];
      if (NARR > 2) mfit_pc_smem[n_pad + i] = a.w[i];
      if (NARR > 3) mfit_pc_smem[2 * n_pad + i] = a.c[i];
    }
  if (STAGED && a.ring_stages > 0) {
    if ((threadIdx.x & 31) == 0) {  // every warp owns the mbarriers of its own ring
      unsigned long long* bars = reinterpret_cast<unsigned long long*>(
                                     mfit_pc_smem + 3 * n_pad + (size_t)(blockDim.x >> 5) * a.ring_stages * (NPT > 0 ? 32 * NPT : n_pad)) +
                                 (threadIdx.x >> 5) * a.ring_stages;
      for (int s_ = 0; s_ < a.ring_stages; s_++) mfit_mbar_init(&bars[s_], 1);
      mfit_fence_mbar_init();
    }
    __syncthreads();
    mfit_perchain_ring<Model, LL, NPT>(a, mfit_pc_smem, n_pad);
    return;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long w_lo = (long long)a.cb0 * MFIT_CB;
  long long w_hi = w_lo + (long long)a.n_cb * MFIT_CB;
  if (w_hi > a.W) w_hi = a.W;
  long long chain = w_lo + (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  Model m_next;
  if (chain < w_hi) m_next.load(a, chain);
  const double* ybase = LL == MFITK_NORMAL ? a.yw : a.y;
  double ycur[MFIT_PC_U];  // STAGED: the block of y values about to be evaluated
  auto prime = [&](long long ch) {
    const double* yn = ybase + ch * a.pc_stride;
#pragma unroll
    for (int u = 0; u < MFIT_PC_U; u++) {
      const long long j = 32 * u + lane;
      ycur[u] = (STAGED && ch < w_hi && j < a.n) ? yn[j] : 0.0;
    }
  };
  if (STAGED) prime(chain);
  for (; chain < w_hi; chain += warps) {
    Model m = m_next;
    if (chain + warps < w_hi) m_next.load(a, chain + warps);  // the next chain's parameters, one chain ahead
    if (a.skip != nullptr && a.skip[chain] != 0) {
      if (STAGED) prime(chain + warps);
      continue;
    }
    m.setup();
    const double* xs = STAGED ? mfit_pc_smem : a.x + chain * a.x_stride;
    const double* ys = (LL == MFITK_NORMAL ? a.yw : a.y) + chain * a.pc_stride;
    const double* ws = STAGED ? mfit_pc_smem + n_pad : a.w + chain * a.w_stride;
    const double* cs = STAGED ? mfit_pc_smem + 2 * n_pad : a.c + chain * a.w_stride;
    double sum;
    // theta is warp-uniform here, so is the model's verdict on its arguments
    const bool unguarded = !GUARDED || ((MfitTerm<LL>::USES == 0 || (LL == MFITK_POISSON && m.positive(a))) && m.safe(a));
    if (STAGED) {
      // Software pipeline over 256-point blocks, across chain boundaries: the y values of the NEXT block (of this chain,
      // or the first block of this warp's next chain) are requested before the current block is evaluated, so a warp
      // always has a block in flight instead of alternating between waiting for memory and computing.
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      MfitGuard gd;
      gd.reset();
      MfitNoGuard ng;
      for (long long i0 = 0; i0 < a.n; i0 += 32 * MFIT_PC_U) {
        double ynext[MFIT_PC_U];
        {
          long long nch = chain, ni0 = i0 + 32 * MFIT_PC_U;
          if (ni0 >= a.n) { nch = chain + warps; ni0 = 0; }
          const double* yn = ybase + nch * a.pc_stride;
#pragma unroll
          for (int u = 0; u < MFIT_PC_U; u++) {
            const long long j = ni0 + 32 * u + lane;
            ynext[u] = (nch < w_hi && j < a.n) ? yn[j] : 0.0;
          }
        }
#pragma unroll
        for (int u = 0; u < MFIT_PC_U; u++) {
          const long long j = i0 + 32 * u + lane;
          if (j < a.n) {
            const double w = NARR > 2 ? ws[j] : 0.0, c = NARR > 3 ? cs[j] : 0.0;
            if (unguarded) acc[u & 3] = MfitTerm<LL>::template add<false>(acc[u & 3], mfit_model_scalar<Model, false>(m, xs[j], ng), ycur[u], w, c, ng);
            else acc[u & 3] = MfitTerm<LL>::template add<false>(acc[u & 3], mfit_model_scalar<Model, false>(m, xs[j], gd), ycur[u], w, c, gd);
          }
        }
#pragma unroll
        for (int u = 0; u < MFIT_PC_U; u++) ycur[u] = ynext[u];
      }
      sum = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) sum = sum + __shfl_down_sync(0xffffffffu, sum, off);
      if (!unguarded && __any_sync(0xffffffffu, gd.tripped() ? 1 : 0))  // rare: this chain again with libdevice
        sum = mfit_perchain_pass<Model, LL, true, true>(m, xs, ys, ws, cs, a.n, lane, gd);
    } else if (unguarded) {
      MfitNoGuard ng;
      sum = mfit_perchain_pass<Model, LL, false, false>(m, xs, ys, ws, cs, a.n, lane, ng);
    } else {
      MfitGuard gd;
      gd.reset();
      sum = mfit_perchain_pass<Model, LL, false, false>(m, xs, ys, ws, cs, a.n, lane, gd);
      if (__any_sync(0xffffffffu, gd.tripped() ? 1 : 0))
        sum = mfit_perchain_pass<Model, LL, true, false>(m, xs, ys, ws, cs, a.n, lane, gd);
    }
    if (lane == 0) a.partial[chain] = sum;  // one segment
  }
}

// the host picks STAGED (separate kernels, so that the staged one keeps its lower register count):
// a.x_stride == 0 && a.w_stride == 0 && a.n <= MFIT_PC_STAGE_MAX

// ---------------------------------------------------------------- fit curve and max / min envelope over the trace
// walker-get-data-and-fit (mcmc-fitting.lisp:1231-1255): y-fit = (apply fn x params) over x-fit, and for every x the
// largest and smallest model value over the selected steps of the walk (all-ys / max-ys / min-ys :1250-1252).  The
// chain kernels have gathered each chain's distinct selected steps as parameter columns (k_envelope_select), so a
// "chain" of Model::load is a column here.  One CTA row per chain, MFIT_FIT_PTS x-fit points per thread held in
// registers, the columns are CTA-uniform (broadcast loads, the next one requested before the current one is
// evaluated).  libdevice math throughout (SLOW forms): this runs once per plot, not once per step.
#define MFIT_FIT_THREADS 128
#define MFIT_FIT_PTS 4
struct MfitFitArgs {
  MfitEvalArgs a;        /* the data set's model slots (idx, degree, xmin, xmax); theta = gathered columns
                            [chains of this launch][P][cols], stride = cols */
  long long chain_pitch; /* P * cols */
  long long cols;        /* columns per chain; the last one holds the parameters of the fit curve */
  const int* n_sel;      /* [chains of this launch] selected distinct steps */
  const double* xfit;    /* [n_fit] */
  int n_fit;
  double* ymax;          /* [chains of this launch][n_fit] */
  double* ymin;
  double* yfit;
};

template <class Model>
__device__ __forceinline__ void mfit_fit_envelope_body(const MfitFitArgs& f) {
  const long long wl = blockIdx.x;
  MfitEvalArgs a = f.a;
  a.theta = f.a.theta + (size_t)wl * f.chain_pitch;
  a.stride = f.cols;
  a.W = (int)f.cols;
  const int n = f.n_sel[wl];
  const int i0 = (int)blockIdx.y * (MFIT_FIT_THREADS * MFIT_FIT_PTS) + (int)threadIdx.x;
  double x[MFIT_FIT_PTS], mx[MFIT_FIT_PTS], mn[MFIT_FIT_PTS];
#pragma unroll
  for (int u = 0; u < MFIT_FIT_PTS; u++) {
    const int i = i0 + u * MFIT_FIT_THREADS;
    x[u] = i < f.n_fit ? f.xfit[i] : 0.0;
    mx[u] = mn[u] = __longlong_as_double(0x7ff8000000000000LL);  // no step selected: NaN
  }
  MfitNoGuard ng;
  Model m_next;
  m_next.load(a, n > 0 ? 0 : f.cols - 1);
  for (int j = 0; j < n; j++) {
    Model m = m_next;
    m_next.load(a, j + 1 < n ? j + 1 : f.cols - 1);
    m.setup();
#pragma unroll
    for (int u = 0; u < MFIT_FIT_PTS; u++) {
      const double v = mfit_model_scalar<Model, true>(m, x[u], ng);
      if (j == 0) { mx[u] = v; mn[u] = v; }       // (reduce #'max y) starts from the first element
      else { mx[u] = v > mx[u] ? v : mx[u]; mn[u] = v < mn[u] ? v : mn[u]; }
    }
  }
  m_next.setup();  // the fit's own parameters
#pragma unroll
  for (int u = 0; u < MFIT_FIT_PTS; u++) {
    const int i = i0 + u * MFIT_FIT_THREADS;
    if (i < f.n_fit) {
      const size_t o = (size_t)wl * f.n_fit + i;
      f.yfit[o] = mfit_model_scalar<Model, true>(m_next, x[u], ng);
      f.ymax[o] = mx[u];
      f.ymin[o] = mn[u];
    }
  }
}

#endif  // MFIT_EVAL_CUH
