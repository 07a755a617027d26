// chain_kernels.cu -- per-chain side of the Metropolis-Hastings step, one thread per chain.
//
// Device form of hot loop #1 of the reference (mcmc-fitting.lisp:902-942): proposal theta' = theta + L.z
// (get-covariant-sample :679-700), the accept test of walker-take-step (:1091-1094), walker-modify
// :add-step (:549-555), the temperature schedule, auto-stop and acceptance-window l-matrix update of
// walker-adaptive-steps-full (:905-942) and the forward-step covariance + Cholesky re-estimate
// (:497-502, :263-280, :614-643, :583-598, :888-895).
//
// This translation unit is compiled with -fmad=false and keeps the reference's operation order
// (sequential sums, separate multiply and add) so that, given the same injected normals/uniforms and the
// same log-probabilities, proposals, l-matrices and decisions are bit-identical to the CPU oracle.
#include <cuda_runtime.h>
#include <math.h>
#include "mfit_chain.h"
#include "mfit_internal.h"

namespace {

// ------------------------------------------------------------------ Philox4x32-10 (counter based)
struct U4 { unsigned x, y, z, w; };
__device__ __forceinline__ U4 philox4x32_10(U4 c, unsigned k0, unsigned k1) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    U4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += W0;
    k1 += W1;
  }
  return c;
}
__device__ __forceinline__ double u01(unsigned hi, unsigned lo) {  // (0,1), 53 bits
  unsigned long long v = ((unsigned long long)hi << 32) | lo;
  return ((double)(v >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
// stream layout: counter = (step_lo, step_hi, chain_lo, chain_hi16 | lane << 16); lane 0xFFFF = accept uniform,
// lane m < 0xFFFF = normals 2m, 2m+1 (Box-Muller)
__device__ __forceinline__ void philox_pair(unsigned long long seed, long long gchain, long long step, unsigned lane,
                                            double* a, double* b) {
  U4 c;
  c.x = (unsigned)step;
  c.y = (unsigned)((unsigned long long)step >> 32);
  c.z = (unsigned)gchain;
  c.w = ((unsigned)((unsigned long long)gchain >> 32) & 0xFFFFu) | (lane << 16);
  U4 r = philox4x32_10(c, (unsigned)seed, (unsigned)(seed >> 32));
  *a = u01(r.x, r.y);
  *b = u01(r.z, r.w);
}
__device__ __forceinline__ void philox_normals(unsigned long long seed, long long gchain, long long step, int P,
                                               double* z) {
  for (int m = 0; 2 * m < P; m++) {
    double u1, u2, s, c;
    philox_pair(seed, gchain, step, (unsigned)m, &u1, &u2);
    double rad = sqrt(-2.0 * log(u1));
    sincospi(2.0 * u2, &s, &c);
    z[2 * m] = rad * c;
    if (2 * m + 1 < P) z[2 * m + 1] = rad * s;
  }
}
__device__ __forceinline__ double philox_uniform(unsigned long long seed, long long gchain, long long step) {
  double a, b;
  philox_pair(seed, gchain, step, 0xFFFFu, &a, &b);
  return a;
}

// ------------------------------------------------------------------ trace ring access
__device__ __forceinline__ double tprob(const MfitChains& c, long long row, long long w) {
  return c.trace_prob[(size_t)(row % c.cap) * c.stride + w];
}
__device__ __forceinline__ double ttheta(const MfitChains& c, long long row, int i, long long w) {
  return c.trace_theta[((size_t)(row % c.cap) * c.P + i) * c.stride + w];
}
// (walker-get w :get :steps :take take): the newest min(length, take) rows
__device__ __forceinline__ long long wlen(const MfitChains& c, long long w, long long take) {
  long long m = c.length[w] < take ? c.length[w] : take;
  if (m > c.rows[w]) m = c.rows[w];
  if (m > c.cap) m = c.cap;
  return m;
}
// walker-get :acceptance :506-508 -> exact count / len
__device__ void acc_window(const MfitChains& c, long long w, long long take, long long* cnt, long long* len) {
  long long m = wlen(c, w, take);
  long long r = c.rows[w] - 1;
  long long n = 1;  // the oldest element of the window is always kept
  unsigned long long prev = (unsigned long long)__double_as_longlong(tprob(c, r, w));
  for (long long k = 1; k < m; k++) {
    unsigned long long b = (unsigned long long)__double_as_longlong(tprob(c, r - k, w));
    if (b != prev) n++;
    prev = b;
  }
  *cnt = m > 0 ? n : 0;
  *len = m;
}
// (< acc thr) / (> acc thr) with thr a single-float literal compared exactly against the rational count/len
__device__ __forceinline__ bool acc_lt(long long cnt, long long len, float thr) {
  return (double)cnt < (double)thr * (double)len;
}
__device__ __forceinline__ bool acc_gt(long long cnt, long long len, float thr) {
  return (double)cnt > (double)thr * (double)len;
}
// stable-probs-p :880-885
__device__ bool stable_probs(const MfitChains& c, long long w, long long S) {
  long long m = wlen(c, w, S);
  if (m < 200) return false;
  long long hi = c.rows[w] - 1, lo = c.rows[w] - m;
  double early = tprob(c, hi, w);
  for (long long j = hi - 199; j <= hi; j++) early = fmax(early, tprob(c, j, w));
  double late = tprob(c, lo, w);
  for (long long j = lo; j < lo + 200; j++) late = fmax(late, tprob(c, j, w));
  double mn = tprob(c, lo, w);
  for (long long j = lo; j <= hi; j++) mn = fmin(mn, tprob(c, j, w));
  double spread = early - mn;
  return (fabs(early - late) < 0.5) && (4.0 < spread && spread < 9.0);
}

__device__ __forceinline__ double* ll_ptr(const MfitChains& c, int d) {
  if (c.xch) {
    const MfitDsDev& s = c.ds[d];
    return const_cast<double*>(c.xch) + (size_t)c.xpar * c.xpar_stride + ((size_t)(s.xrank * c.xslots + s.xslot) * 2 + 0) * c.stride;
  }
  return c.lld + (size_t)d * c.stride;
}
__device__ __forceinline__ double* lp_ptr(const MfitChains& c, int d) {
  if (c.xch) {
    const MfitDsDev& s = c.ds[d];
    return const_cast<double*>(c.xch) + (size_t)c.xpar * c.xpar_stride + ((size_t)(s.xrank * c.xslots + s.xslot) * 2 + 1) * c.stride;
  }
  return c.lpd + (size_t)d * c.stride;
}
// Per-data-set terms are read through L2 (never a stale L1 line): with the peer exchange other GPUs write them.
__device__ __forceinline__ double term_load(const double* p) { return __ldcg(p); }

// ------------------------------------------------------------------ peer exchange (data sets sharded over GPUs)
__device__ __forceinline__ unsigned long long ld_acquire_sys(const volatile unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// every block, before it reads terms other ranks own: wait until all of them have delivered exchange c.xepoch.
// Bounded (a few seconds): a peer that never delivers must not hang the GPU; the host reports it.
__device__ void peer_wait(const MfitChains& c) {
  if (c.xranks > 0) {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      for (int r = 0; r < c.xranks; r++) {
        if (r == c.xrank) continue;
        while (ld_acquire_sys(c.xflag + r) < c.xepoch + 1) {
          if (clock64() - t0 > 20000000000LL) { *c.xerr = 1; break; }
          __nanosleep(200);
        }
      }
    }
    __syncthreads();
  }
}
// this rank's rows (ll and lp of every data set it owns, chains [w_lo, w_hi) of this block) -> every peer's buffer
__device__ void peer_push(const MfitChains& c, long long w, bool valid) {
  if (c.xranks <= 0) return;
  if (valid)
    for (int d = 0; d < c.D; d++) {
      const MfitDsDev& s = c.ds[d];
      if (!s.owned) continue;
      const size_t off = (size_t)c.xpar * c.xpar_stride + ((size_t)(s.xrank * c.xslots + s.xslot) * 2) * c.stride + w;
      const double ll = c.xch[off], lp = c.xch[off + c.stride];
      for (int q = 0; q < c.xranks; q++) {
        if (q == c.xrank) continue;
        c.xpeer[q][off] = ll;
        c.xpeer[q][off + c.stride] = lp;
      }
    }
  __threadfence_system();   // this thread's remote stores are visible system-wide ...
  __syncthreads();
  if (threadIdx.x == 0) {   // ... before the last block to get here tells the peers
    if (atomicAdd(c.xdone, 1) == (int)gridDim.x - 1) {
      *c.xdone = 0;
      __threadfence_system();
      for (int q = 0; q < c.xranks; q++)
        if (q != c.xrank) st_release_sys(c.xflag_peer[q] + c.xrank, c.xepoch + 1);
    }
  }
}

// per-segment partial sums -> dataset log-likelihood.  The likelihood kernel cuts a data set into up to 1024
// segments (fine-grained work items, see mfit_eval.cuh), so the sum over segments is itself parallel: a block of
// MFIT_RED_T threads handles 32 chains x MFIT_RED_S slices; slice k adds the segments [k B, (k+1) B), B =
// ceil(n_seg / MFIT_RED_S), in order, and the slice sums are then added in slice order.  The shape depends on n_seg
// only (i.e. on the point count), so every bit of the result is independent of chain count, grid and GPU count.
#define MFIT_RED_S 16
#define MFIT_RED_T (32 * MFIT_RED_S)
// called by all MFIT_RED_T threads of the block (w may be >= c.W: such threads only take part in the barriers);
// the result is valid in the threads of slice 0
template <int S>
__device__ double reduce_ll(const MfitChains& c, int d, long long w, int slice, double (*red)[32]) {
  const MfitDsDev& s = c.ds[d];
  const int lane = threadIdx.x & 31;
  const int B = (s.n_seg + S - 1) / S;
  double sum = 0.0;
  if (w < c.w1) {
    const int lo = slice * B, hi = min(s.n_seg, lo + B);
    if (lo < hi) {
      const double* p = s.partial + (size_t)lo * c.stride + w;
      sum = *p;
      int k = lo + 1;
      for (; k + 8 <= hi; k += 8) {  // eight loads in flight, added in order
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = p[(size_t)(k - lo + j) * c.stride];
#pragma unroll
        for (int j = 0; j < 8; j++) sum = sum + v[j];
      }
      for (; k < hi; k++) sum = sum + p[(size_t)(k - lo) * c.stride];
    }
  }
  if (S > 1) {
    __syncthreads();  // red is free again
    red[slice][lane] = sum;
    __syncthreads();
    if (slice == 0 && B > 0) {
      const int used = (s.n_seg + B - 1) / B;
      for (int k = 1; k < used; k++) sum = sum + red[k][lane];
    }
  }
  const double cst = (s.cst_pc != nullptr && w < c.W) ? s.cst_pc[w] : s.cst;
  switch (s.ll_kind) {
    case MFIT_LL_NORMAL: return cst + (-0.5 * sum);
    case MFIT_LL_NORMAL_CUTOFF: return sum;
    default: return cst + sum;
  }
}

// prior-bounds-let with constant bounds, body = bounds-total (:346-369); flat = 0d0 (:340-343)
__device__ double builtin_prior(const MfitChains& c, const MfitDsDev& s, const double* theta, long long w) {
  if (s.prior_kind != 1) return 0.0;
  double total = 0.0;
  for (int k = 0; k < s.nb; k++) {
    int id = s.b_idx[k];
    double p = id >= 0 ? theta[(size_t)id * c.stride + w] : 0.0;
    double lo = s.b_lo[k], hi = s.b_hi[k], b;
    if (lo < p && p < hi)
      b = 0.0;
    else {
      double a1 = fabs(p - hi), a2 = fabs(p - lo);
      double m = a1;
      if (a2 < a1) m = a2;
      b = -1e10 * (exp(m * 1e-5) - 1.0);
    }
    total = k == 0 ? b : total + b;
  }
  return total;
}

__device__ void scale_L(const MfitChains& c, long long w, double s) {  // scale-array :605-611
  const int PP = c.P * c.P;
  for (int e = 0; e < PP; e++) {
    double* p = c.L + (size_t)e * c.stride + w;
    *p = s * *p;
  }
}

// a chain that needs its l-matrix re-estimated goes on its group's work list: k_reest is launched with a small fixed grid
// that walks the list (a CTA per chain of the whole set, nearly all of which return at once, cost 50 us per step at 131072 chains)
__device__ __forceinline__ void flag_reest(const MfitChains& c, long long w) {
  c.need_reest[w] = 1;
  if (c.group < 0) return;   // k_run_small re-estimates in place
  const int k = atomicAdd(c.reest_count + c.group, 1);
  c.reest_list[c.w0 + k] = (int)w;
}

// ------------------------------------------------------------------ top of loop iteration: shutdown test + proposal
// Chains flagged need_reest are skipped here: k_reest rebuilds their l-matrix first and then runs this body itself.
__device__ void pre_body(const MfitChains& c, const MfitRunCfg& r, long long w) {
  if (c.done[w] || c.need_reest[w]) return;
  const int P = c.P;
  if (r.mode == MFIT_MODE_ADAPTIVE) {  // :905-917
    long long i = c.loop_i[w];
    bool sh = c.shutting[w] != 0;
    bool trig = (!sh && (r.n - i) < r.tail);
    if (!trig && r.auto_mode != MFIT_AUTO_NONE && !sh && i > 2 * r.S && (i % 1000 == 0)) {
      long long cnt, len;
      acc_window(c, w, 1000, &cnt, &len);
      if (acc_gt(cnt, len, 0.2f) && (double)cnt < 0.5 * (double)len) trig = stable_probs(c, w, r.S);
    }
    if (trig) {
      c.temperature[w] = 1.0;
      c.shutting[w] = 1;
      c.loop_i[w] = r.n - r.tail;
    }
  }
  // get-covariant-sample :679-700
  double z[MFIT_MAX_P];
  if (r.z) {
    const double* zp = r.z + ((size_t)w * r.nstream + r.t) * P;
    for (int j = 0; j < P; j++) z[j] = zp[j];
  } else
    philox_normals(r.seed, r.chain_base + w, c.age[w], P, z);
  for (int i = 0; i < P; i++) {
    double s = 0.0;
    for (int j = 0; j < P; j++) s = s + c.L[(size_t)(i * P + j) * c.stride + w] * z[j];
    c.prop_theta[(size_t)i * c.stride + w] = s + c.cur_theta[(size_t)i * c.stride + w];
  }
  for (int d = 0; d < c.D; d++) {
    const MfitDsDev& s = c.ds[d];
    if (s.owned && s.prior_kind != 2) lp_ptr(c, d)[w] = builtin_prior(c, s, c.prop_theta, w);
  }
}

// ------------------------------------------------------------------ after the likelihood: accept, push, adapt
__device__ void post_body(const MfitChains& c, const MfitRunCfg& r, long long w) {
  if (c.done[w]) return;
  const int P = c.P;
  // walker-make-step :1067-1070
  double ll = 0.0, lp = 0.0;
  for (int d = 0; d < c.D; d++) {
    double v = term_load(ll_ptr(c, d) + w);
    ll = d == 0 ? v : ll + v;
  }
  for (int d = 0; d < c.D; d++) {
    double v = term_load(lp_ptr(c, d) + w);
    lp = d == 0 ? v : lp + v;
  }
  const double p1 = ll + lp;
  const double p0 = c.cur_prob[w];
  const double T = c.temperature[w];
  bool acc = p1 > p0;  // :1091
  if (!acc && !r.greedy) {
    double u = r.u ? r.u[(size_t)w * r.nstream + r.t] : philox_uniform(r.seed, r.chain_base + w, c.age[w]);
    acc = (p1 - p0) / T > log(u);  // :1092
  }
  if (acc) {
    for (int i = 0; i < P; i++) c.cur_theta[(size_t)i * c.stride + w] = c.prop_theta[(size_t)i * c.stride + w];
    c.cur_prob[w] = p1;
    if (p1 > c.best_prob[w]) {  // :553-555
      for (int i = 0; i < P; i++) c.best_theta[(size_t)i * c.stride + w] = c.prop_theta[(size_t)i * c.stride + w];
      c.best_prob[w] = p1;
    }
  }
  // walker-modify :add-step :549-552
  {
    const long long row = c.rows[w];
    const size_t slot = (size_t)(row % c.cap);
    for (int i = 0; i < P; i++)
      c.trace_theta[(slot * P + i) * c.stride + w] = c.cur_theta[(size_t)i * c.stride + w];
    c.trace_prob[slot * c.stride + w] = c.cur_prob[w];
    c.rows[w] = row + 1;
    c.length[w] += 1;
    c.age[w] += 1;
    c.steps_run[w] += 1;
  }
  long long i = c.loop_i[w];
  if (r.mode == MFIT_MODE_ADAPTIVE) {
    const bool sh = c.shutting[w] != 0;
    if (!sh && i < r.Tn) c.temperature[w] = r.temps[i];  // :920-921
    if (r.mwl && i == c.reset_index[w]) {                 // :923-927
      if (c.length[w] > r.mwl) {
        c.length[w] = r.mwl;
        c.reset_index[w] += 1 + r.mwl;
      } else
        c.reset_index[w] += 1 + (r.mwl - c.length[w]);
    }
    if (i > 0) {  // :929-942
      bool enter = false, have = false;
      long long cnt = 0, len = 1;
      if (i % 200 == 0) {
        acc_window(c, w, 200, &cnt, &len);
        have = true;
        if (acc_lt(cnt, len, 0.2f) || acc_gt(cnt, len, 0.4f)) enter = true;
      }
      if (!enter && !sh && (i % (2 * r.S) == 0)) enter = true;
      if (enter) {
        if (!have) acc_window(c, w, 200, &cnt, &len);
        if (acc_gt(cnt, len, 0.2f) && acc_lt(cnt, len, 0.4f))
          flag_reest(c, w);
        else if (acc_lt(cnt, len, 0.2f)) {
          scale_L(c, w, 0.1);
          c.l_updates[w] += 1;
        } else if (acc_gt(cnt, len, 0.4f)) {
          scale_L(c, w, 1.9);
          c.l_updates[w] += 1;
        }
      }
    }
    i++;
    if (i >= r.n) c.done[w] = 1;
  } else if (r.mode == MFIT_MODE_MANY) {
    i++;
    if (i >= r.n) c.done[w] = 1;
  } else {  // walker-sample-region :960-969
    if (i > 0 && i % 20 == 0) {
      long long cnt, len;
      acc_window(c, w, 50, &cnt, &len);
      if (cnt * 50 == len) { scale_L(c, w, 0.25); c.l_updates[w] += 1; }
      else if (cnt * 50 > 4 * len) { scale_L(c, w, 1.7); c.l_updates[w] += 1; }
    }
    i++;
    if (i >= 3000) c.done[w] = 1;
  }
  c.loop_i[w] = i;
}

__global__ void k_pre(const MfitChains c, const MfitRunCfg r) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < c.W) pre_body(c, r, w);
}
// FIN: the per-segment sums of this rank's data sets are reduced here (blocks of 32 chains x S slices; S = 1, one thread
// per chain, when every data set has only a few segments -- walker sets with per-walker data have one); the per-chain
// work that follows runs in the threads of slice 0.  Without FIN (data sets sharded over ranks: k_finalize and the
// all-gather have run) one thread per chain.
template <bool FIN, bool FUSE, int S>
__global__ void k_post(const MfitChains c, MfitRunCfg r) {
  long long w;
  if (FIN) {
    __shared__ double red[S][32];
    const int slice = S > 1 ? threadIdx.x >> 5 : 0;
    w = S > 1 ? c.w0 + (long long)blockIdx.x * 32 + (threadIdx.x & 31) : c.w0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int d = 0; d < c.D; d++) {
      if (!c.ds[d].owned) continue;
      const double v = reduce_ll<S>(c, d, w, slice, red);
      if (slice == 0 && w < c.w1) ll_ptr(c, d)[w] = v;
    }
    if (slice != 0) return;
  } else {
    peer_wait(c);   // the other ranks' terms of this step have arrived (no-op with a host-side collective)
    w = c.w0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  }
  if (w >= c.w1) return;
  post_body(c, r, w);
  if (FUSE) {
    r.t += 1;
    if (c.xpar_stride != 0) {   // the next proposal's priors belong to the next exchange: the other half of the buffer
      MfitChains cn = c;
      cn.xpar ^= 1;
      pre_body(cn, r, w);
    } else
      pre_body(c, r, w);
  }
}
__global__ void k_finalize(const MfitChains c) {
  __shared__ double red[MFIT_RED_S][32];
  const int slice = threadIdx.x >> 5;
  const long long w = (long long)blockIdx.x * 32 + (threadIdx.x & 31);
  for (int d = 0; d < c.D; d++) {
    if (!c.ds[d].owned) continue;
    const double v = reduce_ll<MFIT_RED_S>(c, d, w, slice, red);
    if (slice == 0 && w < c.W) ll_ptr(c, d)[w] = v;
  }
  peer_push(c, w, slice == 0 && w < c.W);   // the threads of slice 0 wrote the rows; every thread takes part in the barrier
}
// kernels that read the exchanged terms outside k_post (first step, force-take-step, eval_logp's read-out) are preceded by this
__global__ void k_peer_wait(const MfitChains c) { peer_wait(c); }

// ------------------------------------------------------------------ run set-up (:866-901)
__global__ void k_init_run(const MfitChains c, const MfitRunCfg r, const double* l_shared) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= c.W) return;
  const int P = c.P, PP = P * P;
  c.shutting[w] = 0;
  c.need_reest[w] = 0;
  c.steps_run[w] = 0;
  c.l_updates[w] = 0;
  c.reset_index[w] = 10000;
  if (r.mode == MFIT_MODE_ADAPTIVE) {
    c.loop_i[w] = 1;
    c.done[w] = (1 >= r.n) ? 1 : 0;
    c.temperature[w] = r.temperature0;
    if (l_shared) {
      for (int e = 0; e < PP; e++) c.L[(size_t)e * c.stride + w] = l_shared[e];
    } else {
      for (int e = 0; e < PP; e++) {  // diagonal-covariance of the most likely params :899-900
        int i = e / P, j = e % P;
        c.L[(size_t)e * c.stride + w] = (i == j) ? c.best_theta[(size_t)i * c.stride + w] : 0.0;
      }
      long long cnt, len;
      acc_window(c, w, 100, &cnt, &len);
      if (!(c.length[w] < r.S || acc_lt(cnt, len, 0.1f))) flag_reest(c, w);  // :897-901
    }
  } else if (r.mode == MFIT_MODE_MANY) {
    c.loop_i[w] = 0;
    c.done[w] = (0 >= r.n) ? 1 : 0;
    c.temperature[w] = 1.0;
    const double* lw = l_shared + (r.l_per_chain ? (size_t)w * PP : 0);   // :851 makes the default l-matrix per walker
    for (int e = 0; e < PP; e++) c.L[(size_t)e * c.stride + w] = lw[e];
  } else {
    c.loop_i[w] = 1;
    c.done[w] = 0;
    c.temperature[w] = 1.0;
    for (int e = 0; e < PP; e++) {  // :959
      int i = e / P, j = e % P;
      double v = (i == j) ? c.best_theta[(size_t)i * c.stride + w] : 0.0;
      c.L[(size_t)e * c.stride + w] = r.sr_scale * v;
    }
  }
}

// ------------------------------------------------------------------ get-optimal-mcmc-l-matrix :888-895, one CTA per chain
// returns 1 when the l-matrix was replaced or rescaled (counts as an update)
__device__ int reest_body(const MfitChains& c, const MfitRunCfg& r, long long w, unsigned char* sm) {
  const int P = c.P, PP = P * P;
  double* avg = reinterpret_cast<double*>(sm);
  double* cov = avg + P;
  double* Ln = cov + PP;
  int* fwd = reinterpret_cast<int*>(Ln + PP);
  __shared__ int m_sh, ok_sh;
  const int tid = threadIdx.x, nth = blockDim.x;
  const double scale = (2.38 * 2.38) / (double)P;  // (/ (expt 2.38d0 2) num-params)

  if (r.sampling_opt == MFIT_OPT_BEST_VALUE) {  // :895
    for (int e = tid; e < PP; e += nth) {
      int i = e / P, j = e % P;
      double v = (i == j) ? c.best_theta[(size_t)i * c.stride + w] : 0.0;
      c.L[(size_t)e * c.stride + w] = 1e-5 * v;
    }
    return 1;
  }
  const long long mwin = wlen(c, w, r.S);
  const long long hi = c.rows[w] - 1, lo = c.rows[w] - mwin;
  if (tid == 0) {  // walker-get :forward-steps :497-502, newest first
    int m = 0;
    for (long long j = hi; j >= lo + 1; j--)
      if (tprob(c, j, w) > tprob(c, j - 1, w)) fwd[m++] = (int)(j - lo);
    m_sh = m;
  }
  __syncthreads();
  const int m = m_sh;
  if (m == 1) return 0;  // 0x0 result fails the dimension check :936 -> l-matrix kept
  if (m == 0) {          // (elt nil 0) -> type-error handler returns l-matrix, scaled in place
    for (int e = tid; e < PP; e += nth) {
      double* p = c.L + (size_t)e * c.stride + w;
      *p = scale * *p;
    }
    return 1;
  }
  const int n = m - 1;  // number of differences (older - newer), :263-275
#define MFIT_DIFF(k, i) (ttheta(c, lo + fwd[(k) + 1], (i), w) - ttheta(c, lo + fwd[(k)], (i), w))
  for (int i = tid; i < P; i += nth) {  // :626
    double s = MFIT_DIFF(0, i);
    for (int k = 1; k < n; k++) s = s + MFIT_DIFF(k, i);
    avg[i] = s / (double)n;
  }
  __syncthreads();
  for (int e = tid; e < PP; e += nth) {  // :636-643
    const int i = e / P, j = e % P;
    const double ai = avg[i], aj = avg[j];
    double s = 0.0;
    for (int k = 0; k < n; k++) s = s + ((MFIT_DIFF(k, i) - ai) * (MFIT_DIFF(k, j) - aj)) / (double)n;
    cov[e] = s;
  }
#undef MFIT_DIFF
  for (int e = tid; e < PP; e += nth) Ln[e] = 0.0;
  __syncthreads();
  if (tid == 0) {  // cholesky-decomp :583-598
    int ok = 1;
    for (int i = 0; i < P; i++)
      for (int k = 0; k <= i; k++) {
        double tmp = 0.0;
        for (int j = 0; j < k; j++) tmp = tmp + Ln[i * P + j] * Ln[k * P + j];
        double v;
        if (i == k) {
          double d = cov[i * P + k] - tmp;
          v = sqrt(d > 0.0 ? d : (d != d ? d : 0.0));
        } else
          v = (cov[i * P + k] - tmp) / Ln[k * P + k];
        Ln[i * P + k] = v;
        if (!isfinite(v)) ok = 0;
      }
    ok_sh = ok;
  }
  __syncthreads();
  if (ok_sh) {
    for (int e = tid; e < PP; e += nth) c.L[(size_t)e * c.stride + w] = scale * Ln[e];
  } else {  // FPO / DB0 handlers :892-893
    for (int e = tid; e < PP; e += nth) {
      double* p = c.L + (size_t)e * c.stride + w;
      *p = scale * *p;
    }
  }
  return 1;
}

// r.t is the take-step index of the step that FOLLOWS the re-estimate; the chain's proposal for it is made here
__global__ void k_reest(const MfitChains c, const MfitRunCfg r) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int n = *(volatile int*)(c.reest_count + c.group);
  for (int k = blockIdx.x; k < n; k += gridDim.x) {   // the order of the list does not matter: chains do not interact
    const long long w = c.reest_list[c.w0 + k];
    const int upd = reest_body(c, r, w, sm);
    __syncthreads();  // l-matrix writes of every thread are visible to thread 0
    if (threadIdx.x == 0) {
      c.need_reest[w] = 0;
      c.l_updates[w] += upd;
      pre_body(c, r, w);
    }
    __syncthreads();  // shared memory is reused by the next chain
  }
  if (threadIdx.x == 0) {  // the last CTA out empties the list for the next step
    __threadfence();
    if (atomicAdd(c.reest_count + 4 + c.group, 1) == (int)gridDim.x - 1) {
      c.reest_count[c.group] = 0;
      c.reest_count[4 + c.group] = 0;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------ small problems: many steps per launch
// The reference's own use (mcmc-fit: ONE walker, a few hundred to a few thousand points, test.lisp:14-24) is three kernel
// launches per step on the path above -- launch latency, not arithmetic.  Here a CTA owns a chain for `steps` steps: its
// threads stride over the points of every data set (built-in families, libdevice math, explicit fma as in the likelihood
// kernels), a fixed-shape tree adds the per-thread sums, thread 0 runs the very same post_body / pre_body as the three-launch
// path and the CTA the same reest_body.  Only the order of the likelihood sum differs from the large kernel's (thread-strided
// instead of segment-wise): within 1e-12 of the oracle like it, not bit-identical to it.
__device__ double small_model(const MfitSmallDs& s, const double* th, double x) {
  switch (s.family) {
    case MFIT_FAM_POLY: {
      double r = th[s.idx[s.degree]];
      for (int k = s.degree - 1; k >= 0; k--) r = fma(x, r, th[s.idx[k]]);
      return r;
    }
    case MFIT_FAM_EXPO: return fma(th[s.idx[0]], exp(th[s.idx[1]] * x), th[s.idx[2]]);
    case MFIT_FAM_SINUS: return fma(th[s.idx[0]], sin(fma(th[s.idx[1]], x, th[s.idx[2]])), th[s.idx[3]]);
    case MFIT_FAM_LORENTZ: {
      const double g2 = th[s.idx[2]] * th[s.idx[2]], d = x - th[s.idx[1]];
      return (th[s.idx[0]] * g2) / fma(d, d, g2) + th[s.idx[3]];
    }
    case MFIT_FAM_GAUSS_LIN: {
      const double z = (x - th[s.idx[1]]) * (1.0 / th[s.idx[2]]);
      return fma(th[s.idx[0]], exp(-0.5 * (z * z)), fma(th[s.idx[4]], x, th[s.idx[3]]));
    }
    default: {  // MFIT_FAM_NV_MULTI, degree = K dips: scale_1..K mu_1..K sigma bg0
      const int K = s.degree;
      const double g2 = th[s.idx[2 * K]] * th[s.idx[2 * K]];
      double r = th[s.idx[2 * K + 1]];
      for (int k = 0; k < K; k++) {
        const double d = x - th[s.idx[K + k]];
        r -= (th[s.idx[k]] * g2) / fma(d, d, g2);
      }
      return r;
    }
  }
}

__global__ void __launch_bounds__(MFIT_SMALL_THREADS)
k_run_small(MfitChains c, MfitRunCfg r, const __grid_constant__ MfitSmallArgs a, long long steps) {
  extern __shared__ __align__(16) unsigned char sm[];   // reest_body's scratch
  __shared__ double s_th[MFIT_MAX_P];
  __shared__ double s_red[MFIT_SMALL_THREADS];
  __shared__ int s_go;
  const long long w = c.w0 + blockIdx.x;
  if (w >= c.w1) return;
  const int tid = threadIdx.x;
  c.group = -1;   // re-estimates are done right here: no work list
  for (long long s = 0; s < steps; s++, r.t++) {
    if (tid == 0) s_go = c.done[w] ? 0 : 1;
    for (int i = tid; i < c.P; i += MFIT_SMALL_THREADS) s_th[i] = c.prop_theta[(size_t)i * c.stride + w];
    __syncthreads();
    if (!s_go) break;
    // walker-make-step :1067-1070 for the proposal: every data set's log-likelihood
    for (int d = 0; d < a.n_ds; d++) {
      const MfitSmallDs& ds = a.d[d];
      double acc = 0.0;
      for (long long i = tid; i < ds.n; i += MFIT_SMALL_THREADS) {
        const double f = small_model(ds, s_th, ds.x[i]);
        if (ds.ll_kind == MFIT_LL_NORMAL) {
          const double rr = fma(-f, ds.w[i], ds.yw[i]);
          acc = fma(rr, rr, acc);
        } else if (ds.ll_kind == MFIT_LL_NORMAL_CUTOFF) {
          const double rr = (ds.y[i] - f) * ds.w[i];
          double t = fma(-0.5 * rr, rr, ds.c[i]);
          t = (t > -5000.0) ? t : ((t != t) ? t : -5000.0);
          acc = acc + t;
        } else
          acc = acc + fma(ds.y[i], log(f), -f);
      }
      s_red[tid] = acc;
      __syncthreads();
      for (int h = MFIT_SMALL_THREADS / 2; h > 0; h >>= 1) {  // fixed-shape tree
        if (tid < h) s_red[tid] = s_red[tid] + s_red[tid + h];
        __syncthreads();
      }
      if (tid == 0) {
        const MfitDsDev& dd = c.ds[d];
        const double sum = s_red[0];
        ll_ptr(c, d)[w] = ds.n == 0 ? 0.0 : dd.ll_kind == MFIT_LL_NORMAL ? dd.cst + (-0.5 * sum) : dd.ll_kind == MFIT_LL_NORMAL_CUTOFF ? sum : dd.cst + sum;
      }
      __syncthreads();
    }
    if (tid == 0) {
      __threadfence_block();
      post_body(c, r, w);   // accept / push / adapt of this step (may flag a re-estimate)
    }
    __syncthreads();
    MfitRunCfg rn = r;
    rn.t = r.t + 1;
    if (c.need_reest[w]) {  // CTA-uniform: written by thread 0 before the barrier
      const int upd = reest_body(c, rn, w, sm);
      __syncthreads();
      if (tid == 0) {
        c.need_reest[w] = 0;
        c.l_updates[w] += upd;
      }
    }
    if (tid == 0) pre_body(c, rn, w);   // shutdown test + proposal of the next step
    __syncthreads();
  }
}

// re-map a trace ring to a larger capacity (slot = row % cap changes with cap)
__global__ void k_trace_regrow(const MfitChains c, double* new_theta, double* new_prob, long long new_cap) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= c.W) return;
  long long m = c.rows[w] < c.cap ? c.rows[w] : c.cap;
  for (long long k = 0; k < m; k++) {
    const long long row = c.rows[w] - 1 - k;
    const size_t so = (size_t)(row % c.cap), sn = (size_t)(row % new_cap);
    for (int i = 0; i < c.P; i++) new_theta[(sn * c.P + i) * c.stride + w] = c.trace_theta[(so * c.P + i) * c.stride + w];
    new_prob[sn * c.stride + w] = c.trace_prob[so * c.stride + w];
  }
}

// ------------------------------------------------------------------ walker-create first step / force-take-step
__global__ void k_first_step(const MfitChains c) {  // :1148-1162; prop_theta = theta0, lld/lpd evaluated
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= c.W) return;
  const int P = c.P;
  double ll = 0.0, lp = 0.0;
  for (int d = 0; d < c.D; d++) { double v = term_load(ll_ptr(c, d) + w); ll = d == 0 ? v : ll + v; }
  for (int d = 0; d < c.D; d++) { double v = term_load(lp_ptr(c, d) + w); lp = d == 0 ? v : lp + v; }
  const double p = ll + lp;
  for (int i = 0; i < P; i++) {
    double v = c.prop_theta[(size_t)i * c.stride + w];
    c.cur_theta[(size_t)i * c.stride + w] = v;
    c.best_theta[(size_t)i * c.stride + w] = v;
    c.trace_theta[(size_t)i * c.stride + w] = v;
  }
  c.cur_prob[w] = p;
  c.best_prob[w] = p;
  c.trace_prob[w] = p;
  c.rows[w] = 1;
  c.length[w] = 1;
  c.age[w] = 1;
  c.done[w] = 0;
  c.shutting[w] = 0;
  c.need_reest[w] = 0;
  c.steps_run[w] = 0;
  c.l_updates[w] = 0;
  c.loop_i[w] = 0;
  c.temperature[w] = 1.0;
  c.reset_index[w] = 10000;
}
__global__ void k_force_step(const MfitChains c) {  // :1124-1129; prop_theta = cur_theta re-evaluated
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= c.W) return;
  const int P = c.P;
  double ll = 0.0, lp = 0.0;
  for (int d = 0; d < c.D; d++) { double v = term_load(ll_ptr(c, d) + w); ll = d == 0 ? v : ll + v; }
  for (int d = 0; d < c.D; d++) { double v = term_load(lp_ptr(c, d) + w); lp = d == 0 ? v : lp + v; }
  const double p = ll + lp;
  c.cur_prob[w] = p;
  if (p > c.best_prob[w]) {
    for (int i = 0; i < P; i++) c.best_theta[(size_t)i * c.stride + w] = c.cur_theta[(size_t)i * c.stride + w];
    c.best_prob[w] = p;
  }
  const long long row = c.rows[w];
  const size_t slot = (size_t)(row % c.cap);
  for (int i = 0; i < P; i++) c.trace_theta[(slot * P + i) * c.stride + w] = c.cur_theta[(size_t)i * c.stride + w];
  c.trace_prob[slot * c.stride + w] = p;
  c.rows[w] = row + 1;
  c.length[w] += 1;
  c.age[w] += 1;
}
// priors of an arbitrary theta batch (eval_logp, first step): built-in kinds only
__global__ void k_priors(const MfitChains c, const double* theta) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= c.W) return;
  for (int d = 0; d < c.D; d++) {
    const MfitDsDev& s = c.ds[d];
    if (s.owned && s.prior_kind != 2) lp_ptr(c, d)[w] = builtin_prior(c, s, theta, w);
  }
}
__global__ void k_eval_finalize(const MfitChains c, double* prob, double* ll_out, double* lp_out) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= c.W) return;
  double ll = 0.0, lp = 0.0;
  for (int d = 0; d < c.D; d++) {
    double v = term_load(ll_ptr(c, d) + w);
    if (ll_out) ll_out[(size_t)w * c.D + d] = v;
    ll = d == 0 ? v : ll + v;
  }
  for (int d = 0; d < c.D; d++) {
    double v = term_load(lp_ptr(c, d) + w);
    if (lp_out) lp_out[(size_t)w * c.D + d] = v;
    lp = d == 0 ? v : lp + v;
  }
  if (prob) prob[w] = ll + lp;
}

__global__ void k_count_active(const MfitChains c, int* out) {
  __shared__ int s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  int local = 0;
  for (long long w = threadIdx.x; w < c.W; w += blockDim.x) local += c.done[w] ? 0 : 1;
  atomicAdd(&s, local);
  __syncthreads();
  if (threadIdx.x == 0) *out = s;
}
__global__ void k_keep_walks(const MfitChains c, long long keep) {  // :568-569
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < c.W && c.length[w] > keep) c.length[w] = keep;  // (subseq walk 0 keep) of a shorter walk is an error in the reference: never grow
}
// walker-modify :burn-walks :566-567 (the walk loses its OLDEST burn-number steps), :reset :570-573 ((last walk): the walk
// shrinks to its OLDEST step, which also becomes the last step) and :reset-to-most-likely :574-578.  The ring keeps the
// newest `length` rows as the walk, so the surviving step of the two resets is pushed as a new row of a walk of length 1.
__global__ void k_modify(const MfitChains c, int mode, long long number) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= c.W) return;
  const int P = c.P;
  if (mode == MFIT_MODIFY_BURN_WALKS) {
    const long long held = wlen(c, w, 0x7fffffffffffffffLL);
    c.length[w] = held > number ? held - number : 0;
    return;
  }
  const long long held = wlen(c, w, 0x7fffffffffffffffLL);
  double th[MFIT_MAX_P], p;
  if (mode == MFIT_MODIFY_RESET) {
    if (held < 1) return;
    const long long oldest = c.rows[w] - held;
    for (int i = 0; i < P; i++) th[i] = ttheta(c, oldest, i, w);
    p = tprob(c, oldest, w);
  } else {
    for (int i = 0; i < P; i++) th[i] = c.best_theta[(size_t)i * c.stride + w];
    p = c.best_prob[w];
  }
  const long long slot = c.rows[w] % c.cap;
  for (int i = 0; i < P; i++) {
    c.trace_theta[((size_t)slot * P + i) * c.stride + w] = th[i];
    c.cur_theta[(size_t)i * c.stride + w] = th[i];
  }
  c.trace_prob[(size_t)slot * c.stride + w] = p;
  c.cur_prob[w] = p;
  c.rows[w] += 1;
  c.length[w] = 1;
}
__global__ void k_stats(const MfitChains c, mfit_chain_stats* out) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= c.W) return;
  mfit_chain_stats s;
  s.length = c.length[w];
  s.age = c.age[w];
  s.steps_this_run = c.steps_run[w];
  s.last_prob = c.cur_prob[w];
  s.best_prob = c.best_prob[w];
  s.temperature = c.temperature[w];
  long long cnt, len;
  acc_window(c, w, 200, &cnt, &len);
  s.accept_200 = len > 0 ? (double)cnt / (double)len : 0.0;
  s.shutting_down = c.shutting[w];
  s.l_updates = c.l_updates[w];
  out[w] = s;
}
// newest-first gather of the trace: theta [W][take][P], prob [W][take]
__global__ void k_trace_gather(const MfitChains c, long long take, double* theta, double* prob, long long* nvalid) {
  long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long long)c.W * take) return;
  const long long r = id / c.W, w = id % c.W;  // adjacent threads = adjacent chains (coalesced ring reads)
  const long long m = wlen(c, w, take);
  if (r == 0 && nvalid) nvalid[w] = m;
  const double nanv = __longlong_as_double(0x7ff8000000000000LL);
  if (r < m) {
    const long long row = c.rows[w] - 1 - r;
    if (theta)
      for (int i = 0; i < c.P; i++) theta[((size_t)w * take + r) * c.P + i] = ttheta(c, row, i, w);
    if (prob) prob[(size_t)w * take + r] = tprob(c, row, w);
  } else {
    if (theta)
      for (int i = 0; i < c.P; i++) theta[((size_t)w * take + r) * c.P + i] = nanv;
    if (prob) prob[(size_t)w * take + r] = nanv;
  }
}
// walker-load :989-1001 revived (checkpoint / resume): the inverse of k_trace_gather.  Row r (newest first) of chain w
// goes to absolute row n_valid[w] - 1 - r of its ring; rows the ring cannot hold (r >= cap) are the oldest and drop out.
__global__ void k_trace_scatter(const MfitChains c, long long rows, const double* theta, const double* prob,
                                const long long* n_valid) {
  long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long long)c.W * rows) return;
  const long long r = id / c.W, w = id % c.W;  // adjacent threads = adjacent chains (coalesced ring writes)
  const long long m = n_valid[w];
  if (r >= m || r >= c.cap) return;
  const long long slot = (m - 1 - r) % c.cap;
  for (int i = 0; i < c.P; i++) c.trace_theta[((size_t)slot * c.P + i) * c.stride + w] = theta[((size_t)w * rows + r) * c.P + i];
  c.trace_prob[(size_t)slot * c.stride + w] = prob[(size_t)w * rows + r];
}
// the per-chain slots that go with a restored walk: last step = newest row, length, age, and the most likely step --
// the saved one, or (walker-modify :add-walks :558-563) the first strictly largest prob going from the oldest row up
__global__ void k_restore_state(const MfitChains c, const long long* n_valid, const long long* age, const double* best_theta,
                                const double* best_prob) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= c.W) return;
  const int P = c.P;
  const long long m = n_valid[w];
  const long long held = m < c.cap ? m : c.cap;
  c.rows[w] = m;
  c.length[w] = held;
  c.age[w] = age ? age[w] : m;
  const long long newest = m - 1;
  for (int i = 0; i < P; i++) c.cur_theta[(size_t)i * c.stride + w] = ttheta(c, newest, i, w);
  c.cur_prob[w] = tprob(c, newest, w);
  long long brow = m - held;
  double bp = tprob(c, brow, w);
  for (long long row = m - held + 1; row <= newest; row++) {
    const double p = tprob(c, row, w);
    if (p > bp) { bp = p; brow = row; }
  }
  const bool given = best_theta != nullptr && best_prob != nullptr && !(bp > best_prob[w]);
  for (int i = 0; i < P; i++)
    c.best_theta[(size_t)i * c.stride + w] = given ? best_theta[(size_t)w * P + i] : ttheta(c, brow, i, w);
  c.best_prob[w] = given ? best_prob[w] : bp;
  c.done[w] = 0;
  c.shutting[w] = 0;
  c.need_reest[w] = 0;
  c.steps_run[w] = 0;
  c.l_updates[w] = 0;
  c.loop_i[w] = 0;
  c.temperature[w] = 1.0;
  c.reset_index[w] = 10000;
}
__global__ void k_philox_stream(unsigned long long seed, long long chain_base, long long W, long long first,
                                long long nsteps, int P, double* z, double* u) {
  long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= W * nsteps) return;
  const long long w = id / nsteps, t = id % nsteps;
  double zz[MFIT_MAX_P];
  philox_normals(seed, chain_base + w, first + t, P, zz);
  for (int j = 0; j < P; j++) z[(size_t)id * P + j] = zz[j];
  u[id] = philox_uniform(seed, chain_base + w, first + t);
}
__global__ void k_transpose_in(const double* src, double* dst, long long W, int P, long long stride) {
  long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= W * P) return;
  const long long w = id % W;
  const int i = (int)(id / W);
  dst[(size_t)i * stride + w] = src[(size_t)w * P + i];
}
__global__ void k_transpose_out(const double* src, double* dst, long long W, int P, long long stride) {
  long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= W * P) return;
  const long long w = id % W;
  const int i = (int)(id / W);
  dst[(size_t)w * P + i] = src[(size_t)i * stride + w];
}
__global__ void k_l_read(const MfitChains c, double* out) {
  long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int PP = c.P * c.P;
  if (id >= (long long)c.W * PP) return;
  const long long w = id % c.W;
  const int e = (int)(id / c.W);
  out[(size_t)w * PP + e] = c.L[(size_t)e * c.stride + w];
}
__global__ void k_flush(double* buf, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += step) buf[i] = (double)i;
}

// ------------------------------------------------------------------ walker-get analytics over the trace (:487-543)
// One CTA per chain over the newest `take` rows of its trace ring.  The two covariance-type results keep the
// reference's operation order (sequential sums in list order = newest first, divide inside the sum, :614-643;
// Cholesky :583-598), so they are bit-identical to the CPU oracle; selection and counting are order-free.
//   sel: [W][take + 1] int scratch (row offsets of the selected steps, newest first)
__device__ unsigned long long dbits(double v) { return (unsigned long long)__double_as_longlong(v); }

// compact the offsets k in [0, m) (k = 0 is the newest row) for which pred(k) holds, in order; returns the count
template <class Pred>
__device__ int block_select(int m, int* out, Pred pred, int* sh /* [blockDim.x + 1] */, int limit = 0x7fffffff) {
  const int tid = threadIdx.x, nth = blockDim.x;
  int base = 0;
  for (int k0 = 0; k0 < m; k0 += nth) {
    const int k = k0 + tid;
    const int flag = (k < m && pred(k)) ? 1 : 0;
    sh[tid] = flag;
    __syncthreads();
    // inclusive scan (blockDim.x <= 128: a simple log-step scan)
    for (int off = 1; off < nth; off <<= 1) {
      int v = tid >= off ? sh[tid - off] : 0;
      __syncthreads();
      sh[tid] += v;
      __syncthreads();
    }
    if (flag && base + sh[tid] - 1 < limit) out[base + sh[tid] - 1] = k;
    const int total = sh[nth - 1];
    __syncthreads();
    base += total;
  }
  return base < limit ? base : limit;
}

// lplist-covariance :614-643 of n rows (accessor val(k, i), k in list order) -> cov[P*P]; avg is scratch [P]
template <class Val>
__device__ void block_covariance(int n, int P, Val val, double* avg, double* cov) {
  const int tid = threadIdx.x, nth = blockDim.x;
  for (int i = tid; i < P; i += nth) {
    double s = val(0, i);
    for (int k = 1; k < n; k++) s = s + val(k, i);
    avg[i] = s / (double)n;
  }
  __syncthreads();
  for (int e = tid; e < P * P; e += nth) {
    const int i = e / P, j = e % P;
    const double ai = avg[i], aj = avg[j];
    double s = 0.0;
    for (int k = 0; k < n; k++) s = s + ((val(k, i) - ai) * (val(k, j) - aj)) / (double)n;
    cov[e] = s;
  }
  __syncthreads();
}

__global__ void k_analytics(const MfitChains c, long long take, int* sel_all, double* ml_theta, double* ml_prob,
                            double* acceptance, double* cov_out, double* l_out) {
  const long long w = blockIdx.x;
  if (w >= c.W) return;
  extern __shared__ __align__(16) unsigned char sm[];
  const int P = c.P, PP = P * P, tid = threadIdx.x, nth = blockDim.x;
  double* avg = reinterpret_cast<double*>(sm);
  double* cov = avg + P;
  double* Ln = cov + PP;
  int* sh = reinterpret_cast<int*>(Ln + PP);        // [nth + 1]
  double* redp = reinterpret_cast<double*>(sh + nth + 2);  // [nth]
  int* redk = reinterpret_cast<int*>(redp + nth);          // [nth]
  int* sel = sel_all + (size_t)w * (take + 1);
  const int m = (int)wlen(c, w, take);
  const long long newest = c.rows[w] - 1;
  auto prob_at = [&](int k) { return tprob(c, newest - k, w); };
  const double nanv = __longlong_as_double(0x7ff8000000000000LL);

  // :most-likely-step -- (reduce (lambda (x y) (if (> x y) x y)) steps): the largest prob, of equal ones the oldest
  if (ml_prob || ml_theta) {
    double bp = -INFINITY;
    int bk = -1;
    for (int k = tid; k < m; k += nth) {
      const double p = prob_at(k);
      if (p >= bp) { bp = p; bk = k; }   // k ascending = towards older rows
    }
    redp[tid] = bp;
    redk[tid] = bk;
    __syncthreads();
    if (tid == 0) {
      for (int t = 1; t < nth; t++)
        if (redk[t] >= 0 && (redk[0] < 0 || redp[t] > redp[0] || (redp[t] == redp[0] && redk[t] > redk[0]))) {
          redp[0] = redp[t];
          redk[0] = redk[t];
        }
      if (redk[0] < 0) redk[0] = m - 1;  // every prob is NaN: reduce ends on the oldest step
    }
    __syncthreads();
    const int bk0 = redk[0];
    if (ml_prob && tid == 0) ml_prob[w] = m > 0 ? prob_at(bk0) : nanv;
    if (ml_theta)
      for (int i = tid; i < P; i += nth) ml_theta[(size_t)w * P + i] = m > 0 ? ttheta(c, newest - bk0, i, w) : nanv;
    __syncthreads();
  }
  // :unique-steps :491-495 (a step whose prob differs from the next older one's; the oldest always) and :acceptance :506-508
  const int n_uni = block_select(m, sel, [&](int k) { return k == m - 1 || dbits(prob_at(k)) != dbits(prob_at(k + 1)); }, sh);
  if (acceptance && tid == 0) acceptance[w] = m > 0 ? (double)n_uni / (double)m : nanv;
  if (cov_out) {
    __syncthreads();  // sel visible
    if (n_uni > 0) {
      block_covariance(n_uni, P, [&](int k, int i) { return ttheta(c, newest - sel[k], i, w); }, avg, cov);
      for (int e = tid; e < PP; e += nth) cov_out[(size_t)w * PP + e] = cov[e];
    } else
      for (int e = tid; e < PP; e += nth) cov_out[(size_t)w * PP + e] = nanv;
    __syncthreads();
  }
  // :l-matrix :543 = cholesky(covariance(differences (older - newer) of the :forward-steps :497-502))
  if (l_out) {
    const int n_fwd = block_select(m, sel, [&](int k) { return k + 1 < m && prob_at(k) > prob_at(k + 1); }, sh);
    __syncthreads();
    const int n = n_fwd - 1;
    if (n >= 1) {
      block_covariance(n, P, [&](int k, int i) { return ttheta(c, newest - sel[k + 1], i, w) - ttheta(c, newest - sel[k], i, w); },
                       avg, cov);
      for (int e = tid; e < PP; e += nth) Ln[e] = 0.0;
      __syncthreads();
      if (tid == 0) {  // cholesky-decomp :583-598
        for (int i = 0; i < P; i++)
          for (int k = 0; k <= i; k++) {
            double tmp = 0.0;
            for (int j = 0; j < k; j++) tmp = tmp + Ln[i * P + j] * Ln[k * P + j];
            if (i == k) {
              const double d = cov[i * P + k] - tmp;
              Ln[i * P + k] = sqrt(d > 0.0 ? d : (d != d ? d : 0.0));
            } else
              Ln[i * P + k] = (cov[i * P + k] - tmp) / Ln[k * P + k];
          }
      }
      __syncthreads();
      for (int e = tid; e < PP; e += nth) l_out[(size_t)w * PP + e] = Ln[e];
    } else
      for (int e = tid; e < PP; e += nth) l_out[(size_t)w * PP + e] = nanv;  // the reference signals an error here
  }
}

// :median-params :516-523 = nth-percentile 50 (:1493-1504) of every parameter over the newest `take` rows: a 4-bit
// radix select per (chain, parameter), adjacent lanes = adjacent chains so that every pass over the ring is coalesced
__device__ __forceinline__ unsigned long long order_key(double v) {
  const unsigned long long b = dbits(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}
__global__ void k_median(const MfitChains c, long long take, double* out) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long w = id % c.stride;
  const int i = (int)(id / c.stride);
  if (i >= c.P || w >= c.W) return;
  const int m = (int)wlen(c, w, take);
  if (m == 0) { out[(size_t)w * c.P + i] = __longlong_as_double(0x7ff8000000000000LL); return; }
  const long long newest = c.rows[w] - 1;
  long long rank = (m - 1) / 2;          // (floor (* 50 (- len 1) 1/100))
  const bool two = ((m - 1) & 1) != 0;   // non-zero remainder: mean of the elements at pos and pos + 1
  unsigned long long prefix = 0;
  for (int shift = 60; shift >= 0; shift -= 4) {
    unsigned cnt[16];
#pragma unroll
    for (int b = 0; b < 16; b++) cnt[b] = 0;
    for (int k = 0; k < m; k++) {
      const unsigned long long key = order_key(ttheta(c, newest - k, i, w));
      const bool match = shift == 60 || (key >> (shift + 4)) == prefix;
      const unsigned d = (unsigned)(key >> shift) & 15u;
#pragma unroll
      for (int b = 0; b < 16; b++) cnt[b] += (match && d == (unsigned)b) ? 1u : 0u;
    }
    unsigned d = 0;
#pragma unroll
    for (int b = 0; b < 16; b++) {
      if (d == (unsigned)b && rank >= cnt[b] && b < 15) { rank -= cnt[b]; d = b + 1; }
    }
    prefix = (prefix << 4) | d;
  }
  const double e1 = key_value(prefix);
  double res = e1;
  if (two) {  // the next order statistic: e1 again if it is repeated beyond `rank`, else the smallest larger value
    long long le = 0;
    unsigned long long next = ~0ull;
    for (int k = 0; k < m; k++) {
      const unsigned long long key = order_key(ttheta(c, newest - k, i, w));
      if (key <= prefix) le++;
      else if (key < next) next = key;
    }
    const long long pos = (m - 1) / 2;
    const double e2 = le >= pos + 2 ? e1 : key_value(next);
    res = (e1 + e2) / 2.0;
  }
  out[(size_t)w * c.P + i] = res;
}

// walker-get-data-and-fit :1248-1249: sorted-walks = the ceil(0.66 take) most probable entries of the WHOLE walk (repeated
// entries of a rejected proposal counted), reduced here to its distinct steps -- only the set matters to the max / min
// envelope -- and gathered as parameter columns for the model kernel (mfit_fit_envelope_body):
//   out_theta [chains of this launch][P][cols]: column j = j-th selected distinct step, newest first; column cols - 1 =
//   the parameters of the fit curve itself (fit[w][P], most likely / median of the newest `take` steps)
// One CTA per chain: an 8 x 8-bit radix select over the prob column of the ring finds the prob of the M-th best entry,
// a block-wide compaction keeps one row per run of equal probs at or above it.  (Entries tied with the M-th are all
// taken: they are repeats of one step, which the reference's sort keeps at least once.)
__device__ __forceinline__ unsigned long long prob_key(double p) { return p != p ? 0ull : order_key(p); }  // NaN sorts last
__global__ void k_envelope_select(const MfitChains c, long long take, long long w_first, long long cols, const double* fit,
                                  double* out_theta, int* n_sel, int* sel_all) {
  const long long wl = blockIdx.x, w = w_first + wl;
  if (w >= c.W) return;
  __shared__ unsigned hist[256];
  __shared__ int scan[128 + 1];
  __shared__ unsigned long long s_prefix;
  __shared__ long long s_rank;
  const int tid = threadIdx.x, nth = blockDim.x, P = c.P;
  int* sel = sel_all + (size_t)wl * cols;
  const int m_all = (int)wlen(c, w, 0x7fffffffffffffffLL);     // (walker-get walker :get :steps): every step held
  const long long tk = wlen(c, w, take);                       // take clamped to the walker length :1233-1235
  const long long newest = c.rows[w] - 1;
  double* outw = out_theta + (size_t)wl * P * cols;
  for (int i = tid; i < P; i += nth) outw[(size_t)i * cols + cols - 1] = fit[(size_t)w * P + i];
  const float prod = __fmul_rn(0.66f, (float)tk);              // (* 0.66 take): single-float
  long long M = (long long)ceilf(prod);
  if (m_all == 0 || M < 1) {
    if (tid == 0) n_sel[wl] = 0;
    return;
  }
  if (M > m_all) M = m_all;
  if (tid == 0) { s_prefix = 0ull; s_rank = M; }
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int b = tid; b < 256; b += nth) hist[b] = 0u;
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    for (int k = tid; k < m_all; k += nth) {
      const unsigned long long key = prob_key(tprob(c, newest - k, w));
      if (shift == 56 || (key >> (shift + 8)) == prefix) atomicAdd(&hist[(unsigned)(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      long long rank = s_rank;
      int b = 255;
      for (; b > 0; b--) {
        if (rank <= (long long)hist[b]) break;
        rank -= hist[b];
      }
      s_rank = rank;
      s_prefix = (prefix << 8) | (unsigned long long)b;
    }
    __syncthreads();
  }
  const unsigned long long thr = s_prefix;
  const int n = block_select(m_all, sel, [&](int k) {
    const double p = tprob(c, newest - k, w);
    return prob_key(p) >= thr && (k == m_all - 1 || dbits(p) != dbits(tprob(c, newest - k - 1, w)));
  }, scan, (int)cols - 1);
  __syncthreads();
  for (int i = 0; i < P; i++)
    for (int j = tid; j < n; j += nth) outw[(size_t)i * cols + j] = ttheta(c, newest - sel[j], i, w);
  if (tid == 0) n_sel[wl] = n;
}

inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

// ------------------------------------------------------------------ launchers
void mfit_launch_init_run(const MfitChains& c, const MfitRunCfg& r, const double* l_shared, cudaStream_t st) {
  k_init_run<<<blocks_for(c.W, 128), 128, 0, st>>>(c, r, l_shared);
}
void mfit_launch_run_small(const MfitChains& c, const MfitRunCfg& r, const MfitSmallArgs& a, long long steps, cudaStream_t st) {
  const int P = c.P;
  size_t smem = sizeof(double) * (P + 2 * (size_t)P * P) + sizeof(int) * (size_t)(r.S + 1);
  static size_t configured = 0;
  if (smem > 40 * 1024 && smem > configured) {
    cudaFuncSetAttribute(k_run_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  k_run_small<<<(unsigned)(c.w1 - c.w0), MFIT_SMALL_THREADS, smem, st>>>(c, r, a, steps);
}
void mfit_launch_reest(const MfitChains& c, const MfitRunCfg& r, cudaStream_t st) {
  const int P = c.P;
  size_t smem = sizeof(double) * (P + 2 * (size_t)P * P) + sizeof(int) * (size_t)(r.S + 1);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(k_reest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  const long long chains = c.w1 - c.w0;
  k_reest<<<(unsigned)(chains < 296 ? (chains > 0 ? chains : 1) : 296), 128, smem, st>>>(c, r);
}
void mfit_launch_pre(const MfitChains& c, const MfitRunCfg& r, cudaStream_t st) {
  k_pre<<<blocks_for(c.W, 128), 128, 0, st>>>(c, r);
}
void mfit_launch_post(const MfitChains& c, const MfitRunCfg& r, bool fuse_pre, cudaStream_t st) {
  const bool fin = c.xch == nullptr;
  const unsigned g = blocks_for(c.w1 - c.w0, 128);
  if (fin) {
    if (c.max_seg <= 8) {  // few segments: one thread per chain
      if (fuse_pre) k_post<true, true, 1><<<g, 128, 0, st>>>(c, r);
      else k_post<true, false, 1><<<g, 128, 0, st>>>(c, r);
    } else {
      const unsigned gr = blocks_for(c.w1 - c.w0, 32);
      if (fuse_pre) k_post<true, true, MFIT_RED_S><<<gr, MFIT_RED_T, 0, st>>>(c, r);
      else k_post<true, false, MFIT_RED_S><<<gr, MFIT_RED_T, 0, st>>>(c, r);
    }
  } else {
    if (fuse_pre) k_post<false, true, 1><<<g, 128, 0, st>>>(c, r);
    else k_post<false, false, 1><<<g, 128, 0, st>>>(c, r);
  }
}
void mfit_launch_finalize(const MfitChains& c, cudaStream_t st) {
  k_finalize<<<blocks_for(c.W, 32), MFIT_RED_T, 0, st>>>(c);
}
void mfit_launch_peer_wait(const MfitChains& c, cudaStream_t st) {
  if (c.xranks > 0) k_peer_wait<<<1, 32, 0, st>>>(c);
}
void mfit_launch_first_step(const MfitChains& c, cudaStream_t st) {
  k_first_step<<<blocks_for(c.W, 128), 128, 0, st>>>(c);
}
void mfit_launch_force_step(const MfitChains& c, cudaStream_t st) {
  k_force_step<<<blocks_for(c.W, 128), 128, 0, st>>>(c);
}
void mfit_launch_priors(const MfitChains& c, const double* theta, cudaStream_t st) {
  k_priors<<<blocks_for(c.W, 128), 128, 0, st>>>(c, theta);
}
void mfit_launch_count_active(const MfitChains& c, int* out, cudaStream_t st) {
  k_count_active<<<1, 1024, 0, st>>>(c, out);
}
void mfit_launch_keep_walks(const MfitChains& c, long long keep, cudaStream_t st) {
  k_keep_walks<<<blocks_for(c.W, 128), 128, 0, st>>>(c, keep);
}
void mfit_launch_modify(const MfitChains& c, int mode, long long number, cudaStream_t st) {
  k_modify<<<blocks_for(c.W, 128), 128, 0, st>>>(c, mode, number);
}
void mfit_launch_stats(const MfitChains& c, void* stats_dev, cudaStream_t st) {
  k_stats<<<blocks_for(c.W, 128), 128, 0, st>>>(c, (mfit_chain_stats*)stats_dev);
}
void mfit_launch_trace_gather(const MfitChains& c, long long take, double* theta, double* prob, long long* nvalid,
                              cudaStream_t st) {
  k_trace_gather<<<blocks_for((long long)c.W * take, 256), 256, 0, st>>>(c, take, theta, prob, nvalid);
}
void mfit_launch_trace_restore(const MfitChains& c, long long rows, const double* theta, const double* prob,
                               const long long* n_valid, const long long* age, const double* best_theta,
                               const double* best_prob, cudaStream_t st) {
  k_trace_scatter<<<blocks_for((long long)c.W * rows, 256), 256, 0, st>>>(c, rows, theta, prob, n_valid);
  k_restore_state<<<blocks_for(c.W, 128), 128, 0, st>>>(c, n_valid, age, best_theta, best_prob);
}
void mfit_launch_philox_stream(unsigned long long seed, long long chain_base, long long W, long long first,
                               long long nsteps, int P, double* z, double* u, cudaStream_t st) {
  k_philox_stream<<<blocks_for(W * nsteps, 128), 128, 0, st>>>(seed, chain_base, W, first, nsteps, P, z, u);
}
void mfit_launch_transpose_in(const double* src, double* dst, long long W, int P, long long stride, cudaStream_t st) {
  k_transpose_in<<<blocks_for(W * P, 256), 256, 0, st>>>(src, dst, W, P, stride);
}
void mfit_launch_transpose_out(const double* src, double* dst, long long W, int P, long long stride, cudaStream_t st) {
  k_transpose_out<<<blocks_for(W * P, 256), 256, 0, st>>>(src, dst, W, P, stride);
}
void mfit_launch_eval_finalize(const MfitChains& c, double* prob, double* ll, double* lp, cudaStream_t st) {
  k_eval_finalize<<<blocks_for(c.W, 128), 128, 0, st>>>(c, prob, ll, lp);
}
void mfit_launch_l_read(const MfitChains& c, double* out, cudaStream_t st) {
  k_l_read<<<blocks_for((long long)c.W * c.P * c.P, 256), 256, 0, st>>>(c, out);
}
void mfit_launch_trace_regrow(const MfitChains& c, double* new_theta, double* new_prob, long long new_cap,
                              cudaStream_t st) {
  k_trace_regrow<<<blocks_for(c.W, 128), 128, 0, st>>>(c, new_theta, new_prob, new_cap);
}
void mfit_launch_flush(void* buf, size_t bytes, cudaStream_t st) {
  k_flush<<<148 * 8, 256, 0, st>>>((double*)buf, bytes / 8);
}
void mfit_launch_analytics(const MfitChains& c, long long take, int* sel, double* ml_theta, double* ml_prob,
                           double* acceptance, double* cov, double* l_out, cudaStream_t st) {
  const int nth = 128;
  const size_t smem = sizeof(double) * (c.P + 2 * (size_t)c.P * c.P) + sizeof(int) * (nth + 2) + sizeof(double) * nth +
                      sizeof(int) * nth + 16;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(k_analytics, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  k_analytics<<<(unsigned)c.W, nth, smem, st>>>(c, take, sel, ml_theta, ml_prob, acceptance, cov, l_out);
}
void mfit_launch_envelope_select(const MfitChains& c, long long take, long long w_first, long long n_chains, long long cols,
                                 const double* fit, double* out_theta, int* n_sel, int* sel, cudaStream_t st) {
  k_envelope_select<<<(unsigned)n_chains, 128, 0, st>>>(c, take, w_first, cols, fit, out_theta, n_sel, sel);
}
void mfit_launch_median(const MfitChains& c, long long take, double* out, cudaStream_t st) {
  k_median<<<blocks_for(c.stride * c.P, 128), 128, 0, st>>>(c, take, out);
}
