// api.cu -- host side of libmfit_cuda.so: the C ABI of include/mfit_cuda.h.
//
// What the reference does with Lisp lists inside one SBCL image (walker-create :1132-1163,
// walker-make-step :1067-1070, walker-adaptive-steps-full :862-942) is here a context that owns
// flat FP64 device arrays and enqueues kernels on one CUDA stream:
//
//   per MH step:   eval kernel per dataset (mfit_eval.cuh)      -- hot loop #2, >99 % of the time
//                  [ncclAllGather of ll/lp when datasets are sharded over ranks]
//                  k_post  (accept, push, adapt; + next proposal)   \  hot loop #1,
//                  k_reest (forward-step covariance + Cholesky)     /  chain_kernels.cu
//
// The host never looks at chain state inside a run except for a lagging, non-blocking poll of
// "how many chains are still active" (auto-stop) and the stop flag (mfit-walker-estop :860).
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "mfit_chain.h"
#include "mfit_eval.cuh"
#include "mfit_internal.h"

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync on the context's stream, release
// threshold raised so that freed blocks stay cached): walker-create / trace reads allocate and free hundreds of MB per
// call, which costs tens of milliseconds with cudaMalloc / cudaFree and next to nothing from a warm pool (the pool's
// release threshold is raised to 16 GB in mfit_ctx_create).  Everything
// a context does is enqueued on its one stream, so stream-ordered frees are safe.
static thread_local cudaStream_t g_alloc_stream = nullptr;
static inline cudaError_t mfit_pool_malloc(void** p, size_t bytes) { return cudaMallocAsync(p, bytes, g_alloc_stream); }
static inline cudaError_t mfit_pool_free(void* p) { return cudaFreeAsync(p, g_alloc_stream); }
#define cudaMalloc(p, bytes) mfit_pool_malloc((void**)(p), (bytes))
#define cudaFree(p) mfit_pool_free(p)

namespace {

std::string g_create_error;

struct Dataset {
  long long n = -1, n_pad = 0;
  bool x_is_sequence = false;  // uploaded through mfit_dataset_upload_xd (a one-component sequence is still a sequence)
  int x_dims = 1;        // components of every x (multi-dimensional inputs :1137); x holds x_dims arrays of n_pad
  int model_xdims = 0;   // 0: the model takes x as a number; k > 0: it reads components 0 .. k-1 of a sequence
  double *x = nullptr, *y = nullptr, *yw = nullptr, *w = nullptr, *c = nullptr;
  double xmin = 0.0, xmax = 0.0;  // -inf / +inf when some x is not finite
  double cst_normal = 0.0, cst_poisson = 0.0;
  bool poisson_ok = false;     // every y is a count (non-negative integer-valued, <= 5e6)
  bool poisson_ready = false;  // cst_poisson / cst_poisson_pc hold -sum log y_i!
  double ymax = 0.0;
  // walker set with per-walker data: W_pc chains, chain w owns row w of y / yw / w / c (and of x unless x_shared)
  long long W_pc = 0;
  bool x_shared = false, sigma_shared = false;
  bool sigma_uniform = false;   // one number for every point of every walker (or no :data-error at all)
  double *cst_normal_pc = nullptr, *cst_poisson_pc = nullptr;  // [W_pc] device
  cudaKernel_t k_user_pc[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // [kind + 3 * staged]
  int family = -1, degree = 0;
  int idx[MFIT_MAXK];
  int ll_kind = MFIT_LL_NORMAL;
  int prior_kind = 0;
  int nb = 0;
  int* b_idx = nullptr;
  double *b_lo = nullptr, *b_hi = nullptr;
  // :log-liklihood given as a per-point lambda (mfit_likelihood_sexpr): compiled with the model when first needed
  std::string ll_text, model_code;     // the lambda; the lowered MfitUser struct of a user :function
  cudaLibrary_t lib_term = nullptr;
  cudaKernel_t k_term = nullptr;
  int term_narr = 2;                   // arrays the term reads per point: 2 (x, y) or 4 (+ the log-factorial table in the c slot)
  double* lf = nullptr;                // [n_pad] log y_i! (single-float accumulation) for such a term
  double err_head[16];                 // first elements of the stddev list ((elt error k) in the lambda)
  int n_err_head = 0;
  cudaLibrary_t lib_model = nullptr, lib_prior = nullptr;
  cudaKernel_t k_user[3] = {nullptr, nullptr, nullptr};
  cudaKernel_t k_user_fit = nullptr;
  cudaKernel_t k_prior = nullptr;
  int owner = 0, xslot = 0;
  bool owned = true;
  int tiles_per_seg = 1, n_seg = 0;
};

// buffers needed to evaluate W parameter vectors against every dataset
struct Workspace {
  long long W = 0, stride = 0;
  double *lld = nullptr, *lpd = nullptr;
  std::vector<double*> partial;
  MfitDsDev* ds_dev = nullptr;
  double* xch = nullptr;
  int xslots = 0;
  bool peer = false;   // xch is the context's peer-mapped exchange buffer (not owned by the workspace)
};

// Peer exchange state of a context whose data sets are sharded over GPUs (mfit_peer_setup / mfit_peer_attach): one plain
// cudaMalloc block [flags 512 B][done, err 512 B][two halves of the exchange buffer], mapped into every peer.
struct Peers {
  bool ready = false;
  int rank = 0, n = 0;
  long long max_stride = 0;
  int max_slots = 0;
  char* base = nullptr;
  double* xch = nullptr;
  unsigned long long* flags = nullptr;
  int *xdone = nullptr, *xerr = nullptr;
  double* xpeer[8] = {nullptr};
  unsigned long long* flag_peer[8] = {nullptr};
  void* opened[8] = {nullptr};        // IPC mappings to close
  unsigned long long epoch = 0;       // exchanges completed (all ranks make the same calls, so it agrees everywhere)
};
constexpr size_t kPeerHeader = 1024;

struct EventPair { cudaEvent_t a, b; };

// launch shape of a shared-data likelihood kernel (mfit_eval.cuh): persistent CTAs of `warps` independent warps
struct EvalShape { int warps = 0, stages = 0, ctas = 0; size_t smem = 0; };

// What every context on a device can share, made once per process: the fixed cost of a context (walker-create is called
// per fit) was dominated by things that do not depend on it -- cudaGetDeviceProperties (~1 ms), two 16 MB pinned bounce
// buffers (~10 ms each to pin, again to unpin), kernel attribute / occupancy queries, streams and events for the chain
// groups.  Contexts borrow them; a mutex serialises the borrowers of the bounce buffers and stream sets.
struct StreamSet { cudaStream_t g[3]; cudaEvent_t fork_ev, join_ev[3], poll_ev[4], stage_ev[2]; int* pinned_ints; };
struct DeviceShared {
  std::mutex mu;
  bool probed = false;
  int cc_major = 0, cc_minor = 0, sms = 0, smem_optin = 0, smem_sm = 0;
  char* stage[2] = {nullptr, nullptr};      // pinned bounce buffers for large device -> host reads
  std::mutex stage_mu;
  std::vector<cudaEvent_t> slice_ev;        // one per 1 MB slice of a small read (d2h_staged)
  std::vector<StreamSet*> free_sets;        // stream / event sets not in use by a context
  std::map<const void*, int> resident;      // per-walker eval kernel -> CTAs resident on the whole device
  std::map<const void*, EvalShape> shape;   // shared-data eval kernel -> launch shape (built-in kernels only)
};
DeviceShared* device_shared(int device) {
  static std::mutex mu;
  static std::map<int, DeviceShared*> all;
  std::lock_guard<std::mutex> lock(mu);
  auto it = all.find(device);
  if (it == all.end()) it = all.emplace(device, new DeviceShared()).first;
  return it->second;
}

}  // namespace

struct mfit_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  long long launches = 0;
  int D = 0, P = 0;
  std::vector<Dataset> ds;
  int seg_tiles_override = 0;
  std::map<const void*, int> resident;  // per-walker eval kernel -> CTAs resident on the whole device (NVRTC kernels of this context)
  std::map<const void*, EvalShape> shape;  // shared-data eval kernel -> launch shape (NVRTC kernels of this context)
  unsigned long long* mode_dev = nullptr;  // [4] tiles per kernel mode, counted while timing is on (mfit_eval_mode_tiles)
  int* sched_dev = nullptr;             // [2 per group] work-item counter / finished-CTA counter of the eval kernels (zero between launches)
  // A run splits the chains into two or three groups that advance on separate streams: the per-chain kernels of one
  // group (segment reduction, accept/adapt, next proposal -- latency-bound, one warp per SM) and the tail of its
  // likelihood kernel then overlap the likelihood kernels of the others.  Chains never interact: results are unchanged.
  static constexpr int kMaxGroups = 4;
  DeviceShared* shared = nullptr;
  StreamSet* set = nullptr;             // borrowed from `shared` for the life of the context
  cudaStream_t gstream[kMaxGroups] = {nullptr, nullptr, nullptr, nullptr};  // [0] is `stream`
  cudaEvent_t fork_ev = nullptr, join_ev[kMaxGroups] = {nullptr, nullptr, nullptr, nullptr};
  int n_groups = 1;
  int groups_override = 0;              // mfit_set_groups: 0 = automatic
  int small_mode = 0;                   // mfit_set_small_path: 0 = automatic, 1 = never, 2 = whenever the problem qualifies

  // chains
  bool have_chains = false;
  MfitChains ch;
  Workspace ws;
  std::vector<void*> chain_allocs;
  unsigned long long seed = 0;
  long long chain_base = 0;
  bool cap_fixed = false;
  long long rows_upper = 0;  // no chain has pushed more trace rows than this
  double* temps_dev = nullptr;
  long long temps_len = 0;
  double *z_dev = nullptr, *u_dev = nullptr;
  size_t z_bytes = 0, u_bytes = 0;
  double* lshared_dev = nullptr;
  int* active_dev = nullptr;
  int* active_host = nullptr;  // pinned [4]
  int* xflag_dev = nullptr;    // [n_ranks][2] (chains still active, stop requested) of every rank: the collective stop test
  int* xflag_host = nullptr;   // pinned copy
  cudaEvent_t poll_ev[4];
  bool poll_ev_made = false;

  // multi-GPU
  Peers peers;
  void* comm = nullptr;
  int rank = 0, n_ranks = 1;
  bool sharded = false;
  int xslots = 0;

  std::atomic<int> stop{0};
  std::atomic<long long>* stop_at = nullptr;   // several ranks in one process: the poll boundary at which all of them stop

  // an open run (mfit_run_begin ... mfit_run_advance)
  bool run_open = false;
  MfitRunCfg run_cfg;
  long long run_t = 0, run_max = 0;

  // measurement
  bool timing = false;
  std::vector<EventPair> ev;
  size_t ev_used = 0;
  double eval_seconds = 0.0;
  long long eval_launches = 0;
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  // pinned bounce buffers for large device -> host reads (traces): pageable destinations are filled from here
  char* stage[2] = {nullptr, nullptr};   // = shared->stage while a read is in progress
  cudaEvent_t stage_ev[2];
};

namespace {

int fail(mfit_ctx* c, int code, const char* fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  else g_create_error = buf;
  return code;
}

#define CU(call)                                                                                              \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess)                                                                                    \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? MFIT_ERR_NOMEM : MFIT_ERR_CUDA, "%s: %s", #call,     \
                  cudaGetErrorString(e_));                                                                    \
  } while (0)

long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

void free_dev(void* p) { if (p) cudaFree(p); }

void dataset_free_term(Dataset& d);
void dataset_free_data(Dataset& d) {
  dataset_free_term(d);
  free_dev(d.x);  // x, y, yw, w, c are one allocation (mfit_dataset_upload_xd)
  free_dev(d.cst_normal_pc); free_dev(d.cst_poisson_pc);
  d.x = d.y = d.yw = d.w = d.c = nullptr;
  d.cst_normal_pc = d.cst_poisson_pc = nullptr;
  d.W_pc = 0;
  d.x_shared = d.sigma_shared = d.sigma_uniform = false;
  d.n = -1;
}
void dataset_free_prior(Dataset& d) {
  free_dev(d.b_idx); free_dev(d.b_lo); free_dev(d.b_hi);
  d.b_idx = nullptr; d.b_lo = d.b_hi = nullptr;
  d.nb = 0;
  if (d.lib_prior) { cudaLibraryUnload(d.lib_prior); d.lib_prior = nullptr; d.k_prior = nullptr; }
}
void dataset_free_term(Dataset& d) {  // the compiled likelihood term depends on the model and on the data (error list, y table)
  if (d.lib_term) { cudaLibraryUnload(d.lib_term); d.lib_term = nullptr; }
  d.k_term = nullptr;
  free_dev(d.lf);
  d.lf = nullptr;
}
void dataset_free_model(Dataset& d) {
  dataset_free_term(d);
  d.model_code.clear();
  if (d.lib_model) { cudaLibraryUnload(d.lib_model); d.lib_model = nullptr; }
  d.k_user[0] = d.k_user[1] = d.k_user[2] = nullptr;
  for (auto& k : d.k_user_pc) k = nullptr;
  d.k_user_fit = nullptr;
  d.model_xdims = 0;
}

void workspace_free(Workspace& w) {
  free_dev(w.lld); free_dev(w.lpd); free_dev(w.ds_dev);
  if (!w.peer) free_dev(w.xch);
  for (double* p : w.partial) free_dev(p);
  w = Workspace();
}

void chains_free(mfit_ctx* ctx) {
  for (void* p : ctx->chain_allocs) free_dev(p);
  ctx->chain_allocs.clear();
  workspace_free(ctx->ws);
  ctx->have_chains = false;
}

// log-factorial :379-380 -- (reduce (lambda (x y) (+ x (log y))) (up-to n)): single-float accumulation
double log_factorial_ref(long long n, std::vector<float>& memo) {
  if (n < (long long)memo.size()) return (double)memo[(size_t)n];
  if (memo.empty()) memo.push_back(0.0f);
  for (long long j = (long long)memo.size(); j <= n; j++) memo.push_back(memo[(size_t)j - 1] + (float)std::log((double)j));
  return (double)memo[(size_t)n];
}

// Sum of the theta-independent terms hoisted out of the likelihood (-ln(2 pi)/2 - ln sigma_i, -log k_i!).  With a
// scalar :data-error every term is the SAME number, and a plain running sum then rounds the same way at every step:
// the drift grows like N eps, not sqrt(N) eps (measured: 1.3e-10 relative on C5's 1e7 points, test_gpu_fullsize.py).
// The reference adds these terms together with the random -z^2/2 (:400), which keeps its own sum within ~3e-14 of
// exact; the hoisted constant is therefore formed exactly-rounded-ish (Neumaier).
struct CompSum {
  double s = 0.0, c = 0.0;
  void add(double v) {
    const double t = s + v;
    c += std::fabs(s) >= std::fabs(v) ? (s - t) + v : (v - t) + s;
    s = t;
  }
  double value() const { return s + c; }
};

// segment geometry is a function of the point count alone, so that partial sums -- and therefore
// every log-likelihood bit -- do not depend on chain count, SM count or the number of GPUs
void plan_segments(mfit_ctx* ctx, Dataset& d) {
  if (d.W_pc > 0) {  // one warp sums a whole (short) data set
    d.tiles_per_seg = 1;
    d.n_seg = d.n > 0 ? 1 : 0;
    return;
  }
  const long long n_tiles = (d.n + MFIT_TILE - 1) / MFIT_TILE;
  // at most MFIT_MAX_SEG segments, and up to five tiles per work item as long as that leaves >= 64 items per
  // chain-block: per-item set-up costs ~4 % with one-tile items.  Inside a run the granularity hardly matters, because
  // the tail of one chain group's kernel overlaps the next group's (C2, bench.py with MFIT_SEG_TILES = 1 / 2 / 3 / 5 / 6 /
  // 7 / 8 / 10 / 12: 0.471 / 0.460 / 0.445 / 0.443 / 0.442 / 0.442 / 0.442 / 0.444 / 0.442 ms per step), but a launch on
  // its own (walker-make-step batches, the kernel-timing phase of bench.py) ends with a partly empty last wave: all
  // 4096 C2 chains are 1568 items of eight tiles on 444 resident CTAs = four waves of eight tile-times where 28.2
  // would do (0.497 ms), 2512 items of five tiles = six waves of five (0.439 ms; tools/seg_sweep.sh).
  long long tps = (n_tiles + MFIT_MAX_SEG - 1) / MFIT_MAX_SEG;
  const long long want = n_tiles / 64 < 5 ? n_tiles / 64 : 5;
  if (tps < want) tps = want;
  if (ctx->seg_tiles_override > 0) tps = ctx->seg_tiles_override;
  else if (const char* e = getenv("MFIT_SEG_TILES")) { if (atoi(e) > 0) tps = atoi(e); }  // tuning aid
  if (tps < 1) tps = 1;
  d.tiles_per_seg = (int)tps;
  d.n_seg = (int)((n_tiles + tps - 1) / tps);
}

int workspace_make(mfit_ctx* ctx, Workspace& w, long long W) {
  workspace_free(w);
  w.W = W;
  w.stride = round_up(W > 0 ? W : 1, MFIT_CB);
  const int D = ctx->D;
  CU(cudaMalloc(&w.lld, sizeof(double) * D * w.stride));
  CU(cudaMalloc(&w.lpd, sizeof(double) * D * w.stride));
  CU(cudaMemsetAsync(w.lld, 0, sizeof(double) * D * w.stride, ctx->stream));
  CU(cudaMemsetAsync(w.lpd, 0, sizeof(double) * D * w.stride, ctx->stream));
  w.partial.assign(D, nullptr);
  if (ctx->sharded) {
    w.xslots = ctx->xslots;
    const size_t cnt = (size_t)ctx->n_ranks * w.xslots * 2 * w.stride;
    if (ctx->peers.ready) {
      if (w.stride > ctx->peers.max_stride || w.xslots > ctx->peers.max_slots)
        return fail(ctx, MFIT_ERR_ARG, "%lld chains / %d data sets per rank exceed what mfit_peer_setup reserved", (long long)W, w.xslots);
      w.xch = ctx->peers.xch;
      w.peer = true;
    } else {
      CU(cudaMalloc(&w.xch, sizeof(double) * cnt));
      CU(cudaMemsetAsync(w.xch, 0, sizeof(double) * cnt, ctx->stream));
    }
  }
  std::vector<MfitDsDev> host(D);
  for (int d = 0; d < D; d++) {
    Dataset& s = ctx->ds[d];
    MfitDsDev& h = host[d];
    memset(&h, 0, sizeof h);
    h.owned = s.owned ? 1 : 0;
    h.ll_kind = s.ll_kind;
    h.prior_kind = s.prior_kind;
    h.nb = s.nb;
    h.b_idx = s.b_idx;
    h.b_lo = s.b_lo;
    h.b_hi = s.b_hi;
    h.xrank = s.owner;
    h.xslot = s.xslot;
    if (s.owned) {
      plan_segments(ctx, s);
      h.n_seg = s.n_seg;
      h.cst = s.ll_kind == MFIT_LL_NORMAL ? s.cst_normal : s.ll_kind == MFIT_LL_POISSON ? s.cst_poisson : 0.0;
      h.cst_pc = s.W_pc > 0 ? (s.ll_kind == MFIT_LL_NORMAL ? s.cst_normal_pc : s.ll_kind == MFIT_LL_POISSON ? s.cst_poisson_pc : nullptr) : nullptr;
      if (s.W_pc > 0 && s.W_pc != W)
        return fail(ctx, MFIT_ERR_ARG, "dataset %d holds data for %lld walkers, %lld chains requested", d, s.W_pc, (long long)W);
      if (s.n_seg > 0) CU(cudaMalloc(&w.partial[d], sizeof(double) * (size_t)s.n_seg * w.stride));
      h.partial = w.partial[d];
    }
  }
  CU(cudaMalloc(&w.ds_dev, sizeof(MfitDsDev) * (D > 0 ? D : 1)));
  CU(cudaMemcpyAsync(w.ds_dev, host.data(), sizeof(MfitDsDev) * D, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));  // `host` dies with this frame
  return 0;
}

// -sum_i log y_i! of a data set (of every row of a walker set): the table of log-factorial :379-380 -- (reduce (lambda (x y)
// (+ x (log y))) (up-to n)), accumulated in single-float -- is built on the host up to the largest count and looked up on
// the device (prep_kernels.cu).  Done when a log-poisson likelihood first needs it.
int ensure_poisson_const(mfit_ctx* ctx, Dataset& s) {
  CU(cudaSetDevice(ctx->device));
  g_alloc_stream = ctx->stream;
  std::vector<float> memo;
  log_factorial_ref((long long)s.ymax, memo);
  std::vector<double> table(memo.begin(), memo.end());
  if (table.empty()) table.push_back(0.0);
  const size_t rows = s.W_pc > 0 ? (size_t)s.W_pc : 1;
  const int blocks = rows > 1 ? (int)(rows < 2048 ? rows : 2048)
                              : (int)std::max<long long>(1, std::min<long long>(592, (s.n_pad + MFIT_PREP_THREADS - 1) / MFIT_PREP_THREADS));
  const size_t n_res = rows > 1 ? rows : (size_t)blocks;
  double* tab_dev = nullptr;
  MfitPrepBlock* dev = nullptr;
  CU(cudaMalloc(&tab_dev, sizeof(double) * table.size()));
  CU(cudaMalloc(&dev, sizeof(MfitPrepBlock) * n_res));
  std::vector<MfitPrepBlock> res(n_res);
  cudaError_t e = cudaMemcpyAsync(tab_dev, table.data(), sizeof(double) * table.size(), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    mfit_launch_poisson_const(s.y, tab_dev, s.n, s.n_pad, (long long)rows, dev, blocks, ctx->stream);
    ctx->launches++;
    e = cudaMemcpyAsync(res.data(), dev, sizeof(MfitPrepBlock) * n_res, cudaMemcpyDeviceToHost, ctx->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(tab_dev); cudaFree(dev);
  if (e != cudaSuccess) return fail(ctx, MFIT_ERR_CUDA, "log-factorial constants: %s", cudaGetErrorString(e));
  if (rows > 1) {
    std::vector<double> pc(rows);
    for (size_t r = 0; r < rows; r++) pc[r] = res[r].sum + res[r].err;
    CU(cudaMemcpyAsync(s.cst_poisson_pc, pc.data(), sizeof(double) * rows, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  } else {
    CompSum t;
    for (const MfitPrepBlock& b : res) { t.add(b.sum); t.c += b.err; }
    s.cst_poisson = t.value();
  }
  s.poisson_ready = true;
  return 0;
}

// compile the data set's likelihood lambda together with its model (NVRTC) and build the log-factorial table if it reads one
int ensure_user_term(mfit_ctx* ctx, Dataset& s, int d) {
  CU(cudaSetDevice(ctx->device));
  g_alloc_stream = ctx->stream;
  std::string code, e;
  int n_err = 0;
  int rc = mfit_lower_to_string(s.ll_text.c_str(), nullptr, 0, 2, code, e, &n_err);
  if (rc) return fail(ctx, rc, "%s", e.c_str());
  if (n_err > s.n_err_head)
    return fail(ctx, MFIT_ERR_ARG, "dataset %d: the likelihood lambda reads element %d of an error list of %d", d, n_err - 1, s.n_err_head);
  s.term_narr = code.find("NARR = 4") != std::string::npos ? 4 : 2;
  const std::string type = s.family == MFIT_FAM_SEXPR ? std::string("MfitUser") : mfit_builtin_model_type(s.family, s.degree);
  if (type.empty()) return fail(ctx, MFIT_ERR_STATE, "dataset %d has no model (:function)", d);
  std::vector<char> cubin;
  rc = mfit_nvrtc_compile(mfit_user_term_source(s.family == MFIT_FAM_SEXPR ? s.model_code : std::string(), type, code, s.err_head,
                                                s.n_err_head), cubin, e);
  if (rc) return fail(ctx, rc, "%s", e.c_str());
  dataset_free_term(s);
  CU(cudaLibraryLoadData(&s.lib_term, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  CU(cudaLibraryGetKernel(&s.k_term, s.lib_term, "mfit_user_eval_term"));
  if (s.term_narr == 4) {  // log y_i! per point
    if (!s.poisson_ok)
      return fail(ctx, MFIT_ERR_ARG, "dataset %d: log-factorial / log-poisson need non-negative integer-valued y (:379-380)", d);
    std::vector<float> memo;
    log_factorial_ref((long long)s.ymax, memo);
    std::vector<double> table(memo.begin(), memo.end());
    if (table.empty()) table.push_back(0.0);
    double* tab_dev = nullptr;
    CU(cudaMalloc(&tab_dev, sizeof(double) * table.size()));
    CU(cudaMalloc(&s.lf, sizeof(double) * (size_t)s.n_pad));
    cudaError_t ce = cudaMemcpyAsync(tab_dev, table.data(), sizeof(double) * table.size(), cudaMemcpyHostToDevice, ctx->stream);
    if (ce == cudaSuccess) {
      mfit_launch_logfact_fill(s.y, tab_dev, s.lf, s.n, s.n_pad, ctx->stream);
      ctx->launches++;
      ce = cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(tab_dev);
    if (ce != cudaSuccess) return fail(ctx, MFIT_ERR_CUDA, "log-factorial table: %s", cudaGetErrorString(ce));
  }
  return 0;
}

int check_problem_ready(mfit_ctx* ctx) {
  if (ctx->D <= 0) return fail(ctx, MFIT_ERR_STATE, "mfit_problem_define has not been called");
  for (int d = 0; d < ctx->D; d++) {
    Dataset& s = ctx->ds[d];
    if (!s.owned) continue;
    if (s.n < 0) return fail(ctx, MFIT_ERR_STATE, "dataset %d has no data", d);
    if (s.family < 0) return fail(ctx, MFIT_ERR_STATE, "dataset %d has no model (:function)", d);
    if (s.ll_kind == MFIT_LL_POISSON && !s.poisson_ok)
      return fail(ctx, MFIT_ERR_ARG, "dataset %d: log-poisson needs non-negative integer-valued y (log-factorial :379-380)", d);
    if (s.ll_kind == MFIT_LL_USER) {
      if (s.W_pc > 0) return fail(ctx, MFIT_ERR_ARG, "dataset %d: a likelihood lambda is not supported for walker sets with per-walker data", d);
      if (!s.k_term) {
        int rc = ensure_user_term(ctx, s, d);
        if (rc) return rc;
      }
    }
    if (s.ll_kind == MFIT_LL_POISSON && !s.poisson_ready) {
      int rc = ensure_poisson_const(ctx, s);
      if (rc) return rc;
    }
    // the reference would signal a type error inside the model (a sequence used as a number, or the reverse)
    if (s.x_dims > 1 && s.model_xdims == 0)
      return fail(ctx, MFIT_ERR_ARG, "dataset %d: x has %d components but the model takes x as a number", d, s.x_dims);
    if (s.model_xdims > 0 && s.W_pc > 0)
      return fail(ctx, MFIT_ERR_ARG, "dataset %d: walker sets with per-walker data take a one-dimensional x", d);
    if (s.model_xdims > s.x_dims || (s.model_xdims > 0 && !s.x_is_sequence))
      return fail(ctx, MFIT_ERR_ARG, "dataset %d: the model reads %d components of x, the data have %s%d", d, s.model_xdims,
                  s.x_is_sequence ? "" : "a plain number, not a sequence of ", s.x_dims);
  }
  return 0;
}

// the arguments one data set contributes to a likelihood launch
void fill_eval_args(MfitEvalArgs& a, const Dataset& s, const Workspace& w, int d, const double* theta, const unsigned char* skip,
                    int cb0, int n_cb) {
  memset(&a, 0, sizeof a);
  a.x = s.x; a.y = s.y; a.yw = s.yw; a.w = s.w;
  a.c = s.ll_kind == MFIT_LL_USER ? s.lf : s.c;   // a user term reads the log-factorial table in that slot
  a.n = s.n;
  a.xd_stride = s.n_pad;
  a.xmin = s.xmin;
  a.xmax = s.xmax;
  a.tiles_per_seg = s.tiles_per_seg;
  a.n_seg = s.n_seg;
  a.n_cb = n_cb >= 0 ? n_cb : (int)(w.stride / MFIT_CB);
  a.cb0 = cb0;
  a.W = (int)w.W;
  a.stride = w.stride;
  a.theta = theta;
  a.skip = skip;
  a.partial = w.partial[d];
  for (int k = 0; k < MFIT_MAXK; k++) a.idx[k] = s.idx[k];
  a.degree = s.degree;
  a.diag = 0;
}

int ensure_sched(mfit_ctx* ctx) {
  if (!ctx->sched_dev) {
    CU(cudaMalloc(&ctx->sched_dev, 2 * mfit_ctx::kMaxGroups * sizeof(int)));
    CU(cudaMemsetAsync(ctx->sched_dev, 0, 2 * mfit_ctx::kMaxGroups * sizeof(int), ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

// walker set with per-walker data: one warp per chain, no staging, no work-item queue
int launch_eval_per_walker(mfit_ctx* ctx, Dataset& s, const Workspace& w, int d, const double* theta, const unsigned char* skip,
                           int cb0, int n_cb, cudaStream_t st) {
  MfitEvalArgs a;
  fill_eval_args(a, s, w, d, theta, skip, cb0, n_cb);
  a.x_stride = s.x_shared ? 0 : s.n_pad;
  a.pc_stride = s.n_pad;
  a.w_stride = s.sigma_shared ? 0 : s.n_pad;
  // shared x / sigma vectors are staged in shared memory (mfit_eval_perchain_body)
  const bool staged = a.x_stride == 0 && a.w_stride == 0 && a.n <= MFIT_PC_STAGE_MAX;
  const size_t n_pad32 = (size_t)((a.n + 31) & ~31LL);
  size_t smem_pc = staged ? sizeof(double) * 3 * n_pad32 : 0;
  // ... and the chains' own y rows go through a per-warp TMA ring when two CTAs per SM still fit with at least two
  // stages (mfit_perchain_ring): rows of up to ~800 points; longer ones are read straight into registers
  a.ring_stages = 0;
  { const char* dg = getenv("MFIT_PC_DIAG"); a.diag = dg ? atoi(dg) : 0; }
  // Rows of up to 512 points (the reference's spectra have a few hundred, nv-specific.lisp:50-66), built-in family, no
  // cutoff constants: the register form (mfit_reg_row) -- four-warp CTAs, three per SM, every lane keeps the x and 1/sigma
  // of its points in registers across all the chains its warp evaluates.
  int reg_form = 0;       // 0 = no, 2 = rows up to 256 points, 3 = up to 512
  int threads = MFIT_PC_THREADS;
  const bool ring_ok = (long long)ctx->P * w.stride < (1LL << 31);   // mfit_perchain_ring forms idx[k] * stride in 32 bits
  if (ring_ok && staged && a.n > 0 && a.n <= 512 && s.family != MFIT_FAM_SEXPR && s.ll_kind != MFIT_LL_NORMAL_CUTOFF && s.ll_kind != MFIT_LL_USER &&
      (s.sigma_uniform || s.ll_kind == MFIT_LL_POISSON) && getenv("MFIT_NO_PC_RING") == nullptr && getenv("MFIT_NO_PC_REG") == nullptr) {
    reg_form = a.n <= 256 ? 2 : 3;
    threads = MFIT_PCR_THREADS;
    const size_t warps = MFIT_PCR_THREADS / 32, slot = reg_form == 2 ? 256 : 512,
                 budget = (MFIT_PCR_CTAS >= 4 ? 55 : 74) * 1024;  // MFIT_PCR_CTAS CTAs per SM
    for (int S = 4; S >= 2 && a.ring_stages == 0; S--) {
      const size_t need = smem_pc + warps * S * slot * sizeof(double) + warps * S * sizeof(unsigned long long) +
                          warps * MFIT_PCR_BATCH * 33 * sizeof(double);   // ring, its mbarriers, the parked per-lane sums
      if (need <= budget) {
        a.ring_stages = S;
        smem_pc = need;
      }
    }
    if (a.ring_stages == 0) { reg_form = 0; threads = MFIT_PC_THREADS; }
  }
  if (!reg_form && ring_ok && staged && a.n > 0 && getenv("MFIT_NO_PC_RING") == nullptr) {
    const char* bk = getenv("MFIT_PC_RING_BUDGET_KB");  // experiments: shared memory one CTA may take (default: two CTAs per SM)
    const size_t warps = MFIT_PC_THREADS / 32, budget = (bk ? (size_t)atoi(bk) : 110) * 1024;
    for (int S = 3; S >= 2 && a.ring_stages == 0; S--) {
      const size_t need = smem_pc + warps * S * n_pad32 * sizeof(double) + warps * S * sizeof(unsigned long long);
      if (need <= budget) {
        a.ring_stages = S;
        smem_pc = need;
      }
    }
  }
  const void* fpc = s.family == MFIT_FAM_SEXPR ? (const void*)s.k_user_pc[s.ll_kind + (staged ? 3 : 0)]
                                              : mfit_builtin_perchain_kernel(s.family, s.degree, s.ll_kind, reg_form ? reg_form : (int)staged);
  if (!fpc) return fail(ctx, MFIT_ERR_STATE, "no per-walker kernel for dataset %d (family %d degree %d)", d, s.family, s.degree);
  // built-in kernels are the same symbols for every context of the process: their attributes are set once per device
  std::unique_lock<std::mutex> shared_lock(ctx->shared->mu, std::defer_lock);
  const bool user_pc = s.family == MFIT_FAM_SEXPR;
  if (!user_pc) shared_lock.lock();
  std::map<const void*, int>& resident = user_pc ? ctx->resident : ctx->shared->resident;
  auto itp = resident.find(fpc);
  if (itp == resident.end()) {
    const size_t smem_max = 112 * 1024;
    CU(cudaFuncSetAttribute(fpc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    CU(cudaFuncSetAttribute(fpc, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0, sms = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fpc, threads, smem_pc));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    if (per_sm < 1) per_sm = 1;
    itp = resident.emplace(fpc, per_sm * sms).first;
  }
  const int resident_ctas = itp->second;
  if (!user_pc) shared_lock.unlock();
  const long long chains = (long long)a.n_cb * MFIT_CB;
  const long long need = (chains + threads / 32 - 1) / (threads / 32);
  const unsigned gridp = (unsigned)(need < resident_ctas ? need : resident_ctas);
  void* argsp[] = {&a};
  CU(cudaLaunchKernel(fpc, dim3(gridp > 0 ? gridp : 1), dim3(threads), argsp, smem_pc, st));
  ctx->launches++;
  return 0;
}

const void* shared_data_kernel(const Dataset& s) {
  if (s.ll_kind == MFIT_LL_USER) return (const void*)s.k_term;
  return s.family == MFIT_FAM_SEXPR ? (const void*)s.k_user[s.ll_kind] : mfit_builtin_eval_kernel(s.family, s.degree, s.ll_kind);
}

// One launch for the data sets ds[0..n) (same kernel = same model family / lowered lambda and likelihood kind): their
// work items share one queue.
int launch_eval_fused(mfit_ctx* ctx, const int* ds, int n, const Workspace& w, const double* theta, const unsigned char* skip,
                      int group, int cb0, int n_cb, cudaStream_t st) {
  int rc = ensure_sched(ctx);
  if (rc) return rc;
  MfitMultiArgs ma;
  memset(&ma, 0, sizeof ma);
  ma.n_ds = n;
  ma.sched = ctx->sched_dev + 2 * group;
  ma.mode_tiles = ctx->timing ? ctx->mode_dev : nullptr;
  long long items = 0;
  for (int k = 0; k < n; k++) {
    fill_eval_args(ma.d[k], ctx->ds[ds[k]], w, ds[k], theta, skip, cb0, n_cb);
    ma.item_start[k] = (int)items;
    items += (long long)ma.d[k].n_cb * (MFIT_CB / MFIT_WB) * ma.d[k].n_seg;
    if (items > 2000000000LL) return fail(ctx, MFIT_ERR_ARG, "too many work items in one launch (%lld)", items);
  }
  ma.item_start[n] = (int)items;
  if (items == 0) return 0;
  Dataset& s = ctx->ds[ds[0]];
  const void* fn = shared_data_kernel(s);
  if (!fn) return fail(ctx, MFIT_ERR_STATE, "no kernel for dataset %d (family %d degree %d)", ds[0], s.family, s.degree);
  std::unique_lock<std::mutex> shared_lock(ctx->shared->mu, std::defer_lock);
  const bool user_k = s.family == MFIT_FAM_SEXPR || s.ll_kind == MFIT_LL_USER;
  if (!user_k) shared_lock.lock();
  std::map<const void*, EvalShape>& shapes = user_k ? ctx->shape : ctx->shared->shape;
  auto it = shapes.find(fn);
  if (it == shapes.end()) {
    // Persistent CTAs of MFIT_WARPS independent warps, MFIT_MIN_CTAS per SM, a three-stage TMA ring per warp.  The
    // shape is fixed once per kernel; when the stages of a wide data set (many x components, the cutoff constants) do
    // not fit next to the math tables, first a shallower ring, then fewer warps.
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, fn));
    int sms = 0, max_optin = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    CU(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    const long long room = (long long)max_optin - (long long)fa.sharedSizeBytes;
    const int xd = s.model_xdims > 0 ? s.model_xdims : 1;
    // arrays per stage as mfit_eval_smem_bytes counts them by likelihood kind: a user term reads 2 (like poisson) or 4 (like cutoff)
    const int ll_smem = s.ll_kind == MFIT_LL_USER ? (s.term_narr == 4 ? MFIT_LL_NORMAL_CUTOFF : MFIT_LL_POISSON) : s.ll_kind;
    EvalShape sh;
    sh.warps = 0;
    int max_warps = MFIT_WARPS;
    if (fa.maxThreadsPerBlock / 32 < max_warps) max_warps = fa.maxThreadsPerBlock / 32;
    if (const char* e = getenv("MFIT_EVAL_WARPS")) { if (atoi(e) > 0 && atoi(e) < max_warps) max_warps = atoi(e); }  // tuning aid
    int want_stages = MFIT_STAGES;
    if (const char* e = getenv("MFIT_EVAL_STAGES")) { if (atoi(e) > 0 && atoi(e) <= 4) want_stages = atoi(e); }   // tuning aid
    // all MFIT_MIN_CTAS CTAs of an SM must fit: each takes its static tables + the rings (+ 1 KB the system reserves per CTA)
    int smem_sm = 0;
    CU(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, ctx->device));
    const long long per_cta = (long long)smem_sm / MFIT_MIN_CTAS - (long long)fa.sharedSizeBytes - 1024;
    const long long fit = per_cta < room ? per_cta : room;
    for (int w = max_warps; w >= 1 && sh.warps == 0; w = w > 4 ? w - 4 : w - 1)
      for (int st = want_stages; st >= 2 && sh.warps == 0; st--)
        if ((long long)mfit_eval_smem_bytes(ll_smem, xd, w, st) <= fit) { sh.warps = w; sh.stages = st; }
    if (sh.warps == 0)  // fewer CTAs per SM rather than none
      for (int w = max_warps; w >= 1 && sh.warps == 0; w--)
        if ((long long)mfit_eval_smem_bytes(ll_smem, xd, w, 2) <= room) { sh.warps = w; sh.stages = 2; }
    if (sh.warps == 0) return fail(ctx, MFIT_ERR_ARG, "dataset %d: a tile of its %d arrays does not fit in shared memory", ds[0], xd + 3);
    sh.smem = mfit_eval_smem_bytes(ll_smem, xd, sh.warps, sh.stages);
    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem));
    CU(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, sh.warps * 32, sh.smem));
    if (per_sm < 1) return fail(ctx, MFIT_ERR_CUDA, "likelihood kernel of dataset %d cannot be resident (%d warps, %zu bytes)", ds[0], sh.warps, sh.smem);
    // the kernel's __launch_bounds__ say how many CTAs per SM its registers are allotted for (MFIT_MIN_CTAS)
    if (per_sm > MFIT_MIN_CTAS) per_sm = MFIT_MIN_CTAS;
    sh.ctas = per_sm * sms;
    it = shapes.emplace(fn, sh).first;
  }
  const EvalShape sh = it->second;
  if (!user_k) shared_lock.unlock();
  ma.stages = sh.stages;
  void* args[] = {&ma};
  const long long need = (items + sh.warps - 1) / sh.warps;
  const unsigned grid = (unsigned)(need < sh.ctas ? need : sh.ctas);
  CU(cudaLaunchKernel(fn, dim3(grid), dim3(sh.warps * 32), args, sh.smem, st));
  ctx->launches++;
  return 0;
}

// all owned datasets: per-segment partial sums (+ user priors).  Built-in priors are written by the chain kernels.
int eval_all(mfit_ctx* ctx, Workspace& w, const MfitChains& view, const double* theta, const unsigned char* skip,
             int group = 0, int cb0 = 0, int n_cb = -1, cudaStream_t st = nullptr) {
  if (!st) st = ctx->stream;
  const bool timed = ctx->timing && ctx->ev_used < ctx->ev.size();
  if (timed) CU(cudaEventRecord(ctx->ev[ctx->ev_used].a, st));
  int n_eval = 0;
  for (int d = 0; d < ctx->D;) {
    Dataset& s = ctx->ds[d];
    if (!s.owned || s.n == 0 || s.n_seg == 0) { d++; continue; }
    if (s.W_pc > 0) {
      int rc = launch_eval_per_walker(ctx, s, w, d, theta, skip, cb0, n_cb, st);
      if (rc) return rc;
      n_eval++;
      d++;
      continue;
    }
    // batch the following data sets that run the same kernel (global fits: the same family on every data set)
    int batch[MFIT_MAX_FUSED], nb = 0;
    const void* fn = shared_data_kernel(s);
    int e = d;
    for (; e < ctx->D && nb < MFIT_MAX_FUSED; e++) {
      Dataset& t = ctx->ds[e];
      if (!t.owned || t.n == 0 || t.n_seg == 0) continue;
      if (t.W_pc > 0 || shared_data_kernel(t) != fn) break;
      batch[nb++] = e;
    }
    int rc = launch_eval_fused(ctx, batch, nb, w, theta, skip, group, cb0, n_cb, st);
    if (rc) return rc;
    n_eval++;
    d = e;
  }
  if (timed) {
    CU(cudaEventRecord(ctx->ev[ctx->ev_used].b, st));
    ctx->ev_used++;
    ctx->eval_launches += n_eval;
  }
  for (int d = 0; d < ctx->D; d++) {
    Dataset& s = ctx->ds[d];
    if (!s.owned || s.prior_kind != 2) continue;
    double* out = view.xch ? const_cast<double*>(view.xch) + (size_t)view.xpar * view.xpar_stride +
                                 ((size_t)(s.owner * view.xslots + s.xslot) * 2 + 1) * view.stride
                           : view.lpd + (size_t)d * view.stride;
    // the user prior kernel indexes chains from 0: a group passes pointers advanced to its first chain
    const long long w_lo = (long long)cb0 * MFIT_CB;
    long long w_hi = n_cb >= 0 ? w_lo + (long long)n_cb * MFIT_CB : w.W;
    if (w_hi > w.W) w_hi = w.W;
    long long stride = w.stride, Wg = w_hi - w_lo;
    if (Wg <= 0) continue;
    const double* th_g = theta + w_lo;
    const unsigned char* skip_g = skip ? skip + w_lo : nullptr;
    double* out_g = out + w_lo;
    void* args[] = {(void*)&th_g, &stride, &Wg, (void*)&skip_g, &out_g};
    CU(cudaLaunchKernel((const void*)s.k_prior, dim3((unsigned)((Wg + 127) / 128)), dim3(128), args, 0, st));
    ctx->launches++;
  }
  return 0;
}

// the view's exchange fields follow the context's exchange counter (peer exchange only)
void sync_epoch(mfit_ctx* ctx, MfitChains& v) {
  if (v.xranks > 0) {
    v.xepoch = ctx->peers.epoch;
    v.xpar = (int)(ctx->peers.epoch & 1);
  }
}

// After k_finalize: make the other ranks' per-data-set terms available to the kernel enqueued next.  Host-side collective
// (NCCL all-gather of every rank's rows), or -- peer exchange -- nothing to launch: k_finalize has pushed this rank's rows
// into the peers' buffers, k_post waits for theirs itself, any other consumer is preceded by a wait kernel.
int exchange(mfit_ctx* ctx, Workspace& w, const MfitChains* v = nullptr, bool consumer_waits = false) {
  if (!ctx->sharded) return 0;
  if (w.peer) {
    if (!consumer_waits && v) { mfit_launch_peer_wait(*v, ctx->stream); ctx->launches++; }
    return 0;
  }
  if (!ctx->comm) return fail(ctx, MFIT_ERR_STATE, "data sets are sharded but neither NCCL nor peers are attached");
  const size_t per_rank = sizeof(double) * (size_t)w.xslots * 2 * w.stride;
  std::string e;
  int rc = mfit_nccl_allgather_bytes(ctx->comm, (const char*)w.xch + per_rank * ctx->rank, w.xch, per_rank, ctx->stream, e);
  if (rc) return fail(ctx, rc, "%s", e.c_str());
  return 0;
}

MfitChains view_of(mfit_ctx* ctx, Workspace& w) {
  MfitChains v;
  memset(&v, 0, sizeof v);
  v.W = (int)w.W;
  v.P = ctx->P;
  v.D = ctx->D;
  v.stride = w.stride;
  v.w0 = 0;
  v.w1 = w.W;
  v.max_seg = 0;
  for (const Dataset& s : ctx->ds)
    if (s.owned && s.n_seg > v.max_seg) v.max_seg = s.n_seg;
  v.lld = w.lld;
  v.lpd = w.lpd;
  v.xch = w.xch;
  v.xslots = w.xslots;
  if (w.peer) {
    const Peers& p = ctx->peers;
    v.xpar_stride = (long long)p.n * w.xslots * 2 * w.stride;
    v.xrank = p.rank;
    v.xranks = p.n;
    for (int q = 0; q < p.n; q++) { v.xpeer[q] = p.xpeer[q]; v.xflag_peer[q] = p.flag_peer[q]; }
    v.xflag = p.flags;
    v.xdone = p.xdone;
    v.xerr = p.xerr;
    sync_epoch(ctx, v);
  }
  v.ds = w.ds_dev;
  return v;
}

void collect_timing(mfit_ctx* ctx) {
  for (size_t i = 0; i < ctx->ev_used; i++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev[i].a, ctx->ev[i].b) == cudaSuccess) ctx->eval_seconds += ms * 1e-3;
  }
  ctx->ev_used = 0;
}

// device -> host copy of a large block into caller memory that is probably pageable: the DMA goes into two pinned
// bounce buffers (full PCIe rate instead of the driver's pageable path) and the CPU copies chunk i out while chunk
// i + 1 is in flight.  Returns after everything has landed in dst.
constexpr size_t kStageBytes = (size_t)16 << 20;
// pageable destinations are usually untouched pages: the copy out of the bounce buffer is page-fault bound on one
// core, so it is split over a few threads
void copy_out(char* dst, const char* src, size_t bytes) {
  const int nt = bytes >= ((size_t)4 << 20) ? 4 : 1;
  if (nt == 1) { memcpy(dst, src, bytes); return; }
  std::thread th[3];
  const size_t per = (bytes / nt + 4095) & ~(size_t)4095;
  for (int t = 1; t < nt; t++) {
    const size_t lo = per * t, hi = lo + per < bytes ? lo + per : bytes;
    if (lo < hi) th[t - 1] = std::thread([=] { memcpy(dst + lo, src + lo, hi - lo); });
  }
  memcpy(dst, src, per < bytes ? per : bytes);
  for (auto& t : th) if (t.joinable()) t.join();
}
// MFIT_DEBUG_TIMING=1: wall-clock marks of the set-up path on stderr (each mark synchronises the stream first)
struct PhaseClock {
  bool on; double t0; cudaStream_t st;
  explicit PhaseClock(cudaStream_t s);
  void mark(const char* what);
};
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
PhaseClock::PhaseClock(cudaStream_t s) : on(getenv("MFIT_DEBUG_TIMING") != nullptr), t0(now_ms()), st(s) {}
void PhaseClock::mark(const char* what) {
  if (!on) return;
  cudaStreamSynchronize(st);
  const double t = now_ms();
  fprintf(stderr, "  [mfit] %-28s %8.3f ms\n", what, t - t0);
  t0 = t;
}
int d2h_staged(mfit_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes < ((size_t)512 << 10)) {
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
  std::lock_guard<std::mutex> stage_lock(ctx->shared->stage_mu);  // the bounce buffers belong to the device, not to the context
  for (int i = 0; i < 2; i++) {
    if (!ctx->shared->stage[i]) CU(cudaMallocHost(&ctx->shared->stage[i], kStageBytes));
    ctx->stage[i] = ctx->shared->stage[i];
  }
  if (bytes <= kStageBytes) {
    // A few MB (the trace of a short run): one DMA per 1 MB slice into the first bounce buffer, all enqueued at once, and the
    // CPU copies slice i out -- mostly page faults of the caller's fresh pages -- while the later ones are still in flight.
    const size_t slice = (size_t)1 << 20, n = (bytes + slice - 1) / slice;
    std::vector<cudaEvent_t>& evs = ctx->shared->slice_ev;
    while (evs.size() < n) {
      cudaEvent_t ev;
      CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      evs.push_back(ev);
    }
    for (size_t i = 0; i < n; i++) {
      const size_t len = i + 1 < n ? slice : bytes - i * slice;
      CU(cudaMemcpyAsync(ctx->stage[0] + i * slice, (const char*)src + i * slice, len, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaEventRecord(evs[i], ctx->stream));
    }
    for (size_t i = 0; i < n; i++) {
      const size_t len = i + 1 < n ? slice : bytes - i * slice;
      CU(cudaEventSynchronize(evs[i]));
      memcpy((char*)dst + i * slice, ctx->stage[0] + i * slice, len);
    }
    return 0;
  }
  const size_t n_chunks = (bytes + kStageBytes - 1) / kStageBytes;
  auto chunk_len = [&](size_t i) { return i + 1 < n_chunks ? kStageBytes : bytes - i * kStageBytes; };
  for (size_t i = 0; i <= n_chunks; i++) {
    if (i < n_chunks) {
      CU(cudaMemcpyAsync(ctx->stage[i & 1], (const char*)src + i * kStageBytes, chunk_len(i), cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaEventRecord(ctx->stage_ev[i & 1], ctx->stream));
    }
    if (i >= 1) {
      CU(cudaEventSynchronize(ctx->stage_ev[(i - 1) & 1]));
      copy_out((char*)dst + (i - 1) * kStageBytes, ctx->stage[(i - 1) & 1], chunk_len(i - 1));
    }
  }
  return 0;
}

template <class T>
int chain_alloc(mfit_ctx* ctx, T** out, size_t count, bool zero = true) {
  void* p = nullptr;
  CU(cudaMalloc(&p, sizeof(T) * (count ? count : 1)));
  if (zero) CU(cudaMemsetAsync(p, 0, sizeof(T) * (count ? count : 1), ctx->stream));
  ctx->chain_allocs.push_back(p);
  *out = (T*)p;
  return 0;
}

long long min_trace_cap(int P) {
  long long S = 10LL * (P > 50 ? P : 50);
  return (S > 1000 ? S : 1000) + 1;
}

int grow_trace(mfit_ctx* ctx, long long need_rows) {
  MfitChains& c = ctx->ch;
  if (need_rows <= c.cap || ctx->cap_fixed) return 0;
  size_t free_b = 0, total_b = 0;
  CU(cudaMemGetInfo(&free_b, &total_b));
  const size_t row_bytes = sizeof(double) * (size_t)(c.P + 1) * c.stride;
  long long afford = (long long)((free_b * 0.8) / row_bytes);
  long long new_cap = need_rows < afford ? need_rows : afford;
  if (new_cap <= c.cap) return 0;  // keep the ring: the oldest rows are overwritten, like :max-walker-length
  double *nt = nullptr, *np = nullptr;
  CU(cudaMalloc(&nt, sizeof(double) * (size_t)new_cap * c.P * c.stride));
  CU(cudaMalloc(&np, sizeof(double) * (size_t)new_cap * c.stride));
  mfit_launch_trace_regrow(c, nt, np, new_cap, ctx->stream);
  ctx->launches++;
  CU(cudaStreamSynchronize(ctx->stream));
  for (auto& p : ctx->chain_allocs) {
    if (p == c.trace_theta) { cudaFree(p); p = nt; }
    else if (p == c.trace_prob) { cudaFree(p); p = np; }
  }
  c.trace_theta = nt;
  c.trace_prob = np;
  c.cap = new_cap;
  return 0;
}

int upload_opt(mfit_ctx* ctx, const double* host, size_t count, double** dev, size_t* dev_bytes) {
  const size_t bytes = sizeof(double) * count;
  if (*dev_bytes < bytes) {
    free_dev(*dev);
    *dev = nullptr;
    *dev_bytes = 0;
    CU(cudaMalloc(dev, bytes));
    *dev_bytes = bytes;
  }
  CU(cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

// a peer-exchange wait that ran out of time (a rank died or fell out of step) left a mark: report it once
int peer_check(mfit_ctx* ctx) {
  if (!ctx->peers.ready) return 0;
  int err = 0;
  CU(cudaMemcpyAsync(&err, ctx->peers.xerr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (err) {
    CU(cudaMemsetAsync(ctx->peers.xerr, 0, sizeof(int), ctx->stream));
    return fail(ctx, MFIT_ERR_NCCL, "peer exchange timed out: a rank did not deliver its data sets' terms");
  }
  return 0;
}

// set-up of a run (:866-901): driver state per chain, first proposals
int run_begin(mfit_ctx* ctx, MfitRunCfg r, const double* l_shared_host, long long max_steps) {
  MfitChains& c = ctx->ch;
  int rc = check_problem_ready(ctx);
  if (rc) return rc;
  ctx->run_open = false;
  ctx->stop.store(0);
  if (max_steps < 0) max_steps = 0;
  rc = grow_trace(ctx, ctx->rows_upper + max_steps + 1);
  if (rc) return rc;
  const double* l_dev = nullptr;
  double* l_tmp = nullptr;
  if (l_shared_host && r.l_per_chain) {   // W x P x P, used by k_init_run only
    CU(cudaMalloc(&l_tmp, sizeof(double) * (size_t)c.W * c.P * c.P));
    CU(cudaMemcpyAsync(l_tmp, l_shared_host, sizeof(double) * (size_t)c.W * c.P * c.P, cudaMemcpyHostToDevice, ctx->stream));
    l_dev = l_tmp;
  } else if (l_shared_host) {
    if (!ctx->lshared_dev) CU(cudaMalloc(&ctx->lshared_dev, sizeof(double) * MFIT_MAX_P * MFIT_MAX_P));
    CU(cudaMemcpyAsync(ctx->lshared_dev, l_shared_host, sizeof(double) * c.P * c.P, cudaMemcpyHostToDevice, ctx->stream));
    l_dev = ctx->lshared_dev;
  }
  r.seed = ctx->seed;
  r.chain_base = ctx->chain_base;
  r.t = 0;
  sync_epoch(ctx, c);   // the first proposals' priors belong to the exchange of the first step
  mfit_launch_init_run(c, r, l_dev, ctx->stream);
  free_dev(l_tmp);                       // stream-ordered: after k_init_run has read it
  mfit_launch_pre(c, r, ctx->stream);    // proposals of chains that keep their l-matrix
  mfit_launch_reest(c, r, ctx->stream);  // chains resuming a long walk re-estimate first (:897-901)
  ctx->launches += 3;
  if (!ctx->active_dev) CU(cudaMalloc(&ctx->active_dev, sizeof(int) * 4));
  ctx->run_cfg = r;
  ctx->run_t = 0;
  ctx->run_max = max_steps;
  ctx->run_open = true;
  return 0;
}

// Small problems (the reference's own: one walker, a thousand points) are launch-bound on the three-launch step: when every
// data set is shared data with a built-in family and a built-in prior, and a step is only a few microseconds of
// arithmetic, the chains advance many steps per launch, one CTA per chain (chain_kernels.cu k_run_small).
bool small_path(mfit_ctx* ctx, MfitSmallArgs& a) {
  if (ctx->small_mode == 1 || ctx->sharded || ctx->timing || ctx->D > MFIT_SMALL_MAX_DS) return false;
  if (const char* e = getenv("MFIT_SMALL_PATH")) { if (atoi(e) == 0) return false; }
  long long n_total = 0;
  memset(&a, 0, sizeof a);
  a.n_ds = ctx->D;
  for (int d = 0; d < ctx->D; d++) {
    const Dataset& s = ctx->ds[d];
    if (!s.owned || s.W_pc > 0 || s.family == MFIT_FAM_SEXPR || s.family < 0 || s.ll_kind > MFIT_LL_POISSON || s.prior_kind == 2 ||
        s.x_dims != 1 || s.x_is_sequence)
      return false;
    MfitSmallDs& o = a.d[d];
    o.x = s.x; o.y = s.y; o.yw = s.yw; o.w = s.w; o.c = s.c;
    o.n = s.n;
    o.family = s.family; o.degree = s.degree; o.ll_kind = s.ll_kind;
    for (int k = 0; k < 16; k++) o.idx[k] = s.idx[k];
    n_total += s.n;
  }
  if (ctx->small_mode == 2) return true;
  // a CTA of 128 threads per chain: worth it while the large kernel's step (three launches, ~25 us) is mostly latency
  return n_total <= 4096 && (long long)ctx->ch.W * n_total <= (1LL << 24) && ctx->ch.W <= 8192;
}

// up to `budget` more MH steps of the open run; returns when they have completed on the device
int run_advance(mfit_ctx* ctx, long long budget, long long* steps_done, int* finished) {
  MfitChains& c = ctx->ch;
  if (!ctx->run_open) return fail(ctx, MFIT_ERR_STATE, "no open run");
  MfitRunCfg r = ctx->run_cfg;
  ctx->eval_seconds = 0.0;
  ctx->eval_launches = 0;
  ctx->ev_used = 0;
  const long long poll_every = 64;
  bool poll_pending[4] = {false, false, false, false};
  int poll_slot = 0, rc = 0;
  long long t = ctx->run_t;
  const long long t_end = (budget < 0 || t + budget > ctx->run_max) ? ctx->run_max : t + budget;
  bool all_done = false;
  // chain groups (see mfit_ctx::gstream): two when there are enough chains to fill the machine twice over, the data sets
  // are not sharded over ranks (the exchange is a whole-set collective) and kernel times are not being recorded
  // (an overlapped kernel's event interval includes its wait for SM slots)
  const long long n_cb_all = c.stride / MFIT_CB;
  int n_groups = (ctx->sharded || ctx->timing) ? 1 : n_cb_all >= 12 ? 3 : n_cb_all >= 8 ? 2 : 1;
  // walker sets with per-walker data: their likelihood kernel takes one pass over the rows, comparable to the chain kernels
  // of a group; two groups overlap them best (NVSET step 0.182 / 0.173 / 0.181 ms with one / two / three groups)
  for (const Dataset& s : ctx->ds)
    if (s.W_pc > 0 && n_groups > 2) n_groups = 2;
  if (const char* e = getenv("MFIT_GROUPS")) { if (atoi(e) > 0 && !ctx->timing) n_groups = atoi(e); }  // tuning aid
  if (ctx->groups_override > 0) n_groups = ctx->groups_override;
  if (n_groups > mfit_ctx::kMaxGroups) n_groups = mfit_ctx::kMaxGroups;
  if (n_groups > n_cb_all) n_groups = (int)n_cb_all;
  if (ctx->sharded || n_groups < 1) n_groups = 1;
  ctx->n_groups = n_groups;
  ctx->gstream[0] = ctx->stream;
  if (n_groups > 1) {
    if (!ctx->fork_ev) CU(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
    CU(cudaEventRecord(ctx->fork_ev, ctx->stream));  // everything enqueued so far (run set-up, earlier steps) comes first
    for (int g = 1; g < n_groups; g++) {
      if (!ctx->gstream[g]) {
        CU(cudaStreamCreateWithFlags(&ctx->gstream[g], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ctx->join_ev[g], cudaEventDisableTiming));
      }
      CU(cudaStreamWaitEvent(ctx->gstream[g], ctx->fork_ev, 0));
    }
  }
  auto join_groups = [&]() -> int {  // the context's stream waits for what the other groups have enqueued
    for (int g = 1; g < n_groups; g++) {
      CU(cudaEventRecord(ctx->join_ev[g], ctx->gstream[g]));
      CU(cudaStreamWaitEvent(ctx->stream, ctx->join_ev[g], 0));
    }
    return 0;
  };
  struct Group { int cb0, n_cb; long long w0, w1; cudaStream_t st; } grp[mfit_ctx::kMaxGroups];
  for (int g = 0; g < n_groups; g++) {
    const int lo = (int)(n_cb_all * g / n_groups), hi = (int)(n_cb_all * (g + 1) / n_groups);
    grp[g] = {lo, hi - lo, 0, 0, ctx->gstream[g]};
    grp[g].w0 = (long long)lo * MFIT_CB;
    grp[g].w1 = (long long)hi * MFIT_CB;
    if (grp[g].w1 > c.W) grp[g].w1 = c.W;
    if (grp[g].w0 > c.W) grp[g].w0 = c.W;
  }
  bool stop_seen = false;
  MfitSmallArgs small_args;
  if (small_path(ctx, small_args)) {
    // many steps per launch; between launches the host looks at the stop flag and (blocking, it is cheap here) at the chains left
    const long long chunk = 512;
    while (t < t_end && !all_done) {
      if (ctx->stop.load()) break;  // mfit-walker-estop, :904
      const long long k = t_end - t < chunk ? t_end - t : chunk;
      r.t = t;
      mfit_launch_run_small(c, r, small_args, k, ctx->stream);
      mfit_launch_count_active(c, ctx->active_dev, ctx->stream);
      ctx->launches += 2;
      CU(cudaMemcpyAsync(ctx->active_host, ctx->active_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      t += k;
      if (ctx->active_host[0] == 0) all_done = true;
    }
  } else
  for (; t < t_end && !all_done; t++) {
    // mfit-walker-estop, :904.  With the data sets sharded over ranks every step holds a collective, so all ranks must
    // leave the loop at the same t: the flag is then only looked at -- all-gathered -- at the poll boundaries below.
    if (!ctx->sharded && ctx->stop.load()) break;
    r.t = t;
    sync_epoch(ctx, c);
    for (int g = 0; g < n_groups; g++) {
      const Group& G = grp[g];
      if (G.w1 <= G.w0) continue;
      MfitChains cg = c;
      cg.w0 = G.w0;
      cg.w1 = G.w1;
      cg.group = g;
      rc = eval_all(ctx, ctx->ws, c, c.prop_theta, c.done, g, G.cb0, G.n_cb, G.st);
      if (rc) return rc;
      if (ctx->sharded) {
        mfit_launch_finalize(c, ctx->stream);   // peer exchange: also stores this rank's rows into the peers' buffers
        ctx->launches++;
        rc = exchange(ctx, ctx->ws, &c, true);  // k_post waits for the peers' rows itself
        if (rc) return rc;
      }
      mfit_launch_post(cg, r, true, G.st);  // accept/push/adapt of step t, proposal of step t+1
      MfitRunCfg rn = r;
      rn.t = t + 1;
      MfitChains cn = cg;
      if (cn.xpar_stride != 0) cn.xpar ^= 1;   // a re-estimated chain's next proposal: priors of the NEXT exchange
      mfit_launch_reest(cn, rn, G.st);
      ctx->launches += 2;
    }
    if (ctx->ws.peer) ctx->peers.epoch++;
    if (ctx->sharded && (t + 1) % poll_every == 0) {
      // Collective, deterministic stop test: every rank holds the same chains (so the same active count) but its own
      // stop flag and its own host timing.  (active, stop) of all ranks are all-gathered and read back with a blocking
      // copy at the same step on every rank; any stop request, or no chain left, ends the loop everywhere at t + 1.
      const int R = ctx->n_ranks;
      if (!ctx->xflag_dev) {
        CU(cudaMalloc(&ctx->xflag_dev, sizeof(int) * 2 * R));
        if (2 * R > 56) return fail(ctx, MFIT_ERR_ARG, "more than 28 ranks");
        ctx->xflag_host = ctx->active_host + 8;  // the context's pinned words: [0,4) active counts, [8, 8 + 2R) these flags
      }
      ctx->launches++;
      if (ctx->comm) {
        mfit_launch_count_active(c, ctx->xflag_dev + 2 * ctx->rank, ctx->stream);
        ctx->xflag_host[2 * ctx->rank + 1] = ctx->stop.load();
        CU(cudaMemcpyAsync(ctx->xflag_dev + 2 * ctx->rank + 1, ctx->xflag_host + 2 * ctx->rank + 1, sizeof(int),
                           cudaMemcpyHostToDevice, ctx->stream));
        std::string er;
        rc = mfit_nccl_allgather_bytes(ctx->comm, ctx->xflag_dev + 2 * ctx->rank, ctx->xflag_dev, 2 * sizeof(int), ctx->stream, er);
        if (rc) return fail(ctx, rc, "%s", er.c_str());
      } else {
        // peers without a communicator (all ranks in one process, mfit_multi_*): the active count is the same on every
        // rank by construction; a stop request is agreed on through the owner's shared word, TWO boundaries ahead: a peer's
        // host may already have passed its test of boundary b + 1 when rank 0 writes at boundary b (its blocking read of
        // boundary b needs only the GPU work rank 0 enqueued BEFORE this test), but it cannot have reached boundary b + 2,
        // whose predecessor's read waits for steps rank 0 enqueues after the write
        CU(cudaMemsetAsync(ctx->xflag_dev, 0xff, sizeof(int) * 2 * R, ctx->stream));   // "-1": other ranks neither done nor stopping
        mfit_launch_count_active(c, ctx->xflag_dev + 2 * ctx->rank, ctx->stream);
        const long long boundary = (t + 1) / poll_every;
        int stop_now = 0;
        if (ctx->stop_at) {
          long long at = ctx->stop_at->load();
          if (at < 0 && ctx->rank == 0 && ctx->stop.load()) {
            long long want = boundary + 2;
            ctx->stop_at->compare_exchange_strong(at, want);
            at = ctx->stop_at->load();
          }
          stop_now = at >= 0 && boundary >= at;
        }
        ctx->xflag_host[2 * ctx->rank + 1] = stop_now;
        CU(cudaMemcpyAsync(ctx->xflag_dev + 2 * ctx->rank + 1, ctx->xflag_host + 2 * ctx->rank + 1, sizeof(int),
                           cudaMemcpyHostToDevice, ctx->stream));
      }
      CU(cudaMemcpyAsync(ctx->xflag_host, ctx->xflag_dev, sizeof(int) * 2 * R, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      for (int k = 0; k < R; k++) {
        if (ctx->xflag_host[2 * k] == 0) all_done = true;
        if (ctx->xflag_host[2 * k + 1] > 0) stop_seen = true;
      }
      if (stop_seen) { t++; break; }
    } else if ((t + 1) % poll_every == 0) {
      // lagging poll: look at the answer of an earlier query if it has arrived; never block
      for (int s = 0; s < 4; s++)
        if (poll_pending[s] && cudaEventQuery(ctx->poll_ev[s]) == cudaSuccess) {
          poll_pending[s] = false;
          if (ctx->active_host[s] == 0) all_done = true;
        }
      if (!poll_pending[poll_slot]) {
        if ((rc = join_groups()) != 0) return rc;  // the count must see every group's step t
        mfit_launch_count_active(c, ctx->active_dev + poll_slot, ctx->stream);
        ctx->launches++;
        CU(cudaMemcpyAsync(ctx->active_host + poll_slot, ctx->active_dev + poll_slot, sizeof(int),
                           cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaEventRecord(ctx->poll_ev[poll_slot], ctx->stream));
        poll_pending[poll_slot] = true;
        poll_slot = (poll_slot + 1) & 3;
      }
    }
  }
  if ((rc = join_groups()) != 0) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaGetLastError());
  if ((rc = peer_check(ctx)) != 0) return rc;
  collect_timing(ctx);
  ctx->rows_upper += t - ctx->run_t;
  if (steps_done) *steps_done = t - ctx->run_t;
  ctx->run_t = t;
  const bool fin = all_done || t >= ctx->run_max || (ctx->sharded ? stop_seen : ctx->stop.load() != 0);
  if (fin) ctx->run_open = false;
  if (finished) *finished = fin ? 1 : 0;
  return 0;
}

int run_loop(mfit_ctx* ctx, MfitRunCfg r, const double* l_shared_host, long long max_steps, long long* steps_launched) {
  int rc = run_begin(ctx, r, l_shared_host, max_steps);
  if (rc) return rc;
  long long done = 0;
  rc = run_advance(ctx, -1, &done, nullptr);
  ctx->run_open = false;
  if (steps_launched) *steps_launched = done;
  return rc;
}

}  // namespace

// multi.cpp: the word through which the contexts of one process agree on the step at which a stop request takes effect
void mfit_internal_stop_word(mfit_ctx* ctx, std::atomic<long long>* word) { if (ctx) ctx->stop_at = word; }

// =================================================================================== C ABI
extern "C" {

int mfit_abi_version(void) { return MFIT_ABI_VERSION; }

int mfit_ctx_create(int device, void* stream, mfit_ctx** out) {
  if (!out) return fail(nullptr, MFIT_ERR_ARG, "out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, MFIT_ERR_CUDA, "no CUDA device (%s); this library has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= n) return fail(nullptr, MFIT_ERR_ARG, "device %d out of range (have %d)", device, n);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, MFIT_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  DeviceShared* sh = device_shared(device);
  StreamSet* set = nullptr;
  {
    std::lock_guard<std::mutex> lock(sh->mu);
    if (!sh->probed) {
      cudaDeviceGetAttribute(&sh->cc_major, cudaDevAttrComputeCapabilityMajor, device);
      cudaDeviceGetAttribute(&sh->cc_minor, cudaDevAttrComputeCapabilityMinor, device);
      cudaDeviceGetAttribute(&sh->sms, cudaDevAttrMultiProcessorCount, device);
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = 16ull << 30;  // freed blocks stay cached up to 16 GB; beyond that they go back to the driver
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      sh->probed = true;
    }
    if (sh->cc_major != 10)
      return fail(nullptr, MFIT_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, sh->cc_major,
                  sh->cc_minor);
    if (!sh->free_sets.empty()) { set = sh->free_sets.back(); sh->free_sets.pop_back(); }
  }
  if (!set) {  // streams of the chain groups, their fork / join events, the poll and bounce-buffer events, a pinned word or two
    set = new StreamSet();
    bool ok = true;
    for (auto& g : set->g) ok = ok && cudaStreamCreateWithFlags(&g, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&set->fork_ev, cudaEventDisableTiming) == cudaSuccess;
    for (auto& ev : set->join_ev) ok = ok && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
    for (auto& ev : set->poll_ev) ok = ok && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
    for (auto& ev : set->stage_ev) ok = ok && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMallocHost(&set->pinned_ints, 64 * sizeof(int)) == cudaSuccess;
    if (!ok) return fail(nullptr, MFIT_ERR_CUDA, "stream / event set: %s", cudaGetErrorString(cudaGetLastError()));
  }
  mfit_ctx* c = new mfit_ctx();
  c->device = device;
  c->shared = sh;
  c->set = set;
  for (int g = 1; g < mfit_ctx::kMaxGroups; g++) { c->gstream[g] = set->g[g - 1]; c->join_ev[g] = set->join_ev[g - 1]; }
  c->fork_ev = set->fork_ev;
  for (int k = 0; k < 4; k++) c->poll_ev[k] = set->poll_ev[k];
  for (int k = 0; k < 2; k++) c->stage_ev[k] = set->stage_ev[k];
  c->active_host = set->pinned_ints;
  if (stream) c->stream = (cudaStream_t)stream;
  else {
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return fail(nullptr, MFIT_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    c->own_stream = true;
  }
  *out = c;
  return 0;
}

void mfit_ctx_destroy(mfit_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  g_alloc_stream = ctx->stream;
  cudaStreamSynchronize(ctx->stream);
  chains_free(ctx);
  for (auto& d : ctx->ds) { dataset_free_data(d); dataset_free_prior(d); dataset_free_model(d); }
  free_dev(ctx->temps_dev); free_dev(ctx->z_dev); free_dev(ctx->u_dev); free_dev(ctx->lshared_dev);
  free_dev(ctx->active_dev); free_dev(ctx->flush_buf); free_dev(ctx->sched_dev); free_dev(ctx->mode_dev);
  free_dev(ctx->xflag_dev);
  for (void* m : ctx->peers.opened) if (m) cudaIpcCloseMemHandle(m);
  if (ctx->peers.base) (cudaFree)(ctx->peers.base);
  for (auto& p : ctx->ev) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  cudaStreamSynchronize(ctx->stream);  // the stream-ordered frees above
  for (int g = 1; g < mfit_ctx::kMaxGroups; g++) cudaStreamSynchronize(ctx->gstream[g]);
  {  // the stream / event set goes back to the device's pool
    std::lock_guard<std::mutex> lock(ctx->shared->mu);
    ctx->shared->free_sets.push_back(ctx->set);
  }
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* mfit_last_error(mfit_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int mfit_synchronize(mfit_ctx* ctx) {
  if (!ctx) return MFIT_ERR_ARG;
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int64_t mfit_launch_count(mfit_ctx* ctx) { return ctx ? ctx->launches : 0; }
int mfit_stream_handle(mfit_ctx* ctx, void** stream_out) {
  if (!ctx || !stream_out) return MFIT_ERR_ARG;
  *stream_out = (void*)ctx->stream;
  return 0;
}

int mfit_problem_define(mfit_ctx* ctx, int n_datasets, int n_params) {
  if (!ctx) return MFIT_ERR_ARG;
  if (n_datasets < 1 || n_params < 1 || n_params > MFIT_MAX_P)
    return fail(ctx, MFIT_ERR_ARG, "need 1 <= datasets and 1 <= params <= %d", MFIT_MAX_P);
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  chains_free(ctx);
  for (auto& d : ctx->ds) { dataset_free_data(d); dataset_free_prior(d); dataset_free_model(d); }
  ctx->ds.assign(n_datasets, Dataset());
  for (auto& d : ctx->ds) memset(d.idx, 0, sizeof d.idx);
  ctx->D = n_datasets;
  ctx->P = n_params;
  ctx->sharded = false;
  return 0;
}

#define NEED_DS(d)                                                                       \
  if (!ctx) return MFIT_ERR_ARG;                                                          \
  g_alloc_stream = ctx->stream;                                                           \
  if (ctx->D <= 0) return fail(ctx, MFIT_ERR_STATE, "mfit_problem_define has not been called"); \
  if ((d) < 0 || (d) >= ctx->D) return fail(ctx, MFIT_ERR_ARG, "dataset index %d out of range", (d))

int mfit_dataset_upload(mfit_ctx* ctx, int d, const double* x, const double* y, const double* sigma, int64_t n) {
  return mfit_dataset_upload_xd(ctx, d, x, 0, y, sigma, n);
}

// host -> device rows of n doubles laid down n_pad apart, the padding of every row zeroed
static int upload_rows(mfit_ctx* ctx, double* dst, const double* src, size_t n, size_t n_pad, size_t rows) {
  if (n > 0 && rows > 0)
    CU(cudaMemcpy2DAsync(dst, n_pad * sizeof(double), src, n * sizeof(double), n * sizeof(double), rows, cudaMemcpyHostToDevice,
                         ctx->stream));
  if (n_pad > n && rows > 0)
    CU(cudaMemset2DAsync(dst + n, n_pad * sizeof(double), 0, (n_pad - n) * sizeof(double), rows, ctx->stream));
  return 0;
}

// the upload pass (prep_kernels.cu) over a data set already in device memory (sigma in s.c), and its small results
static int prepare_dataset(mfit_ctx* ctx, Dataset& s, size_t rows, size_t x_rows, size_t sig_rows, bool have_sigma,
                           std::vector<MfitPrepBlock>& res) {
  const int blocks = rows > 1 ? (int)(rows < 2048 ? rows : 2048)
                              : (int)std::max<long long>(1, std::min<long long>(592, (s.n_pad + MFIT_PREP_THREADS - 1) / MFIT_PREP_THREADS));
  const size_t n_res = rows > 1 ? rows : (size_t)blocks;
  MfitPrepBlock* dev = nullptr;
  CU(cudaMalloc(&dev, sizeof(MfitPrepBlock) * n_res));
  mfit_launch_prepare(s.x, s.y, s.yw, s.w, s.c, s.n, s.n_pad, (long long)rows, (long long)x_rows, (long long)sig_rows,
                      have_sigma ? 1 : 0, dev, blocks, ctx->stream);
  ctx->launches += (rows > 1 && sig_rows == 1) ? 2 : 1;
  res.resize(n_res);
  cudaError_t e = cudaMemcpyAsync(res.data(), dev, sizeof(MfitPrepBlock) * n_res, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // also: the caller's arrays are free to go
  cudaFree(dev);
  if (e != cudaSuccess) return fail(ctx, MFIT_ERR_CUDA, "dataset upload: %s", cudaGetErrorString(e));
  return 0;
}

// x_dims = 0: every x is a number; x_dims >= 1: every x is a sequence of x_dims numbers, component k of point i at
// x[k * n + i]
int mfit_dataset_upload_xd(mfit_ctx* ctx, int d, const double* x, int x_dims, const double* y, const double* sigma, int64_t n) {
  NEED_DS(d);
  if (n < 0 || (n > 0 && (!x || !y))) return fail(ctx, MFIT_ERR_ARG, "dataset %d: bad x/y/n", d);
  if (x_dims < 0 || x_dims > MFIT_MAX_XDIMS) return fail(ctx, MFIT_ERR_ARG, "dataset %d: x_dims must be 0 .. %d", d, MFIT_MAX_XDIMS);
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  Dataset& s = ctx->ds[d];
  dataset_free_data(s);
  s.n = n;
  s.n_pad = round_up(n > 0 ? n : 1, MFIT_TILE);
  s.x_is_sequence = x_dims > 0;
  s.x_dims = x_dims > 0 ? x_dims : 1;
  const size_t np = (size_t)s.n_pad, bytes = sizeof(double) * np;
  PhaseClock pc(ctx->stream);
  // one allocation: [x_0 .. x_{dims-1}][y][y/sigma][1/sigma][c]
  double* block = nullptr;
  CU(cudaMalloc(&block, bytes * (s.x_dims + 4)));
  s.x = block; s.y = block + np * s.x_dims; s.yw = s.y + np; s.w = s.yw + np; s.c = s.w + np;
  int rc = upload_rows(ctx, s.x, x, (size_t)n, np, (size_t)s.x_dims);
  if (!rc) rc = upload_rows(ctx, s.y, y, (size_t)n, np, 1);
  if (!rc && sigma) rc = upload_rows(ctx, s.c, sigma, (size_t)n, np, 1);
  pc.mark("upload: alloc + H2D");
  std::vector<MfitPrepBlock> res;
  if (!rc) rc = prepare_dataset(ctx, s, 1, 1, 1, sigma != nullptr, res);
  pc.mark("upload: prepare pass");
  if (rc) { dataset_free_data(s); return rc; }
  CompSum cst;
  double xmin = INFINITY, xmax = -INFINITY, ymax = 0.0;
  bool x_finite = true, pois_ok = true;
  for (const MfitPrepBlock& b : res) {  // block order
    cst.add(b.sum);
    cst.c += b.err;
    if (b.xmin < xmin) xmin = b.xmin;
    if (b.xmax > xmax) xmax = b.xmax;
    if (b.ymax > ymax) ymax = b.ymax;
    if (b.bad_x) x_finite = false;
    if (b.bad_pois) pois_ok = false;
  }
  s.xmin = x_finite && n > 0 && !s.x_is_sequence ? xmin : -INFINITY;
  s.xmax = x_finite && n > 0 && !s.x_is_sequence ? xmax : INFINITY;
  s.cst_normal = cst.value();
  s.cst_poisson = 0.0;
  s.poisson_ok = pois_ok;
  s.poisson_ready = false;   // -sum log y_i! is formed when a log-poisson likelihood first needs it (ensure_poisson_const)
  s.ymax = ymax;
  s.n_err_head = (int)std::min<int64_t>(n, 16);
  for (int k = 0; k < s.n_err_head; k++) s.err_head[k] = sigma ? sigma[k] : 1.0;
  if (ctx->have_chains) chains_free(ctx);  // segment geometry changed: chains must be re-initialised
  return 0;
}

// walker set with per-walker data
int mfit_dataset_upload_per_walker(mfit_ctx* ctx, int d, const double* x, const double* y, const double* sigma,
                                   int64_t n, int64_t W, int x_shared, int sigma_shared) {
  NEED_DS(d);
  if (n < 0 || W < 1 || (n > 0 && (!x || !y))) return fail(ctx, MFIT_ERR_ARG, "dataset %d: bad x/y/n/W", d);
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  Dataset& s = ctx->ds[d];
  dataset_free_data(s);
  s.n = n;
  s.n_pad = round_up(n > 0 ? n : 1, 32);
  s.x_dims = 1;
  s.x_is_sequence = false;
  s.W_pc = W;
  s.x_shared = x_shared != 0;
  s.sigma_shared = sigma_shared != 0 || sigma == nullptr;
  s.sigma_uniform = sigma == nullptr;
  if (sigma && s.sigma_shared) {
    s.sigma_uniform = true;
    for (int64_t i = 1; i < n; i++) if (sigma[i] != sigma[0]) { s.sigma_uniform = false; break; }
  }
  const size_t np = (size_t)s.n_pad, rows = (size_t)W, xrows = x_shared ? 1 : rows, wrows = s.sigma_shared ? 1 : rows;
  // one allocation: [x rows][y rows][y/sigma rows][1/sigma rows][c rows]
  double* block = nullptr;
  CU(cudaMalloc(&block, sizeof(double) * np * (xrows + 2 * rows + 2 * wrows)));
  s.x = block; s.y = s.x + np * xrows; s.yw = s.y + np * rows; s.w = s.yw + np * rows; s.c = s.w + np * wrows;
  int rc = upload_rows(ctx, s.x, x, (size_t)n, np, xrows);
  if (!rc) rc = upload_rows(ctx, s.y, y, (size_t)n, np, rows);
  if (!rc && sigma) rc = upload_rows(ctx, s.c, sigma, (size_t)n, np, wrows);
  std::vector<MfitPrepBlock> res;
  if (!rc) rc = prepare_dataset(ctx, s, rows, xrows, wrows, sigma != nullptr, res);
  if (rc) { dataset_free_data(s); return rc; }
  std::vector<double> cst(rows);
  double xmin = INFINITY, xmax = -INFINITY, ymax = 0.0;
  bool x_finite = true, pois_ok = true;
  for (size_t r = 0; r < rows; r++) {
    const MfitPrepBlock& b = res[r];
    cst[r] = b.sum + b.err;
    if (b.xmin < xmin) xmin = b.xmin;
    if (b.xmax > xmax) xmax = b.xmax;
    if (b.ymax > ymax) ymax = b.ymax;
    if (b.bad_x) x_finite = false;
    if (b.bad_pois) pois_ok = false;
  }
  s.xmin = x_finite && n > 0 ? xmin : -INFINITY;
  s.xmax = x_finite && n > 0 ? xmax : INFINITY;
  s.cst_normal = s.cst_poisson = 0.0;
  s.poisson_ok = pois_ok;
  s.poisson_ready = false;
  s.ymax = ymax;
  const size_t cb = sizeof(double) * rows;
  CU(cudaMalloc(&s.cst_normal_pc, cb)); CU(cudaMalloc(&s.cst_poisson_pc, cb));
  CU(cudaMemcpyAsync(s.cst_normal_pc, cst.data(), cb, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(s.cst_poisson_pc, 0, cb, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->have_chains) chains_free(ctx);
  return 0;
}

int mfit_model_builtin(mfit_ctx* ctx, int d, int family, int degree, const int32_t* param_index, int n_idx) {
  NEED_DS(d);
  const int need = mfit_family_nparams(family, degree);
  if (need < 0 || !mfit_builtin_eval_kernel(family, degree, 0))
    return fail(ctx, MFIT_ERR_ARG, "unknown family %d / degree %d", family, degree);
  if (n_idx != need || need > MFIT_MAXK) return fail(ctx, MFIT_ERR_ARG, "family %d takes %d parameters, got %d", family, need, n_idx);
  Dataset& s = ctx->ds[d];
  for (int k = 0; k < n_idx; k++) {
    if (param_index[k] < 0 || param_index[k] >= ctx->P) return fail(ctx, MFIT_ERR_ARG, "param_index[%d]=%d out of range", k, param_index[k]);
    s.idx[k] = param_index[k];
  }
  dataset_free_model(s);
  s.family = family;
  s.degree = degree;
  s.model_xdims = 0;
  return 0;
}

static int load_cubin(mfit_ctx* ctx, const std::vector<char>& cubin, cudaLibrary_t* lib) {
  CU(cudaLibraryLoadData(lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  return 0;
}

int mfit_model_sexpr(mfit_ctx* ctx, int d, const char* lambda_text, const char* const* key_names, int n_keys) {
  NEED_DS(d);
  if (!lambda_text || n_keys != ctx->P) return fail(ctx, MFIT_ERR_ARG, "need the lambda text and exactly P=%d key names", ctx->P);
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  std::string code, e;
  int model_xdims = 0;
  int rc = mfit_lower_to_string(lambda_text, key_names, n_keys, 0, code, e, &model_xdims);
  if (rc) return fail(ctx, rc, "%s", e.c_str());
  std::vector<char> cubin;
  rc = mfit_nvrtc_compile(mfit_user_model_source(code), cubin, e);
  if (rc) return fail(ctx, rc, "%s", e.c_str());
  Dataset& s = ctx->ds[d];
  dataset_free_model(s);
  rc = load_cubin(ctx, cubin, &s.lib_model);
  if (rc) return rc;
  CU(cudaLibraryGetKernel(&s.k_user[0], s.lib_model, "mfit_user_eval_normal"));
  CU(cudaLibraryGetKernel(&s.k_user[1], s.lib_model, "mfit_user_eval_cutoff"));
  CU(cudaLibraryGetKernel(&s.k_user[2], s.lib_model, "mfit_user_eval_poisson"));
  CU(cudaLibraryGetKernel(&s.k_user_pc[0], s.lib_model, "mfit_user_perchain_normal"));
  CU(cudaLibraryGetKernel(&s.k_user_pc[1], s.lib_model, "mfit_user_perchain_cutoff"));
  CU(cudaLibraryGetKernel(&s.k_user_pc[2], s.lib_model, "mfit_user_perchain_poisson"));
  CU(cudaLibraryGetKernel(&s.k_user_pc[3], s.lib_model, "mfit_user_perchain_staged_normal"));
  CU(cudaLibraryGetKernel(&s.k_user_pc[4], s.lib_model, "mfit_user_perchain_staged_cutoff"));
  CU(cudaLibraryGetKernel(&s.k_user_pc[5], s.lib_model, "mfit_user_perchain_staged_poisson"));
  CU(cudaLibraryGetKernel(&s.k_user_fit, s.lib_model, "mfit_user_fit_envelope"));
  s.family = MFIT_FAM_SEXPR;
  s.degree = 0;
  s.model_xdims = model_xdims;
  s.model_code = code;
  return 0;
}

int mfit_likelihood_set(mfit_ctx* ctx, int d, int kind) {
  NEED_DS(d);
  if (kind < 0 || kind > 2) return fail(ctx, MFIT_ERR_ARG, "unknown likelihood kind %d", kind);
  if (ctx->ds[d].ll_kind != kind && ctx->have_chains) chains_free(ctx);
  if (ctx->ds[d].ll_kind == MFIT_LL_USER) { CU(cudaSetDevice(ctx->device)); dataset_free_term(ctx->ds[d]); }
  ctx->ds[d].ll_kind = kind;
  return 0;
}

int mfit_likelihood_sexpr(mfit_ctx* ctx, int d, const char* lambda_text) {
  NEED_DS(d);
  if (!lambda_text) return fail(ctx, MFIT_ERR_ARG, "need the lambda text");
  const int canon = mfit_ll_canonical_kind(lambda_text);
  if (canon >= 0) return mfit_likelihood_set(ctx, d, canon);
  std::string code, e;
  int rc = mfit_lower_to_string(lambda_text, nullptr, 0, 2, code, e);   // syntax errors now, compilation when the model is known
  if (rc) return fail(ctx, rc, "%s", e.c_str());
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  if (ctx->have_chains) chains_free(ctx);
  Dataset& s = ctx->ds[d];
  dataset_free_term(s);
  s.ll_text = lambda_text;
  s.ll_kind = MFIT_LL_USER;
  return 0;
}

int mfit_prior_flat(mfit_ctx* ctx, int d) {
  NEED_DS(d);
  if (ctx->have_chains) chains_free(ctx);
  dataset_free_prior(ctx->ds[d]);
  ctx->ds[d].prior_kind = 0;
  return 0;
}

int mfit_prior_bounds(mfit_ctx* ctx, int d, int n, const int32_t* idx, const double* lo, const double* hi) {
  NEED_DS(d);
  if (n < 0 || (n > 0 && (!idx || !lo || !hi))) return fail(ctx, MFIT_ERR_ARG, "bad bounds arrays");
  for (int k = 0; k < n; k++)
    if (idx[k] >= ctx->P) return fail(ctx, MFIT_ERR_ARG, "bounds idx[%d]=%d out of range", k, idx[k]);
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  if (ctx->have_chains) chains_free(ctx);
  Dataset& s = ctx->ds[d];
  dataset_free_prior(s);
  s.prior_kind = 1;
  s.nb = n;
  if (n > 0) {
    CU(cudaMalloc(&s.b_idx, sizeof(int) * n));
    CU(cudaMalloc(&s.b_lo, sizeof(double) * n));
    CU(cudaMalloc(&s.b_hi, sizeof(double) * n));
    CU(cudaMemcpyAsync(s.b_idx, idx, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(s.b_lo, lo, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(s.b_hi, hi, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // the caller's arrays are free to go
  }
  return 0;
}

int mfit_prior_sexpr(mfit_ctx* ctx, int d, const char* lambda_text, const char* const* key_names, int n_keys) {
  NEED_DS(d);
  if (!lambda_text || n_keys != ctx->P) return fail(ctx, MFIT_ERR_ARG, "need the lambda text and exactly P=%d key names", ctx->P);
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  std::string code, e;
  int rc = mfit_lower_to_string(lambda_text, key_names, n_keys, 1, code, e);
  if (rc) return fail(ctx, rc, "%s", e.c_str());
  std::vector<char> cubin;
  rc = mfit_nvrtc_compile(mfit_user_prior_source(code), cubin, e);
  if (rc) return fail(ctx, rc, "%s", e.c_str());
  if (ctx->have_chains) chains_free(ctx);
  Dataset& s = ctx->ds[d];
  dataset_free_prior(s);
  rc = load_cubin(ctx, cubin, &s.lib_prior);
  if (rc) return rc;
  CU(cudaLibraryGetKernel(&s.k_prior, s.lib_prior, "mfit_user_prior_kernel"));
  s.prior_kind = 2;
  return 0;
}

int mfit_set_segment_tiles(mfit_ctx* ctx, int tiles_per_segment) {
  if (!ctx || tiles_per_segment < 0) return MFIT_ERR_ARG;
  g_alloc_stream = ctx->stream;
  if (ctx->have_chains) chains_free(ctx);
  ctx->seg_tiles_override = tiles_per_segment;
  return 0;
}

int mfit_segment_geometry(mfit_ctx* ctx, int d, int32_t* tiles_per_segment, int32_t* n_segments) {
  NEED_DS(d);
  Dataset& s = ctx->ds[d];
  if (s.n < 0) return fail(ctx, MFIT_ERR_STATE, "dataset %d has no data", d);
  plan_segments(ctx, s);
  if (tiles_per_segment) *tiles_per_segment = s.tiles_per_seg;
  if (n_segments) *n_segments = s.n_seg;
  return 0;
}

// ------------------------------------------------------------------------------------ walker-make-step on a batch
int mfit_eval_logp(mfit_ctx* ctx, const double* theta, int64_t W, double* prob, double* ll, double* lp) {
  if (!ctx) return MFIT_ERR_ARG;
  if (W < 0 || (W > 0 && !theta)) return fail(ctx, MFIT_ERR_ARG, "bad theta/W");
  int rc = check_problem_ready(ctx);
  if (rc) return rc;
  if (W == 0) return 0;
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  Workspace w;
  rc = workspace_make(ctx, w, W);
  if (rc) { workspace_free(w); return rc; }
  const int P = ctx->P, D = ctx->D;
  double *th_rm = nullptr, *th = nullptr, *o_prob = nullptr, *o_ll = nullptr, *o_lp = nullptr;
  auto cleanup = [&]() { free_dev(th_rm); free_dev(th); free_dev(o_prob); free_dev(o_ll); free_dev(o_lp); workspace_free(w); };
#define CUX(call)                                                                          \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) { cleanup(); return fail(ctx, MFIT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } \
  } while (0)
  CUX(cudaMalloc(&th_rm, sizeof(double) * W * P));
  CUX(cudaMalloc(&th, sizeof(double) * P * w.stride));
  CUX(cudaMemsetAsync(th, 0, sizeof(double) * P * w.stride, ctx->stream));
  CUX(cudaMalloc(&o_prob, sizeof(double) * W));
  CUX(cudaMalloc(&o_ll, sizeof(double) * W * D));
  CUX(cudaMalloc(&o_lp, sizeof(double) * W * D));
  CUX(cudaMemcpyAsync(th_rm, theta, sizeof(double) * W * P, cudaMemcpyHostToDevice, ctx->stream));
  mfit_launch_transpose_in(th_rm, th, W, P, w.stride, ctx->stream);
  if (w.peer && ctx->run_open) { cleanup(); return fail(ctx, MFIT_ERR_STATE, "a run is open: its pending terms live in the peer exchange buffer"); }
  MfitChains v = view_of(ctx, w);
  mfit_launch_priors(v, th, ctx->stream);
  ctx->launches += 2;
  ctx->eval_seconds = 0.0;
  ctx->eval_launches = 0;
  ctx->ev_used = 0;
  rc = eval_all(ctx, w, v, th, nullptr);
  if (rc) { cleanup(); return rc; }
  mfit_launch_finalize(v, ctx->stream);
  ctx->launches++;
  rc = exchange(ctx, w, &v);
  if (rc) { cleanup(); return rc; }
  mfit_launch_eval_finalize(v, o_prob, o_ll, o_lp, ctx->stream);
  ctx->launches++;
  if (w.peer) ctx->peers.epoch++;
  if (prob) CUX(cudaMemcpyAsync(prob, o_prob, sizeof(double) * W, cudaMemcpyDeviceToHost, ctx->stream));
  if (ll) CUX(cudaMemcpyAsync(ll, o_ll, sizeof(double) * W * D, cudaMemcpyDeviceToHost, ctx->stream));
  if (lp) CUX(cudaMemcpyAsync(lp, o_lp, sizeof(double) * W * D, cudaMemcpyDeviceToHost, ctx->stream));
  CUX(cudaStreamSynchronize(ctx->stream));
  CUX(cudaGetLastError());
  collect_timing(ctx);
  cleanup();
#undef CUX
  return 0;
}

// ------------------------------------------------------------------------------------ chains
int mfit_chains_init(mfit_ctx* ctx, const double* theta0, int64_t W, uint64_t seed, int64_t chain_id_base,
                     int64_t trace_capacity) {
  if (!ctx) return MFIT_ERR_ARG;
  if (W < 1 || !theta0) return fail(ctx, MFIT_ERR_ARG, "need W >= 1 and theta0");
  if (W > 2000000000LL) return fail(ctx, MFIT_ERR_ARG, "W too large");
  int rc = check_problem_ready(ctx);
  if (rc) return rc;
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  PhaseClock pc(ctx->stream);
  chains_free(ctx);
  rc = workspace_make(ctx, ctx->ws, W);
  if (rc) return rc;
  pc.mark("chains_init: workspace");
  const int P = ctx->P, D = ctx->D;
  MfitChains& c = ctx->ch;
  c = view_of(ctx, ctx->ws);
  const size_t st = (size_t)c.stride;
  // every per-chain array out of ONE zeroed allocation (seventeen allocations and memsets cost more than the first step)
  struct Slot { void** p; size_t bytes; };
  std::vector<Slot> slots;
  size_t total = 0;
#define CA(field, count)                                                            \
  slots.push_back({(void**)&c.field, (sizeof(*c.field) * (size_t)(count) + 255) & ~(size_t)255}); \
  total += slots.back().bytes;
  CA(cur_theta, P * st) CA(cur_prob, st) CA(prop_theta, P * st) CA(best_theta, P * st) CA(best_prob, st)
  CA(L, (size_t)P * P * st) CA(loop_i, st) CA(shutting, st) CA(done, st) CA(need_reest, st) CA(temperature, st)
  CA(length, st) CA(age, st) CA(rows, st) CA(reset_index, st) CA(steps_run, st) CA(l_updates, st)
  CA(reest_list, st) CA(reest_count, 8)
#undef CA
  (void)D;
  {
    char* base = nullptr;
    if ((rc = chain_alloc(ctx, &base, total)) != 0) { chains_free(ctx); return rc; }
    for (const Slot& sl : slots) { *sl.p = base; base += sl.bytes; }
  }
  long long cap = trace_capacity;
  ctx->cap_fixed = trace_capacity > 0;
  const long long mincap = min_trace_cap(P);
  if (cap < mincap) cap = mincap;
  c.cap = cap;
  if ((rc = chain_alloc(ctx, &c.trace_theta, (size_t)cap * P * st, false)) != 0) { chains_free(ctx); return rc; }
  if ((rc = chain_alloc(ctx, &c.trace_prob, (size_t)cap * st, false)) != 0) { chains_free(ctx); return rc; }
  pc.mark("chains_init: chain arrays");
  ctx->seed = seed;
  ctx->chain_base = chain_id_base;
  ctx->have_chains = true;
  // first step, walker-create :1148-1162
  double* th_rm = nullptr;
  CU(cudaMalloc(&th_rm, sizeof(double) * W * P));
  CU(cudaMemcpyAsync(th_rm, theta0, sizeof(double) * W * P, cudaMemcpyHostToDevice, ctx->stream));
  mfit_launch_transpose_in(th_rm, c.prop_theta, W, P, c.stride, ctx->stream);
  sync_epoch(ctx, c);
  mfit_launch_priors(c, c.prop_theta, ctx->stream);
  ctx->launches += 2;
  rc = eval_all(ctx, ctx->ws, c, c.prop_theta, nullptr);
  if (rc) { cudaFree(th_rm); return rc; }
  mfit_launch_finalize(c, ctx->stream);
  ctx->launches++;
  rc = exchange(ctx, ctx->ws, &c);
  if (rc) { cudaFree(th_rm); return rc; }
  pc.mark("chains_init: first eval");
  mfit_launch_first_step(c, ctx->stream);
  if (ctx->ws.peer) ctx->peers.epoch++;
  ctx->rows_upper = 1;
  ctx->launches++;
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(th_rm);
  if (e != cudaSuccess) return fail(ctx, MFIT_ERR_CUDA, "chains_init: %s", cudaGetErrorString(e));
  CU(cudaGetLastError());
  collect_timing(ctx);
  return 0;
}

static int make_cfg(mfit_ctx* ctx, const mfit_adapt_cfg* cfg, MfitRunCfg& r) {
  memset(&r, 0, sizeof r);
  const int P = ctx->P;
  r.mode = MFIT_MODE_ADAPTIVE;
  r.n = cfg->n;                                   // (floor n) :866 -- the ABI takes an integer
  r.S = 10LL * (P > 50 ? P : 50);                 // steps-to-settle :873
  r.Tn = r.n > 10 * r.S ? r.n : 10 * r.S;         // temp-steps :875
  r.tail = r.S > 2000 ? r.S : 2000;               // (max 2000 steps-to-settle) :906
  r.mwl = cfg->max_walker_length > 0 ? cfg->max_walker_length / 2 : 0;  // :868
  r.auto_mode = cfg->auto_mode;
  r.sampling_opt = cfg->sampling_optimization;
  r.greedy = cfg->greedy;
  r.temperature0 = cfg->temperature;
  // temps :878  (max 1 (* (cos (* x pi (+ 1 (* 2 (floor temp-steps 5000))) (/ (* 2 temp-steps)))) temperature))
  // 30000 cosines are half a millisecond of host time per run: the last table is kept per process
  static std::mutex temps_mu;
  static std::vector<double> temps_cache;
  static long long temps_cache_Tn = -1;
  static double temps_cache_T = 0.0;
  std::lock_guard<std::mutex> lock(temps_mu);
  if (temps_cache_Tn != r.Tn || temps_cache_T != cfg->temperature) {
    temps_cache.resize((size_t)r.Tn);
    const double pi = 3.141592653589793;
    const double k = (double)(1 + 2 * (r.Tn / 5000));
    const double inv = 1.0 / (double)(2 * r.Tn);
    for (long long x = 0; x < r.Tn; x++) {
      const double a = (((double)x * pi) * k) * inv;
      const double t = std::cos(a) * cfg->temperature;
      temps_cache[(size_t)x] = t > 1.0 ? t : 1.0;
    }
    temps_cache_Tn = r.Tn;
    temps_cache_T = cfg->temperature;
  }
  const std::vector<double>& temps = temps_cache;
  if (ctx->temps_len < r.Tn) {
    free_dev(ctx->temps_dev);
    ctx->temps_dev = nullptr;
    ctx->temps_len = 0;
    CU(cudaMalloc(&ctx->temps_dev, sizeof(double) * (size_t)r.Tn));
    ctx->temps_len = r.Tn;
  }
  CU(cudaMemcpyAsync(ctx->temps_dev, temps.data(), sizeof(double) * (size_t)r.Tn, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  r.temps = ctx->temps_dev;
  return 0;
}

static int attach_streams(mfit_ctx* ctx, MfitRunCfg& r, const double* z, const double* u, int64_t nstream,
                          long long need, bool need_u) {
  r.z = nullptr;
  r.u = nullptr;
  r.nstream = nstream;
  if (!z && !u) return 0;
  if (!z) return fail(ctx, MFIT_ERR_ARG, "u without z");
  if (nstream < need) return fail(ctx, MFIT_ERR_STREAM, "injected stream holds %lld steps, the run may take %lld", (long long)nstream, need);
  const size_t W = (size_t)ctx->ch.W;
  int rc = upload_opt(ctx, z, W * nstream * ctx->P, &ctx->z_dev, &ctx->z_bytes);
  if (rc) return rc;
  r.z = ctx->z_dev;
  if (u) {
    rc = upload_opt(ctx, u, W * nstream, &ctx->u_dev, &ctx->u_bytes);
    if (rc) return rc;
    r.u = ctx->u_dev;
  } else if (need_u)
    return fail(ctx, MFIT_ERR_ARG, "z injected without u");
  return 0;
}

int mfit_run_adaptive(mfit_ctx* ctx, const mfit_adapt_cfg* cfg, const double* l_matrix, const double* z,
                      const double* u, int64_t nstream, int64_t* steps_launched) {
  if (!ctx || !cfg) return MFIT_ERR_ARG;
  if (!ctx->have_chains) return fail(ctx, MFIT_ERR_STATE, "mfit_chains_init has not been called");
  if (cfg->n < 0) return fail(ctx, MFIT_ERR_ARG, "n < 0");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  MfitRunCfg r;
  int rc = make_cfg(ctx, cfg, r);
  if (rc) return rc;
  // i runs 1..n-1, but the shutdown test (:905-917) jumps back to i = n - tail once: n steps when n > tail,
  // tail steps for 2 <= n <= tail (the jump lands on a non-positive i), none for n < 2
  const long long max_steps = r.n > 1 ? (r.n > r.tail ? r.n : r.tail) : 0;
  rc = attach_streams(ctx, r, z, u, nstream, max_steps, !cfg->greedy);
  if (rc) return rc;
  long long launched = 0;
  rc = run_loop(ctx, r, l_matrix, max_steps, &launched);
  if (steps_launched) *steps_launched = launched;
  return rc;
}

int mfit_run_begin(mfit_ctx* ctx, const mfit_adapt_cfg* cfg, const double* l_matrix, const double* z, const double* u,
                   int64_t nstream) {
  if (!ctx || !cfg) return MFIT_ERR_ARG;
  if (!ctx->have_chains) return fail(ctx, MFIT_ERR_STATE, "mfit_chains_init has not been called");
  if (cfg->n < 0) return fail(ctx, MFIT_ERR_ARG, "n < 0");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  MfitRunCfg r;
  int rc = make_cfg(ctx, cfg, r);
  if (rc) return rc;
  const long long max_steps = r.n > 1 ? (r.n > r.tail ? r.n : r.tail) : 0;  // see mfit_run_adaptive
  rc = attach_streams(ctx, r, z, u, nstream, z ? (nstream < max_steps ? nstream : max_steps) : 0, !cfg->greedy);
  if (rc) return rc;
  return run_begin(ctx, r, l_matrix, max_steps);
}

int mfit_run_advance(mfit_ctx* ctx, int64_t max_steps, int64_t* steps_done, int* finished) {
  if (!ctx) return MFIT_ERR_ARG;
  if (!ctx->have_chains) return fail(ctx, MFIT_ERR_STATE, "mfit_chains_init has not been called");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  if (ctx->run_cfg.z && ctx->run_t + (max_steps < 0 ? ctx->run_max - ctx->run_t : max_steps) > ctx->run_cfg.nstream &&
      ctx->run_t + 1 > ctx->run_cfg.nstream)
    return fail(ctx, MFIT_ERR_STREAM, "injected stream exhausted");
  long long budget = max_steps;
  if (ctx->run_cfg.z) {  // never step past the injected stream
    const long long left = ctx->run_cfg.nstream - ctx->run_t;
    if (budget < 0 || budget > left) budget = left;
  }
  long long done = 0;
  int rc = run_advance(ctx, budget, &done, finished);
  if (steps_done) *steps_done = done;
  return rc;
}

int mfit_run_many_steps(mfit_ctx* ctx, int64_t n, const double* l_matrix, const double* z, const double* u,
                        int64_t nstream) {
  if (!ctx) return MFIT_ERR_ARG;
  if (!ctx->have_chains) return fail(ctx, MFIT_ERR_STATE, "mfit_chains_init has not been called");
  if (!l_matrix) return fail(ctx, MFIT_ERR_ARG, "walker-many-steps needs an l-matrix here (the :median-params default is post-processing)");
  if (n < 0) return fail(ctx, MFIT_ERR_ARG, "n < 0");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  MfitRunCfg r;
  memset(&r, 0, sizeof r);
  r.mode = MFIT_MODE_MANY;
  r.n = n;
  r.S = 10LL * (ctx->P > 50 ? ctx->P : 50);
  r.temperature0 = 1.0;
  int rc = attach_streams(ctx, r, z, u, nstream, n, true);
  if (rc) return rc;
  return run_loop(ctx, r, l_matrix, n, nullptr);
}

int mfit_run_many_steps_per_chain(mfit_ctx* ctx, int64_t n, const double* l_matrices, const double* z, const double* u,
                                  int64_t nstream) {
  if (!ctx) return MFIT_ERR_ARG;
  if (!ctx->have_chains) return fail(ctx, MFIT_ERR_STATE, "mfit_chains_init has not been called");
  if (!l_matrices) return fail(ctx, MFIT_ERR_ARG, "l_matrices is NULL");
  if (n < 0) return fail(ctx, MFIT_ERR_ARG, "n < 0");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  MfitRunCfg r;
  memset(&r, 0, sizeof r);
  r.mode = MFIT_MODE_MANY;
  r.n = n;
  r.S = 10LL * (ctx->P > 50 ? ctx->P : 50);
  r.temperature0 = 1.0;
  r.l_per_chain = 1;
  int rc = attach_streams(ctx, r, z, u, nstream, n, true);
  if (rc) return rc;
  return run_loop(ctx, r, l_matrices, n, nullptr);
}

int mfit_run_sample_region(mfit_ctx* ctx, double initial_scale, const double* z, int64_t nstream) {
  if (!ctx) return MFIT_ERR_ARG;
  if (!ctx->have_chains) return fail(ctx, MFIT_ERR_STATE, "mfit_chains_init has not been called");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  MfitRunCfg r;
  memset(&r, 0, sizeof r);
  r.mode = MFIT_MODE_SAMPLE_REGION;
  r.n = 3000;
  r.S = 10LL * (ctx->P > 50 ? ctx->P : 50);
  r.greedy = 1;
  r.temperature0 = 1.0;
  r.sr_scale = initial_scale;
  int rc = attach_streams(ctx, r, z, nullptr, nstream, 2999, false);
  if (rc) return rc;
  return run_loop(ctx, r, nullptr, 2999, nullptr);
}

int mfit_force_take_step(mfit_ctx* ctx) {
  if (!ctx) return MFIT_ERR_ARG;
  if (!ctx->have_chains) return fail(ctx, MFIT_ERR_STATE, "mfit_chains_init has not been called");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  MfitChains& c = ctx->ch;
  int rc = grow_trace(ctx, ctx->rows_upper + 2);
  if (rc) return rc;
  ctx->rows_upper += 1;
  sync_epoch(ctx, c);
  mfit_launch_priors(c, c.cur_theta, ctx->stream);
  ctx->launches++;
  rc = eval_all(ctx, ctx->ws, c, c.cur_theta, nullptr);
  if (rc) return rc;
  mfit_launch_finalize(c, ctx->stream);
  ctx->launches++;
  rc = exchange(ctx, ctx->ws, &c);
  if (rc) return rc;
  mfit_launch_force_step(c, ctx->stream);
  if (ctx->ws.peer) ctx->peers.epoch++;
  ctx->launches++;
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaGetLastError());
  collect_timing(ctx);
  return 0;
}

int mfit_request_stop(mfit_ctx* ctx) {
  if (!ctx) return MFIT_ERR_ARG;
  ctx->stop.store(1);
  return 0;
}

int mfit_trace_read(mfit_ctx* ctx, int64_t take, double* theta, double* prob, int64_t* n_valid) {
  if (!ctx) return MFIT_ERR_ARG;
  if (!ctx->have_chains) return fail(ctx, MFIT_ERR_STATE, "mfit_chains_init has not been called");
  if (take < 1) return fail(ctx, MFIT_ERR_ARG, "take < 1");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  MfitChains& c = ctx->ch;
  const size_t W = (size_t)c.W;
  double *dt = nullptr, *dp = nullptr;
  long long* dn = nullptr;
  int rc = 0;
  cudaError_t e = cudaSuccess;
  if (theta) e = cudaMalloc(&dt, sizeof(double) * W * take * c.P);
  if (e == cudaSuccess && prob) e = cudaMalloc(&dp, sizeof(double) * W * take);
  if (e == cudaSuccess && n_valid) e = cudaMalloc(&dn, sizeof(long long) * W);
  const bool dbg = getenv("MFIT_DEBUG_TIMING") != nullptr;
  const double t_a = now_ms();
  if (e == cudaSuccess) {
    mfit_launch_trace_gather(c, take, dt, dp, dn, ctx->stream);
    ctx->launches++;
    if (dbg) { cudaStreamSynchronize(ctx->stream); fprintf(stderr, "trace_read: malloc+gather %.1f ms\n", now_ms() - t_a); }
    if (n_valid) e = cudaMemcpyAsync(n_valid, dn, sizeof(long long) * W, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && theta) rc = d2h_staged(ctx, theta, dt, sizeof(double) * W * take * c.P);
    if (e == cudaSuccess && rc == 0 && prob) rc = d2h_staged(ctx, prob, dp, sizeof(double) * W * take);
    if (e == cudaSuccess && rc == 0) e = cudaStreamSynchronize(ctx->stream);
  }
  if (e != cudaSuccess) rc = fail(ctx, e == cudaErrorMemoryAllocation ? MFIT_ERR_NOMEM : MFIT_ERR_CUDA, "trace_read: %s", cudaGetErrorString(e));
  if (dbg) fprintf(stderr, "trace_read: +copies %.1f ms\n", now_ms() - t_a);
  free_dev(dt); free_dev(dp); free_dev(dn);
  if (dbg) fprintf(stderr, "trace_read: +free %.1f ms\n", now_ms() - t_a);
  return rc;
}

int mfit_chains_restore(mfit_ctx* ctx, int64_t rows, const double* theta, const double* prob, const int64_t* n_valid,
                        const int64_t* age, const double* best_theta, const double* best_prob) {
  if (!ctx) return MFIT_ERR_ARG;
  if (!ctx->have_chains) return fail(ctx, MFIT_ERR_STATE, "mfit_chains_init has not been called");
  if (ctx->run_open) return fail(ctx, MFIT_ERR_STATE, "a run is open");
  if (rows < 1 || !theta || !prob || !n_valid) return fail(ctx, MFIT_ERR_ARG, "need rows >= 1, theta, prob and n_valid");
  if ((best_theta == nullptr) != (best_prob == nullptr)) return fail(ctx, MFIT_ERR_ARG, "best_theta and best_prob go together");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  MfitChains& c = ctx->ch;
  const size_t W = (size_t)c.W, P = (size_t)c.P;
  long long most = 0;
  for (size_t w = 0; w < W; w++) {
    if (n_valid[w] < 1 || n_valid[w] > rows) return fail(ctx, MFIT_ERR_ARG, "n_valid[%zu] = %lld outside 1 .. rows", w, (long long)n_valid[w]);
    if (age && age[w] < n_valid[w]) return fail(ctx, MFIT_ERR_ARG, "age[%zu] below the number of steps held", w);
    if (n_valid[w] > most) most = n_valid[w];
  }
  int rc = grow_trace(ctx, most + 1);
  if (rc) return rc;
  double *dt = nullptr, *dp = nullptr, *dbt = nullptr, *dbp = nullptr;
  long long *dn = nullptr, *da = nullptr;
  cudaError_t e = cudaMalloc(&dt, sizeof(double) * W * (size_t)rows * P);
  if (e == cudaSuccess) e = cudaMalloc(&dp, sizeof(double) * W * (size_t)rows);
  if (e == cudaSuccess) e = cudaMalloc(&dn, sizeof(long long) * W);
  if (e == cudaSuccess && age) e = cudaMalloc(&da, sizeof(long long) * W);
  if (e == cudaSuccess && best_theta) e = cudaMalloc(&dbt, sizeof(double) * W * P);
  if (e == cudaSuccess && best_prob) e = cudaMalloc(&dbp, sizeof(double) * W);
  auto up = [&](void* dst, const void* src, size_t bytes) {
    if (e == cudaSuccess && dst) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
  };
  up(dt, theta, sizeof(double) * W * (size_t)rows * P);
  up(dp, prob, sizeof(double) * W * (size_t)rows);
  up(dn, n_valid, sizeof(long long) * W);
  up(da, age, sizeof(long long) * W);
  up(dbt, best_theta, sizeof(double) * W * P);
  up(dbp, best_prob, sizeof(double) * W);
  if (e == cudaSuccess) {
    mfit_launch_trace_restore(c, rows, dt, dp, dn, da, dbt, dbp, ctx->stream);
    ctx->launches += 2;
    e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  free_dev(dt); free_dev(dp); free_dev(dn); free_dev(da); free_dev(dbt); free_dev(dbp);
  if (e != cudaSuccess)
    return fail(ctx, e == cudaErrorMemoryAllocation ? MFIT_ERR_NOMEM : MFIT_ERR_CUDA, "chains_restore: %s", cudaGetErrorString(e));
  ctx->rows_upper = most;
  return 0;
}

int mfit_stats_read(mfit_ctx* ctx, mfit_chain_stats* stats) {
  if (!ctx || !stats) return MFIT_ERR_ARG;
  if (!ctx->have_chains) return fail(ctx, MFIT_ERR_STATE, "mfit_chains_init has not been called");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  MfitChains& c = ctx->ch;
  mfit_chain_stats* d = nullptr;
  CU(cudaMalloc(&d, sizeof(mfit_chain_stats) * c.W));
  mfit_launch_stats(c, d, ctx->stream);
  ctx->launches++;
  cudaError_t e = cudaMemcpyAsync(stats, d, sizeof(mfit_chain_stats) * c.W, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(ctx, MFIT_ERR_CUDA, "stats_read: %s", cudaGetErrorString(e));
  return 0;
}

static int read_soa(mfit_ctx* ctx, const double* src_theta, const double* src_prob, double* theta, double* prob) {
  MfitChains& c = ctx->ch;
  if (theta) {
    double* tmp = nullptr;
    CU(cudaMalloc(&tmp, sizeof(double) * (size_t)c.W * c.P));
    mfit_launch_transpose_out(src_theta, tmp, c.W, c.P, c.stride, ctx->stream);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(theta, tmp, sizeof(double) * (size_t)c.W * c.P, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(ctx, MFIT_ERR_CUDA, "read: %s", cudaGetErrorString(e));
  }
  if (prob) {
    CU(cudaMemcpyAsync(prob, src_prob, sizeof(double) * (size_t)c.W, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

int mfit_best_read(mfit_ctx* ctx, double* theta, double* prob) {
  if (!ctx) return MFIT_ERR_ARG;
  if (!ctx->have_chains) return fail(ctx, MFIT_ERR_STATE, "mfit_chains_init has not been called");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  return read_soa(ctx, ctx->ch.best_theta, ctx->ch.best_prob, theta, prob);
}
int mfit_last_read(mfit_ctx* ctx, double* theta, double* prob) {
  if (!ctx) return MFIT_ERR_ARG;
  if (!ctx->have_chains) return fail(ctx, MFIT_ERR_STATE, "mfit_chains_init has not been called");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  return read_soa(ctx, ctx->ch.cur_theta, ctx->ch.cur_prob, theta, prob);
}

int mfit_l_matrix_read(mfit_ctx* ctx, double* l) {
  if (!ctx || !l) return MFIT_ERR_ARG;
  if (!ctx->have_chains) return fail(ctx, MFIT_ERR_STATE, "mfit_chains_init has not been called");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  MfitChains& c = ctx->ch;
  const size_t cnt = (size_t)c.W * c.P * c.P;
  double* tmp = nullptr;
  CU(cudaMalloc(&tmp, sizeof(double) * cnt));
  mfit_launch_l_read(c, tmp, ctx->stream);
  ctx->launches++;
  cudaError_t e = cudaMemcpyAsync(l, tmp, sizeof(double) * cnt, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(tmp);
  if (e != cudaSuccess) return fail(ctx, MFIT_ERR_CUDA, "l_matrix_read: %s", cudaGetErrorString(e));
  return 0;
}

int mfit_trace_analytics(mfit_ctx* ctx, int64_t take, double* most_likely_theta, double* most_likely_prob,
                         double* median, double* acceptance, double* covariance, double* l_matrix) {
  if (!ctx) return MFIT_ERR_ARG;
  if (!ctx->have_chains) return fail(ctx, MFIT_ERR_STATE, "mfit_chains_init has not been called");
  if (take < 1) return fail(ctx, MFIT_ERR_ARG, "take < 1");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  MfitChains& c = ctx->ch;
  if (take > c.cap) take = c.cap;  // no chain holds more rows than the ring
  const size_t W = (size_t)c.W, P = (size_t)c.P;
  struct Out { double* host; size_t count; double* dev; };
  Out outs[6] = {{most_likely_theta, W * P, nullptr}, {most_likely_prob, W, nullptr}, {median, W * P, nullptr},
                 {acceptance, W, nullptr}, {covariance, W * P * P, nullptr}, {l_matrix, W * P * P, nullptr}};
  int* sel = nullptr;
  cudaError_t e = cudaSuccess;
  for (auto& o : outs)
    if (o.host && e == cudaSuccess) e = cudaMalloc(&o.dev, sizeof(double) * o.count);
  const bool window = most_likely_theta || most_likely_prob || acceptance || covariance || l_matrix;
  if (e == cudaSuccess && window) e = cudaMalloc(&sel, sizeof(int) * W * (size_t)(take + 1));
  if (e == cudaSuccess) {
    if (window) {
      mfit_launch_analytics(c, take, sel, outs[0].dev, outs[1].dev, outs[3].dev, outs[4].dev, outs[5].dev, ctx->stream);
      ctx->launches++;
    }
    if (median) {
      mfit_launch_median(c, take, outs[2].dev, ctx->stream);
      ctx->launches++;
    }
    for (auto& o : outs)
      if (o.host && e == cudaSuccess) e = cudaMemcpyAsync(o.host, o.dev, sizeof(double) * o.count, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  for (auto& o : outs) free_dev(o.dev);
  free_dev(sel);
  if (e != cudaSuccess)
    return fail(ctx, e == cudaErrorMemoryAllocation ? MFIT_ERR_NOMEM : MFIT_ERR_CUDA, "trace_analytics: %s", cudaGetErrorString(e));
  return 0;
}

int mfit_fit_envelope(mfit_ctx* ctx, int d, int64_t take, int which_solution, const double* x_fit, int n_fit,
                      double* y_fit, double* y_max, double* y_min, double* params, int64_t* n_distinct) {
  if (!ctx) return MFIT_ERR_ARG;
  if (!ctx->have_chains) return fail(ctx, MFIT_ERR_STATE, "mfit_chains_init has not been called");
  if (d < 0 || d >= ctx->D) return fail(ctx, MFIT_ERR_ARG, "dataset index %d out of range", d);
  if (take < 1) return fail(ctx, MFIT_ERR_ARG, "take < 1");
  if (!x_fit || n_fit < 1) return fail(ctx, MFIT_ERR_ARG, "x_fit: need at least one point");
  if (which_solution != MFIT_FIT_MOST_LIKELY && which_solution != MFIT_FIT_MEDIAN)
    return fail(ctx, MFIT_ERR_ARG, "which_solution must be MFIT_FIT_MOST_LIKELY or MFIT_FIT_MEDIAN");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  Dataset& s = ctx->ds[d];
  const void* fn = s.family == MFIT_FAM_SEXPR ? (const void*)s.k_user_fit : mfit_builtin_fit_kernel(s.family, s.degree);
  if (!fn) return fail(ctx, MFIT_ERR_STATE, "dataset %d has no model", d);
  if (s.model_xdims > 0) return fail(ctx, MFIT_ERR_ARG, "dataset %d: the fit curve needs a model of a one-dimensional x", d);
  MfitChains& c = ctx->ch;
  if (take > c.cap) take = c.cap;  // no chain holds more rows than the ring
  const size_t W = (size_t)c.W, P = (size_t)c.P;
  // columns per chain: the distinct steps among the ceil(0.66 take) best entries, plus the fit's own parameters
  const long long cols = (long long)std::ceil(0.66f * (float)take) + 1;
  // chains per pass: gathered parameter columns bounded by 512 MB
  long long nb = (long long)((512ull << 20) / (sizeof(double) * P * (size_t)cols));
  if (nb < 1) nb = 1;
  if (nb > (long long)W) nb = (long long)W;
  double *fit = nullptr, *cols_dev = nullptr, *xf = nullptr, *out = nullptr;
  int *sel = nullptr, *nsel = nullptr;
  std::vector<int> nsel_host((size_t)nb);
  cudaError_t e = cudaMalloc(&fit, sizeof(double) * W * P);
  if (e == cudaSuccess) e = cudaMalloc(&cols_dev, sizeof(double) * (size_t)nb * P * (size_t)cols);
  if (e == cudaSuccess) e = cudaMalloc(&xf, sizeof(double) * (size_t)n_fit);
  if (e == cudaSuccess) e = cudaMalloc(&out, sizeof(double) * 3 * (size_t)nb * (size_t)n_fit);
  if (e == cudaSuccess) e = cudaMalloc(&nsel, sizeof(int) * (size_t)nb);
  // scratch of the selection: the window pass of the most likely step wants take + 1 offsets per chain
  const size_t sel_window = which_solution == MFIT_FIT_MOST_LIKELY ? W * (size_t)(take + 1) : 0;
  const size_t sel_cols = (size_t)nb * (size_t)cols;
  if (e == cudaSuccess) e = cudaMalloc(&sel, sizeof(int) * (sel_window > sel_cols ? sel_window : sel_cols));
  if (e == cudaSuccess) e = cudaMemcpyAsync(xf, x_fit, sizeof(double) * (size_t)n_fit, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    // params :1240-1243
    if (which_solution == MFIT_FIT_MOST_LIKELY)
      mfit_launch_analytics(c, take, sel, fit, nullptr, nullptr, nullptr, nullptr, ctx->stream);
    else
      mfit_launch_median(c, take, fit, ctx->stream);
    ctx->launches++;
    if (params) e = cudaMemcpyAsync(params, fit, sizeof(double) * W * P, cudaMemcpyDeviceToHost, ctx->stream);
  }
  for (long long w0 = 0; w0 < (long long)W && e == cudaSuccess; w0 += nb) {
    const long long nw = w0 + nb <= (long long)W ? nb : (long long)W - w0;
    mfit_launch_envelope_select(c, take, w0, nw, cols, fit, cols_dev, nsel, sel, ctx->stream);
    MfitFitArgs f;
    memset(&f, 0, sizeof f);
    f.a.xmin = s.xmin; f.a.xmax = s.xmax; f.a.n = s.n;
    for (int k = 0; k < MFIT_MAXK; k++) f.a.idx[k] = s.idx[k];
    f.a.degree = s.degree;
    f.a.theta = cols_dev;
    f.chain_pitch = (long long)P * cols;
    f.cols = cols;
    f.n_sel = nsel;
    f.xfit = xf;
    f.n_fit = n_fit;
    f.ymax = out;
    f.ymin = out + (size_t)nb * (size_t)n_fit;
    f.yfit = out + 2 * (size_t)nb * (size_t)n_fit;
    void* args[] = {&f};
    const unsigned gy = (unsigned)((n_fit + MFIT_FIT_THREADS * MFIT_FIT_PTS - 1) / (MFIT_FIT_THREADS * MFIT_FIT_PTS));
    e = cudaLaunchKernel(fn, dim3((unsigned)nw, gy), dim3(MFIT_FIT_THREADS), args, 0, ctx->stream);
    ctx->launches += 2;
    const size_t cnt = sizeof(double) * (size_t)nw * (size_t)n_fit, off = (size_t)w0 * (size_t)n_fit;
    if (e == cudaSuccess && y_max) e = cudaMemcpyAsync(y_max + off, f.ymax, cnt, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && y_min) e = cudaMemcpyAsync(y_min + off, f.ymin, cnt, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && y_fit) e = cudaMemcpyAsync(y_fit + off, f.yfit, cnt, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && n_distinct)
      e = cudaMemcpyAsync(nsel_host.data(), nsel, sizeof(int) * (size_t)nw, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // the pass buffers are reused
    if (e == cudaSuccess && n_distinct)
      for (long long k = 0; k < nw; k++) n_distinct[w0 + k] = nsel_host[(size_t)k];
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  free_dev(fit); free_dev(cols_dev); free_dev(xf); free_dev(out); free_dev(nsel); free_dev(sel);
  if (e != cudaSuccess)
    return fail(ctx, e == cudaErrorMemoryAllocation ? MFIT_ERR_NOMEM : MFIT_ERR_CUDA, "fit_envelope: %s", cudaGetErrorString(e));
  return 0;
}

int mfit_keep_walks(mfit_ctx* ctx, int64_t keep) {
  if (!ctx) return MFIT_ERR_ARG;
  if (!ctx->have_chains) return fail(ctx, MFIT_ERR_STATE, "mfit_chains_init has not been called");
  if (keep < 1) return fail(ctx, MFIT_ERR_ARG, "keep < 1");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  mfit_launch_keep_walks(ctx->ch, keep, ctx->stream);
  ctx->launches++;
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int mfit_walker_modify(mfit_ctx* ctx, int mode, int64_t number) {
  if (!ctx) return MFIT_ERR_ARG;
  if (!ctx->have_chains) return fail(ctx, MFIT_ERR_STATE, "mfit_chains_init has not been called");
  if (ctx->run_open) return fail(ctx, MFIT_ERR_STATE, "a run is open");
  if (mode == MFIT_MODIFY_KEEP_WALKS) return mfit_keep_walks(ctx, number);
  if (mode != MFIT_MODIFY_BURN_WALKS && mode != MFIT_MODIFY_RESET && mode != MFIT_MODIFY_RESET_TO_MOST_LIKELY)
    return fail(ctx, MFIT_ERR_ARG, "unknown walker-modify mode %d", mode);
  if (mode == MFIT_MODIFY_BURN_WALKS && number < 0) return fail(ctx, MFIT_ERR_ARG, "burn-number < 0");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  mfit_launch_modify(ctx->ch, mode, number, ctx->stream);
  ctx->launches++;
  if (mode != MFIT_MODIFY_BURN_WALKS) ctx->rows_upper += 1;
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int mfit_philox_stream(mfit_ctx* ctx, uint64_t seed, int64_t chain_id_base, int64_t W, int64_t first_step,
                       int64_t nsteps, int P, double* z, double* u) {
  if (!ctx || !z || !u || W < 1 || nsteps < 1 || P < 1 || P > MFIT_MAX_P) return MFIT_ERR_ARG;
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  double *dz = nullptr, *du = nullptr;
  CU(cudaMalloc(&dz, sizeof(double) * W * nsteps * P));
  cudaError_t e = cudaMalloc(&du, sizeof(double) * W * nsteps);
  if (e == cudaSuccess) {
    mfit_launch_philox_stream(seed, chain_id_base, W, first_step, nsteps, P, dz, du, ctx->stream);
    ctx->launches++;
    e = cudaMemcpyAsync(z, dz, sizeof(double) * W * nsteps * P, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(u, du, sizeof(double) * W * nsteps, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  free_dev(dz); free_dev(du);
  if (e != cudaSuccess) return fail(ctx, MFIT_ERR_CUDA, "philox_stream: %s", cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------------ multi-GPU
int mfit_nccl_unique_id(void* id_out) {
  if (!id_out) return MFIT_ERR_ARG;
  std::string e;
  int rc = mfit_nccl_get_unique_id(id_out, e);
  if (rc) return fail(nullptr, rc, "%s", e.c_str());
  return 0;
}

int mfit_nccl_attach(mfit_ctx* ctx, const void* unique_id, int rank, int n_ranks) {
  if (!ctx || !unique_id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return MFIT_ERR_ARG;
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  // One communicator per (unique id, rank) and process, shared by every context that attaches with it: bringing a
  // communicator up costs tens of milliseconds (ncclCommInitRank is a collective with a bootstrap network round), which a
  // program that builds walker set after walker set must not pay each time.  It lives until the process ends.
  static std::mutex mu;
  static std::map<std::string, void*> comms;
  std::string key((const char*)unique_id, 128);
  key += "/" + std::to_string(rank) + "/" + std::to_string(n_ranks) + "/" + std::to_string(ctx->device);
  std::lock_guard<std::mutex> lock(mu);
  auto it = comms.find(key);
  if (it == comms.end()) {
    void* comm = nullptr;
    std::string e;
    int rc = mfit_nccl_comm_init(&comm, n_ranks, unique_id, rank, e);
    if (rc) return fail(ctx, rc, "%s", e.c_str());
    it = comms.emplace(key, comm).first;
  }
  ctx->comm = it->second;
  ctx->rank = rank;
  ctx->n_ranks = n_ranks;
  return 0;
}

// ---- peer exchange: data sets sharded over the GPUs of one box without a collective call on the step path
struct PeerHandle {   // MFIT_PEER_HANDLE_BYTES
  cudaIpcMemHandle_t ipc;        // 64 bytes
  unsigned long long ptr;        // the block's address in the exporting process
  long long pid;
  int device, rank, n_ranks, max_slots;
  long long max_stride;
};
static_assert(sizeof(PeerHandle) <= MFIT_PEER_HANDLE_BYTES, "PeerHandle does not fit");

int mfit_peer_setup(mfit_ctx* ctx, int rank, int n_ranks, int64_t max_chains, int max_datasets_per_rank, void* handle_out) {
  if (!ctx || !handle_out || n_ranks < 1 || n_ranks > 8 || rank < 0 || rank >= n_ranks || max_chains < 1 || max_datasets_per_rank < 1)
    return ctx ? fail(ctx, MFIT_ERR_ARG, "peer exchange: 1..8 ranks, rank in range, chains and data sets per rank >= 1") : MFIT_ERR_ARG;
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  Peers& p = ctx->peers;
  if (p.base) return fail(ctx, MFIT_ERR_STATE, "mfit_peer_setup has been called already");
  p.rank = rank;
  p.n = n_ranks;
  p.max_stride = round_up(max_chains, MFIT_CB);
  p.max_slots = max_datasets_per_rank;
  const size_t half = sizeof(double) * (size_t)n_ranks * p.max_slots * 2 * (size_t)p.max_stride;
  // plain cudaMalloc: memory of the stream-ordered pool cannot be exported to other processes
  CU((cudaMalloc)((void**)&p.base, kPeerHeader + 2 * half));
  CU(cudaMemsetAsync(p.base, 0, kPeerHeader + 2 * half, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  p.flags = reinterpret_cast<unsigned long long*>(p.base);
  p.xdone = reinterpret_cast<int*>(p.base + 512);
  p.xerr = p.xdone + 1;
  p.xch = reinterpret_cast<double*>(p.base + kPeerHeader);
  PeerHandle h;
  memset(&h, 0, sizeof h);
  CU(cudaIpcGetMemHandle(&h.ipc, p.base));
  h.ptr = (unsigned long long)(uintptr_t)p.base;
  h.pid = (long long)getpid();
  h.device = ctx->device;
  h.rank = rank;
  h.n_ranks = n_ranks;
  h.max_slots = p.max_slots;
  h.max_stride = p.max_stride;
  memset(handle_out, 0, MFIT_PEER_HANDLE_BYTES);
  memcpy(handle_out, &h, sizeof h);
  ctx->rank = rank;
  ctx->n_ranks = n_ranks;
  return 0;
}

int mfit_peer_attach(mfit_ctx* ctx, const void* handles) {
  if (!ctx || !handles) return MFIT_ERR_ARG;
  Peers& p = ctx->peers;
  if (!p.base) return fail(ctx, MFIT_ERR_STATE, "mfit_peer_setup has not been called");
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  for (int q = 0; q < p.n; q++) {
    PeerHandle h;
    memcpy(&h, (const char*)handles + (size_t)q * MFIT_PEER_HANDLE_BYTES, sizeof h);
    if (h.rank != q || h.n_ranks != p.n || h.max_stride != p.max_stride || h.max_slots != p.max_slots)
      return fail(ctx, MFIT_ERR_ARG, "peer handle %d does not match this context's mfit_peer_setup", q);
    char* base = nullptr;
    if (q == p.rank) base = p.base;
    else if (h.pid == (long long)getpid()) {   // the peer lives in this process: its pointer is valid here once peer access is on
      if (h.device != ctx->device) {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, ctx->device, h.device));
        if (!can) return fail(ctx, MFIT_ERR_CUDA, "device %d cannot access device %d", ctx->device, h.device);
        cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, MFIT_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
      }
      base = reinterpret_cast<char*>((uintptr_t)h.ptr);
    } else {
      void* m = nullptr;
      CU(cudaIpcOpenMemHandle(&m, h.ipc, cudaIpcMemLazyEnablePeerAccess));
      p.opened[q] = m;
      base = (char*)m;
    }
    p.flag_peer[q] = reinterpret_cast<unsigned long long*>(base);
    p.xpeer[q] = reinterpret_cast<double*>(base + kPeerHeader);
  }
  p.ready = true;
  if (ctx->have_chains) chains_free(ctx);
  return 0;
}

int mfit_dataset_owners(mfit_ctx* ctx, const int32_t* owner) {
  if (!ctx || !owner) return MFIT_ERR_ARG;
  if (ctx->D <= 0) return fail(ctx, MFIT_ERR_STATE, "mfit_problem_define has not been called");
  if (!ctx->comm && !ctx->peers.ready) return fail(ctx, MFIT_ERR_STATE, "neither mfit_nccl_attach nor mfit_peer_attach has been called");
  g_alloc_stream = ctx->stream;
  std::vector<int> per_rank(ctx->n_ranks, 0);
  for (int d = 0; d < ctx->D; d++) {
    if (owner[d] < 0 || owner[d] >= ctx->n_ranks) return fail(ctx, MFIT_ERR_ARG, "owner[%d]=%d out of range", d, owner[d]);
    ctx->ds[d].owner = owner[d];
    ctx->ds[d].xslot = per_rank[owner[d]]++;
    ctx->ds[d].owned = owner[d] == ctx->rank;
  }
  int slots = 1;
  for (int v : per_rank) slots = v > slots ? v : slots;
  ctx->xslots = slots;
  ctx->sharded = true;
  if (ctx->have_chains) chains_free(ctx);
  return 0;
}

int mfit_nccl_allgather_f64(mfit_ctx* ctx, const double* local_host, int64_t count, double* out_host) {
  if (!ctx || !local_host || !out_host || count < 0) return MFIT_ERR_ARG;
  if (!ctx->comm) return fail(ctx, MFIT_ERR_STATE, "mfit_nccl_attach has not been called");
  if (count == 0) return 0;
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  double* buf = nullptr;
  const size_t per = sizeof(double) * (size_t)count;
  CU(cudaMalloc(&buf, per * ctx->n_ranks));
  int rc = 0;
  cudaError_t e = cudaMemcpyAsync((char*)buf + per * ctx->rank, local_host, per, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    std::string er;
    rc = mfit_nccl_allgather_bytes(ctx->comm, (char*)buf + per * ctx->rank, buf, per, ctx->stream, er);
    if (rc) fail(ctx, rc, "%s", er.c_str());
  }
  if (e == cudaSuccess && rc == 0) e = cudaMemcpyAsync(out_host, buf, per * ctx->n_ranks, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && rc == 0) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(buf);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(ctx, MFIT_ERR_CUDA, "allgather: %s", cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------------ measurement
int mfit_measure_fp64(mfit_ctx* ctx, double* dfma_flops, double* cost_exp, double* cost_log, double* cost_sin,
                      double* cost_div) {
  if (!ctx) return MFIT_ERR_ARG;
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  if (mfit_measure_fp64_dev(ctx->stream, dfma_flops, cost_exp, cost_log, cost_sin, cost_div) != 0)
    return fail(ctx, MFIT_ERR_CUDA, "fp64 micro-benchmark failed: %s", cudaGetErrorString(cudaGetLastError()));
  ctx->launches += 5 * 4;
  return 0;
}

int mfit_set_timing(mfit_ctx* ctx, int on) {
  if (!ctx) return MFIT_ERR_ARG;
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  ctx->timing = on != 0;
  if (ctx->timing && !ctx->mode_dev) {
    CU(cudaMalloc(&ctx->mode_dev, 4 * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(ctx->mode_dev, 0, 4 * sizeof(unsigned long long), ctx->stream));
  }
  if (ctx->timing && ctx->ev.empty()) {
    ctx->ev.resize(4096);
    for (auto& p : ctx->ev) { CU(cudaEventCreate(&p.a)); CU(cudaEventCreate(&p.b)); }
  }
  return 0;
}

int mfit_eval_mode_tiles(mfit_ctx* ctx, int64_t* tiles4) {
  if (!ctx || !tiles4) return MFIT_ERR_ARG;
  for (int k = 0; k < 4; k++) tiles4[k] = 0;
  if (!ctx->mode_dev) return 0;
  CU(cudaSetDevice(ctx->device));
  unsigned long long h[4];
  CU(cudaMemcpyAsync(h, ctx->mode_dev, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < 4; k++) tiles4[k] = (int64_t)h[k];
  CU(cudaMemsetAsync(ctx->mode_dev, 0, sizeof h, ctx->stream));
  return 0;
}

int mfit_eval_kernel_time(mfit_ctx* ctx, double* seconds, int64_t* launches) {
  if (!ctx) return MFIT_ERR_ARG;
  if (seconds) *seconds = ctx->eval_seconds;
  if (launches) *launches = ctx->eval_launches;
  return 0;
}

int mfit_set_small_path(mfit_ctx* ctx, int mode) {
  if (!ctx || mode < 0 || mode > 2) return MFIT_ERR_ARG;
  ctx->small_mode = mode;
  return 0;
}

int mfit_set_groups(mfit_ctx* ctx, int n_groups) {
  if (!ctx || n_groups < 0 || n_groups > mfit_ctx::kMaxGroups) return MFIT_ERR_ARG;
  ctx->groups_override = n_groups;
  return 0;
}

int mfit_flush_l2(mfit_ctx* ctx) {
  if (!ctx) return MFIT_ERR_ARG;
  CU(cudaSetDevice(ctx->device)); g_alloc_stream = ctx->stream;
  if (!ctx->flush_buf) {
    ctx->flush_bytes = (size_t)256 << 20;  // 2x the 126 MB L2
    CU(cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
  }
  mfit_launch_flush(ctx->flush_buf, ctx->flush_bytes, ctx->stream);
  ctx->launches++;
  return 0;
}

}  // extern "C"
