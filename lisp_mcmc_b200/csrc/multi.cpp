// multi.cpp -- several GPUs of one box behind ONE handle (mfit_multi_*, include/mfit_cuda.h).
//
// The reference advances a walker set with `mapc` on one core (nv-specific.lisp:58-66) and has no notion of a second
// worker; SURVEY.md 8(b)/(e) asks for the sharding to live behind the C ABI so that the SBCL host (`:devices`) gets all
// GPUs without Python, MPI or a second process.  This file is that fan-out and nothing else: one mfit_ctx per device, one
// persistent host thread per context (CUDA's current device and the library's allocation stream are per thread), every
// call posted to all of them and joined.  Two shardings, both bit-identical to one GPU:
//   chains    W chains split contiguously (the same split as lisp_mcmc_b200.walker.shard_chains); Philox streams are keyed
//             by global chain id; nothing crosses GPUs during a run; per-chain outputs land in the caller's arrays at the
//             shard's offset, i.e. in global chain order.
//   datasets  data set d lives on GPU d mod n (walker-make-step's sum over data sets, mcmc-fitting.lisp:1067-1070, split);
//             every GPU holds all chains and replays identical proposals; the per-data-set terms travel through the peer
//             exchange (k_finalize stores into the peers' blocks over NVLink, k_post waits on flags), which inside one
//             process needs no IPC handle: the contexts exchange raw pointers.  The n host threads MUST run concurrently
//             here (each GPU's accept kernel waits for the others' terms), which is why the workers are threads and not a loop.
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mfit_internal.h"

namespace {

struct Worker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> job;
  bool has_job = false, done = false, quit = false;
  int rc = 0;
};

}  // namespace

struct mfit_multi {
  int n = 0, mode = MFIT_SHARD_CHAINS;
  std::vector<int> dev;
  std::vector<mfit_ctx*> ctx;
  std::vector<Worker*> workers;
  std::string err;
  int D = 0, P = 0;
  int64_t W = 0;
  int64_t reserve_chains = 131072;     // peer exchange block (datasets mode): chains it is sized for
  std::vector<int64_t> lo, cnt;        // chain shard of every rank (chains mode); datasets mode: 0, W everywhere
  std::vector<char> open;              // run_begin made, not finished (per rank)
  std::atomic<long long> stop_at{-1};  // datasets mode: the poll boundary at which all ranks leave the step loop
  bool peers_ready = false;
  int reserved_slots = 0;              // data sets per GPU the exchange blocks were sized for
};

namespace {

std::string g_multi_create_error;

int mfail(mfit_multi* m, int code, const char* fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (m) m->err = buf;
  else g_multi_create_error = buf;
  return code;
}

void worker_main(Worker* w) {
  std::unique_lock<std::mutex> lock(w->mu);
  for (;;) {
    w->cv.wait(lock, [&] { return w->has_job || w->quit; });
    if (w->quit) return;
    std::function<int()> job = std::move(w->job);
    w->has_job = false;
    lock.unlock();
    int rc = job();
    lock.lock();
    w->rc = rc;
    w->done = true;
    w->cv.notify_all();
  }
}

// fn(rank) on the worker thread of every rank in `ranks` (all when empty), concurrently; first failing rank's status and
// message are kept
int run_on(mfit_multi* m, const std::vector<int>& ranks, const std::function<int(int)>& fn) {
  std::vector<int> rs = ranks;
  if (rs.empty()) for (int r = 0; r < m->n; r++) rs.push_back(r);
  for (int r : rs) {
    Worker* w = m->workers[r];
    std::lock_guard<std::mutex> lock(w->mu);
    w->job = [fn, r]() { return fn(r); };
    w->has_job = true;
    w->done = false;
    w->cv.notify_all();
  }
  int first = 0;
  for (int r : rs) {
    Worker* w = m->workers[r];
    std::unique_lock<std::mutex> lock(w->mu);
    w->cv.wait(lock, [&] { return w->done; });
    if (w->rc != 0 && first == 0) {
      first = w->rc;
      m->err = "GPU " + std::to_string(m->dev[r]) + " (rank " + std::to_string(r) + "): " + mfit_last_error(m->ctx[r]);
    }
  }
  return first;
}

int run_all(mfit_multi* m, const std::function<int(int)>& fn) { return run_on(m, {}, fn); }

// ranks that hold chains (chains mode: a rank's shard may be empty when W < n)
std::vector<int> chain_ranks(mfit_multi* m) {
  std::vector<int> rs;
  for (int r = 0; r < m->n; r++) if (m->cnt[r] > 0) rs.push_back(r);
  return rs;
}

void shard(mfit_multi* m, int64_t W) {
  m->W = W;
  for (int r = 0; r < m->n; r++) {
    if (m->mode == MFIT_SHARD_CHAINS) {
      const int64_t base = W / m->n, rem = W % m->n;
      m->lo[r] = r * base + (r < rem ? r : rem);
      m->cnt[r] = base + (r < rem ? 1 : 0);
    } else {
      m->lo[r] = 0;
      m->cnt[r] = W;
    }
  }
}

bool owns(mfit_multi* m, int r, int d) { return m->mode == MFIT_SHARD_CHAINS || m->n == 1 || d % m->n == r; }

int need_chains(mfit_multi* m) {
  if (!m) return MFIT_ERR_ARG;
  if (m->W <= 0) return mfail(m, MFIT_ERR_STATE, "mfit_multi_chains_init has not been called");
  return 0;
}

}  // namespace

extern "C" {

int mfit_multi_create(int n_gpus, const int* dev_ids, int shard_mode, mfit_multi** out) {
  if (!out) return mfail(nullptr, MFIT_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (n_gpus < 1 || n_gpus > 8) return mfail(nullptr, MFIT_ERR_ARG, "1..8 GPUs, got %d", n_gpus);
  if (shard_mode != MFIT_SHARD_CHAINS && shard_mode != MFIT_SHARD_DATASETS)
    return mfail(nullptr, MFIT_ERR_ARG, "shard_mode must be MFIT_SHARD_CHAINS or MFIT_SHARD_DATASETS");
  for (int a = 0; a < n_gpus; a++)
    for (int b = a + 1; dev_ids && b < n_gpus; b++)
      if (dev_ids[a] == dev_ids[b]) return mfail(nullptr, MFIT_ERR_ARG, "device %d listed twice", dev_ids[a]);
  mfit_multi* m = new mfit_multi();
  m->n = n_gpus;
  m->mode = shard_mode;
  m->lo.assign(n_gpus, 0);
  m->cnt.assign(n_gpus, 0);
  m->open.assign(n_gpus, 0);
  for (int r = 0; r < n_gpus; r++) {
    m->dev.push_back(dev_ids ? dev_ids[r] : r);
    mfit_ctx* c = nullptr;
    int rc = mfit_ctx_create(m->dev[r], nullptr, &c);  // one after the other: the create-error string is process-wide
    if (rc) {
      mfail(nullptr, rc, "GPU %d: %s", m->dev[r], mfit_last_error(nullptr));
      for (mfit_ctx* p : m->ctx) mfit_ctx_destroy(p);
      delete m;
      return rc;
    }
    m->ctx.push_back(c);
  }
  for (int r = 0; r < n_gpus; r++) {
    Worker* w = new Worker();
    w->th = std::thread(worker_main, w);
    m->workers.push_back(w);
  }
  *out = m;
  return 0;
}

void mfit_multi_destroy(mfit_multi* m) {
  if (!m) return;
  run_all(m, [m](int r) { mfit_ctx_destroy(m->ctx[r]); return 0; });
  for (Worker* w : m->workers) {
    { std::lock_guard<std::mutex> lock(w->mu); w->quit = true; w->cv.notify_all(); }
    w->th.join();
    delete w;
  }
  delete m;
}

const char* mfit_multi_last_error(mfit_multi* m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }
int mfit_multi_size(mfit_multi* m) { return m ? m->n : 0; }
mfit_ctx* mfit_multi_ctx(mfit_multi* m, int rank) { return (m && rank >= 0 && rank < m->n) ? m->ctx[rank] : nullptr; }

int mfit_multi_reserve(mfit_multi* m, int64_t max_chains) {
  if (!m || max_chains < 1) return m ? mfail(m, MFIT_ERR_ARG, "max_chains >= 1") : MFIT_ERR_ARG;
  if (m->peers_ready) return mfail(m, MFIT_ERR_STATE, "the exchange blocks exist already: call before mfit_multi_problem_define");
  m->reserve_chains = max_chains;
  return 0;
}

int mfit_multi_problem_define(mfit_multi* m, int n_datasets, int n_params) {
  if (!m) return MFIT_ERR_ARG;
  int rc = run_all(m, [=](int r) { return mfit_problem_define(m->ctx[r], n_datasets, n_params); });
  if (rc) return rc;
  m->D = n_datasets;
  m->P = n_params;
  m->W = 0;
  if (m->mode == MFIT_SHARD_DATASETS && m->n > 1) {
    const int slots = (n_datasets + m->n - 1) / m->n;
    if (m->peers_ready && slots > m->reserved_slots)
      return mfail(m, MFIT_ERR_ARG, "%d data sets need %d exchange slots per GPU, the blocks of this handle hold %d: make a new handle",
                   n_datasets, slots, m->reserved_slots);
    if (!m->peers_ready) {
      std::vector<char> handles((size_t)m->n * MFIT_PEER_HANDLE_BYTES);
      rc = run_all(m, [&](int r) {
        return mfit_peer_setup(m->ctx[r], r, m->n, m->reserve_chains, slots > 0 ? slots : 1, handles.data() + (size_t)r * MFIT_PEER_HANDLE_BYTES);
      });
      if (rc) return rc;
      rc = run_all(m, [&](int r) {
        mfit_internal_stop_word(m->ctx[r], &m->stop_at);
        return mfit_peer_attach(m->ctx[r], handles.data());
      });
      if (rc) return rc;
      m->peers_ready = true;
      m->reserved_slots = slots > 0 ? slots : 1;
    }
    std::vector<int32_t> owner(n_datasets);
    for (int d = 0; d < n_datasets; d++) owner[d] = d % m->n;
    rc = run_all(m, [&](int r) { return mfit_dataset_owners(m->ctx[r], owner.data()); });
    if (rc) return rc;
  }
  return 0;
}

int mfit_multi_dataset_upload(mfit_multi* m, int d, const double* x, int x_dims, const double* y, const double* sigma, int64_t n) {
  if (!m) return MFIT_ERR_ARG;
  std::vector<int> rs;
  for (int r = 0; r < m->n; r++) if (owns(m, r, d)) rs.push_back(r);
  if (rs.empty()) return mfail(m, MFIT_ERR_ARG, "data set %d out of range", d);
  return run_on(m, rs, [=](int r) { return mfit_dataset_upload_xd(m->ctx[r], d, x, x_dims, y, sigma, n); });
}

int mfit_multi_model_builtin(mfit_multi* m, int d, int family, int degree, const int32_t* param_index, int n_idx) {
  if (!m) return MFIT_ERR_ARG;
  return run_all(m, [=](int r) { return mfit_model_builtin(m->ctx[r], d, family, degree, param_index, n_idx); });
}
int mfit_multi_model_sexpr(mfit_multi* m, int d, const char* lambda_text, const char* const* key_names, int n_keys) {
  if (!m) return MFIT_ERR_ARG;
  return run_all(m, [=](int r) { return mfit_model_sexpr(m->ctx[r], d, lambda_text, key_names, n_keys); });
}
int mfit_multi_likelihood_set(mfit_multi* m, int d, int kind) {
  if (!m) return MFIT_ERR_ARG;
  return run_all(m, [=](int r) { return mfit_likelihood_set(m->ctx[r], d, kind); });
}
int mfit_multi_likelihood_sexpr(mfit_multi* m, int d, const char* lambda_text) {
  if (!m) return MFIT_ERR_ARG;
  return run_all(m, [=](int r) { return mfit_likelihood_sexpr(m->ctx[r], d, lambda_text); });
}
int mfit_multi_prior_flat(mfit_multi* m, int d) {
  if (!m) return MFIT_ERR_ARG;
  return run_all(m, [=](int r) { return mfit_prior_flat(m->ctx[r], d); });
}
int mfit_multi_prior_bounds(mfit_multi* m, int d, int n, const int32_t* idx, const double* lo, const double* hi) {
  if (!m) return MFIT_ERR_ARG;
  return run_all(m, [=](int r) { return mfit_prior_bounds(m->ctx[r], d, n, idx, lo, hi); });
}
int mfit_multi_prior_sexpr(mfit_multi* m, int d, const char* lambda_text, const char* const* key_names, int n_keys) {
  if (!m) return MFIT_ERR_ARG;
  return run_all(m, [=](int r) { return mfit_prior_sexpr(m->ctx[r], d, lambda_text, key_names, n_keys); });
}

// walker-make-step over a batch: chains mode splits the batch like the chains; datasets mode evaluates every theta on
// every GPU (each its own data sets, terms exchanged) and rank 0's copy is returned
int mfit_multi_eval_logp(mfit_multi* m, const double* theta, int64_t W, double* prob, double* ll, double* lp) {
  if (!m) return MFIT_ERR_ARG;
  if (W < 0 || (W > 0 && !theta)) return mfail(m, MFIT_ERR_ARG, "bad theta/W");
  const int P = m->P, D = m->D;
  if (m->mode == MFIT_SHARD_DATASETS)
    return run_all(m, [=](int r) { return mfit_eval_logp(m->ctx[r], theta, W, r == 0 ? prob : nullptr, r == 0 ? ll : nullptr, r == 0 ? lp : nullptr); });
  std::vector<int> rs;
  std::vector<int64_t> lo(m->n), cnt(m->n);
  for (int r = 0; r < m->n; r++) {
    const int64_t base = W / m->n, rem = W % m->n;
    lo[r] = r * base + (r < rem ? r : rem);
    cnt[r] = base + (r < rem ? 1 : 0);
    if (cnt[r] > 0) rs.push_back(r);
  }
  if (rs.empty()) return 0;
  return run_on(m, rs, [&](int r) {
    return mfit_eval_logp(m->ctx[r], theta + lo[r] * P, cnt[r], prob ? prob + lo[r] : nullptr, ll ? ll + lo[r] * D : nullptr,
                          lp ? lp + lo[r] * D : nullptr);
  });
}

int mfit_multi_chains_init(mfit_multi* m, const double* theta0, int64_t W, uint64_t seed, int64_t trace_capacity) {
  if (!m) return MFIT_ERR_ARG;
  if (W < 1 || !theta0) return mfail(m, MFIT_ERR_ARG, "bad theta0/W");
  if (m->mode == MFIT_SHARD_DATASETS && m->n > 1 && W > m->reserve_chains)
    return mfail(m, MFIT_ERR_ARG, "%lld chains exceed the %lld the exchange blocks were sized for (mfit_multi_reserve before mfit_multi_problem_define)",
                 (long long)W, (long long)m->reserve_chains);
  shard(m, W);
  std::fill(m->open.begin(), m->open.end(), 0);
  const int P = m->P;
  return run_on(m, chain_ranks(m), [=](int r) {
    return mfit_chains_init(m->ctx[r], theta0 + m->lo[r] * P, m->cnt[r], seed, m->lo[r], trace_capacity);
  });
}

int mfit_multi_run_adaptive(mfit_multi* m, const mfit_adapt_cfg* cfg, const double* l_matrix, int64_t* steps_launched) {
  int rc = need_chains(m);
  if (rc) return rc;
  m->stop_at.store(-1);
  std::vector<int64_t> steps(m->n, 0);
  rc = run_on(m, chain_ranks(m), [&](int r) { return mfit_run_adaptive(m->ctx[r], cfg, l_matrix, nullptr, nullptr, 0, &steps[r]); });
  std::fill(m->open.begin(), m->open.end(), 0);
  if (steps_launched) { *steps_launched = 0; for (int64_t s : steps) if (s > *steps_launched) *steps_launched = s; }
  return rc;
}

int mfit_multi_run_begin(mfit_multi* m, const mfit_adapt_cfg* cfg, const double* l_matrix) {
  int rc = need_chains(m);
  if (rc) return rc;
  m->stop_at.store(-1);
  rc = run_on(m, chain_ranks(m), [=](int r) { return mfit_run_begin(m->ctx[r], cfg, l_matrix, nullptr, nullptr, 0); });
  for (int r = 0; r < m->n; r++) m->open[r] = (rc == 0 && m->cnt[r] > 0) ? 1 : 0;
  return rc;
}

// up to max_steps more iterations on every GPU whose chains are still in the loop; *steps_done = the most any GPU made
// (datasets mode: all make the same number), *finished = every GPU has left the loop
int mfit_multi_run_advance(mfit_multi* m, int64_t max_steps, int64_t* steps_done, int* finished) {
  int rc = need_chains(m);
  if (rc) return rc;
  std::vector<int> rs;
  for (int r = 0; r < m->n; r++) if (m->open[r]) rs.push_back(r);
  std::vector<int64_t> done(m->n, 0);
  std::vector<int> fin(m->n, 1);
  if (!rs.empty()) {
    rc = run_on(m, rs, [&](int r) { return mfit_run_advance(m->ctx[r], max_steps, &done[r], &fin[r]); });
    if (rc) { std::fill(m->open.begin(), m->open.end(), 0); return rc; }
  }
  int all = 1;
  int64_t most = 0;
  for (int r = 0; r < m->n; r++) {
    if (fin[r]) m->open[r] = 0; else all = 0;
    if (done[r] > most) most = done[r];
  }
  if (steps_done) *steps_done = most;
  if (finished) *finished = all;
  return 0;
}

int mfit_multi_run_many_steps(mfit_multi* m, int64_t n, const double* l_matrix) {
  int rc = need_chains(m);
  if (rc) return rc;
  m->stop_at.store(-1);
  return run_on(m, chain_ranks(m), [=](int r) { return mfit_run_many_steps(m->ctx[r], n, l_matrix, nullptr, nullptr, 0); });
}

// mfit-walker-estop :860.  Called from the host thread while another thread may sit in mfit_multi_run_advance: only
// atomics are touched.  Chains mode: every GPU stops on its own; datasets mode: the GPUs agree on a step (api.cu).
int mfit_multi_request_stop(mfit_multi* m) {
  if (!m) return MFIT_ERR_ARG;
  for (mfit_ctx* c : m->ctx) mfit_request_stop(c);
  return 0;
}

int mfit_multi_trace_read(mfit_multi* m, int64_t take, double* theta, double* prob, int64_t* n_valid) {
  int rc = need_chains(m);
  if (rc) return rc;
  if (take < 0) return mfail(m, MFIT_ERR_ARG, "take < 0");
  const int P = m->P;
  if (m->mode == MFIT_SHARD_DATASETS) return run_on(m, {0}, [=](int r) { return mfit_trace_read(m->ctx[r], take, theta, prob, n_valid); });
  return run_on(m, chain_ranks(m), [=](int r) {
    return mfit_trace_read(m->ctx[r], take, theta ? theta + m->lo[r] * take * P : nullptr, prob ? prob + m->lo[r] * take : nullptr,
                           n_valid ? n_valid + m->lo[r] : nullptr);
  });
}

int mfit_multi_stats_read(mfit_multi* m, mfit_chain_stats* stats) {
  int rc = need_chains(m);
  if (rc) return rc;
  if (!stats) return mfail(m, MFIT_ERR_ARG, "stats is NULL");
  if (m->mode == MFIT_SHARD_DATASETS) return run_on(m, {0}, [=](int r) { return mfit_stats_read(m->ctx[r], stats); });
  return run_on(m, chain_ranks(m), [=](int r) { return mfit_stats_read(m->ctx[r], stats + m->lo[r]); });
}

int mfit_multi_best_read(mfit_multi* m, double* theta, double* prob) {
  int rc = need_chains(m);
  if (rc) return rc;
  const int P = m->P;
  if (m->mode == MFIT_SHARD_DATASETS) return run_on(m, {0}, [=](int r) { return mfit_best_read(m->ctx[r], theta, prob); });
  return run_on(m, chain_ranks(m), [=](int r) {
    return mfit_best_read(m->ctx[r], theta ? theta + m->lo[r] * P : nullptr, prob ? prob + m->lo[r] : nullptr);
  });
}

int mfit_multi_last_read(mfit_multi* m, double* theta, double* prob) {
  int rc = need_chains(m);
  if (rc) return rc;
  const int P = m->P;
  if (m->mode == MFIT_SHARD_DATASETS) return run_on(m, {0}, [=](int r) { return mfit_last_read(m->ctx[r], theta, prob); });
  return run_on(m, chain_ranks(m), [=](int r) {
    return mfit_last_read(m->ctx[r], theta ? theta + m->lo[r] * P : nullptr, prob ? prob + m->lo[r] : nullptr);
  });
}

int mfit_multi_l_matrix_read(mfit_multi* m, double* l) {
  int rc = need_chains(m);
  if (rc) return rc;
  if (!l) return mfail(m, MFIT_ERR_ARG, "l is NULL");
  const int P = m->P;
  if (m->mode == MFIT_SHARD_DATASETS) return run_on(m, {0}, [=](int r) { return mfit_l_matrix_read(m->ctx[r], l); });
  return run_on(m, chain_ranks(m), [=](int r) { return mfit_l_matrix_read(m->ctx[r], l + m->lo[r] * P * P); });
}

int64_t mfit_multi_launch_count(mfit_multi* m) {
  int64_t s = 0;
  if (m) for (mfit_ctx* c : m->ctx) s += mfit_launch_count(c);
  return s;
}

}  // extern "C"

// ---- the trace consumers and editors: per-chain work, so chain shards are served in place at their offsets;
// with sharded data sets every GPU holds all chains (and must edit them alike), and rank 0 -- for the fit envelope the
// owner of the data set -- answers
extern "C" {

int mfit_multi_trace_analytics(mfit_multi* m, int64_t take, double* most_likely_theta, double* most_likely_prob, double* median,
                               double* acceptance, double* covariance, double* l_matrix) {
  int rc = need_chains(m);
  if (rc) return rc;
  const int P = m->P;
  if (m->mode == MFIT_SHARD_DATASETS)
    return run_on(m, {0}, [=](int r) {
      return mfit_trace_analytics(m->ctx[r], take, most_likely_theta, most_likely_prob, median, acceptance, covariance, l_matrix);
    });
  return run_on(m, chain_ranks(m), [=](int r) {
    const int64_t o = m->lo[r];
    return mfit_trace_analytics(m->ctx[r], take, most_likely_theta ? most_likely_theta + o * P : nullptr,
                                most_likely_prob ? most_likely_prob + o : nullptr, median ? median + o * P : nullptr,
                                acceptance ? acceptance + o : nullptr, covariance ? covariance + o * P * P : nullptr,
                                l_matrix ? l_matrix + o * P * P : nullptr);
  });
}

int mfit_multi_fit_envelope(mfit_multi* m, int d, int64_t take, int which_solution, const double* x_fit, int n_fit, double* y_fit,
                            double* y_max, double* y_min, double* params, int64_t* n_distinct) {
  int rc = need_chains(m);
  if (rc) return rc;
  if (d < 0 || d >= m->D) return mfail(m, MFIT_ERR_ARG, "data set %d out of range", d);
  const int P = m->P;
  if (m->mode == MFIT_SHARD_DATASETS)
    return run_on(m, {m->n > 1 ? d % m->n : 0}, [=](int r) {
      return mfit_fit_envelope(m->ctx[r], d, take, which_solution, x_fit, n_fit, y_fit, y_max, y_min, params, n_distinct);
    });
  return run_on(m, chain_ranks(m), [=](int r) {
    const int64_t o = m->lo[r];
    return mfit_fit_envelope(m->ctx[r], d, take, which_solution, x_fit, n_fit, y_fit ? y_fit + o * n_fit : nullptr,
                             y_max ? y_max + o * n_fit : nullptr, y_min ? y_min + o * n_fit : nullptr,
                             params ? params + o * P : nullptr, n_distinct ? n_distinct + o : nullptr);
  });
}

int mfit_multi_walker_modify(mfit_multi* m, int mode, int64_t number) {
  int rc = need_chains(m);
  if (rc) return rc;
  return run_on(m, chain_ranks(m), [=](int r) { return mfit_walker_modify(m->ctx[r], mode, number); });
}

int mfit_multi_keep_walks(mfit_multi* m, int64_t keep) {
  int rc = need_chains(m);
  if (rc) return rc;
  return run_on(m, chain_ranks(m), [=](int r) { return mfit_keep_walks(m->ctx[r], keep); });
}

int mfit_multi_chains_restore(mfit_multi* m, int64_t rows, const double* theta, const double* prob, const int64_t* n_valid,
                              const int64_t* age, const double* best_theta, const double* best_prob) {
  int rc = need_chains(m);
  if (rc) return rc;
  const int P = m->P;
  return run_on(m, chain_ranks(m), [=](int r) {
    const int64_t o = m->lo[r];   // 0 when every GPU holds all chains
    return mfit_chains_restore(m->ctx[r], rows, theta ? theta + o * rows * P : nullptr, prob ? prob + o * rows : nullptr,
                               n_valid ? n_valid + o : nullptr, age ? age + o : nullptr,
                               best_theta ? best_theta + o * P : nullptr, best_prob ? best_prob + o : nullptr);
  });
}

}  // extern "C"
