// mfit_internal.h -- shared declarations inside libmfit_cuda.so
#ifndef MFIT_INTERNAL_H
#define MFIT_INTERNAL_H
#include <stddef.h>
#include <stdint.h>
#include "../../include/mfit_cuda.h"

// prep_kernels.cu: per-block (shared data) or per-row (walker sets) results of the upload pass
#define MFIT_PREP_THREADS 256
struct MfitPrepBlock {
  double sum, err;      /* compensated sum of the block's terms and the collected rounding errors */
  double xmin, xmax, ymax;
  int bad_x, bad_pois;  /* some x not finite / some y not a count */
};

#ifdef __cplusplus
#include <atomic>
#include <string>
#include <vector>

// api.cu, for multi.cpp
void mfit_internal_stop_word(mfit_ctx* ctx, std::atomic<long long>* word);

// eval_builtin.cu
const void* mfit_builtin_eval_kernel(int family, int degree, int ll);
// eval_builtin_ll.cu, one object per likelihood kind
const void* mfit_builtin_eval_kernel_normal(int family, int degree);
const void* mfit_builtin_eval_kernel_cutoff(int family, int degree);
const void* mfit_builtin_eval_kernel_poisson(int family, int degree);
int mfit_family_nparams(int family, int degree);
size_t mfit_eval_smem_bytes(int ll, int x_dims, int warps, int stages);
// eval_perchain.cu
const void* mfit_builtin_perchain_kernel(int family, int degree, int ll, int staged);
// eval_perchain_reg.cu
const void* mfit_builtin_perchain_reg_kernel_normal(int family, int degree, int npt);
const void* mfit_builtin_perchain_reg_kernel_poisson(int family, int degree, int npt);

// eval_fit.cu
const void* mfit_builtin_fit_kernel(int family, int degree);

#ifdef __CUDACC__
void mfit_launch_prepare(const double* x, const double* y, double* yw, double* w, double* c, long long n, long long n_pad,
                         long long rows, long long x_rows, long long sig_rows, int have_sigma, MfitPrepBlock* out, int blocks,
                         cudaStream_t st);
void mfit_launch_logfact_fill(const double* y, const double* table, double* lf, long long n, long long n_pad, cudaStream_t st);
void mfit_launch_poisson_const(const double* y, const double* table, long long n, long long n_pad, long long rows,
                               MfitPrepBlock* out, int blocks, cudaStream_t st);
#endif

// lower.cpp
// x_dims (may be null): 0 when the model uses x as a number, else the number of components of x it reads
int mfit_lower_to_string(const char* text, const char* const* keys, int n_keys, int kind, std::string& out,
                         std::string& err, int* x_dims = nullptr);
// nvrtc_glue.cpp
std::string mfit_user_model_source(const std::string& model_struct);
std::string mfit_user_prior_source(const std::string& prior_fn);
std::string mfit_user_term_source(const std::string& model_struct, const std::string& model_type, const std::string& term_code,
                                  const double* err_head, int n_err);
// lower.cpp: MFIT_LL_* when the likelihood lambda is one of the built-in kinds spelled out, e.g.
// (lambda (y model error) (log-poisson model y)); -1 otherwise
int mfit_ll_canonical_kind(const char* text);
// eval_builtin.cu: C++ type of a built-in family ("MfitGaussLin", "MfitPoly<3>" ...), empty if unknown
std::string mfit_builtin_model_type(int family, int degree);
int mfit_nvrtc_compile(const std::string& src, std::vector<char>& cubin, std::string& err);

// nccl_glue.cpp (libnccl is dlopen'ed on first use)
int mfit_nccl_get_unique_id(void* id128, std::string& err);
int mfit_nccl_comm_init(void** comm, int n_ranks, const void* id128, int rank, std::string& err);
int mfit_nccl_allgather_bytes(void* comm, const void* send, void* recv, size_t bytes_per_rank, void* stream,
                              std::string& err);
void mfit_nccl_comm_destroy(void* comm);
#endif
#endif
