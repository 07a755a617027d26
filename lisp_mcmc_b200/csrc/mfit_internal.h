// mfit_internal.h -- shared declarations inside libmfit_cuda.so
#ifndef MFIT_INTERNAL_H
#define MFIT_INTERNAL_H
#include <stddef.h>
#include <stdint.h>
#include "../../include/mfit_cuda.h"

const void* mfit_builtin_eval_kernel(int family, int degree, int ll);
int mfit_family_nparams(int family, int degree);
size_t mfit_eval_smem_bytes(int ll);

#endif
