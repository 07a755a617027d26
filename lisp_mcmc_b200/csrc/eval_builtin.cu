// eval_builtin.cu -- look-up of the nvcc-compiled log-likelihood kernels of the built-in model families x likelihoods
// (eval_builtin_ll.cu; the user s-expression path instantiates the same template via NVRTC) and what the host needs to know about them.
#include "mfit_eval.cuh"
#include "mfit_models.cuh"
#include "mfit_internal.h"

// the kernels themselves are instantiated in eval_builtin_ll.cu, one object per likelihood kind so that the build compiles them side by side
const void* mfit_builtin_eval_kernel(int family, int degree, int ll) {
  switch (ll) {
    case MFITK_NORMAL: return mfit_builtin_eval_kernel_normal(family, degree);
    case MFITK_CUTOFF: return mfit_builtin_eval_kernel_cutoff(family, degree);
    case MFITK_POISSON: return mfit_builtin_eval_kernel_poisson(family, degree);
  }
  return nullptr;
}

std::string mfit_builtin_model_type(int family, int degree) {
  switch (family) {
    case MFIT_FAM_POLY: return (degree >= 0 && degree <= 7) ? "MfitPoly<" + std::to_string(degree) + ">" : "";
    case MFIT_FAM_EXPO: return "MfitExpo";
    case MFIT_FAM_SINUS: return "MfitSinus";
    case MFIT_FAM_LORENTZ: return "MfitLorentz";
    case MFIT_FAM_GAUSS_LIN: return "MfitGaussLin";
    case MFIT_FAM_NV_MULTI: return (degree >= 1 && degree <= 4) ? "MfitNvMulti<" + std::to_string(degree) + ">" : "";
  }
  return "";
}

int mfit_family_nparams(int family, int degree) {
  switch (family) {
    case MFIT_FAM_POLY: return degree + 1;
    case MFIT_FAM_EXPO: return 3;
    case MFIT_FAM_SINUS: return 4;
    case MFIT_FAM_LORENTZ: return 4;
    case MFIT_FAM_GAUSS_LIN: return 5;
    case MFIT_FAM_NV_MULTI: return 2 * degree + 2;
  }
  return -1;
}

// dynamic shared memory of the shared-data kernel: per warp `stages` stages of narr arrays, one mbarrier per stage, the item ring
size_t mfit_eval_smem_bytes(int ll, int x_dims, int warps, int stages) {
  int narr = (ll == MFIT_LL_NORMAL ? 3 : ll == MFIT_LL_NORMAL_CUTOFF ? 4 : 2) + x_dims - 1;
  return (size_t)warps * ((size_t)stages * (narr * MFIT_TILE * 8 + 8) + MFIT_QN * 4);
}
