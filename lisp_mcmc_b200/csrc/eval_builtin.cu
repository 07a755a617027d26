// eval_builtin.cu -- nvcc-compiled instantiations of the log-likelihood kernel for the built-in
// model families x likelihoods (the user s-expression path instantiates the same template via NVRTC).
#include "mfit_eval.cuh"
#include "mfit_models.cuh"
#include "mfit_internal.h"

template <class Model, int LL>
__global__ void __launch_bounds__(MFIT_THREADS, MFIT_MIN_CTAS) mfit_eval_kernel(const __grid_constant__ MfitMultiArgs a) {
  mfit_eval_body<Model, LL>(a);
}

#define MFIT_ROW(M)                                                         \
  {(const void*)mfit_eval_kernel<M, MFITK_NORMAL>,                        \
   (const void*)mfit_eval_kernel<M, MFITK_CUTOFF>,                 \
   (const void*)mfit_eval_kernel<M, MFITK_POISSON>}

static const void* const k_poly[][3] = {MFIT_ROW(MfitPoly<0>), MFIT_ROW(MfitPoly<1>), MFIT_ROW(MfitPoly<2>),
                                        MFIT_ROW(MfitPoly<3>), MFIT_ROW(MfitPoly<4>), MFIT_ROW(MfitPoly<5>),
                                        MFIT_ROW(MfitPoly<6>), MFIT_ROW(MfitPoly<7>)};
static const void* const k_expo[3] = MFIT_ROW(MfitExpo);
static const void* const k_sinus[3] = MFIT_ROW(MfitSinus);
static const void* const k_lorentz[3] = MFIT_ROW(MfitLorentz);
static const void* const k_gauss[3] = MFIT_ROW(MfitGaussLin);
static const void* const k_nv[][3] = {MFIT_ROW(MfitNvMulti<1>), MFIT_ROW(MfitNvMulti<2>), MFIT_ROW(MfitNvMulti<3>),
                                      MFIT_ROW(MfitNvMulti<4>)};

const void* mfit_builtin_eval_kernel(int family, int degree, int ll) {
  if (ll < 0 || ll > 2) return nullptr;
  switch (family) {
    case MFIT_FAM_POLY: return (degree >= 0 && degree <= 7) ? k_poly[degree][ll] : nullptr;
    case MFIT_FAM_EXPO: return k_expo[ll];
    case MFIT_FAM_SINUS: return k_sinus[ll];
    case MFIT_FAM_LORENTZ: return k_lorentz[ll];
    case MFIT_FAM_GAUSS_LIN: return k_gauss[ll];
    case MFIT_FAM_NV_MULTI: return (degree >= 1 && degree <= 4) ? k_nv[degree - 1][ll] : nullptr;
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
