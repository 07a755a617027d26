// eval_builtin_ll.cu -- nvcc-compiled instantiations of the log-likelihood kernel for the built-in model families, for ONE
// likelihood kind per object (-DMFIT_EB_LL=0 normal / 1 cutoff / 2 poisson -> build/eval_builtin_normal.o, _cutoff.o,
// _poisson.o): the three compile side by side instead of one after the other.
#include "mfit_eval.cuh"
#include "mfit_models.cuh"
#include "mfit_internal.h"

#ifndef MFIT_EB_LL
#define MFIT_EB_LL MFITK_NORMAL
#endif

template <class Model, int LL>
__global__ void __launch_bounds__(MFIT_THREADS, MFIT_MIN_CTAS) mfit_eval_kernel(const __grid_constant__ MfitMultiArgs a) {
  mfit_eval_body<Model, LL>(a);
}

#define MFIT_K(M) (const void*)mfit_eval_kernel<M, MFIT_EB_LL>

static const void* const k_poly[] = {MFIT_K(MfitPoly<0>), MFIT_K(MfitPoly<1>), MFIT_K(MfitPoly<2>), MFIT_K(MfitPoly<3>),
                                     MFIT_K(MfitPoly<4>), MFIT_K(MfitPoly<5>), MFIT_K(MfitPoly<6>), MFIT_K(MfitPoly<7>)};
static const void* const k_nv[] = {MFIT_K(MfitNvMulti<1>), MFIT_K(MfitNvMulti<2>), MFIT_K(MfitNvMulti<3>), MFIT_K(MfitNvMulti<4>)};

#if MFIT_EB_LL == MFITK_NORMAL
const void* mfit_builtin_eval_kernel_normal(int family, int degree) {
#elif MFIT_EB_LL == MFITK_CUTOFF
const void* mfit_builtin_eval_kernel_cutoff(int family, int degree) {
#else
const void* mfit_builtin_eval_kernel_poisson(int family, int degree) {
#endif
  switch (family) {
    case MFIT_FAM_POLY: return (degree >= 0 && degree <= 7) ? k_poly[degree] : nullptr;
    case MFIT_FAM_EXPO: return MFIT_K(MfitExpo);
    case MFIT_FAM_SINUS: return MFIT_K(MfitSinus);
    case MFIT_FAM_LORENTZ: return MFIT_K(MfitLorentz);
    case MFIT_FAM_GAUSS_LIN: return MFIT_K(MfitGaussLin);
    case MFIT_FAM_NV_MULTI: return (degree >= 1 && degree <= 4) ? k_nv[degree - 1] : nullptr;
  }
  return nullptr;
}
