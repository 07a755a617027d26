// eval_fit.cu -- nvcc-compiled instantiations of the fit-curve / envelope kernel (mfit_fit_envelope_body) for the
// built-in model families; user s-expression models get theirs through NVRTC (nvrtc_glue.cpp).
#include "mfit_eval.cuh"
#include "mfit_models.cuh"
#include "mfit_internal.h"

template <class Model>
__global__ void __launch_bounds__(MFIT_FIT_THREADS) mfit_fit_envelope_kernel(const __grid_constant__ MfitFitArgs f) {
  mfit_fit_envelope_body<Model>(f);
}

#define MFIT_FITK(M) (const void*)mfit_fit_envelope_kernel<M>

static const void* const k_poly[] = {MFIT_FITK(MfitPoly<0>), MFIT_FITK(MfitPoly<1>), MFIT_FITK(MfitPoly<2>),
                                     MFIT_FITK(MfitPoly<3>), MFIT_FITK(MfitPoly<4>), MFIT_FITK(MfitPoly<5>),
                                     MFIT_FITK(MfitPoly<6>), MFIT_FITK(MfitPoly<7>)};
static const void* const k_nv[] = {MFIT_FITK(MfitNvMulti<1>), MFIT_FITK(MfitNvMulti<2>), MFIT_FITK(MfitNvMulti<3>),
                                   MFIT_FITK(MfitNvMulti<4>)};

const void* mfit_builtin_fit_kernel(int family, int degree) {
  switch (family) {
    case MFIT_FAM_POLY: return (degree >= 0 && degree <= 7) ? k_poly[degree] : nullptr;
    case MFIT_FAM_EXPO: return MFIT_FITK(MfitExpo);
    case MFIT_FAM_SINUS: return MFIT_FITK(MfitSinus);
    case MFIT_FAM_LORENTZ: return MFIT_FITK(MfitLorentz);
    case MFIT_FAM_GAUSS_LIN: return MFIT_FITK(MfitGaussLin);
    case MFIT_FAM_NV_MULTI: return (degree >= 1 && degree <= 4) ? k_nv[degree - 1] : nullptr;
  }
  return nullptr;
}
