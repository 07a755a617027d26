// eval_perchain.cu -- nvcc-compiled instantiations of the per-walker-data likelihood kernel (mfit_eval_perchain_body)
// for the built-in model families x likelihoods.
#include "mfit_eval.cuh"
#include "mfit_models.cuh"
#include "mfit_internal.h"

template <class Model, int LL, bool STAGED>
__global__ void __launch_bounds__(MFIT_PC_THREADS, STAGED ? MFIT_PC_STAGED_CTAS : 2) mfit_eval_perchain_kernel(const __grid_constant__ MfitEvalArgs a) {
  mfit_eval_perchain_body_t<Model, LL, STAGED>(a);
}

// [likelihood kind + 3 * staged]
#define MFIT_ROW(M)                                                               \
  {(const void*)mfit_eval_perchain_kernel<M, MFITK_NORMAL, false>,               \
   (const void*)mfit_eval_perchain_kernel<M, MFITK_CUTOFF, false>,               \
   (const void*)mfit_eval_perchain_kernel<M, MFITK_POISSON, false>,              \
   (const void*)mfit_eval_perchain_kernel<M, MFITK_NORMAL, true>,                \
   (const void*)mfit_eval_perchain_kernel<M, MFITK_CUTOFF, true>,                \
   (const void*)mfit_eval_perchain_kernel<M, MFITK_POISSON, true>}

static const void* const k_poly[][6] = {MFIT_ROW(MfitPoly<0>), MFIT_ROW(MfitPoly<1>), MFIT_ROW(MfitPoly<2>),
                                        MFIT_ROW(MfitPoly<3>), MFIT_ROW(MfitPoly<4>), MFIT_ROW(MfitPoly<5>),
                                        MFIT_ROW(MfitPoly<6>), MFIT_ROW(MfitPoly<7>)};
static const void* const k_expo[6] = MFIT_ROW(MfitExpo);
static const void* const k_sinus[6] = MFIT_ROW(MfitSinus);
static const void* const k_lorentz[6] = MFIT_ROW(MfitLorentz);
static const void* const k_gauss[6] = MFIT_ROW(MfitGaussLin);
static const void* const k_nv[][6] = {MFIT_ROW(MfitNvMulti<1>), MFIT_ROW(MfitNvMulti<2>), MFIT_ROW(MfitNvMulti<3>),
                                      MFIT_ROW(MfitNvMulti<4>)};

// staged: 0 = rows read straight from global memory, 1 = shared x / sigma staged in shared memory, 2 / 3 = staged with the
// register form for rows of up to 256 / 512 points (normal and poisson likelihoods only)
const void* mfit_builtin_perchain_kernel(int family, int degree, int ll, int staged) {
  if (ll < 0 || ll > 2) return nullptr;
  if (staged >= 2)
    return ll == MFITK_NORMAL    ? mfit_builtin_perchain_reg_kernel_normal(family, degree, staged == 2 ? 8 : 16)
           : ll == MFITK_POISSON ? mfit_builtin_perchain_reg_kernel_poisson(family, degree, staged == 2 ? 8 : 16)
                                 : nullptr;
  ll += staged ? 3 : 0;
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
