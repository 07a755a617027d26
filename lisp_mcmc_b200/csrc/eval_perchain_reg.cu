// eval_perchain_reg.cu -- the register form of the staged per-walker-data kernel (mfit_reg_row in mfit_eval.cuh) for the
// built-in model families: rows of up to 32 NPT points, normal and poisson likelihoods.  Compiled twice, once per
// likelihood (-DMFIT_PCR_LL=0 / 2: build/eval_perchain_reg_normal.o, _poisson.o), so that the build runs them side by side.
#include "mfit_eval.cuh"
#include "mfit_models.cuh"
#include "mfit_internal.h"

// staged, with the chain-invariant x and 1/sigma of a lane's points in registers (mfit_reg_row): rows of up to 32 NPT points
template <class Model, int LL, int NPT>
__global__ void __launch_bounds__(MFIT_PCR_THREADS, MFIT_PCR_CTAS) mfit_eval_perchain_reg_kernel(const __grid_constant__ MfitEvalArgs a) {
  mfit_eval_perchain_body_t<Model, LL, true, NPT>(a);
}

#ifndef MFIT_PCR_LL
#define MFIT_PCR_LL MFITK_NORMAL
#endif
// [NPT 8, NPT 16]
#define MFIT_ROW(M)                                                               \
  {(const void*)mfit_eval_perchain_reg_kernel<M, MFIT_PCR_LL, 8>,                \
   (const void*)mfit_eval_perchain_reg_kernel<M, MFIT_PCR_LL, 16>}

static const void* const k_poly[][2] = {MFIT_ROW(MfitPoly<0>), MFIT_ROW(MfitPoly<1>), MFIT_ROW(MfitPoly<2>),
                                        MFIT_ROW(MfitPoly<3>), MFIT_ROW(MfitPoly<4>), MFIT_ROW(MfitPoly<5>),
                                        MFIT_ROW(MfitPoly<6>), MFIT_ROW(MfitPoly<7>)};
static const void* const k_expo[2] = MFIT_ROW(MfitExpo);
static const void* const k_sinus[2] = MFIT_ROW(MfitSinus);
static const void* const k_lorentz[2] = MFIT_ROW(MfitLorentz);
static const void* const k_gauss[2] = MFIT_ROW(MfitGaussLin);
static const void* const k_nv[][2] = {MFIT_ROW(MfitNvMulti<1>), MFIT_ROW(MfitNvMulti<2>), MFIT_ROW(MfitNvMulti<3>),
                                      MFIT_ROW(MfitNvMulti<4>)};

#if MFIT_PCR_LL == MFITK_NORMAL
const void* mfit_builtin_perchain_reg_kernel_normal(int family, int degree, int npt) {
#else
const void* mfit_builtin_perchain_reg_kernel_poisson(int family, int degree, int npt) {
#endif
  if (npt != 8 && npt != 16) return nullptr;
  const int k = npt == 16 ? 1 : 0;
  switch (family) {
    case MFIT_FAM_POLY: return (degree >= 0 && degree <= 7) ? k_poly[degree][k] : nullptr;
    case MFIT_FAM_EXPO: return k_expo[k];
    case MFIT_FAM_SINUS: return k_sinus[k];
    case MFIT_FAM_LORENTZ: return k_lorentz[k];
    case MFIT_FAM_GAUSS_LIN: return k_gauss[k];
    case MFIT_FAM_NV_MULTI: return (degree >= 1 && degree <= 4) ? k_nv[degree - 1][k] : nullptr;
  }
  return nullptr;
}
