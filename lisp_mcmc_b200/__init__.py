"""lisp_mcmc_b200 -- B200 (sm_100a) backend for the Metropolis-Hastings hot path of afranson/Lisp-MCMC.

The product is csrc/ (hand-written CUDA + the C ABI of include/mfit_cuda.h, built to libmfit_cuda.so);
walker.py mirrors the reference's walker API on top of it, lisp/mfit-cuda.lisp is the SBCL/CFFI host."""
from . import _lib, distributed, families, io  # noqa: F401
from .walker import (Walker, builtin, create_log_liklihood_function, mcmc_fit, shard_chains, shard_datasets, walker_adaptive_steps,  # noqa: F401
                     walker_adaptive_steps_full, walker_create, walker_force_take_step, walker_get, walker_load,
                     walker_many_steps, walker_sample_region, walker_save, walker_set_create, walker_set_delete,
                     walker_set_get, walker_set_load, walker_set_save, linspace)
