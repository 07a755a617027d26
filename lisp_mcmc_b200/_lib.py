"""ctypes binding of libmfit_cuda.so (include/mfit_cuda.h) -- the same entry points the CFFI file
lisp/mfit-cuda.lisp binds from SBCL.  There is no CPU fallback: if the library is missing, or it
finds no sm_100 device, calls raise."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MFIT_LIB") or os.path.join(_HERE, "libmfit_cuda.so")   # MFIT_LIB: an alternative build (tools/build_variant.sh)

OK, ERR_ARG, ERR_CUDA, ERR_SEXPR, ERR_NVRTC, ERR_STATE, ERR_STREAM, ERR_NCCL, ERR_NOMEM = 0, -1, -2, -3, -4, -5, -6, -7, -8
FAM_POLY, FAM_EXPO, FAM_SINUS, FAM_LORENTZ, FAM_GAUSS_LIN, FAM_NV_MULTI, FAM_SEXPR = 0, 1, 2, 3, 4, 5, 100
LL_NORMAL, LL_NORMAL_CUTOFF, LL_POISSON, LL_USER = 0, 1, 2, 3
AUTO_NONE, AUTO_PROB_SETTLE = 0, 1
OPT_COVARIANCE, OPT_BEST_VALUE = 0, 1
FIT_MOST_LIKELY, FIT_MEDIAN = 0, 1
SHARD_CHAINS, SHARD_DATASETS = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_cpp = C.POINTER(C.c_char_p)


class AdaptCfg(C.Structure):
    """mfit_adapt_cfg: the keyword arguments of walker-adaptive-steps-full (mcmc-fitting.lisp:862)."""
    _fields_ = [("n", C.c_int64), ("temperature", C.c_double), ("auto_mode", C.c_int32),
                ("sampling_optimization", C.c_int32), ("max_walker_length", C.c_int64), ("greedy", C.c_int32),
                ("reserved", C.c_int32)]


class ChainStats(C.Structure):
    """mfit_chain_stats: scalar slots of struct walker (mcmc-fitting.lisp:467-479) per chain."""
    _fields_ = [("length", C.c_int64), ("age", C.c_int64), ("steps_this_run", C.c_int64),
                ("last_prob", C.c_double), ("best_prob", C.c_double), ("temperature", C.c_double),
                ("accept_200", C.c_double), ("shutting_down", C.c_int32), ("l_updates", C.c_int32)]


# every symbol include/mfit_cuda.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "mfit_abi_version": (C.c_int, []),
    "mfit_ctx_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "mfit_ctx_destroy": (None, [C.c_void_p]),
    "mfit_last_error": (C.c_char_p, [C.c_void_p]),
    "mfit_synchronize": (C.c_int, [C.c_void_p]),
    "mfit_launch_count": (C.c_int64, [C.c_void_p]),
    "mfit_problem_define": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mfit_dataset_upload": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, C.c_int64]),
    "mfit_dataset_upload_xd": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_int, _dp, _dp, C.c_int64]),
    "mfit_dataset_upload_per_walker": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, C.c_int64, C.c_int64, C.c_int, C.c_int]),
    "mfit_model_builtin": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _ip, C.c_int]),
    "mfit_model_sexpr": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, _cpp, C.c_int]),
    "mfit_likelihood_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mfit_likelihood_sexpr": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p]),
    "mfit_prior_flat": (C.c_int, [C.c_void_p, C.c_int]),
    "mfit_prior_bounds": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _ip, _dp, _dp]),
    "mfit_prior_sexpr": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, _cpp, C.c_int]),
    "mfit_lower_sexpr": (C.c_int, [C.c_char_p, _cpp, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int]),
    "mfit_compile_sexpr": (C.c_int, [C.c_char_p, _cpp, C.c_int, C.c_int, C.c_char_p, C.c_int]),
    "mfit_set_segment_tiles": (C.c_int, [C.c_void_p, C.c_int]),
    "mfit_segment_geometry": (C.c_int, [C.c_void_p, C.c_int, _ip, _ip]),
    "mfit_eval_logp": (C.c_int, [C.c_void_p, _dp, C.c_int64, _dp, _dp, _dp]),
    "mfit_chains_init": (C.c_int, [C.c_void_p, _dp, C.c_int64, C.c_uint64, C.c_int64, C.c_int64]),
    "mfit_run_adaptive": (C.c_int, [C.c_void_p, C.POINTER(AdaptCfg), _dp, _dp, _dp, C.c_int64, _lp]),
    "mfit_run_begin": (C.c_int, [C.c_void_p, C.POINTER(AdaptCfg), _dp, _dp, _dp, C.c_int64]),
    "mfit_run_advance": (C.c_int, [C.c_void_p, C.c_int64, _lp, C.POINTER(C.c_int)]),
    "mfit_run_many_steps": (C.c_int, [C.c_void_p, C.c_int64, _dp, _dp, _dp, C.c_int64]),
    "mfit_run_many_steps_per_chain": (C.c_int, [C.c_void_p, C.c_int64, _dp, _dp, _dp, C.c_int64]),
    "mfit_run_sample_region": (C.c_int, [C.c_void_p, C.c_double, _dp, C.c_int64]),
    "mfit_force_take_step": (C.c_int, [C.c_void_p]),
    "mfit_request_stop": (C.c_int, [C.c_void_p]),
    "mfit_trace_read": (C.c_int, [C.c_void_p, C.c_int64, _dp, _dp, _lp]),
    "mfit_stats_read": (C.c_int, [C.c_void_p, C.POINTER(ChainStats)]),
    "mfit_best_read": (C.c_int, [C.c_void_p, _dp, _dp]),
    "mfit_last_read": (C.c_int, [C.c_void_p, _dp, _dp]),
    "mfit_l_matrix_read": (C.c_int, [C.c_void_p, _dp]),
    "mfit_trace_analytics": (C.c_int, [C.c_void_p, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp]),
    "mfit_fit_envelope": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp, _lp]),
    "mfit_chains_restore": (C.c_int, [C.c_void_p, C.c_int64, _dp, _dp, _lp, _lp, _dp, _dp]),
    "mfit_keep_walks": (C.c_int, [C.c_void_p, C.c_int64]),
    "mfit_walker_modify": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "mfit_philox_stream": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, _dp, _dp]),
    "mfit_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "mfit_nccl_attach": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "mfit_peer_setup": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p]),
    "mfit_peer_attach": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mfit_dataset_owners": (C.c_int, [C.c_void_p, _ip]),
    "mfit_nccl_allgather_f64": (C.c_int, [C.c_void_p, _dp, C.c_int64, _dp]),
    "mfit_multi_create": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "mfit_multi_destroy": (None, [C.c_void_p]),
    "mfit_multi_last_error": (C.c_char_p, [C.c_void_p]),
    "mfit_multi_size": (C.c_int, [C.c_void_p]),
    "mfit_multi_ctx": (C.c_void_p, [C.c_void_p, C.c_int]),
    "mfit_multi_launch_count": (C.c_int64, [C.c_void_p]),
    "mfit_multi_reserve": (C.c_int, [C.c_void_p, C.c_int64]),
    "mfit_multi_problem_define": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mfit_multi_dataset_upload": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_int, _dp, _dp, C.c_int64]),
    "mfit_multi_model_builtin": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _ip, C.c_int]),
    "mfit_multi_model_sexpr": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, _cpp, C.c_int]),
    "mfit_multi_likelihood_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mfit_multi_likelihood_sexpr": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p]),
    "mfit_multi_prior_flat": (C.c_int, [C.c_void_p, C.c_int]),
    "mfit_multi_prior_bounds": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _ip, _dp, _dp]),
    "mfit_multi_prior_sexpr": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, _cpp, C.c_int]),
    "mfit_multi_eval_logp": (C.c_int, [C.c_void_p, _dp, C.c_int64, _dp, _dp, _dp]),
    "mfit_multi_chains_init": (C.c_int, [C.c_void_p, _dp, C.c_int64, C.c_uint64, C.c_int64]),
    "mfit_multi_run_adaptive": (C.c_int, [C.c_void_p, C.POINTER(AdaptCfg), _dp, _lp]),
    "mfit_multi_run_begin": (C.c_int, [C.c_void_p, C.POINTER(AdaptCfg), _dp]),
    "mfit_multi_run_advance": (C.c_int, [C.c_void_p, C.c_int64, _lp, C.POINTER(C.c_int)]),
    "mfit_multi_run_many_steps": (C.c_int, [C.c_void_p, C.c_int64, _dp]),
    "mfit_multi_request_stop": (C.c_int, [C.c_void_p]),
    "mfit_multi_trace_read": (C.c_int, [C.c_void_p, C.c_int64, _dp, _dp, _lp]),
    "mfit_multi_stats_read": (C.c_int, [C.c_void_p, C.POINTER(ChainStats)]),
    "mfit_multi_best_read": (C.c_int, [C.c_void_p, _dp, _dp]),
    "mfit_multi_last_read": (C.c_int, [C.c_void_p, _dp, _dp]),
    "mfit_multi_l_matrix_read": (C.c_int, [C.c_void_p, _dp]),
    "mfit_multi_trace_analytics": (C.c_int, [C.c_void_p, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp]),
    "mfit_multi_fit_envelope": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp, _lp]),
    "mfit_multi_walker_modify": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "mfit_multi_keep_walks": (C.c_int, [C.c_void_p, C.c_int64]),
    "mfit_multi_chains_restore": (C.c_int, [C.c_void_p, C.c_int64, _dp, _dp, _lp, _lp, _dp, _dp]),
    "mfit_measure_fp64": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, _dp]),
    "mfit_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "mfit_eval_kernel_time": (C.c_int, [C.c_void_p, _dp, _lp]),
    "mfit_eval_mode_tiles": (C.c_int, [C.c_void_p, _lp]),
    "mfit_set_small_path": (C.c_int, [C.c_void_p, C.c_int]),
    "mfit_set_groups": (C.c_int, [C.c_void_p, C.c_int]),
    "mfit_flush_l2": (C.c_int, [C.c_void_p]),
    "mfit_stream_handle": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
}

_LIB = None


class MfitError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "mfit error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Load libmfit_cuda.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


def keys_array(keys):
    return (C.c_char_p * len(keys))(*[k.encode() for k in keys])


def lower_sexpr(text, keys, kind=0):
    """The CUDA C a user lambda lowers to (kind 0 = :function, 1 = :log-prior).  No GPU needed."""
    L = lib()
    err = C.create_string_buffer(1024)
    need = L.mfit_lower_sexpr(text.encode(), keys_array(keys), len(keys), kind, None, 0, err, 1024)
    if need < 0:
        raise MfitError(need, err.value.decode())
    out = C.create_string_buffer(need)
    L.mfit_lower_sexpr(text.encode(), keys_array(keys), len(keys), kind, out, need, err, 1024)
    return out.value.decode()


def compile_sexpr(text, keys, kind=0):
    """Lower and NVRTC-compile for sm_100a; returns the cubin size.  No GPU needed."""
    err = C.create_string_buffer(16384)
    n = lib().mfit_compile_sexpr(text.encode(), keys_array(keys), len(keys), kind, err, 16384)
    if n < 0:
        raise MfitError(n, err.value.decode())
    return n
