"""ctypes binding of libbgp.so (include/bgp.h).  There is no CPU fallback: if the shared
library is missing, or no B200 is visible when a handle is created, this raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbgp.so")

BGP_MAX_OPS, BGP_MAX_LEAVES, BGP_MAX_DIM, BGP_MAX_THETA = 24, 4, 64, 80
(OP_CONST, OP_WHITE, OP_RBF, OP_MATERN12, OP_MATERN32, OP_MATERN52, OP_ADD, OP_MUL, OP_POW) = range(1, 10)
FLAG_ZEROABLE_WHITE = 1
PRIOR_NONE, PRIOR_HALFNORMAL_SQRT, PRIOR_ROUNDFLAT, PRIOR_INVGAMMA, PRIOR_NORMAL = range(5)
EXTRACT_L, EXTRACT_LINV, EXTRACT_KINV, EXTRACT_ALPHA = 1, 2, 3, 4
ACQ_EI, ACQ_TTEI, ACQ_MEAN, ACQ_LCB, ACQ_MES = 1, 2, 3, 4, 5


class Op(C.Structure):
    _fields_ = [("code", C.c_int32), ("theta_idx", C.c_int32), ("n_ls", C.c_int32),
                ("flags", C.c_int32), ("value", C.c_double), ("fixed_ls_offset", C.c_int32),
                ("reserved", C.c_int32)]


class Prior(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("p", C.c_double * 6)]


class BgpError(RuntimeError):
    pass


_P = C.c_void_p
_SIGNATURES = {
    "bgp_create": [C.POINTER(_P), C.c_int],
    "bgp_destroy": [_P],
    "bgp_version": [],
    "bgp_set_kernel": [_P, C.POINTER(Op), C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int],
    "bgp_set_warp": [_P, C.c_int],
    "bgp_mcmc_seed_source": [_P, _P],
    "bgp_set_priors": [_P, C.POINTER(Prior), C.c_int],
    "bgp_set_data": [_P, _P, _P, _P, C.c_int, C.c_int, _P],
    "bgp_logprob_batched": [_P, _P, C.c_int, _P, _P, _P, _P, _P],
    "bgp_factor_slab_doubles": [_P],
    "bgp_factorize_batched": [_P, _P, C.c_int, _P, _P, _P, _P, _P],
    "bgp_factor_extract": [_P, _P, _P, C.c_int, _P, _P],
    "bgp_lml_gradient": [_P, _P, _P, _P, _P, _P],
    "bgp_predict_batched": [_P, _P, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_double, C.c_double,
                            _P, _P, _P, C.c_int, _P, _P, C.c_int64, _P],
    "bgp_acq_sweep": [_P, C.c_int, _P, _P, C.c_int, C.c_int, C.c_double, _P, C.c_int, _P, _P, _P, _P, _P],
    "bgp_argmax": [_P, _P, C.c_int, _P, _P],
    "bgp_acq_stats": [_P, _P, _P, C.c_int, C.c_int, _P, _P],
    "bgp_mes_fit": [_P, _P, _P, C.c_int, C.c_int, _P, _P],
    "bgp_ei_best": [_P, _P, _P, C.c_int, C.c_int, C.c_double, _P, C.c_int64, _P, _P],
    "bgp_acq_per_theta": [_P, C.c_int, _P, _P, C.c_int, C.c_int, C.c_double, _P, _P, _P, C.c_int, _P, _P, _P, _P],
    "bgp_acq_combine": [_P, _P, C.c_int, C.c_int, _P, _P, _P],
    "bgp_mcmc_run": [_P, _P, _P, C.c_int, C.c_int, C.c_double, C.c_uint64, _P, _P, _P, _P],
    "bgp_mcmc_split": [_P, C.c_int, C.c_uint64, C.c_int, _P, _P],
    "bgp_mcmc_propose": [_P, _P, _P, C.c_int, C.c_int, C.c_double, C.c_uint64, C.c_int, _P, _P, _P, _P],
    "bgp_mcmc_accept": [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_uint64, C.c_int, _P, _P, _P, _P],
    "bgp_posterior_cov": [_P, _P, _P, _P, C.c_int, C.c_int64, C.c_int, C.c_double, _P, C.c_int64, _P],
    "bgp_dense_cholesky": [_P, _P, C.c_int, C.c_int64, C.c_double, _P, _P, _P],
    "bgp_dense_slab_doubles": [C.c_int],
    "bgp_pvrs_combine": [_P, _P, _P, C.c_int, _P, C.c_int, _P, _P, _P, _P, _P],
    "bgp_vr_combine": [_P, _P, C.c_int, C.c_int64, _P, _P, _P, _P, _P],
    "bgp_slab_trmm": [_P, _P, C.c_int, _P, C.c_int, _P, _P, _P],
    "bgp_acq_sweep_nccl": [_P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P, C.c_int, C.c_double, _P, C.c_int,
                           C.c_double, C.c_double, _P, _P, _P],
    "bgp_dense_cholesky_inplace": [_P, _P, C.c_int, C.c_int64, C.c_double, _P, _P],
    "bgp_dense_trmm": [_P, _P, C.c_int, C.c_int64, _P, C.c_int, _P, _P, _P],
    "bgp_peer_export": [_P, C.c_int, _P],
    "bgp_peer_connect": [_P, _P, C.c_int, C.c_int],
    "bgp_peer_close": [_P],
    "bgp_peer_status": [_P, C.POINTER(C.c_int)],
    "bgp_peer_counters": [_P, C.POINTER(C.c_uint64)],
    "bgp_logprob_launches": [_P],
    "bgp_mcmc_run_sharded": [_P, _P, _P, C.c_int, C.c_int, C.c_double, C.c_uint64, _P, _P, _P, _P],
}
IPC_HANDLE_BYTES = 64
EXPORTED = tuple(_SIGNATURES) + ("bgp_last_error",)

_lib = None


def load():
    """Returns the loaded library; raises BgpError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BgpError(f"{LIB_PATH} is missing: build it with `sh bayes-skopt_b200/build.sh` "
                       "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int64 if name.endswith("slab_doubles") else C.c_int
    lib.bgp_last_error.argtypes = []
    lib.bgp_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().bgp_last_error().decode("utf-8", "replace")
        raise BgpError(f"libbgp {what}: {msg}")
