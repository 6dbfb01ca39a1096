"""ctypes binding of libnfn_b200.so (the C ABI declared in include/nfn_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, this module
raises.  PyTorch is only the tensor carrier (device memory + current stream).
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# NFN_B200_LIB points at an A/B tuning variant built by build.py (tools only)
LIB_PATH = os.environ.get("NFN_B200_LIB") or os.path.join(_PKG, "libnfn_b200.so")

NFN_MAX_FLOWS = 64
NFN_MAX_DIMS = 8
FLOW_CODES = {"planar": 0, "radial": 1, "affine": 2}

_c_float_p = ctypes.c_void_p  # raw device/host addresses
_i64 = ctypes.c_int64


class ChainDesc(ctypes.Structure):
    """struct nfn_chain_desc (include/nfn_b200.h)."""

    _fields_ = [
        ("n_dims", ctypes.c_int32),
        ("n_flows", ctypes.c_int32),
        ("trainable_base", ctypes.c_int32),
        ("flow_type", ctypes.c_uint8 * NFN_MAX_FLOWS),
    ]


class VariationalLayer(ctypes.Structure):
    """nfn_variational_layer (include/nfn_b200.h): one DenseVariational layer's operands for nfn_bayes_train_step."""
    _fields_ = [("posterior", ctypes.c_void_p), ("prior_loc", ctypes.c_void_p), ("eps", ctypes.c_void_p),
                ("w", ctypes.c_void_p), ("dw", ctypes.c_void_p), ("dposterior", ctypes.c_void_p),
                ("dprior_loc", ctypes.c_void_p), ("kl", ctypes.c_void_p), ("prior_scale", ctypes.c_float),
                ("kl_grad", ctypes.c_float), ("n", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class EventXform(ctypes.Structure):
    """struct nfn_event_xform (include/nfn_b200.h): y normalisation / noise / Jacobian shift / exp fused into a head."""

    _fields_ = [
        ("mean", ctypes.c_float * NFN_MAX_DIMS),
        ("std", ctypes.c_float * NFN_MAX_DIMS),
        ("noise_std", ctypes.c_float),
        ("logp_shift", ctypes.c_float),
        ("seed", ctypes.c_uint64),
        ("offset", ctypes.c_uint64),
        ("offset_dev", ctypes.c_void_p),
        ("flags", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


XF_NORMALISE, XF_NOISE, XF_EXP = 1, 2, 4
_xf_p = ctypes.POINTER(EventXform)


class NfnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libnfn_b200 error %d: %s" % (code, msg))
        self.code = code


# name -> (restype, argtypes); every symbol include/nfn_b200.h declares
SIGNATURES = {
    "nfn_version": (ctypes.c_int, []),
    "nfn_last_error": (ctypes.c_char_p, []),
    "nfn_set_math_mode": (ctypes.c_int, [ctypes.c_int]),
    "nfn_set_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    "nfn_get_option": (ctypes.c_int, [ctypes.c_char_p]),
    "nfn_launch_count_reset": (_i64, []),
    "nfn_host_release": (ctypes.c_int, []),
    "nfn_chain_param_size": (ctypes.c_int, [ctypes.POINTER(ChainDesc)]),
    "nfn_chain_is_specialized": (ctypes.c_int, [ctypes.POINTER(ChainDesc)]),
    "nfn_jit_cache_size": (ctypes.c_int, []),
    "nfn_jit_compile_check": (_i64, [ctypes.POINTER(ChainDesc), ctypes.c_int]),
    "nfn_chain_forward": (ctypes.c_int, [ctypes.POINTER(ChainDesc), _c_float_p, _c_float_p, _i64,
                                        _c_float_p, _i64, ctypes.c_void_p]),
    "nfn_chain_forward_grid": (ctypes.c_int, [ctypes.POINTER(ChainDesc), _c_float_p, _c_float_p, _i64,
                                             _c_float_p, _i64, ctypes.c_void_p]),
    "nfn_chain_forward_backward": (ctypes.c_int, [ctypes.POINTER(ChainDesc), _c_float_p, _c_float_p, _i64,
                                                 _c_float_p, ctypes.c_float, _c_float_p, _c_float_p,
                                                 _c_float_p, ctypes.c_void_p, _c_float_p, _i64,
                                                 ctypes.c_void_p]),
    "nfn_peer_region_bytes": (_i64, [ctypes.c_int, ctypes.c_int]),
    "nfn_peer_alloc": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p]),
    "nfn_peer_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "nfn_peer_close": (ctypes.c_int, [ctypes.c_void_p]),
    "nfn_peer_free": (ctypes.c_int, [ctypes.c_void_p]),
    "nfn_peer_comm_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p)]),
    "nfn_peer_comm_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "nfn_peer_allreduce": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, _c_float_p, ctypes.c_void_p]),
    "nfn_peer_set_deferred": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "nfn_peer_flush": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "nfn_peer_status": (ctypes.c_int, [ctypes.c_void_p]),
    "nfn_chain_forward_backward_peer": (ctypes.c_int, [ctypes.POINTER(ChainDesc), _c_float_p, _c_float_p, _i64,
                                                      _c_float_p, ctypes.c_float, _c_float_p, _c_float_p,
                                                      _c_float_p, ctypes.c_int, ctypes.c_void_p, _c_float_p,
                                                      _i64, ctypes.c_void_p]),
    "nfn_dense_chain_forward": (ctypes.c_int, [ctypes.POINTER(ChainDesc), ctypes.c_int, _c_float_p, _c_float_p,
                                              _c_float_p, _c_float_p, _i64, _c_float_p, _i64, ctypes.c_void_p]),
    "nfn_dense_chain_forward_backward": (ctypes.c_int, [ctypes.POINTER(ChainDesc), ctypes.c_int, _c_float_p,
                                                       _c_float_p, _c_float_p, _c_float_p, _i64, _c_float_p,
                                                       ctypes.c_float, _c_float_p, _c_float_p, _c_float_p,
                                                       _c_float_p, ctypes.c_void_p, _i64, ctypes.c_void_p]),
    "nfn_jit_dense_compile_check": (_i64, [ctypes.POINTER(ChainDesc), ctypes.c_int, ctypes.c_int]),
    "nfn_jit_dense_tc5_compile_check": (_i64, [ctypes.POINTER(ChainDesc), ctypes.c_int, ctypes.c_int]),
    "nfn_dense_act_supported": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "nfn_dense_act_forward": (ctypes.c_int, [_c_float_p, _c_float_p, _c_float_p, _i64, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, _c_float_p, ctypes.c_void_p]),
    "nfn_dense_act_backward": (ctypes.c_int, [_c_float_p, _c_float_p, _c_float_p, _c_float_p, _i64, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p, _c_float_p,
                                             ctypes.c_void_p]),
    "nfn_flow_forward": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p, _i64,
                                       _c_float_p, _c_float_p, _i64, ctypes.c_void_p]),
    "nfn_mdn_forward": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p, _i64,
                                      _c_float_p, _i64, ctypes.c_void_p]),
    "nfn_mdn_forward_backward": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p, _i64,
                                               _c_float_p, ctypes.c_float, _c_float_p, _c_float_p,
                                               _c_float_p, ctypes.c_void_p, _c_float_p, _i64,
                                               ctypes.c_void_p]),
    "nfn_kmn_forward": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p, _i64,
                                      _c_float_p, _c_float_p, _c_float_p, _i64, ctypes.c_void_p]),
    "nfn_kmn_forward_backward": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p, _i64,
                                               _c_float_p, _c_float_p, _c_float_p, ctypes.c_float,
                                               _c_float_p, _c_float_p, _c_float_p, _c_float_p,
                                               ctypes.c_void_p, _i64, ctypes.c_void_p]),
    "nfn_chain_forward_x": (ctypes.c_int, [ctypes.POINTER(ChainDesc), _c_float_p, _c_float_p, _i64,
                                          _c_float_p, _i64, _xf_p, ctypes.c_void_p]),
    "nfn_chain_forward_grid_x": (ctypes.c_int, [ctypes.POINTER(ChainDesc), _c_float_p, _c_float_p, _i64,
                                               _c_float_p, _i64, _xf_p, ctypes.c_void_p]),
    "nfn_chain_forward_backward_x": (ctypes.c_int, [ctypes.POINTER(ChainDesc), _c_float_p, _c_float_p, _i64,
                                                   _c_float_p, ctypes.c_float, _c_float_p, _c_float_p,
                                                   _c_float_p, ctypes.c_void_p, _c_float_p, _i64, _xf_p,
                                                   ctypes.c_void_p]),
    "nfn_dense_chain_forward_x": (ctypes.c_int, [ctypes.POINTER(ChainDesc), ctypes.c_int, _c_float_p, _c_float_p,
                                                _c_float_p, _c_float_p, _i64, _c_float_p, _i64, _xf_p,
                                                ctypes.c_void_p]),
    "nfn_dense_chain_forward_backward_x": (ctypes.c_int, [ctypes.POINTER(ChainDesc), ctypes.c_int, _c_float_p,
                                                         _c_float_p, _c_float_p, _c_float_p, _i64, _c_float_p,
                                                         ctypes.c_float, _c_float_p, _c_float_p, _c_float_p,
                                                         _c_float_p, ctypes.c_void_p, _i64, _xf_p, ctypes.c_void_p]),
    "nfn_dense_act_forward_x": (ctypes.c_int, [_c_float_p, _c_float_p, _c_float_p, _c_float_p, _c_float_p, _i64,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_float_p, ctypes.c_void_p]),
    "nfn_dense_act_backward_x": (ctypes.c_int, [_c_float_p, _c_float_p, _c_float_p, _c_float_p, _c_float_p, _c_float_p,
                                               _i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p,
                                               _c_float_p, ctypes.c_void_p]),
    "nfn_mdn_forward_x": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p, _i64,
                                        _c_float_p, _i64, _xf_p, ctypes.c_void_p]),
    "nfn_mdn_forward_backward_x": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p, _i64,
                                                 _c_float_p, ctypes.c_float, _c_float_p, _c_float_p,
                                                 _c_float_p, ctypes.c_void_p, _c_float_p, _i64, _xf_p,
                                                 ctypes.c_void_p]),
    "nfn_bayes_train_step": (ctypes.c_int, [ctypes.POINTER(ChainDesc), ctypes.c_int, ctypes.c_int, _i64, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p, _c_float_p,
                                           _c_float_p, _i64, ctypes.POINTER(VariationalLayer),
                                           ctypes.POINTER(VariationalLayer), ctypes.c_float, _c_float_p, _c_float_p,
                                           _c_float_p, ctypes.c_void_p, _xf_p, ctypes.c_void_p]),
    "nfn_variational_sample": (ctypes.c_int, [_c_float_p, _c_float_p, ctypes.c_float, _c_float_p, ctypes.c_int,
                                             ctypes.c_int, _c_float_p, ctypes.c_void_p, ctypes.c_void_p]),
    "nfn_variational_sample_backward": (ctypes.c_int, [_c_float_p, _c_float_p, ctypes.c_float, _c_float_p, _c_float_p,
                                                      _c_float_p, ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p,
                                                      ctypes.c_void_p]),
    "nfn_dense_act_forward_draws": (ctypes.c_int, [_c_float_p, _c_float_p, _c_float_p, _c_float_p, ctypes.c_int, _i64,
                                                  ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_float_p,
                                                  ctypes.c_void_p]),
    "nfn_dense_act_backward_draws": (ctypes.c_int, [_c_float_p, _c_float_p, _c_float_p, _c_float_p, _c_float_p,
                                                   ctypes.c_int, _i64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                   ctypes.c_int, _c_float_p, ctypes.c_void_p]),
    "nfn_dense_chain_forward_draws_x": (ctypes.c_int, [ctypes.POINTER(ChainDesc), ctypes.c_int, ctypes.c_int, _i64,
                                                      _c_float_p, _c_float_p, _c_float_p, _c_float_p, _i64, _c_float_p,
                                                      _xf_p, ctypes.c_void_p]),
    "nfn_dense_chain_forward_backward_draws_x": (ctypes.c_int, [ctypes.POINTER(ChainDesc), ctypes.c_int, ctypes.c_int,
                                                               _i64, _c_float_p, _c_float_p, _c_float_p, _c_float_p,
                                                               _i64, _c_float_p, ctypes.c_float, _c_float_p, _c_float_p,
                                                               _c_float_p, _c_float_p, ctypes.c_void_p, _xf_p,
                                                               ctypes.c_void_p]),
    "nfn_dense_mdn_forward_x": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p,
                                              _c_float_p, _c_float_p, _i64, _c_float_p, _i64, _xf_p, ctypes.c_void_p]),
    "nfn_dense_mdn_forward_backward_x": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_float_p,
                                                       _c_float_p, _c_float_p, _c_float_p, _i64, _c_float_p,
                                                       ctypes.c_float, _c_float_p, _c_float_p, _c_float_p,
                                                       _c_float_p, ctypes.c_void_p, _i64, _xf_p, ctypes.c_void_p]),
    "nfn_dense_mdn_forward_draws_x": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i64,
                                                    _c_float_p, _c_float_p, _c_float_p, _c_float_p, _i64, _c_float_p,
                                                    _xf_p, ctypes.c_void_p]),
    "nfn_dense_mdn_forward_backward_draws_x": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i64,
                                                             _c_float_p, _c_float_p, _c_float_p, _c_float_p, _i64,
                                                             _c_float_p, ctypes.c_float, _c_float_p, _c_float_p,
                                                             _c_float_p, _c_float_p, ctypes.c_void_p, _xf_p,
                                                             ctypes.c_void_p]),
    "nfn_dense_kmn_forward_x": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p,
                                              _c_float_p, _c_float_p, _i64, _c_float_p, _c_float_p, _c_float_p, _i64,
                                              _xf_p, ctypes.c_void_p]),
    "nfn_dense_kmn_forward_backward_x": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_float_p,
                                                       _c_float_p, _c_float_p, _c_float_p, _i64, _c_float_p,
                                                       _c_float_p, _c_float_p, ctypes.c_float, _c_float_p,
                                                       _c_float_p, _c_float_p, _c_float_p, _c_float_p,
                                                       ctypes.c_void_p, _i64, _xf_p, ctypes.c_void_p]),
    "nfn_jit_dense_kmn_compile_check": (_i64, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "nfn_jit_dense_mdn_compile_check": (_i64, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "nfn_kmn_forward_x": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p, _i64,
                                        _c_float_p, _c_float_p, _c_float_p, _i64, _xf_p, ctypes.c_void_p]),
    "nfn_kmn_forward_backward_x": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p, _i64,
                                                 _c_float_p, _c_float_p, _c_float_p, ctypes.c_float,
                                                 _c_float_p, _c_float_p, _c_float_p, _c_float_p,
                                                 ctypes.c_void_p, _i64, _xf_p, ctypes.c_void_p]),
    "nfn_logmeanexp_draws": (ctypes.c_int, [_c_float_p, _i64, _i64, _c_float_p, ctypes.c_void_p]),
    "nfn_chain_forward_host": (ctypes.c_int, [ctypes.POINTER(ChainDesc), _c_float_p, _c_float_p, _i64,
                                             _c_float_p, _i64]),
    "nfn_chain_forward_backward_host": (ctypes.c_int, [ctypes.POINTER(ChainDesc), _c_float_p, _c_float_p,
                                                      _i64, _c_float_p, ctypes.c_float, _c_float_p,
                                                      _c_float_p, ctypes.POINTER(ctypes.c_double),
                                                      _c_float_p, _i64]),
    "nfn_mdn_forward_backward_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p,
                                                    _i64, _c_float_p, ctypes.c_float, _c_float_p,
                                                    _c_float_p, ctypes.POINTER(ctypes.c_double), _i64]),
}

_lib = None


def load():
    """Load libnfn_b200.so (built in-tree by ``normalizingflownetwork_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -m normalizingflownetwork_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc < 0:
        raise NfnError(rc, load().nfn_last_error().decode("utf-8", "replace"))
    return rc


def make_desc(flow_types, n_dims, trainable_base_dist):
    flow_types = list(flow_types)
    if len(flow_types) > NFN_MAX_FLOWS:
        raise ValueError("at most %d flows are supported, got %d" % (NFN_MAX_FLOWS, len(flow_types)))
    if not 1 <= int(n_dims) <= NFN_MAX_DIMS:
        raise ValueError("n_dims must be in 1..%d, got %r" % (NFN_MAX_DIMS, n_dims))
    d = ChainDesc()
    d.n_dims = int(n_dims)
    d.n_flows = len(flow_types)
    d.trainable_base = 1 if trainable_base_dist else 0
    for i, f in enumerate(flow_types):
        d.flow_type[i] = FLOW_CODES[f]
    return d


def ptr(t):
    """Raw address of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def current_stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(
            "%s must be a CUDA tensor: normalizingflownetwork_b200 has no CPU path "
            "(got device %s)" % (name, t.device)
        )
