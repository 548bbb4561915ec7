"""ctypes binding of libnasr_ctc.so — the only door to the kernels.  No fallback: a missing or
unloadable library is an error, never a reason to compute somewhere else."""
from __future__ import annotations

import ctypes
import os

from ._build import LIB_PATH

OK = 0
ERR_INVALID_ARGUMENT, ERR_WORKSPACE_TOO_SMALL, ERR_CUDA, ERR_UNSUPPORTED = 1, 2, 3, 4
ST_LABEL_OUT_OF_RANGE, ST_SEQ_LEN_OUT_OF_RANGE, ST_NOT_ENOUGH_TIME, ST_NO_VALID_PATH = 1, 2, 4, 8
ABI_VERSION = 1

_vp, _i, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t

# name -> (restype, argtypes); every symbol include/nasr_ctc.h declares
SIGNATURES = {
    "nasr_abi_version": (_i, []),
    "nasr_last_error": (ctypes.c_char_p, []),
    "nasr_launch_count": (ctypes.c_uint64, []),
    "nasr_debug_config": (_i, [_i, _i]),
    "nasr_debug_profile": (_i, [_vp]),
    "nasr_allreduce_scalars": (_i, [_vp, _vp, _i, _vp]),
    "nasr_labels_coo_to_csr_i32": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "nasr_ctc_workspace_bytes": (_i, [_i, _i, _i, _i, ctypes.POINTER(_sz)]),
    "nasr_ctc_loss_grad_f32": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp,
                                    _sz, _vp]),
    "nasr_ctc_loss_grad_strided_f32": (_i, [_vp, _i, _i, _i, ctypes.c_longlong, ctypes.c_longlong, _vp, _vp, _i, _vp,
                                            _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "nasr_ctc_loss_grad_dl": (_i, [_vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nasr_ctc_greedy_decode_i64": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "nasr_ctc_greedy_decode_strided_i64": (_i, [_vp, _i, _i, _i, ctypes.c_longlong, ctypes.c_longlong, _vp, _i, _i,
                                                _vp, _vp, _vp, _vp]),
    "nasr_ctc_beam_workspace_bytes": (_i, [_i, _i, _i, _i, ctypes.POINTER(_sz)]),
    "nasr_ctc_beam_search_i64": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "nasr_ctc_beam_search_strided_i64": (_i, [_vp, _i, _i, _i, ctypes.c_longlong, ctypes.c_longlong, _vp, _i, _i,
                                              _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "nasr_hyp_to_sparse_i64": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "nasr_edit_distance_i64": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "nasr_edit_distance_csr_i64": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "nasr_batch_sums_f64": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "nasr_affine_logits_f32": (_i, [_vp, ctypes.c_longlong, _i, ctypes.c_longlong, _vp, _vp, _i, _vp,
                                    ctypes.c_longlong, _vp]),
    "nasr_affine_workspace_bytes": (_i, [ctypes.c_longlong, _i, _i, ctypes.POINTER(_sz)]),
    "nasr_affine_backward_f32": (_i, [_vp, ctypes.c_longlong, _i, ctypes.c_longlong, _vp, _i, _vp, ctypes.c_longlong,
                                      _vp, ctypes.c_longlong, _vp, _vp, _vp, _sz, _vp]),
    "nasr_host_ctx_create": (_i, [_i, _i, _i, _i, _i, ctypes.POINTER(_vp)]),
    "nasr_host_ctx_destroy": (None, [_vp]),
    "nasr_host_ctx_set_decoder": (_i, [_vp, _i, _i]),
    "nasr_host_ctx_pinned_logits": (_vp, [_vp]),
    "nasr_host_ctx_pinned_grad": (_vp, [_vp]),
    "nasr_host_ctc_step": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp,
                                _vp, _vp, _vp]),
}

_LIB = None


class NasrError(RuntimeError):
    pass


def load(path=None):
    """Load the shared library (once) and type its entry points."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = path or os.environ.get("NASR_CTC_LIB", LIB_PATH)
    if not os.path.exists(path):
        raise NasrError(
            "libnasr_ctc.so not found at %s — build it with `python -m neuralasr_b200._build` "
            "(or __graft_entry__.build()). There is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.nasr_abi_version() != ABI_VERSION:
        raise NasrError("libnasr_ctc.so ABI %d, binding expects %d" % (lib.nasr_abi_version(), ABI_VERSION))
    _LIB = lib
    return lib


def check(rc, what=""):
    if rc != OK:
        msg = load().nasr_last_error().decode("utf-8", "replace")
        kind = {1: "invalid argument", 2: "workspace too small", 3: "CUDA error", 4: "unsupported"}.get(rc, "error")
        if rc == ERR_INVALID_ARGUMENT:
            raise ValueError("%s: %s" % (what or "nasr", msg))
        raise NasrError("%s: %s: %s" % (what or "nasr", kind, msg))


def launch_count():
    return int(load().nasr_launch_count())
