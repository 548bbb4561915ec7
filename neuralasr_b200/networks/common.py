"""Drop-in ``loss`` / ``decoding`` / ``label_error_rate`` backed by sm_100a CUDA kernels.

Mirrors the three helpers every CTC model of the reference ends with
(``networks/tfnetwork.py:58-70``; module-level names of the README-era ``networks/common.py``,
``README.md:85-86,130``), same argument order minus ``self``:

    loss(logits, labels, seq_len)        <- create_loss   : reduce_mean(tf.nn.ctc_loss(labels, logits, seq_len))
    decoding(logits, seq_len)            <- create_model  : ctc_greedy_decoder(logits, seq_len) -> (decoded[0], log_prob)
    label_error_rate(model, labels)      <- create_metric : reduce_mean(edit_distance(cast(model, int32), labels))

``logits``: float32 CUDA tensor, time-major ``[T, B, C]``; ``labels``: the sparse triple
``(indices, values, shape)`` that ``utils.sparse_tuple_from`` builds (numpy or torch), or a prepared
:class:`LabelsCSR`; ``seq_len``: int32 ``[B]``.  Blank is ``C-1`` (reference convention,
``preprocess_mfcc.py:81-92``).

torch is used for device memory and streams only; all arithmetic happens in ``libnasr_ctc.so``
(``include/nasr_ctc.h``).  There is no CPU path.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib
from ..utils import sparse_to_csr

_WORKSPACES = {}


# --------------------------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------------------------
class LabelsCSR:
    """Device-resident CSR form of a label SparseTensor triple (values i32[N], offsets i32[B+1])."""

    __slots__ = ("values", "offsets", "max_len", "batch", "host_values", "host_offsets")

    def __init__(self, values, offsets, max_len, batch, host_values=None, host_offsets=None):
        self.values, self.offsets, self.max_len, self.batch = values, offsets, int(max_len), int(batch)
        self.host_values, self.host_offsets = host_values, host_offsets


def _to_numpy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def prepare_labels(labels, device):
    """Sparse triple -> :class:`LabelsCSR` on ``device`` (host prefix sum + one small H2D copy)."""
    if isinstance(labels, LabelsCSR):
        return labels
    if hasattr(labels, "indices") and hasattr(labels, "values") and hasattr(labels, "dense_shape"):
        labels = (labels.indices, labels.values, labels.dense_shape)     # SparseTensorValue-like
    indices, values, shape = labels
    if isinstance(indices, torch.Tensor) and indices.is_cuda and isinstance(values, torch.Tensor) and values.is_cuda:
        return prepare_labels_device(indices, values, shape)
    vals, offs, max_len = sparse_to_csr((_to_numpy(indices), _to_numpy(values), _to_numpy(shape)))
    dvals = torch.from_numpy(vals if vals.size else np.zeros(1, np.int32)).to(device, non_blocking=True)
    doffs = torch.from_numpy(offs).to(device, non_blocking=True)
    return LabelsCSR(dvals, doffs, max_len, offs.size - 1, vals, offs)


def prepare_labels_device(indices, values, shape, row0=0, rows=None, check=False):
    """A label triple that already lives on the GPU -> :class:`LabelsCSR` without touching the host
    (``nasr_labels_coo_to_csr_i32``): ``indices`` int64 ``[N,2]`` and ``values`` int32 ``[N]`` CUDA tensors in
    ``sparse_tuple_from``'s order (reference ``utils.py:44-58``), ``shape`` the host pair ``[B, max_len]``.
    ``row0`` / ``rows`` take a tower's block of rows, re-based (``tf.sparse_split(axis=0)``,
    ``tfnetwork.py:97-99``).  ``check=True`` reads the order flag back (one sync) and raises like TF would."""
    lib = _lib.load()
    dev = indices.device
    shape = [int(v) for v in (shape.tolist() if hasattr(shape, "tolist") else shape)]
    B = shape[0] - int(row0) if rows is None else int(rows)
    N = int(values.numel())
    if indices.dtype != torch.int64 or tuple(indices.shape) != (N, 2):
        raise ValueError("indices must be int64 [N, 2]")
    indices = indices.contiguous()
    values = values.to(torch.int32).contiguous()
    offs = torch.empty(B + 1, dtype=torch.int32, device=dev)
    info = torch.empty(3, dtype=torch.int32, device=dev)
    out = values if int(row0) == 0 and B == shape[0] else torch.empty(max(N, 1), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nasr_labels_coo_to_csr_i32(_ptr(indices), _ptr(values), N, int(row0), B, _ptr(offs),
                                                  None if out is values else _ptr(out), _ptr(info),
                                                  _stream_ptr(dev)), "nasr_labels_coo_to_csr")
    if check and int(info[0].item()):
        raise ValueError("labels: indices are not in row-major order (tf.SparseTensor canonical ordering)")
    if N == 0:
        out = torch.zeros(1, dtype=torch.int32, device=dev)
    lab = LabelsCSR(out, offs, shape[1], B)
    lab_info = info
    lab.host_values = lab_info       # (kept alive; the host copies do not exist on this path)
    return lab


def _check_logits(logits):
    if not isinstance(logits, torch.Tensor) or not logits.is_cuda:
        raise ValueError("logits must be a CUDA torch.Tensor (there is no CPU path)")
    if logits.dtype != torch.float32 or logits.dim() != 3:
        raise ValueError("logits must be float32 [T, B, C] (time-major), got %s %s"
                         % (logits.dtype, tuple(logits.shape)))
    if logits.shape[2] > 1 and logits.stride(2) != 1:
        raise ValueError("logits: the class axis must be dense (stride 1); frames and utterances may be "
                         "strided, e.g. model_output_btc.transpose(0, 1) needs no .contiguous()")
    return logits


def batch_major(logits_btc):
    """View a batch-major ``[B, T, C]`` model output as the ``[T, B, C]`` the helpers take — no copy: the
    kernels read and write through the strides, so the transpose every reference model tail does before
    ``create_loss`` (``bilstm_ctc_net.py:48``, ``lstm_ctc_net.py:43``, ``wavenet.py:171``) disappears."""
    return logits_btc.transpose(0, 1)


def _seq_len_tensor(seq_len, device, B):
    if isinstance(seq_len, torch.Tensor):
        t = seq_len.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()
    else:
        t = torch.from_numpy(np.ascontiguousarray(seq_len, dtype=np.int32)).to(device, non_blocking=True)
    if t.dim() != 1 or t.numel() != B:
        raise ValueError("seq_len must have shape [B=%d], got %s" % (B, tuple(t.shape)))
    return t


def _ws_key(device):
    """One workspace per (device, stream): calls queued on different streams must not share retry flags,
    checkpoints or the beam trie, and a buffer that is regrown is only ever in use by the stream that owns it."""
    return (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)


def _workspace(device, nbytes):
    key = _ws_key(device)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def debug_config(path=0, split_frames=0):
    """Test hook (``nasr_debug_config``): 0 = throughput kernel + robust retry (default), 1 = robust
    kernel only, 2 = throughput kernel only; ``split_frames`` forces the forward half's length."""
    _lib.check(_lib.load().nasr_debug_config(int(path), int(split_frames)), "nasr_debug_config")


def retry_flags(device, B):
    """Test hook: which utterances of the last loss call the throughput kernel handed to the robust one
    (the first B int32 of the workspace; undefined if the throughput kernel did not run)."""
    ws = _WORKSPACES[_ws_key(device)]
    off = (-ws.data_ptr()) % 256
    return ws[off:off + 4 * B].view(torch.int32).clone()


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


# --------------------------------------------------------------------------------------------
# errors: TF raises InvalidArgumentError out of sess.run; here per-utterance flags come back
# --------------------------------------------------------------------------------------------
def status_message(b, flags):
    msgs = []
    if flags & _lib.ST_LABEL_OUT_OF_RANGE:
        msgs.append("Saw a non-null label (index >= num_classes - 1) following a null label, batch: %d" % b)
    if flags & _lib.ST_SEQ_LEN_OUT_OF_RANGE:
        msgs.append("sequence_length(%d) <= max_time violated" % b)
    if flags & _lib.ST_NOT_ENOUGH_TIME:
        msgs.append("Not enough time for target transition sequence, batch: %d" % b)
    if flags & _lib.ST_NO_VALID_PATH and not msgs:
        msgs.append("No valid path found, batch: %d (loss is +inf)" % b)
    return "; ".join(msgs)


def check_status(status, raise_on_no_valid_path=False):
    """Raise ``ValueError`` with TF-like text if any utterance was flagged (synchronises)."""
    st = status.detach().cpu().numpy()
    mask = _lib.ST_LABEL_OUT_OF_RANGE | _lib.ST_SEQ_LEN_OUT_OF_RANGE | _lib.ST_NOT_ENOUGH_TIME
    if raise_on_no_valid_path:
        mask |= _lib.ST_NO_VALID_PATH
    bad = np.nonzero(st & mask)[0]
    if bad.size:
        raise ValueError(status_message(int(bad[0]), int(st[bad[0]])))


# --------------------------------------------------------------------------------------------
# CTC loss + gradient
# --------------------------------------------------------------------------------------------
def ctc_loss_and_grad(logits, labels, seq_len, grad_loss=None, want_grad=True, out_grad=None,
                      blank=None, use_dlpack=False):
    """One fused launch: per-utterance loss ``[B]``, gradient ``[T,B,C]`` (scaled by ``grad_loss[b]``,
    unit if None) and status flags ``[B]``.  Asynchronous on the current stream."""
    lib = _lib.load()
    logits = _check_logits(logits)
    T, B, C = logits.shape
    dev = logits.device
    blank = C - 1 if blank is None else int(blank)
    lab = prepare_labels(labels, dev)
    if lab.batch != B:
        raise ValueError("labels dense_shape[0]=%d but logits batch=%d" % (lab.batch, B))
    sl = _seq_len_tensor(seq_len, dev, B)
    loss_b = torch.empty(B, dtype=torch.float32, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    grad = None
    if want_grad:
        grad = out_grad if out_grad is not None else torch.empty_strided(logits.shape, logits.stride(),
                                                                         dtype=torch.float32, device=dev)
        if grad.shape != logits.shape or grad.dtype != torch.float32 or grad.stride() != logits.stride():
            raise ValueError("out_grad must be a float32 tensor with the shape and strides of logits")
    gl = None
    if grad_loss is not None:
        gl = grad_loss.to(device=dev, dtype=torch.float32).contiguous()
        if gl.numel() != B:
            raise ValueError("grad_loss must have B elements")
    need = ctypes.c_size_t(0)
    _lib.check(lib.nasr_ctc_workspace_bytes(T, B, C, lab.max_len, ctypes.byref(need)), "ctc workspace")
    ws = _workspace(dev, need.value)
    with torch.cuda.device(dev):
        if use_dlpack:
            keep = []

            def dl(t):
                if t is None:
                    return None
                cap = t.__dlpack__()
                keep.append(cap)
                ctypes.pythonapi.PyCapsule_GetPointer.restype = ctypes.c_void_p
                ctypes.pythonapi.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
                return ctypes.c_void_p(ctypes.pythonapi.PyCapsule_GetPointer(cap, b"dltensor"))

            rc = lib.nasr_ctc_loss_grad_dl(dl(logits), dl(lab.values), dl(lab.offsets), lab.max_len,
                                           dl(sl), blank, dl(loss_b), dl(grad), dl(gl), dl(status),
                                           dl(ws), _stream_ptr(dev))
            for cap in keep:   # we never took ownership: run the producer's deleter ourselves
                _release_capsule(cap)
        else:
            rc = lib.nasr_ctc_loss_grad_strided_f32(_ptr(logits), T, B, C, logits.stride(0), logits.stride(1),
                                                    _ptr(lab.values), _ptr(lab.offsets), lab.max_len, _ptr(sl),
                                                    blank, _ptr(loss_b), _ptr(grad), _ptr(gl), _ptr(status),
                                                    _ptr(ws), ws.numel(), _stream_ptr(dev))
    _lib.check(rc, "nasr_ctc_loss_grad")
    return loss_b, grad, status


class _DLManagedTensor(ctypes.Structure):
    pass


_DLManagedTensor._fields_ = [
    ("data", ctypes.c_void_p), ("device_type", ctypes.c_int32), ("device_id", ctypes.c_int32),
    ("ndim", ctypes.c_int32), ("dtype_code", ctypes.c_uint8), ("dtype_bits", ctypes.c_uint8),
    ("dtype_lanes", ctypes.c_uint16), ("shape", ctypes.c_void_p), ("strides", ctypes.c_void_p),
    ("byte_offset", ctypes.c_uint64), ("manager_ctx", ctypes.c_void_p),
    ("deleter", ctypes.CFUNCTYPE(None, ctypes.c_void_p)),
]


def _release_capsule(cap):
    """A DLPack capsule still named "dltensor" is unconsumed: call its deleter, then rename it so the
    capsule destructor does not run the deleter a second time."""
    api = ctypes.pythonapi
    api.PyCapsule_GetPointer.restype = ctypes.c_void_p
    api.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
    api.PyCapsule_SetName.argtypes = [ctypes.py_object, ctypes.c_char_p]
    p = api.PyCapsule_GetPointer(cap, b"dltensor")
    mt = ctypes.cast(p, ctypes.POINTER(_DLManagedTensor)).contents
    api.PyCapsule_SetName(cap, b"used_dltensor")
    if mt.deleter:
        mt.deleter(p)


class _CtcLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, seq_len, check):
        B = logits.shape[1]
        # d(mean loss)/d(logits) = (1/B) * per-utterance gradient (tfnetwork.py:59 reduce_mean)
        gl = torch.full((B,), 1.0 / B, dtype=torch.float32, device=logits.device)
        loss_b, grad, status = ctc_loss_and_grad(logits.detach(), labels, seq_len, grad_loss=gl,
                                                 want_grad=logits.requires_grad)
        if check:
            check_status(status)
        ctx.save_for_backward(grad if grad is not None else torch.empty(0, device=logits.device))
        sums = torch.empty(4, dtype=torch.float64, device=logits.device)
        lib = _lib.load()
        with torch.cuda.device(logits.device):
            _lib.check(lib.nasr_batch_sums_f64(_ptr(loss_b), None, None, B, _ptr(sums),
                                               _stream_ptr(logits.device)), "nasr_batch_sums")
        mean = (sums[0] / B).to(torch.float32)
        ctx.mark_non_differentiable(loss_b, status)
        return mean, loss_b, status

    @staticmethod
    def backward(ctx, g_mean, _g_loss_b, _g_status):
        (grad,) = ctx.saved_tensors
        return grad * g_mean, None, None, None


def loss(logits, labels, seq_len, check=True):
    """Mean CTC loss over the batch — drop-in for ``create_loss`` (``networks/tfnetwork.py:58-59``).

    Differentiable w.r.t. ``logits`` (the backward is the gradient the same launch already produced).
    ``check=True`` reads the status flags back and raises ``ValueError`` for what TF reports as
    ``InvalidArgumentError``; pass ``check=False`` on a hot loop to stay asynchronous.
    """
    mean, per_utt, status = _CtcLossFn.apply(_check_logits(logits), labels, seq_len, bool(check))
    mean.per_utterance = per_utt      # float32 [B]
    mean.status = status              # int32 [B] NASR_ST_* flags
    return mean


# --------------------------------------------------------------------------------------------
# the model tails' affine projection
# --------------------------------------------------------------------------------------------
def _rows2d(t, name, cols=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2:
        raise ValueError("%s must be a float32 CUDA [rows, %s] tensor (there is no CPU path)" % (name, cols or "K"))
    if t.shape[1] > 1 and t.stride(1) != 1:
        t = t.contiguous()
    if t.shape[0] > 1 and t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


def affine_logits(H, W, b=None, out=None):
    """``tf.matmul(outputs, W) + b`` of the reference's tails (``bilstm_ctc_net.py:33-45``): ``H`` float32 CUDA
    ``[rows, K]`` (rows may be strided), ``W`` ``[K, C]``, ``b`` ``[C]`` -> ``[rows, C]`` (no autograd: see
    :func:`affine_projection`)."""
    H = _rows2d(H, "H")
    if W.dim() != 2 or W.shape[0] != H.shape[1] or W.dtype != torch.float32 or not W.is_cuda:
        raise ValueError("W must be a float32 CUDA [K=%d, C] tensor, got %s" % (H.shape[1], tuple(W.shape)))
    W = W.contiguous()
    rows, K = H.shape
    C = W.shape[1]
    if b is not None:
        if b.shape != (C,) or b.dtype != torch.float32 or not b.is_cuda:
            raise ValueError("b must be a float32 CUDA [C=%d] tensor" % C)
        b = b.contiguous()
    if out is None:
        out = torch.empty((rows, C), dtype=torch.float32, device=H.device)
    lib = _lib.load()
    with torch.cuda.device(H.device):
        _lib.check(lib.nasr_affine_logits_f32(_ptr(H), rows, K, H.stride(0) if rows > 1 else K, _ptr(W), _ptr(b), C,
                                              _ptr(out), out.stride(0) if rows > 1 else C,
                                              _stream_ptr(H.device)), "nasr_affine_logits_f32")
    return out


def affine_backward(H, W, dlogits, want_dH=True, want_dW=True, want_db=True, dense_dH=False):
    """The three gradients of :func:`affine_logits`: ``(dH [rows, K], dW [K, C], db [C])`` (``None`` where not
    wanted).  ``dH`` is a ``[:, :K]`` view of rows pitched at a multiple of 32 floats unless ``dense_dH`` asks for a
    contiguous tensor (what autograd hands on without a copy)."""
    H = _rows2d(H, "H")
    dlogits = _rows2d(dlogits, "dlogits", "C")
    W = W.contiguous()
    rows, K = H.shape
    C = W.shape[1]
    dev = H.device
    # dH rows start on 128-byte lines: the row pitch is K rounded up to 32 floats and the result is the [:, :K] view.
    # (Measured at 256000 x 500: 0.137 ms at a pitch of 512, 0.166 at 504, 0.208 for the dense pitch of 500 floats,
    # whose odd rows start 16 bytes into a sector so that every 128-byte piece the kernel stores ends in two
    # half-written sectors.)
    pitch = K if dense_dH else (K + 31) // 32 * 32
    dH = torch.empty((rows, pitch), dtype=torch.float32, device=dev)[:, :K] if want_dH else None
    dW = torch.empty((K, C), dtype=torch.float32, device=dev) if want_dW else None
    db = torch.empty((C,), dtype=torch.float32, device=dev) if want_db else None
    lib = _lib.load()
    ws, nbytes = None, 0
    if want_dW or want_db:
        need = ctypes.c_size_t(0)
        with torch.cuda.device(dev):
            _lib.check(lib.nasr_affine_workspace_bytes(rows, K, C, ctypes.byref(need)), "nasr_affine_workspace_bytes")
        nbytes = int(need.value)
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nasr_affine_backward_f32(_ptr(H), rows, K, H.stride(0) if rows > 1 else K, _ptr(W), C,
                                                _ptr(dlogits), dlogits.stride(0) if rows > 1 else C, _ptr(dH),
                                                dH.stride(0) if (want_dH and rows > 1) else K,
                                                _ptr(dW), _ptr(db), _ptr(ws), nbytes, _stream_ptr(dev)),
                   "nasr_affine_backward_f32")
    return dH, dW, db


class _AffineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, H, W, b):
        ctx.save_for_backward(H, W)
        ctx.has_bias = b is not None
        return affine_logits(H.detach(), W.detach(), None if b is None else b.detach())

    @staticmethod
    def backward(ctx, g):
        H, W = ctx.saved_tensors
        dH, dW, db = affine_backward(H, W, g, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                     ctx.has_bias and ctx.needs_input_grad[2], dense_dH=True)
        return dH, dW, db


def affine_projection(outputs, W, b, batch_size):
    """The tail every CTC model of the reference ends with (``bilstm_ctc_net.py:31-48``, ``lstm_ctc_net.py:26-43``)::

        outputs = tf.reshape(outputs, [-1, num_hidden]); logits = tf.matmul(outputs, W) + b
        logits = tf.reshape(logits, [batch_s, -1, num_classes]); logits = tf.transpose(logits, (1, 0, 2))

    ``outputs`` float32 CUDA ``[..., num_hidden]`` (batch-major, any leading shape — the BiLSTM tail reshapes the
    (forward, backward) pair, which stacks both directions along the rows).  Returns the time-major ``[T', B, C]``
    VIEW of the batch-major result: no transpose is executed, ``loss`` / ``decoding`` read it through its strides.
    Differentiable w.r.t. ``outputs``, ``W`` and ``b``."""
    K = W.shape[0]
    H = outputs.reshape(-1, K)
    logits = _AffineFn.apply(H, W, b)
    return batch_major(logits.reshape(int(batch_size), -1, W.shape[1]))


# --------------------------------------------------------------------------------------------
# greedy decode
# --------------------------------------------------------------------------------------------
class DecodedSparse:
    """``decoded[0]`` of ``tf.nn.ctc_greedy_decoder``: unpacks / indexes as
    ``(indices i64[M,2], values i64[M], dense_shape i64[2])``; materialised lazily because M is
    data dependent (one host sync).  ``hyp`` [B,T] / ``hyp_len`` [B] keep the dense device form."""

    def __init__(self, hyp, hyp_len):
        self.hyp, self.hyp_len = hyp, hyp_len
        self._triple = None

    def _materialise(self):
        if self._triple is None:
            lib = _lib.load()
            B, T = self.hyp.shape
            dev = self.hyp.device
            offs = torch.zeros(B + 1, dtype=torch.int32, device=dev)
            torch.cumsum(self.hyp_len, 0, out=offs[1:])
            M = int(offs[-1].item())
            indices = torch.empty((M, 2), dtype=torch.int64, device=dev)
            values = torch.empty(M, dtype=torch.int64, device=dev)
            shape = torch.empty(2, dtype=torch.int64, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.nasr_hyp_to_sparse_i64(_ptr(self.hyp), self.hyp.stride(0), _ptr(offs), B, _ptr(indices),
                                                      _ptr(values), _ptr(shape), _stream_ptr(dev)),
                           "nasr_hyp_to_sparse")
            self._triple = (indices, values, shape)
        return self._triple

    indices = property(lambda self: self._materialise()[0])
    values = property(lambda self: self._materialise()[1])
    dense_shape = property(lambda self: self._materialise()[2])

    def __iter__(self):
        return iter(self._materialise())

    def __getitem__(self, i):
        return self._materialise()[i]

    def __len__(self):
        return 3


def decoding(logits, seq_len, merge_repeated=True, blank=None):
    """Greedy CTC decode — drop-in for ``create_model`` with the greedy op of
    ``networks/tfnetwork.py:63``.  Returns ``(decoded, neg_sum_logits[B,1])`` like the reference's
    ``(model[0], log_prob)``."""
    lib = _lib.load()
    logits = _check_logits(logits)
    T, B, C = logits.shape
    dev = logits.device
    blank = C - 1 if blank is None else int(blank)
    sl = _seq_len_tensor(seq_len, dev, B)
    hyp = torch.empty((B, T), dtype=torch.int64, device=dev)
    hyp_len = torch.empty(B, dtype=torch.int32, device=dev)
    nsl = torch.empty((B, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nasr_ctc_greedy_decode_strided_i64(_ptr(logits), T, B, C, logits.stride(0),
                                                          logits.stride(1), _ptr(sl), blank,
                                                          int(bool(merge_repeated)), _ptr(hyp), _ptr(hyp_len),
                                                          _ptr(nsl), _stream_ptr(dev)), "nasr_ctc_greedy_decode")
    return DecodedSparse(hyp, hyp_len), nsl


# --------------------------------------------------------------------------------------------
# beam search decode
# --------------------------------------------------------------------------------------------
_BEAM_WORKSPACES = {}


def beam_decoding(logits, seq_len, beam_width=100, top_paths=1, merge_repeated=True, blank=None):
    """CTC beam search — ``tf.nn.ctc_beam_search_decoder(logits, seq_len, beam_width, top_paths,
    merge_repeated)``, the op ``create_model`` runs (``networks/tfnetwork.py:62``).  Returns
    ``(decoded, log_probability)`` like TF: ``decoded`` is a list of ``top_paths`` :class:`DecodedSparse`
    (best first), ``log_probability`` float32 ``[B, top_paths]``."""
    lib = _lib.load()
    logits = _check_logits(logits)
    T, B, C = logits.shape
    dev = logits.device
    blank = C - 1 if blank is None else int(blank)
    W, P = int(beam_width), int(top_paths)
    sl = _seq_len_tensor(seq_len, dev, B)
    hyp = torch.empty((B, P, T), dtype=torch.int64, device=dev)
    hyp_len = torch.empty((B, P), dtype=torch.int32, device=dev)
    log_prob = torch.empty((B, P), dtype=torch.float32, device=dev)
    need = ctypes.c_size_t(0)
    _lib.check(lib.nasr_ctc_beam_workspace_bytes(T, B, C, W, ctypes.byref(need)), "beam workspace")
    key = _ws_key(dev)
    ws = _BEAM_WORKSPACES.get(key)
    if ws is None or ws.numel() < need.value:
        ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=dev)
        _BEAM_WORKSPACES[key] = ws
    with torch.cuda.device(dev):
        _lib.check(lib.nasr_ctc_beam_search_strided_i64(_ptr(logits), T, B, C, logits.stride(0), logits.stride(1),
                                                        _ptr(sl), blank, W, P, int(bool(merge_repeated)),
                                                        _ptr(hyp), _ptr(hyp_len), _ptr(log_prob), _ptr(ws),
                                                        ws.numel(), _stream_ptr(dev)), "nasr_ctc_beam_search")
    decoded = [DecodedSparse(hyp[:, p, :], hyp_len[:, p].contiguous()) for p in range(P)]
    return decoded, log_prob


def create_model_beam(logits, seq_len):
    """``create_model`` exactly as the snapshot has it (``networks/tfnetwork.py:61-64``): beam search with TF's
    defaults, returning ``(decoded[0], log_prob)``."""
    decoded, log_prob = beam_decoding(logits, seq_len)
    return decoded[0], log_prob


# --------------------------------------------------------------------------------------------
# label error rate
# --------------------------------------------------------------------------------------------
def edit_distance(model, labels, normalize=True):
    """Per-utterance ``(dist i32[B], ler f32[B])`` — ``tf.edit_distance(cast(model, int32), labels)``."""
    lib = _lib.load()
    if isinstance(model, DecodedSparse):
        dev = model.hyp.device
        B, T = model.hyp.shape
        lab = prepare_labels(labels, dev)
        if lab.batch != B:
            raise ValueError("hypothesis batch %d != labels batch %d" % (B, lab.batch))
        dist = torch.empty(B, dtype=torch.int32, device=dev)
        ler = torch.empty(B, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.nasr_edit_distance_i64(_ptr(model.hyp), model.hyp.stride(0), _ptr(model.hyp_len),
                                                  _ptr(lab.values), _ptr(lab.offsets), lab.max_len, B,
                                                  int(bool(normalize)), _ptr(dist), _ptr(ler),
                                                  _stream_ptr(dev)), "nasr_edit_distance")
        return dist, ler
    # a raw SparseTensor triple (e.g. LAS's dense_to_sparse output, las.py:116-117)
    indices, values, shape = model
    if isinstance(labels, LabelsCSR):
        dev = labels.values.device
    elif isinstance(values, torch.Tensor) and values.is_cuda:
        dev = values.device
    else:
        dev = torch.device("cuda", torch.cuda.current_device())
    hv, ho, hmax = sparse_to_csr((_to_numpy(indices), _to_numpy(values), _to_numpy(shape)))
    lab = prepare_labels(labels, dev)
    B = ho.size - 1
    if lab.batch != B:
        raise ValueError("hypothesis batch %d != labels batch %d" % (B, lab.batch))
    dhv = torch.from_numpy((hv if hv.size else np.zeros(1, np.int32)).astype(np.int64)).to(dev)
    dho = torch.from_numpy(ho).to(dev)
    dist = torch.empty(B, dtype=torch.int32, device=dev)
    ler = torch.empty(B, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nasr_edit_distance_csr_i64(_ptr(dhv), _ptr(dho), hmax, _ptr(lab.values),
                                                  _ptr(lab.offsets), lab.max_len, B,
                                                  int(bool(normalize)), _ptr(dist), _ptr(ler),
                                                  _stream_ptr(dev)), "nasr_edit_distance_csr")
    return dist, ler


def batch_sums(loss_b=None, ler=None, dist=None):
    """``float64[4]`` device vector ``[sum loss, sum ler, sum dist, B]`` (what each rank all-reduces)."""
    lib = _lib.load()
    ref = next(t for t in (loss_b, ler, dist) if t is not None)
    B = ref.numel()
    sums = torch.empty(4, dtype=torch.float64, device=ref.device)
    with torch.cuda.device(ref.device):
        _lib.check(lib.nasr_batch_sums_f64(_ptr(loss_b), _ptr(ler), _ptr(dist), B, _ptr(sums),
                                           _stream_ptr(ref.device)), "nasr_batch_sums")
    return sums


def label_error_rate(model, labels):
    """Mean normalised edit distance — drop-in for ``create_metric`` (``networks/tfnetwork.py:66-70``).
    Returns a 0-dim float32 CUDA tensor with ``.per_utterance`` (ler[B]) and ``.distances`` (i32[B])."""
    dist, ler = edit_distance(model, labels, normalize=True)
    sums = batch_sums(ler=ler, dist=dist)
    mean = (sums[1] / ler.numel()).to(torch.float32)
    mean.per_utterance = ler
    mean.distances = dist
    return mean


# reference-era aliases (the snapshot's method names, networks/tfnetwork.py:58,61,66).  ``create_model`` is what the
# snapshot's method of that name runs -- the beam search decoder, width 100, returning (decoded[0], log_prob)
# (tfnetwork.py:62,64); ``decoding`` / ``model`` keep the README-era name for the greedy decoder the north star asks for
# (the commented alternative on tfnetwork.py:63).
create_loss = loss
create_model = create_model_beam
create_metric = label_error_rate
model = decoding
