"""HOST-buffer entry: numpy in, numpy out, one call per step (the reference's feed_dict world,
``networks/tfnetwork.py:183-190``: host arrays in, ``(loss, mean_ler)`` out).

Thin wrapper over ``nasr_host_ctx_*`` / ``nasr_host_ctc_step`` of the C-ABI: the context owns device
buffers, pinned staging and a stream; a step is H2D(logits, labels, seq_len) -> loss+grad kernel ->
greedy decode -> edit distance -> D2H(loss, grad, status, hyp_len, dist, ler) -> stream sync.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def bind_to_gpu_numa_node(device):
    """Restrict this process to the CPUs NVML reports as local to GPU ``device``, so that the pinned staging a
    :class:`HostContext` allocates next (and the threads that fill it) sit on the NUMA node the GPU's PCIe link
    hangs off.  With one process per GPU on a two-socket box this keeps H2D/D2H traffic off the socket
    interconnect.  Returns the CPU set, or None if NVML or the affinity call is unavailable (never an error)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[device]) if visible and visible.split(",")[device].isdigit() else int(device)
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


class HostContext:
    def __init__(self, device, max_T, max_B, max_C, max_label_len, decoder="greedy", beam_width=100):
        """``decoder``: ``"greedy"`` (``tfnetwork.py:63``) or ``"beam"`` (``tfnetwork.py:62``: what the snapshot's
        ``train``/``evaluate`` fetch; ``neg_sum_logits`` of a step is then the top path's log probability)."""
        self._lib = _lib.load()
        self._ctx = ctypes.c_void_p()
        _lib.check(self._lib.nasr_host_ctx_create(int(device), int(max_T), int(max_B), int(max_C),
                                                  int(max_label_len), ctypes.byref(self._ctx)),
                   "nasr_host_ctx_create")
        self.max = (int(max_T), int(max_B), int(max_C))
        n = self.max[0] * self.max[1] * self.max[2]
        fl = ctypes.POINTER(ctypes.c_float)
        self.pinned_logits = np.ctypeslib.as_array(
            ctypes.cast(self._lib.nasr_host_ctx_pinned_logits(self._ctx), fl), shape=(n,))
        self.pinned_grad = np.ctypeslib.as_array(
            ctypes.cast(self._lib.nasr_host_ctx_pinned_grad(self._ctx), fl), shape=(n,))
        self.set_decoder(decoder, beam_width)

    def set_decoder(self, decoder="greedy", beam_width=100):
        if decoder not in ("greedy", "beam"):
            raise ValueError("decoder must be 'greedy' or 'beam', got %r" % (decoder,))
        _lib.check(self._lib.nasr_host_ctx_set_decoder(self._ctx, int(decoder == "beam"), int(beam_width)),
                   "nasr_host_ctx_set_decoder")
        self.decoder = decoder

    def step(self, logits, label_values, label_offsets, seq_len, blank=None, grad_loss=None,
             want_grad=True, want_decode=True, want_hyp=False, grad_out=None):
        """``logits`` float32 [T,B,C] (a view of ``pinned_logits`` skips one host copy).
        Returns a dict of numpy arrays."""
        T, B, C = logits.shape
        if logits.dtype != np.float32 or not logits.flags.c_contiguous:
            logits = np.ascontiguousarray(logits, dtype=np.float32)
        blank = C - 1 if blank is None else int(blank)
        lv = np.ascontiguousarray(label_values, dtype=np.int32)
        lo = np.ascontiguousarray(label_offsets, dtype=np.int32)
        sl = np.ascontiguousarray(seq_len, dtype=np.int32)
        gl = None if grad_loss is None else np.ascontiguousarray(grad_loss, dtype=np.float32)
        loss = np.empty(B, np.float32)
        status = np.empty(B, np.int32)
        grad = None
        if want_grad:
            grad = grad_out if grad_out is not None else self.pinned_grad[: T * B * C].reshape(T, B, C)
        hyp = np.empty((B, T), np.int64) if (want_decode and want_hyp) else None
        hyp_len = np.empty(B, np.int32) if want_decode else None
        nsl = np.empty(B, np.float32) if want_decode else None
        dist = np.empty(B, np.int32) if want_decode else None
        ler = np.empty(B, np.float32) if want_decode else None
        rc = self._lib.nasr_host_ctc_step(self._ctx, _p(logits), T, B, C, _p(lv), _p(lo), _p(sl), blank,
                                          _p(gl), _p(loss), _p(grad), _p(status), _p(hyp), _p(hyp_len),
                                          _p(nsl), _p(dist), _p(ler))
        _lib.check(rc, "nasr_host_ctc_step")
        return dict(loss=loss, grad=grad, status=status, hyp=hyp, hyp_len=hyp_len,
                    neg_sum_logits=nsl, dist=dist, ler=ler)

    def close(self):
        if self._ctx:
            self._lib.nasr_host_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
