"""Step functions of the hot path with the reference's return conventions.

``TensorFlowNetwork.train / validate / evaluate / decode`` (``networks/tfnetwork.py:166-190``) each make one
``sess.run`` that fetches loss, mean label error rate and/or the decoded ``SparseTensorValue`` and hand numpy
scalars / the flat ``.values`` array back to ``train.py`` / ``decode.py``.  The acoustic model is out of
scope here, so these take the model tail's logits instead of MFCCs; everything after the logits is the same
contract:

    train(logits, labels, seq_len, labels_len)    -> (loss_val, mean_ler_value)        tfnetwork.py:183-190
    validate(...)                                 -> [loss, mean_ler]                  tfnetwork.py:166-170
    evaluate(...)                                 -> (decoded.values, loss, ler)       tfnetwork.py:172-177
    decode(logits, seq_len)                       -> decoded.values                    tfnetwork.py:179-181

``labels`` is the dense padded ``[B, Lmax]`` array + ``labels_len`` that ``DataSet.get_next_batch`` yields
(``dataset.py:48-82``); it goes through ``sparse_tuple_from`` exactly like in the reference.  With a process
group initialised the scalars are reduced over the towers (``towers.all_reduce_sums``).
"""
from __future__ import annotations

import numpy as np
import torch

from . import towers
from .networks import common
from .utils import sparse_tuple_from


class CtcHead:
    """Everything a CTC model tail does after it has logits (``bilstm_ctc_net.py:47-52``)."""

    def __init__(self, time_major=True, decoder="greedy", beam_width=100):
        """``decoder``: ``"greedy"`` (the north star's ``decoding``, ``tfnetwork.py:63``) or ``"beam"`` (what the
        snapshot's ``create_model`` runs, ``tfnetwork.py:62``: beam search, width 100, top path)."""
        if decoder not in ("greedy", "beam"):
            raise ValueError("decoder must be 'greedy' or 'beam', got %r" % (decoder,))
        self.time_major = time_major
        self.decoder, self.beam_width = decoder, int(beam_width)
        self.global_step = 0

    def create_model(self, logits, seq_len):
        """``(decoded[0], log_prob)`` as ``create_model`` returns them (``tfnetwork.py:61-64``)."""
        if self.decoder == "beam":
            decoded, log_prob = common.beam_decoding(logits, seq_len, beam_width=self.beam_width)
            return decoded[0], log_prob
        return common.decoding(logits, seq_len)

    def _view(self, logits):
        return logits if self.time_major else common.batch_major(logits)

    def create_network_tail(self, logits, labels, seq_len):
        """``(loss, model, log_prob, ler)`` — the four values every CTC ``create_network`` returns after
        ``logits`` (``lstm_ctc_net.py:44-47``)."""
        x = self._view(logits)
        lab = common.prepare_labels(labels, x.device)
        loss = common.loss(x, lab, seq_len)
        model, log_prob = self.create_model(x.detach(), seq_len)
        ler = common.label_error_rate(model, lab)
        return loss, model, log_prob, ler

    def create_network(self, outputs, W, b, labels, seq_len, batch_size):
        """What ``create_network`` of the reference's CTC models does after the recurrent layers
        (``bilstm_ctc_net.py:31-52``, ``lstm_ctc_net.py:26-47``): reshape the outputs to ``[-1, num_hidden]``, the affine
        projection ``outputs . W + b``, reshape to ``[batch_s, -1, num_classes]``, time-major, then the three helpers.
        Returns the reference's five values ``(logits, loss, model, prob, ler)``; ``logits`` is the time-major view of
        the batch-major projection result (no transpose is executed) and the loss backpropagates to ``outputs``, ``W``
        and ``b``."""
        logits = common.affine_projection(outputs, W, b, batch_size)
        lab = common.prepare_labels(labels, logits.device)
        loss = common.loss(logits, lab, seq_len)
        model, log_prob = self.create_model(logits.detach(), seq_len)
        ler = common.label_error_rate(model, lab)
        return logits, loss, model, log_prob, ler

    def _scalars(self, loss, ler):
        sums = common.batch_sums(loss_b=loss.per_utterance, ler=ler.per_utterance, dist=ler.distances)
        mean_loss, mean_ler, _, _ = towers.step_scalars(towers.all_reduce_sums(sums))
        return np.float32(mean_loss), np.float32(mean_ler)

    def train(self, logits, labels, seq_len, labels_len):
        """One training step of the path: loss + backward into ``logits.grad`` (if it requires grad), decode,
        label error rate.  Returns ``(loss_val, mean_ler_value)``."""
        self.global_step += 1
        loss, _, _, ler = self.create_network_tail(logits, sparse_tuple_from(labels, labels_len), seq_len)
        if logits.requires_grad:
            loss.backward()
        return self._scalars(loss, ler)

    def validate(self, logits, labels, seq_len, labels_len):
        with torch.no_grad():
            loss, _, _, ler = self.create_network_tail(logits, sparse_tuple_from(labels, labels_len), seq_len)
        return list(self._scalars(loss, ler))

    def evaluate(self, logits, labels, seq_len, labels_len):
        with torch.no_grad():
            loss, model, _, ler = self.create_network_tail(logits, sparse_tuple_from(labels, labels_len), seq_len)
        mean_loss, mean_ler = self._scalars(loss, ler)
        return model.values.cpu().numpy(), mean_loss, mean_ler

    def decode(self, logits, seq_len):
        with torch.no_grad():
            model, _ = self.create_model(self._view(logits), seq_len)
        return model.values.cpu().numpy()


def dense_to_sparse(dense, eos_token=0):
    """``tf.contrib.layers.dense_to_sparse``: the non-``eos_token`` entries of a dense ``[B, L]`` id matrix as
    a sparse triple — how the LAS network feeds ``create_metric`` (``networks/las.py:116-117``)."""
    dense = np.asarray(dense.detach().cpu() if isinstance(dense, torch.Tensor) else dense)
    rows, cols = np.nonzero(dense != eos_token)
    indices = np.stack([rows, cols], 1).astype(np.int64)
    return indices, dense[rows, cols], np.asarray(dense.shape, dtype=np.int64)
