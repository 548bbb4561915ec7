"""CPU oracle (test infrastructure, never shipped code) for the model tails' affine projection.

Restates ``networks/bilstm_ctc_net.py:31-48`` / ``lstm_ctc_net.py:26-43`` of the reference: reshape the recurrent
outputs to ``[-1, num_hidden]``, ``tf.matmul(outputs, W) + b``, reshape to ``[batch_s, -1, num_classes]``, transpose to
time-major — and the gradients TensorFlow's ``MatMul`` / ``BiasAdd`` gradient functions derive from it.  float64 numpy:
the kernels (float32, 3xTF32 products) are compared against it with a bound proportional to ``|H|.|W|``.  Only
``tests/`` and ``bench.py``'s CPU legs import this file.
"""
import numpy as np


def affine_logits(H, W, b=None):
    """``[rows, K] x [K, C] + [C]`` in float64."""
    out = np.asarray(H, np.float64) @ np.asarray(W, np.float64)
    if b is not None:
        out = out + np.asarray(b, np.float64)
    return out


def affine_backward(H, W, dlogits):
    """-> (dH = dL.W^T, dW = H^T.dL, db = column sums of dL) in float64."""
    H, W, dL = (np.asarray(a, np.float64) for a in (H, W, dlogits))
    return dL @ W.T, H.T @ dL, dL.sum(axis=0)


def tail(outputs, W, b, batch_size):
    """The whole tail: time-major ``[T', B, C]`` logits (a transposed copy, as the reference materialises it)."""
    K = W.shape[0]
    logits = affine_logits(np.asarray(outputs).reshape(-1, K), W, b)
    return np.ascontiguousarray(logits.reshape(int(batch_size), -1, W.shape[1]).transpose(1, 0, 2))
