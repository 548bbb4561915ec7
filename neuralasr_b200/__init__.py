"""neuralasr_b200 — B200-native CTC loss / greedy decode / label-error-rate path.

Drop-in for the three helpers every CTC network of zeahmed/NeuralASR ends with
(``networks/tfnetwork.py:58-70``; README-era ``networks/common.py``):

    from neuralasr_b200.networks.common import loss, decoding, label_error_rate

The work is done by hand-written sm_100a CUDA kernels in ``lib/libnasr_ctc.so``
(sources in ``csrc/``, C-ABI in ``include/nasr_ctc.h``).  There is no CPU path:
importing the kernels' binding without the built library raises.
"""
__version__ = "0.1.0"
