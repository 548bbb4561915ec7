"""Host-side mirror of the reference's network helper interface for the CTC path."""
from .common import (LabelsCSR, DecodedSparse, check_status, ctc_loss_and_grad, decoding,  # noqa: F401
                     edit_distance, label_error_rate, loss, prepare_labels, batch_sums)
