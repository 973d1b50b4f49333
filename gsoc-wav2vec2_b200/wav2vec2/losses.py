"""CTC loss behind the reference's ``CTCLoss(config, model_input_shape, division_factor)`` surface
(src/wav2vec2/losses.py:4-56), computed by the sm_100a alpha-beta kernel (csrc/ctc.cu)."""
import torch

from . import ops


class CTCLoss:
    def __init__(self, config, model_input_shape, division_factor=1):
        self.kernal_sizes = config.kernal_sizes
        self.strides = config.strides
        self.pad_id = config.pad_id
        self.division_factor = division_factor
        self.model_input_shape = model_input_shape
        self.last_grad = None

    def _get_logit_length(self, input_length):
        # losses.py:47-56
        for k, s in zip(self.kernal_sizes, self.strides):
            input_length = 1 + (input_length - k) // s
        return input_length

    def __call__(self, labels, hidden_states, return_grad=False):
        """labels [B, S] int (pad_id-padded), hidden_states [B, T', vocab] fp32 logits -> scalar loss.

        Like the reference (losses.py:29-30) every utterance is scored over the constant frame count
        derived from ``model_input_shape``; it must equal the logits' time dimension.
        """
        T = hidden_states.shape[1]
        expected = self._get_logit_length(int(self.model_input_shape[1]))
        if expected != T:
            raise ValueError(f"logits have {T} frames but model_input_shape implies {expected}")
        per_sample, grad = ops.ctc_loss(hidden_states, labels.to(hidden_states.device), self.pad_id,
                                        1.0 / float(self.division_factor), want_grad=return_grad)
        loss = per_sample.sum()     # Keras Reduction.SUM (losses.py:6)
        self.last_grad = grad
        return (loss, grad) if return_grad else loss

    call = __call__
