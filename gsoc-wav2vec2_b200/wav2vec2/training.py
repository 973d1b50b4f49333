"""Stage-1 fine-tuning step of the reference recipe on B200 kernels.

Reference (src/main.py:210-223): ``model.layers[0].trainable = False`` freezes the whole Wav2Vec2 body, so the only
trainable variables are ``lm_head/kernel`` and ``lm_head/bias`` (24 608 parameters for the base model);
``optimizer = Adam(1e-3)``; ``loss = CTCLoss(config, input_shape, division_factor=global_batch)``; Keras ``fit`` then runs,
per step: forward (training=True) -> CTC loss -> gradients -> cross-replica SUM all-reduce (MirroredStrategy/TPUStrategy,
main.py:141-156) -> Adam.  Here every arithmetic step is a kernel behind the C ABI:
forward (all the inference kernels), ``w2v2_ctc_loss`` (loss + d loss / d logits), ``w2v2_lm_head_wgrad``, ONE
``torch.distributed.all_reduce`` over the flat gradient buffer (NCCL on GPUs), ``w2v2_adam``.

Not covered yet (stage 2, main.py:234-250): gradients through the encoder.  Dropout RNG is not implemented, so the step
requires ``config.dropout == 0`` (SpecAugment, main.py/modeling.py:193-199, is applied when enabled).
"""
import torch
import torch.distributed as dist

from . import ops
from .losses import CTCLoss
from .modeling import Wav2Vec2ForCTC


class Stage1Trainer:
    def __init__(self, model: Wav2Vec2ForCTC, loss_fn: CTCLoss, learning_rate=1e-3, beta_1=0.9, beta_2=0.999,
                 epsilon=1e-7):
        self.model, self.loss_fn = model, loss_fn
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.t = 0
        k, b = model.variables["lm_head/kernel"], model.variables["lm_head/bias"]
        self.d, self.V = k.shape
        # trainable variables live in ONE flat fp32 buffer (a single all-reduce message, a single Adam launch)
        self.flat_w = torch.cat([k.reshape(-1), b.reshape(-1)]).contiguous()
        model.variables["lm_head/kernel"] = self.flat_w[: self.d * self.V].view(self.d, self.V)
        model.variables["lm_head/bias"] = self.flat_w[self.d * self.V:]
        self.flat_g = torch.zeros_like(self.flat_w)
        self.m = torch.zeros_like(self.flat_w)
        self.v = torch.zeros_like(self.flat_w)
        for name in model.trainable:
            model.trainable[name] = name.startswith("lm_head/")

    @torch.no_grad()
    def step(self, speech, labels, attention_mask=None):
        """One optimisation step on this rank's shard; returns this rank's (already 1/division_factor-scaled) loss."""
        model = self.model
        logits, hidden = model.forward_with_hidden(speech, attention_mask=attention_mask, training=True)
        loss, grad_logits = self.loss_fn(labels, logits, return_grad=True)
        B, T, V = logits.shape
        gk = self.flat_g[: self.d * self.V].view(self.d, self.V)
        gb = self.flat_g[self.d * self.V:]
        ops.lm_head_wgrad(hidden, grad_logits.view(B * T, V), gk, gb)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)          # the step's single collective
        self.t += 1
        lr_t = self.lr * (1.0 - self.b2 ** self.t) ** 0.5 / (1.0 - self.b1 ** self.t)
        ops.adam(self.flat_w, self.flat_g, self.m, self.v, lr_t, self.b1, self.b2, self.eps)
        model._repack_lm_head()
        return loss
