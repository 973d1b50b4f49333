"""The reference's two-stage fine-tuning recipe (src/main.py:192-258, src/training_utils.py:23-48) on the B200 train steps.

    stage 1  (main.py:203-223)  backbone frozen, only ``lm_head`` trains, Adam(stage1_lr), ``stage1_epochs`` epochs
    stage 2  (main.py:227-250)  everything but the conv extractor trains, Adam(stage2_lr1); the per-epoch scheduler of
                                training_utils.py:23-31 switches to ``stage2_lr2`` after ``stage2_transition_epochs``
    per epoch (training_utils.py:35-45) a weights checkpoint is written (here: ``save_pretrained`` directories
    ``<ckpt_path>_stage1`` / ``<ckpt_path>_stage2`` instead of ``tf_model`` files), the validation loss is evaluated, and
    every ``logging_steps`` batches the running loss / lr are handed to ``log_fn`` (wandb in the reference).

Host-side control flow only: every step is ``Stage1Trainer.step`` / ``Stage2Trainer.step`` (kernels behind the C ABI, one
gradient all-reduce per step under ``torch.distributed``).  ``train_data`` / ``val_data`` are callables returning an
iterable of ``(speech [B, L] fp32, labels [B, S] int32)`` batches per epoch (the reference's tf.data pipelines,
data_utils.py, are out of scope - SURVEY 8f rank 4).
"""
from dataclasses import dataclass
from typing import Callable, Dict, Iterable, List, Optional, Tuple

import torch

from .losses import CTCLoss
from .modeling import Wav2Vec2ForCTC
from .training import Stage1Trainer, Stage2Trainer

Batches = Callable[[], Iterable[Tuple[torch.Tensor, torch.Tensor]]]


@dataclass
class FineTuneArgs:
    """The training knobs of src/main.py:28-62 that the recipe itself uses, with the reference's defaults."""
    stage1_lr: float = 1e-3
    stage1_epochs: int = 15
    stage2_lr1: float = 1e-4
    stage2_lr2: float = 5e-5
    stage2_transition_epochs: int = 10
    stage2_epochs: int = 15
    logging_steps: int = 16
    ckpt_path: Optional[str] = None
    seed: int = 42


def stage2_learning_rate(epoch: int, args: FineTuneArgs) -> float:
    """training_utils.py:23-25: ``lr1 if epoch <= transition_epochs else lr2`` (Keras epochs count from 0)."""
    return args.stage2_lr1 if epoch <= args.stage2_transition_epochs else args.stage2_lr2


@torch.no_grad()
def evaluate(model: Wav2Vec2ForCTC, loss_fn: CTCLoss, data: Iterable[Tuple[torch.Tensor, torch.Tensor]]) -> float:
    """Mean validation loss over the batches (Keras ``model.evaluate`` / ``validation_data``): eval-mode forward."""
    total, n = 0.0, 0
    for speech, labels in data:
        logits = model(speech.to(model.device), training=False)
        total += float(loss_fn(labels.to(model.device), logits))
        n += 1
    return total / max(n, 1)


def _run_stage(trainer, epochs, train_data, val_data, model, loss_fn, args, stage, lr_of_epoch, log_fn, history):
    for epoch in range(epochs):
        if lr_of_epoch is not None:
            trainer.lr = lr_of_epoch(epoch)                     # tf.keras.callbacks.LearningRateScheduler
        running, seen = 0.0, 0
        # one batch of look-ahead: stage 2 runs the NEXT batch's frozen-extractor forward while its gradient all-reduce is in flight
        lookahead = isinstance(trainer, Stage2Trainer)
        it = iter(train_data())
        nxt = next(it, None)
        if nxt is not None:
            nxt = (nxt[0].to(model.device), nxt[1].to(model.device))
        step = -1
        while nxt is not None:
            step += 1
            (speech, labels), nxt = nxt, next(it, None)
            if nxt is not None:
                nxt = (nxt[0].to(model.device), nxt[1].to(model.device))
            if lookahead:
                loss = float(trainer.step(speech, labels, next_speech=None if nxt is None else nxt[0]))
            else:
                loss = float(trainer.step(speech, labels))
            running, seen = running + loss, seen + 1
            if log_fn is not None and step % args.logging_steps == 0:
                log_fn({"stage": stage, "epoch": epoch, "step": step, "loss": running / seen, "lr": trainer.lr})
        entry = {"stage": stage, "epoch": epoch, "loss": running / max(seen, 1), "lr": trainer.lr}
        if val_data is not None:
            entry["val_loss"] = evaluate(model, loss_fn, val_data())
        if args.ckpt_path:
            model.save_pretrained(f"{args.ckpt_path}_stage{stage}")   # ModelCheckpoint(save_freq="epoch")
        history.append(entry)
        if log_fn is not None:
            log_fn(entry)


def fine_tune(model: Wav2Vec2ForCTC, loss_fn: CTCLoss, train_data: Batches, val_data: Optional[Batches] = None,
              args: Optional[FineTuneArgs] = None, log_fn: Optional[Callable[[Dict], None]] = None) -> List[Dict]:
    """Run stage 1 then stage 2; returns the per-epoch history (loss, val_loss, lr) like ``History.history``."""
    args = args or FineTuneArgs()
    history: List[Dict] = []
    if args.stage1_epochs > 0:
        t1 = Stage1Trainer(model, loss_fn, learning_rate=args.stage1_lr)
        _run_stage(t1, args.stage1_epochs, train_data, val_data, model, loss_fn, args, 1, None, log_fn, history)
    if args.stage2_epochs > 0:
        for name in model.trainable:                              # main.py:231: model.trainable = True, then the extractor is frozen
            model.trainable[name] = True
        t2 = Stage2Trainer(model, loss_fn, learning_rate=args.stage2_lr1, seed=args.seed)
        _run_stage(t2, args.stage2_epochs, train_data, val_data, model, loss_fn, args, 2,
                   lambda e: stage2_learning_rate(e, args), log_fn, history)
    return history
