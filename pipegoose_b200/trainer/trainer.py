"""A minimal working trainer (the reference's Trainer.fit/train bodies are ``pass``,
trainer/trainer.py:13-35): epochs over a dataloader of ``{"input_ids", ["attention_mask"], ["labels"]}``
batches, loss/backward/step, callbacks, tokens/s logging with device-side timing."""
from __future__ import annotations

import time
from typing import List, Optional

import torch
from torch import nn

from pipegoose_b200.trainer.callback import Callback
from pipegoose_b200.trainer.logger import DistributedLogger
from pipegoose_b200.trainer.state import TrainerStage, TrainerState, TrainerStatus


class Trainer:
    def __init__(self, module: nn.Module, train_loader, eval_loader=None, optim=None, num_epochs: int = 1,
                 callbacks: Optional[List[Callback]] = None, loggers: Optional[List[DistributedLogger]] = None,
                 parallel_context=None, log_every: int = 10):
        self.module = module
        self.train_loader = train_loader
        self.eval_loader = eval_loader
        self.optim = optim
        self.num_epochs = num_epochs
        self.callbacks = sorted(callbacks or [], key=lambda c: c.order)
        self.loggers = loggers or []
        self.parallel_context = parallel_context
        self.log_every = log_every
        self.state = TrainerState()

    def _device(self):
        return next(self.module.parameters()).device

    def _call(self, name, *args):
        for cb in self.callbacks:
            getattr(cb, name)(self, *args)

    def _log(self, msg):
        for lg in self.loggers:
            lg.info(msg)

    def train_step(self, batch) -> torch.Tensor:
        dev = self._device()
        batch = {k: (v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
        labels = batch.pop("labels", batch["input_ids"])
        out = self.module(**batch, labels=labels)
        loss = out.loss if hasattr(out, "loss") else out[0]
        self.optim.zero_grad()
        loss.backward()
        self.optim.step()
        self.state.tokens_seen += int(batch["input_ids"].numel())
        return loss

    def fit(self):
        self.state.status = TrainerStatus.RUNNING
        self._call("on_fit_start")
        self.train()
        self._call("on_fit_end")
        self.state.status = TrainerStatus.FINISHED
        return self.state

    def train(self):
        self.state.stage = TrainerStage.TRAINING
        self.module.train()
        t0, tok0 = time.time(), self.state.tokens_seen
        for epoch in range(self.num_epochs):
            self.state.epoch = epoch
            self._call("on_epoch_start")
            for batch in self.train_loader:
                loss = self.train_step(batch)
                self.state.step += 1
                if self.state.step % self.log_every == 0:
                    self.state.last_loss = float(loss.item())
                    dt = max(time.time() - t0, 1e-9)
                    self._log(f"step {self.state.step} loss {self.state.last_loss:.4f} "
                              f"tokens/s {(self.state.tokens_seen - tok0) / dt:.0f}")
                self._call("on_step_end", loss)
            self._call("on_epoch_end")

    @torch.no_grad()
    def evaluate(self) -> float:
        assert self.eval_loader is not None
        self.state.stage = TrainerStage.VALIDATING
        self.module.eval()
        dev = self._device()
        total, n = 0.0, 0
        for batch in self.eval_loader:
            batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
            labels = batch.pop("labels", batch["input_ids"])
            out = self.module(**batch, labels=labels)
            total += float((out.loss if hasattr(out, "loss") else out[0]).item())
            n += 1
        return total / max(n, 1)
