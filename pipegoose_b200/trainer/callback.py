class Callback:
    """Hooks into the trainer loop (the reference only declares on_fit_start/on_fit_end)."""

    order = 0

    def on_fit_start(self, trainer, pl_module=None):
        pass

    def on_fit_end(self, trainer, pl_module=None):
        pass

    def on_epoch_start(self, trainer):
        pass

    def on_epoch_end(self, trainer):
        pass

    def on_step_end(self, trainer, loss: float):
        pass

    def on_evaluate(self, trainer, eval_loss: float):
        """After a periodic evaluation (``Trainer(eval_every=N)``) or the one at the end of ``fit``."""
        pass
