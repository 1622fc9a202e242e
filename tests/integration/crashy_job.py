"""A 2-replica training job for tests/integration/test_fault_tolerance.py: on its FIRST launch data-parallel rank 1 dies
(``os._exit``, no clean-up, as after a hardware fault) in the middle of step 4; the launcher restarts the job and the
Trainer resumes from the last complete checkpoint."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import torch

from pipegoose_b200.distributed import ParallelContext, ParallelMode
from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import DataParallel, TensorParallel
from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
from pipegoose_b200.trainer import Callback, Trainer


def main():
    workdir, mode = sys.argv[1], sys.argv[2]
    crash = mode in ("crash", "hang")
    ctx = ParallelContext.from_torch(tensor_parallel_size=1, pipeline_parallel_size=1, data_parallel_size=2, backend="gloo")
    dp_rank = ctx.get_local_rank(ParallelMode.DATA)
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=64, hidden_size=32, n_layer=1, n_head=4))
    model = DataParallel(TensorParallel(model, ctx).parallelize(), ctx).parallelize()
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-2), ctx)
    g = torch.Generator().manual_seed(5 + dp_rank)
    data = [{"input_ids": torch.randint(0, 64, (2, 8), generator=g)} for _ in range(8)]
    marker = os.path.join(workdir, "crashed_once")

    started_at = []

    class Crash(Callback):
        def on_fit_start(self, trainer):
            started_at.append(trainer.state.step)          # 0, or the step of the checkpoint this launch resumed from

        def on_step_end(self, trainer, loss):
            if crash and dp_rank == 1 and trainer.state.step == 4 and not os.path.exists(marker):
                open(marker, "w").write("x")
                if mode == "hang":
                    import time

                    time.sleep(3600)              # a wedged rank: alive, silent — only the peers' watchdog can tell
                os._exit(17)                      # the other replica is left inside its next all-reduce

    trainer = Trainer(model, data, optim=optim, parallel_context=ctx, callbacks=[Crash()], log_every=1,
                      checkpoint_dir=os.path.join(workdir, "ckpt") if crash else None, checkpoint_every=3, resume=crash,
                      watchdog_timeout_s=8 if mode == "hang" else 30)
    state = trainer.fit()
    if ctx.get_global_rank() == 0:
        checksum = float(sum(p.detach().double().sum() for p in model.parameters()))
        with open(os.path.join(workdir, f"result_{mode}.json"), "w") as f:
            json.dump({"step": state.step, "loss": state.last_loss, "checksum": checksum,
                       "restarts": int(os.environ.get("TORCHELASTIC_RESTART_COUNT", "0")), "started_at": started_at[0]}, f)
    ctx.destroy()


if __name__ == "__main__":
    main()
