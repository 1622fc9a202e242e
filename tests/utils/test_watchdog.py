"""Failure detection helpers (utils/watchdog.py): a peer that stops heart-beating is reported, a clean exit is not,
the default handler turns a hang into RankFailure, step_deadline dumps stacks on overrun."""
import tempfile
import time

import pytest
import torch.distributed as dist

from pipegoose_b200.testing.utils import init_parallel_context, spawn
from pipegoose_b200.utils.watchdog import RankFailure, RankWatchdog, step_deadline


def run_watchdog(rank, world_size, port, scenario):
    ctx = init_parallel_context(rank, world_size, port, 1, 1, world_size)
    seen = []
    handler = None if scenario == "interrupt" else seen.append
    wd = RankWatchdog(ctx, timeout_s=1.0, interval_s=0.1, on_failure=handler, tag=f"t-{scenario}").start()
    dist.barrier()
    if rank == 1:
        # "silent": the heartbeat thread dies without saying goodbye (a crashed / wedged rank);  "clean": orderly exit
        wd.stop(announce=(scenario == "clean"))
        time.sleep(2.5)
    else:
        if scenario == "interrupt":
            with pytest.raises(RankFailure) as info:
                with wd.translate():
                    time.sleep(30)   # stands for a collective that would block forever
            assert info.value.dead_ranks == [1]
        else:
            deadline = time.monotonic() + 3.0
            while time.monotonic() < deadline and not wd.failed:
                time.sleep(0.05)
            if scenario == "silent":
                assert wd.failed == [1] and seen == [[1]]
                with pytest.raises(RankFailure):
                    wd.check()
            else:
                assert wd.failed == [] and seen == []
                wd.check()
        wd.stop()
    dist.barrier()
    ctx.destroy()


@pytest.mark.parametrize("scenario", ["silent", "clean", "interrupt"])
def test_rank_watchdog(scenario):
    spawn(run_watchdog, world_size=2, scenario=scenario)


def test_step_deadline_dumps_stacks_on_overrun():
    with tempfile.TemporaryFile(mode="w+") as f:
        with step_deadline(0.2, file=f):
            time.sleep(0.6)
        f.seek(0)
        assert "test_step_deadline_dumps_stacks_on_overrun" in f.read()
    with tempfile.TemporaryFile(mode="w+") as f:
        with step_deadline(5.0, file=f):
            pass
        time.sleep(0.05)
        f.seek(0)
        assert f.read() == ""
