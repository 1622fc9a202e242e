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


def run_stall(rank, world_size, port):
    from pipegoose_b200.utils.watchdog import STALL_EXIT_CODE  # noqa: F401

    ctx = init_parallel_context(rank, world_size, port, 1, 1, world_size)
    wd = RankWatchdog(ctx, timeout_s=30.0, interval_s=0.1, stall_timeout_s=1.0, tag="stall").start()
    for _ in range(8):          # progress every 0.25 s for 2 s: twice the deadline, nothing happens
        time.sleep(0.25)
        wd.tick()
    assert wd.ticks == 8
    time.sleep(30)              # the main thread is wedged: the process must end with STALL_EXIT_CODE long before this returns
    raise AssertionError("the stalled process was not ended")


def test_stalled_main_thread_ends_the_process_with_the_restart_exit_code():
    from torch.multiprocessing import ProcessExitedException

    from pipegoose_b200.utils.watchdog import STALL_EXIT_CODE

    t0 = time.monotonic()
    with pytest.raises(ProcessExitedException) as info:
        spawn(run_stall, world_size=1)
    assert info.value.exit_code == STALL_EXIT_CODE and time.monotonic() - t0 < 25
