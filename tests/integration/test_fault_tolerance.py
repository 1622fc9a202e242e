"""Failure recovery end to end (the reference has neither): a replica dies mid-run without any
clean-up, ``torchrun --max-restarts`` relaunches the job, ``Trainer(resume=True)`` continues from the last complete
checkpoint, and the final parameters equal those of a run that never crashed.

The relaunch only works because ``ParallelContext`` gives every restart attempt its own key space in the launcher's
store (distributed/parallel_context.py::init_global_dist): stock ``init_process_group("gloo")`` + ``new_group`` reads the
dead ranks' addresses of the previous attempt and fails with "connectFullMesh ... Connection refused"."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _launch(workdir, mode, max_restarts):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1", "--nnodes=1",
           "--nproc-per-node=2", f"--max-restarts={max_restarts}", "--monitor-interval", "0.5",
           os.path.join(HERE, "crashy_job.py"), workdir, mode]
    return _run_in_own_group(cmd, timeout=420)


def _run_in_own_group(cmd, timeout):
    """``subprocess.run`` that, on a timeout, ends the launcher AND its ranks (a killed torchrun leaves its workers behind):
    the job gets its own process group, which is signalled as a whole."""
    import signal

    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        proc.communicate()
        raise
    return subprocess.CompletedProcess(cmd, proc.returncode, out, err)


@pytest.mark.timeout(900)
def test_job_survives_a_dead_replica(tmp_path):
    workdir = str(tmp_path)
    clean = _launch(workdir, "clean", 0)
    assert clean.returncode == 0, clean.stderr[-3000:]
    crashed = _launch(workdir, "crash", 1)
    assert crashed.returncode == 0, crashed.stderr[-3000:]
    assert os.path.exists(os.path.join(workdir, "crashed_once"))
    want = json.load(open(os.path.join(workdir, "result_clean.json")))
    got = json.load(open(os.path.join(workdir, "result_crash.json")))
    assert got["restarts"] == 1 and want["restarts"] == 0
    assert got["step"] == want["step"] == 8
    assert want["started_at"] == 0 and got["started_at"] == 3               # checkpoints at 3 and 6; the crash came in 4
    assert got["loss"] == pytest.approx(want["loss"], abs=1e-5)
    assert got["checksum"] == pytest.approx(want["checksum"], abs=1e-4)


@pytest.mark.timeout(900)
def test_job_survives_a_hung_replica(tmp_path):
    """The replica does not die, it stops making progress (it sleeps inside a callback; its heartbeat THREAD keeps beating,
    so no peer would ever declare it dead, and its peer waits in the next all-reduce).  Both processes' progress
    deadlines (``RankWatchdog(stall_timeout_s=...)``, ticked by the Trainer) expire, the processes end with
    STALL_EXIT_CODE, the launcher starts the second attempt, which resumes from the checkpoint of step 3."""
    workdir = str(tmp_path)
    clean = _launch(workdir, "clean", 0)
    assert clean.returncode == 0, clean.stderr[-3000:]
    hung = _launch(workdir, "hang", 1)
    assert hung.returncode == 0, hung.stderr[-3000:]
    assert "no progress for" in hung.stderr
    want = json.load(open(os.path.join(workdir, "result_clean.json")))
    got = json.load(open(os.path.join(workdir, "result_hang.json")))
    assert got["restarts"] == 1 and got["started_at"] == 3 and got["step"] == 8
    assert got["checksum"] == pytest.approx(want["checksum"], abs=1e-4)


@pytest.mark.timeout(600)
def test_two_launcher_agents_like_two_nodes(tmp_path):
    """``torchrun --nnodes 2`` (two agents, one rank each: LOCAL_RANK 0 on both, LOCAL_WORLD_SIZE 1 < WORLD_SIZE 2), the
    shape of a multi-node job, against the single-agent run of the same job."""
    from pipegoose_b200.testing.utils import find_free_port

    one, two = str(tmp_path / "one"), str(tmp_path / "two")
    os.makedirs(one), os.makedirs(two)
    assert _launch(one, "clean", 0).returncode == 0
    port = find_free_port()
    procs = [subprocess.Popen([sys.executable, "-m", "torch.distributed.run", "--nnodes=2", f"--node-rank={n}",
                               "--nproc-per-node=1", "--rdzv-backend", "c10d", "--rdzv-endpoint", f"127.0.0.1:{port}",
                               "--rdzv-id", "two-agents", "--local-addr", "127.0.0.1",
                               os.path.join(HERE, "crashy_job.py"), two, "clean"],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
             for n in (0, 1)]
    try:
        outs = [p.communicate(timeout=400) for p in procs]
    except subprocess.TimeoutExpired:
        import signal

        for p in procs:
            try:
                os.killpg(p.pid, signal.SIGKILL)
            except ProcessLookupError:
                pass
        raise
    assert [p.returncode for p in procs] == [0, 0], outs[0][1][-2000:] + outs[1][1][-2000:]
    want = json.load(open(os.path.join(one, "result_clean.json")))
    got = json.load(open(os.path.join(two, "result_clean.json")))
    assert got["step"] == 8 and got["checksum"] == pytest.approx(want["checksum"], abs=1e-6)
