"""Job runtime of pipeline parallelism: packages, jobs + callbacks, queues, worker pool, stage-to-stage
transport, progress tracker and group handshake (parity with the reference's tests under
tests/nn/pipeline_parallel/{job,sync}/ and test_{worker,queue,comm,package}.py — real gloo processes,
nothing mocked)."""
import time
from queue import Queue

import pytest
import torch
import torch.distributed as dist
from torch import nn

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.nn.pipeline_parallel import queue as Q
from pipegoose_b200.nn.pipeline_parallel._comm import RECV_QUEUE, recv_package, send_package
from pipegoose_b200.nn.pipeline_parallel._job.backward import BackwardJob
from pipegoose_b200.nn.pipeline_parallel._job.callback import Callback, CallbackEvent
from pipegoose_b200.nn.pipeline_parallel._job.creator import create_job, schedule_backward_execution, schedule_backward_job
from pipegoose_b200.nn.pipeline_parallel._job.forward import ForwardJob
from pipegoose_b200.nn.pipeline_parallel._job.job import Job, JobStatus
from pipegoose_b200.nn.pipeline_parallel._job.job_type import JobType
from pipegoose_b200.nn.pipeline_parallel._job.register import add_job_to_queue
from pipegoose_b200.nn.pipeline_parallel._package import Metadata, Package, TrainingMetadata
from pipegoose_b200.nn.pipeline_parallel._worker import WorkerManager
from pipegoose_b200.nn.pipeline_parallel.exception import PipelineNoSavedActivationError, PipelineNoSavedInput
from pipegoose_b200.nn.pipeline_parallel.sync.callback import Callback as ClockCallback
from pipegoose_b200.nn.pipeline_parallel.sync.handshake import (
    ParallelGroupHandshake,
    ProgressTracker,
    get_progress_tracker,
)
from pipegoose_b200.testing.utils import init_parallel_context, spawn


def make_package(data, microbatch_idx=0, partition_idx=0, job_type=JobType.FORWARD, src=0, dst=0, training=True):
    return Package(data, Metadata(microbatch_idx, partition_idx, job_type, TrainingMetadata(training, training), src, dst))


# ------------------------------------------------------------------------------------------ package / job
def test_package_carries_data_and_metadata():
    x = torch.randn(2, 3)
    p = make_package(x, microbatch_idx=1, partition_idx=2, src=3, dst=4)
    assert p.data is x
    m = p.metadata
    assert (m.microbatch_idx, m.partition_idx, m.src, m.dst) == (1, 2, 3, 4)
    assert m.job_type is JobType.FORWARD and m.training.is_training and m.training.is_grad_enabled
    assert p.clone_metadata(partition_idx=5).partition_idx == 5 and m.partition_idx == 2


class _Doubler(Job):
    def run_compute(self):
        return self.function(self.input.data)


def test_job_lifecycle_and_callback_order():
    log = []

    class A(Callback):
        order = 2

        def after_create(self):
            log.append("A.create")

        def before_compute(self):
            log.append("A.before")

        def after_compute(self):
            log.append("A.after")

    class B(Callback):
        order = 1

        def after_compute(self):
            log.append("B.after")
            assert self.job.status is JobStatus.EXECUTED

    job = _Doubler(lambda t: t * 2, make_package(torch.ones(2)), cbs=[A, B()])
    assert job.status is JobStatus.PENDING and len(job.key) == 15 and log == ["A.create"]
    out = job.compute()
    assert torch.equal(out, torch.full((2,), 2.0)) and job.output is out
    assert log == ["A.create", "A.before", "B.after", "A.after"]
    job.remove_cb(A)
    assert all(not isinstance(c, A) for c in job.cbs) and len(job.cbs) == 1
    with pytest.raises(AssertionError):
        job._run_callback("after_compute")
    assert CallbackEvent.AFTER_COMPUTE.value == "after_compute"


def test_failing_job_is_marked_failed():
    seen = []

    class OnFail(Callback):
        def on_failure(self):
            seen.append(type(self.job.error))

    def boom(_):
        raise ValueError("x")

    job = _Doubler(boom, make_package(torch.ones(1)), cbs=[OnFail])
    with pytest.raises(ValueError):
        job.compute()
    assert job.status is JobStatus.FAILED and seen == [ValueError]


def test_register_and_queues():
    q = Queue()
    job = _Doubler(lambda t: t, make_package(torch.ones(1)))
    add_job_to_queue(job, q)
    assert q.qsize() == 1 and q.get() is job
    with pytest.raises(AssertionError):
        add_job_to_queue("not a job", q)
    Q.clear_all()
    x = torch.randn(3, requires_grad=False)
    Q.save_input_activations(x, 0, 1)
    assert Q.InputActivations.is_saved(0, 1) and Q.get_input_activations(0, 1).requires_grad
    y = torch.randn(3, requires_grad=True) * 2
    Q.save_output_activations(y, 0, 1)
    assert Q.SavedActivation.is_saved(0, 1)
    assert Q.get_output_activations(0, 1, is_pipeline=True) is y
    detached = Q.get_output_activations(0, 1)
    assert detached.requires_grad and detached.grad_fn is None
    with pytest.raises(PipelineNoSavedActivationError):
        Q.get_output_activations(7, 7)
    with pytest.raises(PipelineNoSavedInput):
        Q.get_input_activations(7, 7)
    Q.clear_all()


# ------------------------------------------------------------------------------------------ worker pool
def test_worker_manager_executes_and_grows():
    pending, selected = Queue(), Queue()
    done = []

    class SlowJob:
        def compute(self):
            time.sleep(0.05)
            done.append(1)

    class BadJob:
        def compute(self):
            raise RuntimeError("bad")

    wm = WorkerManager(num_workers=1, min_workers=1, max_workers=3, pending_jobs=pending, selected_jobs=selected)
    wm.spawn()
    assert wm.pending_jobs is pending and wm.selected_jobs is selected
    assert all(w.is_alive() and not w.is_running for w in wm.worker_pool)
    for _ in range(6):
        pending.put(SlowJob())
    pending.put(BadJob())
    assert wm.wait_idle(timeout=20)
    assert len(done) == 6
    assert 1 <= len(wm.worker_pool) <= 3
    assert len(wm.failed_jobs) == 1 and isinstance(wm.failed_jobs[0][1], RuntimeError)
    assert all(w.is_alive() for w in wm.worker_pool), "a failing job must not kill its worker"
    pool = list(wm.worker_pool)
    wm.destroy()
    assert all(not w.is_alive() for w in pool)


# ------------------------------------------------------------------------------------------ transport + jobs
def run_send_recv_package(rank, world_size, port, dtype):
    ctx = init_parallel_context(rank, world_size, port, 1, world_size, 1)
    if rank == 0:
        pkg = make_package(torch.arange(12, dtype=dtype).view(3, 4), microbatch_idx=2, partition_idx=1,
                           job_type=JobType.BACKWARD, src=0, dst=1)
        send_package(pkg, ctx)
    else:
        pkg = recv_package(0, ctx)
        assert RECV_QUEUE.get_nowait() is pkg
        assert torch.equal(pkg.data, torch.arange(12, dtype=dtype).view(3, 4)) and pkg.data.dtype == dtype
        m = pkg.metadata
        assert (m.microbatch_idx, m.partition_idx, m.job_type, m.src, m.dst) == (2, 1, JobType.BACKWARD, 0, 1)
        assert m.training.is_training and m.training.is_grad_enabled
    ctx.destroy()


@pytest.mark.parametrize("dtype", [torch.float32, torch.int64])
def test_send_recv_package(dtype):
    spawn(run_send_recv_package, world_size=2, dtype=dtype)


def run_pipeline_of_jobs(rank, world_size, port, state_dicts, batch, ref_grads, ref_input_grad, ref_loss, use_callback=False,
                         use_trigger=False):
    """Two stages, GPipe order, every step a job created by ``create_job`` from a package."""
    ctx = init_parallel_context(rank, world_size, port, 1, 2, 1)
    Q.clear_all()
    torch.manual_seed(0)
    stage = nn.Sequential(nn.Linear(8, 8), nn.Tanh())
    stage.load_state_dict(state_dicts[rank])
    n_mb = 2
    chunks = batch.chunk(n_mb)
    losses = []
    if rank == 0:
        for i in range(n_mb):
            job = create_job(stage, make_package(chunks[i].clone(), i, 0, src=0, dst=0), ctx)
            assert isinstance(job, ForwardJob)
            job.compute()
            assert job.status is JobStatus.DONE and job.output.metadata.partition_idx == 1 and job.output.metadata.dst == 1
        for i in reversed(range(n_mb)):
            pkg = recv_package(1, ctx)
            assert pkg.metadata.job_type is JobType.BACKWARD and pkg.metadata.partition_idx == 0
            job = create_job(stage, pkg, ctx)
            assert isinstance(job, BackwardJob)
            job.compute()
            assert job.output.data is not None  # gradient w.r.t. the pipeline input
            if i == 0:
                assert torch.allclose(job.output.data, ref_input_grad[:chunks[0].shape[0]], atol=1e-6)
    else:
        outs = []
        for i in range(n_mb):
            pkg = recv_package(0, ctx)
            assert pkg.metadata.microbatch_idx == i and pkg.metadata.partition_idx == 1
            job = create_job(stage, pkg, ctx, schedule_backward=use_callback)
            job.compute()
            outs.append(job.output)
        for i in reversed(range(n_mb)):
            # loss.backward() only records d loss / d output; the backward job replays it through the stage
            if use_trigger:    # event-driven: loss.backward() itself leaves the backward job in the pending queue
                assert Q.JobQueue.PENDING_JOBS.empty()
                y = schedule_backward_job(outs[i], parallel_context=ctx).data
                loss = y.pow(2).sum() / batch.shape[0]
                losses.append(loss.item())
                loss.backward()
                bjob = Q.JobQueue.PENDING_JOBS.get_nowait()
                assert isinstance(bjob, BackwardJob) and Q.JobQueue.PENDING_JOBS.empty()
                m = bjob.input.metadata
                assert (m.microbatch_idx, m.partition_idx, m.job_type) == (i, 1, JobType.BACKWARD)
                bjob.compute()
                continue
            if use_callback:   # ScheduleBackwardJobCallback already swapped the output for the recording wrapper
                y = outs[i].data
                assert y is Q._SAVED_SCHEDULED_ACTIVATIONS[(i, 1)] and y.requires_grad
            else:
                y = schedule_backward_execution(outs[i]).data
            loss = y.pow(2).sum() / batch.shape[0]
            losses.append(loss.item())
            loss.backward()
            grad = Q.get_grad_loss(i, 1)
            bjob = create_job(stage, make_package(grad, i, 1, JobType.BACKWARD, src=1, dst=1), ctx)
            bjob.compute()
        assert abs(sum(losses) - ref_loss) < 1e-5
    for name, p in stage.named_parameters():
        assert torch.allclose(p.grad, ref_grads[rank][name], atol=1e-5), (rank, name)
    ctx.destroy()


@pytest.mark.parametrize("use_callback", [False, True, "trigger"])
def test_forward_backward_jobs_match_sequential_execution(use_callback):
    torch.manual_seed(1)
    stages = [nn.Sequential(nn.Linear(8, 8), nn.Tanh()) for _ in range(2)]
    batch = torch.randn(6, 8)
    x = batch.clone().requires_grad_(True)
    loss = stages[1](stages[0](x)).pow(2).sum() / batch.shape[0]
    loss.backward()
    ref_grads = [{n: p.grad.clone() for n, p in s.named_parameters()} for s in stages]
    spawn(run_pipeline_of_jobs, world_size=2, state_dicts=[s.state_dict() for s in stages], batch=batch,
          ref_grads=ref_grads, ref_input_grad=x.grad.clone(), ref_loss=loss.item(), use_callback=use_callback is True,
          use_trigger=use_callback == "trigger")


# ------------------------------------------------------------------------------------------ sync
def run_progress_tracker(rank, world_size, port, tp, pp, dp):
    ctx = init_parallel_context(rank, world_size, port, tp, pp, dp)
    n_clocks = world_size
    progress = {c: {r: False for r in range(world_size)} for c in range(n_clocks)}
    fired = []

    class OnClock(ClockCallback):
        def after_new_clock_cycle(self, progress, clock_idx):
            fired.append(clock_idx)

    tracker = ProgressTracker(0, callbacks=[OnClock()], parallel_context=ctx, parallel_mode=ParallelMode.GLOBAL)
    assert get_progress_tracker() is tracker
    if rank == tracker.master_rank:
        tracker.initiate(progress)
    dist.barrier()
    assert tracker.is_initiated() and tracker.clock_idx == 0
    assert tracker.progress == progress and not tracker.is_all_confirmed(clock_idx=0)
    dist.barrier()  # nobody confirms before everyone has looked at the untouched table
    for clock in range(n_clocks):
        tracker.confirm(rank)
        assert tracker.is_confirmed(rank, clock_idx=clock)
        tracker.wait_for_clock(clock)  # blocks on the store, no polling
        assert tracker.is_all_confirmed(clock_idx=clock)
        if clock + 1 < n_clocks:
            assert not tracker.is_all_confirmed(clock_idx=clock + 1)
        assert tracker.clock_idx == clock + 1
        dist.barrier()
    assert tracker.progress == {c: {r: True for r in range(world_size)} for c in range(n_clocks)}
    assert fired == list(range(1, n_clocks + 1))
    # a second schedule (the reference re-initiates the tracker for the backward pass)
    if rank == tracker.master_rank:
        tracker.initiate({0: {r: False for r in range(world_size)}})
    dist.barrier()
    assert tracker.is_initiated() and tracker.clock_idx == 0
    dist.barrier()
    tracker.confirm(rank)
    tracker.wait_for_clock(0)
    assert tracker.clock_idx == 1
    ctx.destroy()


@pytest.mark.parametrize("tp,pp,dp", [(1, 2, 1), (2, 2, 1)])
def test_progress_tracker(tp, pp, dp):
    spawn(run_progress_tracker, world_size=tp * pp * dp, tp=tp, pp=pp, dp=dp)


def run_handshake(rank, world_size, port, tp, pp, dp):
    ctx = init_parallel_context(rank, world_size, port, tp, pp, dp)
    hs = ParallelGroupHandshake(ctx, ParallelMode.TENSOR, master_rank=0)
    hs.initiate()
    dist.barrier()
    assert hs.is_initiated() and not hs.is_confirmed()
    hs.confirm()
    assert hs.is_confirmed()
    hs.barrier()
    # after the barrier every member of this TENSOR group had confirmed
    for r in range(ctx.get_world_size(ParallelMode.TENSOR)):
        assert hs._store.check([f"e0/r{r}"])
    hs.barrier()  # a second epoch works without re-construction
    ctx.destroy()


def test_parallel_group_handshake():
    spawn(run_handshake, world_size=4, tp=2, pp=1, dp=2)


# ------------------------------------------------------------------------------------------ engine made of jobs
def run_job_engine(rank, world_size, port, pp, state, ids, ref_loss, ref_grads):
    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
    from pipegoose_b200.nn import PipelineParallel

    ctx = init_parallel_context(rank, world_size, port, 1, pp, 1)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4))
    model.load_state_dict(state)
    names = {id(p): n for n, p in model.named_parameters()}
    model = PipelineParallel(model, num_microbatches=4, parallel_context=ctx, runtime="jobs").parallelize()
    engine = model._pg_pipeline_engine
    for _ in range(2):  # a second step re-uses the workers and starts new tracker rounds
        out = model(ids, labels=ids)
        assert torch.allclose(out.loss, ref_loss, atol=1e-5), (out.loss, ref_loss)  # broadcast from the last stage
        out.loss.backward()   # installs the gradients the jobs computed (they survive a zero_grad() in between)
        for p in model._pg_pipeline_stage.parameters():
            assert torch.allclose(p.grad, ref_grads[names[id(p)]], atol=2e-5), names[id(p)]
    # the tracker saw every task of the last (backward) schedule (earlier stages may still be finishing theirs)
    engine.tracker.wait_for_clock(len(engine.tracker.progress) - 1)
    assert all(all(done.values()) for done in engine.tracker.progress.values())
    assert not engine.worker_manager.failed_jobs
    engine.destroy()
    ctx.destroy()


@pytest.mark.parametrize("pp", [2, 4])
def test_job_runtime_engine_matches_sequential_training_step(pp):
    import copy

    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM

    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4))
    ids = torch.randint(0, 96, (8, 8))
    loss = torch.stack([model(c, labels=c).loss for c in ids.chunk(4)]).mean()
    loss.backward()
    spawn(run_job_engine, world_size=pp, pp=pp, state=copy.deepcopy(model.state_dict()), ids=ids, ref_loss=loss.detach(),
          ref_grads={n: p.grad.clone() for n, p in model.named_parameters()})


def run_job_engine_uneven(rank, world_size, port, state, ids, labels, ref_loss, ref_grads):
    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
    from pipegoose_b200.nn import PipelineParallel

    ctx = init_parallel_context(rank, world_size, port, 1, 3, 1)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=6, n_head=4))
    model.load_state_dict(state)
    names = {id(p): n for n, p in model.named_parameters()}
    model = PipelineParallel(model, num_microbatches=2, parallel_context=ctx, runtime="jobs").parallelize()
    out = model(ids, labels=labels)
    assert torch.allclose(out.loss, ref_loss, atol=1e-5), (out.loss, ref_loss)
    out.loss.backward()
    for p in model._pg_pipeline_stage.parameters():
        assert torch.allclose(p.grad, ref_grads[names[id(p)]], atol=2e-5), names[id(p)]
    model._pg_pipeline_engine.destroy()
    ctx.destroy()


def test_job_runtime_with_uneven_micro_batches_and_ignored_labels():
    """5 sequences in 2 micro-batches over 3 stages (a middle stage sees both shapes), one sequence partly ignored."""
    import copy

    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM

    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=6, n_head=4))
    ids = torch.randint(0, 96, (5, 8))
    labels = ids.clone()
    labels[4, 4:] = -100
    loss = model(ids, labels=labels).loss
    loss.backward()
    spawn(run_job_engine_uneven, world_size=3, state=copy.deepcopy(model.state_dict()), ids=ids, labels=labels,
          ref_loss=loss.detach(), ref_grads={n: p.grad.clone() for n, p in model.named_parameters()})
