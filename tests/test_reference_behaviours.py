"""One small test per behaviour the reference pins with its own one-behaviour tests, for the spots where this suite
only had a combined test (see docs/TESTING.md, "Reference test file -> where the same behaviour is pinned here"):
reference tests/core/bucket/test_bucket.py, tests/nn/pipeline_parallel/job/test_callback.py, test_queue.py,
test_scheduler.py.  Single process, no spawn."""
from queue import Queue

import pytest
import torch

from pipegoose_b200.core.bucket.bucket import Bucket
from pipegoose_b200.core.bucket.exception import BucketClosedError, BucketFullError
from pipegoose_b200.nn.pipeline_parallel import queue as Q
from pipegoose_b200.nn.pipeline_parallel._job.callback import Callback, CallbackEvent
from pipegoose_b200.nn.pipeline_parallel._job.job import Job, JobStatus
from pipegoose_b200.nn.pipeline_parallel._job.job_type import JobType
from pipegoose_b200.nn.pipeline_parallel._job.register import add_job_to_queue
from pipegoose_b200.nn.pipeline_parallel._package import Metadata, Package, TrainingMetadata
from pipegoose_b200.nn.pipeline_parallel.scheduler import GPipeScheduler

SIZE = 512


# ---------------------------------------------------------------------------------------------------- bucket
def test_bucket_new_tensor_lives_in_the_bucket_not_in_its_old_storage():
    t = torch.randn(3, 5)
    before = t.clone()
    old = t.untyped_storage().data_ptr()
    bucket = Bucket(SIZE, torch.float32)
    assert (bucket.size, bucket.dtype, bucket.available_size, len(bucket), bucket.is_full) == (SIZE, torch.float32, SIZE, 0, False)
    out = bucket.add_tensor(t)
    assert isinstance(out, torch.Tensor) and torch.equal(out, before)
    assert bucket.available_size == SIZE - 15 and len(bucket) == 1
    assert out.untyped_storage().data_ptr() == bucket.storage().data_ptr() != old


def test_bucket_rejects_a_tensor_larger_than_itself():
    bucket = Bucket(SIZE, torch.float32)
    with pytest.raises(BucketFullError):
        bucket.add_tensor(torch.randn(2, SIZE))
    assert len(bucket) == 0 and bucket.is_free


def test_bucket_rejects_a_tensor_larger_than_the_space_left():
    bucket = Bucket(SIZE, torch.float32)
    bucket.add_tensor(torch.randn(SIZE - 1))
    with pytest.raises(BucketFullError):
        bucket.add_tensor(torch.randn(2))
    bucket.add_tensor(torch.randn(1))   # exactly the space left still fits
    assert bucket.is_full and bucket.available_size == 0


def test_bucket_closed():
    bucket = Bucket(SIZE, torch.float32)
    assert bucket.is_closed is False
    bucket.close()
    with pytest.raises(BucketClosedError):
        bucket.add_tensor(torch.randn(4))
    assert bucket.is_closed is True


def test_bucket_rejects_another_dtype():
    bucket = Bucket(SIZE, torch.float32)
    with pytest.raises(AssertionError):
        bucket.add_tensor(torch.randn(4).half())
    with pytest.raises(AssertionError):
        bucket.add_tensor([1.0, 2.0])


def test_bucket_clear_empties_and_reopens():
    bucket = Bucket(SIZE, torch.float32)
    bucket.add_tensor(torch.randn(10)), bucket.add_tensor(torch.randn(20))
    bucket.close()
    bucket.clear()
    assert bucket.available_size == SIZE and len(bucket) == 0 and not bucket.is_closed
    assert torch.count_nonzero(bucket._flat) == 0
    with pytest.raises(AssertionError):
        bucket.clear()   # nothing left to clear


# ---------------------------------------------------------------------------------------------------- job callbacks
class _Square(Job):
    def run_compute(self):
        return self.function(self.input.data)


def _package(x):
    return Package(x, Metadata(0, 0, JobType.FORWARD, TrainingMetadata(True, True), 0, 0))


def test_job_callbacks_see_the_job():
    seen = {}

    class Peek(Callback):
        def after_create(self):
            seen["key"], seen["status"] = self.job.key, self.job.status

        def after_compute(self):
            seen["output"] = self.job.output

    job = _Square(lambda t: t * t, _package(torch.tensor([3.0])), cbs=[Peek])
    assert seen == {"key": job.key, "status": JobStatus.PENDING}
    job.compute()
    assert seen["output"] is job.output and float(seen["output"]) == 9.0


def test_job_callbacks_run_in_order_whatever_the_registration_order():
    log = []

    def cb(name, order_):
        class C(Callback):
            order = order_

            def before_compute(self):
                log.append(name)
        return C

    job = _Square(lambda t: t, _package(torch.ones(1)), cbs=[cb("third", 3), cb("first", 1), cb("second", 2)])
    job.compute()
    assert log == ["first", "second", "third"]


def test_job_callbacks_add_and_remove_one():
    class Mark(Callback):
        def after_compute(self):
            self.job.marked = True

    job = _Square(lambda t: t, _package(torch.ones(1)))
    n = len(job.cbs)
    job.add_cb(Mark)                       # a class is instantiated, an instance is taken as is
    assert len(job.cbs) == n + 1 and isinstance(job.cbs[-1], Mark) and job.cbs[-1].job is job
    job.remove_cb(Mark)
    assert len(job.cbs) == n
    inst = Mark()
    job.add_cb(inst)
    assert job.cbs[-1] is inst
    job.remove_cb(inst)
    assert inst not in job.cbs
    with pytest.raises(AssertionError):
        job.add_cb(object())


def test_job_callbacks_add_and_remove_a_list():
    class A(Callback):
        pass

    class B(Callback):
        pass

    job = _Square(lambda t: t, _package(torch.ones(1)))
    n = len(job.cbs)
    job.add_cbs([A, B()])
    assert len(job.cbs) == n + 2
    job.remove_cbs([A, B])
    assert len(job.cbs) == n
    assert {e.value for e in CallbackEvent} >= {"after_create", "before_compute", "after_compute"}


def test_job_goes_into_a_queue_once():
    q = Queue()
    job = _Square(lambda t: t, _package(torch.ones(1)))
    add_job_to_queue(job, q)
    assert q.qsize() == 1 and q.get_nowait() is job and q.empty()


# ---------------------------------------------------------------------------------------------------- activation stores
def test_activation_queues():
    Q.clear_all()
    try:
        a, b = torch.randn(4), torch.randn(4, requires_grad=True) * 3
        # inputs: kept (a stage may read them again), returned as a leaf that can take a gradient
        Q.save_input_activations(a, microbatch_idx=2, partition_idx=1)
        assert Q.InputActivations.is_saved(2, 1) and not Q.InputActivations.is_saved(1, 2)
        got = Q.get_input_activations(2, 1)
        assert got.requires_grad and torch.equal(got, a) and Q.InputActivations.is_saved(2, 1)
        # outputs: with is_pipeline the graph is kept, otherwise a detached leaf
        Q.save_output_activations(b, microbatch_idx=2, partition_idx=1)
        assert Q.get_output_activations(2, 1, is_pipeline=True) is b
        leaf = Q.get_output_activations(2, 1)
        assert leaf.grad_fn is None and leaf.requires_grad and torch.equal(leaf, b)
        # SavedActivation.get_saved_activations consumes the entry (the backward job takes it exactly once)
        key = Q.SavedActivation.get_key(2, 1)
        assert Q.SavedActivation.get_saved_activations(key) is b and not Q.SavedActivation.is_saved(2, 1)
        Q.SavedActivation.save_activations(key, b, is_by_schedule=True)   # goes to the store of the backward triggers
        assert not Q.SavedActivation.is_saved(2, 1) and Q._SAVED_SCHEDULED_ACTIVATIONS[key] is b
        # the gradient of the loss of a micro-batch is consumed by its backward job
        g = torch.ones(4)
        Q.save_grad_loss(g, 0, 3)
        assert Q.get_grad_loss(0, 3) is g
        with pytest.raises(KeyError):
            Q.get_grad_loss(0, 3)
    finally:
        Q.clear_all()


# ---------------------------------------------------------------------------------------------------- GPipe clocks
@pytest.mark.parametrize("m,n", [(4, 3), (5, 2), (1, 4)])
def test_gpipe_forward_clocks(m, n):
    fwd = GPipeScheduler(m, n).get_forward_schedules()
    assert len(fwd) == m + n - 1
    for c, clock in enumerate(fwd):
        assert [(t.microbatch_idx, t.partition_idx) for t in clock] == [(c - p, p) for p in range(n) if 0 <= c - p < m]
        assert all(t.job_type is JobType.FORWARD for t in clock)
    # at most one task per partition per clock, every (micro-batch, partition) exactly once
    assert sorted((t.microbatch_idx, t.partition_idx) for c in fwd for t in c) == [(i, p) for i in range(m) for p in range(n)]


@pytest.mark.parametrize("m,n", [(4, 3), (5, 2), (1, 4)])
def test_gpipe_backward_clocks_mirror_the_forward_ones(m, n):
    sch = GPipeScheduler(m, n)
    fwd, bwd = sch.get_forward_schedules(), sch.get_backward_schedules()
    assert len(bwd) == len(fwd) == sch.total_backward_clock_cycles
    for f, b in zip(reversed(fwd), bwd):
        assert [(t.microbatch_idx, t.partition_idx) for t in f] == [(t.microbatch_idx, t.partition_idx) for t in b]
        assert all(t.job_type is JobType.BACKWARD for t in b)
    # the first backward clock is the last partition on the last micro-batch; the last one is partition 0, micro-batch 0
    assert (bwd[0][0].microbatch_idx, bwd[0][0].partition_idx) == (m - 1, n - 1)
    assert (bwd[-1][0].microbatch_idx, bwd[-1][0].partition_idx) == (0, 0)


def test_gpipe_full_schedule_is_forward_then_backward():
    sch = GPipeScheduler(3, 2)
    clocks = sch.get_schedules()
    assert clocks == sch.get_forward_schedules() + sch.get_backward_schedules()
    assert sch.total_clock_cycles == len(clocks) == 2 * (3 + 2 - 1)
    order = sch.get_stage_order(1)
    assert [(t.job_type, t.microbatch_idx) for t in order] == [(JobType.FORWARD, 0), (JobType.FORWARD, 1), (JobType.FORWARD, 2),
                                                               (JobType.BACKWARD, 2), (JobType.BACKWARD, 1), (JobType.BACKWARD, 0)]


# ---------------------------------------------------------------------------------------------------- small parity names
def test_schedule_record_and_module_constants():
    from pipegoose_b200.nn.pipeline_parallel.partitioner import INPUT_NAMES
    from pipegoose_b200.nn.pipeline_parallel.pipeline_engine import PipelineEngine, Schedule
    from pipegoose_b200.nn.pipeline_parallel.sync.handshake import Handshake
    from pipegoose_b200.nn.pipeline_parallel.task import Task
    from pipegoose_b200.testing.utils import N_MICROBATCHES, N_PARTITIONS

    s = Schedule(JobType.BACKWARD, 2, 5)    # the reference's field order: job type, partition, micro-batch
    assert (s.partition_idx, s.microbatch_idx) == (2, 5)
    t = s.to_task()
    assert t == Task(JobType.BACKWARD, 5, 2) and Schedule.from_task(t) == s
    assert INPUT_NAMES == ["input_ids", "attention_mask"] and PipelineEngine.MASTER_RANK == 0
    assert (N_PARTITIONS, N_MICROBATCHES) == (3, 5)
    assert Handshake.master_rank is None and Handshake.parallel_context is None


# ---------------------------------------------------------------------------------------------------- found by running the
# reference's own tests against this package (tools/run_reference_tests.py)
def test_expert_loss_accepts_plain_numbers():
    import torch.nn.functional as F

    from pipegoose_b200.nn.expert_parallel import ExpertLoss
    from pipegoose_b200.nn.expert_parallel.expert_context import ExpertContext

    store = ExpertContext.get_instance()
    store.pop_all_aux_loss(), store.pop_all_z_loss()
    logits, target = torch.randn(6, 3), torch.randn(6, 3)
    loss_fn = ExpertLoss(torch.nn.MSELoss(), aux_weight=0.1, z_weight=0.2)
    store.push_aux_loss(1.5), store.push_z_loss(2.5)                       # floats, as the reference's test pushes them
    store.push_aux_loss(torch.tensor(0.5)), store.push_z_loss(torch.tensor(1.0))
    assert loss_fn.aux_loss[0] == 1.5 and len(loss_fn.z_loss) == 2
    got = loss_fn(logits, target)
    assert torch.allclose(got, F.mse_loss(logits, target) + 0.1 * 2.0 + 0.2 * 3.5)
    assert store.aux_loss == [] and store.z_loss == []


def test_save_grad_loss_returns_the_package():
    from pipegoose_b200.nn.pipeline_parallel._job.backward import save_grad_loss

    Q.clear_all()
    pkg = _package(torch.randn(3, 2))
    pkg.metadata.microbatch_idx, pkg.metadata.partition_idx = 1, 2
    out = save_grad_loss(pkg)
    assert out is pkg and out.data.requires_grad
    out.data.pow(2).sum().backward()
    assert isinstance(Q.get_grad_loss(1, 2), torch.Tensor)
    Q.clear_all()


def _run_oversized_and_grad_inputs(rank, world_size, port):
    import torch.distributed as dist

    from pipegoose_b200.core.bucket.dist import BucketDistributor
    from pipegoose_b200.core.bucket.utils import mb_size_to_num_elements
    from pipegoose_b200.distributed.functional import all_gather
    from pipegoose_b200.distributed.parallel_mode import ParallelMode
    from pipegoose_b200.testing.utils import init_parallel_context

    ctx = init_parallel_context(rank, world_size, port, 1, 1, world_size)
    n = mb_size_to_num_elements(0.001, torch.float32)
    big = torch.arange(2 * n, dtype=torch.float32)
    BucketDistributor(dist.all_reduce, 0.001, ctx).execute(big, ParallelMode.DATA)
    # larger than a bucket: reduced on its own and COMPLETE when execute() returns (no flush)
    assert torch.equal(big, torch.arange(2 * n, dtype=torch.float32) * world_size)
    leaf = torch.tensor(float(rank), requires_grad=True)      # a 0-d tensor that requires grad (reference test_functional)
    assert all_gather(leaf, dim=0, parallel_context=ctx, parallel_mode=ParallelMode.DATA).tolist() == [0.0, 1.0]
    ctx.destroy()


def test_oversized_tensors_complete_in_execute_and_all_gather_takes_grad_leaves():
    from pipegoose_b200.testing.utils import spawn

    spawn(_run_oversized_and_grad_inputs, world_size=2)


def _run_shared_model(rank, world_size, port, model, ref, x):
    from pipegoose_b200.optim import DistributedOptimizer
    from pipegoose_b200.testing.utils import init_parallel_context

    ctx = init_parallel_context(rank, world_size, port, 1, 1, world_size)
    opt = DistributedOptimizer(torch.optim.Adam(model.parameters()), ctx)
    opt.zero_grad()
    model(x).sum().backward()
    opt.step()
    for p, q in zip(model.parameters(), ref.parameters()):
        assert torch.allclose(p, q), "a rank wrote shared parameters while another one was still reading them"
    ctx.destroy()


def test_zero1_on_a_model_in_shared_memory():
    """A module passed to the ranks as a spawn argument lives in shared memory: every rank works on the SAME parameter
    storage (the reference's tests/optim/zero/test_optim.py).  Owners must not step before everybody finished backward."""
    import copy

    from pipegoose_b200.testing.utils import spawn

    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.ReLU(), torch.nn.Linear(512, 64))
    x = torch.randn(128, 256)
    ref = copy.deepcopy(model)
    o = torch.optim.Adam(ref.parameters())
    ref(x).sum().backward()
    o.step()
    spawn(_run_shared_model, world_size=4, model=model, ref=ref, x=x)
