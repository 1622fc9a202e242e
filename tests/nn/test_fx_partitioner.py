"""Graph partitioning of models the structural partitioner does not know (reference: partitioner.py:146-244 traces and
cuts any model; here ``torch.fx`` + single-activation cuts): stage chains reproduce the model, stages are balanced, values
derived from the inputs are re-computed per stage, and the pipeline engine trains such a model like sequential code."""
import copy

import pytest
import torch
from torch import nn

from pipegoose_b200.nn import PipelineParallel
from pipegoose_b200.nn.pipeline_parallel.fx_partitioner import GraphPartitioner, GraphStage, NoLegalCut, _minmax_cuts
from pipegoose_b200.nn.pipeline_parallel.partitioner import UniformPartitioner
from pipegoose_b200.nn.pipeline_parallel.scheduler import SchedulerType
from pipegoose_b200.testing.utils import init_parallel_context, spawn


class _Block(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.ln, self.fc1, self.fc2 = nn.LayerNorm(d), nn.Linear(d, 4 * d), nn.Linear(4 * d, d)

    def forward(self, x, mask):
        return x + self.fc2(torch.relu(self.fc1(self.ln(x)))) * mask.unsqueeze(-1)


class _TokenModel(nn.Module):
    """Not a family the structural partitioner knows: no ``transformer`` attribute, custom argument names."""

    def __init__(self, n_blocks=6, d=16, vocab=50):
        super().__init__()
        self.emb = nn.Embedding(vocab, d)
        self.blocks = nn.ModuleList([_Block(d) for _ in range(n_blocks)])
        self.head = nn.Linear(d, vocab)
        self.register_buffer("scale", torch.tensor(2.0))
        self.vocab = vocab

    def forward(self, tokens, mask, labels=None):
        m = mask.float() * self.scale          # input-derived: every stage re-computes it
        x = self.emb(tokens)
        for b in self.blocks:
            x = b(x, m)
        logits = self.head(x)
        if labels is not None:
            return nn.functional.cross_entropy(logits.view(-1, self.vocab), labels.view(-1))
        return logits


def _chain(stages, first, **inputs):
    x = first
    for st in stages:
        x = st(x, **inputs)
    return x


@pytest.mark.parametrize("n", [2, 3, 4])
def test_stage_chain_reproduces_the_model_and_is_balanced(n):
    torch.manual_seed(0)
    net = _TokenModel()
    ids, mask = torch.randint(0, 50, (2, 5)), torch.ones(2, 5)
    mask[1, 3:] = 0
    stages = GraphPartitioner(net, None, n_partitions=n).split()
    assert len(stages) == n and all(isinstance(s, GraphStage) for s in stages)
    assert torch.allclose(_chain(stages, ids, mask=mask, labels=ids), net(ids, mask, ids))
    # every parameter lives in exactly one stage; the embedding does not count towards the balance
    owned = [id(p) for s in stages for p in s.parameters()]
    assert sorted(owned) == sorted(id(p) for p in net.parameters())
    sizes = [sum(p.numel() for p in s.parameters() if p is not net.emb.weight) for s in stages]
    block = sum(p.numel() for p in net.blocks[0].parameters())
    assert max(sizes) - min(sizes) <= block
    # the mask is read by every stage from the micro-batch, only the last stage reads the labels
    assert all("mask" in s.stage_inputs for s in stages)
    assert ["labels" in s.stage_inputs for s in stages] == [False] * (n - 1) + [True]
    assert stages[0].is_first and stages[-1].is_last and stages[0].first_input_name == "tokens"


def test_concrete_args_select_the_traced_branch():
    torch.manual_seed(0)
    net = _TokenModel(n_blocks=4)
    ids, mask = torch.randint(0, 50, (3, 4)), torch.ones(3, 4)
    stages = GraphPartitioner(net, None, n_partitions=2, concrete_args={"labels": None}).split()
    assert torch.allclose(_chain(stages, ids, mask=mask), net(ids, mask))


def test_convolutional_model_without_named_inputs():
    torch.manual_seed(0)

    class Wrapped(nn.Module):
        def __init__(self):
            super().__init__()
            self.c1, self.c2, self.c3 = nn.Conv2d(3, 8, 3, padding=1), nn.Conv2d(8, 8, 3, padding=1), nn.Conv2d(8, 4, 3, padding=1)
            self.fc = nn.Linear(4 * 6 * 6, 10)

        def forward(self, img):
            y = torch.relu(self.c1(img))
            y = torch.relu(self.c2(y)) + y
            y = torch.relu(self.c3(y))
            return self.fc(y.flatten(1))

    net = Wrapped()
    img = torch.randn(2, 3, 6, 6)
    for n in (2, 3):
        stages = GraphPartitioner(net, None, n_partitions=n).split()
        assert torch.allclose(_chain(stages, img), net(img), atol=1e-6)
        assert all(s.stage_inputs == () for s in stages)


def test_long_skip_connection_limits_the_cuts():
    class Skip(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b, self.c = nn.Linear(4, 4), nn.Linear(4, 4), nn.Linear(4, 4)

        def forward(self, x):
            h = self.a(x)
            return self.c(self.b(h)) + h     # h is live until the very end: nothing between a and the sum can be cut alone

    net = Skip()
    with pytest.raises(NoLegalCut, match="exactly one activation"):
        GraphPartitioner(net, None, n_partitions=3).split()
    stages = GraphPartitioner(net, None, n_partitions=2).split()   # the one legal cut: right after ``a``
    x = torch.randn(3, 4)
    assert torch.allclose(_chain(stages, x), net(x))
    assert [sum(p.numel() for p in s.parameters()) for s in stages] == [20, 40]


def test_uniform_partitioner_falls_back_to_the_graph_and_explains_failures():
    class Ctx:
        pipeline_parallel_size = 2

    torch.manual_seed(0)
    net = _TokenModel(n_blocks=4)
    stages = UniformPartitioner(net, Ctx()).split()
    assert len(stages) == 2 and isinstance(stages[0], GraphStage)

    class DataDependent(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = nn.Linear(4, 4), nn.Linear(4, 4)

        def forward(self, x):
            return self.a(x) if x.sum() > 0 else self.b(x)

    with pytest.raises(NotImplementedError, match="register_family.*torch.fx could not trace"):
        UniformPartitioner(DataDependent(), Ctx()).split()


@pytest.mark.parametrize("costs,n,want", [([1, 1, 1, 1], 2, [0, 2, 4]), ([5, 1, 1, 1, 1, 1], 2, [0, 1, 6]),
                                          ([1, 2160, 2160, 2160, 2160, 2160, 2160, 850], 2, [0, 4, 8]),
                                          ([3, 3, 3], 3, [0, 1, 2, 3])])
def test_minmax_cuts(costs, n, want):
    assert _minmax_cuts(costs, n) == want


def run_graph_pipeline(rank, world_size, port, pp, sched, state, ids, mask, ref_loss, ref_grads):
    ctx = init_parallel_context(rank, world_size, port, 1, pp, 1)
    model = _TokenModel()
    model.load_state_dict(state)
    names = {id(p): n for n, p in model.named_parameters()}
    custom = None if pp == 2 else (lambda m, c: GraphPartitioner(m, c, leaf_modules=(_Block,)))   # both entry points
    model = PipelineParallel(model, num_microbatches=4, parallel_context=ctx, scheduler_type=sched, partitioner=custom).parallelize()
    assert isinstance(model._pg_pipeline_stage, GraphStage)
    out = model(ids, mask=mask, labels=ids)
    assert torch.allclose(out.loss, ref_loss, atol=1e-5)
    out.loss.backward()
    n_checked = 0
    for p in model._pg_pipeline_stage.parameters():
        assert torch.allclose(p.grad, ref_grads[names[id(p)]], atol=2e-5), names[id(p)]
        n_checked += 1
    assert n_checked > 0
    with torch.no_grad():
        assert torch.allclose(model(ids, mask=mask, labels=ids).loss, ref_loss, atol=1e-5)
    ctx.destroy()


@pytest.mark.parametrize("pp,sched", [(2, SchedulerType.ONE_F_ONE_B), (3, SchedulerType.GPIPE)])
def test_pipeline_engine_trains_a_graph_partitioned_model(pp, sched):
    torch.manual_seed(0)
    model = _TokenModel()
    ids = torch.randint(0, 50, (8, 6))
    mask = (torch.rand(8, 6) > 0.2).float()
    losses = [model(i, m, i) for i, m in zip(ids.chunk(4), mask.chunk(4))]
    loss = torch.stack(losses).mean()
    loss.backward()
    spawn(run_graph_pipeline, world_size=pp, pp=pp, sched=sched, state=copy.deepcopy(model.state_dict()), ids=ids, mask=mask,
          ref_loss=loss.detach(), ref_grads={n: p.grad.clone() for n, p in model.named_parameters()})


def test_parameters_shared_between_stages_are_refused_unless_they_are_the_tied_embedding():
    class Shared(nn.Module):
        def __init__(self, expose):
            super().__init__()
            self.emb, self.mid, self.head = nn.Embedding(20, 8), nn.Linear(8, 8), nn.Linear(8, 20, bias=False)
            self.head.weight = self.emb.weight
            self.expose = expose

        def get_input_embeddings(self):
            return self.emb if self.expose else None

        def get_output_embeddings(self):
            return self.head if self.expose else None

        def forward(self, tokens):
            return self.head(torch.tanh(self.mid(self.emb(tokens))))

    with pytest.raises(NoLegalCut, match="emb.weight is used by pipeline stages"):
        GraphPartitioner(Shared(False), None, n_partitions=2).split()
    stages = GraphPartitioner(Shared(True), None, n_partitions=2).split()     # the engine sums the tied table's gradient
    assert any(p is stages[1].graph_module.head.weight for p in stages[0].parameters())


def test_leaf_modules_keep_untraceable_blocks_as_single_nodes():
    class Gate(nn.Module):   # data-dependent control flow: torch.fx cannot look inside
        def __init__(self):
            super().__init__()
            self.a, self.b = nn.Linear(6, 6), nn.Linear(6, 6)

        def forward(self, x):
            return x + (self.a(x) if float(x.detach().sum()) > 0 else self.b(x))

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.inp = nn.Linear(3, 6)
            self.gates = nn.ModuleList([Gate() for _ in range(4)])

        def forward(self, x):
            x = self.inp(x)
            for g in self.gates:
                x = g(x)
            return x

    torch.manual_seed(0)
    net = Net()
    with pytest.raises(Exception):
        GraphPartitioner(net, None, n_partitions=2).split()
    stages = GraphPartitioner(net, None, n_partitions=2, leaf_modules=(Gate,)).split()
    for x in (torch.randn(5, 3), -torch.rand(5, 3) - 5.0):
        assert torch.allclose(_chain(stages, x), net(x), atol=1e-6)
    assert [sum(isinstance(m, Gate) for m in s.modules()) for s in stages] == [2, 2]


class _UNet(nn.Module):
    """Long skip connections: between the encoder and the decoder TWO or THREE activations are live."""

    def __init__(self, d=8):
        super().__init__()
        self.e1, self.e2, self.e3 = nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, d)
        self.mid = nn.Linear(d, d)
        self.d3, self.d2, self.d1 = nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, d)
        self.out = nn.Linear(d, 1)

    def forward(self, x, target=None):
        a = torch.tanh(self.e1(x))
        b = torch.tanh(self.e2(a))
        c = torch.tanh(self.e3(b))
        m = torch.tanh(self.mid(c))
        y = torch.tanh(self.d3(m) + c)
        y = torch.tanh(self.d2(y) + b)
        y = torch.tanh(self.d1(y) + a)
        pred = self.out(y).squeeze(-1)
        if target is not None:
            return ((pred - target) ** 2).mean()
        return pred


def test_cuts_through_skip_connections_pack_several_activations():
    torch.manual_seed(0)
    net = _UNet()
    x, target = torch.randn(6, 8), torch.randn(6)
    # single-activation cuts exist only before / after the whole U: one stage gets the encoder, the middle and the decoder
    single = GraphPartitioner(net, None, n_partitions=4).split()
    assert max(sum(p.numel() for p in s.parameters()) for s in single) >= 5 * 72
    stages = GraphPartitioner(net, None, n_partitions=4, max_boundary_tensors=4).split()
    assert any(s.multi_out for s in stages) and [s.multi_out for s in stages[:-1]] == [s.multi_in for s in stages[1:]]
    sizes = [sum(p.numel() for p in s.parameters()) for s in stages]
    assert max(sizes) <= 2 * 72 + 9       # 7 equal layers + the head over 4 stages
    # chained by hand with tuples
    for s in stages:
        s.pack_outputs = False
    assert torch.allclose(_chain(stages, x, target=target), net(x, target), atol=1e-6)
    # chained through packed buffers, as the engine does: metadata from the producer, gradients through pack / unpack
    for s in stages:
        s.pack_outputs = True
    net.zero_grad()
    net(x, target).backward()
    want = {n: p.grad.clone() for n, p in net.named_parameters()}
    net.zero_grad()
    h = x
    for i, s in enumerate(stages):
        if s.multi_in:
            assert h.dim() == 1
            s.set_in_meta("k", stages[i - 1].last_out_meta)
            s.select_boundary("k")
        h = s(h, target=target)
    assert torch.allclose(h, net(x, target), atol=1e-6)
    h.backward()
    for n, p in net.named_parameters():
        assert torch.allclose(p.grad, want[n], atol=1e-6), n
    with pytest.raises(TypeError, match="floating-point"):
        GraphStage.pack([torch.zeros(2), torch.zeros(2, dtype=torch.long)])


def run_unet_pipeline(rank, world_size, port, pp, state, x, target, ref_loss, ref_grads):
    ctx = init_parallel_context(rank, world_size, port, 1, pp, 1)
    model = _UNet()
    model.load_state_dict(state)
    names = {id(p): n for n, p in model.named_parameters()}
    model = PipelineParallel(model, num_microbatches=3, parallel_context=ctx, scheduler_type=SchedulerType.ONE_F_ONE_B,
                             partitioner=lambda m, c: GraphPartitioner(m, c, max_boundary_tensors=4)).parallelize()
    # labels= switches the engine to its training schedule; the model's own name for them travels as a named input
    out = model(x, target=target, labels=target)
    assert torch.allclose(out.loss, ref_loss, atol=1e-6)
    out.loss.backward()
    for p in model._pg_pipeline_stage.parameters():
        assert torch.allclose(p.grad, ref_grads[names[id(p)]], atol=1e-6), names[id(p)]
    # a second step with a different micro-batch size: new handshake, new boundary metadata
    out2 = model(x[:3], target=target[:3], labels=target[:3])
    assert torch.isfinite(out2.loss)
    ctx.destroy()


@pytest.mark.parametrize("pp", [2, 4])
def test_pipeline_engine_moves_packed_boundaries(pp):
    torch.manual_seed(0)
    model = _UNet()
    x, target = torch.randn(6, 8), torch.randn(6)
    losses = [model(a, b) for a, b in zip(x.chunk(3), target.chunk(3))]
    loss = torch.stack(losses).mean()
    loss.backward()
    spawn(run_unet_pipeline, world_size=pp, pp=pp, state=copy.deepcopy(model.state_dict()), x=x, target=target,
          ref_loss=loss.detach(), ref_grads={n: p.grad.clone() for n, p in model.named_parameters()})


class _RandomMaskModel(nn.Module):
    """A random keep-mask drawn from the inputs alone (no parameter upstream) and used by every block: a stage must not
    draw its own."""

    def __init__(self, d=8):
        super().__init__()
        self.fc = nn.ModuleList([nn.Linear(d, d) for _ in range(4)])
        self.drop = nn.Dropout(0.5)

    def forward(self, x):
        keep = self.drop(torch.ones_like(x))                    # module without parameters, input-derived argument
        noise = torch.rand_like(x)                              # function of the input alone
        h = x
        for fc in self.fc:
            h = fc(h) * keep + noise
        return h


def test_random_values_are_computed_once_and_carried():
    torch.manual_seed(0)
    net = _RandomMaskModel().train()
    x = torch.randn(3, 8)
    def where(stages, what):
        return [any(n.target == what for n in s.graph_module.graph.nodes) for s in stages]

    # single-activation cuts: the only one is right behind the mask (before the noise is drawn) — everything else has
    # mask + noise + hidden state live.  Three stages cannot be cut that way.
    two = GraphPartitioner(net, None, n_partitions=2).split()
    assert where(two, "drop") == [True, False] and where(two, torch.rand_like) == [False, True]
    with pytest.raises(NoLegalCut):
        GraphPartitioner(net, None, n_partitions=3).split()
    # with room for three tensors the mask and the noise travel with the hidden state instead of being drawn again
    three = GraphPartitioner(net, None, n_partitions=3, max_boundary_tensors=3).split()
    assert sum(where(three, "drop")) == 1 and sum(where(three, torch.rand_like)) == 1
    assert all(sum(p.numel() for p in s.parameters()) > 0 for s in three)
    for stages in (two, three):
        for s in stages:
            s.pack_outputs = False
        torch.manual_seed(7)
        want = net(x)
        torch.manual_seed(7)
        assert torch.allclose(_chain(stages, x, x=x), want)      # (later stages read the input by name, as in the engine)


class _InputSkipModel(nn.Module):
    """``forward(self, x, labels)``: the raw input is added again at the end (every stage can read it from the micro-batch),
    and the argument is called ``x`` like the stages' own carried value."""

    def __init__(self, d=8):
        super().__init__()
        self.fc = nn.ModuleList([nn.Linear(d, d) for _ in range(4)])

    def forward(self, x, labels=None):
        h = x
        for fc in self.fc:
            h = torch.tanh(fc(h))
        out = h + x
        if labels is not None:
            return ((out - labels) ** 2).mean()
        return out


def run_input_skip(rank, world_size, port, state, x, y, ref_loss, ref_grads):
    ctx = init_parallel_context(rank, world_size, port, 1, world_size, 1)
    model = _InputSkipModel()
    model.load_state_dict(state)
    names = {id(p): n for n, p in model.named_parameters()}
    model = PipelineParallel(model, num_microbatches=2, parallel_context=ctx,
                             partitioner=lambda m, c: GraphPartitioner(m, c)).parallelize()
    out = model(x, labels=y)
    assert torch.allclose(out.loss, ref_loss, atol=1e-6)
    out.loss.backward()
    for p in model._pg_pipeline_stage.parameters():
        assert torch.allclose(p.grad, ref_grads[names[id(p)]], atol=1e-6), names[id(p)]
    ctx.destroy()


def test_engine_feeds_the_raw_input_to_later_stages_by_name():
    torch.manual_seed(0)
    model = _InputSkipModel()
    x, y = torch.randn(4, 8), torch.randn(4, 8)
    loss = torch.stack([model(a, b) for a, b in zip(x.chunk(2), y.chunk(2))]).mean()
    loss.backward()
    spawn(run_input_skip, world_size=2, state=copy.deepcopy(model.state_dict()), x=x, y=y, ref_loss=loss.detach(),
          ref_grads={n: p.grad.clone() for n, p in model.named_parameters()})
