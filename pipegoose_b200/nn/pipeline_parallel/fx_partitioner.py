"""Graph-based pipeline partitioning of an arbitrary traceable ``nn.Module`` (parity: the part of the reference's
nn/pipeline_parallel/partitioner.py:146-244 that is not tied to a model family — trace, balance by parameter count,
rebuild one ``GraphModule`` per shard).

The reference traces 🤗 models with ``transformers.utils.fx`` and threads every value that crosses a cut through the
shards as a tuple.  The pipeline engine here moves ONE activation tensor per micro-batch between neighbouring stages
(static shapes, one NCCL send/recv pair per boundary), so this partitioner looks for the cuts a pipeline wants anyway:

* the model is traced with ``torch.fx`` (``symbolic_trace`` or a tracer / ``GraphModule`` the caller supplies);
* every node is classified as *input-derived* (a forward argument, a buffer, or a parameter-free DETERMINISTIC function
  of those: masks, position ids, sequence lengths, ALiBi slopes, ...) or as an *activation* (anything downstream of a
  parameter, and anything drawn from a random generator — dropout, ``torch.rand`` — which must be computed once);
* a cut between two nodes is legal when exactly one activation is live across it — the residual stream at a block
  boundary, the hidden state between two layers of an MLP or a CNN.  Input-derived values never cross a cut: every
  stage that needs one re-computes it from the micro-batch's inputs, which every stage holds (no transfer at all).
  ``max_boundary_tensors=k`` also admits cuts with up to ``k`` live activations (long skip connections, U-Nets, a
  (hidden, residual) pair): the stage packs them into ONE flat buffer for the transfer and the next stage unpacks it —
  the reference's "every value that crosses the cut" tuple, still one send/recv pair per boundary;
* the legal cuts divide the graph into segments; the segments are balanced over the stages by parameter count
  (embeddings excluded, as in the reference) so that the largest stage is as small as possible.

Each stage is a ``GraphModule`` rooted at the original model, so it owns exactly the sub-modules / parameters its
nodes use.  ``stage(x, **inputs)``: ``x`` is the first input (first stage) or the carried activation (a tuple / the
packed buffer when several cross); ``inputs`` are the model's forward arguments by name (``stage.stage_inputs`` lists the
ones this stage reads).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Set

import torch
from torch import fx, nn

from pipegoose_b200.nn.pipeline_parallel.partitioner import BasePartitioner


class NoLegalCut(NotImplementedError):
    """The traced graph has fewer single-activation cut points than the pipeline has boundaries."""


def _minmax_cuts(costs: List[int], n_parts: int) -> List[int]:
    """Boundaries ``[0, b1, ..., len(costs)]`` of the contiguous split into ``n_parts`` non-empty groups whose largest
    group is smallest (exact, O(parts * items^2): graphs have a few hundred cut points at most); ties go to the split
    that is also best for the remaining groups."""
    m = len(costs)
    assert m >= n_parts
    prefix = [0]
    for c in costs:
        prefix.append(prefix[-1] + c)
    inf = float("inf")
    best = [[inf] * (m + 1) for _ in range(n_parts + 1)]   # best[k][i]: first i items in k groups
    arg = [[0] * (m + 1) for _ in range(n_parts + 1)]
    best[0][0] = 0
    for k in range(1, n_parts + 1):
        for i in range(k, m - (n_parts - k) + 1):
            for j in range(k - 1, i):
                if best[k - 1][j] == inf:
                    continue
                v = max(best[k - 1][j], prefix[i] - prefix[j])
                if v < best[k][i]:
                    best[k][i], arg[k][i] = v, j
    bounds, i = [m], m
    for k in range(n_parts, 0, -1):
        i = arg[k][i]
        bounds.append(i)
    return bounds[::-1]


def _arg_nodes(node: fx.Node) -> List[fx.Node]:
    return list(node.all_input_nodes)


# Values drawn from a random generator must be computed ONCE: a stage that re-computed them "from the inputs" would
# draw other numbers than its neighbour (a random mask, dropout on position features).  They count as activations.
_RANDOM_FUNCTIONS = {torch.rand, torch.randn, torch.randint, torch.rand_like, torch.randn_like, torch.randint_like,
                     torch.randperm, torch.bernoulli, torch.multinomial, torch.normal, torch.poisson, torch.dropout,
                     torch.nn.functional.dropout, torch.nn.functional.dropout1d, torch.nn.functional.dropout2d,
                     torch.nn.functional.dropout3d, torch.nn.functional.alpha_dropout,
                     torch.nn.functional.feature_alpha_dropout, torch.nn.functional.gumbel_softmax,
                     torch.nn.functional.rrelu}
_RANDOM_METHODS = {"bernoulli", "bernoulli_", "normal_", "uniform_", "random_", "exponential_", "geometric_",
                   "cauchy_", "log_normal_", "multinomial", "dropout"}
_RANDOM_MODULES = (nn.Dropout, nn.Dropout1d, nn.Dropout2d, nn.Dropout3d, nn.AlphaDropout, nn.FeatureAlphaDropout, nn.RReLU)


def _draws_random_numbers(gm: fx.GraphModule, node: fx.Node) -> bool:
    if node.op == "call_function":
        return node.target in _RANDOM_FUNCTIONS
    if node.op == "call_method":
        return node.target in _RANDOM_METHODS
    if node.op == "call_module":
        obj = gm
        for part in node.target.split("."):
            obj = getattr(obj, part)
        return isinstance(obj, _RANDOM_MODULES)
    return False


class GraphStage(nn.Module):
    """One shard of a traced model.  ``forward(x, **inputs)`` (see the module docstring).

    Boundaries with several activations: the stage returns them as ONE packed 1-D tensor (``pack``) and remembers their
    shapes in ``last_out_meta``; the next stage unpacks what it receives with the metadata given to ``set_in_meta`` (the
    pipeline engine exchanges it once per micro-batch shape during its handshake).  Called with a tuple — stages
    chained by hand — a stage takes the values as they are."""

    def __init__(self, graph_module: fx.GraphModule, carried_names: Sequence[str], n_out: int, stage_inputs: Sequence[str],
                 first_input_name: str, is_first: bool, is_last: bool):
        super().__init__()
        self.graph_module = graph_module
        self.carried_names = tuple(carried_names)    # placeholders of the incoming activation(s)
        self.n_out = n_out                           # activations handed to the next stage (0 on the last stage)
        self.stage_inputs = tuple(stage_inputs)      # forward arguments of the model this stage reads by name
        self.first_input_name = first_input_name     # the model's first forward argument (the engine's ``input_ids`` slot)
        self.is_first, self.is_last = is_first, is_last
        self.multi_in, self.multi_out = len(self.carried_names) > 1, n_out > 1
        self.pack_outputs = True                     # False: return the tuple itself (chaining stages in one process)
        self._in_meta: Dict = {}
        self._key = None
        self.last_out_meta = None

    # -- packed boundaries
    def set_in_meta(self, key, meta) -> None:
        self._in_meta[key] = [(tuple(shape), dtype) for shape, dtype in meta]

    def select_boundary(self, key) -> None:
        self._key = key

    @staticmethod
    def pack(tensors: Sequence[torch.Tensor]):
        for t in tensors:
            if not (isinstance(t, torch.Tensor) and t.is_floating_point()):
                raise TypeError("only floating-point tensors can cross a pipeline cut as activations, got "
                                f"{type(t).__name__}{'' if not isinstance(t, torch.Tensor) else ' of ' + str(t.dtype)}")
        dtype = tensors[0].dtype
        for t in tensors[1:]:
            dtype = torch.promote_types(dtype, t.dtype)
        meta = [(tuple(t.shape), t.dtype) for t in tensors]
        return torch.cat([t.reshape(-1).to(dtype) for t in tensors]), meta

    def unpack(self, flat: torch.Tensor):
        meta = self._in_meta.get(self._key)
        if meta is None and len(self._in_meta) == 1:
            meta = next(iter(self._in_meta.values()))
        assert meta is not None, "no boundary metadata: the pipeline engine's handshake (or set_in_meta) must come first"
        sizes = [int(torch.Size(shape).numel()) for shape, _ in meta]
        return [part.view(shape).to(dtype) for part, (shape, dtype) in zip(flat.split(sizes), meta)]

    def forward(self, x, /, **inputs):     # (positional-only: a traced model may call ITS first argument ``x`` too)
        if self.multi_in:
            values = list(x) if isinstance(x, (tuple, list)) else self.unpack(x)
            kwargs = dict(zip(self.carried_names, values))
        else:
            kwargs = {self.carried_names[0]: x}
        for name in self.stage_inputs:
            if name not in kwargs:
                kwargs[name] = inputs.get(name)
        out = self.graph_module(**kwargs)
        if self.multi_out:
            if not self.pack_outputs:
                return tuple(out)
            out, self.last_out_meta = self.pack(list(out))
        return out


class _LeafTracer(fx.Tracer):
    def __init__(self, leaf_types):
        super().__init__()
        self._leaf_types = tuple(leaf_types)

    def is_leaf_module(self, m: nn.Module, module_qualified_name: str) -> bool:
        return isinstance(m, self._leaf_types) or super().is_leaf_module(m, module_qualified_name)


class GraphPartitioner(BasePartitioner):
    """``GraphPartitioner(model, parallel_context).split() -> [GraphStage]`` for any model ``torch.fx`` can trace.

    ``concrete_args`` / ``tracer`` are handed to the tracing step (a model whose forward branches on ``labels is None``
    needs ``concrete_args={"labels": None}`` or a real ``labels`` placeholder, exactly as with ``torch.fx`` itself); a
    ready ``GraphModule`` is taken as is.  ``leaf_modules``: module classes to keep as single graph nodes.
    ``max_boundary_tensors``: see the module docstring."""

    def __init__(self, model: nn.Module, parallel_context, concrete_args: Optional[Dict] = None,
                 tracer: Optional[fx.Tracer] = None, n_partitions: Optional[int] = None,
                 leaf_modules: Sequence[type] = (), max_boundary_tensors: int = 1):
        self.module = model
        self.parallel_context = parallel_context
        self.concrete_args = concrete_args
        self.tracer = tracer
        self._n = n_partitions
        assert max_boundary_tensors >= 1
        self.max_boundary_tensors = max_boundary_tensors   # activations that may cross one cut (packed into one buffer)
        # module classes the tracer does not look into (a block whose forward torch.fx cannot trace — data-dependent
        # control flow, Python-side caches — is still one node of the graph, and block boundaries are the cuts anyway)
        self.leaf_modules = tuple(leaf_modules)
        if self.leaf_modules and tracer is None:
            self.tracer = _LeafTracer(self.leaf_modules)

    # ------------------------------------------------------------------ tracing and node classes
    def trace(self) -> fx.GraphModule:
        if isinstance(self.module, fx.GraphModule):
            return self.module
        if self.tracer is not None:
            graph = self.tracer.trace(self.module, concrete_args=self.concrete_args)
            return fx.GraphModule(self.module, graph)
        return fx.symbolic_trace(self.module, concrete_args=self.concrete_args)

    @staticmethod
    def _fetch(root: nn.Module, target: str):
        obj = root
        for part in target.split("."):
            obj = getattr(obj, part)
        return obj

    def _node_cost(self, gm: fx.GraphModule, node: fx.Node, seen: Set[int]) -> int:
        """Trainable parameters a node brings into its shard (each parameter counted once; embeddings count 0, the
        reference's rule: they skew the balance of the blocks)."""
        params: List[nn.Parameter] = []
        if node.op == "call_module":
            mod = self._fetch(gm, node.target)
            if isinstance(mod, nn.Embedding):
                for p in mod.parameters():
                    seen.add(id(p))
                return 0
            params = list(mod.parameters())
        elif node.op == "get_attr":
            obj = self._fetch(gm, node.target)
            if isinstance(obj, nn.Parameter):
                params = [obj]
        cost = 0
        for p in params:
            if id(p) not in seen:
                seen.add(id(p))
                cost += p.numel()
        return cost

    def _input_derived(self, gm: fx.GraphModule, nodes: List[fx.Node]) -> Set[fx.Node]:
        """Nodes every stage can compute for itself: the forward arguments (every stage holds the micro-batch), buffers,
        and parameter-free functions of those."""
        free: Set[fx.Node] = set()
        for n in nodes:
            if n.op == "placeholder":
                free.add(n)
            elif n.op == "get_attr":
                if not isinstance(self._fetch(gm, n.target), nn.Parameter):
                    free.add(n)
            elif _draws_random_numbers(gm, n):
                continue                      # computed once, carried like an activation
            elif n.op in ("call_function", "call_method"):
                if all(d in free for d in _arg_nodes(n)):
                    free.add(n)
            elif n.op == "call_module":
                mod = self._fetch(gm, n.target)
                if next(mod.parameters(), None) is None and all(d in free for d in _arg_nodes(n)):
                    free.add(n)
        return free

    # ------------------------------------------------------------------ cut points
    def legal_cuts(self, gm: fx.GraphModule):
        """``(nodes, free, cuts)``: ``cuts[k] = (position, carried nodes)`` — cutting before ``nodes[position]`` hands
        exactly those activations (at most ``max_boundary_tensors``, in graph order) to the next stage."""
        nodes = [n for n in gm.graph.nodes if n.op != "output"]
        output = next(n for n in gm.graph.nodes if n.op == "output")
        assert any(n.op == "placeholder" for n in nodes), "the traced forward takes no argument"
        free = self._input_derived(gm, nodes)
        index = {n: i for i, n in enumerate(nodes)}
        index[output] = len(nodes)
        last_use = {}
        for n in nodes:
            if n in free:
                continue
            uses = [index[u] for u in n.users if u in index]
            last_use[n] = max(uses) if uses else -1
        first_compute = next((i for i, n in enumerate(nodes) if n not in free), len(nodes))
        cuts = []
        live: List[fx.Node] = []
        for pos in range(1, len(nodes)):
            prev = nodes[pos - 1]
            if prev not in free:
                live.append(prev)
            live = [n for n in live if last_use[n] >= pos]
            if pos <= first_compute:
                continue        # nothing computed yet: an empty first stage is no stage
            if 1 <= len(live) <= self.max_boundary_tensors:
                cuts.append((pos, tuple(live)))
        return nodes, free, cuts

    # ------------------------------------------------------------------ split
    def _n_partitions(self) -> int:
        return self._n if self._n is not None else self.parallel_context.pipeline_parallel_size

    def split(self, input_names: Optional[List[str]] = None) -> List[nn.Module]:
        n = self._n_partitions()
        gm = self.trace()
        nodes, free, cuts = self.legal_cuts(gm)
        if len(cuts) < n - 1:
            k = self.max_boundary_tensors
            raise NoLegalCut(
                f"{type(self.module).__name__}: {n} pipeline stages need {n - 1} cut points where "
                f"{'exactly one activation tensor is' if k == 1 else f'at most {k} activation tensors are'} live, the traced "
                f"graph has {len(cuts)}.  (Values computed from the forward arguments alone are re-computed by each stage "
                "and do not count; long skip connections do — max_boundary_tensors admits cuts through them.)")
        seen: Set[int] = set()
        costs = [self._node_cost(gm, node, seen) for node in nodes]
        # segments between consecutive legal cuts, balanced like blocks
        edges = [0] + [pos for pos, _ in cuts] + [len(nodes)]
        seg_costs = [sum(costs[edges[i]:edges[i + 1]]) or 1 for i in range(len(edges) - 1)]
        b = _minmax_cuts(seg_costs, n)
        carried_at = {pos: carried for pos, carried in cuts}
        bounds = [edges[i] for i in b]          # node positions where each stage starts (+ the end)
        first = next(x for x in nodes if x.op == "placeholder")
        output = next(x for x in gm.graph.nodes if x.op == "output")
        stages = []
        for s in range(n):
            lo, hi = bounds[s], bounds[s + 1]
            carried_in = (first,) if s == 0 else carried_at[lo]
            carried_out = () if s == n - 1 else carried_at[hi]
            stages.append(self._build_stage(gm, nodes, free, lo, hi, carried_in, carried_out, output, first,
                                            is_first=(s == 0), is_last=(s == n - 1)))
        self._check_shared_parameters(stages)
        return stages

    def _check_shared_parameters(self, stages: List["GraphStage"]) -> None:
        """A parameter used by two stages would get only a part of its gradient on each of them.  The one sharing the
        pipeline engine completes itself is the tied input / output embedding of a model that exposes it through
        ``get_input_embeddings()`` / ``get_output_embeddings()`` (first and last stage); anything else is refused."""
        names = {id(p): n for n, p in self.module.named_parameters()}
        tied = None
        get_in, get_out = getattr(self.module, "get_input_embeddings", None), getattr(self.module, "get_output_embeddings", None)
        if callable(get_in) and callable(get_out):
            emb, head = get_in(), get_out()
            if emb is not None and head is not None and getattr(head, "weight", None) is getattr(emb, "weight", 0):
                tied = id(emb.weight)
        owners: Dict[int, List[int]] = {}
        for s, stage in enumerate(stages):
            for p in stage.parameters():
                owners.setdefault(id(p), []).append(s)
        for pid, where in owners.items():
            if len(where) > 1 and not (pid == tied and where == [0, len(stages) - 1]):
                raise NoLegalCut(f"parameter {names.get(pid, '?')} is used by pipeline stages {where}: its gradient would be "
                                 "split between them (only a tied input/output embedding exposed through "
                                 "get_input_embeddings()/get_output_embeddings() is summed across stages)")

    def _build_stage(self, gm, nodes, free, lo, hi, carried_in, carried_out, output, first, is_first, is_last) -> GraphStage:
        graph = fx.Graph()
        env: Dict[fx.Node, fx.Node] = {}
        if is_first:
            carried_names = [first.name]
        elif len(carried_in) == 1:
            carried_names = ["carried_activation"]
        else:
            carried_names = [f"carried_activation_{i}" for i in range(len(carried_in))]
        for node, name in zip(carried_in, carried_names):
            env[node] = graph.placeholder(name)
        body = [x for x in nodes[lo:hi] if x not in free]
        # input-derived values the body (or the model's output) needs, in graph order, with their own dependencies
        needed: Set[fx.Node] = set()
        stack = [d for x in body for d in _arg_nodes(x)]
        if is_last:
            stack += _arg_nodes(output)
        while stack:
            d = stack.pop()
            if d in free and d not in needed and d not in env:
                needed.add(d)
                stack.extend(_arg_nodes(d))
        stage_inputs = []
        for x in nodes:
            if x in needed and x.op == "placeholder":
                env[x] = graph.placeholder(x.name, default_value=None)
                stage_inputs.append(x.name)
        for x in nodes:
            if x in needed and x.op != "placeholder":
                env[x] = graph.node_copy(x, lambda a: env[a])
        for x in body:
            missing = [d for d in _arg_nodes(x) if d not in env]
            if missing:
                raise NoLegalCut(f"stage [{lo}, {hi}) reads {[m.name for m in missing]} which an earlier stage computed and "
                                 "did not hand over (internal error in legal_cuts)")
            env[x] = graph.node_copy(x, lambda a: env[a])
        if is_last:
            graph.output(fx.node.map_arg(output.args[0], lambda a: env[a]))
        elif len(carried_out) == 1:
            graph.output(env[carried_out[0]])
        else:
            graph.output(tuple(env[c] for c in carried_out))
        graph.lint()
        sub = fx.GraphModule(gm, graph, class_name=f"{type(self.module).__name__}Stage")
        return GraphStage(sub, carried_names, len(carried_out), stage_inputs, first.name, is_first, is_last)
