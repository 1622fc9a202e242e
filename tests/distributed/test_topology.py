import pytest

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.distributed.topology import Topology


def test_parallel_mode_members():
    assert ParallelMode.GLOBAL.value == "global"
    assert ParallelMode.TENSOR.value == "tensor"
    assert ParallelMode.PIPELINE.value == "pipeline"
    assert ParallelMode.DATA.value == "data"
    assert ParallelMode.EXPERT_DATA.value == "expert"


def test_layout_tp2_pp2_dp2():
    # the layout asserted by the reference's initializer tests (tests/distributed/_initializers/*)
    topo = Topology(8, 2, 2, 2)
    assert topo.groups(ParallelMode.TENSOR) == [[0, 1], [2, 3], [4, 5], [6, 7]]
    assert topo.groups(ParallelMode.DATA) == [[0, 2], [1, 3], [4, 6], [5, 7]]
    assert topo.groups(ParallelMode.PIPELINE) == [[0, 4], [1, 5], [2, 6], [3, 7]]
    assert topo.groups(ParallelMode.EXPERT) == topo.groups(ParallelMode.TENSOR)
    assert topo.groups(ParallelMode.EXPERT_DATA) == topo.groups(ParallelMode.DATA)
    assert topo.groups(ParallelMode.GLOBAL) == [list(range(8))]


@pytest.mark.parametrize("tp,pp,dp", [(1, 1, 1), (2, 1, 1), (2, 4, 2), (8, 1, 1), (2, 2, 2), (1, 4, 2)])
def test_every_rank_is_in_exactly_one_group_per_mode(tp, pp, dp):
    world = tp * pp * dp
    topo = Topology(world, tp, pp, dp)
    for mode in ParallelMode:
        seen = sorted(r for g in topo.groups(mode) for r in g)
        assert seen == list(range(world))
    for r in range(world):
        c = topo.coord(r)
        assert topo.rank_of(c.pp, c.dp, c.tp) == r
        assert topo.local_rank(r, ParallelMode.TENSOR) == c.tp
        assert topo.local_rank(r, ParallelMode.DATA) == c.dp
        assert topo.local_rank(r, ParallelMode.PIPELINE) == c.pp


def test_invalid_sizes():
    with pytest.raises(AssertionError):
        Topology(8, 2, 2, 3)


def _run_node_probe(rank, world_size, port):
    import os

    from pipegoose_b200.distributed.parallel_mode import ParallelMode
    from pipegoose_b200.distributed.symmetric import _NODE_CACHE, peers_share_a_node
    from pipegoose_b200.testing.utils import init_parallel_context

    ctx = init_parallel_context(rank, world_size, port, 2, 1, 2)
    assert peers_share_a_node(ctx, ParallelMode.TENSOR) and peers_share_a_node(ctx, ParallelMode.DATA)
    _NODE_CACHE.clear()
    os.environ["PIPEGOOSE_B200_FAKE_NODE"] = str(rank // 2)   # pretend ranks {0,1} and {2,3} sit on two hosts
    assert peers_share_a_node(ctx, ParallelMode.TENSOR)        # [0,1] / [2,3]
    assert not peers_share_a_node(ctx, ParallelMode.DATA)      # [0,2] / [1,3]: peer mapping impossible -> NCCL paths
    ctx.destroy()


def test_groups_that_span_hosts_are_detected():
    from pipegoose_b200.testing.utils import spawn

    spawn(_run_node_probe, world_size=4)


def test_exhaustive_small_layouts_are_orthogonal_partitions():
    """Every (tp, pp, dp) with up to 48 ranks: groups of a mode partition the world into equally sized groups, groups
    of two different modes meet in exactly one rank (the layouts are orthogonal axes of one grid), tensor-parallel peers
    are adjacent ranks (they share NVLink inside a node) and the pipeline axis is the outermost one."""
    axes = {ParallelMode.TENSOR: "tp", ParallelMode.PIPELINE: "pp", ParallelMode.DATA: "dp"}
    n = 0
    for tp in (1, 2, 3, 4, 8):
        for pp in (1, 2, 3, 4):
            for dp in (1, 2, 3, 4, 6):
                world = tp * pp * dp
                if world > 48:
                    continue
                n += 1
                topo = Topology(world, tp, pp, dp)
                size = {ParallelMode.TENSOR: tp, ParallelMode.PIPELINE: pp, ParallelMode.DATA: dp}
                for mode in axes:
                    groups = topo.groups(mode)
                    assert all(len(g) == size[mode] for g in groups) and len(groups) == world // size[mode]
                    assert sorted(r for g in groups for r in g) == list(range(world))
                    assert all(g == sorted(g) for g in groups)
                for g in topo.groups(ParallelMode.TENSOR):
                    assert g == list(range(g[0], g[0] + tp))
                for g in topo.groups(ParallelMode.PIPELINE):
                    assert all(b - a == tp * dp for a, b in zip(g, g[1:]))
                for m1 in axes:
                    for m2 in axes:
                        if m1 is m2 or size[m1] == 1 or size[m2] == 1:
                            continue
                        for g1 in topo.groups(m1):
                            for g2 in topo.groups(m2):
                                assert len(set(g1) & set(g2)) <= 1
                        # ... and through every rank passes exactly one group of each mode
                for r in range(world):
                    for mode in axes:
                        assert sum(r in g for g in topo.groups(mode)) == 1
    assert n > 50
