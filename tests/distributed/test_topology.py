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
