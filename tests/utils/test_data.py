"""Input pipeline helpers (utils/data.py): the 3-D-parallel sampler, token-file datasets, the device prefetcher."""
import os

import pytest
import torch
from torch.utils.data import DataLoader

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.testing.utils import init_parallel_context, spawn
from pipegoose_b200.utils.data import (DevicePrefetcher, TokenFileDataset, build_dataloader, data_parallel_sampler,
                                       write_token_file)


def test_token_file_dataset(tmp_path):
    path = str(tmp_path / "tokens.bin")
    tokens = torch.arange(0, 103) % 50000
    write_token_file(path, tokens)
    assert os.path.getsize(path) == 103 * 2
    ds = TokenFileDataset(path, seq_len=10)
    assert len(ds) == 10 and ds.n_tokens == 103
    assert torch.equal(ds[3]["input_ids"], tokens[30:40]) and ds[3]["input_ids"].dtype == torch.int64
    assert torch.equal(ds[-1]["input_ids"], tokens[90:100])
    with pytest.raises(IndexError):
        ds[10]
    overlapping = TokenFileDataset(path, seq_len=10, stride=5)
    assert len(overlapping) == 19 and torch.equal(overlapping[1]["input_ids"], tokens[5:15])
    with pytest.raises(ValueError, match="do not fit"):
        write_token_file(path, [70000])
    write_token_file(path, [70000, 1, 2], dtype="uint32")
    assert TokenFileDataset(path, 3, dtype="uint32")[0]["input_ids"].tolist() == [70000, 1, 2]
    with pytest.raises(ValueError, match="fewer than one sequence"):
        TokenFileDataset(path, 4, dtype="uint32")
    # worker processes re-open the file themselves (the memmap is not pickled)
    write_token_file(path, tokens)
    batches = list(DataLoader(TokenFileDataset(path, 10), batch_size=5, num_workers=2))
    assert torch.equal(torch.cat([b["input_ids"] for b in batches]).reshape(-1), tokens[:100])


def test_device_prefetcher_keeps_order_and_lookahead():
    pulled = []

    def source():
        for i in range(5):
            pulled.append(i)
            yield {"input_ids": torch.full((2, 4), i), "meta": [torch.tensor(i), "text"]}

    class Loader:
        sampler = "the-sampler"

        def __iter__(self):
            return source()

        def __len__(self):
            return 5

    pf = DevicePrefetcher(Loader(), "cpu", depth=2)
    assert len(pf) == 5 and pf.sampler == "the-sampler"
    seen = []
    for batch in pf:
        seen.append(int(batch["input_ids"][0, 0]))
        # when the caller gets batch i, batches i+1 and i+2 have been pulled from the loader already
        assert pulled[-1] == min(4, seen[-1] + 2)
        assert batch["meta"][1] == "text" and int(batch["meta"][0]) == seen[-1]
    assert seen == [0, 1, 2, 3, 4]
    assert pf.bytes_per_batch == 2 * 4 * 8 + 8
    assert list(DevicePrefetcher([], "cpu")) == []
    assert [int(b) for b in DevicePrefetcher([torch.tensor(7)], "cpu", depth=4)] == [7]
    assert [int(b) for b in pf.__class__([torch.tensor(1), torch.tensor(2)], "cpu", depth=1)] == [1, 2]


def run_sampler(rank, world_size, port, path):
    ctx = init_parallel_context(rank, world_size, port, 2, 1, 2)        # tp 2 x dp 2
    ds = TokenFileDataset(path, seq_len=4)
    loader = build_dataloader(ds, ctx, batch_size=2, shuffle=True)
    assert loader.sampler.num_replicas == 2 and loader.sampler.rank == ctx.get_local_rank(ParallelMode.DATA)
    per_epoch = []
    for epoch in range(2):
        loader.sampler.set_epoch(epoch)
        per_epoch.append(torch.cat([b["input_ids"] for b in loader]))
    mine = torch.stack(per_epoch)
    everyone = [torch.empty_like(mine) for _ in range(world_size)]
    torch.distributed.all_gather(everyone, mine)
    by_rank = {r: everyone[r] for r in range(world_size)}
    tp_peers = ctx.get_ranks_in_group(ParallelMode.TENSOR)
    dp_peers = ctx.get_ranks_in_group(ParallelMode.DATA)
    for r in tp_peers:          # one replica = one batch stream
        assert torch.equal(by_rank[r], mine)
    other = [r for r in dp_peers if r != rank][0]
    for e in range(2):          # replicas: disjoint halves that cover the dataset
        a, b = {tuple(x.tolist()) for x in mine[e]}, {tuple(x.tolist()) for x in by_rank[other][e]}
        assert not (a & b) and len(a | b) == len(ds)
    assert not torch.equal(mine[0], mine[1])                            # another permutation per epoch
    assert data_parallel_sampler(ds, ctx, shuffle=False).rank == ctx.get_local_rank(ParallelMode.DATA)
    ctx.destroy()


def test_build_dataloader_shards_over_the_data_group_only(tmp_path):
    path = str(tmp_path / "tokens.bin")
    write_token_file(path, torch.arange(64))
    spawn(run_sampler, world_size=4, path=path)


def test_trainer_sets_the_sampler_epoch():
    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
    from pipegoose_b200.optim import FusedAdam
    from pipegoose_b200.trainer import Trainer

    epochs = []

    class Sampler:
        def set_epoch(self, e):
            epochs.append(e)

    class Loader(list):
        sampler = Sampler()

    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=64, hidden_size=32, n_layer=1, n_head=4))
    data = Loader({"input_ids": torch.randint(0, 64, (2, 8))} for _ in range(2))
    Trainer(model, data, optim=FusedAdam(model.parameters(), lr=1e-2), num_epochs=3).fit()
    assert epochs == [0, 1, 2]


def run_resume_with_sampler(rank, world_size, port, path, ckp_dir):
    """Interrupted in the middle of the SECOND epoch of a shuffled, sharded loader; the resumed run must see the same
    batches in the same order as the uninterrupted one (sampler epoch + position restored)."""
    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
    from pipegoose_b200.nn import DataParallel, TensorParallel
    from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
    from pipegoose_b200.trainer import Callback, Trainer

    ctx = init_parallel_context(rank, world_size, port, 1, 1, 2)
    ds = TokenFileDataset(path, seq_len=8)                          # 24 sequences -> 12 per replica -> 6 batches per epoch

    class Seen(Callback):
        def __init__(self):
            self.batches = []

    def build(**kw):
        torch.manual_seed(0)
        model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=1, n_head=4))
        model = DataParallel(TensorParallel(model, ctx).parallelize(), ctx).parallelize()
        optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-2), ctx)
        loader = build_dataloader(ds, ctx, batch_size=2, shuffle=True)
        trainer = Trainer(model, loader, optim=optim, parallel_context=ctx, num_epochs=3, **kw)
        seen = []
        inner = trainer.train_step

        def spy(batch):
            seen.append(batch["input_ids"].clone())
            return inner(batch)

        trainer.train_step = spy
        return model, trainer, seen

    model_a, trainer_a, seen_a = build()
    assert trainer_a.fit().step == 18
    model_b, trainer_b, seen_b = build(checkpoint_dir=ckp_dir, checkpoint_every=3, max_steps=9)    # stops in epoch 1
    assert trainer_b.fit().step == 9
    model_c, trainer_c, seen_c = build(checkpoint_dir=ckp_dir, resume=True)
    assert trainer_c.fit().step == 18
    assert len(seen_b) == 9 and len(seen_c) == 9
    for got, want in zip(seen_b + seen_c, seen_a):
        assert torch.equal(got, want)
    for (n, a), (_, c) in zip(model_a.named_parameters(), model_c.named_parameters()):
        assert torch.allclose(a, c, atol=1e-6), n
    ctx.destroy()


def test_trainer_resumes_a_shuffled_sharded_loader_mid_epoch(tmp_path):
    path = str(tmp_path / "tokens.bin")
    write_token_file(path, torch.randint(0, 96, (24 * 8,), generator=torch.Generator().manual_seed(3)))
    spawn(run_resume_with_sampler, world_size=2, path=path, ckp_dir=str(tmp_path / "run"))
