"""CPU, world_size 2 over gloo: the host side of the multi-GPU paths — bucket planning + the
bucket walk of the data-parallel gradient all-reduce, and the shard / gather logic of extraction."""
import os
import socket

import numpy as np
import pandas as pd
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeGradBuffer:
    """Same interface as training.GradBuffer (entries / offsets / flat), CPU memory."""

    def __init__(self, sizes):
        self.entries = [n for n, _ in sizes]
        self.offsets, total = {}, 0
        for n, k in sizes:
            self.offsets[n] = (total, k, (k,))
            total += (k + 7) // 8 * 8
        self.flat = torch.zeros(total)


class _FakeModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(5))


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from stonkgs_b200.dp import DataParallel
        model = _FakeModel()
        with torch.no_grad():
            model.w.fill_(float(rank + 1))
        # device-side state derived from the weights BEFORE the broadcast (an optimizer or a forward built it): it must
        # be dropped, because the broadcast writes through .data and does not bump Parameter._version
        model._dev_state = {"bf16 copies of rank-local weights": rank}
        model._special_rows_version = 7
        dp = DataParallel(model, bucket_mb=4096 * 4 / (1 << 20), wire_dtype=torch.float32)   # 4096-element buckets
        assert torch.equal(model.w.data, torch.ones(5))                                     # rank 0's weights everywhere
        assert model._dev_state is None and model._special_rows_version is None
        sizes = [("w_ent", 10000), ("w_text", 3000), ("t_w", 700), ("t_b", 2), ("l1.w2", 2500), ("l1.b2", 3),
                 ("l0.w2", 2500), ("emb_b", 7)]
        gb = _FakeGradBuffer(sizes)
        g = torch.Generator().manual_seed(100 + rank)
        local = torch.randn(gb.flat.numel(), generator=g)
        gb.flat.copy_(local)
        dp.begin(gb)
        names = [b.names for b in dp.buckets]
        assert names[0] == ["w_ent"] and sum(len(n) for n in names) == len(sizes)          # big segment ships alone, first
        assert dp.buckets[0].start == 0 and all(a.end == b.start for a, b in zip(dp.buckets, dp.buckets[1:]))
        for n, _ in sizes:
            dp.on_ready(n)
        dp.finish(gb)
        both = [torch.randn(gb.flat.numel(), generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
        torch.testing.assert_close(gb.flat, sum(both) / world)
        # no_sync: nothing is reduced
        gb.flat.copy_(local)
        with dp.no_sync():
            dp.begin(gb)
            for n, _ in sizes:
                dp.on_ready(n)
            dp.finish(gb)
        assert torch.equal(gb.flat, local)
        # overlap=False: the same buckets are reduced in finish(), after the whole backward has been enqueued
        dp2 = DataParallel(model, bucket_mb=4096 * 4 / (1 << 20), wire_dtype=torch.float32, overlap=False)
        gb.flat.copy_(local)
        dp2.begin(gb)
        for n, _ in sizes:
            dp2.on_ready(n)
        assert torch.equal(gb.flat, local)                     # nothing has moved yet
        dp2.finish(gb)
        torch.testing.assert_close(gb.flat, sum(both) / world)
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


def _embed_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from stonkgs_b200.embeddings import get_stonkgs_embeddings
        n = 11
        ids = np.arange(n * 512, dtype=np.int64).reshape(n, 512)
        df = pd.DataFrame({"input_ids": list(ids), "attention_mask": list(np.ones_like(ids)),
                           "token_type_ids": list(np.zeros_like(ids))}, index=[f"row{i}" for i in range(n)])
        seen = []

        def fake_embed(i, m, t):
            seen.append(i[:, 0] // 512)
            return np.repeat((i[:, :1] // 512).astype(np.float32), 768, axis=1)   # embedding = row number

        res = get_stonkgs_embeddings(df, _embed_fn=fake_embed)
        mine = np.concatenate(seen)
        lo, hi = (0, 6) if rank == 0 else (6, 11)
        assert list(mine) == list(range(lo, hi))                                   # contiguous shard per rank
        got = np.asarray(res["embedding"].tolist())
        assert got.shape == (n, 768) and list(got[:, 0]) == list(range(n)) and list(res.index) == list(range(n))   # fresh RangeIndex like the reference
        # list_of_indices are POSITIONS (reference: preprocessed_df.iloc[idx]) whatever the frame's own index is
        some = get_stonkgs_embeddings(df, list_of_indices=[7, 2, 9, 4], _embed_fn=fake_embed)
        assert list(np.asarray(some["embedding"].tolist())[:, 0]) == [7, 2, 9, 4] and list(some.index) == [0, 1, 2, 3]
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("worker", [_dp_worker, _embed_worker])
def test_world_size_2_gloo(worker):
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert dict(out) == {0: "ok", 1: "ok"}
