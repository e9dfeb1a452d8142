"""N>1 host logic on CPU: world_size-2 gloo processes shard units, combine timings (MAX) and units (SUM), gather logits."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fhe_linformer_b200 import shard


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = list(shard.my_units(total, rank, world))
    seconds = 1.0 + rank          # rank 1 is the slow one
    tot, worst, rate = shard.combine(len(mine), seconds)
    logits = shard.gather_logits(np.arange(20) + 100 * rank)
    q.put((rank, mine, tot, worst, rate, [l.tolist() for l in logits]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_and_combine():
    world, total = 2, 257
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs: p.join(timeout=60)
    units = sorted(u for r in res for u in r[1])
    assert units == list(range(total))                      # every unit exactly once
    assert abs(len(res[0][1]) - len(res[1][1])) <= 1
    for r in res:
        assert r[2] == total and r[3] == 2.0 and abs(r[4] - total / 2.0) < 1e-12   # SUM of units, MAX of time
        assert r[5][0][:3] == [0, 1, 2] and r[5][1][:3] == [100, 101, 102]


def test_partition_edges():
    assert list(shard.my_units(5, 0, 8)) == [0] and list(shard.my_units(5, 7, 8)) == []
    assert [len(shard.my_units(256, r, 8)) for r in range(8)] == [32] * 8
    assert sum(len(shard.my_units(1000, r, 3)) for r in range(3)) == 1000
    tot, worst, rate = shard.combine(10, 2.0)                # no process group: identity
    assert (tot, worst, rate) == (10.0, 2.0, 5.0)


def test_limb_ranges_of_the_sharded_key_switch_cover_the_extended_basis():
    """Host logic of sharded.py without a GPU: every limb of Q_l u P is owned by exactly one rank, for ragged splits too."""
    from types import SimpleNamespace
    from fhe_linformer_b200 import sharded
    eng = SimpleNamespace(K=7, N=16, alpha=7)
    for l in (28, 13, 2, 1):
        for world in (1, 2, 3, 4, 8):
            owned = []
            for r in range(world):
                st = sharded._RankState(eng, l, r, world, torch.device("cpu"))
                owned += [t for first, count in st.ranges() for t in range(first, first + count)]
            assert sorted(owned) == list(range(l + 7)), (l, world)


def _gather_worker(rank, world, port, q):
    """The exchange of the limb-sharded key switch on CPU tensors: DistComm.all_gather of padded shares + sharded.assemble."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fhe_linformer_b200 import sharded
    comm = sharded.DistComm()
    ok = True
    for total, lead in ((7, 2), (28, 2), (5, 1), (1, 2)):            # K special limbs / l digit limbs over the ranks, ragged splits included
        sizes, pad = sharded.share_sizes(total, world)
        mine = shard.my_units(total, rank, world)
        full = torch.arange(lead * total * 4, dtype=torch.int64).reshape(lead, total, 4) * 7 + 3     # what every rank must end with
        share = torch.zeros((lead, pad, 4), dtype=torch.int64)
        share[:, :len(mine)] = full[:, mine.start:mine.stop]
        got = sharded.assemble(comm.all_gather(share), total, world, 1)
        ok = ok and bool((got == full).all())
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_of_limb_shares_with_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs: p.join(timeout=60)
    assert res == [(0, True), (1, True)]
