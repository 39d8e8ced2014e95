"""CPU, world_size 2 (gloo): the multi-GPU plan of SURVEY.md 8(e) -- stripe the list entries over the
ranks, every rank answers all queries on its stripe, all-gather the per-rank top-k, merge with id
de-duplication -- reproduces the unsharded result. The per-rank search and the merge are done by the
oracle here (no GPU in this container); the striping and gather layout are the product's
(lira_ann_search_b200.parallel)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from helpers import merge_ranks_numpy as _merge_ranks_numpy, synth
    from lira_ann_search_b200.parallel import stripe_assignment
    O.set_num_threads(2)
    x_d, x_q = synth(3000, 16, 40, seed=5, integer=True)
    B, k = 12, 10
    rng = np.random.RandomState(3)
    d2b = np.stack([rng.randint(0, B, len(x_d)), np.where(rng.rand(len(x_d)) < 0.4, rng.randint(0, B, len(x_d)), -1)], 1)
    poff = np.arange(len(x_q) + 1, dtype=np.int64) * 5
    pids = np.concatenate([rng.choice(B, 5, replace=False) for _ in range(len(x_q))]).astype(np.int32)
    off, ids, vecs = O.build_lists_from_data_2_bkt(x_d, stripe_assignment(d2b, B, rank, world), B)
    I_loc, D_loc, cmp_loc = O.search(off, ids, vecs, x_q, poff, pids, k, O.L2, O.F64, 1)
    gD = [torch.empty(D_loc.shape, dtype=torch.float32) for _ in range(world)]
    gI = [torch.empty(I_loc.shape, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gD, torch.from_numpy(D_loc))
    dist.all_gather(gI, torch.from_numpy(I_loc))
    cmp_t = torch.from_numpy(cmp_loc.copy())
    dist.all_reduce(cmp_t)
    D, I = _merge_ranks_numpy(np.stack([t.numpy() for t in gD]), np.stack([t.numpy() for t in gI]), k)
    off0, ids0, vecs0 = O.build_lists_from_data_2_bkt(x_d, d2b, B)
    I_ref, D_ref, cmp_ref = O.search(off0, ids0, vecs0, x_q, poff, pids, k, O.L2, O.F64, 1)
    ok = bool(np.array_equal(I, I_ref) and np.array_equal(D, D_ref) and np.array_equal(cmp_t.numpy(), cmp_ref))
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()


def test_striped_search_allgather_merge_world2():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


def test_stripes_partition_every_list():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from lira_ann_search_b200.parallel import stripe_assignment, stripe_csr
    rng = np.random.RandomState(0)
    N, B = 4000, 9
    d2b = np.stack([rng.randint(0, B, N), np.where(rng.rand(N) < 0.3, rng.randint(0, B, N), -1)], 1).astype(np.int32)
    d2b[:50, 1] = d2b[:50, 0]  # same bucket twice: collapses like std::unique (search.cpp:383-386)
    x = rng.randn(N, 4).astype(np.float32)
    off, ids, _ = O.build_lists_from_data_2_bkt(x, d2b, B)
    for world in (2, 3, 8):
        parts = [O.build_lists_from_data_2_bkt(x, stripe_assignment(d2b, B, r, world), B) for r in range(world)]
        for b in range(B):
            merged = np.sort(np.concatenate([p[1][p[0][b]:p[0][b + 1]] for p in parts]))
            assert np.array_equal(merged, ids[off[b]:off[b + 1]])
            sizes = [p[0][b + 1] - p[0][b] for p in parts]
            assert max(sizes) - min(sizes) <= 1
        o2 = [stripe_csr(off, ids, r, world) for r in range(world)]
        assert sum(int(o[0][-1]) for o in o2) == off[-1]
        for b in range(B):
            merged = np.sort(np.concatenate([o[1][o[0][b]:o[0][b + 1]] for o in o2]))
            assert np.array_equal(merged, ids[off[b]:off[b + 1]])
    # empty lists, also trailing ones, are legal (the reference skips empty buckets)
    off_e = np.array([0, 3, 5, 5, 5], np.int64)
    ids_e = np.arange(5, dtype=np.int32)
    for world in (1, 2, 3):
        got = [stripe_csr(off_e, ids_e, r, world) for r in range(world)]
        assert all(len(o[0]) == 5 and o[0][-1] == len(o[1]) for o in got)
        assert np.array_equal(np.sort(np.concatenate([o[1] for o in got])), ids_e)
        assert all(o[0][3] == o[0][2] == o[0][4] for o in got)
