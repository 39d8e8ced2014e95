"""Multi-GPU query phase: one process per GPU, list ENTRIES striped across the ranks,
queries and the probing model replicated. Each rank answers every query on its stripe; the only
exchange is an all-gather of the per-rank top-k lists ((score, id) packed in 64-bit keys, Q*k*8 bytes
per rank) followed by a k-way merge with id de-duplication (the two copies of a redundantly stored
vector can sit on different GPUs). The reference has no multi-process code; this is new work.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi as C


def stripe_assignment(data_2_bkt, n_bkt, rank, world):
    """Per-rank view of data_2_bkt[N, n_mul]: entry j (in the bucket's sorted-unique id order of
    search.cpp:380-386) of every bucket belongs to rank j % world; everything else becomes -1.
    The union over ranks is exactly the original assignment, stripes are disjoint and balanced to
    within one entry per bucket for any probe pattern."""
    d2b = np.asarray(data_2_bkt)
    n, n_mul = d2b.shape
    flat_b = d2b.reshape(-1).astype(np.int64)
    flat_i = np.repeat(np.arange(n, dtype=np.int64), n_mul)
    valid = flat_b >= 0
    if np.any(flat_b[valid] >= n_bkt):
        raise ValueError("bucket id out of range.")
    key = flat_b * n + flat_i
    key[~valid] = np.iinfo(np.int64).max
    order = np.argsort(key, kind="stable")
    sk = key[order]
    first = np.ones(len(sk), bool)
    first[1:] = sk[1:] != sk[:-1]  # duplicates of (bucket, id) collapse like std::unique
    first &= sk != np.iinfo(np.int64).max
    sb = np.where(first, sk // n, -1)
    # position of each kept entry inside its bucket
    counts = np.bincount(sb[first], minlength=n_bkt)
    starts = np.zeros(n_bkt + 1, np.int64)
    np.cumsum(counts, out=starts[1:])
    pos = np.full(len(sk), -1, np.int64)
    pos[first] = np.arange(int(first.sum())) - starts[sb[first]]
    keep_sorted = first & (pos % world == rank)
    keep = np.zeros(len(sk), bool)
    keep[order] = keep_sorted
    out = np.where(keep.reshape(n, n_mul), d2b, -1)
    return out.astype(d2b.dtype)


def stripe_csr(list_offsets, list_ids, rank, world):
    """Same striping for an explicit CSR (cluster_ids order): returns (offsets, ids) of this rank."""
    off = np.asarray(list_offsets, np.int64)
    ids = np.asarray(list_ids)
    B = len(off) - 1
    lst = np.repeat(np.arange(B), np.diff(off))      # list of every entry (empty lists, also trailing ones, simply do not occur)
    pos = np.arange(off[-1]) - off[:-1][lst]
    keep = pos % world == rank
    sizes = np.bincount(lst[keep], minlength=B).astype(np.int64)
    new_off = np.zeros(B + 1, np.int64)
    np.cumsum(sizes, out=new_off[1:])
    return new_off, ids[keep]


_BUF = {}


def _buffers(Q, k, world, device):
    """Grow-only scratch of the merge (keys of this rank, gathered keys) per (shape, device): a step of the serving loop
    allocates nothing."""
    import torch
    key = (Q, k, world, str(device))
    b = _BUF.get(key)
    if b is None:
        b = (torch.empty((Q, k), dtype=torch.int64, device=device), torch.empty((world, Q, k), dtype=torch.int64, device=device))
        _BUF[key] = b
    return b


def pack_keys(D, I, metric, device, out=None):
    """(D[Q,k] f32, I[Q,k] i64) torch CUDA tensors -> int64 tensor of ordered (score, id) keys."""
    import torch
    keys = torch.empty(D.shape, dtype=torch.int64, device=D.device) if out is None else out
    from .engine import _stream_handle
    st = _stream_handle(D.device)
    m = C.METRIC_IP if str(metric).lower() in ("1", "ip", "inner_product") else C.METRIC_L2
    C.check(C.lib().lira_pack_keys_dev(D.data_ptr(), I.data_ptr(), D.numel(), m, keys.data_ptr(), device, st))
    return keys


def allgather_merge(D, I, k, metric="L2", dedup=True, device=0, group=None, out=None):
    """Per-rank (D, I) -> global top-k on every rank: NCCL all-gather of the packed keys + merge kernel. `out` = (D, I) CUDA
    tensors of an earlier call are filled in place."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    Q = D.shape[0]
    mine, gathered = _buffers(Q, D.shape[1], world, D.device)
    keys = pack_keys(D.contiguous(), I.contiguous(), metric, device, out=mine)
    dist.all_gather_into_tensor(gathered, keys, group=group)
    if out is None:
        out = (torch.empty((Q, k), dtype=torch.float32, device=keys.device), torch.empty((Q, k), dtype=torch.int64, device=keys.device))
    D_out, I_out = out
    from .engine import _stream_handle
    st = _stream_handle(keys.device)
    m = C.METRIC_IP if str(metric).lower() in ("1", "ip", "inner_product") else C.METRIC_L2
    C.check(C.lib().lira_merge_ranks_dev(gathered.data_ptr(), world, Q, int(k), m, int(bool(dedup)),
                                         D_out.data_ptr(), I_out.data_ptr(), device, st))
    return D_out, I_out
