#!/usr/bin/env python
"""One GPU's share of config 5 (BigANN-100M shape sharded over 8 GPUs): N = 12.5 M base vectors x 128, full 2x
redundancy (every vector in its two nearest partitions, LIRA_largescale.py:320-329) = 25 M list entries, B = 1024,
10 000 queries, k = 10. Everything is generated and indexed on the device (plain torch, untimed plumbing); the timed
part is the library's grouped list scan with explicit probe sets (nprobe nearest centroids, the paper's IVF baseline).
Checks the ids of a query sample against a brute-force scan of the probed entries and prints one JSON line.

    python tools/scale_test.py [--N 12500000] [--Q 10000] [--B 1024] [--nprobe 16] [--steps 5]

Under torchrun (WORLD_SIZE > 1) the entries of every list are striped over the ranks (entry j of a list -> rank j mod world,
SURVEY.md 8e): every rank builds the same dataset, keeps its stripe, answers every query on it, and the per-rank top-k lists
are merged by an NCCL all-gather + lira_merge_ranks_dev; the merged result is checked against the brute-force scan.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/scale_test.py --N 8000000
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lira_ann_search_b200 as L

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=12_500_000)
ap.add_argument("--d", type=int, default=128)
ap.add_argument("--Q", type=int, default=10_000)
ap.add_argument("--B", type=int, default=1024)
ap.add_argument("--nprobe", type=int, default=16)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--check", type=int, default=64)
ap.add_argument("--real", action="store_true", help="real-valued unit vectors (DEEP style) instead of small integers: approximate tensor-core mode")
ap.add_argument("--compare-cuda-cores", action="store_true", help="also time the fp32 CUDA-core scan")
args = ap.parse_args()
L._cabi.require_gpu()
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
if world > 1:
    import torch.distributed as tdist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    tdist.init_process_group("nccl", device_id=dev)
N, d, Q, B, k = args.N, args.d, args.Q, args.B, args.k
t0 = time.time()
g = torch.Generator(device=dev).manual_seed(43)
ncomp = 4096
centres = torch.randn(ncomp, d, device=dev, generator=g)
logw = 0.5 * torch.randn(ncomp, device=dev, generator=g)
w = torch.softmax(logw, 0)


def gen(n, seed):
    gg = torch.Generator(device=dev).manual_seed(seed)
    comp = torch.multinomial(w, n, replacement=True, generator=gg)
    x = centres[comp] + 0.35 * torch.randn(n, d, device=dev, generator=gg)
    if args.real:
        return x / x.norm(dim=1, keepdim=True)             # unit vectors, like DEEP
    return torch.clamp(torch.round(32 * x + 64), 0, 255)   # integer valued, like SIFT / BigANN


CH = 1_000_000 if d <= 256 else 125_000
x_d = torch.empty((N, d), dtype=torch.float32, device=dev)
for c, s in enumerate(range(0, N, CH)):
    x_d[s:s + CH] = gen(min(CH, N - s), 43 * 1_000_003 + c)
x_q = gen(Q, 50)
# centroids: a few Lloyd iterations on a sample (plumbing)
samp = x_d[torch.randperm(N, device=dev, generator=g)[:262_144]]
cent = samp[:B].clone()
for _ in range(8):
    a = torch.cat([torch.cdist(samp[i:i + 32768], cent).argmin(1) for i in range(0, len(samp), 32768)])
    s = torch.zeros_like(cent).index_add_(0, a, samp)
    n = torch.bincount(a, minlength=B).clamp(min=1).unsqueeze(1)
    cent = s / n
# two nearest partitions per vector
b2 = torch.empty((N, 2), dtype=torch.int64, device=dev)
ACH = max(16384, (1 << 28) // B)   # rows per assignment step: the [rows, B] distance matrix stays near 1 GiB
for s in range(0, N, ACH):
    b2[s:s + ACH] = torch.cdist(x_d[s:s + ACH], cent).topk(2, dim=1, largest=False).indices
ids = torch.arange(N, device=dev).repeat_interleave(2)
key = b2.reshape(-1) * N + ids
order = torch.argsort(key)
ids = ids[order].to(torch.int32)
sizes = torch.bincount(b2.reshape(-1), minlength=B)
off = torch.zeros(B + 1, dtype=torch.int64, device=dev)
off[1:] = torch.cumsum(sizes, 0)
del key, order, b2
E = int(off[-1])
vecs = torch.empty((E, d), dtype=torch.float32, device=dev)
for s in range(0, E, 4 * CH):
    vecs[s:s + 4 * CH] = x_d[ids[s:s + 4 * CH].long()]
del x_d
torch.cuda.synchronize()
print(f"[scale] data + lists on the device: {time.time() - t0:.1f}s; E = {E} entries, list sizes {int(sizes.min())}..{int(sizes.max())}",
      file=sys.stderr, flush=True)
if world > 1:
    # this rank's stripe of every list: entries at positions j = rank (mod world) of the list
    pos = torch.arange(E, device=dev) - torch.repeat_interleave(off[:-1], sizes)
    mine = (pos % world) == rank
    my_vecs, my_ids = vecs[mine].contiguous(), ids[mine].contiguous()
    my_sizes = torch.clamp((sizes - rank + world - 1) // world, min=0)
    my_off = torch.zeros(B + 1, dtype=torch.int64, device=dev)
    my_off[1:] = torch.cumsum(my_sizes, 0)
    del pos, mine
else:
    my_vecs, my_ids, my_off = vecs, ids, off
index = L.LiraIndex.from_device(my_vecs, my_ids, my_off.cpu().numpy(), d, "L2", device=local)
expect_mode = "none" if d > 1024 else ("approximate" if args.real else "exact")   # d > 1024: the fp32 CUDA-core scan answers
assert index.tensor_core_mode == expect_mode
# probe sets: nprobe nearest centroids
pids = torch.cdist(x_q, cent).topk(args.nprobe, dim=1, largest=False).indices.to(torch.int32).reshape(-1).contiguous()
poff = (torch.arange(Q + 1, device=dev, dtype=torch.int64) * args.nprobe).contiguous()
index.set_timing(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def answer():
    D, I, cmp_ = index.search_dev(x_q, poff, pids, k)
    if world > 1:
        from lira_ann_search_b200.parallel import allgather_merge
        D, I = allgather_merge(D, I, k, "L2", dedup=True, device=local)
    return D, I, cmp_


for _ in range(3):
    D, I, cmp_ = answer()
torch.cuda.synchronize()
assert index.last_path == ("cuda-core" if expect_mode == "none" else "tensor-core")
ms, scan_ms, scan_bytes, scan_pairs = [], [], [], []
for _ in range(args.steps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        tdist.barrier()
    e0.record()
    D, I, cmp_ = answer()
    e1.record()
    e1.synchronize()
    ms.append(e0.elapsed_time(e1))
    tm = index.last_timing()
    scan_ms.append(tm["scan_ms"]); scan_bytes.append(tm["scan_bytes"]); scan_pairs.append(tm["scan_pairs"])
# check a sample against a brute-force scan of the probed entries (distinct ids, ties by id)
Ih, Dh = I.cpu().numpy(), D.cpu().numpy()
offh = off.cpu().numpy()
bad = 0
for qi in range(min(args.check, Q)):
    lists = pids[qi * args.nprobe:(qi + 1) * args.nprobe].cpu().numpy()
    rows = torch.cat([torch.arange(offh[b], offh[b + 1], device=dev) for b in lists])
    dist = ((vecs[rows] - x_q[qi]) ** 2).sum(1)
    gid = ids[rows].long()
    gid_h, dist_h = gid.cpu().numpy(), dist.cpu().numpy()
    o = np.lexsort((gid_h, dist_h))   # by (distance, id)
    gid_s, dist_s = gid_h[o], dist_h[o]
    _, first = np.unique(gid_s, return_index=True)
    keep = np.sort(first)[:k]
    if args.real:   # ids may differ at fp32 ties only: compare the distances rank by rank
        if not np.allclose(dist_s[keep], Dh[qi], rtol=1e-5, atol=1e-6):
            bad += 1
    elif not (np.array_equal(gid_s[keep], Ih[qi]) and np.array_equal(dist_s[keep], Dh[qi])):
        bad += 1
simt_ms = None
if args.compare_cuda_cores:
    index.set_use_tensor_cores(False)
    D2, I2, _ = index.search_dev(x_q, poff, pids, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    D2, I2, _ = index.search_dev(x_q, poff, pids, k)
    e1.record()
    e1.synchronize()
    simt_ms = e0.elapsed_time(e1)
    same_ids = float((I2 == I).all(1).float().mean())
    index.set_use_tensor_cores(True)
if world > 1:   # the slowest rank's time counts; every rank scanned its stripe
    t = torch.tensor([float(np.mean(ms))], device=dev)
    tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
    ms = [float(t.item())]
    c = cmp_.double().clone()
    tdist.all_reduce(c)
    cmp_ = c
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
s_ms = float(np.mean(scan_ms))
if rank == 0:
  print(json.dumps({"n_gpus": world, "sharding": "list entries striped over the ranks + NCCL all-gather + merge" if world > 1 else "none", "workload": "unit vectors, full 2x redundancy (DEEP style)" if args.real else "config 5 shard (1/8 of BigANN-100M shape, full 2x redundancy)", "N": N, "entries": E, "Q": Q, "B": B,
                  "nprobe": args.nprobe, "k": k, "ms_per_batch": float(np.mean(ms)), "qps": Q / (float(np.mean(ms)) * 1e-3),
                  "scan_kernel_ms": s_ms, "scan_algorithmic_bytes": float(np.mean(scan_bytes)),
                  "scan_gbs_algorithmic": float(np.mean(scan_bytes)) / (s_ms * 1e-3) / 1e9,
                  "roofline_frac_of_measured_hbm": float(np.mean(scan_bytes)) / (s_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                  "scan_tflops_algorithmic": 2.0 * d * float(np.mean(scan_pairs)) / (s_ms * 1e-3) / 1e12,
                  "tensor_frac_of_measured_bf16_sustained": 2.0 * d * float(np.mean(scan_pairs)) / (s_ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"],
                  "mean_entries_scanned_per_query": float(cmp_.double().mean()),
                  "checked_queries": min(args.check, Q), "mismatches": bad, "redo": index.last_redo,
                  "tensor_core_mode": index.tensor_core_mode,
                  "cuda_core_ms_per_batch": simt_ms, "rows_with_identical_ids_vs_cuda_cores": same_ids if simt_ms else None}), flush=True)
if world > 1:
    tdist.destroy_process_group()
