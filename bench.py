#!/usr/bin/env python
"""bench.py -- QPS of the LIRA query phase at recall@10 >= 0.95 (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                      # this repo's CUDA path, config 1
    torchrun --nproc-per-node N bench.py --gpus N --steps K --warmup W  # N > 1: the sharded 100M-vector shape
    python bench.py --impl reference --gpus N --steps K --warmup W      # the reference's own CPU search.cpp
    python bench.py --prepare [--gpus N]                                # build the workload + operating point only

N = 1, `config.workload = "sift1m-shape"` (BASELINE.json configs[0], the configuration the metric is quoted on):
1M x 128 fp32 integer-valued Gaussian-mixture base (synthetic: no datasets on the box), 10k queries, B = 1024 K-Means
partitions, LIRA probing model trained here with the reference loop shape (BCELoss + Adam), 3 % learned redundancy
(n_mul = 2), k = 10, L2. One step = the whole query phase for the 10k-query batch: centroid features -> MLP ->
threshold select -> grouped list scan -> dedup merge.

N > 1, default, `config.workload = "bigann-shape-sharded"` (BASELINE.json configs[4]: 100 M x 128 vectors, every one stored in
its two nearest partitions = 200 M list entries, split over the ranks: STRONG scaling, queries/s should grow with N): every
rank generates only its own 100 M / N vectors on the device (chunk-seeded generator), so every one of the B = 1024 lists is
striped across the ranks by vector id. Every rank answers all 10k queries on its stripe; the per-rank top-k lists are
all-gathered over NCCL (packed 64-bit keys) and merged with id de-duplication by lira_merge_ranks_dev -- the collective and
the merge are inside the timed region. The line also carries the time one rank needs for its stripe alone
(`share_alone_ms`: what the exchange step adds is ms_per_step - share_alone_ms). `--share S` keeps S vectors per rank instead
(the dataset grows with N: weak scaling of the dataset). `--shard queries` (index replicas of config 1, no collective) and
`--workload sift1m --shard lists` (config 1 striped) are kept behind flags.

`--workload knn` (one GPU): BASELINE.json configs[1], compute_knn's exact ground truth on SIFT1M shape (1 M x 128 base, 10k
queries, k = 100) through lira_knn_search_dev, plus the reference program's own use (self-kNN, k + 1 = 11) on a sample.

The reference arm never imports this repo's package or loads liblira_b200.so: it reads the cached workload and operating
point (built in a separate `--prepare` process when missing) and times oracle/_ref/search_ref, the reference's unmodified
search.cpp, on a bounded sample that both arms state in the same `config` dict.
"""
import argparse
import hashlib
import json
import os
import re
import socket
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CACHE = os.environ.get("LIRA_BENCH_CACHE", "/tmp/lira_bench_cache")
SEED = 43
GEN = {"kind": "gaussian mixture, integer valued", "components": 4096, "weight_lognormal_sigma": 0.5, "sigma": 0.8,
       "affine": "clip(round(16 x + 100), 0, 255)", "seed": SEED}
MLP_KEYS = ("distance_net.0", "distance_net.2", "vector_net.0", "vector_net.2", "fc.0", "fc.2")
REF_SAMPLE = int(os.environ.get("LIRA_REF_SAMPLE", "1000"))   # queries per step of the reference arm (config 1)
SHARE = 12_500_000        # vectors per rank of the sharded workload in its weak-scaling form (--share)
TOTAL = 100_000_000       # the BigANN-100M-shape dataset (BASELINE.json configs[4]): N > 1 default, split over the ranks
CHUNK = 500_000           # rows per generator chunk (chunk c has its own Philox seed: any rank can produce any chunk)
REF_SHARD_ROWS = 1_000_000   # reference arm at N > 1: search.cpp runs on rank 0's first two chunks
REF_SHARD_QUERIES = 200


# ---------------------------------------------------------------------------------------------
# build-side plumbing (untimed; plain torch on the GPU -- shared by both arms, no library code)
# ---------------------------------------------------------------------------------------------
def make_mlp(B, d):
    """Same architecture and parameter names as the reference's MLP_2_Input (model_probing.py:5-39)."""
    import torch.nn as nn
    import torch

    def two(n_in, n_h, n_out, last):
        return nn.Sequential(nn.Linear(n_in, n_h), nn.ReLU(), nn.Linear(n_h, n_out), last)

    class MLP_2_Input(nn.Module):
        def __init__(self):
            super().__init__()
            self.distance_net = two(B, 128, 64, nn.ReLU())
            self.vector_net = two(d, 128, 64, nn.ReLU())
            self.fc = two(128, 128, B, nn.Sigmoid())

        def forward(self, x_dist, x_vec):
            return self.fc(torch.cat((self.distance_net(x_dist), self.vector_net(x_vec)), dim=1))

    return MLP_2_Input()


def kmeans_assign(x_t, c_t, chunk=262144):
    import torch
    out = torch.empty(x_t.shape[0], dtype=torch.int64, device=x_t.device)
    c2 = (c_t * c_t).sum(1)[None, :]
    for a in range(0, x_t.shape[0], chunk):
        xb = x_t[a:a + chunk]
        out[a:a + chunk] = ((xb * xb).sum(1)[:, None] + c2 - 2.0 * xb @ c_t.T).argmin(1)
    return out


def kmeans_train(x_t, k, niter=20, seed=1234):
    """Lloyd on at most 256 k sampled points (faiss.Kmeans shape: utils.py:321-330)."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    n = x_t.shape[0]
    if n > 256 * k:
        x_t = x_t[torch.randperm(n, generator=g)[:256 * k].to(x_t.device)]
        n = x_t.shape[0]
    c = x_t[torch.randperm(n, generator=g)[:k].to(x_t.device)].clone()
    for _ in range(niter):
        a = kmeans_assign(x_t, c)
        cnt = torch.bincount(a, minlength=k).float()
        s = torch.zeros_like(c).index_add_(0, a, x_t)
        nz = cnt > 0
        c[nz] = s[nz] / cnt[nz, None]
        if (~nz).any():
            idx = torch.randint(0, n, (int((~nz).sum()),), generator=g).to(x_t.device)
            c[~nz] = x_t[idx]
    return c


def mixture(d, dev):
    import torch
    g = torch.Generator(device=dev).manual_seed(SEED * 1_000_003)
    centres = torch.randn(GEN["components"], d, generator=g, device=dev)
    w = torch.exp(GEN["weight_lognormal_sigma"] * torch.randn(GEN["components"], generator=g, device=dev))
    return g, centres, w


def draw(m, centres, w, g):
    import torch
    c = torch.multinomial(w, m, replacement=True, generator=g)
    x = centres[c] + GEN["sigma"] * torch.randn(m, centres.shape[1], generator=g, device=centres.device)
    return torch.clamp(torch.round(16 * x + 100), 0, 255)


def knn_torch(qs, x_d, kk):
    import torch
    out = torch.empty(qs.shape[0], kk, dtype=torch.int64, device=qs.device)
    bn = (x_d * x_d).sum(1)
    for a in range(0, qs.shape[0], 2048):
        qb = qs[a:a + 2048]
        dist = bn[None, :] - 2.0 * qb @ x_d.T  # + |q|^2, constant per row (integer data: exact in fp32)
        out[a:a + 2048] = dist.topk(kk, largest=False).indices
    return out


def train_probing_model(x_d, cent, assign, tr_idx, knn_tr, dev, log, t0):
    """Scaler over all rows of x_d + the MLP trained on (features, vector) -> partitions holding a 10-NN
    (LIRA_smallscale.py:299-329 shape: BCELoss, Adam, batches of 512)."""
    import torch
    N, B, d = x_d.shape[0], cent.shape[0], x_d.shape[1]
    n_tr = tr_idx.shape[0]
    labels = torch.zeros(n_tr, B, device=dev)
    labels.scatter_(1, assign[knn_tr], 1.0)
    s1 = torch.zeros(B, dtype=torch.float64, device=dev)
    s2 = torch.zeros(B, dtype=torch.float64, device=dev)
    for a in range(0, N, 65536):
        f = torch.cdist(x_d[a:a + 65536], cent).double()
        s1 += f.sum(0)
        s2 += (f * f).sum(0)
    mean = s1 / N
    scale = torch.sqrt(torch.clamp(s2 / N - mean * mean, min=0))
    scale[scale == 0] = 1.0
    mean32, scale32 = mean.float(), scale.float()
    torch.manual_seed(SEED)
    model = make_mlp(B, d).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    crit = torch.nn.BCELoss()
    xf = (torch.cdist(x_d[tr_idx], cent) - mean32) / scale32
    xv = x_d[tr_idx]
    for epoch in range(40):
        perm = torch.randperm(n_tr, device=dev)
        for a in range(0, n_tr, 512):
            idx = perm[a:a + 512]
            opt.zero_grad()
            loss = crit(model(xf[idx], xv[idx]), labels[idx])
            loss.backward()
            opt.step()
    log(f"[bench] probing model trained (last loss {loss.item():.4f}): {time.time() - t0:.1f}s")
    return model.eval(), mean32, scale32


def weights_of(model):
    sd = model.state_dict()
    out = []
    for kk in MLP_KEYS:
        out.append(sd[kk + ".weight"].float().cpu().numpy())
        out.append(sd[kk + ".bias"].float().cpu().numpy())
    return out


def make_workload(N=1_000_000, d=128, Q=10_000, B=1024, k=10, redundancy_ratio=0.03, dev="cuda:0", log=print):
    """Config 1 (cached on disk). Returns (dict of numpy arrays, cache path)."""
    import torch
    tag = f"sift1m_N{N}_d{d}_Q{Q}_B{B}_k{k}_r{redundancy_ratio}_s{SEED}_v4"
    path = os.path.join(CACHE, tag)
    names = ["x_d", "x_q", "gt", "centroids", "scaler_mean", "scaler_scale", "data_2_bkt"] + [f"mlp_{i}" for i in range(12)]
    if all(os.path.exists(os.path.join(path, n + ".npy")) for n in names):
        log(f"[bench] workload cache hit: {path}")
        return {n: np.load(os.path.join(path, n + ".npy")) for n in names}, path
    t0 = time.time()
    g, centres, w = mixture(d, dev)
    x_d = draw(N, centres, w, g)
    x_q = draw(Q, centres, w, g)
    gt = knn_torch(x_q, x_d, 100)
    log(f"[bench] data + ground truth: {time.time() - t0:.1f}s")
    cent = kmeans_train(torch.as_tensor(x_d.cpu().numpy(), device=dev), B, niter=20)
    cent = torch.as_tensor(cent.cpu().numpy(), device=dev)
    assign = kmeans_assign(x_d, cent)
    log(f"[bench] kmeans: {time.time() - t0:.1f}s")
    # training set: a 30 % sample of base points with their exact 10-NN (labels: partitions holding a kNN)
    n_tr = min(N, 300_000)
    tr_idx = torch.randperm(N, generator=torch.Generator().manual_seed(SEED))[:n_tr].to(dev)
    knn_tr = knn_torch(x_d[tr_idx], x_d, k + 1)[:, 1:]
    model, mean32, scale32 = train_probing_model(x_d, cent, assign, tr_idx, knn_tr, dev, log, t0)
    # learned redundancy, n_mul = 2 (LIRA_smallscale.py:77-97, 331-354): the redundancy_ratio fraction of
    # points with the largest predicted nprobe get a second partition chosen by the model
    top2 = torch.empty(N, 2, dtype=torch.int64, device=dev)
    neff = torch.empty(N, dtype=torch.int64, device=dev)
    with torch.no_grad():
        for a in range(0, N, 65536):
            xb = x_d[a:a + 65536]
            s = model((torch.cdist(xb, cent) - mean32) / scale32, xb)
            neff[a:a + 65536] = (s > 0.5).sum(1)
            top2[a:a + 65536] = s.topk(2).indices
    order = torch.argsort(neff, descending=True, stable=True)[:int(N * redundancy_ratio)]
    d2b = torch.full((N, 2), -1, dtype=torch.int64, device=dev)
    d2b[:, 0] = assign
    cur = assign[order]
    t1, t2 = top2[order, 0], top2[order, 1]
    second = torch.where(cur != t1, t1, torch.where(neff[order] > 1, t2, torch.full_like(t1, -1)))
    second = torch.where(neff[order] > 0, second, torch.full_like(second, -1))
    d2b[order, 1] = second
    log(f"[bench] redundancy: {int((second >= 0).sum())} second copies: {time.time() - t0:.1f}s")
    out = {"x_d": x_d.cpu().numpy().astype(np.float32), "x_q": x_q.cpu().numpy().astype(np.float32),
           "gt": gt.cpu().numpy().astype(np.int32), "centroids": cent.cpu().numpy().astype(np.float32),
           "scaler_mean": mean32.cpu().numpy(), "scaler_scale": scale32.cpu().numpy(),
           "data_2_bkt": d2b.cpu().numpy().astype(np.int32)}
    for i, wv in enumerate(weights_of(model)):
        out[f"mlp_{i}"] = wv
    os.makedirs(path, exist_ok=True)
    for n, v in out.items():
        np.save(os.path.join(path, n + ".npy"), v)
    del x_d, model
    torch.cuda.empty_cache()
    return out, path


# ---- sharded workload: chunk-seeded generator, model from rank 0's first two chunks ---------------------
def shard_chunk(c, d, centres, w, dev):
    import torch
    g = torch.Generator(device=dev).manual_seed(SEED * 1_000_003 + 1 + c)
    return draw(CHUNK, centres, w, g)


def shard_queries(Q, d, centres, w, dev):
    import torch
    g = torch.Generator(device=dev).manual_seed(SEED * 1_000_003 + 7_000_000)
    return draw(Q, centres, w, g)


def make_shard_model(d, B, k, dev, log):
    """Centroids, scaler and probing model of the sharded workload, from rank 0's first two chunks (1 M vectors), cached.
    Identical for every N (so the partitions do not move when ranks are added)."""
    import torch
    path = os.path.join(CACHE, f"bigann_model_d{d}_B{B}_k{k}_s{SEED}_v1")
    names = ["centroids", "scaler_mean", "scaler_scale"] + [f"mlp_{i}" for i in range(12)]
    if all(os.path.exists(os.path.join(path, n + ".npy")) for n in names):
        return {n: np.load(os.path.join(path, n + ".npy")) for n in names}, path
    t0 = time.time()
    _, centres, w = mixture(d, dev)
    x = torch.cat([shard_chunk(0, d, centres, w, dev), shard_chunk(1, d, centres, w, dev)])
    cent = kmeans_train(x, B, niter=20)
    assign = kmeans_assign(x, cent)
    n_tr = 300_000
    tr_idx = torch.randperm(x.shape[0], generator=torch.Generator().manual_seed(SEED))[:n_tr].to(dev)
    knn_tr = knn_torch(x[tr_idx], x, k + 1)[:, 1:]
    model, mean32, scale32 = train_probing_model(x, cent, assign, tr_idx, knn_tr, dev, log, t0)
    out = {"centroids": cent.cpu().numpy().astype(np.float32), "scaler_mean": mean32.cpu().numpy(),
           "scaler_scale": scale32.cpu().numpy()}
    for i, wv in enumerate(weights_of(model)):
        out[f"mlp_{i}"] = wv
    os.makedirs(path, exist_ok=True)
    for n, v in out.items():
        np.save(os.path.join(path, n + ".npy"), v)
    del x, model
    torch.cuda.empty_cache()
    return out, path


def recall_at(ids, gt, k):
    hit = (gt[:, :k, None] == ids[:, None, :k]).any(-1)
    return float(hit.sum(1).mean() / k)


def source_hash():
    """Hash of the kernel sources: ncu-derived numbers under profiles/ are only quoted when they were taken on this code."""
    h = hashlib.sha256()
    cs = os.path.join(ROOT, "lira-ann-search_b200", "csrc")
    for f in sorted(os.listdir(cs)):
        if f.endswith(".cuh"):   # (the kernels live in the .cuh files; lira_b200.cu is host code)
            h.update(open(os.path.join(cs, f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(kernel_key):
    """dram bytes per launch of the dominant kernel from the newest ncu --set full capture whose summary
    (profiles/r*_scan_traffic.json, written by tools/make_profiles.py) carries the hash of the CURRENT kernel sources."""
    best = None
    pd = os.path.join(ROOT, "profiles")
    for f in sorted(os.listdir(pd)) if os.path.isdir(pd) else []:
        if re.match(r"r\d+_scan_traffic.*\.json$", f):
            try:
                j = json.load(open(os.path.join(pd, f)))
            except Exception:
                continue
            if j.get("source_hash") == source_hash() and j.get("workload") == kernel_key:
                best = float(j["dram_bytes_per_launch"])
    return best


# ---------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------------------------
class Clocks:
    """SM clock and throttle reasons sampled through NVML (the library behind nvidia-smi's clocks.sm /
    clocks_event_reasons.* fields) from a polling thread: the timed region of this bench is a few milliseconds,
    shorter than one `nvidia-smi -lms` period, so the samples are taken in-process every ~0.5 ms. mark() brackets
    the timed region; the summary is over the samples inside it."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu=0):
        self.gpu, self.rows, self.stop_flag, self.thread, self.h = gpu, [], False, None, None
        self.t0 = self.t1 = None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception as e:  # no NVML: the line then says so
            print(f"[bench] clocks: NVML unavailable ({e})", file=sys.stderr)
            self.h = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.perf_counter(), float(mhz), int(rs)))
            except Exception:
                pass
            time.sleep(0.0005)

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1.0)
        inside = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= self.t1]
        scope = "timed region"
        if not inside:  # (should not happen: the poll period is far below one step)
            inside, scope = self.rows, "warm-up + timed region"
        sm = [r[1] for r in inside]
        reasons = sorted({name for r in inside for name, bit in self.REASONS if r[2] & bit})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(sm), "scope": scope, "source": "NVML, polled in-process every ~0.5 ms"}


# ---------------------------------------------------------------------------------------------
# config dicts: built from the workload + cached operating point only, so both arms print the same one
# ---------------------------------------------------------------------------------------------
def config_sift1m(wl, op, k, world, shard):
    N, d = wl["x_d"].shape
    par = "single GPU" if world == 1 else (
        f"lists striped over {world} ranks + NCCL all-gather + merge" if shard == "lists" else
        f"{world} index replicas, one {len(wl['x_q'])}-query batch per rank and step, no data-path collective")
    return {"workload": "sift1m-shape", "N": int(N), "d": int(d), "Q": int(len(wl["x_q"])), "B": int(wl["centroids"].shape[0]),
            "k": k, "n_mul": 2, "redundancy_ratio": 0.03, "generator": GEN,
            "select": "b200: score > threshold, ids de-duplicated before top-k (LIRA_smallscale.py:206-214); reference search.cpp: "
                      "score >= threshold + argmax fallback, duplicates removed after top-k (search.cpp:448-513); the threshold is "
                      "the largest one of the sweep where BOTH reach the recall target",
            "threshold": op["threshold"], "recall_at_10": op["recall_b200"], "recall_at_10_reference_semantics": op["recall_ref"],
            "avg_nprobe": op["avg_nprobe"], "avg_cmp": op["avg_cmp"],
            "reference_sample": f"first {REF_SAMPLE} of {len(wl['x_q'])} queries per step",
            "l2_between_steps": "flushed (256 MiB write); probed lists per step also exceed the 126 MB L2",
            "parallelism": par}


def config_shard(meta, op, world, share=None):
    SHARE = share if share else (TOTAL // world if world > 1 else 12_500_000)   # (shadows the module constant: vectors per rank of THIS run)
    return {"workload": "bigann-shape-sharded", "N": int(world * SHARE), "vectors_per_rank": SHARE, "d": meta["d"], "Q": meta["Q"],
            "B": meta["B"], "k": meta["k"], "n_mul": 2, "redundancy": "full 2x: every vector is stored in its two nearest partitions",
            "list_entries": int(2 * world * SHARE), "generator": dict(GEN, chunk_rows=CHUNK, chunk_seed="43 * 1000003 + 1 + chunk"),
            "model": "centroids + probing model from rank 0's first 1 M vectors, the same for every N",
            "select": "score > threshold, ids de-duplicated before top-k", "threshold": op["threshold"],
            "recall_at_10": op["recall_b200"], "avg_nprobe": op["avg_nprobe"], "avg_cmp": op["avg_cmp"],
            "reference_sample": f"search.cpp on rank 0's first {REF_SHARD_ROWS} vectors (1/{world * SHARE // REF_SHARD_ROWS} of the dataset), "
                                f"first {REF_SHARD_QUERIES} queries per step",
            "l2_between_steps": "flushed (256 MiB write); the probed lists (GBs per rank) exceed the 126 MB L2",
            "parallelism": f"every list striped over {world} ranks by vector id (each rank generates and holds only its {SHARE} vectors) "
                           f"+ NCCL all-gather of packed top-k keys + merge kernel, inside the timed region"}


def load_op(path):
    try:
        return json.load(open(os.path.join(path, "op.json")))
    except Exception:
        return None


def save_op(path, op):
    os.makedirs(path, exist_ok=True)
    tmp = os.path.join(path, f"op.json.{os.getpid()}")
    json.dump(op, open(tmp, "w"))
    os.replace(tmp, os.path.join(path, "op.json"))


def shard_share(args, world):
    """vectors per rank: --share given = weak scaling (the dataset grows with N); default = the 100 M-vector dataset split over the ranks"""
    if args.share:
        return args.share
    return TOTAL // world if world > 1 else SHARE   # (one GPU cannot hold the fp32 rows of all 200 M entries: one eighth of them)


def shard_op_path(world, args):
    return os.path.join(CACHE, f"bigann_op_N{world}_S{shard_share(args, world)}_Q{args.Q}_B{args.B}_k{args.k}_r{args.recall}_v2")


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU search.cpp (oracle/_ref/search_ref), else the oracle port.
# Nothing here imports lira_ann_search_b200 or loads liblira_b200.so.
# ---------------------------------------------------------------------------------------------
def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def prepare_in_subprocess(args, world, log):
    """The operating point needs the GPU path (recall of the actual search): a separate process builds and caches it."""
    cmd = [sys.executable]
    if world > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                "--master-port", str(free_port())]
    cmd += [os.path.join(ROOT, "bench.py"), "--prepare", "--gpus", str(world), "--k", str(args.k), "--N", str(args.N),
            "--Q", str(args.Q), "--B", str(args.B), "--recall", str(args.recall), "--share", str(args.share)]
    if args.workload:
        cmd += ["--workload", args.workload]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_PORT", "MASTER_ADDR",
                                                             "GROUP_RANK", "ROLE_RANK", "LOCAL_WORLD_SIZE", "ROLE_WORLD_SIZE",
                                                             "TORCHELASTIC_RUN_ID", "GROUP_WORLD_SIZE", "ROLE_NAME")}
    t0 = time.time()
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    log(f"[bench] --prepare subprocess rc={r.returncode} in {time.time() - t0:.1f}s")
    if r.returncode != 0:
        log(r.stderr[-2000:])
    return r.returncode == 0


def write_torchscript(wl, B, d, path):
    import torch
    model = make_mlp(B, d)
    sd = {}
    for i, kk in enumerate(MLP_KEYS):
        sd[kk + ".weight"] = torch.as_tensor(wl[f"mlp_{2 * i}"])
        sd[kk + ".bias"] = torch.as_tensor(wl[f"mlp_{2 * i + 1}"])
    model.load_state_dict(sd)
    torch.jit.save(torch.jit.script(model.eval()), path)


def write_xvecs(path, arr):
    arr = np.ascontiguousarray(arr)
    rec = np.empty((arr.shape[0], arr.shape[1] + 1), np.int32)
    rec[:, 0] = arr.shape[1]
    rec[:, 1:] = arr.view(np.int32)
    rec.tofile(path)


def time_search_cpp(args, art, x_q, gt, thr, log):
    """art: dict with centroids, data_2_bkt, x_d, scaler_*, mlp_*. Returns (qps, recall, kind, cores_used)."""
    k = args.k
    exe = os.path.join(ROOT, "oracle", "_ref", "search_ref")
    cores = os.cpu_count() or 1
    passes = args.warmup + args.steps
    if os.path.exists(exe):
        with tempfile.TemporaryDirectory() as td:
            pfx = os.path.join(td, "bench")
            for n in ("centroids", "data_2_bkt", "x_d", "scaler_mean", "scaler_scale"):
                np.save(f"{pfx}_{n}.npy", art[n])
            B, d = art["centroids"].shape
            write_torchscript(art, B, d, pfx + "_mlp_2_input.pt")
            ds = os.path.join(td, "data", "bench")
            os.makedirs(ds)
            write_xvecs(os.path.join(ds, "bench_query.fvecs"), x_q.astype(np.float32))
            write_xvecs(os.path.join(ds, "bench_groundtruth.ivecs"), gt.astype(np.int32))
            # one process, `passes` thresholds 1e-7 apart: the artifacts load once, every pass is one step
            step = 1e-7
            cmd = [exe, "--dataset", "bench", "--data_path", os.path.join(td, "data"), "--artifacts_dir", td,
                   "--prefix", "bench", "--k", str(k), "--metric", "L2", "--num_threads", str(cores),
                   "--t_min", repr(thr), "--t_max", repr(thr + step * (passes - 1) + step / 4), "--t_step", repr(step)]
            t0 = time.time()
            txt = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
            log(f"[bench] reference search.cpp ran in {time.time() - t0:.1f}s")
        qps = [float(x) for x in re.findall(r"QPS\s*:\s*([-+0-9.eE]+)", txt)][args.warmup:]
        rec = [float(x) for x in re.findall(r"avg_recall\s*:\s*([-+0-9.eE]+)", txt)][args.warmup:]
        # search.cpp has no OpenMP pragma: one thread per query by construction, whatever --num_threads says
        return float(np.mean(qps)), float(np.mean(rec)), "reference", 1
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    off, ids, vecs = O.build_lists_from_data_2_bkt(art["x_d"], art["data_2_bkt"], art["centroids"].shape[0])
    w = [art[f"mlp_{i}"] for i in range(12)]
    times = []
    for _ in range(passes):
        t0 = time.perf_counter()
        f = O.features_cpp(x_q, art["centroids"], art["scaler_mean"], art["scaler_scale"])
        _, probs, _ = O.mlp_forward(f, x_q, w)
        poff, pids = O.select(probs.astype(np.float32), O.SELECT_GE_ARGMAX, thr)
        out_ids, _, _ = O.search(off, ids, vecs, x_q, poff, pids, k, O.L2, O.F32, 0)
        times.append(time.perf_counter() - t0)
    return len(x_q) / float(np.mean(times[args.warmup:])), recall_at(out_ids, gt, k), "port", O.num_threads()


def run_reference(args, world, log):
    import torch
    dev = "cuda:0"
    cores = os.cpu_count() or 1
    sharded = world > 1 and args.workload == "bigann"
    if not sharded:
        wl, wl_path = make_workload(args.N, 128, args.Q, args.B, args.k, dev=dev, log=log)
        op = load_op(wl_path)
        if op is None and prepare_in_subprocess(args, 1, log):
            op = load_op(wl_path)
        if op is None:
            op = {"threshold": 0.02, "recall_b200": None, "recall_ref": None, "avg_nprobe": None, "avg_cmp": None,
                  "note": "operating point not available (prepare failed): first grid threshold"}
        n = min(REF_SAMPLE, len(wl["x_q"]))
        qps, rec, kind, used = time_search_cpp(args, wl, wl["x_q"][:n], wl["gt"][:n], op["threshold"], log)
        config = config_sift1m(wl, op, args.k, world, args.shard)
        sample = f"first {n} of {len(wl['x_q'])} queries per step, threshold {op['threshold']:g}, recall@{args.k} {rec:.4f}"
        ms = 1e3 * n / qps
    else:
        d, B, k, Q = 128, args.B, args.k, args.Q
        mdl, _ = make_shard_model(d, B, k, dev, log)
        op = load_op(shard_op_path(world, args))
        if op is None and prepare_in_subprocess(args, world, log):
            op = load_op(shard_op_path(world, args))
        if op is None:
            op = {"threshold": 0.02, "recall_b200": None, "avg_nprobe": None, "avg_cmp": None,
                  "note": "operating point not available (prepare failed): default threshold"}
        _, centres, w = mixture(d, dev)
        x = torch.cat([shard_chunk(c, d, centres, w, dev) for c in range(REF_SHARD_ROWS // CHUNK)])
        x_q = shard_queries(Q, d, centres, w, dev)[:REF_SHARD_QUERIES]
        cent = torch.as_tensor(mdl["centroids"], device=dev)
        b2 = torch.cat([torch.cdist(x[a:a + 131072], cent).topk(2, dim=1, largest=False).indices for a in range(0, x.shape[0], 131072)])
        gt = knn_torch(x_q, x, 100)
        art = dict(mdl)
        art["x_d"] = x.cpu().numpy().astype(np.float32)
        art["data_2_bkt"] = b2.cpu().numpy().astype(np.int32)
        qps, rec, kind, used = time_search_cpp(args, art, x_q.cpu().numpy(), gt.cpu().numpy(), op["threshold"], log)
        config = config_shard({"d": d, "Q": Q, "B": B, "k": k}, op, world, shard_share(args, world))
        sample = (f"rank 0's first {REF_SHARD_ROWS} vectors only (1/{world * shard_share(args, world) // REF_SHARD_ROWS} of the dataset; search.cpp's time per "
                  f"query grows linearly with the probed entries), first {REF_SHARD_QUERIES} queries per step, threshold "
                  f"{op['threshold']:g}, recall@{k} on that sub-sample {rec:.4f}")
        ms = 1e3 * REF_SHARD_QUERIES / qps
    line = {"metric": "qps_at_recall10_ge_0.95", "value": qps, "unit": "queries/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong" if (sharded and not args.share) else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": used, "kind": kind, "sample": sample, "host_cores": cores},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------
def timed_steps(step, args, stream, flush, index, dist, clocks, rank, sync_step):
    """W untimed warm-up steps, then exactly K steps: per-step CUDA events on `stream`, L2 flushed between steps (outside the
    events), barrier + synchronize on both sides, max over ranks. `step` only ENQUEUES a batch (lira_probe_search_enqueue_dev
    [+ the NCCL all-gather and the merge]): consecutive steps queue up behind each other on the stream like the batches of a
    serving loop, and lira_index_finish checks every batch's status words after the loop. Kernel-level timings come from
    three extra, individually synchronised steps (`sync_step`) after the timed region."""
    import torch
    import lira_ann_search_b200 as L
    # the library's per-kernel events sit BETWEEN the launches of a batch (and cut the programmatic dependent launch chain
    # there): they are recorded in the three extra steps below only, the timed region runs the launch sequence a caller gets
    index.set_timing(False)
    for _ in range(args.warmup):
        flush.zero_()
        step()
    index.finish()
    torch.cuda.synchronize()
    launches0 = L.launch_count()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    clocks.mark_begin()
    t_wall = time.perf_counter()
    res = None
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record(stream)
        res = step()
        ev[i][1].record(stream)
    index.finish()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    wall = time.perf_counter() - t_wall
    clocks.mark_end()
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    launches = L.launch_count() - launches0
    if dist is not None:
        t = torch.tensor([total_ms], device=flush.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    tms = []
    index.set_timing(True)
    for _ in range(3):
        flush.zero_()
        sync_step()
        torch.cuda.synchronize()
        tms.append(index.last_timing())
    return res, total_ms, wall, launches, tms


def find_op_sift1m(index, model, wl, args, dev, gather_merge, log):
    """Largest threshold where this path (score > thr, dedup before top-k) AND the reference's semantics (score >= thr +
    argmax, duplicates removed after top-k: search.cpp:448-513) reach the recall target -- on all queries and, for the
    reference semantics, also on the first REF_SAMPLE queries the reference arm times. Coarse 0.02 grid
    (LIRA_smallscale.py:199), then steps of 0.002 between the last passing and the first failing grid point."""
    import torch
    import lira_ann_search_b200 as L
    Q, B, k = len(wl["x_q"]), wl["centroids"].shape[0], args.k
    d_q = torch.as_tensor(wl["x_q"], device=dev)
    scores = torch.zeros((Q, (B + 3) // 4 * 4), dtype=torch.float32, device=dev)
    scores[:, :B] = torch.as_tensor(model.scores(wl["x_q"]), device=dev)
    gt = wl["gt"]
    sweep = []

    def probe(thr):
        D, I, npb, cmp_ = index.select_search_dev(scores, d_q, L.SELECT_GT, thr, k, True)
        D, I = gather_merge(D, I)
        Dr, Ir, _, _ = index.select_search_dev(scores, d_q, L.SELECT_GE_ARGMAX, thr, k, False)
        Dr, Ir = gather_merge(Dr, Ir)
        torch.cuda.synchronize()
        Ih, Irh = I.cpu().numpy(), Ir.cpu().numpy()
        r = {"threshold": thr, "recall_b200": recall_at(Ih, gt, k), "recall_ref": recall_at(Irh, gt, k),
             "recall_ref_sample": recall_at(Irh[:REF_SAMPLE], gt[:REF_SAMPLE], k),
             "avg_nprobe": float(npb.float().mean()), "avg_cmp": float(cmp_.float().mean())}
        r["ok"] = min(r["recall_b200"], r["recall_ref"], r["recall_ref_sample"]) >= args.recall
        sweep.append(r)
        return r

    best, fail = None, None
    for i in range(1, 41):
        r = probe(round(0.02 * i, 2))
        if not r["ok"]:
            fail = r
            break
        best = r
    lo = best["threshold"] if best else 0.0
    hi = fail["threshold"] if fail else None
    if hi is not None:
        t = lo + 0.002
        while t < hi - 1e-9:
            r = probe(round(t, 3))
            if not r["ok"]:
                break
            best = r
            t += 0.002
    if best is None:   # even the first grid point fails: go below it
        for t in (0.015, 0.01, 0.005, 0.002, 0.001):
            r = probe(t)
            if r["ok"]:
                best = r
                break
    if best is None:
        best = sweep[0]
        log(f"[bench] WARNING: recall target {args.recall} not reached (recall {best['recall_b200']:.4f} at {best['threshold']})")
    op = dict(best)
    op["sweep"] = [(s["threshold"], s["recall_b200"], s["recall_ref"], s["avg_nprobe"]) for s in sweep]
    log(f"[bench] operating point: threshold {op['threshold']} recall@{k} {op['recall_b200']:.4f} (reference semantics "
        f"{op['recall_ref']:.4f}) nprobe {op['avg_nprobe']:.2f}")
    return op


def run_sift1m(args, rank, world, local, dist, log):
    import torch
    import lira_ann_search_b200 as L
    dev = f"cuda:{local}"
    if rank == 0:
        wl, wl_path = make_workload(args.N, 128, args.Q, args.B, args.k, dev=dev, log=log)
    if dist is not None:
        dist.barrier()
    if rank != 0:
        wl, wl_path = make_workload(args.N, 128, args.Q, args.B, args.k, dev=dev, log=log)
    N, d = wl["x_d"].shape
    Q, B, k = len(wl["x_q"]), wl["centroids"].shape[0], args.k
    weights = [wl[f"mlp_{i}"] for i in range(12)]
    d2b = wl["data_2_bkt"]
    shard_lists = world > 1 and args.shard == "lists"
    if shard_lists:
        from lira_ann_search_b200.parallel import stripe_assignment
        d2b = stripe_assignment(d2b, B, rank, world)
    index = L.LiraIndex.from_data_2_bkt(wl["x_d"], d2b, B, "L2", device=local)
    model = L.LiraModel.from_arrays(wl["centroids"], wl["scaler_mean"], wl["scaler_scale"], weights, device=local)
    d_q = torch.as_tensor(wl["x_q"], device=dev)
    gt = wl["gt"]

    def gather_merge(D, I):
        if not shard_lists:
            return D, I
        from lira_ann_search_b200.parallel import allgather_merge
        return allgather_merge(D, I, k, "L2", dedup=True, device=local)

    op = load_op(wl_path)
    if op is None or abs(op.get("target", -1) - args.recall) > 1e-12:
        op = find_op_sift1m(index, model, wl, args, dev, gather_merge, log)
        op["target"] = args.recall
        if rank == 0:
            save_op(wl_path, op)
    else:
        log(f"[bench] operating point (cached): threshold {op['threshold']} recall@{k} {op['recall_b200']:.4f} nprobe {op['avg_nprobe']:.2f}")
    thr = op["threshold"]
    if args.prepare:
        return

    out = None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    index.set_timing(True)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(stream)

    def step():
        nonlocal out
        out = index.probe_search_enqueue_dev(model, d_q, L.SELECT_GT, thr, k, True, out=out)
        return gather_merge(out[0], out[1])

    def sync_step():
        nonlocal out
        out = index.probe_search_dev(model, d_q, L.SELECT_GT, thr, k, True, out=out)

    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    (D, I), total_ms, wall, launches, tms = timed_steps(step, args, stream, flush, index, dist, clocks, rank, sync_step)
    clk = clocks.stop() if rank == 0 else None
    rec = recall_at(I.cpu().numpy(), gt, k)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    s_ms = float(np.mean([t["scan_ms"] for t in tms]))
    trio_ms = float(np.mean([t.get("scan_total_ms", t["scan_ms"]) for t in tms]))
    s_bytes = float(np.mean([t["scan_bytes"] for t in tms]))
    achieved = s_bytes / (s_ms * 1e-3) / 1e9
    flops = 2.0 * d * float(np.mean([t["scan_pairs"] for t in tms]))

    # ---- the same scan in its HBM-bound regime: 1024-query batches (~10 queries per probed list, where the
    # >= 70 % HBM target is meaningful; at 10 000 queries per batch the scan is bound by the operand stream out of L2) ----
    small = None
    if Q >= 2048:
        qs = 1024
        out_s = None
        s_ms_l, s_by_l = [], []
        for i in range(3 + 5):
            flush.zero_()
            out_s = index.probe_search_dev(model, d_q[:qs], L.SELECT_GT, thr, k, True, out=out_s)
            torch.cuda.synchronize()
            tm = index.last_timing()
            if i >= 3:
                s_ms_l.append(tm["scan_ms"]); s_by_l.append(tm["scan_bytes"])
        ach = float(np.mean(s_by_l)) / (float(np.mean(s_ms_l)) * 1e-3) / 1e9
        small = {"Q": qs, "kernel_ms": float(np.mean(s_ms_l)), "algorithmic_bytes": float(np.mean(s_by_l)), "achieved": ach,
                 "frac": ach / hbm_peak, "unit": "GB/s",
                 "note": "first 1024 queries of the batch as one batch; the kernel streams the fp16 copy of the rows, so the PHYSICAL "
                         "fraction of the HBM peak is about half of this algorithmic (fp32-byte) one"}

    e2e = measure_e2e(args, index, model, wl["x_q"], thr, k, gather_merge if shard_lists else None, flush, dist, dev)
    if rank != 0:
        return

    # ---- CPU baseline (the oracle port, bounded sample, same probe sets): at N = 1 only ---------------
    cpu_baseline = None
    if world == 1:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O
        ns = min(args.cpu_sample, Q)
        off, ids, vecs = O.build_lists_from_data_2_bkt(wl["x_d"], wl["data_2_bkt"], B)
        t0 = time.perf_counter()
        f = O.features_cpp(wl["x_q"][:ns], wl["centroids"], wl["scaler_mean"], wl["scaler_scale"])
        _, probs, _ = O.mlp_forward(f, wl["x_q"][:ns], weights)
        poff, pids = O.select(probs.astype(np.float32), O.SELECT_GT, thr)
        cids, _, _ = O.search(off, ids, vecs, wl["x_q"][:ns], poff, pids, k, O.L2, O.F32, 1)
        cpu_s = time.perf_counter() - t0
        same = float(np.mean((e2e["ids"][:ns] == cids).all(1)))
        cpu_baseline = {"value": ns / cpu_s, "unit": "queries/s", "cores": O.num_threads(), "kind": "port",
                        "sample": f"first {ns} of {Q} queries, same threshold; ids identical to GPU on {same} of rows"}
    jobs = 1 if (world == 1 or shard_lists) else world   # query batches answered per step by the whole job
    tensor_path = index.last_path == "tensor-core"
    line = {
        "metric": "qps_at_recall10_ge_0.95", "value": jobs * Q * args.steps / (total_ms * 1e-3), "unit": "queries/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if shard_lists else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_sift1m(wl, op, k, world, args.shard), "recall_at_10": rec,
        "e2e": {"value": jobs * Q / e2e["pinned_s"], "unit": "queries/s", "h2d_bytes_per_step": int(Q * d * 4),
                "d2h_bytes_per_step": int(Q * k * 12 + Q * 12), "pageable_value": jobs * Q / e2e["pageable_s"],
                "steps_timed": e2e["pinned_s_steps"], "median_step_value": jobs * Q / e2e["pinned_s_median"],
                "api": e2e["api"]},
        "gpu_launches": int(launches), "clocks": clk,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": measured_traffic("sift1m-shape"),
                     "kernel": "tc_scan_kernel<false, false> (tcgen05 list scan, fp16 operands)" if tensor_path else "scan_lists_kernel",
                     "kernel_ms": s_ms, "scan_total_ms": trio_ms, "frac_scan_total": s_bytes / (trio_ms * 1e-3) / 1e9 / hbm_peak,
                     "scan_total_is": "seed + filter + refine (every kernel of the scan as SURVEY.md 8d defines it)",
                     "hbm_bound_point": small, "algorithmic_bytes": s_bytes,
                     "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                     "tensor_companion": {"flops": flops, "achieved_tflops": flops / (s_ms * 1e-3) / 1e12,
                                          "peak_tflops": float(peaks.get("bf16_tflops", 1590.0)),
                                          "note": "2*d*pairs (algorithmic) against the measured cuBLAS bf16 burst figure; the kernel's operands are fp16"}},
        "wall_s_timed_region": wall,
    }
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    print(json.dumps(line), flush=True)


def measure_e2e(args, index, model, x_q, thr, k, gather_merge, flush, dist, dev):
    """The same step through the host-buffer C ABI call (H2D of the queries, D2H of the results inside the timed region):
    from pinned memory (the contract's definition) and from an ordinary pageable numpy array (what a reference caller has).
    Consecutive batches are submitted through the asynchronous submit / wait pair when the library has it, so the copies of
    batch i+1 overlap the kernels of batch i; the time per step is the steady-state period of that pipeline."""
    import torch
    import lira_ann_search_b200 as L
    Q, d = x_q.shape
    index.set_timing(False)   # (a caller does not ask for the per-kernel events: they sit between the launches of a batch)
    pin_q = torch.empty((Q, d), dtype=torch.float32).pin_memory()
    pin_q.copy_(torch.as_tensor(x_q))
    res = {}
    # (a step is ~0.4 ms of host wall clock: enough repetitions that one scheduler hiccup on the host does not move the mean)
    # warm-up: 3 steps, or 10 for the pipelined API (its first batches allocate the staging slots and time the direct upload of the
    # pinned array to decide whether to stage it: one-off host-side costs of a handle, milliseconds against a 0.4 ms step)
    n_warm = 3 if gather_merge is not None else 10
    n_rep = n_warm + (max(6, min(args.steps, 20)) if gather_merge is not None else max(60, min(10 * args.steps, 200)))
    pipelined = hasattr(index, "probe_search_submit") and gather_merge is None
    res["api"] = "lira_probe_search_submit / lira_probe_search_wait, three batches in flight" if pipelined else "lira_probe_search"
    for name, q_host in (("pinned_s", pin_q.numpy()), ("pageable_s", np.array(x_q, copy=True))):
        outs = [None, None]
        ts = []
        ids = None
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        if pipelined:
            # steady state with two batches ahead: submit(i + 2) before wait(i); the period between completed batches is timed
            outs = [None] * 4
            index.probe_search_submit(model, q_host, L.SELECT_GT, thr, k, True, slot=0)
            index.probe_search_submit(model, q_host, L.SELECT_GT, thr, k, True, slot=1)
            for i in range(n_rep):
                t0 = time.perf_counter()
                index.probe_search_submit(model, q_host, L.SELECT_GT, thr, k, True, slot=(i + 2) & 3)
                outs[i & 3] = index.probe_search_wait(slot=i & 3, out=outs[i & 3])
                ts.append(time.perf_counter() - t0)
            for i in (n_rep, n_rep + 1):
                outs[i & 3] = index.probe_search_wait(slot=i & 3, out=outs[i & 3])
            ids = outs[0][1]
        elif gather_merge is not None:
            # sharded: pinned / pageable host queries -> device, the batch, the NCCL all-gather + merge, merged results -> host
            d_q = torch.empty((Q, d), dtype=torch.float32, device=dev)
            q_t = torch.as_tensor(q_host)
            D_h = torch.empty((Q, k), dtype=torch.float32).pin_memory()
            I_h = torch.empty((Q, k), dtype=torch.int64).pin_memory()
            dev_out = None
            for i in range(n_rep):
                flush.zero_()
                torch.cuda.synchronize()
                if dist is not None:
                    dist.barrier()
                t0 = time.perf_counter()
                d_q.copy_(q_t, non_blocking=True)
                dev_out = index.probe_search_enqueue_dev(model, d_q, L.SELECT_GT, thr, k, True, out=dev_out)
                Dg, Ig = gather_merge(dev_out[0], dev_out[1])
                D_h.copy_(Dg, non_blocking=True)
                I_h.copy_(Ig, non_blocking=True)
                index.finish()
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            ids = I_h.numpy().copy()
            res["api"] = "host queries -> lira_probe_search_enqueue_dev -> NCCL all-gather + lira_merge_ranks_dev -> host results, one batch at a time"
        else:
            host_out = None
            for i in range(n_rep):
                flush.zero_()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                host_out = index.probe_search(model, q_host, L.SELECT_GT, thr, k, True, out=host_out)
                ts.append(time.perf_counter() - t0)
            ids = host_out[1]
        s = float(np.mean(ts[n_warm:]))
        if dist is not None:
            t = torch.tensor([s], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s = float(t.item())
        res[name] = s
        res[name + "_median"] = float(np.median(ts[n_warm:]))
        res[name + "_steps"] = len(ts) - n_warm
        res["ids"] = ids
    return res


def run_shard(args, rank, world, local, dist, log):
    """N > 1 default: the BigANN-100M-shape dataset (100 M vectors, 200 M list entries) split over the ranks, every list striped
    by vector id: strong scaling, queries/s grows with N. --share S: S vectors per rank instead (weak scaling)."""
    import torch
    import lira_ann_search_b200 as L
    from lira_ann_search_b200.parallel import allgather_merge
    dev = f"cuda:{local}"
    d, B, k, Q = 128, args.B, args.k, args.Q
    t0 = time.time()
    if rank == 0:
        mdl, _ = make_shard_model(d, B, k, dev, log)
    if dist is not None:
        dist.barrier()
    if rank != 0:
        mdl, _ = make_shard_model(d, B, k, dev, log)
    _, centres, w = mixture(d, dev)
    share = shard_share(args, world)
    n_chunks = share // CHUNK
    x = torch.empty((share, d), dtype=torch.float32, device=dev)
    for j in range(n_chunks):
        x[j * CHUNK:(j + 1) * CHUNK] = shard_chunk(rank * n_chunks + j, d, centres, w, dev)
    x_q = shard_queries(Q, d, centres, w, dev)
    base_id = rank * share
    # exact ground truth of the whole dataset: per-rank exact kNN (library, base adopted on the device) + the cross-rank merge
    kn = L.KnnIndex(x, "L2")
    Dg, Ig = kn.search_dev(x_q, k)
    Ig = torch.where(Ig >= 0, Ig + base_id, Ig)
    if dist is not None:
        Dg, Ig = allgather_merge(Dg, Ig, k, "L2", dedup=True, device=local)
    torch.cuda.synchronize()
    gt = Ig.cpu().numpy()
    kn.close()
    del kn
    log(f"[bench] rank 0: {share} vectors + exact ground truth over {world * share}: {time.time() - t0:.1f}s")
    # the two nearest partitions of every vector: exact 2-NN against the centroid table (library kNN, device in / device out)
    cent = torch.as_tensor(mdl["centroids"], device=dev)
    kc = L.KnnIndex(cent, "L2")
    b2 = torch.empty((share, 2), dtype=torch.int64, device=dev)
    for a in range(0, share, 2_000_000):
        kc.search_dev(x[a:a + 2_000_000], 2, out=(torch.empty((min(2_000_000, share - a), 2), dtype=torch.float32, device=dev), b2[a:a + 2_000_000]))
    torch.cuda.synchronize()
    kc.close()
    del kc
    ids = (torch.arange(share, device=dev, dtype=torch.int64) + base_id).repeat_interleave(2)
    key = b2.reshape(-1) * (1 << 32) + ids
    order = torch.argsort(key)
    sizes = torch.bincount(b2.reshape(-1), minlength=B)
    off = torch.zeros(B + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(sizes, 0)
    ids32 = ids[order].to(torch.int32)
    local_rows = (ids[order] - base_id)
    del key, order, b2, ids
    E = int(off[-1])
    vecs = torch.empty((E, d), dtype=torch.float32, device=dev)
    for s in range(0, E, 4_000_000):
        vecs[s:s + 4_000_000] = x[local_rows[s:s + 4_000_000]]
    del x, local_rows
    torch.cuda.empty_cache()
    index = L.LiraIndex.from_device(vecs, ids32, off.cpu().numpy(), d, "L2", device=local)
    weights = [mdl[f"mlp_{i}"] for i in range(12)]
    model = L.LiraModel.from_arrays(mdl["centroids"], mdl["scaler_mean"], mdl["scaler_scale"], weights, device=local)
    log(f"[bench] rank 0: {E} list entries in {B} lists ({int(sizes.min())}..{int(sizes.max())} per list), mode {index.tensor_core_mode}: {time.time() - t0:.1f}s")

    def gather_merge(D, I):
        if dist is None:
            return D, I
        return allgather_merge(D, I, k, "L2", dedup=True, device=local)

    # ---- operating point: largest threshold with recall@10 >= target against the exact global ground truth ----
    x_q_h = x_q.cpu().numpy()
    scores = torch.zeros((Q, (B + 3) // 4 * 4), dtype=torch.float32, device=dev)
    scores[:, :B] = torch.as_tensor(model.scores(x_q_h), device=dev)
    op_path = shard_op_path(world, args)

    def probe(thr):
        D, I, npb, cmp_ = index.select_search_dev(scores, x_q, L.SELECT_GT, thr, k, True)
        D, I = gather_merge(D, I)
        torch.cuda.synchronize()
        c = cmp_.double().mean().reshape(1)
        if dist is not None:
            dist.all_reduce(c)
        return {"threshold": thr, "recall_b200": recall_at(I.cpu().numpy(), gt, k), "avg_nprobe": float(npb.float().mean()),
                "avg_cmp": float(c.item())}

    op = load_op(op_path)
    if op is None or op.get("share") != share:
        sweep = []
        best, prev_fail = None, None
        for thr in (0.5, 0.3, 0.2, 0.1, 0.05, 0.02, 0.01, 0.005, 0.002, 0.001, 0.0005, 0.0002, 0.0001):
            r = probe(thr)
            sweep.append(r)
            if r["recall_b200"] >= args.recall:
                best = r
                break
            prev_fail = r
        if best is None:
            best = sweep[-1]
            log(f"[bench] WARNING: recall target not reached, recall {best['recall_b200']:.4f}")
        elif prev_fail is not None:   # bisect (geometrically) between the passing and the failing threshold
            lo, hi = best["threshold"], prev_fail["threshold"]
            for _ in range(5):
                mid = float(np.sqrt(lo * hi))
                r = probe(mid)
                sweep.append(r)
                if r["recall_b200"] >= args.recall:
                    best, lo = r, mid
                else:
                    hi = mid
        op = dict(best)
        op["share"] = share
        op["sweep"] = [(s["threshold"], s["recall_b200"], s["avg_nprobe"]) for s in sweep]
        if rank == 0:
            save_op(op_path, op)
    thr = op["threshold"]
    log(f"[bench] operating point: threshold {thr:g} recall@{k} {op['recall_b200']:.4f} nprobe {op['avg_nprobe']:.2f}: {time.time() - t0:.1f}s")
    if args.prepare:
        return

    out = None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    index.set_timing(True)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(stream)

    def local_step():
        nonlocal out
        out = index.probe_search_dev(model, x_q, L.SELECT_GT, thr, k, True, out=out)
        return out[0], out[1]

    merged = None

    def step():
        nonlocal out, merged
        out = index.probe_search_enqueue_dev(model, x_q, L.SELECT_GT, thr, k, True, out=out)
        if dist is not None:
            merged = allgather_merge(out[0], out[1], k, "L2", dedup=True, device=local, out=merged)
            return merged
        return out[0], out[1]

    # one rank's share alone (no collective): what a single GPU needs for its stripe
    for _ in range(3):
        local_step()
    alone = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        local_step()
        e1.record(stream)
        e1.synchronize()
        alone.append(e0.elapsed_time(e1))
    alone_ms = torch.tensor([float(np.mean(alone))], device=dev)
    if dist is not None:
        dist.all_reduce(alone_ms, op=dist.ReduceOp.MAX)
    alone_ms = float(alone_ms.item())

    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    (D, I), total_ms, wall, launches, tms = timed_steps(step, args, stream, flush, index, dist, clocks, rank, local_step)
    clk = clocks.stop() if rank == 0 else None
    Ih = I.cpu().numpy()
    rec = recall_at(Ih, gt, k)

    # ---- merged ids against a brute-force scan of the probed entries on a query sample (every rank scans its stripe with
    # torch, the candidates are gathered and merged on the host: distinct ids, ties by id) ----
    ns = min(args.check, Q)
    offh = off.cpu().numpy()
    sc_h = scores[:ns, :B].cpu().numpy()
    cand = []
    for qi in range(ns):
        lists = np.nonzero(sc_h[qi] > np.float32(thr))[0]
        if len(lists) == 0:
            cand.append((np.empty(0, np.float32), np.empty(0, np.int64)))
            continue
        rows = torch.cat([torch.arange(offh[b], offh[b + 1], device=dev) for b in lists])
        dist_ = ((vecs[rows] - x_q[qi]) ** 2).sum(1)
        gid = ids32[rows].long()
        kk = min(4 * k, dist_.numel())
        top = torch.topk(dist_, kk, largest=False)
        # everything at the kk-th distance or below (ties) so that the host-side order by (distance, id) is exact
        keep = dist_ <= top.values[-1]
        cand.append((dist_[keep].cpu().numpy(), gid[keep].cpu().numpy()))
    if dist is not None:
        allc = [None] * world
        dist.all_gather_object(allc, cand)
    else:
        allc = [cand]
    bad = 0
    if rank == 0:
        Dh = D.cpu().numpy()
        for qi in range(ns):
            dd = np.concatenate([allc[r][qi][0] for r in range(world)])
            gg = np.concatenate([allc[r][qi][1] for r in range(world)])
            o = np.lexsort((gg, dd))
            gs, dsrt = gg[o], dd[o]
            _, first = np.unique(gs, return_index=True)
            keep = np.sort(first)[:k]
            exp_i = np.full(k, -1, np.int64)
            exp_i[:len(keep)] = gs[keep]
            if not np.array_equal(exp_i, Ih[qi]) or not np.array_equal(dsrt[keep], Dh[qi][:len(keep)]):
                bad += 1

    e2e = measure_e2e(args, index, model, x_q_h, thr, k, gather_merge if dist is not None else None, flush, dist, dev)
    nccl_ranks = world if dist is not None else 1
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1356.7))
    s_ms = float(np.mean([t["scan_ms"] for t in tms]))
    trio_ms = float(np.mean([t.get("scan_total_ms", t["scan_ms"]) for t in tms]))
    s_bytes = float(np.mean([t["scan_bytes"] for t in tms]))
    flops = 2.0 * d * float(np.mean([t["scan_pairs"] for t in tms]))
    ach_tf = flops / (s_ms * 1e-3) / 1e12
    meta = {"d": d, "Q": Q, "B": B, "k": k}
    cfg = config_shard(meta, op, world, share)
    strong = not args.share and world > 1
    line = {
        "metric": "qps_at_recall10_ge_0.95", "value": Q * args.steps / (total_ms * 1e-3), "unit": "queries/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg, "recall_at_10": rec,
        "scaling_note": {"what_is_fixed": ("the dataset: 100 M vectors = 200 M list entries split over the ranks, and the query batch; ideal strong "
                                          "scaling multiplies queries/s by n_gpus. bench.py --gpus 1 runs the SIFT1M-shape workload "
                                          "(BASELINE.json configs[0]) instead, so compare N = 2, 4, 8 with each other: queries/s x n_gpus / 2") if strong else
                                         ("vectors_per_rank (the dataset grows n_gpus-fold, the query batch is fixed): ideal weak scaling keeps "
                                          "queries/s constant"),
                         "share_alone_ms": alone_ms, "ms_per_step": total_ms / args.steps,
                         "entries_scanned_per_s": op["avg_cmp"] * Q * args.steps / (total_ms * 1e-3),
                         "note": "share_alone_ms = one rank answering the batch on its own stripe, no collective (max over ranks); "
                                 "ms_per_step adds the NCCL all-gather and the merge kernel"},
        "check": {"queries": ns, "mismatches_vs_brute_force": bad, "how": "merged ids and distances against a torch brute-force scan of the "
                  "probed entries of every rank, merged on the host (distinct ids, ties by id)"},
        "comm": {"nranks": nccl_ranks, "collective_in_timed_region": "all_gather_into_tensor of Q*k packed 64-bit keys per rank (NCCL)",
                 "bytes_per_rank_and_step": int(Q * k * 8)},
        "e2e": {"value": Q / e2e["pinned_s"], "unit": "queries/s", "h2d_bytes_per_step": int(Q * d * 4),
                "d2h_bytes_per_step": int(Q * k * 12 + Q * 12), "pageable_value": Q / e2e["pageable_s"], "api": e2e["api"]},
        "gpu_launches": int(launches), "clocks": clk,
        "roofline": {"bound": "tensor", "achieved": ach_tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach_tf / tf_peak,
                     "traffic": measured_traffic("bigann-shape-sharded"),
                     "kernel": "tc_scan_kernel<false, false> (tcgen05 list scan, fp16 operands)" if index.last_path == "tensor-core" else "scan_lists_kernel",
                     "kernel_ms": s_ms, "scan_total_ms": trio_ms, "flops_algorithmic": flops,
                     "why_tensor": "about Q * nprobe / B queries share every probed list, far above the ~16 where the list stream binds "
                                   "(SURVEY.md 8d): the filter kernel is bound by the tensor pipe",
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback",
                     "hbm_companion": {"algorithmic_bytes": s_bytes, "achieved_gbs": s_bytes / (s_ms * 1e-3) / 1e9,
                                       "frac": s_bytes / (s_ms * 1e-3) / 1e9 / hbm_peak, "peak": hbm_peak}},
        "wall_s_timed_region": wall,
    }
    print(json.dumps(line), flush=True)



# ---------------------------------------------------------------------------------------------
# --workload knn: BASELINE.json configs[1] -- compute_knn's exact ground truth on SIFT1M shape
# ---------------------------------------------------------------------------------------------
def knn_data(N, Q, d, dev):
    import torch
    _, centres, w = mixture(d, dev)
    x = torch.cat([shard_chunk(c, d, centres, w, dev) for c in range((N + CHUNK - 1) // CHUNK)])[:N].contiguous()
    return x, shard_queries(Q, d, centres, w, dev)


def knn_cpu_baseline(x_h, q_h, k, n, log):
    """the oracle's exact kNN (plain C restatement of compute_knn.cpp:208-259, OpenMP over the queries) on the first n queries"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    O.knn(x_h[:20000], q_h[:4], k, O.L2, O.F32, 0)
    t0 = time.perf_counter()
    D, I = O.knn(x_h, q_h[:n], k, O.L2, O.F32, 0)
    dt = time.perf_counter() - t0
    return n / dt, cores, D, I


def run_knn_reference(args, log):
    import torch
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    d, k, N, Q = 128, 100, args.N, args.Q
    x, x_q = knn_data(N, Q, d, dev)
    n = 64
    qps, cores, _, _ = knn_cpu_baseline(x.cpu().numpy(), x_q.cpu().numpy(), k, n, log)
    line = {"metric": "exact_knn_queries_per_s", "value": qps, "unit": "queries/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n / qps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_knn(N, Q, d, k),
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": f"first {n} of {Q} queries against all {N} base vectors; compute_knn.cpp itself needs Faiss "
                                       "(absent here), so this is the oracle's restatement of its exact branch, OpenMP over the queries"},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def config_knn(N, Q, d, k):
    return {"workload": "compute_knn-exact-ground-truth", "N": N, "d": d, "Q": Q, "k": k, "metric": "L2",
            "generator": dict(GEN, chunk_rows=CHUNK, chunk_seed="43 * 1000003 + 1 + chunk"),
            "reference": "compute_knn.cpp:208-259 (IndexFlatL2.add + search in 10 000-row batches); --queries mode of bin/compute_knn",
            "l2_between_steps": "flushed (256 MiB write); the base (128 MB as bytes) is streamed once per 512 queries"}


def run_knn(args, rank, local, log):
    import torch
    import lira_ann_search_b200 as L
    dev = f"cuda:{local}"
    d, k, N, Q = 128, 100, args.N, args.Q
    x, x_q = knn_data(N, Q, d, dev)
    kn = L.KnnIndex(x, "L2")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = kn.search_dev(x_q, k, stream=stream.cuda_stream)
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        kn.search_dev(x_q, k, out=out, stream=stream.cuda_stream)
    torch.cuda.synchronize()
    launches0 = L.launch_count()
    clocks = Clocks(local)
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    clocks.mark_begin()
    t_wall = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record(stream)
        kn.search_dev(x_q, k, out=out, stream=stream.cuda_stream)
        ev[i][1].record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_wall
    clocks.mark_end()
    clk = clocks.stop()
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    launches = L.launch_count() - launches0
    scan_kind, redo = kn.last_scan_kind, kn.last_redo
    D, I = out[0].cpu().numpy(), out[1].cpu().numpy()
    # the reference program's own use: self-kNN with k + 1 = 11 (compute_knn.cpp:237), first 10 000 base rows as the batch
    self_ms = []
    o2 = kn.search_dev(x[:10000], 11, stream=stream.cuda_stream)
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        kn.search_dev(x[:10000], 11, out=o2, stream=stream.cuda_stream)
        e1.record(stream)
        e1.synchronize()
        self_ms.append(e0.elapsed_time(e1))
    self_ok = bool((o2[1][:, 0].cpu().numpy() == np.arange(10000)).mean() > 0.99)   # (exact duplicates may put a twin first)
    # end to end from host buffers: lira_knn_search uploads the queries and brings D / I back inside the timed region
    x_q_h = x_q.cpu().numpy()
    kn.search(x_q_h, k)
    e2e_s = []
    for _ in range(5):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Dh, Ih = kn.search(x_q_h, k)
        e2e_s.append(time.perf_counter() - t0)
    e2e_t = float(np.median(e2e_s))
    # check: the oracle on a query sample (bit-identical on integer data), and the CPU baseline in one go
    n = 64
    cpu_qps, cores, D_ref, I_ref = knn_cpu_baseline(x.cpu().numpy(), x_q_h, k, n, log)
    same = bool(np.array_equal(I[:n], I_ref) and np.array_equal(D[:n], D_ref) and np.array_equal(Ih[:n], I_ref))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf_peak = float(peaks.get("bf16_tflops", 1590.0))
    ms = total_ms / args.steps
    flops = 2.0 * N * Q * d
    ach = flops / (ms * 1e-3) / 1e12
    line = {
        "metric": "exact_knn_queries_per_s", "value": Q / (ms * 1e-3), "unit": "queries/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 operands, s32 accumulation (exact)" if scan_kind == "u8" else "f16 operands, f32 accumulation (exact on integer data)",
        "data": "synthetic", "config": config_knn(N, Q, d, k),
        "check": {"queries": n, "ids_and_distances_identical_to_oracle": same, "queries_redone_on_cuda_cores": int(redo)},
        "self_knn_k11": {"batch": 10000, "ms_per_batch": float(np.mean(self_ms)), "self_is_first": self_ok,
                         "whole_base_estimate_s": float(np.mean(self_ms)) * 1e-3 * N / 10000,
                         "tflops_algorithmic": 2.0 * N * 10000 * d / (float(np.mean(self_ms)) * 1e-3) / 1e12},
        "e2e": {"value": Q / e2e_t, "unit": "queries/s", "h2d_bytes_per_step": int(Q * d * 4), "d2h_bytes_per_step": int(Q * k * 12),
                "api": "lira_knn_search (host queries in, host D / I out; base resident behind the handle)"},
        "gpu_launches": int(launches), "clocks": clk,
        "roofline": {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak, "traffic": None,
                     "flops_algorithmic": flops, "kernel": f"u8_scan_kernel (tcgen05.mma kind::i8), scan kind {scan_kind}",
                     "what": "2 Q N d over the WHOLE search (seed pass over 16 base segments + filter pass + refine), against the measured "
                             "cuBLAS bf16 burst figure (the i8 pipe's nominal peak is twice that)",
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops" if peaks else "fallback"},
        "wall_s_timed_region": wall,
        "cpu_baseline": {"value": cpu_qps, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"first {n} of {Q} queries against all {N} base vectors (oracle restatement of compute_knn.cpp:208-259, "
                                   "OpenMP over the queries)"},
    }
    print(json.dumps(line), flush=True)

# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--prepare", action="store_true", help="build / cache the workload and the operating point, then exit")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--N", type=int, default=1_000_000)
    ap.add_argument("--Q", type=int, default=10_000)
    ap.add_argument("--B", type=int, default=1024)
    ap.add_argument("--recall", type=float, default=0.95)
    ap.add_argument("--cpu-sample", type=int, default=2000)
    ap.add_argument("--check", type=int, default=32, help="sharded workload: queries checked against brute force")
    ap.add_argument("--share", type=int, default=0, help="sharded workload: vectors per rank (multiple of 500 000) = weak scaling; "
                                                         "default 0 = the 100 M-vector dataset split over the ranks (strong scaling)")
    ap.add_argument("--workload", default=None, choices=["sift1m", "bigann", "knn"],
                    help="default: sift1m (config 1) on one GPU, bigann (config 5: 100 M vectors split over the ranks, lists striped) on several; "
                         "knn: config 2 (compute_knn exact ground truth, k = 100) on one GPU")
    ap.add_argument("--shard", default=None, choices=["queries", "lists"],
                    help="sift1m on N > 1: index replicas + sharded query stream (queries) or striped lists + NCCL merge (lists)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload is None:
        args.workload = "sift1m" if (world == 1 or args.shard is not None) else "bigann"
    if args.shard is None:
        args.shard = "queries" if args.workload == "sift1m" else "lists"
    log = (lambda *a: print(*a, file=sys.stderr, flush=True)) if rank == 0 else (lambda *a: None)
    if args.impl == "reference":
        if rank == 0:
            if args.workload == "knn":
                run_knn_reference(args, log)
            else:
                run_reference(args, world, log)
        return

    import torch
    import lira_ann_search_b200 as L
    L._cabi.require_gpu()
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    try:
        if args.workload == "knn":
            if rank == 0:
                run_knn(args, rank, local, log)
        elif args.workload == "bigann":
            run_shard(args, rank, world, local, dist, log)
        else:
            run_sift1m(args, rank, world, local, dist, log)
    finally:
        if dist is not None:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
