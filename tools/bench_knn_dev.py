#!/usr/bin/env python
"""Device-resident timing of the exact kNN handle (lira_knn_create_dev / lira_knn_search_dev) on SIFT1M-shape synthetic data:
tools/bench_knn_dev.py [N] [Q] [k] -- CUDA events around search_dev, base and queries already in HBM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lira_ann_search_b200 as L
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
g = torch.Generator(device="cuda"); g.manual_seed(0)
base = torch.randint(0, 256, (N, 128), device="cuda", generator=g).float()
q = torch.randint(0, 256, (Q, 128), device="cuda", generator=g).float()
idx = L.KnnIndex(base, "L2")
for it in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    D, I = idx.search_dev(q, k, stream=torch.cuda.current_stream().cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"knn dev N={N} Q={Q} k={k}: {ms:.2f} ms, {2*N*Q*128/ms/1e9:.1f} TFLOP/s algorithmic, path {idx.last_path} redo {idx.last_redo}", flush=True)
b = base[:200000]; qq = q[:512]
d = (qq * qq).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2 * qq @ b.T
ref = torch.topk(d, k, largest=False)
i2 = L.KnnIndex(b.contiguous(), "L2")
Dm, Im = i2.search_dev(qq.contiguous(), k)
print("check vs torch (200k x 512): ids equal", bool((torch.sort(ref.indices, 1).values == torch.sort(Im, 1).values).float().mean() > 0.999))
