#!/usr/bin/env python
"""Time lira_knn (exact kNN, compute_knn path) on SIFT1M-shape synthetic data: tools/bench_knn.py [N] [Q] [k]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lira_ann_search_b200 as L
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 11
rng = np.random.RandomState(0)
base = rng.randint(0, 256, (N, 128)).astype(np.float32)
q = rng.randint(0, 256, (Q, 128)).astype(np.float32)
for it in range(3):
    t0 = time.perf_counter()
    D, I = L.knn(base, q, k, "L2")
    dt = time.perf_counter() - t0
    print(f"knn N={N} Q={Q} k={k}: {dt*1e3:.1f} ms end to end (host buffers), {2*N*Q*128/dt/1e12:.2f} TFLOP/s algorithmic")
import torch
b = torch.as_tensor(base[:200000], device="cuda"); qq = torch.as_tensor(q[:512], device="cuda")
d = (qq * qq).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2 * qq @ b.T
Dm, Im = L.knn(base[:200000], q[:512], k, "L2")
ref = torch.topk(d, k, largest=False)
print("check vs torch (200k x 512): ids equal", bool((torch.sort(ref.indices, 1).values.cpu().numpy() == np.sort(Im, 1)).mean() > 0.999))
