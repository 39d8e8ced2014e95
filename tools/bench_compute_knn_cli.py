#!/usr/bin/env python
"""bin/compute_knn on a SIFT1M-shape dataset (1 M x 128 integer-valued vectors), the reference's own use: self-kNN k = 10
written as {ds}/knn_cache/{ds}-data_self_knn10-n1000000.bin. Prints the program's timings and checks a sample of rows
against a brute-force scan.   python tools/bench_compute_knn_cli.py [N] [k]"""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lira_ann_search_b200 as L
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rng = np.random.RandomState(0)
centres = rng.randn(1024, 128).astype(np.float32)
x = np.clip(np.round(24 * (centres[rng.randint(0, 1024, N)] + 0.5 * rng.randn(N, 128).astype(np.float32)) + 100), 0, 255).astype(np.float32)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with tempfile.TemporaryDirectory() as td:
    os.makedirs(os.path.join(td, "syn"))
    L.write_xvecs(os.path.join(td, "syn", "syn_base.fvecs"), x)
    t0 = time.time()
    r = subprocess.run([os.path.join(root, "bin", "compute_knn"), "syn", td, str(k), "0"], capture_output=True, text=True)
    wall = time.time() - t0
    print(r.stdout[-600:], r.stderr[-300:])
    knn = np.fromfile(os.path.join(td, "syn", "knn_cache", f"syn-data_self_knn{k}-n{N}.bin"), dtype=np.int32).reshape(N, k)
xb = torch.as_tensor(x, device="cuda")
rows = rng.choice(N, 256, replace=False)
q = xb[rows]
d = (q * q).sum(1)[:, None] + (xb * xb).sum(1)[None, :] - 2 * q @ xb.T
ref = torch.topk(d, k + 1, largest=False)
dist_mine = torch.gather(d, 1, torch.as_tensor(knn[rows].astype(np.int64), device="cuda"))
ok = bool(torch.allclose(torch.sort(dist_mine, 1).values, torch.sort(ref.values[:, 1:], 1).values))
print(f"compute_knn N={N} k={k}: wall {wall:.2f}s (file read included); sampled rows have the brute-force distances: {ok}")
