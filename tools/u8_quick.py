import sys, os, time
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import lira_ann_search_b200 as L
import oracle as O
from helpers import synth, random_lists
from test_gpu_parity import lists_csr
for metric in (O.L2, O.IP):
  for k, d in [(10, 128), (10, 96), (1, 20), (10, 200), (16, 256), (10, 8)]:
    rng = np.random.RandomState(17 + k + d)
    x_d, x_q = synth(30000, d, 700, seed=31 + d, integer=True)
    B = 24
    cl = random_lists(len(x_d), B, rng, redundancy=0.5, empty=(5,))
    cl[3] = cl[3][:7]
    off, ids, vecs = lists_csr(x_d, cl)
    nprobe = rng.randint(0, 7, len(x_q))
    poff = np.zeros(len(x_q) + 1, np.int64); np.cumsum(nprobe, out=poff[1:])
    pids = np.concatenate([rng.choice(B, n, replace=False) for n in nprobe] + [np.empty(0, int)]).astype(np.int32)
    index = L.LiraIndex.from_csr(x_d, off, ids, metric)
    D, I, cmp_ = index.search(x_q, poff, pids, k)
    I_ref, D_ref, cmp_ref = O.search(off, ids, vecs, x_q, poff, pids, k, metric, O.F64, 1)
    print(metric, k, d, index.last_path, index.last_redo, np.array_equal(I, I_ref), np.array_equal(D, D_ref), flush=True)
for k in (10, 17, 100):
    x_d, x_q = synth(70_000, 64, 300, seed=100 + k, integer=True)
    index = L.KnnIndex(x_d, O.L2)
    D, I = index.search(x_q, k)
    D_ref, I_ref = O.knn(x_d, x_q, k, O.L2, O.F64, 0)
    print('knn', k, index.last_path, index.last_redo, np.array_equal(I, I_ref), np.array_equal(D, D_ref), flush=True)
