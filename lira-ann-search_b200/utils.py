"""Host-side mirror of the reference's utils.py for the query phase and ground-truth path.
Same function names, argument meaning and return shapes; the arithmetic that the reference hands
to faiss-cpu / scipy goes to liblira_b200 instead. Citations are into /root/reference/utils.py.

Index construction: K-Means runs in the library (lira_kmeans_train: assignment on the kNN path,
update kernel), the scaler statistics and the redundancy rule have device forms (engine.feature_stats_dev,
engine.mul_partition_dev); label building stays numpy as in the reference.
"""
from __future__ import annotations

import glob
import os
import time

import numpy as np

from . import engine

SEED = 43  # utils.py:15-21


# ---------------------------------------------------------------------------------------------
# I/O (a13) -- utils.py:23-88
# ---------------------------------------------------------------------------------------------
def read_xvecs(file_path, dtype="float32"):
    """.fvecs / .ivecs / .bvecs reader: each record is int32 d followed by d values."""
    if not os.path.exists(file_path):
        raise FileNotFoundError(f"File not found: {file_path}")
    dt = np.dtype(dtype)
    if dt.itemsize == 4:
        raw = np.memmap(file_path, dtype="int32", mode="r")
        d = int(raw[0])
        return raw.view(dt).reshape(-1, d + 1)[:, 1:]
    raw = np.memmap(file_path, dtype="uint8", mode="r")  # .bvecs: 4-byte header + d bytes
    d = int(raw[:4].view("int32")[0])
    return raw.reshape(-1, d + 4)[:, 4:]


def write_xvecs(file_path, arr):
    arr = np.ascontiguousarray(arr)
    if arr.dtype.itemsize != 4:
        raise ValueError("write_xvecs handles 4-byte element types (fvecs / ivecs)")
    n, d = arr.shape
    rec = np.empty((n, d + 1), np.int32)
    rec[:, 0] = d
    rec[:, 1:] = arr.view(np.int32)
    rec.tofile(file_path)


def load_data(dataset_name, data_path="/data/vector_datasets"):
    """-> (x_d, x_q, gt_ids or None); {ds}_base.fvecs (fallback _learn), _query.fvecs, _groundtruth.ivecs."""
    ds_dir = os.path.join(data_path, dataset_name)
    base = os.path.join(ds_dir, f"{dataset_name}_base.fvecs")
    if not os.path.exists(base):
        base = os.path.join(ds_dir, f"{dataset_name}_learn.fvecs")
    x_d = np.ascontiguousarray(read_xvecs(base, "float32"))
    x_q = np.ascontiguousarray(read_xvecs(os.path.join(ds_dir, f"{dataset_name}_query.fvecs"), "float32"))
    gt_file = os.path.join(ds_dir, f"{dataset_name}_groundtruth.ivecs")
    gt = read_xvecs(gt_file, "int32") if os.path.exists(gt_file) else None
    print(f"Loaded dataset '{dataset_name}':\n  Base vectors: {x_d.shape}\n  Query vectors: {x_q.shape}")
    if gt is not None:
        print(f"  Ground truth: {gt.shape}")
    return x_d, x_q, gt


def fprint(message, file=None):
    print(message)
    if file:
        print(message, file=file)


# ---------------------------------------------------------------------------------------------
# centroid-distance features (a1, a2) -- utils.py:98-180
# ---------------------------------------------------------------------------------------------
def get_dist_cid(data, kmeans, n_bkt, device=0):
    """Euclidean distance of every row to every centroid, fp32 [n, n_bkt] (utils.py:98-118)."""
    return engine.centroid_features(data, kmeans.centroids, device=device)


class _Scaler:
    """The two StandardScaler fields the reference persists (utils.py:171-175)."""

    def __init__(self, mean_, scale_):
        self.mean_, self.scale_ = mean_, scale_


def _fit_scaler(x_d, kmeans, n_bkt, device, batch=65536):
    # fp64 mean / variance over all data rows, as StandardScaler.fit / partial_fit accumulate them
    n = x_d.shape[0]
    s1 = np.zeros(n_bkt, np.float64)
    s2 = np.zeros(n_bkt, np.float64)
    for a in range(0, n, batch):
        dist = get_dist_cid(x_d[a:a + batch], kmeans, n_bkt, device).astype(np.float64)
        s1 += dist.sum(0)
        s2 += (dist * dist).sum(0)
    mean = s1 / n
    var = np.maximum(s2 / n - mean * mean, 0.0)
    scale = np.sqrt(var)
    scale[scale < 10 * np.finfo(np.float64).eps] = 1.0  # sklearn _handle_zeros_in_scale
    return _Scaler(mean, scale)


def get_scaled_dist(x_d, x_q, kmeans, n_bkt, cfg=None, device=0, return_data=True):
    """-> (distances_data_scaled [n_d,B] f32, distances_query_scaled [n_q,B] f32); side effect: writes
    {pth_log}/{file_name}_scaler_mean.npy / _scaler_scale.npy (utils.py:120-180).
    return_data=False skips materialising the n_d x B matrix (query phase only)."""
    scaler = _fit_scaler(x_d, kmeans, n_bkt, device)
    mean32, scale32 = scaler.mean_.astype(np.float32), scaler.scale_.astype(np.float32)
    q_scaled = engine.centroid_features(x_q, kmeans.centroids, mean32, scale32, device)
    d_scaled = engine.centroid_features(x_d, kmeans.centroids, mean32, scale32, device) if return_data else None
    if cfg is not None:   # (LIRA_largescale.py:258 calls without cfg, which the reference's signature does not allow)
        os.makedirs(cfg.pth_log, exist_ok=True)
        np.save(os.path.join(cfg.pth_log, f"{cfg.file_name}_scaler_mean.npy"), mean32)
        np.save(os.path.join(cfg.pth_log, f"{cfg.file_name}_scaler_scale.npy"), scale32)
    return d_scaled, q_scaled


def get_scaled_dist_data(x_d, kmeans, n_bkt, device=0):
    """Standardised centroid distances of one batch of data rows with a scaler fitted ON THAT BATCH (utils.py:182-215,
    called once per 1 000 000-row redundancy batch at LIRA_largescale.py:323): every batch is standardised with its own
    mean / std, not with the scaler the model was trained with -- reproduced as the reference does it."""
    scaler = _fit_scaler(x_d, kmeans, n_bkt, device)
    return engine.centroid_features(x_d, kmeans.centroids, scaler.mean_.astype(np.float32), scaler.scale_.astype(np.float32),
                                    device)


# ---------------------------------------------------------------------------------------------
# ground truth / self-kNN (a11) -- utils.py:223-319
# ---------------------------------------------------------------------------------------------
def compute_data_knn(x_data, cfg, data_path="/data/vector_datasets", device=0):
    """Cache lookup order as the reference (newest *_ivf_nprobe*.bin, exact .bin, .npy), else exact
    brute force on the GPU with k+1 neighbours and column 0 dropped (utils.py:293-310)."""
    cache_dir = os.path.join(data_path, cfg.dataset, "knn_cache")
    os.makedirs(cache_dir, exist_ok=True)
    n = len(x_data)
    for pattern in (f"{cfg.dataset}-data_self_knn{cfg.k}-n{n}_ivf_nprobe*.bin",
                    f"{cfg.dataset}-data_self_knn{cfg.k}-n{n}.bin"):
        hits = glob.glob(os.path.join(cache_dir, pattern))
        if hits:
            path = max(hits, key=os.path.getctime)
            print(f"Loading precomputed KNN from cache: {os.path.basename(path)}")
            return np.fromfile(path, dtype=np.int32).reshape(n, cfg.k)
    npy = os.path.join(cache_dir, f"{cfg.dataset}-data_self_knn{cfg.k}-n{n}.npy")
    if os.path.exists(npy):
        return np.load(npy).astype(int)
    t0 = time.time()
    knn = np.zeros((n, cfg.k), np.int32)
    batch = min(10000, n)  # utils.py:301
    metric = "inner_product" if cfg.dis_metric == "inner_product" else "L2"
    index = engine.KnnIndex(x_data, metric, device)   # index.add(x_data): uploaded once (utils.py:294-298)
    for a in range(0, n, batch):
        _, ids = index.search(x_data[a:a + batch], cfg.k + 1)
        knn[a:a + batch] = ids[:, 1:cfg.k + 1]
    index.close()
    print(f"KNN computation completed in {time.time() - t0:.2f}s")
    np.save(npy, knn)
    return knn


# ---------------------------------------------------------------------------------------------
# partitions (build side; needed to make synthetic indexes) -- utils.py:321-330
# ---------------------------------------------------------------------------------------------
class Kmeans:
    """faiss.Kmeans with the fields the reference reads (.centroids, .index): Lloyd on at most 256 * k sampled points, niter
    iterations, L2 assignment (utils.py:321-325). Runs in liblira_b200: the assignment step is the exact 1-NN of the kNN path
    against the centroid table, the update step a segmented mean on the device (lira_kmeans_train)."""

    def __init__(self, d, k, niter=20, verbose=False, seed=1234, device="cuda:0"):
        self.d, self.k, self.niter, self.verbose, self.seed, self.device = d, k, niter, verbose, seed, device
        self.centroids = None

    @property
    def _dev_index(self):
        return int(str(self.device).split(":")[1]) if ":" in str(self.device) else 0

    @property
    def index(self):
        """faiss.Kmeans.index: a flat L2 index over the centroids; `.search(x, 1)` is the partition assignment the
        reference uses (utils.py:325, LIRA_largescale.py:294). Exact kNN on the GPU (engine.KnnIndex)."""
        return _CentroidIndex(self)

    def train(self, x, init_centroids=None):
        self.centroids = engine.kmeans_train(x, self.k, self.niter, self.seed, init_centroids, self._dev_index)
        return self


class _CentroidIndex:
    def __init__(self, km):
        self.km = km

    @property
    def ntotal(self):
        return len(self.km.centroids)

    def search(self, x, k):
        dev_i = self.km._dev_index
        x = np.ascontiguousarray(x, np.float32)
        idx = engine.KnnIndex(np.ascontiguousarray(self.km.centroids, np.float32), "L2", dev_i)
        D = np.empty((len(x), k), np.float32)
        I = np.empty((len(x), k), np.int64)
        for a in range(0, len(x), 1 << 20):
            D[a:a + (1 << 20)], I[a:a + (1 << 20)] = idx.search(x[a:a + (1 << 20)], k)
        idx.close()
        return D, I


def build_kmeans_index(x_data, n_bkt, device="cuda:0"):
    """-> (kmeans, data_2_bkt [n,1], cluster_cnts [B], cluster_ids list[B] of id lists) (utils.py:321-330)."""
    n_d, dim = x_data.shape
    kmeans = Kmeans(dim, n_bkt, niter=20, device=device).train(x_data)
    a = kmeans.index.search(x_data, 1)[1].reshape(-1)
    cluster_cnts = np.bincount(a, minlength=n_bkt)
    order = np.argsort(a, kind="stable")
    bounds = np.zeros(n_bkt + 1, np.int64)
    np.cumsum(cluster_cnts, out=bounds[1:])
    cluster_ids = [order[bounds[b]:bounds[b + 1]].tolist() for b in range(n_bkt)]
    return kmeans, a.reshape(-1, 1), cluster_cnts, cluster_ids


# ---------------------------------------------------------------------------------------------
# recall helpers (a9) -- utils.py:354-405
# ---------------------------------------------------------------------------------------------
class KnnDistrIds:
    """knn_distr_id of get_knn_distr_redundancy without Q*B python lists. `member[q, j, c]` is the bucket
    of the c-th copy of ground-truth id knn[q, j] (-1 = none). Indexing [q][b] (or [q, b]) returns the list
    the reference would hold there (ids repeated once per matching copy, utils.py:375-377)."""

    def __init__(self, knn, member):
        self.knn, self.member = np.asarray(knn), np.asarray(member)

    def ids(self, q, b):
        hit = self.member[q] == b
        return np.repeat(self.knn[q], hit.shape[1])[hit.reshape(-1)].tolist()

    def __getitem__(self, key):
        if isinstance(key, tuple):
            return self.ids(*key)
        return _Row(self, key)


class _Row:
    def __init__(self, parent, q):
        self.parent, self.q = parent, q

    def __getitem__(self, b):
        return self.parent.ids(self.q, b)


def get_knn_distr_redundancy(knn, data_2_bkt, cfg):
    """-> (knn_distr_cnt [n, B] int, knn_distr_id) over all n_mul columns, ignoring -1 (utils.py:354-379)."""
    knn = np.asarray(knn)
    d2b = np.asarray(data_2_bkt)
    member = d2b[knn]  # [n, k, n_mul]
    n = knn.shape[0]
    cnt = np.zeros((n, cfg.n_bkt), dtype=int)
    flat = member.reshape(n, -1)
    rows = np.repeat(np.arange(n), flat.shape[1])
    ok = flat.reshape(-1) >= 0
    np.add.at(cnt, (rows[ok], flat.reshape(-1)[ok]), 1)
    return cnt, KnnDistrIds(knn, member)


def get_knn_labels_data_only(knn, data_2_bkt, cfg):
    """labels[i, b] = 1 iff sample i has a kNN in bucket b (utils.py:381-405)."""
    cnt, _ = get_knn_distr_redundancy(knn, data_2_bkt, cfg)
    return (cnt != 0).astype(np.uint8)


# ---------------------------------------------------------------------------------------------
# inverted lists (a6) -- utils.py:407-429
# ---------------------------------------------------------------------------------------------
def create_flat_indexes(x_d, xd_id_bkts, cfg, dis_metric: str = "L2", device=0):
    """One flat (exact) index per bucket. All buckets live in ONE device-resident LiraIndex; the returned
    list holds per-bucket views with the faiss surface the drivers use (.search, .ntotal)."""
    index = engine.LiraIndex.from_cluster_ids(x_d, xd_id_bkts, dis_metric, device)
    return index.views()


def create_inner_indexes(x_d, cluster_ids, cfg, device=0):
    return create_flat_indexes(x_d, cluster_ids, cfg, dis_metric=cfg.dis_metric, device=device)


def get_idle_gpu():
    """Index of the GPU with the least memory in use (utils.py:90-96 asks nvidia-smi; NVML is the library behind it)."""
    import pynvml
    pynvml.nvmlInit()
    used = [pynvml.nvmlDeviceGetMemoryInfo(pynvml.nvmlDeviceGetHandleByIndex(i)).used for i in range(pynvml.nvmlDeviceGetCount())]
    return int(np.argmin(used))


def per_query(all_outputs, knn_distr_cnt_query, cluster_cnts, n_bkt, cfg, nq_test=100, recall_target=0.98):
    """Smallest top-nprobe (1..19) reaching the recall target per query (utils.py:502-519). Returns a
    DataFrame with columns q_id, nprobe, cmp and writes the reference's per-query CSV."""
    import pandas as pd
    import torch
    all_outputs = torch.as_tensor(np.asarray(all_outputs))
    rows = []
    for q_id in range(min(nq_test, len(all_outputs))):
        nprobe, cmp_ = 0, 0
        for probe_m in range(1, 20):
            mb = all_outputs[q_id].topk(probe_m).indices.cpu().numpy()
            if knn_distr_cnt_query[q_id, mb].sum() / cfg.k >= recall_target:
                nprobe, cmp_ = probe_m, int(np.asarray(cluster_cnts)[mb].sum())
                break
        rows.append({"q_id": q_id, "nprobe": nprobe, "cmp": cmp_})
    df = pd.DataFrame(rows, columns=["q_id", "nprobe", "cmp"])
    df.to_csv(cfg.pth_log + f"{cfg.dataset}-k={cfg.k}-ML_kmeans={n_bkt}_perquery.csv", index=False)
    return df
