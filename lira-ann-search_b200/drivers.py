"""The two LIRA drivers on the B200 library: the `__main__` flows of the reference's LIRA_smallscale.py (:246-379) and
LIRA_largescale.py (:184-354) with their `Config` dataclasses and argv (`--dataset sift --n_bkt 1024 --k 10 ...`).

The stages are the package's mirrors of the reference functions (same names, same call order): load_data ->
compute_data_knn -> build_kmeans_index -> get_knn_labels_data_only / get_knn_distr_redundancy -> get_scaled_dist ->
MLP_2_Input training (PyTorch, as in the reference) -> model_evaluate -> create_inner_indexes -> get_cmp_recall ->
query_tuning, before and after the redundancy assignment (mul_partition_by_model). Everything the reference hands to
faiss-cpu / scipy runs in liblira_b200 on the GPU.

Two additions, both optional: `init_centroids` / `init_model` warm-start the partitions and the probing model from files
(`.npy` centroids, a `state_dict` saved with torch.save), and `n_epoch = 0` then skips training -- that is how the tests run
the whole driver on the golden toy dataset and compare the emitted `..._tuning_threshold/model_{0,1}.csv` with the CSVs of
the reference's own run. And LIRA_largescale's `get_scaled_dist(xd_sub, x_q, kmeans, n_bkt)` call, which misses the `cfg`
argument in the reference (LIRA_largescale.py:258 against utils.py:120) and cannot run there, is made with `cfg`.
"""
from __future__ import annotations

import argparse
import dataclasses
import os
import time
from dataclasses import dataclass

import numpy as np

from . import engine, query, utils
from .utils import fprint


# ---------------------------------------------------------------------------------------------
# Config (LIRA_smallscale.py:27-75, LIRA_largescale.py:27-49)
# ---------------------------------------------------------------------------------------------
@dataclass
class Config:
    method_name: str = "LIRA_RE"
    dataset: str = None          # required, e.g. 'sift'
    data_path: str = "/data/vector_datasets"
    dis_metric: str = "L2"
    k: int = None                # required
    n_bkt: int = None            # required
    n_epoch: int = 10
    batch_size: int = 64
    n_mul: int = 2
    repa_step: int = 10
    redundancy_ratio: float = 0.03
    duplicate_type: str = "model"   # 'None' | 'model'
    pth_log: str = None
    file_name: str = None
    log_name: str = None
    df_name: str = None
    # warm start (not in the reference)
    init_centroids: str = None
    init_model: str = None
    lr: float = 0.0001

    def update(self):
        if self.dataset is None:
            raise ValueError("--dataset is required, e.g. --dataset sift")
        if self.k is None:
            raise ValueError("--k is required, e.g. --k 10")
        if self.n_bkt is None:
            raise ValueError("--n_bkt is required, e.g. --n_bkt 64")
        m = (self.dis_metric or "L2").lower()
        if m in ("l2", "euclidean", "euclidean_distance"):
            self.dis_metric = "L2"
        elif m in ("ip", "inner_product", "dot", "dot_product"):
            self.dis_metric = "inner_product"
        else:
            print(f"warning: unknown metric '{self.dis_metric}' kept as given; supported: 'L2', 'inner_product'")
        if self.pth_log is None:
            self.pth_log = f"./logs/{self.dataset}/ML_kmeans_RE_FLAT/"
        self.file_name = (f"{self.dataset}-k={self.k}-ML_kmeans={self.n_bkt}_FLAT_Metric={self.dis_metric}"
                          f"_ReType={self.duplicate_type}_ReRatio={self.redundancy_ratio}")
        self.log_name = f"{self.file_name}.txt"
        self.df_name = f"{self.file_name}.csv"


@dataclass
class LargeConfig:
    method_name: str = "LIRA_fullRE_subtrain"
    dataset: str = "deep50M"
    data_path: str = "/data/vector_datasets"
    dis_metric: str = None
    k: int = 100
    n_bkt: int = 1024
    n_epoch: int = 30
    batch_size: int = 512
    n_mul: int = 2
    repa_step: int = 1            # full redundancy
    duplicate_type: str = "model"
    redundancy_ratio: float = 1.0   # (query_tuning's log line prints it; every point is treated in the large-scale flow)
    pth_log: str = None
    file_name: str = None
    log_name: str = None
    df_name: str = None
    init_centroids: str = None
    init_model: str = None
    lr: float = 0.0001
    sub_fraction: float = 0.01    # LIRA_largescale.py:204: the model is trained on 1 % of the data
    batch_redundancy: int = 1_000_000   # LIRA_largescale.py:319

    def update(self):
        if self.dis_metric is None:
            self.dis_metric = "L2"
        if self.pth_log is None:
            self.pth_log = f"./logs/{self.dataset}/{self.method_name}_FLAT/"
        self.file_name = f"{self.dataset}-k={self.k}-ML_kmeans={self.n_bkt}_FLAT_ReType={self.duplicate_type}"
        self.log_name = f"{self.file_name}.txt"
        self.df_name = f"{self.file_name}.csv"


def parse_config(cls, argv=None):
    """`--field value` for every dataclass field, like HfArgumentParser(Config).parse_args_into_dataclasses()[0]."""
    ap = argparse.ArgumentParser()
    for f in dataclasses.fields(cls):
        typ = {"int": int, "float": float, "str": str}.get(f.type if isinstance(f.type, str) else f.type.__name__, str)
        ap.add_argument(f"--{f.name}", type=typ, default=f.default)
    ns = ap.parse_args(argv)
    cfg = cls(**vars(ns))
    cfg.update()
    return cfg


# ---------------------------------------------------------------------------------------------
# cal_metrics (LIRA_smallscale.py:99-143)
# ---------------------------------------------------------------------------------------------
PD_COLS = ["Epoch", "Accuracy", "Hit Rate", "nprobe predict", "nprobe target", "KNN Recall", "KNN Computations"]   # LIRA_smallscale.py:317


def cal_metrics(all_predicts, all_targets, epoch, knn_distr_id, cluster_id, results_df, loss, knn=100, fw=None):
    """Probing metrics of one epoch appended to results_df: accuracy, hit rate (TP / (TP + FN), nan rows skipped), mean
    predicted / target nprobe, kNN recall of the predicted probe sets (union of the ground-truth ids found in the probed
    partitions / k) and the reference's always-zero 'KNN Computations' column."""
    import pandas as pd
    import torch
    P, T = all_predicts.bool(), all_targets.bool()
    nprobe_pred = P.sum(1).float().mean().item()
    nprobe_tgt = T.sum(1).float().mean().item()
    accuracy = float((P == T).float().mean().item())
    hit = (P & T).sum(1).float() / T.sum(1).float()
    hit_rate = torch.nanmean(hit).item()
    member = knn_distr_id.member                      # [Q, k, n_mul] partitions holding each ground-truth id
    Pn = P.numpy()
    qq = np.arange(len(Pn))[:, None, None]
    found = ((member >= 0) & Pn[qq, np.where(member >= 0, member, 0)]).any(-1)   # id lies in a predicted partition
    ids = knn_distr_id.knn
    srt = np.sort(ids, 1)
    if (srt[:, 1:] == srt[:, :-1]).any():             # np.unique semantics for repeated ground-truth ids
        recalls = np.array([len(set(ids[q][found[q]].tolist())) for q in range(len(ids))]) / knn
    else:
        recalls = found.sum(1) / knn
    recall_avg, cmp_avg = float(np.mean(recalls)), 0.0
    fprint(f"| Epoch {epoch} | Loss {loss:.4f} | Accuracy {accuracy:.4f} | Hit Rate {hit_rate:.4f} | nprobe predict "
           f"{nprobe_pred:.4f} | nprobe target {nprobe_tgt:.4f} | KNN Recall {recall_avg:.4f} | KNN Computations {cmp_avg:.4f} |", fw)
    row = pd.DataFrame({"Epoch": [epoch], "Loss": [loss], "Accuracy": [accuracy], "Hit Rate": [hit_rate],
                        "nprobe predict": [nprobe_pred], "nprobe target": [nprobe_tgt], "KNN Recall": [recall_avg],
                        "KNN Computations": [cmp_avg]}).round(4)
    # (the reference starts from an empty frame without the 'Loss' column, so 'Loss' ends up LAST in the saved CSV)
    if results_df is None:
        results_df = pd.DataFrame(columns=PD_COLS)
    if len(results_df) == 0:
        return row[list(results_df.columns) + [c for c in row.columns if c not in results_df.columns]]
    return pd.concat([results_df, row], ignore_index=True)


# ---------------------------------------------------------------------------------------------
# shared stages
# ---------------------------------------------------------------------------------------------
def _partition(x, n_bkt, cfg, device):
    """build_kmeans_index (utils.py:321-330), or the assignment to given centroids when cfg.init_centroids is set."""
    if cfg.init_centroids:
        km = utils.Kmeans(x.shape[1], n_bkt, device=f"cuda:{device}")
        km.centroids = np.ascontiguousarray(np.load(cfg.init_centroids), np.float32)
        a = km.index.search(x, 1)[1].reshape(-1)
        cnts = np.bincount(a, minlength=n_bkt)
        order = np.argsort(a, kind="stable")
        bounds = np.zeros(n_bkt + 1, np.int64)
        np.cumsum(cnts, out=bounds[1:])
        return km, a.reshape(-1, 1), cnts, [order[bounds[b]:bounds[b + 1]].tolist() for b in range(n_bkt)]
    return utils.build_kmeans_index(x, n_bkt, device=f"cuda:{device}")


def _train(cfg, n_bkt, dim, train_loader, test_loader, knn_distr_id_query, cluster_ids, device, fw):
    import torch
    from .model_probing import MLP_2_Input, model_evaluate, model_train
    model = MLP_2_Input(input_dim1=n_bkt, input_dim2=dim, output_dim=n_bkt).to(device)
    if cfg.init_model:
        model.load_state_dict(torch.load(cfg.init_model, map_location=device))
    criterion = torch.nn.BCELoss()
    optimizer = torch.optim.Adam(model.parameters(), lr=cfg.lr)
    t0 = time.perf_counter()
    all_targets, all_predicts, loss_test, all_outputs = model_evaluate(model, test_loader, criterion, device)
    fprint(f"Epoch -1, Test Loss: {loss_test}, time_test: {time.perf_counter() - t0}", fw)
    results_df = cal_metrics(all_predicts, all_targets, -1, knn_distr_id_query, cluster_ids, None, loss_test, knn=cfg.k, fw=fw)
    for epoch in range(cfg.n_epoch):
        t0 = time.perf_counter()
        loss_train = model_train(model, train_loader, device, optimizer, criterion)
        t_train = time.perf_counter() - t0
        t0 = time.perf_counter()
        all_targets, all_predicts, loss_test, all_outputs = model_evaluate(model, test_loader, criterion, device)
        fprint(f"Epoch {epoch}, Train Loss: {loss_train}, Test Loss: {loss_test}, time_train: {t_train}, "
               f"time_test: {time.perf_counter() - t0}", fw)
        results_df = cal_metrics(all_predicts, all_targets, epoch, knn_distr_id_query, cluster_ids, results_df, loss_test,
                                 knn=cfg.k, fw=fw)
    return model, criterion, all_outputs, results_df


def _loaders(dist_d, x_d, labels_d, dist_q, x_q, labels_q, batch_size):
    import torch
    from torch.utils.data import DataLoader, TensorDataset
    tr = TensorDataset(torch.tensor(dist_d, dtype=torch.float32), torch.tensor(x_d, dtype=torch.float32),
                       torch.tensor(labels_d, dtype=torch.float32))
    te = TensorDataset(torch.tensor(dist_q, dtype=torch.float32), torch.tensor(x_q, dtype=torch.float32),
                       torch.tensor(labels_q, dtype=torch.float32))
    return DataLoader(tr, batch_size=batch_size, shuffle=False), DataLoader(te, batch_size=batch_size, shuffle=False)


# ---------------------------------------------------------------------------------------------
# LIRA_smallscale.py:246-379
# ---------------------------------------------------------------------------------------------
def run_smallscale(cfg: Config, device_index=None):
    import torch
    from .model_probing import model_evaluate
    n_bkt = cfg.n_bkt
    os.makedirs(cfg.pth_log, exist_ok=True)
    fw = open(cfg.pth_log + cfg.log_name, "a", encoding="utf-8")
    x_d, x_q, gt_ids = utils.load_data(cfg.dataset, data_path=cfg.data_path)
    if gt_ids is None:
        raise ValueError(f"Ground truth file not found for dataset {cfg.dataset}. Please ensure {cfg.dataset}_groundtruth.ivecs exists.")
    fprint(f">> dataset: {cfg.dataset}, data_sizes: {x_d.shape}, query_size: {x_q.shape}, n_bkt: {n_bkt}, knn: {cfg.k}, "
           f"metric: {cfg.dis_metric}", fw)
    n_d, dim = x_d.shape
    n_q = x_q.shape[0]
    dev_i = utils.get_idle_gpu() if device_index is None else device_index
    device = f"cuda:{dev_i}"
    fprint(f">> device: {device}", fw)
    fprint(f">> distance metric: {cfg.dis_metric}", fw)
    fprint(">> begin data preprocessing", fw)
    knn_data = utils.compute_data_knn(x_d, cfg, data_path=cfg.data_path, device=dev_i)
    knn_query = np.asarray(gt_ids[:, :cfg.k])
    fprint(f">> using precomputed ground truth with shape: {knn_query.shape}", fw)

    # (2) initial partitioning
    data_2_bkt = np.full((n_d, cfg.n_mul), -1)
    t0 = time.perf_counter()
    kmeans, single, cluster_cnts, cluster_ids = _partition(x_d, n_bkt, cfg, dev_i)
    data_2_bkt[:, :1] = single
    fprint(f">> build kmeans index time: {time.perf_counter() - t0}", fw)

    # (3) probing model
    t0 = time.perf_counter()
    labels_data = utils.get_knn_labels_data_only(knn_data, data_2_bkt, cfg)
    fprint(f">> get knn distribution time: {time.perf_counter() - t0}", fw)
    knn_distr_cnt_query, knn_distr_id_query = utils.get_knn_distr_redundancy(knn_query, data_2_bkt, cfg)
    labels_query = (knn_distr_cnt_query != 0).astype(np.uint8)
    t0 = time.perf_counter()
    dist_d, dist_q = utils.get_scaled_dist(x_d, x_q, kmeans, n_bkt, cfg, device=dev_i)
    t_dist = time.perf_counter() - t0
    fprint(f">> get scaled distance time: {t_dist}", fw)
    fprint(f">> get scaled distance time of queries: {t_dist * n_q / (n_q + n_d)}", fw)
    train_loader, test_loader = _loaders(dist_d, x_d, labels_data, dist_q, x_q, labels_query, cfg.batch_size)
    model, criterion, all_outputs, results_df = _train(cfg, n_bkt, dim, train_loader, test_loader, knn_distr_id_query,
                                                       cluster_ids, device, fw)

    # (4) redundancy with the probing model
    fprint(f">> begin redundancy with {cfg.duplicate_type}, metric: {cfg.dis_metric}", fw)
    if cfg.duplicate_type == "model":
        _, data_predicts, _, data_partition_score = model_evaluate(model, train_loader, criterion, device)
        nprobe_predicts = torch.sum(data_predicts, axis=1)
        xd_id_sorted_pre = torch.argsort(nprobe_predicts, descending=True, stable=True)
        n_redundancy = int(len(data_predicts) * cfg.redundancy_ratio)
        fprint(f">> redundancy ratio: {cfg.redundancy_ratio * 100:.1f}%, redundant vectors: {n_redundancy}/{len(data_predicts)}", fw)
        fprint(">> baseline (no redundancy) ...", fw)
        inner = utils.create_inner_indexes(x_d, cluster_ids, cfg, device=dev_i)
        search_time, cmp_all, found = query.get_cmp_recall(inner, x_q, cluster_ids, cfg)
        query.query_tuning(all_outputs, knn_distr_id_query, found, search_time, cmp_all, cfg, fw, part=0)
        query.mul_partition_by_model(data_partition_score, data_predicts, xd_id_sorted_pre, data_2_bkt, cluster_cnts, cluster_ids,
                                     begin=0, end=n_redundancy)
        knn_distr_cnt_query, knn_distr_id_query = utils.get_knn_distr_redundancy(knn_query, data_2_bkt, cfg)
        fprint(f">> after redundancy ({n_redundancy} vectors) ...", fw)
        inner = utils.create_inner_indexes(x_d, cluster_ids, cfg, device=dev_i)
        search_time, cmp_all, found = query.get_cmp_recall(inner, x_q, cluster_ids, cfg)
        query.query_tuning(all_outputs, knn_distr_id_query, found, search_time, cmp_all, cfg, fw, part=1)
    elif cfg.duplicate_type == "None":
        t0 = time.perf_counter()
        inner = utils.create_inner_indexes(x_d, cluster_ids, cfg, device=dev_i)
        fprint(f">> build flat index time: {time.perf_counter() - t0}", fw)
        t0 = time.perf_counter()
        search_time, cmp_all, found = query.get_cmp_recall(inner, x_q, cluster_ids, cfg)
        fprint(f">> search time: {time.perf_counter() - t0}", fw)
        query.query_tuning(all_outputs, knn_distr_id_query, found, search_time, cmp_all, cfg, fw)
    fprint("finish!", fw)
    results_df.to_csv(cfg.pth_log + cfg.df_name, index=False)
    fw.close()
    return {"data_2_bkt": data_2_bkt, "cluster_ids": cluster_ids, "all_outputs": all_outputs, "kmeans": kmeans, "model": model}


# ---------------------------------------------------------------------------------------------
# LIRA_largescale.py:184-354
# ---------------------------------------------------------------------------------------------
def run_largescale(cfg: LargeConfig, device_index=None):
    import torch
    n_bkt = cfg.n_bkt
    os.makedirs(cfg.pth_log, exist_ok=True)
    fw = open(cfg.pth_log + cfg.log_name, "a", encoding="utf-8")
    x_d, x_q, gt_ids = utils.load_data(cfg.dataset, data_path=cfg.data_path)
    if gt_ids is None:
        raise ValueError(f"Ground truth file not found for dataset {cfg.dataset}. Please ensure {cfg.dataset}_groundtruth.ivecs exists.")
    fprint(f">> dataset: {cfg.dataset}, data_sizes: {x_d.shape}, query_size: {x_q.shape}, n_bkt: {n_bkt}, knn: {cfg.k}", fw)
    n_d, dim = x_d.shape
    n_q = x_q.shape[0]
    dev_i = utils.get_idle_gpu() if device_index is None else device_index
    device = f"cuda:{dev_i}"
    fprint(f">> device: {device}", fw)

    # the training subset (LIRA_largescale.py:203-208)
    nd_sub = max(int(n_d * cfg.sub_fraction), 1)
    np.random.seed(43)
    sub_idx = np.random.choice(range(len(x_d)), nd_sub, replace=False)
    xd_sub = np.ascontiguousarray(x_d[sub_idx])
    fprint(">> begin data preprocessing", fw)
    knn_data_sub = utils.compute_data_knn(xd_sub, cfg, data_path=cfg.data_path, device=dev_i)
    # ground truth of the queries ON THE SUBSET, cached (LIRA_largescale.py:214-234)
    cache_dir = os.path.join(cfg.data_path, cfg.dataset, "knn_cache")
    os.makedirs(cache_dir, exist_ok=True)
    cache_file = os.path.join(cache_dir, f"{cfg.dataset}-query_on_subset_knn{cfg.k}-nsub{nd_sub}.npy")
    if not os.path.exists(cache_file):
        print("Computing query KNN on data subset...")
        index_flat = engine.KnnIndex(xd_sub, "inner_product" if cfg.dis_metric == "inner_product" else "L2", dev_i)
        _, knn_query_sub = index_flat.search(x_q, cfg.k)
        index_flat.close()
        np.save(cache_file, knn_query_sub)
        print(f"Cached query-on-subset KNN to: {cache_file}")
    else:
        knn_query_sub = np.load(cache_file).astype(int)
        print(f"Loaded cached query-on-subset KNN from: {cache_file}")

    # (2) initial partitioning of the subset
    data_2_bkt_sub = np.full((nd_sub, cfg.n_mul), -1)
    t0 = time.perf_counter()
    kmeans, single_sub, cluster_cnts, cluster_ids = _partition(xd_sub, n_bkt, cfg, dev_i)
    data_2_bkt_sub[:, :1] = single_sub
    fprint(f">> build kmeans index time: {time.perf_counter() - t0}", fw)

    # (3) probing model on the subset
    t0 = time.perf_counter()
    cnt_d_sub, _ = utils.get_knn_distr_redundancy(knn_data_sub, data_2_bkt_sub, cfg)
    fprint(f">> get knn distribution time: {time.perf_counter() - t0}", fw)
    cnt_q_sub, knn_distr_id_query_sub = utils.get_knn_distr_redundancy(knn_query_sub, data_2_bkt_sub, cfg)
    labels_data = np.where(cnt_d_sub != 0, 1, cnt_d_sub)
    labels_query = np.where(cnt_q_sub != 0, 1, cnt_q_sub)
    t0 = time.perf_counter()
    dist_d_sub, dist_q_sub = utils.get_scaled_dist(xd_sub, x_q, kmeans, n_bkt, cfg, device=dev_i)
    t_dist = time.perf_counter() - t0
    fprint(f">> get scaled distance time: {t_dist}", fw)
    fprint(f">> get scaled distance time of queries: {t_dist * n_q / (n_q + n_d)}", fw)
    train_loader, test_loader = _loaders(dist_d_sub, xd_sub, labels_data, dist_q_sub, x_q, labels_query, cfg.batch_size)
    model, criterion, all_outputs, results_df = _train(cfg, n_bkt, dim, train_loader, test_loader, knn_distr_id_query_sub,
                                                       cluster_ids, device, fw)

    # (4) partition the full data (LIRA_largescale.py:293-299)
    data_2_bkt = np.full((n_d, cfg.n_mul), -1)
    single = kmeans.index.search(x_d, 1)[1]
    data_2_bkt[:, :1] = single
    flat = single.reshape(-1)
    cluster_cnts = np.bincount(flat, minlength=n_bkt)
    order = np.argsort(flat, kind="stable")
    bounds = np.zeros(n_bkt + 1, np.int64)
    np.cumsum(cluster_cnts, out=bounds[1:])
    cluster_ids = [order[bounds[b]:bounds[b + 1]].tolist() for b in range(n_bkt)]
    knn_query = np.asarray(gt_ids[:, :cfg.k])
    fprint(f">> using precomputed ground truth with shape: {knn_query.shape}", fw)
    _, knn_distr_id_query = utils.get_knn_distr_redundancy(knn_query, data_2_bkt, cfg)

    fprint(f">> begin redundancy with {cfg.duplicate_type}", fw)
    if cfg.duplicate_type == "model":
        t0 = time.perf_counter()
        inner = utils.create_inner_indexes(x_d, cluster_ids, cfg, device=dev_i)
        fprint(f">> build flat index time: {time.perf_counter() - t0}", fw)
        _, cmp_all, found = query.get_cmp_recall(inner, x_q, cluster_ids, cfg)
        query.query_tuning_large(all_outputs, knn_distr_id_query, found, cmp_all, cfg)
        del inner
        # every point gets its partitions from the model, in batches (LIRA_largescale.py:319-329). Each batch stays on the device:
        # scaler statistics of THIS batch (utils.py:182-215) -> standardised centroid distances -> the model's forward (PyTorch,
        # as model_infer does) -> the redundancy rule (lira_mul_partition_dev); only the partition columns come back
        cent_dev = torch.as_tensor(np.ascontiguousarray(kmeans.centroids, np.float32), device=device)
        model.eval()
        for start_idx in range(0, n_d, cfg.batch_redundancy):
            end_idx = min(start_idx + cfg.batch_redundancy, n_d)
            x_b = torch.as_tensor(np.ascontiguousarray(x_d[start_idx:end_idx], np.float32), device=device)
            mean, var = engine.feature_stats_dev(x_b, cent_dev)
            scale = np.sqrt(var)
            scale[scale < 10 * np.finfo(np.float64).eps] = 1.0
            feats = engine.centroid_features_dev(x_b, cent_dev, torch.as_tensor(mean.astype(np.float32), device=device),
                                                 torch.as_tensor(scale.astype(np.float32), device=device))
            with torch.no_grad():
                score = torch.cat([model(feats[a:a + 65536], x_b[a:a + 65536]) for a in range(0, len(x_b), 65536)])
            query.mul_partition_by_model_large(score, None, np.arange(start_idx, end_idx), start_idx, data_2_bkt, cluster_cnts, cluster_ids)
            del x_b, feats, score
        _, knn_distr_id_query = utils.get_knn_distr_redundancy(knn_query, data_2_bkt, cfg)
        inner = utils.create_inner_indexes(x_d, cluster_ids, cfg, device=dev_i)
        _, cmp_all, found = query.get_cmp_recall(inner, x_q, cluster_ids, cfg)
        print(">> after redundancy")
        query.query_tuning_large(all_outputs, knn_distr_id_query, found, cmp_all, cfg, part=1)
    elif cfg.duplicate_type == "None":
        t0 = time.perf_counter()
        inner = utils.create_inner_indexes(x_d, cluster_ids, cfg, device=dev_i)
        fprint(f">> build flat index time: {time.perf_counter() - t0}", fw)
        t0 = time.perf_counter()
        _, cmp_all, found = query.get_cmp_recall(inner, x_q, cluster_ids, cfg)
        fprint(f">> search time: {time.perf_counter() - t0}", fw)
        query.query_tuning_large(all_outputs, knn_distr_id_query, found, cmp_all, cfg)
    fprint("finish!", fw)
    results_df.to_csv(cfg.pth_log + cfg.df_name, index=False)
    fw.close()
    return {"data_2_bkt": data_2_bkt, "cluster_ids": cluster_ids, "all_outputs": all_outputs, "kmeans": kmeans, "model": model}
