"""TEST INFRASTRUCTURE -- generate tests/golden/*.npz by running the REFERENCE'S OWN code.

Runs only in the build container (needs /root/reference). Two sources of truth are captured:

  (1) the reference's Python hot path, imported UNMODIFIED from /root/reference with the
      fake-faiss / prettytable shims on sys.path (oracle/ref_shim): utils.build_kmeans_index,
      utils.get_scaled_dist, model_probing.MLP_2_Input + model_train + model_evaluate,
      utils.get_knn_distr_redundancy, LIRA_smallscale.mul_partition_by_model,
      utils.create_inner_indexes, LIRA_smallscale.get_cmp_recall, LIRA_smallscale.query_tuning;
  (2) the reference's C++ search.cpp compiled unmodified (oracle/_ref/search_ref, see
      oracle/Makefile), run on artifacts written the way index.py:144-192 writes them.

Usage:  python oracle/make_golden.py            (writes tests/golden/toy_l2.npz, toy_ip.npz)
The fixtures are committed; the GPU box never needs /root/reference.
"""
import io
import os
import re
import subprocess
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("LIRA_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "ref_shim"))
sys.path.insert(0, HERE)
sys.path.append(REF)

import oracle as O  # noqa: E402


def synth(n, d, nq, seed, integer=True, ncomp=40, sigma=0.9):
    """Small Gaussian-mixture dataset in the style of SURVEY.md 8(d): integer-valued like SIFT
    when integer=True (creates exact distance ties on purpose)."""
    rng = np.random.RandomState(seed)
    centres = rng.randn(ncomp, d)
    w = rng.lognormal(0, 0.5, ncomp)
    w /= w.sum()

    def draw(m):
        c = rng.choice(ncomp, m, p=w)
        x = centres[c] + sigma * rng.randn(m, d)
        if integer:
            return np.clip(np.round(32 * x + 64), 0, 255).astype(np.float32)
        return x.astype(np.float32)

    return draw(n), draw(nq)


def write_xvecs(path, arr, dtype):
    arr = np.ascontiguousarray(arr, dtype)
    n, d = arr.shape
    out = np.empty((n, d + 1), dtype=np.int32)
    out[:, 0] = d
    out[:, 1:] = arr.view(np.int32)
    out.tofile(path)


def run_reference_python(x_d, x_q, gt, n_bkt, k, metric, workdir, seed=0, epochs=3):
    import torch

    import LIRA_smallscale as S  # reference, unmodified
    import model_probing as MP  # reference, unmodified
    import utils as U  # reference, unmodified

    torch.manual_seed(seed)
    np.random.seed(seed)
    cfg = S.Config(dataset="toy", k=k, n_bkt=n_bkt, dis_metric=metric, redundancy_ratio=0.25, batch_size=64)
    cfg.update()
    cfg.pth_log = os.path.join(workdir, "logs") + "/"
    os.makedirs(cfg.pth_log, exist_ok=True)
    fw = io.StringIO()
    S.fw = fw  # cal_metrics reads a module global (LIRA_smallscale.py:129)

    n_d, dim = x_d.shape
    # self-kNN for training labels: utils.compute_data_knn fallback arithmetic (utils.py:293-310)
    index = (U.faiss.IndexFlatIP if metric == "inner_product" else U.faiss.IndexFlatL2)(dim)
    index.add(x_d)
    _, knn_self = index.search(x_d, k + 1)
    knn_data = knn_self[:, 1:k + 1].astype(np.int32)
    knn_query = gt[:, :k]

    data_2_bkt = np.full((n_d, cfg.n_mul), -1)
    kmeans, single, cluster_cnts, cluster_ids = U.build_kmeans_index(x_d, n_bkt)
    data_2_bkt[:, :1] = single
    labels_data = U.get_knn_labels_data_only(knn_data, data_2_bkt, cfg)
    cnt_q0, ids_q0 = U.get_knn_distr_redundancy(knn_query, data_2_bkt, cfg)
    labels_query = (cnt_q0 != 0).astype(np.uint8)
    dist_d, dist_q = U.get_scaled_dist(x_d, x_q, kmeans, n_bkt, cfg)
    mean32 = np.load(os.path.join(cfg.pth_log, f"{cfg.file_name}_scaler_mean.npy"))
    scale32 = np.load(os.path.join(cfg.pth_log, f"{cfg.file_name}_scaler_scale.npy"))

    from torch.utils.data import DataLoader, TensorDataset
    tr = TensorDataset(torch.tensor(dist_d), torch.tensor(x_d), torch.tensor(labels_data, dtype=torch.float32))
    te = TensorDataset(torch.tensor(dist_q), torch.tensor(x_q), torch.tensor(labels_query, dtype=torch.float32))
    trl = DataLoader(tr, batch_size=cfg.batch_size, shuffle=False)
    tel = DataLoader(te, batch_size=cfg.batch_size, shuffle=False)
    model = MP.MLP_2_Input(input_dim1=n_bkt, input_dim2=dim, output_dim=n_bkt)
    crit = torch.nn.BCELoss()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    for _ in range(epochs):
        MP.model_train(model, trl, "cpu", opt, crit)
    _, all_predicts, _, all_outputs = MP.model_evaluate(model, tel, crit, "cpu")

    out = {}
    # ---- part 0: no redundancy
    idx0 = U.create_inner_indexes(x_d, cluster_ids, cfg)
    _, cmp0, found0 = S.get_cmp_recall(idx0, x_q, cluster_ids, cfg)
    S.query_tuning(all_outputs, ids_q0, found0, np.ones_like(cmp0, dtype=float), cmp0, cfg, fw, part=0)
    out["lists0_off"], out["lists0_ids"], _ = O.build_lists_from_cluster_ids(x_d, cluster_ids)
    out["d2b0"] = data_2_bkt.astype(np.int32).copy()
    out["found0"], out["cmp0"] = found0.astype(np.int64), cmp0.astype(np.int64)

    # ---- redundancy (LIRA_smallscale.py:331-354) then part 1
    _, data_predicts, _, data_score = MP.model_evaluate(model, trl, crit, "cpu")
    order = torch.argsort(torch.sum(data_predicts, axis=1), descending=True)
    n_red = int(n_d * cfg.redundancy_ratio)
    # inputs of the redundancy assignment, for the vectorised mirror (lira_ann_search_b200.query.mul_partition_by_model)
    out["red_score"] = data_score.numpy().astype(np.float32)
    out["red_order"] = order.numpy().astype(np.int64)
    out["red_end"] = np.int64(n_red)
    out["red_cnts_before"] = np.asarray(cluster_cnts, np.int64).copy()
    S.mul_partition_by_model(data_score, data_predicts, order, data_2_bkt, cluster_cnts, cluster_ids, begin=0, end=n_red)
    out["red_cnts_after"] = np.asarray(cluster_cnts, np.int64).copy()
    cnt_q1, ids_q1 = U.get_knn_distr_redundancy(knn_query, data_2_bkt, cfg)
    idx1 = U.create_inner_indexes(x_d, cluster_ids, cfg)
    _, cmp1, found1 = S.get_cmp_recall(idx1, x_q, cluster_ids, cfg)
    S.query_tuning(all_outputs, ids_q1, found1, np.ones_like(cmp1, dtype=float), cmp1, cfg, fw, part=1)
    out["lists1_off"], out["lists1_ids"], _ = O.build_lists_from_cluster_ids(x_d, cluster_ids)
    out["d2b1"] = data_2_bkt.astype(np.int32).copy()
    out["found1"], out["cmp1"] = found1.astype(np.int64), cmp1.astype(np.int64)
    out["knn_cnt0"], out["knn_cnt1"] = cnt_q0.astype(np.int64), cnt_q1.astype(np.int64)

    import pandas as pd
    for part in (0, 1):
        df = pd.read_csv(os.path.join(cfg.pth_log, cfg.file_name + "_tuning_threshold", f"{cfg.duplicate_type}_{part}.csv"))
        out[f"tuning{part}"] = df[["threshold", "nprobe", "Recall", "Computations"]].to_numpy(np.float64)

    out.update(centroids=kmeans.centroids.astype(np.float32), scaler_mean=mean32, scaler_scale=scale32,
               dist_q_scaled=dist_q.astype(np.float32), all_outputs=all_outputs.numpy().astype(np.float32),
               knn_self=knn_data)
    for i, w in enumerate(O.mlp_weights_from_state_dict(model.state_dict())):
        out[f"mlp_{i}"] = w
    return out, cfg, model, kmeans, data_2_bkt


def run_reference_cpp(cfg, model, out, x_d, x_q, gt, k, metric, workdir):
    """index.py:144-192 artifact layout -> oracle/_ref/search_ref -> parsed stdout."""
    import torch
    art = os.path.join(workdir, "art")
    os.makedirs(art, exist_ok=True)
    p = os.path.join(art, cfg.file_name)
    np.save(p + "_centroids.npy", out["centroids"])
    np.save(p + "_data_2_bkt.npy", out["d2b1"].astype(np.int32))
    np.save(p + "_x_d.npy", x_d.astype(np.float32))
    np.save(p + "_scaler_mean.npy", out["scaler_mean"].astype(np.float32))
    np.save(p + "_scaler_scale.npy", out["scaler_scale"].astype(np.float32))
    torch.jit.save(torch.jit.script(model.eval()), p + "_mlp_2_input.pt")
    ds = os.path.join(workdir, "data", "toy")
    os.makedirs(ds, exist_ok=True)
    write_xvecs(os.path.join(ds, "toy_query.fvecs"), x_q, np.float32)
    write_xvecs(os.path.join(ds, "toy_groundtruth.ivecs"), gt, np.int32)
    exe = os.path.join(HERE, "_ref", "search_ref")
    txt = subprocess.run([exe, "--dataset", "toy", "--data_path", os.path.join(workdir, "data"),
                          "--artifacts_dir", art, "--prefix", cfg.file_name, "--k", str(k),
                          "--metric", metric, "--num_threads", "1"],
                         check=True, capture_output=True, text=True).stdout
    rows = []
    for blk in txt.split("----------------------------------------"):
        m = {key: re.search(key + r"\s*:\s*([-+0-9.eE]+)", blk) for key in
             ("Threshold", "avg_recall", "avg_nprobe", "avg_cmp")}
        if all(m.values()):
            rows.append([float(m[key].group(1)) for key in ("Threshold", "avg_recall", "avg_nprobe", "avg_cmp")])
    return np.array(rows, np.float64)


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, metric, integer in (("toy_l2", "L2", True), ("toy_ip", "inner_product", False)):
        n, d, nq, B, k = 4000, 20, 64, 32, 10
        x_d, x_q = synth(n, d, nq, seed=43, integer=integer)
        _, gt = O.knn(x_d, x_q, 20, O.IP if metric == "inner_product" else O.L2, O.F64)
        gt = gt.astype(np.int32)
        with tempfile.TemporaryDirectory() as wd:
            cwd = os.getcwd()
            os.chdir(wd)
            try:
                out, cfg, model, kmeans, d2b = run_reference_python(x_d, x_q, gt, B, k, metric, wd)
                out["cpp_rows"] = run_reference_cpp(cfg, model, out, x_d, x_q, gt, k, metric, wd)
            finally:
                os.chdir(cwd)
        out.update(x_d=x_d, x_q=x_q, gt=gt, k=np.int64(k), n_bkt=np.int64(B),
                   metric=np.int64(1 if metric == "inner_product" else 0))
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes; cpp rows", out["cpp_rows"].shape,
              "tuning rows", out["tuning1"].shape)


if __name__ == "__main__":
    main()
