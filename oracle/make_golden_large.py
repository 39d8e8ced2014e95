"""TEST INFRASTRUCTURE -- tests/golden/toy_large.npz: the LIRA_largescale.py flow run with the REFERENCE'S OWN
functions, imported unmodified from /root/reference (fake-faiss / prettytable shims of oracle/ref_shim on sys.path).

The reference's `__main__` cannot run here as a script (get_idle_gpu() shells out to nvidia-smi and the model is
moved to "cuda:N", LIRA_largescale.py:197-198), so this file calls the same functions in the same order
(LIRA_largescale.py:184-354) on a toy dataset, on the CPU, and records every intermediate the mirrors in
lira-ann-search_b200/ are tested against:

  utils.compute_data_knn (python fallback + the .npy cache name, utils.py:223-319), the query-on-subset kNN
  (LIRA_largescale.py:217-234), utils.build_kmeans_index, utils.get_scaled_dist (called with cfg: the 4-argument call
  at LIRA_largescale.py:258 is a bug in the reference), model_probing.model_train / model_evaluate /
  model_infer, kmeans.index.search (LIRA_largescale.py:294), LIRA_largescale.get_cmp_recall / query_tuning /
  mul_partition_by_model, utils.get_scaled_dist_data (per-batch scaler, utils.py:182-215), utils.per_query.

Deviations from the script, all forced by the toy size and stated here: the training subset is 1/2 of the base
instead of 1/100 (60 points cannot train anything), the redundancy batch is 2500 points instead of 1 000 000 (so that
the per-batch scaler quirk is exercised by three batches), n_epoch = 3, device = "cpu".

Usage:  python oracle/make_golden_large.py
"""
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("LIRA_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "ref_shim"))
sys.path.insert(0, HERE)
sys.path.append(REF)

import oracle as O  # noqa: E402
from make_golden import synth  # noqa: E402

SUB_DIV, BATCH_RED, N_EPOCH = 2, 2500, 3


def csr(cluster_ids):
    off = np.zeros(len(cluster_ids) + 1, np.int64)
    np.cumsum([len(c) for c in cluster_ids], out=off[1:])
    ids = np.concatenate([np.asarray(c, np.int64) for c in cluster_ids] + [np.empty(0, np.int64)]).astype(np.int32)
    return off, ids


def main():
    import pandas as pd
    import torch
    from torch.utils.data import DataLoader, TensorDataset

    import LIRA_largescale as S  # reference, unmodified
    import model_probing as MP  # reference, unmodified
    import utils as U  # reference, unmodified

    n, d, nq, B, k = 6000, 16, 120, 16, 10
    x_d, x_q = synth(n, d, nq, seed=47, integer=True)
    _, gt = O.knn(x_d, x_q, 20, O.L2, O.F64)
    gt = gt.astype(np.int32)
    out = dict(x_d=x_d, x_q=x_q, gt=gt, k=np.int64(k), n_bkt=np.int64(B), sub_div=np.int64(SUB_DIV),
               batch_redundancy=np.int64(BATCH_RED), n_epoch=np.int64(N_EPOCH))
    with tempfile.TemporaryDirectory() as wd:
        cwd = os.getcwd()
        os.chdir(wd)
        try:
            torch.manual_seed(0)
            cfg = S.Config(dataset="toyl", k=k, n_bkt=B, batch_size=64, n_epoch=N_EPOCH)
            cfg.data_path = os.path.join(wd, "data")
            cfg.update()
            cfg.pth_log = os.path.join(wd, "logs") + "/"
            cfg.file_name = f"toyl-k={k}-ML_kmeans={B}_FLAT_ReType=model"
            os.makedirs(cfg.pth_log, exist_ok=True)
            fw = io.StringIO()
            S.fw = fw
            n_d, dim = x_d.shape
            device = "cpu"

            # LIRA_largescale.py:200-205 (1/100 in the script)
            nd_sub = int(n_d / SUB_DIV)
            np.random.seed(43)
            sub_idx = np.random.choice(range(len(x_d)), nd_sub, replace=False)
            xd_sub = x_d[sub_idx]
            out["sub_idx"] = np.asarray(sub_idx, np.int64)

            # :209 self-kNN of the subset (python fallback -> .npy cache)
            knn_data_sub = U.compute_data_knn(xd_sub, cfg, data_path=cfg.data_path)
            cache_dir = os.path.join(cfg.data_path, cfg.dataset, "knn_cache")
            assert os.listdir(cache_dir) == [f"toyl-data_self_knn{k}-n{nd_sub}.npy"]
            out["knn_data_sub"] = np.asarray(knn_data_sub, np.int32)
            # :217-234 query kNN on the subset
            index_flat = U.faiss.IndexFlatL2(dim)
            index_flat.add(xd_sub)
            _, knn_query_sub = index_flat.search(x_q, cfg.k)
            out["knn_query_sub"] = np.asarray(knn_query_sub, np.int64)

            # :237-241
            data_2_bkt_sub = np.full((nd_sub, cfg.n_mul), -1)
            kmeans, single_sub, cluster_cnts, cluster_ids = U.build_kmeans_index(xd_sub, B)
            data_2_bkt_sub[:, :1] = single_sub
            out["centroids"] = kmeans.centroids.astype(np.float32)
            # :246-252
            cnt_d_sub, _ = U.get_knn_distr_redundancy(knn_data_sub, data_2_bkt_sub, cfg)
            cnt_q_sub, ids_q_sub = U.get_knn_distr_redundancy(knn_query_sub, data_2_bkt_sub, cfg)
            labels_data = np.where(cnt_d_sub != 0, 1, cnt_d_sub)
            labels_query = np.where(cnt_q_sub != 0, 1, cnt_q_sub)
            # :258 (with cfg)
            dist_d_sub, dist_q = U.get_scaled_dist(xd_sub, x_q, kmeans, B, cfg)
            out["scaler_mean"] = np.load(os.path.join(cfg.pth_log, f"{cfg.file_name}_scaler_mean.npy"))
            out["scaler_scale"] = np.load(os.path.join(cfg.pth_log, f"{cfg.file_name}_scaler_scale.npy"))
            out["dist_q_scaled"] = dist_q.astype(np.float32)
            out["dist_d_sub_scaled"] = dist_d_sub.astype(np.float32)

            tr = TensorDataset(torch.tensor(dist_d_sub, dtype=torch.float32), torch.tensor(xd_sub, dtype=torch.float32),
                               torch.tensor(labels_data, dtype=torch.float32))
            te = TensorDataset(torch.tensor(dist_q, dtype=torch.float32), torch.tensor(x_q, dtype=torch.float32),
                               torch.tensor(labels_query, dtype=torch.float32))
            trl = DataLoader(tr, batch_size=cfg.batch_size, shuffle=False)
            tel = DataLoader(te, batch_size=cfg.batch_size, shuffle=False)
            model = MP.MLP_2_Input(input_dim1=B, input_dim2=dim, output_dim=B).to(device)
            crit = torch.nn.BCELoss()
            opt = torch.optim.Adam(model.parameters(), lr=1e-3)   # (the script's 1e-4 needs its 30 epochs)
            for w_i, w in enumerate(O.mlp_weights_from_state_dict(model.state_dict())):
                out[f"mlp_init_{w_i}"] = w
            results_df = pd.DataFrame(columns=['Epoch', 'Accuracy', 'Hit Rate', 'nprobe predict', 'nprobe target', 'KNN Recall', 'KNN Computations'])
            all_targets, all_predicts, loss_test, all_outputs = MP.model_evaluate(model, tel, crit, device)
            results_df = S.cal_metrics(all_predicts, all_targets, -1, ids_q_sub, cluster_ids, results_df, loss_test, knn=cfg.k)
            for epoch in range(cfg.n_epoch):
                MP.model_train(model, trl, device, opt, crit)
                all_targets, all_predicts, loss_test, all_outputs = MP.model_evaluate(model, tel, crit, device)
                results_df = S.cal_metrics(all_predicts, all_targets, epoch, ids_q_sub, cluster_ids, results_df, loss_test, knn=cfg.k)
            out["all_outputs"] = all_outputs.numpy().astype(np.float32)
            out["all_predicts"] = all_predicts.numpy()
            out["all_targets"] = all_targets.numpy().astype(np.float32)
            out["loss_test"] = np.float64(loss_test)
            out["metrics"] = results_df.to_numpy(np.float64)
            for w_i, w in enumerate(O.mlp_weights_from_state_dict(model.state_dict())):
                out[f"mlp_{w_i}"] = w

            # :292-299 partition the full data
            data_2_bkt = np.full((n_d, cfg.n_mul), -1)
            _, single = kmeans.index.search(x_d, 1)
            data_2_bkt[:, :1] = single
            cluster_cnts = np.bincount(single.flatten(), minlength=B)
            cluster_ids = [[] for _ in range(cfg.n_bkt)]
            for idx, cid in enumerate(single.flatten()):
                cluster_ids[cid].append(idx)
            out["assign_full"] = single.flatten().astype(np.int32)
            out["labels_query_sub"] = labels_query.astype(np.uint8)
            knn_query = gt[:, :cfg.k]
            cnt_q, ids_q = U.get_knn_distr_redundancy(knn_query, data_2_bkt, cfg)

            # :310-318 before redundancy
            idx0 = U.create_inner_indexes(x_d, cluster_ids, cfg)
            _, cmp0, found0 = S.get_cmp_recall(idx0, x_q, cluster_ids, cfg)
            S.cmp_distr_all = cmp0   # query_tuning reads a module global (LIRA_largescale.py:165)
            S.query_tuning(all_outputs, ids_q, found0, cfg)
            out["lists0_off"], out["lists0_ids"] = csr(cluster_ids)
            out["found0"], out["cmp0"] = found0.astype(np.int64), cmp0.astype(np.int64)
            out["knn_cnt0"] = cnt_q.astype(np.int64)

            # :320-329 full redundancy, in batches with their OWN scaler (utils.py:182-215)
            first = True
            score_all = np.zeros((n_d, B), np.float32)
            for start_idx in range(0, n_d, BATCH_RED):
                end_idx = min(start_idx + BATCH_RED, n_d)
                xd_batch = x_d[start_idx:end_idx]
                dist_b = U.get_scaled_dist_data(xd_batch, kmeans, B)
                if first:
                    out["dist_batch0_scaled"] = np.asarray(dist_b, np.float32)
                tl = DataLoader(TensorDataset(torch.tensor(dist_b, dtype=torch.float32), torch.tensor(xd_batch, dtype=torch.float32)),
                                batch_size=cfg.batch_size, shuffle=False)
                data_predicts, data_score = MP.model_infer(model, tl, device)
                score_all[start_idx:end_idx] = data_score.numpy()
                first = False
                S.mul_partition_by_model(data_score, data_predicts, np.arange(start_idx, end_idx), start_idx, data_2_bkt,
                                         cluster_cnts, cluster_ids)
            out["score_all"] = score_all
            out["d2b1"] = data_2_bkt.astype(np.int32)
            out["cnts1"] = np.asarray(cluster_cnts, np.int64)
            out["lists1_off"], out["lists1_ids"] = csr(cluster_ids)
            cnt_q1, ids_q1 = U.get_knn_distr_redundancy(knn_query, data_2_bkt, cfg)
            idx1 = U.create_inner_indexes(x_d, cluster_ids, cfg)
            _, cmp1, found1 = S.get_cmp_recall(idx1, x_q, cluster_ids, cfg)
            S.cmp_distr_all = cmp1
            S.query_tuning(all_outputs, ids_q1, found1, cfg, part=1)
            out["found1"], out["cmp1"] = found1.astype(np.int64), cmp1.astype(np.int64)
            out["knn_cnt1"] = cnt_q1.astype(np.int64)
            for part in (0, 1):
                df = pd.read_csv(os.path.join(cfg.pth_log, cfg.file_name + "_tuning_threshold", f"{cfg.duplicate_type}_{part}.csv"))
                out[f"tuning{part}"] = df[["threshold", "nprobe", "Recall", "Computations"]].to_numpy(np.float64)

            # utils.per_query (utils.py:502-519)
            U.per_query(all_outputs, cnt_q1, np.asarray(cluster_cnts), B, cfg)
            dfp = pd.read_csv(cfg.pth_log + f"{cfg.dataset}-k={cfg.k}-ML_kmeans={B}_perquery.csv")
            out["per_query"] = dfp[["q_id", "nprobe", "cmp"]].to_numpy(np.int64)
        finally:
            os.chdir(cwd)
    path = os.path.join(ROOT, "tests", "golden", "toy_large.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    print("tuning1 head:\n", out["tuning1"][:5], "\nper_query head:\n", out["per_query"][:5], "\nmetrics:\n", out["metrics"])


if __name__ == "__main__":
    main()
