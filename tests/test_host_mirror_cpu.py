"""CPU: the host-side mirrors of the reference's Python call shapes (no GPU, no compute call into liblira_b200), against
fixtures produced by the reference's own unmodified functions (oracle/make_golden.py, oracle/make_golden_large.py)."""
import os
import types

import numpy as np
import pytest
import torch

import lira_ann_search_b200 as L


def _cfg(**kw):
    return types.SimpleNamespace(**kw)


def _lists(z, which):
    off, ids = z[f"lists{which}_off"], z[f"lists{which}_ids"]
    return [ids[off[b]:off[b + 1]].tolist() for b in range(len(off) - 1)]


def test_mul_partition_by_model_large_matches_reference(golden):
    """LIRA_largescale.py:51-72 over three redundancy batches (each scored with its own per-batch scaler)."""
    z = golden("toy_large")
    n, B, step = len(z["x_d"]), int(z["n_bkt"]), int(z["batch_redundancy"])
    d2b = np.full((n, 2), -1, np.int64)
    d2b[:, 0] = z["assign_full"]
    cluster_ids = _lists(z, 0)
    cnts = np.bincount(z["assign_full"], minlength=B).astype(np.int64)
    for a in range(0, n, step):
        e = min(a + step, n)
        score = torch.as_tensor(z["score_all"][a:e])
        L.mul_partition_by_model_large(score, score > 0.5, np.arange(a, e), a, d2b, cnts, cluster_ids)
    assert np.array_equal(d2b, z["d2b1"])
    assert np.array_equal(cnts, z["cnts1"])
    assert cluster_ids == _lists(z, 1)


@pytest.mark.parametrize("part", [0, 1])
def test_query_tuning_large_matches_reference_csv(golden, part, tmp_path):
    """LIRA_largescale.py:151-179: thresholds arange(0.1, 1, 0.02), columns threshold,nprobe,Recall,Computations."""
    import pandas as pd
    z = golden("toy_large")
    k, B = int(z["k"]), int(z["n_bkt"])
    cfg = _cfg(n_bkt=B, k=k, pth_log=str(tmp_path) + "/", file_name="toyl", duplicate_type="model")
    d2b = np.full((len(z["x_d"]), 2), -1, np.int64)
    d2b[:, 0] = z["assign_full"]
    if part == 1:
        d2b = z["d2b1"].astype(np.int64)
    cnt, ids = L.get_knn_distr_redundancy(z["gt"][:, :k], d2b, cfg)
    assert np.array_equal(cnt, z[f"knn_cnt{part}"])
    df = L.query_tuning_large(z["all_outputs"], ids, z[f"found{part}"], z[f"cmp{part}"], cfg, part=part)
    np.testing.assert_allclose(df.to_numpy(np.float64), z[f"tuning{part}"], rtol=1e-9, atol=1e-12)
    back = pd.read_csv(os.path.join(str(tmp_path), "toyl_tuning_threshold", f"model_{part}.csv"))
    assert list(back.columns) == ["threshold", "nprobe", "Recall", "Computations"] and len(back) == 45


def test_per_query_matches_reference_csv(golden, tmp_path):
    """utils.py:502-519: smallest top-nprobe (1..19) per query reaching recall 0.98; 0 when none does."""
    import pandas as pd
    z = golden("toy_large")
    k, B = int(z["k"]), int(z["n_bkt"])
    cfg = _cfg(k=k, dataset="toyl", pth_log=str(tmp_path) + "/")
    df = L.per_query(torch.as_tensor(z["all_outputs"]), z["knn_cnt1"], z["cnts1"], B, cfg)
    assert np.array_equal(df.to_numpy(np.int64), z["per_query"])
    back = pd.read_csv(os.path.join(str(tmp_path), f"toyl-k={k}-ML_kmeans={B}_perquery.csv"))
    assert np.array_equal(back.to_numpy(np.int64), z["per_query"])


def test_model_evaluate_and_model_infer_call_shapes(golden):
    """model_probing.py:86-156: 4-tuple (targets, predicts, mean loss, outputs) / 2-tuple (predicts, outputs), CPU tensors,
    predicts = outputs > 0.5, loss = mean of the per-batch BCELoss means; values vs the reference's own run."""
    from torch.utils.data import DataLoader, TensorDataset
    z = golden("toy_large")
    B, d = int(z["n_bkt"]), z["x_d"].shape[1]
    model = L.MLP_2_Input(B, d, B)
    keys = ("distance_net.0", "distance_net.2", "vector_net.0", "vector_net.2", "fc.0", "fc.2")
    sd = {}
    for i, kk in enumerate(keys):
        sd[kk + ".weight"] = torch.as_tensor(z[f"mlp_{2 * i}"])
        sd[kk + ".bias"] = torch.as_tensor(z[f"mlp_{2 * i + 1}"])
    model.load_state_dict(sd)   # same parameter names as the reference's module
    xdist, xvec = torch.as_tensor(z["dist_q_scaled"]), torch.as_tensor(z["x_q"])
    labels = torch.as_tensor(z["labels_query_sub"], dtype=torch.float32)
    loader = DataLoader(TensorDataset(xdist, xvec, labels), batch_size=64, shuffle=False)
    out = L.model_evaluate(model, loader, torch.nn.BCELoss(), "cpu")
    assert isinstance(out, tuple) and len(out) == 4
    targets, predicts, loss, outputs = out
    assert all(t.device.type == "cpu" for t in (targets, predicts, outputs)) and isinstance(loss, float)
    np.testing.assert_allclose(outputs.numpy(), z["all_outputs"], rtol=1e-5, atol=1e-6)
    assert predicts.dtype == torch.bool and np.array_equal(predicts.numpy(), z["all_predicts"])
    assert np.array_equal(targets.numpy(), z["all_targets"])
    assert loss == pytest.approx(float(z["loss_test"]), rel=1e-5)
    out2 = L.model_infer(model, DataLoader(TensorDataset(xdist, xvec), batch_size=50, shuffle=False), "cpu")
    assert isinstance(out2, tuple) and len(out2) == 2
    np.testing.assert_allclose(out2[1].numpy(), z["all_outputs"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(out2[0].numpy(), z["all_outputs"] > 0.5) or np.array_equal(out2[0].numpy(), out2[1].numpy() > 0.5)


def test_compute_data_knn_cache_lookup_order(tmp_path):
    """utils.py:245-272: newest `*_ivf_nprobe*.bin` first, then the exact `.bin`, then the `.npy`; int32 raw [n, k]."""
    n, k = 50, 4
    cfg = _cfg(dataset="toy", k=k, dis_metric="L2")
    cache = tmp_path / "toy" / "knn_cache"
    cache.mkdir(parents=True)
    x = np.zeros((n, 3), np.float32)
    a = np.full((n, k), 1, np.int32)
    np.save(cache / f"toy-data_self_knn{k}-n{n}.npy", a)
    assert np.array_equal(L.compute_data_knn(x, cfg, str(tmp_path)), a)
    b = np.full((n, k), 2, np.int32)
    b.tofile(cache / f"toy-data_self_knn{k}-n{n}.bin")
    got = L.compute_data_knn(x, cfg, str(tmp_path))
    assert got.dtype == np.int32 and np.array_equal(got, b)
    c = np.full((n, k), 3, np.int32)
    c.tofile(cache / f"toy-data_self_knn{k}-n{n}_ivf_nprobe8.bin")
    assert np.array_equal(L.compute_data_knn(x, cfg, str(tmp_path)), c)
    e = np.full((n, k), 4, np.int32)
    e.tofile(cache / f"toy-data_self_knn{k}-n{n}_ivf_nprobe64.bin")
    os.utime(cache / f"toy-data_self_knn{k}-n{n}_ivf_nprobe64.bin")
    assert np.array_equal(L.compute_data_knn(x, cfg, str(tmp_path)), e)   # the most recently created IVF cache wins
    # another n or k is another cache entry: nothing matches, and without a GPU the fallback must fail loudly
    if L._cabi.lib().lira_device_count() == 0:
        with pytest.raises(L.LiraError):
            L.compute_data_knn(x[:10], cfg, str(tmp_path))


def test_xvecs_io_and_load_data(tmp_path):
    """Appendix B formats: fvecs / ivecs (int32 d + d 4-byte values), bvecs (int32 d + d bytes); load_data's file names,
    the `_learn` fallback, the optional ground truth and the FileNotFoundError of a missing file (utils.py:23-88)."""
    rng = np.random.RandomState(0)
    ds = tmp_path / "toy"
    ds.mkdir()
    x_d, x_q = rng.randn(30, 7).astype(np.float32), rng.randn(5, 7).astype(np.float32)
    gt = rng.randint(0, 30, (5, 10)).astype(np.int32)
    L.write_xvecs(str(ds / "toy_learn.fvecs"), x_d)
    L.write_xvecs(str(ds / "toy_query.fvecs"), x_q)
    a, b, g = L.load_data("toy", str(tmp_path))
    assert np.array_equal(a, x_d) and np.array_equal(b, x_q) and g is None and a.flags.c_contiguous
    L.write_xvecs(str(ds / "toy_base.fvecs"), x_d[:20])
    L.write_xvecs(str(ds / "toy_groundtruth.ivecs"), gt)
    a, b, g = L.load_data("toy", str(tmp_path))
    assert np.array_equal(a, x_d[:20]) and np.array_equal(g, gt) and g.dtype == np.int32
    raw = np.fromfile(str(ds / "toy_query.fvecs"), np.int32).reshape(5, 8)
    assert np.all(raw[:, 0] == 7)
    bv = rng.randint(0, 256, (9, 12)).astype(np.uint8)
    rec = np.zeros((9, 16), np.uint8)
    rec[:, :4] = np.frombuffer(np.int32(12).tobytes(), np.uint8)
    rec[:, 4:] = bv
    rec.tofile(str(ds / "toy.bvecs"))
    assert np.array_equal(L.read_xvecs(str(ds / "toy.bvecs"), "uint8"), bv)
    with pytest.raises(FileNotFoundError):
        L.read_xvecs(str(ds / "missing.fvecs"))
    with pytest.raises(ValueError):
        L.write_xvecs(str(ds / "x.bvecs"), bv)


def test_threshold_is_compared_in_fp32():
    """LIRA_smallscale.py:206 compares a torch fp32 tensor with np.float64(thr): evaluated in fp32, so a score equal to
    float32(thr) is NOT selected (checked against torch itself)."""
    import oracle as O
    s = np.array([[np.float32(0.1), 0.5, 0.0999]], np.float32)
    assert not bool((torch.as_tensor(s) > np.float64(0.1))[0, 0])
    poff, pids = O.select(s, O.SELECT_GT, 0.1)
    assert pids.tolist() == [1]
    poff, pids = O.select(s, O.SELECT_GE_ARGMAX, 0.1)
    assert pids.tolist() == [0, 1]


def test_cal_metrics_and_configs_match_the_reference(golden):
    """drivers.cal_metrics against the metrics table the reference's own cal_metrics (LIRA_smallscale.py:99-143) produced in
    the large-scale golden run, and the Config name / argv conventions of both drivers."""
    import torch
    import lira_ann_search_b200 as L
    z = golden("toy_large")
    B, k = int(z["n_bkt"]), int(z["k"])
    xs = z["x_d"][z["sub_idx"]]
    d2 = ((xs[:, None, :].astype(np.float64) - z["centroids"][None].astype(np.float64)) ** 2).sum(-1)
    d2b_sub = np.full((len(xs), 2), -1)
    d2b_sub[:, 0] = d2.argmin(1)

    class C:
        n_bkt = B
    _, ids_q = L.get_knn_distr_redundancy(z["knn_query_sub"], d2b_sub, C)
    df = L.cal_metrics(torch.as_tensor(z["all_predicts"]), torch.as_tensor(z["all_targets"]), 7, ids_q, None, None,
                       float(z["loss_test"]), knn=k)
    assert list(df.columns) == ["Epoch", "Accuracy", "Hit Rate", "nprobe predict", "nprobe target", "KNN Recall", "KNN Computations", "Loss"]
    ref = z["metrics"][-1]
    assert np.allclose(df.to_numpy(np.float64)[0, 1:], ref[1:], atol=2e-4)
    df2 = L.cal_metrics(torch.as_tensor(z["all_predicts"]), torch.as_tensor(z["all_targets"]), 8, ids_q, None, df, 0.5, knn=k)
    assert len(df2) == 2 and list(df2.columns) == list(df.columns)
    # Config conventions (LIRA_smallscale.py:44-75)
    c = L.parse_config(L.Config, ["--dataset", "sift", "--n_bkt", "1024", "--k", "10", "--dis_metric", "dot"])
    assert c.dis_metric == "inner_product" and c.pth_log == "./logs/sift/ML_kmeans_RE_FLAT/"
    assert c.file_name == "sift-k=10-ML_kmeans=1024_FLAT_Metric=inner_product_ReType=model_ReRatio=0.03"
    assert c.df_name == c.file_name + ".csv" and c.log_name == c.file_name + ".txt"
    with pytest.raises(ValueError):
        L.parse_config(L.Config, ["--n_bkt", "64", "--k", "10"])
    cl = L.parse_config(L.LargeConfig, [])
    assert cl.dis_metric == "L2" and cl.k == 100 and cl.n_bkt == 1024 and cl.batch_size == 512 and cl.n_epoch == 30
    assert cl.file_name == "deep50M-k=100-ML_kmeans=1024_FLAT_ReType=model"
