"""The two command-line programs with the reference's argv and file formats:
bin/compute_knn (compute_knn.cpp) and bin/search (search.cpp). GPU tests: they call the C ABI."""
import os
import re
import subprocess

import numpy as np
import pytest

import oracle as O
from helpers import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")


def _need(name):
    path = os.path.join(BIN, name)
    if not os.path.exists(path):
        pytest.fail(f"{path} is missing: run __graft_entry__.build()")
    return path


def test_compute_knn_writes_the_reference_cache_files(tmp_path):
    import lira_ann_search_b200 as L
    exe = _need("compute_knn")
    x_d, _ = synth(3000, 24, 1, seed=4, integer=True)
    ds = tmp_path / "toy"
    ds.mkdir()
    L.write_xvecs(str(ds / "toy_base.fvecs"), x_d)
    r = subprocess.run([exe, "toy", str(tmp_path), "10", "0", "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = ds / "knn_cache" / "toy-data_self_knn10-n3000.bin"   # compute_knn.cpp:262-271 naming
    knn = np.fromfile(out, dtype=np.int32).reshape(3000, 10)
    _, I_ref = O.knn(x_d, x_d, 11, O.L2, O.F64)
    assert np.array_equal(knn, I_ref[:, 1:])                       # column 0 (self) dropped (:254-259)
    meta = dict(l.split(": ", 1) for l in open(str(out) + ".meta").read().splitlines())
    assert meta["dataset"] == "toy" and meta["n"] == "3000" and meta["dim"] == "24" and meta["k"] == "10"
    assert meta["method"] == "flat_exact" and {"read_time", "build_time", "search_time", "total_time"} <= set(meta)
    # missing dataset: message + exit code 1 like compute_knn.cpp:118-122
    r = subprocess.run([exe, "nope", str(tmp_path), "10"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Cannot find base file" in r.stderr


def test_compute_knn_ivf_branch(tmp_path):
    """compute_knn <ds> <path> <k> <nprobe != 0>: the reference's IVF-approximate branch (compute_knn.cpp:150-203): its nlist rule,
    the `_ivf_nprobe{p}` suffix utils.compute_data_knn looks for first, the extra .meta keys. Probing every list must give the
    exact result; a small nprobe a good approximation; no nprobe argument the automatic one (:190-199)."""
    import lira_ann_search_b200 as L
    exe = _need("compute_knn")
    x_d, _ = synth(3000, 24, 1, seed=4, integer=True)
    ds = tmp_path / "toy"
    ds.mkdir()
    L.write_xvecs(str(ds / "toy_base.fvecs"), x_d)
    _, I_ref = O.knn(x_d, x_d, 11, O.L2, O.F64)
    nlist = 54                                                     # min(int(sqrt(3000)), 256)
    for nprobe, arg in ((nlist, [str(nlist)]), (8, ["8", "4"]), (16, [])):   # 16 = min(max(54 // 4, 16), 64): automatic
        r = subprocess.run([exe, "toy", str(tmp_path), "10"] + arg, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        out = ds / "knn_cache" / f"toy-data_self_knn10-n3000_ivf_nprobe{nprobe}.bin"
        knn = np.fromfile(out, dtype=np.int32).reshape(3000, 10)
        meta = dict(l.split(": ", 1) for l in open(str(out) + ".meta").read().splitlines())
        assert meta["method"] == "ivf_approximate" and meta["n_clusters"] == str(nlist) and meta["nprobe"] == str(nprobe)
        assert meta["probe_ratio"].endswith("%")
        if nprobe == nlist:
            assert np.array_equal(knn, I_ref[:, 1:])
        else:
            rec = np.mean([len(set(knn[i]) & set(I_ref[i, 1:])) / 10 for i in range(3000)])
            assert rec > 0.9, rec
    # utils.compute_data_knn takes the newest *_ivf_nprobe*.bin before anything else (utils.py:245-266)
    cfg = type("C", (), {"dataset": "toy", "k": 10, "dis_metric": "L2"})
    got = L.compute_data_knn(x_d, cfg, data_path=str(tmp_path))
    assert got.shape == (3000, 10) and got.dtype == np.int32


def test_compute_knn_query_ground_truth_mode(tmp_path):
    """compute_knn <ds> <path> <k> --queries: exact ground truth of the query set as {ds}_groundtruth.ivecs (the layout
    utils.read_xvecs / search.cpp:read_ivecs read), k = 100 like config 2."""
    import lira_ann_search_b200 as L
    exe = _need("compute_knn")
    x_d, x_q = synth(20000, 32, 300, seed=9, integer=True)
    ds = tmp_path / "toy"
    ds.mkdir()
    L.write_xvecs(str(ds / "toy_base.fvecs"), x_d)
    L.write_xvecs(str(ds / "toy_query.fvecs"), x_q)
    r = subprocess.run([exe, "toy", str(tmp_path), "100", "--queries"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    gt = L.read_xvecs(str(ds / "toy_groundtruth.ivecs"), dtype="int32")
    assert gt.shape == (300, 100)
    D_ref, I_ref = O.knn(x_d, x_q, 100, O.L2, O.F64)
    assert np.array_equal(gt, I_ref)   # integer data: distances exact, ties by id


@pytest.mark.parametrize("case", ["toy_l2", "toy_ip"])
def test_search_matches_reference_search_cpp_stdout(golden, tmp_path, case):
    """bin/search on artifacts in index.py's layout against the stdout of the UNMODIFIED reference search.cpp
    (tests/golden: cpp_rows = Threshold, avg_recall, avg_nprobe, avg_cmp per threshold)."""
    import torch
    import lira_ann_search_b200 as L
    exe = _need("search")
    z = golden(case)
    k, B, metric = int(z["k"]), int(z["n_bkt"]), int(z["metric"])
    d = z["x_d"].shape[1]
    pfx = str(tmp_path / "art" / "toy")
    os.makedirs(os.path.dirname(pfx))
    np.save(pfx + "_centroids.npy", z["centroids"].astype(np.float32))
    np.save(pfx + "_data_2_bkt.npy", z["d2b1"].astype(np.int32))
    np.save(pfx + "_x_d.npy", z["x_d"].astype(np.float32))
    np.save(pfx + "_scaler_mean.npy", z["scaler_mean"].astype(np.float32))
    np.save(pfx + "_scaler_scale.npy", z["scaler_scale"].astype(np.float32))
    model = L.MLP_2_Input(B, d, B)
    keys = ("distance_net.0", "distance_net.2", "vector_net.0", "vector_net.2", "fc.0", "fc.2")
    sd = {}
    for i, kk in enumerate(keys):
        sd[kk + ".weight"] = torch.as_tensor(z[f"mlp_{2 * i}"])
        sd[kk + ".bias"] = torch.as_tensor(z[f"mlp_{2 * i + 1}"])
    model.load_state_dict(sd)
    torch.jit.save(torch.jit.script(model.eval()), pfx + "_mlp_2_input.pt")
    ds = tmp_path / "data" / "toy"
    ds.mkdir(parents=True)
    L.write_xvecs(str(ds / "toy_query.fvecs"), z["x_q"])
    L.write_xvecs(str(ds / "toy_groundtruth.ivecs"), z["gt"].astype(np.int32))
    r = subprocess.run([exe, "--dataset", "toy", "--data_path", str(tmp_path / "data"), "--artifacts_dir", os.path.dirname(pfx),
                        "--prefix", "toy", "--k", str(k), "--metric", "inner_product" if metric == O.IP else "L2"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    rows = list(zip(*[[float(x) for x in re.findall(key + r"\s*:\s*([-+0-9.eE]+)", r.stdout)]
                      for key in ("Threshold   ", "avg_recall", "avg_nprobe", "avg_cmp")]))
    ref = z["cpp_rows"]
    assert len(rows) == len(ref) == 40 and "QPS" in r.stdout and r.stdout.rstrip().endswith("Done.")
    Q = len(z["x_q"])
    scores = L.LiraModel.from_arrays(z["centroids"], z["scaler_mean"], z["scaler_scale"],
                                     [z[f"mlp_{i}"] for i in range(12)]).scores(z["x_q"])
    for row, rr in zip(rows, ref):
        edge = int((np.abs(scores.astype(np.float64) - rr[0]) < 2e-6).sum())   # scores on the threshold may fall either way
        assert abs(row[0] - rr[0]) < 1e-5
        assert abs(row[2] - rr[2]) <= edge / Q + 1e-4 * max(1.0, rr[2])          # stdout carries ~6 significant digits
        if edge == 0:
            assert abs(row[3] - rr[3]) <= 1e-4 * max(1.0, rr[3])
            assert abs(row[1] - rr[1]) <= (1e-5 if case == "toy_l2" else 2.0 / (k * Q)) + 1e-6
    # unreadable artifacts: "[Error] ..." + exit code 1 (search.cpp:552-555)
    r = subprocess.run([exe, "--dataset", "toy", "--prefix", "missing", "--artifacts_dir", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("[Error]")
