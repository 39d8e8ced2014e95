"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the
same seeded inputs and against the golden fixtures produced by the reference's own code.
Bars: ids bit-exact except for distance ties; fp32 distances within 1e-5 relative; recall identical."""
import numpy as np
import pytest

import oracle as O
from helpers import RTOL, assert_topk_equiv, random_lists, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import lira_ann_search_b200 as L
    L._cabi.require_gpu()
    return L


def lists_csr(x_d, cluster_ids):
    return O.build_lists_from_cluster_ids(x_d, cluster_ids)


# ---------------------------------------------------------------------------------------------
# a1/a2 features, a3 model
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["toy_l2", "toy_ip"])
def test_features_match_reference(L, golden, case):
    z = golden(case)
    raw = L.centroid_features(z["x_q"], z["centroids"])
    np.testing.assert_allclose(raw, O.features_cpp(z["x_q"], z["centroids"]), rtol=2e-6, atol=1e-6)
    got = L.centroid_features(z["x_q"], z["centroids"], z["scaler_mean"], z["scaler_scale"])
    np.testing.assert_allclose(got, z["dist_q_scaled"], rtol=1e-4, atol=1e-4)  # vs the reference's Python
    np.testing.assert_allclose(got, O.features_cpp(z["x_q"], z["centroids"], z["scaler_mean"], z["scaler_scale"]),
                               rtol=1e-4, atol=2e-5)  # vs the C++ twin's fp32 arithmetic


def test_features_odd_shapes_and_zero_scale(L):
    rng = np.random.RandomState(1)
    q, c = rng.randn(131, 10).astype(np.float32), rng.randn(37, 10).astype(np.float32)  # d % 4 != 0, B % 4 != 0
    mean, scale = rng.rand(37).astype(np.float32), rng.rand(37).astype(np.float32) + 0.5
    scale[5] = 0.0  # search.cpp:246: scale == 0 acts as 1
    np.testing.assert_allclose(L.centroid_features(q, c, mean, scale), O.features_cpp(q, c, mean, scale),
                               rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("case", ["toy_l2", "toy_ip"])
def test_model_scores_match_reference(L, golden, case):
    z = golden(case)
    w = [z[f"mlp_{i}"] for i in range(12)]
    model = L.LiraModel.from_arrays(z["centroids"], z["scaler_mean"], z["scaler_scale"], w)
    scores, feats = model.scores(z["x_q"], return_features=True)
    np.testing.assert_allclose(feats, z["dist_q_scaled"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(scores, z["all_outputs"], rtol=1e-4, atol=2e-6)  # reference torch forward
    _, probs, _ = O.mlp_forward(O.features_cpp(z["x_q"], z["centroids"], z["scaler_mean"], z["scaler_scale"]),
                                z["x_q"], w)
    np.testing.assert_allclose(scores, probs, rtol=1e-4, atol=2e-6)  # fp64 arbiter


# ---------------------------------------------------------------------------------------------
# a6/a7 per-list search and the all-pairs sweep
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case,part", [("toy_l2", 0), ("toy_l2", 1), ("toy_ip", 1)])
def test_get_cmp_recall_matches_reference(L, golden, case, part):
    z = golden(case)
    k, metric = int(z["k"]), int(z["metric"])
    off, ids = z[f"lists{part}_off"], z[f"lists{part}_ids"]
    index = L.LiraIndex.from_csr(z["x_d"], off, ids, metric)
    found, cmp_ = index.scan_all_pairs(z["x_q"], k)
    assert np.array_equal(cmp_, z[f"cmp{part}"])
    ref = z[f"found{part}"]
    if case == "toy_l2":  # integer data: fp32 arithmetic is exact, so the result must be bit-identical
        assert np.array_equal(found, ref)
    else:  # float data: identical except where two candidates tie within the tolerance
        bad = np.nonzero((found != ref).any(-1))
        for qi, b in zip(*bad):
            q = z["x_q"][qi].astype(np.float64)
            a = (z["x_d"][found[qi, b]].astype(np.float64) * q).sum(1)
            r = (z["x_d"][ref[qi, b]].astype(np.float64) * q).sum(1)
            np.testing.assert_allclose(a, r, rtol=RTOL, atol=RTOL)


@pytest.mark.parametrize("metric", [O.L2, O.IP])
@pytest.mark.parametrize("k", [1, 10, 33, 100])
def test_list_search_is_indexflat(L, metric, k):
    rng = np.random.RandomState(k)
    x_d, x_q = synth(1500, 20, 70, seed=7 + k, integer=(metric == O.L2))
    cl = random_lists(len(x_d), 6, rng, empty=(2,))
    cl[4] = cl[4][:5]  # shorter than k for k >= 10
    off, ids, vecs = lists_csr(x_d, cl)
    index = L.LiraIndex.from_cluster_ids(x_d, cl, metric)
    assert [index.ntotal(b) for b in range(6)] == [len(c) for c in cl]
    for b in range(6):
        D, I = index.list_search(b, x_q, k)
        D_ref, I_ref = O.list_search(vecs[off[b]:off[b + 1]], x_q, k, metric, O.F64)
        assert_topk_equiv(D, I, D_ref, I_ref, x_q, vecs[off[b]:off[b + 1]], metric)
        if metric == O.L2:  # integer data: exact, including tie order (lower position first)
            assert np.array_equal(I, I_ref)


def test_list_search_views_have_the_faiss_surface(L):
    x_d, x_q = synth(600, 16, 5, seed=3)
    cl = random_lists(600, 4, np.random.RandomState(0))

    class Cfg:
        n_bkt, k, dis_metric = 4, 7, "L2"
    inner = L.create_inner_indexes(x_d, cl, Cfg)
    assert len(inner) == 4 and inner[1].ntotal == len(cl[1])
    D, I = inner[1].search(x_q[0].reshape(1, -1), 7)
    D_ref, I_ref = O.list_search(x_d[cl[1]], x_q[:1], 7, O.L2, O.F64)
    assert np.array_equal(I, I_ref) and np.allclose(D, D_ref)
    _, cmp_, found = L.get_cmp_recall(inner, x_q, cl, Cfg)
    off, ids, vecs = lists_csr(x_d, cl)
    f_ref, c_ref = O.scan_all_pairs(off, ids, vecs, x_q, 7, O.L2, O.F64)
    assert np.array_equal(found, f_ref) and np.array_equal(cmp_, c_ref)


# ---------------------------------------------------------------------------------------------
# a10 online search with explicit probe sets
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric,integer", [(O.L2, True), (O.L2, False), (O.IP, False)])
@pytest.mark.parametrize("dedup", [1, 0])
def test_search_matches_oracle(L, metric, integer, dedup):
    rng = np.random.RandomState(11)
    x_d, x_q = synth(5000, 24, 333, seed=5, integer=integer)
    B, k = 40, 10
    cl = random_lists(len(x_d), B, rng, redundancy=0.5, empty=(7,))
    off, ids, vecs = lists_csr(x_d, cl)
    nprobe = rng.randint(0, 9, len(x_q))  # ragged, some queries probe nothing
    poff = np.zeros(len(x_q) + 1, np.int64)
    np.cumsum(nprobe, out=poff[1:])
    pids = np.concatenate([rng.choice(B, n, replace=False) for n in nprobe] + [np.empty(0, int)]).astype(np.int32)
    index = L.LiraIndex.from_csr(x_d, off, ids, metric)
    D, I, cmp_ = index.search(x_q, poff, pids, k, dedup=bool(dedup))
    I_ref, D_ref, cmp_ref = O.search(off, ids, vecs, x_q, poff, pids, k, metric, O.F64, dedup)
    assert np.array_equal(cmp_, cmp_ref)
    if dedup:
        assert_topk_equiv(D, I, D_ref, I_ref, x_q, x_d, metric)
        if integer:
            assert np.array_equal(I, I_ref)
    else:  # select-then-collapse: compare as sets per query (slot stealing depends on exact tie order)
        if integer:
            assert np.array_equal(I, I_ref)
        else:
            assert np.mean([set(a[a >= 0]) == set(b[b >= 0]) for a, b in zip(I, I_ref)]) > 0.98


def test_search_large_group_tiles_and_k100(L):
    """Every query probes the same few lists -> groups larger than one 64-row tile; k = 100 path."""
    x_d, x_q = synth(4000, 32, 203, seed=9, integer=False)
    B, k = 8, 100
    cl = random_lists(len(x_d), B, np.random.RandomState(2), redundancy=1.0)
    off, ids, vecs = lists_csr(x_d, cl)
    poff = np.arange(len(x_q) + 1, dtype=np.int64) * 3
    pids = np.tile(np.array([1, 4, 6], np.int32), len(x_q))
    index = L.LiraIndex.from_csr(x_d, off, ids, O.L2)
    D, I, _ = index.search(x_q, poff, pids, k)
    I_ref, D_ref, _ = O.search(off, ids, vecs, x_q, poff, pids, k, O.L2, O.F64, 1)
    assert_topk_equiv(D, I, D_ref, I_ref, x_q, x_d, O.L2)


def test_duplicate_vectors_and_dimension_padding(L):
    """Exact duplicates (ties broken by position / id) and d not a multiple of 4."""
    rng = np.random.RandomState(4)
    base = rng.randint(0, 4, (300, 6)).astype(np.float32)  # many exact ties
    base = np.concatenate([base, base[:50]])  # duplicated rows
    x_q = rng.randint(0, 4, (40, 6)).astype(np.float32)
    cl = random_lists(len(base), 5, rng, redundancy=0.3)
    off, ids, vecs = lists_csr(base, cl)
    index = L.LiraIndex.from_csr(base, off, ids, O.L2)
    found, _ = index.scan_all_pairs(x_q, 10)
    f_ref, _ = O.scan_all_pairs(off, ids, vecs, x_q, 10, O.L2, O.F64)
    assert np.array_equal(found, f_ref)
    poff = np.arange(len(x_q) + 1, dtype=np.int64) * 5
    pids = np.tile(np.arange(5, dtype=np.int32), len(x_q))
    D, I, _ = index.search(x_q, poff, pids, 10)
    I_ref, D_ref, _ = O.search(off, ids, vecs, x_q, poff, pids, 10, O.L2, O.F64, 1)
    assert np.array_equal(I, I_ref) and np.array_equal(D, D_ref)


def test_empty_inputs(L):
    x_d, x_q = synth(200, 8, 4, seed=1)
    cl = [[], list(range(200)), []]
    index = L.LiraIndex.from_cluster_ids(x_d, cl, O.L2)
    D, I = index.list_search(0, x_q, 3)
    assert (I == -1).all() and np.isinf(D).all()
    D, I, cmp_ = index.search(x_q, np.zeros(5, np.int64), np.empty(0, np.int32), 3)
    assert (I == -1).all() and (cmp_ == 0).all()
    found, cmp_ = index.scan_all_pairs(x_q, 3)
    assert (found[:, 0] == -1).all() and (found[:, 2] == -1).all() and (cmp_[:, 1] == 200).all()
    D, I = index.list_search(1, x_q[:0], 3)
    assert D.shape == (0, 3)


# ---------------------------------------------------------------------------------------------
# the whole query phase
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["toy_l2", "toy_ip"])
def test_query_phase_matches_reference_search_cpp(L, golden, case):
    """features -> MLP -> '>=' select + argmax fallback -> scan -> select-then-collapse -> recall, against the
    stdout of the unmodified reference search.cpp on the same artifacts (tests/golden/*.npz: cpp_rows)."""
    z = golden(case)
    k, B, metric = int(z["k"]), int(z["n_bkt"]), int(z["metric"])
    w = [z[f"mlp_{i}"] for i in range(12)]
    index = L.LiraIndex.from_data_2_bkt(z["x_d"], z["d2b1"], B, metric)
    model = L.LiraModel.from_arrays(z["centroids"], z["scaler_mean"], z["scaler_scale"], w)
    scores = model.scores(z["x_q"])
    rows = L.search_sweep(index, model, z["x_q"], z["gt"], k, mode=L.SELECT_GE_ARGMAX, dedup=False, out=None)
    ref = z["cpp_rows"]
    assert len(rows) == len(ref) == 40
    Q = len(z["x_q"])
    for row, r in zip(rows, ref):
        # a score within 1e-6 of the threshold may fall on either side (fp32 MLP rounding)
        edge = int((np.abs(scores.astype(np.float64) - row["Threshold"]) < 1e-6).sum())
        assert abs(row["avg_nprobe"] - r[2]) <= edge / Q + 1e-5 * max(1.0, r[2])
        if edge == 0:
            assert abs(row["avg_cmp"] - r[3]) <= 1e-5 * max(1.0, r[3])
            assert abs(row["avg_recall"] - r[1]) <= (1e-5 if case == "toy_l2" else 2.0 / (k * Q))


@pytest.mark.parametrize("mode,value", [(0, 0.3), (1, 0.3), (1, 0.999999), (2, 1), (2, 5), (2, 40)])
def test_probe_search_selection_modes(L, golden, mode, value):
    z = golden("toy_l2")
    k, B, metric = int(z["k"]), int(z["n_bkt"]), int(z["metric"])
    w = [z[f"mlp_{i}"] for i in range(12)]
    index = L.LiraIndex.from_data_2_bkt(z["x_d"], z["d2b1"], B, metric)
    model = L.LiraModel.from_arrays(z["centroids"], z["scaler_mean"], z["scaler_scale"], w)
    scores = model.scores(z["x_q"])
    D, I, nprobe, cmp_ = index.probe_search(model, z["x_q"], mode, value, k)
    # oracle selection on the GPU's own fp32 scores (selection is exact arithmetic on given scores)
    poff, pids = O.select(scores, mode, value)
    assert np.array_equal(nprobe, np.diff(poff))
    off, ids, vecs = O.build_lists_from_data_2_bkt(z["x_d"], z["d2b1"], B)
    I_ref, D_ref, cmp_ref = O.search(off, ids, vecs, z["x_q"], poff, pids, k, metric, O.F64, 1)
    assert np.array_equal(cmp_, cmp_ref)
    assert np.array_equal(I, I_ref) and np.allclose(D, D_ref, rtol=RTOL)


def test_query_tuning_matches_reference_csv(L, golden, tmp_path):
    """create_inner_indexes -> get_cmp_recall -> query_tuning with the reference's call shapes against the
    CSV the reference's own query_tuning wrote for the same index (tests/golden: tuning1)."""
    z = golden("toy_l2")
    k, B = int(z["k"]), int(z["n_bkt"])
    off, ids = z["lists1_off"], z["lists1_ids"]
    cluster_ids = [ids[off[b]:off[b + 1]].tolist() for b in range(B)]

    class Cfg:
        n_bkt, dis_metric, dataset, redundancy_ratio, duplicate_type = B, "L2", "toy", 0.25, "model"
        pth_log, file_name = str(tmp_path) + "/", "toy"
    Cfg.k = k
    inner = L.create_inner_indexes(z["x_d"], cluster_ids, Cfg)
    search_time, cmp_all, found = L.get_cmp_recall(inner, z["x_q"], cluster_ids, Cfg)
    assert found.shape == (len(z["x_q"]), B, k) and np.array_equal(found, z["found1"])
    cnt, knn_ids = L.get_knn_distr_redundancy(z["gt"][:, :k], z["d2b1"], Cfg)
    assert np.array_equal(cnt, z["knn_cnt1"])
    df = L.query_tuning(z["all_outputs"], knn_ids, found, search_time, cmp_all, Cfg, None, part=1)
    got = df[["threshold", "nprobe", "Recall", "Computations"]].to_numpy(np.float64)
    np.testing.assert_allclose(got, z["tuning1"], rtol=1e-9, atol=1e-12)
    assert (df["QPS"] > 0).all()


# ---------------------------------------------------------------------------------------------
# a11 exact kNN
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric,integer,k", [(O.L2, True, 11), (O.L2, False, 100), (O.IP, False, 10)])
def test_knn_matches_oracle(L, metric, integer, k):
    x_d, x_q = synth(20000, 32, 150, seed=21, integer=integer)
    D, I = L.knn(x_d, x_q, k, metric)
    D_ref, I_ref = O.knn(x_d, x_q, k, metric, O.F64)
    assert_topk_equiv(D, I, D_ref, I_ref, x_q, x_d, metric)
    if integer:
        assert np.array_equal(I, I_ref)


def test_self_knn_drops_column_zero_like_compute_knn(L):
    x_d, _ = synth(9000, 16, 1, seed=2, integer=False)
    _, I = L.knn(x_d, x_d[:500], 6, O.L2)
    assert np.array_equal(I[:, 0], np.arange(500))  # self first (distance 0)
    _, I_ref = O.knn(x_d, x_d[:500], 6, O.L2, O.F64)
    assert np.mean(I[:, 1:] == I_ref[:, 1:]) > 0.999


# ---------------------------------------------------------------------------------------------
# tensor-core (tcgen05) scan path: must return the same bits as the exact CUDA-core scan
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric", [O.L2, O.IP])
@pytest.mark.parametrize("k,d", [(10, 128), (10, 96), (100, 64), (1, 20), (10, 200), (16, 256), (10, 8)])
def test_tensor_core_scan_is_bit_identical(L, metric, k, d):
    rng = np.random.RandomState(17 + k + d)
    x_d, x_q = synth(30000, d, 700, seed=31 + d, integer=True)
    B = 24
    cl = random_lists(len(x_d), B, rng, redundancy=0.5, empty=(5,))
    cl[3] = cl[3][:7]  # a list shorter than k
    off, ids, vecs = lists_csr(x_d, cl)
    nprobe = rng.randint(0, 7, len(x_q))
    poff = np.zeros(len(x_q) + 1, np.int64)
    np.cumsum(nprobe, out=poff[1:])
    pids = np.concatenate([rng.choice(B, n, replace=False) for n in nprobe] + [np.empty(0, int)]).astype(np.int32)
    index = L.LiraIndex.from_csr(x_d, off, ids, metric)
    assert index.tensor_core_eligible
    D, I, cmp_ = index.search(x_q, poff, pids, k)
    assert index.last_path == "tensor-core"
    index.set_use_tensor_cores(False)
    D2, I2, cmp2 = index.search(x_q, poff, pids, k)
    assert index.last_path == "cuda-core"
    assert np.array_equal(I, I2) and np.array_equal(D, D2) and np.array_equal(cmp_, cmp2)
    I_ref, D_ref, cmp_ref = O.search(off, ids, vecs, x_q, poff, pids, k, metric, O.F64, 1)
    assert np.array_equal(I, I_ref) and np.array_equal(D, D_ref) and np.array_equal(cmp_, cmp_ref)
    # select-then-collapse semantics as well
    index.set_use_tensor_cores(True)
    D0, I0, _ = index.search(x_q, poff, pids, k, dedup=False)
    I0_ref, D0_ref, _ = O.search(off, ids, vecs, x_q, poff, pids, k, metric, O.F64, 0)
    assert np.array_equal(I0, I0_ref) and np.array_equal(D0, D0_ref)


@pytest.mark.parametrize("metric", [O.L2, O.IP])
@pytest.mark.parametrize("k", [1, 10, 16])
def test_tensor_core_region_compaction_without_seed(L, metric, k, monkeypatch):
    """Every bound starts at +inf (seed pass skipped): each (query, list, half) region fills up and is compacted in the
    kernel again and again, duplicated vectors put ties at the bound (-> exact redo). Results must not change."""
    monkeypatch.setenv("LIRA_TC_NO_SEED", "1")
    rng = np.random.RandomState(5 + k)
    x_d, x_q = synth(20000, 64, 600, seed=77, integer=True)
    x_d[1000:1400] = x_d[600:1000]  # exact duplicates under different ids: ties everywhere
    B = 6
    cl = random_lists(len(x_d), B, rng, redundancy=0.3)
    off, ids, vecs = lists_csr(x_d, cl)
    nprobe = rng.randint(1, 5, len(x_q))
    poff = np.zeros(len(x_q) + 1, np.int64)
    np.cumsum(nprobe, out=poff[1:])
    pids = np.concatenate([rng.choice(B, n, replace=False) for n in nprobe]).astype(np.int32)
    index = L.LiraIndex.from_csr(x_d, off, ids, metric)
    D, I, cmp_ = index.search(x_q, poff, pids, k)
    assert index.last_path == "tensor-core"
    I_ref, D_ref, cmp_ref = O.search(off, ids, vecs, x_q, poff, pids, k, metric, O.F64, 1)
    assert np.array_equal(I, I_ref) and np.array_equal(D, D_ref) and np.array_equal(cmp_, cmp_ref)


@pytest.mark.parametrize("metric", [O.L2, O.IP])
@pytest.mark.parametrize("d,scale", [(96, 1.0), (32, 37.5), (100, 1e-3), (250, 1.0), (400, 5.0)])
def test_tensor_core_approximate_mode_for_real_valued_data(L, metric, d, scale):
    """Real-valued data: fp16 filter with a rigorous error margin + exact fp32 re-rank of every survivor. Ids must equal the
    oracle's except for fp32 ties, distances within 1e-5 relative, at any data scale (the shadow copy is rescaled)."""
    rng = np.random.RandomState(d)
    x_d, x_q = synth(30000, d, 600, seed=8 + d, integer=False)
    x_d, x_q = (x_d * scale).astype(np.float32), (x_q * scale).astype(np.float32)
    if d == 96:   # DEEP-style: unit vectors
        x_d /= np.linalg.norm(x_d, axis=1, keepdims=True)
        x_q /= np.linalg.norm(x_q, axis=1, keepdims=True)
    B = 16
    cl = random_lists(len(x_d), B, rng, redundancy=0.4, empty=(2,))
    off, ids, vecs = lists_csr(x_d, cl)
    index = L.LiraIndex.from_csr(x_d, off, ids, metric)
    assert index.tensor_core_eligible and index.tensor_core_mode == "approximate"
    nprobe = rng.randint(0, 6, len(x_q))
    poff = np.zeros(len(x_q) + 1, np.int64)
    np.cumsum(nprobe, out=poff[1:])
    pids = np.concatenate([rng.choice(B, n, replace=False) for n in nprobe] + [np.empty(0, int)]).astype(np.int32)
    for k in (10, 1, 16):
        D, I, cmp_ = index.search(x_q, poff, pids, k)
        assert index.last_path == "tensor-core"
        I_ref, D_ref, cmp_ref = O.search(off, ids, vecs, x_q, poff, pids, k, metric, O.F64, 1)
        assert_topk_equiv(D, I, D_ref, I_ref, x_q, x_d, metric)
        assert np.array_equal(cmp_, cmp_ref)
    # k > 16 and the pinned CUDA-core scan agree as well
    D, I, _ = index.search(x_q, poff, pids, 40)
    assert index.last_path == "cuda-core"
    index.set_use_tensor_cores(False)
    D2, I2, _ = index.search(x_q, poff, pids, 10)
    assert index.last_path == "cuda-core"
    I_ref, D_ref, _ = O.search(off, ids, vecs, x_q, poff, pids, 10, metric, O.F64, 1)
    assert_topk_equiv(D2, I2, D_ref, I_ref, x_q, x_d, metric)


def test_tensor_core_exact_mode_falls_back_for_real_valued_queries(L):
    x_d, x_q = synth(8000, 32, 300, seed=8, integer=False)  # real-valued queries
    cl = random_lists(len(x_d), 8, np.random.RandomState(1))
    xi, _ = synth(8000, 32, 300, seed=8, integer=True)      # integer base: exact mode
    off, ids, vecs = lists_csr(xi, cl)
    poff = np.arange(len(x_q) + 1, dtype=np.int64) * 2
    pids = np.tile(np.array([0, 3], np.int32), len(x_q))
    index2 = L.LiraIndex.from_csr(xi, off, ids, O.L2)
    assert index2.tensor_core_eligible and index2.tensor_core_mode == "exact"
    D, I, _ = index2.search(x_q, poff, pids, 10)
    assert index2.last_path == "cuda-core"   # the batch check sends it to the CUDA cores
    I_ref, D_ref, _ = O.search(off, ids, vecs, x_q, poff, pids, 10, O.L2, O.F64, 1)
    assert_topk_equiv(D, I, D_ref, I_ref, x_q, xi, O.L2)


def test_tensor_core_query_phase_matches_cuda_cores(L, golden):
    """Whole query phase (features -> MLP -> select -> scan) through both scan implementations."""
    z = golden("toy_l2")
    k, B, metric = int(z["k"]), int(z["n_bkt"]), int(z["metric"])
    w = [z[f"mlp_{i}"] for i in range(12)]
    index = L.LiraIndex.from_data_2_bkt(z["x_d"], z["d2b1"], B, metric)
    model = L.LiraModel.from_arrays(z["centroids"], z["scaler_mean"], z["scaler_scale"], w)
    q = np.tile(z["x_q"], (6, 1))  # 384 queries: above the tensor-core batch threshold
    for mode, value in [(0, 0.1), (1, 0.5), (2, 6)]:
        index.set_use_tensor_cores(True)
        a = index.probe_search(model, q, mode, value, k)
        assert index.last_path == "tensor-core"
        index.set_use_tensor_cores(False)
        b = index.probe_search(model, q, mode, value, k)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_select_without_round_trip_truncation_and_inexact_batches(L):
    """Threshold selection on the tensor-core path sizes its buffers with an upper bound (64 lists per query) instead of a
    host round trip. Queries that select more lists than that, and batches that turn out not to be exact in fp16, must be
    answered again transparently: same results as the CUDA-core path, which always takes the round trip."""
    import torch
    rng = np.random.RandomState(3)
    x_d, x_q = synth(40000, 64, 512, seed=5, integer=True)
    B = 100
    cl = random_lists(len(x_d), B, rng, redundancy=0.2)
    off, ids, vecs = lists_csr(x_d, cl)
    index = L.LiraIndex.from_csr(x_d, off, ids, O.L2)
    scores = rng.rand(len(x_q), B).astype(np.float32)   # threshold 0.25 -> ~75 lists per query (> 64)
    dev = torch.device("cuda:0")
    d_s, d_q = torch.as_tensor(scores, device=dev), torch.as_tensor(x_q, device=dev)
    for thr in (0.25, 0.9, 0.25):   # truncated (cap grows), small probe sets, large again (now below the cap)
        index.set_use_tensor_cores(True)
        a = [t.cpu().numpy() for t in index.select_search_dev(d_s, d_q, L.SELECT_GT, thr, 10)]
        assert index.last_path == "tensor-core"
        index.set_use_tensor_cores(False)
        b = [t.cpu().numpy() for t in index.select_search_dev(d_s, d_q, L.SELECT_GT, thr, 10)]
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        poff = np.zeros(len(x_q) + 1, np.int64)
        np.cumsum((scores > thr).sum(1), out=poff[1:])
        pids = np.nonzero(scores > thr)[1].astype(np.int32)
        I_ref, D_ref, cmp_ref = O.search(off, ids, vecs, x_q, poff, pids, 10, O.L2, O.F64, 1)
        assert np.array_equal(a[1], I_ref) and np.array_equal(a[0], D_ref) and np.array_equal(a[3], cmp_ref)
    # a real-valued query batch on an integer index: the optimistic tensor-core attempt is void, CUDA cores answer
    index.set_use_tensor_cores(True)
    xq_f = (x_q + 0.25).astype(np.float32)
    d_qf = torch.as_tensor(xq_f, device=dev)
    a = [t.cpu().numpy() for t in index.select_search_dev(d_s, d_qf, L.SELECT_GT, 0.9, 10)]
    assert index.last_path == "cuda-core"
    poff = np.zeros(len(x_q) + 1, np.int64)
    np.cumsum((scores > 0.9).sum(1), out=poff[1:])
    pids = np.nonzero(scores > 0.9)[1].astype(np.int32)
    I_ref, D_ref, _ = O.search(off, ids, vecs, xq_f, poff, pids, 10, O.L2, O.F64, 1)
    assert_topk_equiv(a[0], a[1], D_ref, I_ref, xq_f, x_d, O.L2)


@pytest.mark.parametrize("metric", [O.L2, O.IP])
def test_fused_front_end_matches_unfused_flow(L, metric, monkeypatch):
    """lira_probe_search: the fused front end (selection inside the last layer's epilogue, scatter that writes the fp16 query
    rows, no scores in HBM) against the unfused flow (scores -> select_kernel -> scans -> scatter -> gather) on the same
    model: identical ids, distances, nprobe and cmp. Covers score > thr and score >= thr + argmax (with queries that select
    nothing), a threshold that selects more than the optimistic 64 partitions per query (the cap is raised and the batch
    runs again), and a ragged batch size."""
    rng = np.random.RandomState(11)
    B, d, k = 100, 64, 10
    x_d, x_q = synth(30000, d, 777, seed=9, integer=True)
    cl = random_lists(len(x_d), B, rng, redundancy=0.3)
    off, ids, vecs = lists_csr(x_d, cl)
    index = L.LiraIndex.from_csr(x_d, off, ids, metric)
    shapes = [(128, B), (128,), (64, 128), (64,), (128, d), (128,), (64, 128), (64,), (128, 128), (128,), (B, 128), (B,)]
    w = [(rng.randn(*s) * (0.5 / np.sqrt(s[-1]) if len(s) == 2 else 0.1)).astype(np.float32) for s in shapes]
    w[4] = (w[4] / 128.0).astype(np.float32)   # raw vectors are O(100): keep vector_net's activations O(1)
    cent = x_d[rng.choice(len(x_d), B, replace=False)]
    f = O.features_cpp(x_d[:2000], cent, None, None)
    mean = f.mean(0).astype(np.float32)
    scale = f.std(0).astype(np.float32)
    model = L.LiraModel.from_arrays(cent, mean, scale, w)
    s = model.scores(x_q)
    lo, hi = float(np.quantile(s, 0.2)), float(np.quantile(s.max(1), 0.5))
    for mode, value in [(L.SELECT_GT, float(np.quantile(s, 0.9))), (L.SELECT_GE_ARGMAX, hi), (L.SELECT_GT, lo), (L.SELECT_GE_ARGMAX, float(np.quantile(s, 0.95)))]:
        monkeypatch.delenv("LIRA_NO_FUSED", raising=False)
        a = index.probe_search(model, x_q, mode, value, k)
        assert index.last_path == "tensor-core"
        monkeypatch.setenv("LIRA_NO_FUSED", "1")
        b = index.probe_search(model, x_q, mode, value, k)
        assert index.last_path == "tensor-core"
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        # and against the oracle's scan of the same probe sets (selection replayed on the library's own scores)
        sel = (s > np.float32(value)) if mode == L.SELECT_GT else (s >= np.float32(value))
        if mode == L.SELECT_GE_ARGMAX:
            none = ~sel.any(1)
            sel[none, s[none].argmax(1)] = True
        edge = np.abs(s - np.float32(value)) < 1e-6   # (scores are recomputed per call bit-identically; still, skip exact edges)
        rows = ~edge.any(1)
        poff = np.zeros(len(x_q) + 1, np.int64)
        np.cumsum(sel.sum(1), out=poff[1:])
        pids = np.nonzero(sel)[1].astype(np.int32)
        I_ref, D_ref, cmp_ref = O.search(off, ids, vecs, x_q, poff, pids, k, metric, O.F64, 1)
        assert np.array_equal(a[1][rows], I_ref[rows]) and np.array_equal(a[3][rows], cmp_ref[rows])
        assert np.array_equal(a[2][rows], sel.sum(1)[rows])
    monkeypatch.delenv("LIRA_NO_FUSED", raising=False)


def test_asynchronous_forms_match_the_synchronous_call(L):
    """lira_probe_search_enqueue_dev + lira_index_finish and lira_probe_search_submit / _wait against lira_probe_search on the
    same batches: several batches in flight, a batch whose optimistic run is void (more than 64 partitions selected: answered
    again by finish / wait), a real-valued batch on an integer index, and a batch too small for the fused flow."""
    import torch
    rng = np.random.RandomState(21)
    B, d, k = 100, 64, 10
    x_d, x_q = synth(30000, d, 600, seed=13, integer=True)
    cl = random_lists(len(x_d), B, rng, redundancy=0.3)
    off, ids, vecs = lists_csr(x_d, cl)
    index = L.LiraIndex.from_csr(x_d, off, ids, O.L2)
    shapes = [(128, B), (128,), (64, 128), (64,), (128, d), (128,), (64, 128), (64,), (128, 128), (128,), (B, 128), (B,)]
    w = [(rng.randn(*s) * (0.5 / np.sqrt(s[-1]) if len(s) == 2 else 0.1)).astype(np.float32) for s in shapes]
    w[4] = (w[4] / 128.0).astype(np.float32)
    cent = x_d[rng.choice(len(x_d), B, replace=False)]
    f = O.features_cpp(x_d[:2000], cent, None, None)
    model = L.LiraModel.from_arrays(cent, f.mean(0).astype(np.float32), f.std(0).astype(np.float32), w)
    s = model.scores(x_q)
    thr_small, thr_big = float(np.quantile(s, 0.9)), float(np.quantile(s, 0.2))   # ~10 / ~80 partitions per query
    batches = [(x_q, thr_small), (x_q[:300], thr_small), (x_q, thr_big), ((x_q + 0.25).astype(np.float32), thr_small), (x_q[:100], thr_small)]
    ref = [index.probe_search(model, q, L.SELECT_GT, t, k) for q, t in batches]
    index2 = L.LiraIndex.from_csr(x_d, off, ids, O.L2)   # fresh handle: the partition cap starts at 64 again
    dev = torch.device("cuda:0")
    # device form: everything enqueued, one finish
    outs = [index2.probe_search_enqueue_dev(model, torch.as_tensor(q, device=dev), L.SELECT_GT, t, k) for q, t in batches]
    index2.finish()
    torch.cuda.synchronize()
    for o, r in zip(outs, ref):
        for x, y in zip(o, r):
            assert np.array_equal(x.cpu().numpy(), y)
    # host form: two slots, submit(i + 1) before wait(i)
    index3 = L.LiraIndex.from_csr(x_d, off, ids, O.L2)
    got = []
    index3.probe_search_submit(model, batches[0][0], L.SELECT_GT, batches[0][1], k, slot=0)
    for i in range(len(batches)):
        if i + 1 < len(batches):
            index3.probe_search_submit(model, batches[i + 1][0], L.SELECT_GT, batches[i + 1][1], k, slot=(i + 1) & 1)
        got.append(index3.probe_search_wait(slot=i & 1))
    for o, r in zip(got, ref):
        for x, y in zip(o, r):
            assert np.array_equal(x, y)


# ---------------------------------------------------------------------------------------------
# tensor-core (tcgen05, error-compensated TF32) forward of the probing model vs the fp32 CUDA-core kernels
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["toy_l2", "toy_ip"])
def test_model_tensor_core_forward_matches_cuda_cores(L, golden, case):
    z = golden(case)
    w = [z[f"mlp_{i}"] for i in range(12)]
    model = L.LiraModel.from_arrays(z["centroids"], z["scaler_mean"], z["scaler_scale"], w)
    q = np.tile(z["x_q"], (5, 1))[:301]  # not a multiple of the 128-row tile
    model.set_use_tensor_cores(True)
    s_tc, f_tc = model.scores(q, return_features=True)
    model.set_use_tensor_cores(False)
    s_cc, f_cc = model.scores(q, return_features=True)
    # both are fp32-accurate evaluations of the same network: compare with the fp64 arbiter's tolerance
    np.testing.assert_allclose(f_tc, f_cc, rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(s_tc, s_cc, rtol=1e-4, atol=2e-6)
    _, probs, _ = O.mlp_forward(O.features_cpp(q, z["centroids"], z["scaler_mean"], z["scaler_scale"]), q, w)
    np.testing.assert_allclose(s_tc, probs, rtol=1e-4, atol=2e-6)


def test_model_tensor_core_forward_large_offset_data(L):
    """SIFT-like data far from the origin: the centred expansion must keep fp32-level feature accuracy."""
    rng = np.random.RandomState(3)
    B, d = 40, 128
    x_d, x_q = synth(6000, d, 257, seed=12, integer=True)   # values 0..255, norms ~1e6
    cent = x_d[rng.choice(len(x_d), B, replace=False)] + rng.rand(B, d).astype(np.float32)
    f = O.features_cpp(x_d[:2000], cent)
    mean, scale = f.mean(0).astype(np.float32), f.std(0).astype(np.float32)
    shapes = [(128, B), (128,), (64, 128), (64,), (128, d), (128,), (64, 128), (64,), (128, 128), (128,), (B, 128), (B,)]
    w = [(rng.randn(*s) * (0.02 if len(s) == 2 else 0.1)).astype(np.float32) for s in shapes]
    model = L.LiraModel.from_arrays(cent, mean, scale, w)
    s_tc, f_tc = model.scores(x_q, return_features=True)
    f64 = O.features_py(x_q, cent, mean.astype(np.float64), scale.astype(np.float64))  # fp64 cdist arbiter
    np.testing.assert_allclose(f_tc, f64, rtol=1e-5, atol=2e-5)
    _, probs, _ = O.mlp_forward(f64, x_q, w)
    np.testing.assert_allclose(s_tc, probs, rtol=1e-4, atol=2e-6)


def test_knn_tensor_core_path_matches_oracle(L):
    """Exact kNN over base segments on the tensor-core scan (integer data, >= 256 queries, k <= 16)."""
    x_d, x_q = synth(40000, 64, 300, seed=77, integer=True)
    k = 11  # compute_knn's k + 1 for k = 10 (compute_knn.cpp:237)
    D, I = L.knn(x_d, x_q, k, O.L2)
    D_ref, I_ref = O.knn(x_d, x_q, k, O.L2, O.F64)
    assert np.array_equal(I, I_ref) and np.array_equal(D, D_ref.astype(np.float32))
    # self-kNN shape: the query itself comes first
    D2, I2 = L.knn(x_d, x_d[:512], k, O.L2)
    D2_ref, I2_ref = O.knn(x_d, x_d[:512], k, O.L2, O.F64)
    assert np.array_equal(I2, I2_ref) and np.array_equal(D2, D2_ref.astype(np.float32))


# ---------------------------------------------------------------------------------------------
# BASELINE.json full size (config 1 shape: 1 M x 128, B = 1024, 10 000 queries, k = 10): size-independent properties
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("real", [False, True])
def test_full_size_properties(L, real):
    """At the full config-1 size the oracle is too slow for every query, so the result is checked through properties:
    rows sorted best-first, ids distinct and members of the probed lists, distances equal to the exact value recomputed
    from the returned ids, cmp = sum of the probed list sizes, idempotence, the tensor-core scan equal to the CUDA-core
    scan on every row (bit-identical for integer data, up to fp32 ties for real-valued data), and the oracle on a sample."""
    rng = np.random.RandomState(11)
    N, d, Q, B, k, nprobe = 1_000_000, 128, 10_000, 1024, 10, 10
    centres = rng.randn(256, d).astype(np.float32)
    x_d = centres[rng.randint(0, 256, N)] + 0.5 * rng.randn(N, d).astype(np.float32)
    x_q = centres[rng.randint(0, 256, Q)] + 0.5 * rng.randn(Q, d).astype(np.float32)
    if real:
        x_d /= np.linalg.norm(x_d, axis=1, keepdims=True)
        x_q /= np.linalg.norm(x_q, axis=1, keepdims=True)
    else:
        x_d = np.clip(np.round(24 * x_d + 100), 0, 255).astype(np.float32)
        x_q = np.clip(np.round(24 * x_q + 100), 0, 255).astype(np.float32)
    d2b = np.full((N, 2), -1, np.int32)
    d2b[:, 0] = rng.randint(0, B, N)
    red = rng.choice(N, N * 3 // 100, replace=False)           # 3 % of the points get a second list
    d2b[red, 1] = (d2b[red, 0] + 1 + rng.randint(0, B - 1, len(red))) % B
    index = L.LiraIndex.from_data_2_bkt(x_d, d2b, B, O.L2)
    assert index.tensor_core_mode == ("approximate" if real else "exact")
    pids = np.stack([rng.choice(B, nprobe, replace=False) for _ in range(Q)]).astype(np.int32)
    poff = np.arange(Q + 1, dtype=np.int64) * nprobe
    D, I, cmp_ = index.search(x_q, poff, pids.reshape(-1), k)
    assert index.last_path == "tensor-core"
    # sorted best-first, full rows, distinct ids
    assert np.all(np.diff(D, axis=1) >= 0) and np.all(I >= 0)
    assert all(len(set(r)) == k for r in I[::97].tolist())
    # every id lives in one of the probed lists of its query
    member = (d2b[I][:, :, :, None] == pids[:, None, None, :]).any(axis=(2, 3))
    assert member.all()
    # distances are the exact values of the returned ids
    exact = ((x_q[:, None, :].astype(np.float64) - x_d[I].astype(np.float64)) ** 2).sum(-1)
    assert np.allclose(D, exact, rtol=1e-5, atol=1e-6)
    # Computations column: sum of the probed list sizes
    sizes = index.list_sizes()
    assert np.array_equal(cmp_, sizes[pids].sum(1))
    # idempotence, and the CUDA-core scan returns the same rows
    D2, I2, _ = index.search(x_q, poff, pids.reshape(-1), k)
    assert np.array_equal(D, D2) and np.array_equal(I, I2)
    index.set_use_tensor_cores(False)
    Dc, Ic, cc = index.search(x_q, poff, pids.reshape(-1), k)
    assert index.last_path == "cuda-core" and np.array_equal(cc, cmp_)
    if real:
        assert np.allclose(D, Dc, rtol=1e-5, atol=1e-6) and (I == Ic).all(1).mean() > 0.999
    else:
        assert np.array_equal(D, Dc) and np.array_equal(I, Ic)
    # the oracle on a sample of the queries
    off, ids, vecs = O.build_lists_from_data_2_bkt(x_d, d2b, B)
    sel = np.arange(0, Q, 250)
    po = np.arange(len(sel) + 1, dtype=np.int64) * nprobe
    I_ref, D_ref, _ = O.search(off, ids, vecs, x_q[sel], po, pids[sel].reshape(-1), k, O.L2, O.F64, 1)
    if real:
        assert_topk_equiv(D[sel], I[sel], D_ref, I_ref, x_q[sel], x_d, O.L2)
    else:
        assert np.array_equal(I[sel], I_ref) and np.array_equal(D[sel], D_ref)


def test_gist_shape_top_nprobe_sweep(L):
    """Config 3 shape (GIST: d = 960, real-valued, top-nprobe selection, nprobe swept): beyond d = 256 the query tile is
    streamed through the ring together with the list rows (both operands, one K block per slot); approximate mode +
    exact re-rank. Ids / distances vs the oracle, and vs the fp32 CUDA-core scan."""
    import torch
    rng = np.random.RandomState(960)
    x_d, x_q = synth(12000, 960, 300, seed=3, integer=False)
    x_d, x_q = (0.1 * x_d + 0.5).astype(np.float32), (0.1 * x_q + 0.5).astype(np.float32)   # [0, 1]-ish like GIST
    B, k = 32, 10
    cl = random_lists(len(x_d), B, rng, redundancy=0.1)
    off, ids, vecs = lists_csr(x_d, cl)
    index = L.LiraIndex.from_csr(x_d, off, ids, O.L2)
    assert index.tensor_core_mode == "approximate"
    scores = rng.rand(len(x_q), B).astype(np.float32)
    dev = torch.device("cuda:0")
    d_s, d_q = torch.as_tensor(scores, device=dev), torch.as_tensor(x_q, device=dev)
    for nprobe in (1, 2, 8, 32):
        index.set_use_tensor_cores(True)
        D, I, npb, cmp_ = [t.cpu().numpy() for t in index.select_search_dev(d_s, d_q, L.SELECT_TOPN, nprobe, k)]
        assert index.last_path == "tensor-core" and np.all(npb == nprobe)
        poff, pids = O.select(scores, O.SELECT_TOPN, nprobe)
        I_ref, D_ref, cmp_ref = O.search(off, ids, vecs, x_q, poff, pids, k, O.L2, O.F64, 1)
        assert_topk_equiv(D, I, D_ref, I_ref, x_q, x_d, O.L2)
        assert np.array_equal(cmp_, cmp_ref)
        index.set_use_tensor_cores(False)
        Dc, Ic, _, _ = [t.cpu().numpy() for t in index.select_search_dev(d_s, d_q, L.SELECT_TOPN, nprobe, k)]
        assert index.last_path == "cuda-core"
        assert_topk_equiv(Dc, Ic, D_ref, I_ref, x_q, x_d, O.L2)
    # integer data of the same width: exact mode through the streaming ring, bit-identical
    xi_d, xi_q = synth(6000, 300, 300, seed=4, integer=True)
    cl = random_lists(len(xi_d), 8, rng, redundancy=0.1)
    off, ids, vecs = lists_csr(xi_d, cl)
    index = L.LiraIndex.from_csr(xi_d, off, ids, O.L2)
    assert index.tensor_core_mode == "exact"
    poff = np.arange(len(xi_q) + 1, dtype=np.int64) * 3
    pids = np.stack([rng.choice(8, 3, replace=False) for _ in range(len(xi_q))]).astype(np.int32).reshape(-1)
    D, I, _ = index.search(xi_q, poff, pids, k)
    assert index.last_path == "tensor-core"
    I_ref, D_ref, _ = O.search(off, ids, vecs, xi_q, poff, pids, k, O.L2, O.F64, 1)
    assert np.array_equal(I, I_ref) and np.array_equal(D, D_ref)
    # beyond d = 1024: CUDA cores
    x_w, _ = synth(600, 1100, 4, seed=5, integer=False)
    assert L.LiraIndex.from_csr(x_w, np.array([0, 600], np.int64), np.arange(600, dtype=np.int32), O.L2).tensor_core_mode == "none"


@pytest.mark.parametrize("Q", [256, 257, 383, 1000])
def test_tensor_core_batch_size_edges(L, Q):
    """Batch sizes around the tensor-core threshold (256) and off every tile size; lists of 0, 1 and 129 entries; a query
    that probes nothing and one that probes the same list twice."""
    rng = np.random.RandomState(Q)
    x_d, x_q = synth(9000, 48, Q, seed=Q, integer=True)
    B = 12
    cl = random_lists(len(x_d), B, rng, redundancy=0.3, empty=(4,))
    cl[5] = cl[5][:1]
    cl[6] = cl[6][:129]
    off, ids, vecs = lists_csr(x_d, cl)
    nprobe = rng.randint(1, 5, Q)
    nprobe[0] = 0
    sets = [rng.choice(B, n, replace=False) for n in nprobe]
    sets[1] = np.array([6, 6, 5])   # a repeated list
    poff = np.zeros(Q + 1, np.int64)
    np.cumsum([len(x) for x in sets], out=poff[1:])
    pids = np.concatenate(sets + [np.empty(0, int)]).astype(np.int32)
    index = L.LiraIndex.from_csr(x_d, off, ids, O.L2)
    for k in (10, 3):
        D, I, cmp_ = index.search(x_q, poff, pids, k)
        assert index.last_path == "tensor-core"
        I_ref, D_ref, cmp_ref = O.search(off, ids, vecs, x_q, poff, pids, k, O.L2, O.F64, 1)
        assert np.array_equal(I, I_ref) and np.array_equal(D, D_ref) and np.array_equal(cmp_, cmp_ref)



@pytest.mark.parametrize("metric", [O.L2, O.IP])
@pytest.mark.parametrize("k,d", [(10, 128), (10, 96), (1, 20), (16, 64), (10, 8), (10, 200)])
def test_byte_scan_is_bit_identical(L, metric, k, d, monkeypatch):
    """LIRA_U8_SEARCH=1: explicit probe sets on the kind::i8 scan (one byte per component, int32 accumulators). Lists longer
    than one row segment (4096 rows), a group of more than 512 queries (several work items per list), an empty list, a list
    shorter than k, ids stored in two lists: ids, distances and cmp equal the oracle's and the CUDA-core scan's bit for bit."""
    monkeypatch.setenv("LIRA_U8_SEARCH", "1")
    rng = np.random.RandomState(170 + k + d)
    x_d, x_q = synth(40000, d, 900, seed=310 + d, integer=True)
    B = 12
    cl = random_lists(len(x_d), B, rng, redundancy=0.5, empty=(5,))
    cl[3] = cl[3][:7]  # a list shorter than k
    off, ids, vecs = lists_csr(x_d, cl)
    assert (np.diff(off) > 4096).any()
    nprobe = rng.randint(0, 7, len(x_q))
    nprobe[:700] = np.maximum(nprobe[:700], 1)
    poff = np.zeros(len(x_q) + 1, np.int64)
    np.cumsum(nprobe, out=poff[1:])
    sets = [rng.choice(B, n, replace=False) for n in nprobe]
    for i in range(700):
        sets[i][0] = 0 if 0 not in sets[i][1:] else sets[i][0]   # list 0 is probed by more than 512 queries
    pids = np.concatenate(sets + [np.empty(0, int)]).astype(np.int32)
    index = L.LiraIndex.from_csr(x_d, off, ids, metric)
    assert index.byte_scan_eligible == (d <= 128)   # (the byte scan takes d <= 128; longer rows keep the fp16 scan)
    D, I, cmp_ = index.search(x_q, poff, pids, k)
    assert index.last_path == "tensor-core" and index.last_scan_kind == ("u8" if d <= 128 else "fp16")
    I_ref, D_ref, cmp_ref = O.search(off, ids, vecs, x_q, poff, pids, k, metric, O.F64, 1)
    assert np.array_equal(I, I_ref) and np.array_equal(D, D_ref) and np.array_equal(cmp_, cmp_ref)
    D0, I0, _ = index.search(x_q, poff, pids, k, dedup=False)
    I0_ref, D0_ref, _ = O.search(off, ids, vecs, x_q, poff, pids, k, metric, O.F64, 0)
    assert np.array_equal(I0, I0_ref) and np.array_equal(D0, D0_ref)
    # a query batch that is not byte valued falls back to the fp16 scan (same handle), same result
    x_q2 = x_q.copy()
    x_q2[0, 0] = 300.0
    D2, I2, _ = index.search(x_q2, poff, pids, k)
    assert index.last_scan_kind == "fp16"
    I2_ref, D2_ref, _ = O.search(off, ids, vecs, x_q2, poff, pids, k, metric, O.F64, 1)
    assert np.array_equal(I2, I2_ref) and np.array_equal(D2, D2_ref)


def test_byte_scan_without_seed_overflows_to_the_exact_path(L, monkeypatch):
    """LIRA_TC_NO_SEED=1: every bound is +inf, every region overflows, every query is redone by the CUDA-core scan."""
    monkeypatch.setenv("LIRA_U8_SEARCH", "1")
    monkeypatch.setenv("LIRA_TC_NO_SEED", "1")
    rng = np.random.RandomState(9)
    x_d, x_q = synth(20000, 64, 400, seed=78, integer=True)
    B = 6
    cl = random_lists(len(x_d), B, rng, redundancy=0.3)
    off, ids, vecs = lists_csr(x_d, cl)
    nprobe = rng.randint(1, 5, len(x_q))
    poff = np.zeros(len(x_q) + 1, np.int64)
    np.cumsum(nprobe, out=poff[1:])
    pids = np.concatenate([rng.choice(B, n, replace=False) for n in nprobe]).astype(np.int32)
    index = L.LiraIndex.from_csr(x_d, off, ids, O.L2)
    D, I, cmp_ = index.search(x_q, poff, pids, 10)
    assert index.last_scan_kind == "u8" and index.last_redo == len(x_q)
    I_ref, D_ref, cmp_ref = O.search(off, ids, vecs, x_q, poff, pids, 10, O.L2, O.F64, 1)
    assert np.array_equal(I, I_ref) and np.array_equal(D, D_ref) and np.array_equal(cmp_, cmp_ref)


_PDL_CHILD = r'''
import sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import lira_ann_search_b200 as L
z = np.load(sys.argv[2])
index = L.LiraIndex.from_csr(z["x_d"], z["off"], z["ids"], int(z["metric"]))
model = L.LiraModel.from_arrays(z["cent"], z["mean"], z["scale"], [z[f"w{i}"] for i in range(12)])
out = {}
for j, (mode, value) in enumerate(zip(z["modes"], z["values"])):
    D, I, n, c = index.probe_search(model, z["x_q"], int(mode), float(value), 10)
    assert index.last_path == "tensor-core"
    out[f"D{j}"], out[f"I{j}"], out[f"n{j}"], out[f"c{j}"] = D, I, n, c
np.savez(sys.argv[3], **out)
'''


def test_programmatic_launch_chain_changes_nothing(L, tmp_path):
    """The batch's kernels are chained by programmatic dependent launch (each kernel's CTAs start while the previous one
    drains and wait in griddepcontrol.wait): a process with LIRA_NO_PDL=1 (every launch fully serialised) must produce the
    same bytes -- ids, distances, nprobe, cmp -- for both selection modes."""
    import os
    import subprocess
    import sys
    rng = np.random.RandomState(31)
    B, d, k = 100, 64, 10
    x_d, x_q = synth(30000, d, 700, seed=17, integer=True)
    cl = random_lists(len(x_d), B, rng, redundancy=0.3)
    off, ids, vecs = lists_csr(x_d, cl)
    shapes = [(128, B), (128,), (64, 128), (64,), (128, d), (128,), (64, 128), (64,), (128, 128), (128,), (B, 128), (B,)]
    w = [(rng.randn(*s) * (0.5 / np.sqrt(s[-1]) if len(s) == 2 else 0.1)).astype(np.float32) for s in shapes]
    w[4] = (w[4] / 128.0).astype(np.float32)
    cent = x_d[rng.choice(len(x_d), B, replace=False)]
    f = O.features_cpp(x_d[:2000], cent, None, None)
    mean, scale = f.mean(0).astype(np.float32), f.std(0).astype(np.float32)
    model = L.LiraModel.from_arrays(cent, mean, scale, w)
    s = model.scores(x_q)
    modes = [L.SELECT_GT, L.SELECT_GE_ARGMAX]
    values = [float(np.quantile(s, 0.9)), float(np.quantile(s.max(1), 0.5))]
    inp, outp = str(tmp_path / "in.npz"), str(tmp_path / "out.npz")
    np.savez(inp, x_d=x_d, x_q=x_q, off=off, ids=ids, metric=O.L2, cent=cent, mean=mean, scale=scale, modes=np.array(modes),
             values=np.array(values), **{f"w{i}": w[i] for i in range(12)})
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, LIRA_NO_PDL="1")
    subprocess.run([sys.executable, "-c", _PDL_CHILD, root, inp, outp], check=True, env=env, timeout=600)
    z = np.load(outp)
    index = L.LiraIndex.from_csr(x_d, off, ids, O.L2)
    for j, (mode, value) in enumerate(zip(modes, values)):
        D, I, n, c = index.probe_search(model, x_q, mode, value, k)
        assert index.last_path == "tensor-core"
        assert np.array_equal(I, z[f"I{j}"]) and np.array_equal(D, z[f"D{j}"])
        assert np.array_equal(n, z[f"n{j}"]) and np.array_equal(c, z[f"c{j}"])
