"""GPU: the build-side pieces -- K-Means on the device, scaler statistics and centroid features without a host
round trip, the redundancy rule kernel -- against numpy restatements / the reference-pinned host mirrors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import lira_ann_search_b200 as L
    L._cabi.require_gpu()
    return L


def lloyd_numpy(x, c, niter):
    x64 = x.astype(np.float64)
    for _ in range(niter):
        d2 = (x64 * x64).sum(1)[:, None] - 2.0 * x64 @ c.astype(np.float64).T + (c.astype(np.float64) ** 2).sum(1)[None, :]
        a = d2.argmin(1)
        for b in range(len(c)):
            m = a == b
            if m.any():
                c[b] = x[m].astype(np.float64).mean(0)
    return c, a


@pytest.mark.parametrize("integer", [True, False])
def test_kmeans_matches_a_numpy_lloyd_from_the_same_start(L, integer):
    from helpers import synth
    x, _ = synth(10000, 24, 8, seed=17, integer=integer, ncomp=30)   # <= 256 points per centroid: no subsampling
    B = 40
    rng = np.random.RandomState(0)
    init = x[rng.choice(len(x), B, replace=False)].copy()
    got = L.engine.kmeans_train(x, B, niter=6, init_centroids=init)
    ref, a_ref = lloyd_numpy(x, init.copy(), 6)
    # fp32 atomic sums against fp64 means; an assignment may flip where two centroids are equally near
    assert np.allclose(got, ref, rtol=2e-3, atol=2e-3 * np.abs(ref).max()), np.abs(got - ref).max()
    # the objective of the library's centroids is that of the reference run
    def obj(c):
        x64, c64 = x.astype(np.float64), c.astype(np.float64)
        d2 = (x64 ** 2).sum(1)[:, None] - 2.0 * x64 @ c64.T + (c64 ** 2).sum(1)[None]
        return d2.min(1).mean()
    assert abs(obj(got) - obj(ref)) <= 1e-3 * obj(ref)


def test_kmeans_default_start_subsampling_and_empty_clusters(L):
    """build_kmeans_index shape (utils.py:321-330): random start, more than 256 points per centroid (subsampled), and a start
    that leaves clusters empty (duplicated initial centroids) which must be re-seeded."""
    from helpers import synth
    x, _ = synth(30000, 16, 8, seed=5, integer=False, ncomp=12)
    km = L.utils.Kmeans(16, 8, niter=10).train(x)       # 30000 > 256 * 8: subsampled
    d2 = ((x[:, None, :] - km.centroids[None]) ** 2).sum(-1)
    base = ((x - x.mean(0)) ** 2).sum(1).mean()
    assert d2.min(1).mean() < 0.6 * base
    assert len(np.unique(d2.argmin(1))) == 8
    kmeans, single, cnts, ids = L.build_kmeans_index(x, 8)
    assert single.shape == (len(x), 1) and cnts.sum() == len(x) and sum(len(i) for i in ids) == len(x)
    assert np.array_equal(single[:, 0], ((x[:, None, :].astype(np.float64) - kmeans.centroids[None].astype(np.float64)) ** 2).sum(-1).argmin(1))
    init = np.repeat(x[:4], 2, axis=0).copy()           # 8 centroids, pairwise equal: 4 clusters start empty
    c = L.engine.kmeans_train(x, 8, niter=8, init_centroids=init)
    d2 = ((x[:, None, :] - c[None]) ** 2).sum(-1)
    assert len(np.unique(d2.argmin(1))) == 8


def test_feature_stats_and_device_features_match_the_host_mirror(L, golden):
    import torch
    z = golden("toy_large")
    dev = torch.device("cuda:0")
    xs = z["x_d"][z["sub_idx"]]
    x_dev, c_dev = torch.as_tensor(xs, device=dev), torch.as_tensor(z["centroids"], device=dev)
    mean, var = L.engine.feature_stats_dev(x_dev, c_dev)
    assert np.allclose(mean, z["scaler_mean"], rtol=2e-6) and np.allclose(np.sqrt(var), z["scaler_scale"], rtol=2e-5)
    f = L.engine.centroid_features_dev(torch.as_tensor(z["x_q"], device=dev), c_dev, torch.as_tensor(z["scaler_mean"], device=dev),
                                       torch.as_tensor(z["scaler_scale"], device=dev))
    assert np.allclose(f.cpu().numpy(), z["dist_q_scaled"], atol=2e-5)
    f0 = L.engine.centroid_features_dev(x_dev[:100], c_dev)
    assert np.allclose(f0.cpu().numpy(), L.centroid_features(xs[:100], z["centroids"]), rtol=1e-6)


@pytest.mark.parametrize("n_mul", [2, 3, 5])
def test_redundancy_rule_kernel_matches_the_pinned_host_mirror(L, n_mul):
    """lira_mul_partition_dev against query.mul_partition_by_model_large's numpy path (itself pinned against the reference's
    loop in tests/test_oracle_golden.py / test_host_mirror_cpu.py): same data_2_bkt, counts and list appends, with score ties,
    rows without any prediction and current partitions inside and outside the best n_mul."""
    import torch
    import lira_ann_search_b200.query as Qm
    rng = np.random.RandomState(n_mul)
    n, B = 5000, 37
    score = rng.rand(n, B).astype(np.float32)
    score[rng.rand(n, B) < 0.3] = np.float32(0.25)      # ties
    score[:200] *= np.float32(0.4)                       # nothing above 0.5
    score[200:400, :] = np.float32(0.75)                 # everything tied and predicted
    cur = rng.randint(0, B, n)
    cur[400:1400] = score[400:1400].argmax(1)            # current partition = the best one
    def start():
        d2b = np.full((n, n_mul), -1)
        d2b[:, 0] = cur
        cnts = np.bincount(cur, minlength=B).astype(np.int64)
        ids = [np.nonzero(cur == b)[0].tolist() for b in range(B)]
        return d2b, cnts, ids
    t = np.arange(1000, 1000 + n)                        # global ids offset from the rows
    a = start()
    big = np.full((1000 + n, n_mul), -1); big[1000:] = a[0]
    Qm.mul_partition_by_model_large(torch.as_tensor(score), torch.as_tensor(score > 0.5), t, 1000, big, a[1], a[2])
    b = start()
    big2 = np.full((1000 + n, n_mul), -1); big2[1000:] = b[0]
    Qm.mul_partition_by_model_large(torch.as_tensor(score, device="cuda:0"), None, t, 1000, big2, b[1], b[2])
    assert np.array_equal(big, big2)
    assert np.array_equal(a[1], b[1])
    assert a[2] == b[2]
