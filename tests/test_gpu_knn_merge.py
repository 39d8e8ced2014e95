"""-m gpu: the fast paths behind the kNN and multi-GPU numbers, through the C ABI, against the oracle.

  * exact kNN on the tensor-core scan with k > 16 (exhaustive probe sets: CUDA-core two-segment seed, 64-slot regions
    without compaction, refine_topk_kernel<4>, overflow -> exact redo) -- compute_knn.cpp:208-259 with --queries k = 100;
  * exact kNN with k <= 16 on a base of >= 524 288 rows (65 536-row segments);
  * the self-kNN column-0 drop with exact duplicates (compute_knn.cpp:254-259);
  * lira_pack_keys_dev + lira_merge_ranks_dev with ids present on several ranks.
"""
import numpy as np
import pytest

import oracle as O
from helpers import assert_topk_equiv, merge_ranks_numpy, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import lira_ann_search_b200 as lib
    lib._cabi.require_gpu()
    return lib


@pytest.mark.parametrize("metric", [O.L2, O.IP])
@pytest.mark.parametrize("k", [17, 100])
def test_knn_exhaustive_tensor_core_path_k_above_16(L, metric, k):
    """>= 256 integer queries, N >= 65 536: the tcgen05 scan answers (last_path), ids and distances equal the oracle's
    direct form bit for bit (integer data: every product and sum is exact) and Faiss's BLAS form within 1e-5."""
    x_d, x_q = synth(70_000, 64, 300, seed=100 + k, integer=True)
    index = L.KnnIndex(x_d, metric)
    D, I = index.search(x_q, k)
    assert index.last_path == "tensor-core" and index.last_scan_kind == "u8"   # byte-valued base: the kind::i8 scan
    D_ref, I_ref = O.knn(x_d, x_q, k, metric, O.F64, 0)
    assert np.array_equal(I, I_ref) and np.array_equal(D, D_ref)
    D_blas, I_blas = O.knn(x_d, x_q, k, metric, O.F32, 1)   # |x|^2 + |y|^2 - 2 x.y in fp32 (faiss nx >= 20 path)
    assert_topk_equiv(D, I, D_blas, I_blas, x_q, x_d, metric)
    # the CUDA-core scan returns the same rows
    index.set_use_tensor_cores(False)
    Dc, Ic = index.search(x_q, k)
    assert index.last_path == "cuda-core" and np.array_equal(Ic, I) and np.array_equal(Dc, D)
    # one-shot entry point
    D1, I1 = L.knn(x_d, x_q, k, metric)
    assert np.array_equal(I1, I) and np.array_equal(D1, D)


@pytest.mark.parametrize("scan", ["fp16", "u8"])
def test_knn_k100_overflow_is_redone_exactly(L, scan, monkeypatch):
    """fp16 scan: the static bound of the k > 16 path comes from the first two base segments. Here they hold a far-away
    cluster, so the bound is loose, the 64-slot candidate regions of the near segments overflow and the flagged queries are
    answered again by the exact CUDA-core scan (last_redo > 0): the result must not change. Byte scan: its seed pools
    the first 16 segments, here the whole base, so nothing overflows; same result."""
    if scan == "fp16":
        monkeypatch.setenv("LIRA_NO_U8", "1")   # (read when the index is created)
    rng = np.random.RandomState(5)
    d, k = 32, 100
    far = np.clip(np.round(rng.randn(16_384, d) * 4 + 200), 0, 255)
    near = np.clip(np.round(rng.randn(60_000, d) * 4 + 40), 0, 255)
    x_d = np.concatenate([far, near]).astype(np.float32)
    x_q = np.clip(np.round(rng.randn(400, d) * 4 + 40), 0, 255).astype(np.float32)
    index = L.KnnIndex(x_d, O.L2)
    D, I = index.search(x_q, k)
    assert index.last_path == "tensor-core" and index.last_scan_kind == scan
    assert index.last_redo > 0 if scan == "fp16" else index.last_redo == 0
    D_ref, I_ref = O.knn(x_d, x_q, k, O.L2, O.F64, 0)
    assert np.array_equal(I, I_ref) and np.array_equal(D, D_ref)


def test_knn_long_segments_on_a_large_base(L):
    """N >= 524 288 and k <= 16: 65 536-row segments. Oracle on every query."""
    x_d, x_q = synth(600_000, 32, 320, seed=9, integer=True, ncomp=200)
    k = 11   # compute_knn's k + 1 for k = 10 (compute_knn.cpp:237)
    index = L.KnnIndex(x_d, O.L2)
    D, I = index.search(x_q, k)
    assert index.last_path == "tensor-core"
    D_ref, I_ref = O.knn(x_d, x_q, k, O.L2, O.F64, 0)
    assert np.array_equal(I, I_ref) and np.array_equal(D, D_ref)
    # the same handle with k > 16 afterwards (the base is re-cut into 8192-row segments, nothing is uploaded again)
    D2, I2 = index.search(x_q[:256], 40)
    D2_ref, I2_ref = O.knn(x_d, x_q[:256], 40, O.L2, O.F64, 0)
    assert index.last_path == "tensor-core" and np.array_equal(I2, I2_ref) and np.array_equal(D2, D2_ref)
    # real-valued base, k <= 16: approximate filter + exact re-rank
    xr_d, xr_q = synth(530_000, 24, 300, seed=10, integer=False, ncomp=100)
    ir = L.KnnIndex(xr_d, O.L2)
    Dr, Ir = ir.search(xr_q, 10)
    assert ir.last_path == "tensor-core"
    Dr_ref, Ir_ref = O.knn(xr_d, xr_q, 10, O.L2, O.F64, 0)
    assert_topk_equiv(Dr, Ir, Dr_ref, Ir_ref, xr_q, xr_d, O.L2)


def test_self_knn_with_exact_duplicates(L):
    """compute_knn.cpp:237-259: search k + 1, drop column 0 (self ASSUMED first). With exact duplicates the row itself
    is not always column 0: ties go to the lower id on both sides, so the dropped column and the kept ones agree."""
    x_d, _ = synth(66_000, 16, 1, seed=2, integer=True)
    x_d[1000:1400] = x_d[:400]          # 400 exact duplicates of earlier rows
    x_d[30_000:30_050] = x_d[29_000]    # a run of 50 identical rows + the original
    k = 10
    index = L.KnnIndex(x_d, O.L2)
    sel = np.r_[0:400, 1000:1400, 29_000:29_001, 30_000:30_050, 50_000:50_200]
    D, I = index.search(x_d[sel], k + 1)
    assert index.last_path == "tensor-core"
    D_ref, I_ref = O.knn(x_d, x_d[sel], k + 1, O.L2, O.F64, 0)
    assert np.array_equal(I, I_ref) and np.array_equal(D, D_ref)
    assert np.all(D[:, 0] == 0)
    assert np.array_equal(I[400:800, 0], np.arange(400))           # a duplicate's column 0 is its EARLIER twin, not itself
    assert np.array_equal(I[400:800, 1], np.arange(1000, 1400))    # ... and the row itself is what survives the drop
    knn = I[:, 1:k + 1]
    assert knn.shape == (len(sel), k)


@pytest.mark.parametrize("R", [2, 8])
@pytest.mark.parametrize("k,metric", [(10, O.L2), (100, O.L2), (10, O.IP)])
@pytest.mark.parametrize("dedup", [True, False])
def test_merge_ranks_dev_matches_numpy_merge(L, R, k, metric, dedup):
    """Per-rank sorted top-k lists with the same id on several ranks (the two copies of a redundantly stored vector),
    short lists (-1 padded) and a query with no result at all -> lira_pack_keys_dev + lira_merge_ranks_dev."""
    import torch
    from lira_ann_search_b200 import _cabi as C
    from lira_ann_search_b200.parallel import pack_keys
    rng = np.random.RandomState(R * 1000 + k)
    Q, n_ids = 500, 4 * k
    score_of = rng.randint(0, 50, (Q, n_ids)).astype(np.float32)   # many ties; the score is a function of (query, id)
    D_all = np.full((R, Q, k), np.inf, np.float32)
    I_all = np.full((R, Q, k), -1, np.int64)
    for r in range(R):
        for q in range(Q):
            n = 0 if q == 7 else rng.randint(0, k + 1)
            ids = rng.choice(n_ids, n, replace=False)
            sc = score_of[q, ids]
            order = np.lexsort((ids, sc))   # ascending (score, id): what the per-rank search emits
            I_all[r, q, :n] = ids[order]
            D_all[r, q, :n] = sc[order]
    sign = -1.0 if metric == O.IP else 1.0
    dev = torch.device("cuda:0")
    Dt = torch.as_tensor(sign * np.where(I_all >= 0, D_all, 0), device=dev).contiguous()
    It = torch.as_tensor(I_all, device=dev).contiguous()
    keys = pack_keys(Dt.reshape(R * Q, k), It.reshape(R * Q, k), metric, 0)
    D_out = torch.empty((Q, k), dtype=torch.float32, device=dev)
    I_out = torch.empty((Q, k), dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream or 1
    C.check(C.lib().lira_merge_ranks_dev(keys.data_ptr(), R, Q, k, metric, int(dedup), D_out.data_ptr(), I_out.data_ptr(), 0, st))
    torch.cuda.synchronize()
    if dedup:
        D_ref, I_ref = merge_ranks_numpy(D_all, I_all, k, dedup=True)
    else:   # search.cpp:499-513: the k best of the multiset first, equal ids collapsed afterwards (rows may come up short)
        D_ref = np.full((Q, k), np.inf, np.float32)
        I_ref = np.full((Q, k), -1, np.int64)
        for q in range(Q):
            cand = sorted((float(D_all[r, q, j]), int(I_all[r, q, j])) for r in range(R) for j in range(k) if I_all[r, q, j] >= 0)[:k]
            uniq = sorted(set(cand))
            for w, (dd, ii) in enumerate(uniq):
                D_ref[q, w], I_ref[q, w] = dd, ii
    got_I, got_D = I_out.cpu().numpy(), sign * D_out.cpu().numpy()
    assert np.array_equal(got_I, I_ref)
    assert np.array_equal(np.where(got_I >= 0, got_D, np.inf), D_ref)
    assert np.all(got_I[7] == -1)


def test_knn_byte_scan_far_seed_segments_overflow_is_redone(L):
    """Byte scan, exhaustive probe sets: the pooled seed comes from the first 16 base segments (131 072 rows). Here they
    hold a far-away cluster, so T[q] is loose, the 128-slot regions of the near segments overflow and the flagged queries
    go to the exact CUDA-core scan. The result must not change."""
    rng = np.random.RandomState(6)
    d, k = 32, 100
    far = np.clip(np.round(rng.randn(16 * 8192, d) * 4 + 200), 0, 255)
    near = np.clip(np.round(rng.randn(30_000, d) * 4 + 40), 0, 255)
    x_d = np.concatenate([far, near]).astype(np.float32)
    x_q = np.clip(np.round(rng.randn(300, d) * 4 + 40), 0, 255).astype(np.float32)
    index = L.KnnIndex(x_d, O.L2)
    D, I = index.search(x_q, k)
    assert index.last_scan_kind == "u8" and index.last_redo > 0
    D_ref, I_ref = O.knn(x_d, x_q, k, O.L2, O.F64, 0)
    assert np.array_equal(I, I_ref) and np.array_equal(D, D_ref)


@pytest.mark.parametrize("metric", [O.L2, O.IP])
def test_knn_fp16_scan_still_serves_byte_valued_data(L, metric, monkeypatch):
    """LIRA_NO_U8=1 at create time: the same data on the fp16 scan (kind::f16), identical rows."""
    monkeypatch.setenv("LIRA_NO_U8", "1")
    x_d, x_q = synth(70_000, 64, 300, seed=117, integer=True)
    index = L.KnnIndex(x_d, metric)
    D, I = index.search(x_q, 17)
    assert index.last_scan_kind == "fp16"
    D_ref, I_ref = O.knn(x_d, x_q, 17, metric, O.F64, 0)
    assert np.array_equal(I, I_ref) and np.array_equal(D, D_ref)
