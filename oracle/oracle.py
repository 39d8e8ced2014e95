"""TEST INFRASTRUCTURE -- Python face of the CPU oracle (oracle/lira_oracle.c + numpy).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product package (lira-ann-search_b200/) never does.

Pinning: see the header of lira_oracle.c. Host-side helpers below restate, in vectorised
numpy, the pure-Python loops of the reference and are themselves checked against the
reference's own functions (imported with the fake-faiss shim) in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liblira_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int)
_i64p = ctypes.POINTER(ctypes.c_int64)
_lp = ctypes.POINTER(ctypes.c_long)


def build(force: bool = False) -> str:
    """Compile oracle/lira_oracle.c -> oracle/_build/liblira_oracle.so (gcc, seconds)."""
    src = os.path.join(_HERE, "lira_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_recall.restype = ctypes.c_double
        _lib.oracle_select.restype = ctypes.c_long
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(ctypes.c_int(int(n)))


# ---------------------------------------------------------------------------------------------
# a1/a2 features
# ---------------------------------------------------------------------------------------------

def features_cpp(q, centroids, mean=None, scale=None):
    """search.cpp:220-250 twin (fp32 loop, sqrtf, (d-mean)/scale with scale==0 -> 1)."""
    q, c = _f32(q), _f32(centroids)
    Q, d = q.shape
    B = c.shape[0]
    out = np.empty((Q, B), np.float32)
    m = None if mean is None else _f32(mean)
    s = None if scale is None else _f32(scale)
    lib().oracle_features_cpp(_p(q, _f32p), ctypes.c_long(Q), _p(c, _f32p), ctypes.c_long(B),
                              ctypes.c_long(d), _p(m, _f32p), _p(s, _f32p), _p(out, _f32p))
    return out


def features_py(x, centroids, mean64=None, scale64=None):
    """utils.py:98-118 (cdist fp64 -> fp32) + StandardScaler.transform (utils.py:142-167)."""
    x, c = _f32(x), _f32(centroids)
    n, d = x.shape
    B = c.shape[0]
    out = np.empty((n, B), np.float32)
    m = None if mean64 is None else np.ascontiguousarray(mean64, np.float64)
    s = None if scale64 is None else np.ascontiguousarray(scale64, np.float64)
    lib().oracle_features_py(_p(x, _f32p), ctypes.c_long(n), _p(c, _f32p), ctypes.c_long(B),
                             ctypes.c_long(d), _p(m, _f64p), _p(s, _f64p), _p(out, _f32p))
    return out


def scaler_fit(dist_f32):
    """sklearn StandardScaler().fit on an fp32 matrix: mean_/var_ accumulated in fp64,
    scale_ = sqrt(var_) with zeros replaced by 1 (sklearn _handle_zeros_in_scale).
    Returns (mean64, scale64). utils.py:140-141."""
    x = np.asarray(dist_f32, dtype=np.float64)
    mean = x.mean(axis=0)
    var = x.var(axis=0)
    scale = np.sqrt(var)
    scale[scale < 10 * np.finfo(np.float64).eps] = 1.0
    return mean, scale


# ---------------------------------------------------------------------------------------------
# a3 probing model
# ---------------------------------------------------------------------------------------------

MLP_KEYS = ("distance_net.0", "distance_net.2", "vector_net.0", "vector_net.2", "fc.0", "fc.2")


def mlp_weights_from_state_dict(sd):
    """state_dict of model_probing.MLP_2_Input -> flat list [W1,b1,...,W6,b6] fp32 numpy."""
    out = []
    for k in MLP_KEYS:
        out.append(_f32(sd[k + ".weight"].detach().cpu().numpy() if hasattr(sd[k + ".weight"], "detach") else sd[k + ".weight"]))
        out.append(_f32(sd[k + ".bias"].detach().cpu().numpy() if hasattr(sd[k + ".bias"], "detach") else sd[k + ".bias"]))
    return out


def mlp_forward(x_dist, x_vec, weights):
    """model_probing.py:33-39 in fp64. Returns (logits64[n,B], probs64[n,B], hidden64[n,128])."""
    x_dist, x_vec = _f32(x_dist), _f32(x_vec)
    n, B = x_dist.shape
    d = x_vec.shape[1]
    W = [_f32(w) for w in weights]
    Bout = W[10].shape[0]
    logits = np.empty((n, Bout), np.float64)
    probs = np.empty((n, Bout), np.float64)
    hid = np.empty((n, 128), np.float64)
    lib().oracle_mlp_forward(_p(x_dist, _f32p), _p(x_vec, _f32p), ctypes.c_long(n), ctypes.c_long(B),
                             ctypes.c_long(d), ctypes.c_long(Bout), *[_p(w, _f32p) for w in W],
                             _p(logits, _f64p), _p(probs, _f64p), _p(hid, _f64p))
    return logits, probs, hid


# ---------------------------------------------------------------------------------------------
# a5 selection
# ---------------------------------------------------------------------------------------------

SELECT_GT, SELECT_GE_ARGMAX, SELECT_TOPN = 0, 1, 2


def select(scores, mode, value):
    """Returns CSR (probe_offsets int64[Q+1], probe_ids int32[P])."""
    s = _f32(scores)
    Q, B = s.shape
    off = np.empty(Q + 1, np.int64)
    ids = np.empty(Q * B, np.int32)
    n = lib().oracle_select(_p(s, _f32p), ctypes.c_long(Q), ctypes.c_long(B), ctypes.c_int(mode),
                            ctypes.c_double(float(value)), _p(off, _lp), _p(ids, _i32p))
    return off, ids[:n].copy()


# ---------------------------------------------------------------------------------------------
# a6 inverted lists (CSR)
# ---------------------------------------------------------------------------------------------

def build_lists_from_cluster_ids(x_d, cluster_ids):
    """utils.py:407-422: list b holds x_d[cluster_ids[b]] in that order (local idx i <->
    cluster_ids[b][i]). Returns (offsets int64[B+1], ids int32[E], vecs f32[E,d])."""
    B = len(cluster_ids)
    sizes = np.array([len(c) for c in cluster_ids], np.int64)
    off = np.zeros(B + 1, np.int64)
    np.cumsum(sizes, out=off[1:])
    ids = np.empty(off[-1], np.int32)
    for b, c in enumerate(cluster_ids):
        ids[off[b]:off[b + 1]] = np.asarray(c, np.int64)
    vecs = _f32(np.asarray(x_d)[ids]) if len(ids) else np.empty((0, np.asarray(x_d).shape[1]), np.float32)
    return off, ids, vecs


def build_lists_from_data_2_bkt(x_d, data_2_bkt, n_bkt):
    """search.cpp:368-403: every (point, column) with bucket >= 0 joins that bucket, then the
    bucket's ids are sorted and made unique."""
    d2b = np.asarray(data_2_bkt)
    n, n_mul = d2b.shape
    pts = np.repeat(np.arange(n, dtype=np.int64), n_mul)
    bk = d2b.reshape(-1).astype(np.int64)
    keep = bk >= 0
    if np.any(bk[keep] >= n_bkt):
        raise RuntimeError("bucket id out of range.")  # search.cpp:375-377
    pairs = np.unique(np.stack([bk[keep], pts[keep]], 1), axis=0)  # sorted by (bucket, id), unique
    sizes = np.bincount(pairs[:, 0], minlength=n_bkt)
    off = np.zeros(n_bkt + 1, np.int64)
    np.cumsum(sizes, out=off[1:])
    ids = pairs[:, 1].astype(np.int32)
    vecs = _f32(np.asarray(x_d)[ids])
    return off, ids, vecs


# ---------------------------------------------------------------------------------------------
# a7 / a10 scans
# ---------------------------------------------------------------------------------------------

L2, IP = 0, 1
F32, F64 = 0, 1


def list_search(list_vecs, q, k, metric=L2, prec=F64):
    """Faiss IndexFlat{L2,IP}.search restated. Returns (D f32[nq,k], I i64[nq,k])."""
    v, q = _f32(list_vecs), _f32(q)
    nq, d = q.shape
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    lib().oracle_list_search(_p(v, _f32p), ctypes.c_long(v.shape[0]), ctypes.c_long(d), _p(q, _f32p),
                             ctypes.c_long(nq), ctypes.c_int(k), ctypes.c_int(metric),
                             ctypes.c_int(prec), _p(D, _f32p), _p(I, _i64p))
    return D, I


def scan_all_pairs(off, ids, vecs, q, k, metric=L2, prec=F64):
    """get_cmp_recall (LIRA_smallscale.py:145-174). Returns (found i64[Q,B,k], cmp i64[Q,B])."""
    off = np.ascontiguousarray(off, np.int64)
    ids = np.ascontiguousarray(ids, np.int32)
    vecs, q = _f32(vecs), _f32(q)
    Q, d = q.shape
    B = len(off) - 1
    found = np.empty((Q, B, k), np.int64)
    cmp_ = np.empty((Q, B), np.int64)
    lib().oracle_scan_all_pairs(_p(off, _lp), _p(ids, _i32p), _p(vecs, _f32p), ctypes.c_long(B),
                                ctypes.c_long(d), _p(q, _f32p), ctypes.c_long(Q), ctypes.c_int(k),
                                ctypes.c_int(metric), ctypes.c_int(prec), _p(found, _i64p), _p(cmp_, _i64p))
    return found, cmp_


def search(off, ids, vecs, q, probe_off, probe_ids, k, metric=L2, prec=F64, dedup=1):
    """search.cpp:468-514. Returns (ids i64[Q,k], dist f32[Q,k], cmp i64[Q])."""
    off = np.ascontiguousarray(off, np.int64)
    ids = np.ascontiguousarray(ids, np.int32)
    vecs, q = _f32(vecs), _f32(q)
    probe_off = np.ascontiguousarray(probe_off, np.int64)
    probe_ids = np.ascontiguousarray(probe_ids, np.int32)
    Q, d = q.shape
    B = len(off) - 1
    out_ids = np.empty((Q, k), np.int64)
    out_d = np.empty((Q, k), np.float32)
    out_cmp = np.empty(Q, np.int64)
    lib().oracle_search(_p(off, _lp), _p(ids, _i32p), _p(vecs, _f32p), ctypes.c_long(B), ctypes.c_long(d),
                        _p(q, _f32p), ctypes.c_long(Q), _p(probe_off, _lp), _p(probe_ids, _i32p),
                        ctypes.c_int(k), ctypes.c_int(metric), ctypes.c_int(prec), ctypes.c_int(dedup),
                        _p(out_ids, _i64p), _p(out_d, _f32p), _p(out_cmp, _i64p))
    return out_ids, out_d, out_cmp


def recall(ids, gt, k):
    """search.cpp:520-528."""
    ids = np.ascontiguousarray(ids, np.int64)
    gt = np.ascontiguousarray(gt, np.int32)
    return float(lib().oracle_recall(_p(ids, _i64p), ctypes.c_long(ids.shape[0]), ctypes.c_int(k),
                                     _p(gt, _i32p), ctypes.c_long(gt.shape[1])))


def knn(base, query, k, metric=L2, prec=F64, form=0):
    """compute_knn.cpp:208-259 / utils.py:293-310 exact search. Returns (D, I)."""
    base, query = _f32(base), _f32(query)
    N, d = base.shape
    Q = query.shape[0]
    D = np.empty((Q, k), np.float32)
    I = np.empty((Q, k), np.int64)
    lib().oracle_knn(_p(base, _f32p), ctypes.c_long(N), _p(query, _f32p), ctypes.c_long(Q), ctypes.c_long(d),
                     ctypes.c_int(k), ctypes.c_int(metric), ctypes.c_int(prec), ctypes.c_int(form),
                     _p(D, _f32p), _p(I, _i64p))
    return D, I


# ---------------------------------------------------------------------------------------------
# a8 / a9 host-side recall helpers (vectorised restatements)
# ---------------------------------------------------------------------------------------------

def knn_distr_redundancy(knn_ids, data_2_bkt, n_bkt):
    """utils.py:354-379 get_knn_distr_redundancy, as a boolean membership tensor instead of
    Q*B python lists: member[q, j, c] = bucket of the c-th copy of the j-th GT id (-1 = none).
    cnt[q,b] equals the reference's knn_distr_cnt; set(knn_distr_id[q][b]) equals
    {knn[q,j] : any_c member[q,j,c]==b}."""
    knn_ids = np.asarray(knn_ids)
    d2b = np.asarray(data_2_bkt)
    member = d2b[knn_ids]  # [Q,k,n_mul]
    Q = knn_ids.shape[0]
    cnt = np.zeros((Q, n_bkt), np.int64)
    flat = member.reshape(Q, -1)
    rows = np.repeat(np.arange(Q), flat.shape[1])
    ok = flat.reshape(-1) >= 0
    np.add.at(cnt, (rows[ok], flat.reshape(-1)[ok]), 1)
    return cnt, member


def query_tuning_rows(all_outputs, knn_ids, member, found_aknn_id, cmp_distr_all, k, thresholds):
    """query_tuning (LIRA_smallscale.py:199-220) without the timing column.
    Per threshold: mean nprobe, mean recall = |U_{b probed} set(knn_distr_id[q][b]) n
    set(found[q][b])| / k, mean computations. Returns list of (thr, nprobe, recall, cmp)."""
    all_outputs = np.asarray(all_outputs)
    Q, B = all_outputs.shape
    knn_ids = np.asarray(knn_ids)
    rows = []
    for thr in thresholds:
        probed = all_outputs > np.float32(thr)  # torch fp32 tensor vs scalar: compared in fp32 (LIRA_smallscale.py:206)
        nprobe = probed.sum(1)
        cmp_ = (cmp_distr_all * probed).sum(1)
        rec = np.zeros(Q)
        for q in range(Q):
            got = set()
            for b in np.nonzero(probed[q])[0]:
                in_b = knn_ids[q][(member[q] == b).any(axis=1)]
                if in_b.size:
                    got.update(set(in_b.tolist()).intersection(found_aknn_id[q, b].tolist()))
            rec[q] = len(got) / k
        rows.append((float(thr), float(nprobe.mean()), float(rec.mean()), float(cmp_.mean())))
    return rows
