"""The query-phase functions the reference defines inside its drivers, with the same call shapes:
get_cmp_recall (LIRA_smallscale.py:145-174, LIRA_largescale.py:120-149), query_tuning
(LIRA_smallscale.py:176-241, LIRA_largescale.py:151-179), plus the search.cpp main loop
(search.cpp:413-549) as `search_sweep`.
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import _cabi as C
from .engine import ListView, LiraIndex
from .utils import KnnDistrIds, fprint


def get_cmp_recall(inner_indexes, x_q, xd_id_bkt, cfg):
    """Search every partition for every query.
    -> search_time [n_q, n_bkt] f64, cmp_distr_all [n_q, n_bkt] int, found_aknn_id [n_q, n_bkt, k] int (-1 init).

    `inner_indexes` is what create_inner_indexes returned (per-bucket views of one device index): the
    whole n_q x n_bkt sweep is ONE grouped scan on the GPU instead of n_q*n_bkt faiss calls.
    A per-pair wall time has no meaning for a batched kernel; search_time[q, b] is the batch time
    apportioned by list size (t_batch * n_b / sum_b n_b / n_q), so that query_tuning's QPS column,
    1 / mean_q sum_{b probed} search_time[q, b], is the batch throughput restricted to the probed lists."""
    n_bkt, k = cfg.n_bkt, cfg.k
    x_q = np.ascontiguousarray(x_q, np.float32)
    n_q = len(x_q)
    views = [v for v in inner_indexes if isinstance(v, ListView)]
    if len(views) != n_bkt or any(v.index is not views[0].index for v in views):
        raise ValueError("inner_indexes must come from create_inner_indexes / create_flat_indexes")
    index: LiraIndex = views[0].index
    t0 = time.time()
    found, cmp_ = index.scan_all_pairs(x_q, k)
    elapsed = time.time() - t0
    sizes = index.list_sizes().astype(np.float64)
    total = max(sizes.sum(), 1.0)
    search_time = np.broadcast_to(elapsed * sizes / total / max(n_q, 1), (n_q, n_bkt)).copy()
    # reference dtype: np.zeros(..., dtype=int) / np.full(..., -1) -> int64
    return search_time, cmp_.astype(int), found.astype(int)


def _found_mask(knn_distr_id, found_aknn_id, knn_query=None):
    """in_found[q, j, c]: ground-truth id knn[q, j] was returned by the search of the bucket holding
    its c-th copy. Equivalent to set(knn_distr_id[q][b]) & set(found_aknn_id[q][b]) summed over b."""
    if not isinstance(knn_distr_id, KnnDistrIds):
        raise TypeError("knn_distr_id must come from utils.get_knn_distr_redundancy")
    knn, member = knn_distr_id.knn, knn_distr_id.member
    Q, k, n_mul = member.shape
    qq = np.arange(Q)[:, None, None]
    safe = np.where(member >= 0, member, 0)
    got = found_aknn_id[qq, safe]  # [Q, k, n_mul, kfound]
    return (got == knn[:, :, None, None]).any(-1) & (member >= 0)


def _tuning_rows(all_outputs, knn_distr_id, found_aknn_id, search_time, cmp_distr_all, k, thresholds):
    all_outputs = np.asarray(all_outputs)
    member = knn_distr_id.member
    in_found = _found_mask(knn_distr_id, np.asarray(found_aknn_id))
    safe = np.where(member >= 0, member, 0)
    qq = np.arange(all_outputs.shape[0])[:, None, None]
    rows = []
    for thr in thresholds:
        # LIRA_smallscale.py:206: all_outputs is a torch fp32 tensor there, so the comparison runs in fp32
        probed = all_outputs > np.float32(thr)
        nprobe = probed.sum(1)
        cmp_ = (np.asarray(cmp_distr_all) * probed).sum(1)
        hit = (in_found & probed[qq, safe]).any(-1)  # [Q, k]: id found in >= 1 probed bucket
        # set semantics (LIRA_smallscale.py:210-214): a repeated ground-truth id counts once
        knn = knn_distr_id.knn
        if (np.sort(knn, 1)[:, 1:] == np.sort(knn, 1)[:, :-1]).any():
            rec = np.array([len(set(knn[q][hit[q]].tolist())) for q in range(len(knn))]) / k
        else:
            rec = hit.sum(1) / k
        t = (np.asarray(search_time) * probed).sum(1) if search_time is not None else None
        rows.append((thr, nprobe.mean(), rec.mean(), cmp_.mean(), None if t is None else t.mean()))
    return rows


def query_tuning(all_outputs, knn_distr_id, found_aknn_id, search_time, cmp_distr_all, cfg, fw, part=0,
                 thresholds=None):
    """Threshold sweep of the small-scale driver (LIRA_smallscale.py:176-241): writes
    {pth_log}{file_name}_tuning_threshold/{duplicate_type}_{part}.csv with columns
    threshold,nprobe,Recall,Computations,QPS and returns the DataFrame."""
    import pandas as pd
    thresholds = np.arange(0.02, 0.82, 0.02) if thresholds is None else thresholds
    os.makedirs(cfg.pth_log + cfg.file_name + "_tuning_threshold/", exist_ok=True)
    fprint("", fw)
    fprint("=" * 90, fw)
    fprint(f"Query Tuning Results - Part {part}", fw)
    fprint(f"Dataset: {cfg.dataset}, n_bkt: {cfg.n_bkt}, metric: {cfg.dis_metric}, "
           f"redundancy_ratio: {cfg.redundancy_ratio}", fw)
    fprint(f"Number of queries: {len(all_outputs)}", fw)
    fprint("=" * 90, fw)
    out = []
    for thr, nprobe, rec, cmp_, t in _tuning_rows(all_outputs, knn_distr_id, found_aknn_id, search_time,
                                                  cmp_distr_all, cfg.k, thresholds):
        qps = 1.0 / t if t and t > 0 else 0.0
        fprint(f"threshold: {thr:.3f}, nprobe: {nprobe:.2f}, Recall: {rec:.4f}, Computations: {cmp_:.0f}, "
               f"QPS: {qps:.2f}", fw)
        out.append({"threshold": thr, "nprobe": nprobe, "Recall": rec, "Computations": cmp_, "QPS": qps})
    df = pd.DataFrame(out, columns=["threshold", "nprobe", "Recall", "Computations", "QPS"])
    csv_path = cfg.pth_log + cfg.file_name + f"_tuning_threshold/{cfg.duplicate_type}_{part}.csv"
    df.to_csv(csv_path, index=False)
    fprint(f">> Query tuning CSV saved to: {csv_path}", fw)
    return df


def query_tuning_large(all_outputs, knn_distr_id, found_aknn_id, cmp_distr_all, cfg, part=0, thresholds=None):
    """Large-scale variant (LIRA_largescale.py:151-179): thresholds arange(0.1, 1.0, 0.02), no QPS column.
    (The reference reads cmp_distr_all from a module global; it is an argument here.)"""
    import pandas as pd
    thresholds = np.arange(0.1, 1.0, 0.02) if thresholds is None else thresholds
    out = []
    for thr, nprobe, rec, cmp_, _ in _tuning_rows(all_outputs, knn_distr_id, found_aknn_id, None, cmp_distr_all,
                                                  cfg.k, thresholds):
        print(f"threshold: {thr:.3f}, nprobe: {nprobe}, KNN Recall: {rec:.4f}, KNN Computations: {cmp_:.4f}")
        out.append({"threshold": thr, "nprobe": nprobe, "Recall": rec, "Computations": cmp_})
    df = pd.DataFrame(out, columns=["threshold", "nprobe", "Recall", "Computations"])
    os.makedirs(cfg.pth_log + cfg.file_name + "_tuning_threshold/", exist_ok=True)
    df.to_csv(cfg.pth_log + cfg.file_name + f"_tuning_threshold/{cfg.duplicate_type}_{part}.csv", index=False)
    return df


def recall_at_k(ids, gt, k):
    """search.cpp:520-528: mean over queries of |gt[q,:k] & result[q]| / k."""
    ids, gt = np.asarray(ids)[:, :k], np.asarray(gt)[:, :k]
    hit = (gt[:, :, None] == ids[:, None, :]).any(-1)
    return float(hit.sum(1).mean() / k)


def cpp_thresholds(t_min=0.02, t_max=0.80, t_step=0.02):
    """search.cpp:413: `for (float thr = t_min; thr <= t_max + 1e-6f; thr += t_step)` in fp32."""
    out, thr = [], np.float32(t_min)
    while thr <= np.float32(t_max) + np.float32(1e-6):
        out.append(float(thr))
        thr = np.float32(thr + np.float32(t_step))
    return out


def search_sweep(index: LiraIndex, model, x_q, gt_ids, k, thresholds=None, mode=C.SELECT_GE_ARGMAX, dedup=True,
                 out=print):
    """search.cpp:413-549: per threshold, run the whole query phase over all queries and report
    Threshold / avg_recall / avg_nprobe / avg_cmp / avg_time(q) / QPS. Returns the rows as dicts.
    dedup=False reproduces the reference binary's select-then-collapse behaviour exactly."""
    thresholds = cpp_thresholds() if thresholds is None else thresholds
    x_q = np.ascontiguousarray(x_q, np.float32)
    rows = []
    for thr in thresholds:
        t0 = time.perf_counter()
        _, ids, nprobe, cmp_ = index.probe_search(model, x_q, mode, thr, k, dedup)
        dt = time.perf_counter() - t0
        row = {"Threshold": thr, "avg_recall": recall_at_k(ids, gt_ids, k), "avg_nprobe": float(nprobe.mean()),
               "avg_cmp": float(cmp_.mean()), "avg_time(q)": dt / len(x_q), "QPS": len(x_q) / dt}
        rows.append(row)
        if out:
            out(f"=== Threshold = {thr:g} ===")
            out(f"Threshold    : {thr:g}")
            out(f"avg_recall   : {row['avg_recall']:g}")
            out(f"avg_nprobe   : {row['avg_nprobe']:g}")
            out(f"avg_cmp      : {row['avg_cmp']:g}")
            out(f"avg_time(q)  : {row['avg_time(q)']:g} s")
            out(f"QPS          : {row['QPS']:g} q/s")
            out("----------------------------------------")
    return rows


# ---------------------------------------------------------------------------------------------
# redundancy assignment -- LIRA_smallscale.py:77-97 / LIRA_largescale.py:51-72
# ---------------------------------------------------------------------------------------------
def _as_numpy(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def _mul_partition_rows(score_rows, pred_rows, t, data_2_bkt, cluster_cnts, cluster_ids):
    """Core of both drivers' redundancy assignment, vectorised: score_rows[i] / pred_rows[i] belong to point t[i]."""
    if t.size == 0:
        return
    n_mul = data_2_bkt.shape[1]
    top = np.argsort(-score_rows, axis=1, kind="stable")[:, :n_mul]       # the n_mul best partitions of every point
    width = top.shape[1]
    n_eff = (pred_rows != 0).sum(1)
    n_act = np.minimum(n_mul - 1, n_eff)
    cur = np.asarray(data_2_bkt[t, 0], np.int64)
    hit = top == cur[:, None]
    loc = np.where(hit.any(1), hit.argmax(1), width)                        # beyond the best n_mul: certainly >= n_actual
    keep_cur = loc >= n_act
    n_take = np.where(keep_cur | (n_eff == n_act), n_act, n_act + 1)
    first_col = np.where(keep_cur, 1, 0)
    j = np.arange(width)[None, :]
    chosen = j < n_take[:, None]                                            # [points, rank]: this rank is written
    rows, ranks = np.nonzero(chosen)
    data_2_bkt[t[rows], first_col[rows] + ranks] = top[rows, ranks]
    new = chosen & (top != cur[:, None])                                    # appended to a partition other than the current
    rows, ranks = np.nonzero(new)                                           # row-major: points in order, ranks ascending
    part, pts = top[rows, ranks], t[rows]
    np.add.at(cluster_cnts, part, 1)
    by_part = np.argsort(part, kind="stable")
    part_s, pts_s = part[by_part], pts[by_part]
    cuts = np.flatnonzero(np.diff(part_s)) + 1
    for c, ids in zip(part_s[np.r_[0, cuts]] if part_s.size else [], np.split(pts_s, cuts) if part_s.size else []):
        cluster_ids[int(c)].extend(int(x) for x in ids)


def mul_partition_by_model(data_partition_score, data_predicts, xd_id_sorted_pre, data_2_bkt, cluster_cnts, cluster_ids,
                           begin, end):
    """The reference's per-point Python loop, vectorised (same arguments, same in-place effects on `data_2_bkt`,
    `cluster_cnts` and `cluster_ids`, same append order inside every `cluster_ids[c]`).

    For each point t of xd_id_sorted_pre[begin:end] (LIRA_smallscale.py:79-97): partitions ranked by score, descending;
    n_actual = min(n_mul - 1, #partitions predicted for t); with `loc` the rank of t's current partition:
      loc >= n_actual            -> the n_actual best go to columns 1..n_actual             (current partition kept in column 0)
      else, n_eff == n_actual    -> the n_actual best replace columns 0..n_actual-1
      else                       -> the n_actual + 1 best replace columns 0..n_actual
    and t is appended to every chosen partition other than its current one. Equal scores are ranked by partition id
    (torch.argsort leaves their order unspecified)."""
    score, pred, order = _as_numpy(data_partition_score), _as_numpy(data_predicts), _as_numpy(xd_id_sorted_pre)
    t = np.asarray(order[begin:end], np.int64)
    _mul_partition_rows(score[t], pred[t], t, data_2_bkt, cluster_cnts, cluster_ids)


def _append_new_copies(added, t, cluster_cnts, cluster_ids):
    """added[i, j] = partition newly holding point t[i] in column j (-1: none): the reference's bookkeeping -- counts, and the
    point appended to every such partition, points in order, columns ascending (LIRA_smallscale.py:94-97)."""
    rows, cols = np.nonzero(added >= 0)                                     # row-major: points in order, columns ascending
    part, pts = added[rows, cols].astype(np.int64), t[rows]
    np.add.at(cluster_cnts, part, 1)
    by_part = np.argsort(part, kind="stable")
    part_s, pts_s = part[by_part], pts[by_part]
    cuts = np.flatnonzero(np.diff(part_s)) + 1
    for c, ids in zip(part_s[np.r_[0, cuts]] if part_s.size else [], np.split(pts_s, cuts) if part_s.size else []):
        cluster_ids[int(c)].extend(int(x) for x in ids)


def _mul_partition_device(score, t, data_2_bkt, cluster_cnts, cluster_ids):
    """The same rule evaluated by the library on the device (lira_mul_partition_dev: one warp per point) for scores that are
    already there: only the [rows, n_mul] partition columns travel."""
    import torch
    from . import engine
    dev = score.device
    rows = torch.as_tensor(np.ascontiguousarray(data_2_bkt[t], np.int32), device=dev)
    added = torch.full_like(rows, -1)
    engine.mul_partition_dev(score.float().contiguous(), rows, added)
    torch.cuda.synchronize(dev)
    data_2_bkt[t] = rows.cpu().numpy()
    _append_new_copies(added.cpu().numpy(), t, cluster_cnts, cluster_ids)


def mul_partition_by_model_large(data_partition_score, data_predicts, global_xd_ids, start_idx, data_2_bkt, cluster_cnts,
                                 cluster_ids):
    """LIRA_largescale.py:51-72: the same rule for one batch of the full data -- row i of the score / predict matrices is
    point global_xd_ids[i] = start_idx + i. Vectorised, same in-place effects and append order as the reference loop.
    A CUDA score tensor takes the device form of the rule (data_predicts is then not needed: it is score > 0.5)."""
    t = np.asarray(_as_numpy(global_xd_ids), np.int64)
    if getattr(data_partition_score, "is_cuda", False):
        local = t - int(start_idx)
        sc = data_partition_score if np.array_equal(local, np.arange(len(t))) else data_partition_score[local.tolist()]
        return _mul_partition_device(sc, t, data_2_bkt, cluster_cnts, cluster_ids)
    score, pred = _as_numpy(data_partition_score), _as_numpy(data_predicts)
    local = t - int(start_idx)
    _mul_partition_rows(score[local], pred[local], t, data_2_bkt, cluster_cnts, cluster_ids)
