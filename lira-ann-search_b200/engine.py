"""Device-resident LIRA index and probing model (thin objects over the C ABI).

LiraIndex  <-> utils.create_inner_indexes / search.cpp:368-403 (the inverted lists)
ListView   <-> one faiss.IndexFlat{L2,IP} of the reference's `inner_indexes` list
LiraModel  <-> MLP_2_Input + kmeans.centroids + StandardScaler params (search.cpp:302-338)
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi as C


def _stream_handle(device, stream=None) -> int:
    """cudaStream_t to hand to a *_dev entry point: torch's current stream on `device`. The legacy default
    stream is passed as cudaStreamLegacy (0x1) because NULL means "the handle's own stream" in the C ABI."""
    import torch
    st = stream if stream is not None else torch.cuda.current_stream(device).cuda_stream
    return int(st) if int(st) != 0 else 1


def _metric_code(metric) -> int:
    if isinstance(metric, str):
        return C.METRIC_IP if metric.lower() in ("inner_product", "ip", "dot", "dot_product") else C.METRIC_L2
    return int(metric)


class LiraIndex:
    """All B inverted lists of one index in HBM (CSR: offsets, int32 ids, fp32 vectors)."""

    def __init__(self, handle, n_bkt, dim, metric, device):
        self._h = handle
        self.n_bkt, self.dim, self.metric, self.device = n_bkt, dim, metric, device

    # ---- construction ---------------------------------------------------------------------
    @classmethod
    def from_cluster_ids(cls, x_d, cluster_ids, metric="L2", device=0):
        """utils.py:407-422: list b = x_d[cluster_ids[b]] in that order."""
        C.require_gpu()
        x_d = C.f32(x_d)
        B = len(cluster_ids)
        sizes = np.fromiter((len(c) for c in cluster_ids), np.int64, B)
        off = np.zeros(B + 1, np.int64)
        np.cumsum(sizes, out=off[1:])
        ids = np.empty(max(int(off[-1]), 1), np.int32)
        for b, c in enumerate(cluster_ids):
            if len(c):
                ids[off[b]:off[b + 1]] = np.asarray(c, np.int64)
        return cls.from_csr(x_d, off, ids, metric, device)

    @classmethod
    def from_csr(cls, x_d, list_offsets, list_ids, metric="L2", device=0):
        C.require_gpu()
        x_d = C.f32(x_d)
        off = np.ascontiguousarray(list_offsets, np.int64)
        ids = np.ascontiguousarray(list_ids, np.int32)
        h = ctypes.c_void_p()
        m = _metric_code(metric)
        C.check(C.lib().lira_index_create(C.ptr(x_d, C.c_f32p), x_d.shape[0], x_d.shape[1], C.ptr(off, C.c_i64p),
                                          C.ptr(ids, C.c_i32p), len(off) - 1, m, device, ctypes.byref(h)))
        return cls(h, len(off) - 1, x_d.shape[1], m, device)

    @classmethod
    def from_data_2_bkt(cls, x_d, data_2_bkt, n_bkt, metric="L2", device=0):
        """search.cpp:368-403: buckets from the (N, n_mul) assignment matrix, -1 = empty slot."""
        C.require_gpu()
        x_d = C.f32(x_d)
        d2b = np.ascontiguousarray(data_2_bkt, np.int32)
        h = ctypes.c_void_p()
        m = _metric_code(metric)
        C.check(C.lib().lira_index_create_from_assign(C.ptr(x_d, C.c_f32p), x_d.shape[0], x_d.shape[1],
                                                      C.ptr(d2b, C.c_i32p), d2b.shape[1], n_bkt, m, device,
                                                      ctypes.byref(h)))
        return cls(h, n_bkt, x_d.shape[1], m, device)

    @classmethod
    def from_device(cls, d_vecs, d_ids, list_offsets, dim, metric="L2", device=0):
        """Adopt torch CUDA tensors (vecs [E, ld] fp32, ids [E] int32) without copying."""
        C.require_gpu()
        off = np.ascontiguousarray(list_offsets, np.int64)
        h = ctypes.c_void_p()
        m = _metric_code(metric)
        C.check(C.lib().lira_index_create_dev(d_vecs.data_ptr(), d_vecs.stride(0), dim, C.ptr(off, C.c_i64p),
                                              d_ids.data_ptr(), len(off) - 1, m, device, ctypes.byref(h)))
        obj = cls(h, len(off) - 1, dim, m, device)
        obj._keep = (d_vecs, d_ids)
        return obj

    def close(self):
        if self._h is not None:
            C.lib().lira_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- faiss-like surface ---------------------------------------------------------------
    def ntotal(self, b=-1) -> int:
        return int(C.lib().lira_index_ntotal(self._h, int(b)))

    def list_sizes(self):
        return np.array([self.ntotal(b) for b in range(self.n_bkt)], np.int64)

    def views(self):
        """The reference's `inner_indexes`: one object per bucket with .search / .ntotal."""
        return [ListView(self, b) for b in range(self.n_bkt)]

    def list_search(self, b, q, k):
        q = C.f32(q).reshape(-1, self.dim)
        D = np.empty((q.shape[0], k), np.float32)
        I = np.empty((q.shape[0], k), np.int64)
        C.check(C.lib().lira_index_list_search(self._h, int(b), C.ptr(q, C.c_f32p), q.shape[0], int(k),
                                               C.ptr(D, C.c_f32p), C.ptr(I, C.c_i64p)))
        return D, I

    def scan_all_pairs(self, q, k):
        """get_cmp_recall arithmetic: (found[Q,B,k] int64, cmp[Q,B] int64)."""
        q = C.f32(q).reshape(-1, self.dim)
        Q = q.shape[0]
        found = np.empty((Q, self.n_bkt, k), np.int64)
        cmp_ = np.empty((Q, self.n_bkt), np.int64)
        C.check(C.lib().lira_scan_all_pairs(self._h, C.ptr(q, C.c_f32p), Q, int(k), C.ptr(found, C.c_i64p),
                                            C.ptr(cmp_, C.c_i64p)))
        return found, cmp_

    def search(self, q, probe_offsets, probe_ids, k, dedup=True):
        """search.cpp:468-514 with explicit probe sets. Returns (D, I, cmp)."""
        q = C.f32(q).reshape(-1, self.dim)
        Q = q.shape[0]
        po = np.ascontiguousarray(probe_offsets, np.int64)
        pi = np.ascontiguousarray(probe_ids, np.int32)
        D = np.empty((Q, k), np.float32)
        I = np.empty((Q, k), np.int64)
        cmp_ = np.empty(Q, np.int64)
        C.check(C.lib().lira_search(self._h, C.ptr(q, C.c_f32p), Q, C.ptr(po, C.c_i64p), C.ptr(pi, C.c_i32p), int(k),
                                    int(bool(dedup)), C.ptr(D, C.c_f32p), C.ptr(I, C.c_i64p), C.ptr(cmp_, C.c_i64p)))
        return D, I, cmp_

    def probe_search(self, model, q, mode, value, k, dedup=True, out=None):
        """The whole query phase for a batch of host queries. Returns (D, I, nprobe, cmp); `out` = such a tuple of
        C-contiguous arrays from an earlier call is filled in place (no allocation per batch)."""
        q = C.f32(q).reshape(-1, self.dim)
        Q = q.shape[0]
        if out is not None:
            D, I, npb, cmp_ = out
            if not (D.shape == (Q, k) and D.dtype == np.float32 and I.shape == (Q, k) and I.dtype == np.int64
                    and npb.shape == (Q,) and npb.dtype == np.int32 and cmp_.shape == (Q,) and cmp_.dtype == np.int64
                    and all(a.flags.c_contiguous for a in out)):
                raise ValueError("out: expected (float32[Q,k], int64[Q,k], int32[Q], int64[Q]) C-contiguous arrays")
        else:
            D = np.empty((Q, k), np.float32)
            I = np.empty((Q, k), np.int64)
            npb = np.empty(Q, np.int32)
            cmp_ = np.empty(Q, np.int64)
        C.check(C.lib().lira_probe_search(self._h, model._h, C.ptr(q, C.c_f32p), Q, int(mode), float(value), int(k),
                                          int(bool(dedup)), C.ptr(D, C.c_f32p), C.ptr(I, C.c_i64p),
                                          C.ptr(npb, C.c_i32p), C.ptr(cmp_, C.c_i64p)))
        return D, I, npb, cmp_

    # ---- device-resident entry points (torch CUDA tensors in, torch CUDA tensors out) --------
    def probe_search_dev(self, model, d_q, mode, value, k, dedup=True, out=None, stream=None):
        import torch
        Q = d_q.shape[0]
        if out is None:
            out = (torch.empty((Q, k), dtype=torch.float32, device=d_q.device),
                   torch.empty((Q, k), dtype=torch.int64, device=d_q.device),
                   torch.empty((Q,), dtype=torch.int32, device=d_q.device),
                   torch.empty((Q,), dtype=torch.int64, device=d_q.device))
        D, I, npb, cmp_ = out
        st = _stream_handle(d_q.device, stream)
        C.check(C.lib().lira_probe_search_dev(self._h, model._h, d_q.data_ptr(), d_q.stride(0), Q, int(mode),
                                              float(value), int(k), int(bool(dedup)), D.data_ptr(), I.data_ptr(),
                                              npb.data_ptr(), cmp_.data_ptr(), st))
        return out

    def probe_search_enqueue_dev(self, model, d_q, mode, value, k, dedup=True, out=None, stream=None):
        """probe_search_dev without the host wait: the batch is launched and the call returns. The outputs are valid (in stream
        order) once `finish()` has been called -- it checks the status words of every enqueued batch and answers the rare
        batch whose optimistic run was void again."""
        import torch
        Q = d_q.shape[0]
        if out is None:
            out = (torch.empty((Q, k), dtype=torch.float32, device=d_q.device),
                   torch.empty((Q, k), dtype=torch.int64, device=d_q.device),
                   torch.empty((Q,), dtype=torch.int32, device=d_q.device),
                   torch.empty((Q,), dtype=torch.int64, device=d_q.device))
        D, I, npb, cmp_ = out
        st = _stream_handle(d_q.device, stream)
        self._inflight = getattr(self, "_inflight", [])
        self._inflight.append((d_q, out))   # the tensors must outlive the batch
        C.check(C.lib().lira_probe_search_enqueue_dev(self._h, model._h, d_q.data_ptr(), d_q.stride(0), Q, int(mode),
                                                      float(value), int(k), int(bool(dedup)), D.data_ptr(), I.data_ptr(),
                                                      npb.data_ptr(), cmp_.data_ptr(), st))
        return out

    def finish(self):
        C.check(C.lib().lira_index_finish(self._h))
        self._inflight = []

    def probe_search_submit(self, model, q, mode, value, k, dedup=True, slot=0):
        """Host queries in, asynchronously: returns at once; `probe_search_wait(slot)` delivers the results."""
        q = C.f32(q).reshape(-1, self.dim)
        self._slots = getattr(self, "_slots", {})
        self._slots[slot] = (q, q.shape[0], int(k))   # the array must stay alive until wait()
        C.check(C.lib().lira_probe_search_submit(self._h, model._h, C.ptr(q, C.c_f32p), q.shape[0], int(mode), float(value),
                                                 int(k), int(bool(dedup)), int(slot)))

    def probe_search_wait(self, slot=0, out=None):
        _, Q, k = self._slots.pop(slot)
        if out is None:
            out = (np.empty((Q, k), np.float32), np.empty((Q, k), np.int64), np.empty(Q, np.int32), np.empty(Q, np.int64))
        D, I, npb, cmp_ = out
        C.check(C.lib().lira_probe_search_wait(self._h, int(slot), C.ptr(D, C.c_f32p), C.ptr(I, C.c_i64p), C.ptr(npb, C.c_i32p),
                                               C.ptr(cmp_, C.c_i64p)))
        return out

    def select_search_dev(self, d_scores, d_q, mode, value, k, dedup=True, out=None, stream=None):
        import torch
        Q = d_q.shape[0]
        if out is None:
            out = (torch.empty((Q, k), dtype=torch.float32, device=d_q.device),
                   torch.empty((Q, k), dtype=torch.int64, device=d_q.device),
                   torch.empty((Q,), dtype=torch.int32, device=d_q.device),
                   torch.empty((Q,), dtype=torch.int64, device=d_q.device))
        D, I, npb, cmp_ = out
        st = _stream_handle(d_q.device, stream)
        C.check(C.lib().lira_select_search_dev(self._h, d_scores.data_ptr(), d_scores.stride(0), d_q.data_ptr(),
                                               d_q.stride(0), Q, int(mode), float(value), int(k), int(bool(dedup)),
                                               D.data_ptr(), I.data_ptr(), npb.data_ptr(), cmp_.data_ptr(), st))
        return out

    def search_dev(self, d_q, d_probe_offsets, d_probe_ids, k, dedup=True, stream=None):
        import torch
        Q = d_q.shape[0]
        D = torch.empty((Q, k), dtype=torch.float32, device=d_q.device)
        I = torch.empty((Q, k), dtype=torch.int64, device=d_q.device)
        cmp_ = torch.empty((Q,), dtype=torch.int64, device=d_q.device)
        st = _stream_handle(d_q.device, stream)
        C.check(C.lib().lira_search_dev(self._h, d_q.data_ptr(), d_q.stride(0), Q, d_probe_offsets.data_ptr(),
                                        d_probe_ids.data_ptr(), int(d_probe_ids.numel()), int(k), int(bool(dedup)),
                                        D.data_ptr(), I.data_ptr(), cmp_.data_ptr(), st))
        return D, I, cmp_

    # ---- scan implementation choice ---------------------------------------------------------
    def set_use_tensor_cores(self, enable=True):
        """Pin the exact CUDA-core scan (False) or allow the tcgen05 scan where it is exact (True, default)."""
        C.check(C.lib().lira_index_set_use_tensor_cores(self._h, int(bool(enable))))

    @property
    def last_path(self) -> str:
        return {0: "cuda-core", 1: "tensor-core"}.get(int(C.lib().lira_index_last_path(self._h)), "?")

    @property
    def last_redo(self) -> int:
        return int(C.lib().lira_index_last_redo(self._h))

    @property
    def last_scan_kind(self) -> str:
        """Kernel family of the last batch: "cuda-core", "fp16" (tcgen05 kind::f16 scan) or "u8" (tcgen05 kind::i8 byte scan)."""
        return {0: "cuda-core", 1: "fp16", 2: "u8"}.get(int(C.lib().lira_index_last_scan_kind(self._h)), "?")

    @property
    def byte_scan_eligible(self) -> bool:
        return bool(C.lib().lira_index_byte_scan_eligible(self._h))

    @property
    def tensor_core_eligible(self) -> bool:
        return bool(C.lib().lira_index_tensor_core_eligible(self._h))

    @property
    def tensor_core_mode(self) -> str:
        """'exact' (small-integer data, bit-identical to the CUDA cores), 'approximate' (real-valued data: fp16 filter with
        an error margin + exact fp32 re-rank) or 'none'."""
        return {1: "exact", 2: "approximate"}.get(int(C.lib().lira_index_tensor_core_mode(self._h)), "none")

    # ---- instrumentation ------------------------------------------------------------------
    def set_timing(self, enable=True):
        C.check(C.lib().lira_index_set_timing(self._h, int(enable)))

    def last_timing(self):
        scan_ms, total_ms = ctypes.c_float(), ctypes.c_float()
        nbytes, pairs = ctypes.c_int64(), ctypes.c_int64()
        C.check(C.lib().lira_index_last_timing(self._h, ctypes.byref(scan_ms), ctypes.byref(total_ms),
                                               ctypes.byref(nbytes), ctypes.byref(pairs)))
        trio = ctypes.c_float()
        C.check(C.lib().lira_index_last_scan_total_ms(self._h, ctypes.byref(trio)))
        return {"scan_ms": scan_ms.value, "total_ms": total_ms.value, "scan_bytes": nbytes.value,
                "scan_pairs": pairs.value, "scan_total_ms": trio.value}


class ListView:
    """One bucket seen as a faiss.IndexFlat: `.search(q, k) -> (D, I_local)` and `.ntotal`
    (exactly what LIRA_smallscale.py:168-171 uses)."""

    def __init__(self, index: LiraIndex, b: int):
        self.index, self.b = index, b

    @property
    def ntotal(self):
        return self.index.ntotal(self.b)

    def search(self, q, k):
        return self.index.list_search(self.b, q, k)


class LiraModel:
    """Probing model resident on the device: centroids, scaler and the six Linear layers."""

    def __init__(self, handle, n_bkt, dim, device):
        self._h, self.n_bkt, self.dim, self.device = handle, n_bkt, dim, device

    @classmethod
    def from_arrays(cls, centroids, scaler_mean, scaler_scale, weights, device=0):
        """weights: [W1,b1,...,W6,b6] in MLP_2_Input.state_dict() order (model_probing.py:12-31)."""
        C.require_gpu()
        cent = C.f32(centroids)
        B, d = cent.shape
        mean = None if scaler_mean is None else C.f32(scaler_mean)
        scale = None if scaler_scale is None else C.f32(scaler_scale)
        ws = [C.f32(w) for w in weights]
        assert len(ws) == 12
        shapes = [(128, B), (128,), (64, 128), (64,), (128, d), (128,), (64, 128), (64,), (128, 128), (128,), (B, 128), (B,)]
        for w, s in zip(ws, shapes):
            if tuple(w.shape) != s:
                raise ValueError(f"weight shape {w.shape} != expected {s}")
        arr = (C.c_f32p * 12)(*[C.ptr(w, C.c_f32p) for w in ws])
        h = ctypes.c_void_p()
        C.check(C.lib().lira_model_create(C.ptr(cent, C.c_f32p), C.ptr(mean, C.c_f32p), C.ptr(scale, C.c_f32p), B, d,
                                          arr, device, ctypes.byref(h)))
        return cls(h, B, d, device)

    @classmethod
    def from_torch(cls, model, centroids, scaler_mean, scaler_scale, device=0):
        sd = model.state_dict()
        keys = ("distance_net.0", "distance_net.2", "vector_net.0", "vector_net.2", "fc.0", "fc.2")
        ws = []
        for k in keys:
            ws.append(sd[k + ".weight"].detach().float().cpu().numpy())
            ws.append(sd[k + ".bias"].detach().float().cpu().numpy())
        return cls.from_arrays(centroids, scaler_mean, scaler_scale, ws, device)

    def set_use_tensor_cores(self, enable=True):
        """tcgen05 forward with error-compensated TF32 (default) or the fp32 CUDA-core kernels (False)."""
        C.check(C.lib().lira_model_set_use_tensor_cores(self._h, int(bool(enable))))

    def scores(self, q, return_features=False):
        """all_outputs[Q,B] of model_evaluate / model_infer for raw queries (host arrays)."""
        q = C.f32(q).reshape(-1, self.dim)
        out = np.empty((q.shape[0], self.n_bkt), np.float32)
        feats = np.empty((q.shape[0], self.n_bkt), np.float32) if return_features else None
        C.check(C.lib().lira_model_scores(self._h, C.ptr(q, C.c_f32p), q.shape[0], C.ptr(out, C.c_f32p),
                                          C.ptr(feats, C.c_f32p)))
        return (out, feats) if return_features else out

    def close(self):
        if self._h is not None:
            C.lib().lira_model_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def centroid_features(q, centroids, mean=None, scale=None, device=0):
    """get_dist_cid (+ StandardScaler.transform): utils.py:98-118, 142-167 / search.cpp:220-250."""
    C.require_gpu()
    q, cent = C.f32(q), C.f32(centroids)
    out = np.empty((q.shape[0], cent.shape[0]), np.float32)
    m = None if mean is None else C.f32(mean)
    s = None if scale is None else C.f32(scale)
    C.check(C.lib().lira_centroid_features(C.ptr(q, C.c_f32p), q.shape[0], C.ptr(cent, C.c_f32p), cent.shape[0],
                                           cent.shape[1], C.ptr(m, C.c_f32p), C.ptr(s, C.c_f32p), device,
                                           C.ptr(out, C.c_f32p)))
    return out


class KnnIndex:
    """Base vectors resident on the device for exact kNN: faiss.IndexFlat{L2,IP}(d) + .add(x) + .search(q, k) as
    compute_knn.cpp:208-244 / utils.py:293-310 / LIRA_largescale.py:225-229 use it (global ids, best first)."""

    def __init__(self, base, metric="L2", device=0):
        C.require_gpu()
        self._h = ctypes.c_void_p()
        if hasattr(base, "data_ptr"):   # torch CUDA tensor [N, d] (row stride % 4 == 0): adopted, not copied
            if not base.is_cuda or base.dtype.__str__() != "torch.float32" or base.stride(1) != 1:
                raise ValueError("device base must be a float32 CUDA tensor with unit column stride")
            self.ntotal, self.dim, self.device = base.shape[0], base.shape[1], base.device.index
            self._keep = base
            C.check(C.lib().lira_knn_create_dev(base.data_ptr(), base.stride(0), base.shape[0], base.shape[1], _metric_code(metric),
                                                self.device, ctypes.byref(self._h)))
            return
        base = C.f32(base)
        self.ntotal, self.dim, self.device = base.shape[0], base.shape[1], device
        C.check(C.lib().lira_knn_create(C.ptr(base, C.c_f32p), base.shape[0], base.shape[1], _metric_code(metric), device,
                                        ctypes.byref(self._h)))

    def search(self, query, k):
        query = C.f32(query).reshape(-1, self.dim)
        D = np.empty((query.shape[0], k), np.float32)
        I = np.empty((query.shape[0], k), np.int64)
        C.check(C.lib().lira_knn_search(self._h, C.ptr(query, C.c_f32p), query.shape[0], int(k), C.ptr(D, C.c_f32p),
                                        C.ptr(I, C.c_i64p)))
        return D, I

    def search_dev(self, d_query, k, out=None, stream=None):
        """torch CUDA queries [Q, d] -> (D[Q,k] float32, I[Q,k] int64) CUDA tensors; no host round trip."""
        import torch
        Q = d_query.shape[0]
        if out is None:
            out = (torch.empty((Q, k), dtype=torch.float32, device=d_query.device),
                   torch.empty((Q, k), dtype=torch.int64, device=d_query.device))
        C.check(C.lib().lira_knn_search_dev(self._h, d_query.data_ptr(), d_query.stride(0), Q, int(k), out[0].data_ptr(),
                                            out[1].data_ptr(), _stream_handle(d_query.device, stream)))
        return out

    def set_use_tensor_cores(self, enable=True):
        C.check(C.lib().lira_knn_set_use_tensor_cores(self._h, int(bool(enable))))

    @property
    def last_path(self) -> str:
        return {0: "cuda-core", 1: "tensor-core", 2: "mixed"}.get(int(C.lib().lira_knn_last_path(self._h)), "?")

    @property
    def last_redo(self) -> int:
        return int(C.lib().lira_knn_last_redo(self._h))

    @property
    def last_scan_kind(self) -> str:
        return {0: "cuda-core", 1: "fp16", 2: "u8"}.get(int(C.lib().lira_knn_last_scan_kind(self._h)), "?")

    def close(self):
        if self._h is not None:
            C.lib().lira_knn_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def knn(base, query, k, metric="L2", device=0):
    """Exact kNN (compute_knn.cpp:208-259 / utils.py:293-310). Returns (D[Q,k], I[Q,k] int64)."""
    C.require_gpu()
    base, query = C.f32(base), C.f32(query)
    D = np.empty((query.shape[0], k), np.float32)
    I = np.empty((query.shape[0], k), np.int64)
    C.check(C.lib().lira_knn(C.ptr(base, C.c_f32p), base.shape[0], C.ptr(query, C.c_f32p), query.shape[0],
                             base.shape[1], int(k), _metric_code(metric), device, C.ptr(D, C.c_f32p),
                             C.ptr(I, C.c_i64p)))
    return D, I


def knn_ivf(base, k, nlist, nprobe, seed=1234, device=0):
    """IVF-approximate self-kNN (compute_knn.cpp:158-203). Returns (D[N,k], I[N,k] int64)."""
    C.require_gpu()
    base = C.f32(base)
    D = np.empty((base.shape[0], k), np.float32)
    I = np.empty((base.shape[0], k), np.int64)
    C.check(C.lib().lira_knn_ivf(C.ptr(base, C.c_f32p), base.shape[0], base.shape[1], int(k), int(nlist), int(nprobe), int(seed),
                                 device, C.ptr(D, C.c_f32p), C.ptr(I, C.c_i64p)))
    return D, I


def kmeans_train(x, n_bkt, niter=20, seed=1234, init_centroids=None, device=0):
    """faiss.Kmeans(d, k, niter).train(x) -> centroids [k, d] float32 (utils.py:321-324); Lloyd on the device."""
    C.require_gpu()
    x = C.f32(x)
    out = np.empty((n_bkt, x.shape[1]), np.float32)
    init = None if init_centroids is None else C.f32(init_centroids)
    C.check(C.lib().lira_kmeans_train(C.ptr(x, C.c_f32p), x.shape[0], x.shape[1], int(n_bkt), int(niter), int(seed),
                                      C.ptr(init, C.c_f32p), device, C.ptr(out, C.c_f32p)))
    return out


def centroid_features_dev(d_x, d_centroids, d_mean=None, d_scale=None, out=None, stream=None):
    """get_dist_cid (+ transform) for torch CUDA tensors; returns a CUDA tensor [n, B] (row stride rounded up to 4)."""
    import torch
    n, d = d_x.shape
    B = d_centroids.shape[0]
    Bp = (B + 3) // 4 * 4
    if out is None:
        out = torch.empty((n, Bp), dtype=torch.float32, device=d_x.device)
    C.check(C.lib().lira_centroid_features_dev(d_x.data_ptr(), d_x.stride(0), n, d_centroids.data_ptr(), d_centroids.stride(0), B, d,
                                               None if d_mean is None else d_mean.data_ptr(), None if d_scale is None else d_scale.data_ptr(),
                                               out.data_ptr(), out.stride(0), d_x.device.index, _stream_handle(d_x.device, stream)))
    return out[:, :B]


def feature_stats_dev(d_x, d_centroids, stream=None):
    """StandardScaler statistics of the centroid distances of the rows of d_x: (mean[B], var[B]) float64 numpy arrays."""
    B = d_centroids.shape[0]
    mean, var = np.empty(B, np.float64), np.empty(B, np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    C.check(C.lib().lira_feature_stats_dev(d_x.data_ptr(), d_x.stride(0), d_x.shape[0], d_centroids.data_ptr(), d_centroids.stride(0), B,
                                           d_x.shape[1], mean.ctypes.data_as(dp), var.ctypes.data_as(dp), d_x.device.index,
                                           _stream_handle(d_x.device, stream)))
    return mean, var


def mul_partition_dev(d_score, d_data_2_bkt, d_added, first=0, d_points=None, sigma=0.5, stream=None):
    """The redundancy rule on the device (LIRA_smallscale.py:77-97): d_score [rows, B] float32, d_data_2_bkt [N, n_mul] int32
    (updated in place), d_added [N, n_mul] int32 preset to -1."""
    C.check(C.lib().lira_mul_partition_dev(d_score.data_ptr(), d_score.stride(0), d_score.shape[0], d_score.shape[1], float(sigma),
                                           None if d_points is None else d_points.data_ptr(), int(first), d_data_2_bkt.shape[1],
                                           d_data_2_bkt.data_ptr(), d_added.data_ptr(), d_score.device.index,
                                           _stream_handle(d_score.device, stream)))


def launch_count() -> int:
    return int(C.lib().lira_launch_count())
