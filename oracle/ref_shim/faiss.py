"""TEST INFRASTRUCTURE -- stand-in `faiss` module so that the reference's own Python
(/root/reference/utils.py, LIRA_smallscale.py, ...) can be IMPORTED UNMODIFIED in this
container, where faiss-cpu==1.9.0 (requirements.txt:6) is not installed.

It exposes exactly the surface the reference touches: omp_set_num_threads,
IndexFlatL2 / IndexFlatIP with add / search / ntotal, and Kmeans with train / centroids / index.
The arithmetic is the oracle's restatement of Faiss' published IndexFlat behaviour
(oracle/lira_oracle.c: oracle_list_search); Kmeans is a plain Lloyd restatement (build side,
out of the hot-path scope -- it only has to produce *a* partition for the fixtures).
Used by oracle/make_golden.py and tests/ only.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as _o  # noqa: E402

__version__ = "1.9.0-shim"
_PREC = _o.F32  # fp32 direct-difference scan, as libfaiss fvec_L2sqr / fvec_inner_product


def omp_set_num_threads(n):
    _o.set_num_threads(n)


class _IndexFlat:
    metric = _o.L2

    def __init__(self, d):
        self.d = int(d)
        self._x = np.empty((0, self.d), np.float32)

    @property
    def ntotal(self):
        return self._x.shape[0]

    def add(self, x):
        x = np.ascontiguousarray(x, np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        self._x = np.concatenate([self._x, x], 0) if self.ntotal else x.copy()

    def search(self, x, k):
        x = np.ascontiguousarray(x, np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        return _o.list_search(self._x, x, int(k), self.metric, _PREC)


class IndexFlatL2(_IndexFlat):
    metric = _o.L2


class IndexFlatIP(_IndexFlat):
    metric = _o.IP


class Kmeans:
    """faiss.Kmeans(d, k, niter=20, verbose=True) as used at utils.py:323-325: Lloyd iterations on
    at most 256*k sampled points (max_points_per_centroid), seed 1234, L2 assignment."""

    def __init__(self, d, k, niter=20, verbose=False, seed=1234, max_points_per_centroid=256):
        self.d, self.k, self.niter, self.verbose, self.seed = int(d), int(k), int(niter), verbose, seed
        self.max_points_per_centroid = max_points_per_centroid
        self.centroids = None
        self.index = None

    def _assign(self, x, c):
        x2 = (x.astype(np.float64) ** 2).sum(1)[:, None]
        c2 = (c.astype(np.float64) ** 2).sum(1)[None, :]
        out = np.empty(x.shape[0], np.int64)
        for s in range(0, x.shape[0], 16384):
            e = min(s + 16384, x.shape[0])
            d = x2[s:e] + c2 - 2.0 * (x[s:e].astype(np.float64) @ c.astype(np.float64).T)
            out[s:e] = d.argmin(1)
        return out

    def train(self, x):
        x = np.ascontiguousarray(x, np.float32)
        rng = np.random.RandomState(self.seed)
        n = x.shape[0]
        if n > self.max_points_per_centroid * self.k:
            x = x[rng.choice(n, self.max_points_per_centroid * self.k, replace=False)]
            n = x.shape[0]
        c = x[rng.choice(n, self.k, replace=False)].copy()
        for _ in range(self.niter):
            a = self._assign(x, c)
            cnt = np.bincount(a, minlength=self.k)
            s = np.zeros((self.k, self.d), np.float64)
            np.add.at(s, a, x)
            nz = cnt > 0
            c[nz] = (s[nz] / cnt[nz, None]).astype(np.float32)
            for e in np.nonzero(~nz)[0]:  # empty cluster: split the biggest one
                big = int(cnt.argmax())
                c[e] = c[big] * (1 + 1e-4)
                c[big] = c[big] * (1 - 1e-4)
                cnt[e] = cnt[big] // 2
                cnt[big] -= cnt[e]
        self.centroids = c
        self.index = IndexFlatL2(self.d)
        self.index.add(c)
        return 0.0
