"""Shared helpers for the parity tests."""
import numpy as np

import oracle as O

RTOL = 1e-5  # north-star tolerance: fp32 distances within 1e-5 relative


def synth(n, d, nq, seed=43, integer=True, ncomp=40, sigma=0.9):
    rng = np.random.RandomState(seed)
    centres = rng.randn(ncomp, d)
    w = rng.lognormal(0, 0.5, ncomp)
    w /= w.sum()

    def draw(m):
        c = rng.choice(ncomp, m, p=w)
        x = centres[c] + sigma * rng.randn(m, d)
        if integer:
            return np.clip(np.round(32 * x + 64), 0, 255).astype(np.float32)
        return x.astype(np.float32)

    return draw(n), draw(nq)


def random_lists(n, B, rng, redundancy=0.2, empty=()):
    """Random partition of range(n) into B lists plus `redundancy` extra copies; returns cluster_ids."""
    assign = rng.randint(0, B, n)
    for e in empty:
        assign[assign == e] = (e + 1) % B
    cluster_ids = [np.nonzero(assign == b)[0].tolist() for b in range(B)]
    extra = rng.choice(n, int(n * redundancy), replace=False)
    for i in extra:
        b = int(rng.randint(0, B))
        if b in empty or b == assign[i]:
            continue
        cluster_ids[b].append(int(i))
    return cluster_ids


def pair_value(q, v, metric):
    q, v = q.astype(np.float64), v.astype(np.float64)
    return ((q - v) ** 2).sum(-1) if metric == O.L2 else (q * v).sum(-1)


def assert_topk_equiv(D, I, D_ref, I_ref, q, base, metric, rtol=RTOL):
    """ids bit-exact except for distance ties; distances within rtol (relative) of fp64."""
    D, I, D_ref, I_ref = map(np.asarray, (D, I, D_ref, I_ref))
    assert D.shape == D_ref.shape and I.shape == I_ref.shape
    # padding agrees
    assert np.array_equal(I < 0, I_ref < 0)
    ok = I >= 0
    scale = np.maximum(1.0, np.abs(D_ref[ok].astype(np.float64)))
    assert np.all(np.abs(D[ok].astype(np.float64) - D_ref[ok].astype(np.float64)) <= rtol * scale), "distances differ"
    diff = (I != I_ref) & ok
    for qi, j in zip(*np.nonzero(diff)):
        # a differing id must be a (near-)tie: its exact value equals the reference value at that rank
        mine = pair_value(q[qi], base[I[qi, j]], metric)
        ref = pair_value(q[qi], base[I_ref[qi, j]], metric)
        assert abs(mine - ref) <= rtol * max(1.0, abs(ref)), f"query {qi} rank {j}: id {I[qi, j]} vs {I_ref[qi, j]}"
    for qi in np.unique(np.nonzero(diff)[0]):
        row = I[qi][I[qi] >= 0]
        assert len(set(row.tolist())) == len(row), "duplicate id in a result row"


def merge_ranks_numpy(D_all, I_all, k, dedup=True):
    """[R,Q,k] per-rank top-k lists -> [Q,k]: ascending (score, id), an id counted once."""
    R, Q, _ = D_all.shape
    D_out = np.full((Q, k), np.inf, np.float32)
    I_out = np.full((Q, k), -1, np.int64)
    for q in range(Q):
        cand = sorted({(float(D_all[r, q, j]), int(I_all[r, q, j])) for r in range(R) for j in range(D_all.shape[2])
                       if I_all[r, q, j] >= 0})
        seen, w = set(), 0
        for dd, ii in cand:
            if dedup and ii in seen:
                continue
            seen.add(ii)
            D_out[q, w], I_out[q, w] = dd, ii
            w += 1
            if w == k:
                break
    return D_out, I_out
