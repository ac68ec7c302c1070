"""TEST INFRASTRUCTURE (like the rest of oracle/): numpy restatement of the geographic tile-skip decision made by
range_b200/csrc/retrieval.cu:geo_mask_kernel, and the exactness criterion it has to satisfy.

The reference has no such step (range/range.py:231-236 evaluates every pair); the skip is only allowed to drop
entries whose weight exp(T (g - 1)) / l_g is at most 2^-24 / M for every row of the query tile, so that everything
dropped from a row sums to less than one fp32 ulp of its geo normaliser."""
import numpy as np

EPS = 2e-4          # the kernel's slack for acosf / fp32 dot products


def caps_of(xyz, block=128):
    """(centre, radius) of every `block` consecutive unit vectors (same rule as range_b200.database.tile_caps)"""
    M = xyz.shape[0]
    T = (M + block - 1) // block
    p = np.asarray(xyz, np.float64)
    pad = T * block - M
    if pad:
        p = np.concatenate([p, np.repeat(p[-1:], pad, axis=0)])
    p = p.reshape(T, block, 3)
    c = p.sum(1)
    n = np.linalg.norm(c, axis=1, keepdims=True)
    ok = n[:, 0] > 1e-3
    c = np.where(ok[:, None], c / np.maximum(n, 1e-30), np.array([0.0, 0.0, 1.0]))
    r = np.arccos(np.clip(np.einsum("tbi,ti->tb", p, c), -1, 1)).max(1)
    return c, np.where(ok, r, np.pi)


def skip_mask(q_xyz, db_xyz, M_total, T=40.0, lg=None):
    """bool (query tiles, database tiles): True = the geo term of that tile pair may be skipped.
    lg = None: statistics pass (lower bounds from the caps only); lg (N,) = known normalisers: apply pass."""
    cq, rq = caps_of(q_xyz)
    cd, rd = caps_of(db_xyz)
    rq = rq + EPS
    M = db_xyz.shape[0]
    cnt = np.minimum(128, M - 128 * np.arange(len(rd)))
    ang = np.arccos(np.clip(cq @ cd.T, -1, 1))
    far = ang + rq[:, None] + rd[None] + EPS
    near = ang - rq[:, None] - rd[None] - EPS
    thr_ln = np.log(M_total) + 24 * np.log(2)
    cf = np.where(far < np.pi, np.cos(np.minimum(far, np.pi)), -np.inf)
    glb = np.where(np.isfinite(cf), cf, -1.0).max(1)
    thr = glb - thr_ln / T
    llb = (cnt[None] * np.where(np.isfinite(cf), np.exp(T * (cf - 1)), 0.0)).sum(1)
    thr = np.where(llb > 0, np.maximum(thr, 1 + (np.log(np.maximum(llb, 1e-300)) - thr_ln) / T - 1e-6), thr)
    if lg is not None:
        N = q_xyz.shape[0]
        pad = len(rq) * 128 - N
        lmin = np.concatenate([lg, np.full(pad, np.inf)]).reshape(-1, 128).min(1)
        thr = np.maximum(thr, 1 + (np.log(lmin) - thr_ln) / T - 1e-6)
    return (near > 0) & (np.cos(np.maximum(near, 0)) <= thr[:, None])


def exactness_slack(q_xyz, db_xyz, M_total, mask, T=40.0):
    """smallest (thr_row - g) over every pair inside a skipped tile pair, thr_row from the exact fp64 normaliser:
    >= 0 means no skipped entry carries more than 2^-24 / M_total of its row's geo mass"""
    G = np.asarray(q_xyz, np.float64) @ np.asarray(db_xyz, np.float64).T
    lg = np.exp(T * (G - 1)).sum(1)
    thr_row = 1 + (np.log(lg) - np.log(M_total) - 24 * np.log(2)) / T
    worst = np.inf
    for qt, t in zip(*np.nonzero(mask)):
        g = G[qt * 128:(qt + 1) * 128, t * 128:(t + 1) * 128]
        worst = min(worst, (thr_row[qt * 128:(qt + 1) * 128, None] - g).min())
    return worst, lg
