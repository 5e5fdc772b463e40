"""Seeded synthetic inputs of the shapes BASELINE.json names.

Stand-ins for the descriptor matrices the reference builds at
/root/reference/Points Matching/main.cpp:37-40 (SurfDescriptorExtractor) and the
point lists it builds at main.cpp:89-91 (KeyPoint::convert); the extractor itself
(SURF, nonfree) is out of scope.
"""
import numpy as np


def sift_like(n, seed=1234, dim=128):
    """Integer-valued f32 descriptors with OpenCV-SIFT statistics (0..255, ||.||~512)."""
    rng = np.random.default_rng(seed)
    x = rng.gamma(0.6, 1.0, (n, dim))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    x = np.minimum(x, 0.2)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return np.clip(np.round(512.0 * x), 0, 255).astype(np.float32)


def sift_pair(nq, nt, seed=1234, planted=0.5, noise=4.0):
    """Query/train sets where `planted` of the queries have a noisy copy in train."""
    rng = np.random.default_rng(seed + 7)
    q = sift_like(nq, seed)
    t = sift_like(nt, seed + 1)
    k = min(int(planted * nq), nt)
    if k > 0:
        qi = rng.permutation(nq)[:k]
        ti = rng.permutation(nt)[:k]
        t[ti] = np.clip(q[qi] + np.round(rng.normal(0, noise, (k, q.shape[1]))), 0, 255)
    return q, t.astype(np.float32)


def surf_like(n, seed=77, dim=128):
    """Unit-norm signed float descriptors (default SURF(extended) statistics, D1)."""
    rng = np.random.default_rng(seed)
    x = rng.normal(0, 1, (n, dim))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)


def surf_pair(nq, nt, seed=77, planted=0.5, noise=0.02):
    rng = np.random.default_rng(seed + 7)
    q = surf_like(nq, seed)
    t = surf_like(nt, seed + 1)
    k = min(int(planted * nq), nt)
    if k > 0:
        qi = rng.permutation(nq)[:k]
        ti = rng.permutation(nt)[:k]
        y = q[qi] + rng.normal(0, noise, (k, q.shape[1]))
        t[ti] = (y / np.linalg.norm(y, axis=1, keepdims=True)).astype(np.float32)
    return q, t


def orb_like(n, seed=4321, nbytes=32):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (n, nbytes), dtype=np.uint8)


def orb_pair(nq, nt, seed=4321, planted=0.5, nbytes=32):
    """Random 256-bit rows; `planted` of the queries get a train copy with 16-40 bit flips."""
    rng = np.random.default_rng(seed + 7)
    q = orb_like(nq, seed, nbytes)
    t = orb_like(nt, seed + 1, nbytes)
    k = min(int(planted * nq), nt)
    if k > 0:
        qi = rng.permutation(nq)[:k]
        ti = rng.permutation(nt)[:k]
        nbits = nbytes * 8
        nflip = rng.integers(16, 41, size=k)
        order = np.argsort(rng.random((k, nbits)), axis=1)              # a random permutation of the bit positions per row
        flip = np.zeros((k, nbits), dtype=np.uint8)
        np.put_along_axis(flip, order, (np.arange(nbits)[None, :] < nflip[:, None]).astype(np.uint8), axis=1)
        rows = q[qi] ^ np.packbits(flip, axis=1, bitorder="little")
        t[ti] = rows
    return q, t


def correspondences(n, seed=0, outlier_frac=0.5, noise=0.3):
    """Two-view correspondences: 1920x1080, f=1400, known motion, uniform outliers."""
    rng = np.random.default_rng(seed)
    W, H, f = 1920.0, 1080.0, 1400.0
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1.0]])
    X = np.stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(4, 12, n)], 1)
    rv = np.array([0.05, -0.12, 0.03])
    th = np.linalg.norm(rv)
    k = rv / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)
    t = np.array([1.0, 0.1, 0.2])
    x1 = (K @ X.T).T
    x1 = x1[:, :2] / x1[:, 2:3]
    X2 = (R @ X.T).T + t
    x2 = (K @ X2.T).T
    x2 = x2[:, :2] / x2[:, 2:3]
    x1 = x1 + rng.normal(0, noise, x1.shape)
    x2 = x2 + rng.normal(0, noise, x2.shape)
    n_out = int(outlier_frac * n)
    out_idx = rng.permutation(n)[:n_out]
    x2[out_idx] = np.stack([rng.uniform(0, W, n_out), rng.uniform(0, H, n_out)], 1)
    gt_inlier = np.ones(n, dtype=bool)
    gt_inlier[out_idx] = False
    return x1.astype(np.float32), x2.astype(np.float32), gt_inlier


def sample_index_sets(n_points, n_hyp, m=8, seed=99):
    """[n_hyp, m] int32 minimal-sample index sets, distinct within a row; identical on every rank."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, n_points, (n_hyp, m), dtype=np.int64)
    if n_points >= m:
        for _ in range(64):
            s = np.sort(idx, axis=1)
            bad = np.nonzero((s[:, 1:] == s[:, :-1]).any(axis=1))[0]
            if bad.size == 0:
                break
            idx[bad] = rng.integers(0, n_points, (bad.size, m), dtype=np.int64)
    return idx.astype(np.int32)


def image_pair(n1, n2, seed=0, planted=0.5, outlier_frac=0.0):
    """One synthetic image pair for the end-to-end flow (BASELINE config 5): SIFT-like descriptors with
    `planted` of the queries copied (noisily) into the train set, and keypoint coordinates that obey a
    known two-view geometry on the planted matches (random elsewhere).
    Returns desc1 [n1,128], desc2 [n2,128], kp1 [n1,2], kp2 [n2,2], (qi, ti) planted index lists."""
    rng = np.random.default_rng(seed + 13)
    d1 = sift_like(n1, seed)
    d2 = sift_like(n2, seed + 1)
    k = min(int(planted * n1), n2)
    qi = rng.permutation(n1)[:k]
    ti = rng.permutation(n2)[:k]
    d2[ti] = np.clip(d1[qi] + np.round(rng.normal(0, 4.0, (k, d1.shape[1]))), 0, 255)
    x1, x2, _ = correspondences(k, seed=seed + 2, outlier_frac=outlier_frac)
    kp1 = np.stack([rng.uniform(0, 1920, n1), rng.uniform(0, 1080, n1)], 1).astype(np.float32)
    kp2 = np.stack([rng.uniform(0, 1920, n2), rng.uniform(0, 1080, n2)], 1).astype(np.float32)
    kp1[qi] = x1
    kp2[ti] = x2
    return d1, d2.astype(np.float32), kp1, kp2, (qi, ti)
