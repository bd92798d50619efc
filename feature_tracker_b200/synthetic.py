"""Seeded synthetic workloads (numpy only) shaped like the reference's EuRoC fixture.

Used by the tests and by bench.py; nothing here is part of the hot path.
"""
import numpy as np


def _gauss_kernel(sigma):
    r = max(1, int(3.0 * sigma + 0.5))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    return (k / k.sum()).astype(np.float32)


def _smooth(img, sigma):
    k = _gauss_kernel(sigma)
    r = len(k) // 2
    p = np.pad(img, ((0, 0), (r, r)), mode="reflect")
    out = np.zeros_like(img)
    for i, w in enumerate(k):
        out += w * p[:, i:i + img.shape[1]]
    p = np.pad(out, ((r, r), (0, 0)), mode="reflect")
    out2 = np.zeros_like(img)
    for i, w in enumerate(k):
        out2 += w * p[i:i + img.shape[0], :]
    return out2


def make_image(rows, cols, seed):
    """uint8 image: three octaves of smoothed noise (mean ~95, std ~50) plus 400 bright/dark blobs for corners."""
    rng = np.random.default_rng(seed)
    base = np.zeros((rows, cols), np.float32)
    for sigma, w in ((1.5, 0.5), (4.0, 0.3), (12.0, 0.2)):
        n = _smooth(rng.uniform(-1.0, 1.0, (rows, cols)).astype(np.float32), sigma)
        base += w * n / (n.std() + 1e-6)
    base = 95.0 + 50.0 * base / (base.std() + 1e-6)
    for _ in range(400):
        r, c = int(rng.integers(0, rows)), int(rng.integers(0, cols))
        h, w = int(rng.integers(3, 12)), int(rng.integers(3, 12))
        val = float(rng.choice([-1.0, 1.0]) * rng.uniform(40.0, 90.0))
        if rng.uniform() < 0.5:
            base[max(r - h, 0):r + h, max(c - w, 0):c + w] += val
        else:
            yy, xx = np.ogrid[max(r - h, 0):min(r + h, rows), max(c - h, 0):min(c + h, cols)]
            base[max(r - h, 0):min(r + h, rows), max(c - h, 0):min(c + h, cols)] += val * (((yy - r) ** 2 + (xx - c) ** 2) <= h * h)
    return np.clip(base, 0, 255).astype(np.uint8)


def _bilinear(img, ys, xs):
    rows, cols = img.shape
    ys = np.clip(ys, 0, rows - 1.001)
    xs = np.clip(xs, 0, cols - 1.001)
    y0 = np.floor(ys).astype(np.int64)
    x0 = np.floor(xs).astype(np.int64)
    fy = (ys - y0).astype(np.float32)
    fx = (xs - x0).astype(np.float32)
    f = img.astype(np.float32)
    return (1 - fy) * (1 - fx) * f[y0, x0] + (1 - fy) * fx * f[y0, x0 + 1] + fy * (1 - fx) * f[y0 + 1, x0] + fy * fx * f[y0 + 1, x0 + 1]


def warp_image(ref, seed, max_shift=8.0, max_rot_deg=2.0, max_scale=0.02):
    """cur = ref under a random similarity, with gain / bias / noise.  Returns (cur, fwd) where fwd maps ref uv -> cur uv."""
    rng = np.random.default_rng(seed)
    rows, cols = ref.shape
    tx, ty = rng.uniform(-max_shift, max_shift, 2)
    th = np.deg2rad(rng.uniform(-max_rot_deg, max_rot_deg))
    s = rng.uniform(1.0 - max_scale, 1.0 + max_scale)
    cx, cy = cols / 2.0, rows / 2.0
    a, b = s * np.cos(th), s * np.sin(th)
    # forward: cur = A (ref - c) + c + t ; sample cur at pixel p: ref = A^-1 (p - c - t) + c
    det = a * a + b * b
    ia, ib = a / det, b / det
    yy, xx = np.mgrid[0:rows, 0:cols].astype(np.float32)
    dx, dy = xx - cx - tx, yy - cy - ty
    rx = ia * dx + ib * dy + cx
    ry = -ib * dx + ia * dy + cy
    cur = _bilinear(ref, ry, rx)
    cur = cur * rng.uniform(0.9, 1.1) + rng.uniform(-10.0, 10.0) + rng.normal(0.0, 2.0, cur.shape)
    cur = np.clip(cur, 0, 255).astype(np.uint8)

    def fwd(uv):
        uv = np.asarray(uv, np.float64)
        x, y = uv[:, 0] - cx, uv[:, 1] - cy
        return np.stack([a * x - b * y + cx + tx, b * x + a * y + cy + ty], 1)

    return cur, fwd


def _box(img, r):
    p = np.pad(img, ((r + 1, r), (r + 1, r)), mode="edge").astype(np.float64).cumsum(0).cumsum(1)
    n = 2 * r + 1
    return (p[n:, n:] - p[:-n, n:] - p[n:, :-n] + p[:-n, :-n]).astype(np.float32)


def detect_features(img, n, seed, min_distance=None, border=None, border_fraction=0.05, jitter=True):
    """Top-n Shi-Tomasi corners on a min-distance grid (interior set) + a few uniformly placed border features."""
    rng = np.random.default_rng(seed)
    rows, cols = img.shape
    f = img.astype(np.float32)
    gx = np.zeros_like(f)
    gy = np.zeros_like(f)
    gx[:, 1:-1] = f[:, 2:] - f[:, :-2]
    gy[1:-1, :] = f[2:, :] - f[:-2, :]
    sxx, syy, sxy = _box(gx * gx, 2), _box(gy * gy, 2), _box(gx * gy, 2)
    score = 0.5 * (sxx + syy - np.sqrt((sxx - syy) ** 2 + 4.0 * sxy * sxy))
    if border is None:
        border = 16
    n_border = int(round(n * border_fraction))
    n_interior = n - n_border
    if min_distance is None:
        min_distance = max(2, int(np.sqrt((rows - 2 * border) * (cols - 2 * border) / max(n_interior, 1)) * 0.9))
    cell = min_distance
    r0, c0 = border, border
    gh, gw = (rows - 2 * border) // cell, (cols - 2 * border) // cell
    view = score[r0:r0 + gh * cell, c0:c0 + gw * cell].reshape(gh, cell, gw, cell).transpose(0, 2, 1, 3).reshape(gh, gw, cell * cell)
    arg = view.argmax(2)
    best = view.max(2)
    ys = (np.arange(gh)[:, None] * cell + arg // cell + r0).reshape(-1)
    xs = (np.arange(gw)[None, :] * cell + arg % cell + c0).reshape(-1)
    order = np.argsort(-best.reshape(-1), kind="stable")[:n_interior]
    pts = np.stack([xs[order], ys[order]], 1).astype(np.float32)
    if len(pts) < n_interior:  # not enough cells: pad with uniform interior points
        extra = n_interior - len(pts)
        pts = np.concatenate([pts, np.stack([rng.uniform(border, cols - border, extra), rng.uniform(border, rows - border, extra)], 1).astype(np.float32)])
    if n_border > 0:
        side = rng.integers(0, 4, n_border)
        bx = rng.uniform(0, cols - 1, n_border)
        by = rng.uniform(0, rows - 1, n_border)
        d = rng.uniform(0, border, n_border)
        bx = np.where(side == 0, d, np.where(side == 1, cols - 1 - d, bx))
        by = np.where(side == 2, d, np.where(side == 3, rows - 1 - d, by))
        pts = np.concatenate([pts, np.stack([bx, by], 1).astype(np.float32)])
    if jitter:
        pts = pts + rng.uniform(0.0, 1.0, pts.shape).astype(np.float32)
    return np.ascontiguousarray(pts[:n], dtype=np.float32)


def make_pair(rows, cols, n_features, pair_id, **kw):
    """One frame pair + features, seed = 1234 + pair_id."""
    seed = 1234 + pair_id
    ref = make_image(rows, cols, seed)
    cur, fwd = warp_image(ref, seed + 100003)
    uv = detect_features(ref, n_features, seed + 200003, **kw)
    return ref, cur, uv, fwd


def make_brief_sets(n_ref, n_cur, bits=256, flip=0.1, distractor_fraction=0.2, rows=480, cols=752, seed=7):
    """C4: cur = permuted ref with bits flipped w.p. `flip`, plus distractors; positions uniform, pred = cur_pos + N(0, 10)."""
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 2, (n_ref, bits), dtype=np.uint8)
    n_match = min(n_ref, int(round(n_cur * (1.0 - distractor_fraction))))
    perm = rng.permutation(n_ref)[:n_match]
    cur = np.empty((n_cur, bits), np.uint8)
    cur[:n_match] = ref[perm] ^ (rng.uniform(size=(n_match, bits)) < flip).astype(np.uint8)
    cur[n_match:] = rng.integers(0, 2, (n_cur - n_match, bits), dtype=np.uint8)
    order = rng.permutation(n_cur)
    cur = cur[order]
    truth = np.full(n_ref, -1, np.int64)
    inv = np.empty(n_cur, np.int64)
    inv[order] = np.arange(n_cur)
    truth[perm] = inv[:n_match]
    cur_pos = np.stack([rng.uniform(0, cols - 1, n_cur), rng.uniform(0, rows - 1, n_cur)], 1).astype(np.float32)
    pred = np.stack([rng.uniform(0, cols - 1, n_ref), rng.uniform(0, rows - 1, n_ref)], 1).astype(np.float32)
    has = truth >= 0
    pred[has] = cur_pos[truth[has]] + rng.normal(0.0, 10.0, (int(has.sum()), 2)).astype(np.float32)
    return ref, cur, pred, cur_pos, truth


def make_float_sets(n_ref, n_cur, dim=256, noise=0.2, seed=11):
    """C5: unit-norm N(0,1) descriptors; cur = normalise(ref + noise * N(0,1)) permuted (extra cur rows are random)."""
    rng = np.random.default_rng(seed)
    raw = rng.normal(size=(n_ref, dim)).astype(np.float32)
    ref = raw / np.linalg.norm(raw, axis=1, keepdims=True)
    cur = rng.normal(size=(n_cur, dim)).astype(np.float32)
    m = min(n_ref, n_cur)
    cur[:m] = raw[:m] + noise * rng.normal(size=(m, dim)).astype(np.float32)
    cur /= np.linalg.norm(cur, axis=1, keepdims=True)
    order = rng.permutation(n_cur)
    return ref, np.ascontiguousarray(cur[order])


def make_direct_method_scene(rows, cols, n_features, pair_id, depth=5.0, focal=400.0, **kw):
    """A frame pair for the direct-method pose tracker: the similarity-warped pair of make_pair() read as a camera moving in
    front of a fronto-parallel plane at `depth` (translation ~ image shift * depth / focal, roll ~ image rotation).  Returns
    (ref, cur, ref_uv, K = (fx, fy, cx, cy), p_c_in_ref [n, 3])."""
    ref, cur, uv, _ = make_pair(rows, cols, n_features, pair_id, **kw)
    K = np.array([focal, focal, cols / 2.0, rows / 2.0], np.float32)
    z = np.full(uv.shape[0], depth, np.float32)
    pts = np.stack([(uv[:, 0] - K[2]) / K[0] * z, (uv[:, 1] - K[3]) / K[1] * z, z], axis=1).astype(np.float32)
    return ref, cur, uv, K, pts
