"""CPU checks of the detector / BRIEF checker.  PARITY UNPINNED: feature_detector::FeaturePointHarrisDetector
and feature_detector::BriefDescriptor (called at test/test_descriptor_matcher_brief.cpp:59-76) live in the absent sibling
repository Feature_Detector, so there is nothing of the reference's to compile or compare with.  These tests pin oracle/ftk_oracle.c's
restatement of the published algorithm against an independent numpy restatement, against the properties the sequential selection
must have, and against the committed regression fixture."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits_equal
from feature_tracker_b200 import synthetic as S
from oracle import pyoracle as po


def numpy_response(img, kind, h, k):
    """Same arithmetic in numpy float32: integer tensor sums, then one rounding per operation."""
    f = np.float32
    I = img.astype(np.int64)
    rows, cols = I.shape
    gx = np.zeros_like(I)
    gy = np.zeros_like(I)
    gx[:, 1:-1] = I[:, 2:] - I[:, :-2]
    gy[1:-1, :] = I[2:, :] - I[:-2, :]
    out = np.full((rows, cols), -np.inf, np.float32)
    m = h + 1
    if rows <= 2 * m or cols <= 2 * m:
        return out
    def box(a):
        s = np.zeros((rows - 2 * m, cols - 2 * m), np.int64)
        for dr in range(-h, h + 1):
            for dc in range(-h, h + 1):
                s += a[m + dr:rows - m + dr, m + dc:cols - m + dc]
        return s.astype(np.float32)
    inv = f(1.0) / f(4 * (2 * h + 1) ** 2)
    a, b, c = box(gx * gx) * inv, box(gx * gy) * inv, box(gy * gy) * inv
    if kind == "harris":
        tr = a + c
        r = (a * c - b * b) - f(k) * (tr * tr)
    else:
        d = a - c
        r = f(0.5) * ((a + c) - np.sqrt(d * d + f(4.0) * (b * b)))
    out[m:rows - m, m:cols - m] = r
    return out


@pytest.mark.parametrize("kind", ["harris", "shi_tomasi"])
@pytest.mark.parametrize("half", [1, 2, 3])
def test_response_restatement_vs_numpy(oracle, kind, half):
    for shape, seed in (((97, 131), 5), ((480, 752), 6), ((9, 9), 7), ((4, 40), 8)):
        img = S.make_image(*shape, seed=seed) if min(shape) >= 32 else np.random.default_rng(seed).integers(0, 256, shape, dtype=np.uint8)
        ok, r = oracle.detect_response(po.make_detector_params(kind, half, 0.04, 40.0, 20), img)
        assert ok
        assert bits_equal(r, numpy_response(img, kind, half, 0.04)), (kind, half, shape)


def python_selection(response, min_response, dist, needed, existing=()):
    """The sequential loop, literally: order by (response desc, index asc), take unless a taken / existing feature is near."""
    rows, cols = response.shape
    idx = np.nonzero(response.ravel() >= np.float32(min_response))[0]
    order = idx[np.lexsort((idx, -response.ravel()[idx].astype(np.float64)))]
    taken = []
    blockers = [(int(y), int(x)) for x, y in existing if 0 <= x < cols and 0 <= y < rows]
    for p in order:
        if len(taken) >= needed:
            break
        r, c = divmod(int(p), cols)
        if dist > 0 and any(abs(r - br) < dist and abs(c - bc) < dist for br, bc in blockers):
            continue
        taken.append((c, r))
        blockers.append((r, c))
    return np.array(taken, np.float32).reshape(-1, 2)


@pytest.mark.parametrize("kind,thr,dist,needed", [("harris", 1e5, 9, 60), ("shi_tomasi", 300.0, 5, 1000), ("harris", 40.0, 20, 25), ("shi_tomasi", 1e9, 4, 10),
                                                  ("harris", 5e5, 0, 40), ("harris", 5e5, 1, 40)])
def test_selection_restatement_vs_python(oracle, kind, thr, dist, needed):
    img = S.make_image(120, 160, seed=11)
    prm = po.make_detector_params(kind, 1, 0.04, thr, dist)
    _, resp = oracle.detect_response(prm, img)
    existing = np.array([[40.5, 30.2], [500.0, 3.0], [np.nan, 4.0], [159.9, 119.9], [-0.5, 10.0]], np.float32)
    for ex in (None, existing):
        ok, uv, r = oracle.detect_features(prm, img, needed, existing=ex)
        assert ok
        exp = python_selection(resp, thr, dist, needed, () if ex is None else ex)
        assert np.array_equal(uv, exp), (kind, thr, dist, needed, ex is not None)
        assert bits_equal(r, resp[uv[:, 1].astype(int), uv[:, 0].astype(int)])
        assert (np.diff(r) <= 0).all() and (r >= np.float32(thr)).all()
        if dist > 0 and len(uv) > 1:
            d = np.abs(uv[:, None, :] - uv[None, :, :]).max(axis=2) + np.eye(len(uv)) * 1e9
            assert d.min() >= dist


def test_detector_rejects_bad_parameters(oracle):
    img = S.make_image(64, 64, seed=1)
    ok, _, _ = oracle.detect_features(po.make_detector_params("harris", 0, 0.04, 40.0, 20), img, 10)
    assert not ok
    ok, _, _ = oracle.detect_features(po.make_detector_params("harris", 4, 0.04, 40.0, 20), img, 10)
    assert not ok


def test_brief_restatement_vs_numpy(oracle):
    img = S.make_image(90, 120, seed=21)
    rng = np.random.default_rng(3)
    uv = np.concatenate([rng.uniform(-5, 125, (200, 2)).astype(np.float32), np.array([[8.0, 8.0], [111.99, 81.99], [112.0, 40.0], [np.nan, 10], [7.99, 30.0]], np.float32)])
    for n_bits, half, seed in ((256, 8, 0), (128, 4, 77), (32, 15, 1)):
        pattern = oracle.brief_pattern(n_bits, half, seed)
        assert pattern.min() >= -half and pattern.max() <= half
        assert not ((pattern[:, 0] == pattern[:, 2]) & (pattern[:, 1] == pattern[:, 3])).any()
        assert np.array_equal(pattern, oracle.brief_pattern(n_bits, half, seed))
        ok, desc, valid = oracle.describe_brief(img, uv, pattern, half)
        assert ok
        exp_valid = (uv[:, 0] >= half) & (uv[:, 1] >= half) & (uv[:, 0] < 120 - half) & (uv[:, 1] < 90 - half)
        assert np.array_equal(valid.astype(bool), exp_valid)
        assert valid[200] == (1 if half <= 8 else 0) and valid[203] == 0
        for i in range(len(uv)):
            bits = np.zeros(n_bits, np.uint8)
            if exp_valid[i]:
                r, c = int(uv[i, 1]), int(uv[i, 0])
                bits = (img[r + pattern[:, 0], c + pattern[:, 1]] < img[r + pattern[:, 2], c + pattern[:, 3]]).astype(np.uint8)
            words = np.packbits(bits.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).ravel()
            assert np.array_equal(desc[i], words), (n_bits, i)


def test_detector_matches_regression_fixture(oracle, euroc_golden):
    g = dict(np.load(os.path.join(GOLDEN, "detector_golden.npz")))
    assert np.array_equal(oracle.brief_pattern(256, 8, 0), g["pattern"])
    for name in ("ref", "cur"):
        img = euroc_golden[name]
        for kind in ("harris", "shi_tomasi"):
            ok, uv, resp = oracle.detect_features(po.make_detector_params(kind, 1, 0.04, 40.0, 20), img, 300)
            assert ok and np.array_equal(uv, g[f"{name}_{kind}_uv"]) and bits_equal(resp, g[f"{name}_{kind}_response"])
        ok, desc, valid = oracle.describe_brief(img, g[f"{name}_harris_uv"], g["pattern"], 8)
        assert ok and np.array_equal(desc, g[f"{name}_brief"]) and np.array_equal(valid, g[f"{name}_brief_valid"])


def test_detected_features_match_across_the_euroc_pair(oracle, euroc_golden):
    """End-to-end sanity of the restated front end, the reference's own demo flow (test_descriptor_matcher_brief.cpp:57-95): detect
    in both images, describe, NearbyMatch with the demo's thresholds -- a sizeable share of the corners must pair up."""
    g = dict(np.load(os.path.join(GOLDEN, "detector_golden.npz")))
    def unpack(d):
        return np.unpackbits(d.view(np.uint8).reshape(len(d), -1), axis=1, bitorder="little")
    ok, idx = oracle.match_brief_nearby(unpack(g["ref_brief"]), unpack(g["cur_brief"]), g["ref_harris_uv"], g["cur_harris_uv"], 50, 50, 60.0)
    assert ok
    matched = idx >= 0
    assert matched.sum() >= 100, matched.sum()
    shift = g["cur_harris_uv"][idx[matched]] - g["ref_harris_uv"][matched]
    assert np.abs(shift - np.median(shift, axis=0)).max(axis=1).__lt__(12).mean() > 0.6  # the scene has parallax: the shift is not one vector


def test_product_pattern_and_defaults_match_the_checker(oracle):
    """Host-only entry points of the product library (no GPU needed): the default BRIEF pair list is the checker's, the detector
    defaults are the values the reference's demo sets (test/test_descriptor_matcher_brief.cpp:60-61)."""
    import ctypes as C
    import feature_tracker_b200 as ft
    from feature_tracker_b200 import _capi
    from feature_tracker_b200.api import lib
    for n_bits, half, seed in ((256, 8, 0), (128, 4, 77), (32, 15, 1), (1024, 20, 9), (64, 0, 5), (96, 1, 0xFFFFFFFF)):
        assert np.array_equal(ft.brief_pattern(n_bits, half, seed), oracle.brief_pattern(n_bits, half, seed)), (n_bits, half, seed)
    prm = _capi.DetectorParams()
    lib().ftk_detector_params_default(C.byref(prm))
    assert (prm.kind, prm.half_patch, prm.min_distance) == (0, 1, 20) and prm.min_response == 40.0 and abs(prm.harris_k - 0.04) < 1e-7


def test_selection_is_maximal_when_not_capped(oracle):
    """Size-independent property at the full image size: with more features wanted than exist, every candidate that was not taken has a
    taken feature of higher priority inside its window, and no two taken features block each other."""
    img = S.make_image(480, 752, seed=91)
    thr, dist = 3e4, 15
    prm = po.make_detector_params("harris", 1, 0.04, thr, dist)
    ok, uv, resp = oracle.detect_features(prm, img, 100000)
    _, rmap = oracle.detect_response(prm, img)
    assert ok and 100 < len(uv) < 100000
    taken = np.zeros(img.shape, bool)
    cols_i, rows_i = uv[:, 0].astype(int), uv[:, 1].astype(int)
    taken[rows_i, cols_i] = True
    d = np.abs(uv[:, None, :] - uv[None, :, :]).max(axis=2) + np.eye(len(uv)) * 1e9
    assert d.min() >= dist
    # best taken response within the window of every pixel (dilation by brute force over the taken list)
    best = np.full(img.shape, -np.inf, np.float32)
    for r, c, v in zip(rows_i, cols_i, resp):
        ra, rb, ca, cb = max(0, r - dist + 1), min(479, r + dist - 1), max(0, c - dist + 1), min(751, c + dist - 1)
        np.maximum(best[ra:rb + 1, ca:cb + 1], v, out=best[ra:rb + 1, ca:cb + 1])
    cand = (rmap >= np.float32(thr)) & ~taken
    assert (best[cand] >= rmap[cand]).all()


@pytest.mark.parametrize("kind,use_harris", [("harris", True), ("shi_tomasi", False)])
def test_detector_agrees_with_opencv_good_features(oracle, euroc_golden, kind, use_harris):
    """Independent cross-check (the row stays PARITY UNPINNED -- the reference's detector is absent -- but is no longer only
    self-referential): OpenCV's goodFeaturesToTrack with the same response family (Harris k = 0.04 / minimum eigenvalue), the same
    3 x 3 structure-tensor window and the same minimum distance, on the reference's EuRoC fixture.  OpenCV differentiates with Sobel
    kernels and separates features by Euclidean distance, this detector with central differences and a square window, so the sets
    cannot be equal.  Stated bar: at least 80 % of the 100 strongest and 60 % of all 300 detected corners lie within 3 px of an OpenCV
    corner (measured: 86 % / 69 % Harris, 85 % / 68 % Shi-Tomasi)."""
    cv2 = pytest.importorskip("cv2")
    img = euroc_golden["ref"]
    ok, uv, _ = oracle.detect_features(po.make_detector_params(kind, 1, 0.04, 40.0, 20), img, 300)
    pts = cv2.goodFeaturesToTrack(img, maxCorners=300, qualityLevel=1e-4, minDistance=20, blockSize=3, useHarrisDetector=use_harris, k=0.04).reshape(-1, 2)
    assert ok and len(uv) == 300 and len(pts) == 300
    nearest = np.linalg.norm(uv[:, None, :] - pts[None, :, :], axis=2).min(1)
    assert (nearest[:100] <= 3.0).mean() >= 0.80
    assert (nearest <= 3.0).mean() >= 0.60
