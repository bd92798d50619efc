"""CPU tests of the oracle (plain-C restatement) against (a) the committed golden vectors, which were produced by the
reference's own sources compiled in place, (b) that reference build itself when it is available, and (c) domain
properties.  No GPU needed."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits_equal
from feature_tracker_b200 import synthetic as S
from oracle import pyoracle as po

KLT_COMBOS = [(v, m, h) for v in ("basic", "affine", "lssd") for m in ("inverse", "direct", "fast") for h in (6, 7, 10)]


def golden_levels(g):
    L = int(g["levels"])
    return [g["ref"]] + [g[f"ref_l{l}"] for l in range(1, L)], [g["cur"]] + [g[f"cur_l{l}"] for l in range(1, L)]


def test_pyramid_matches_golden(oracle, euroc_golden):
    g = euroc_golden
    lv = oracle.pyramid_build(g["ref"], int(g["levels"]))
    for l in range(1, int(g["levels"])):
        assert (lv[l] == g[f"ref_l{l}"]).all()
        assert lv[l].shape == (480 >> l, 752 >> l)


def test_pyramid_odd_sizes(oracle):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    lv = oracle.pyramid_build(img, 4)
    assert [a.shape for a in lv] == [(37, 53), (18, 26), (9, 13), (4, 6)]
    a = img.astype(np.uint16)
    exp = ((a[0:36:2, 0:52:2] + a[1:36:2, 0:52:2] + a[0:36:2, 1:52:2] + a[1:36:2, 1:52:2]) >> 2).astype(np.uint8)
    assert (lv[1] == exp).all()


@pytest.mark.parametrize("variant,method,half", KLT_COMBOS)
def test_klt_matches_golden(oracle, euroc_golden, variant, method, half):
    g = euroc_golden
    rl, cl = golden_levels(g)
    ok, uv, st = oracle.klt_track(po.make_params(variant, method, half=half), rl, cl, g["pts"])
    assert ok
    key = f"{variant}_{method}_h{half}"
    assert (st == g[key + "_st"]).all()
    assert bits_equal(uv, g[key + "_uv"])


def test_klt_golden_extras(oracle, euroc_golden):
    g = euroc_golden
    rl, cl = golden_levels(g)
    ok, uv, st = oracle.klt_track(po.make_params("lssd", "fast", half=6, luminance=True), rl, cl, g["pts"])
    assert (st == g["lssd_fast_h6_lum_st"]).all() and bits_equal(uv, g["lssd_fast_h6_lum_uv"])
    pred = g["pts"] + np.float32(3.0)
    for v in ("basic", "affine", "lssd"):
        p = po.make_params(v, "fast", half=6, predict=(0.9995, -0.03, 0.03, 0.9995))
        ok, uv, st = oracle.klt_track(p, rl, cl, g["pts"], cur_uv=pred, single_level=True)
        assert (st == g[f"{v}_fast_h6_single_st"]).all() and bits_equal(uv, g[f"{v}_fast_h6_single_uv"])


@pytest.mark.parametrize("variant", ["basic", "affine", "lssd"])
@pytest.mark.parametrize("method", ["inverse", "direct", "fast"])
def test_klt_matches_reference_build_on_synthetic(oracle, reflib, variant, method):
    """Bit-for-bit agreement with the reference's own .cpp on a seeded synthetic pair incl. border features."""
    ref, cur, uv, _ = S.make_pair(240, 320, 120, pair_id=3, border=12)
    rng = np.random.default_rng(9)
    uv = np.concatenate([uv, np.stack([rng.uniform(-4, 324, 40), rng.uniform(-4, 244, 40)], 1).astype(np.float32)])
    rl, cl = oracle.pyramid_build(ref, 3), oracle.pyramid_build(cur, 3)
    for half, lum, single in [(6, False, False), (4, True, False), (7, False, True)]:
        p = po.make_params(variant, method, half=half, half_col=half + 1, max_points=1000, luminance=lum,
                           predict=(0.9995, -0.03, 0.03, 0.9995) if single else (1, 0, 0, 1))
        a = reflib.klt_track(p, rl, cl, uv, single_level=single)
        b = oracle.klt_track(p, rl, cl, uv, single_level=single)
        assert a[0] == b[0] and (a[2] == b[2]).all() and bits_equal(a[1], b[1])


def test_klt_entry_semantics(oracle):
    """optical_flow.cpp:8-19 + basic_klt.cpp:9,15: empty input, size mismatches, skipped statuses, the point cap."""
    ref, cur, uv, _ = S.make_pair(120, 160, 30, pair_id=5, border=10)
    rl, cl = oracle.pyramid_build(ref, 2), oracle.pyramid_build(cur, 2)
    p = po.make_params("basic", "fast", half=4, max_points=10)
    ok, _, _ = oracle.klt_track(p, rl, cl, np.zeros((0, 2), np.float32))
    assert not ok
    # wrong-sized cur/status are reset; features beyond the cap keep cur = ref and kNotTracked
    ok, out, st = oracle.klt_track(p, rl, cl, uv, cur_uv=uv[:5] + 1, status=np.full(3, 4, np.uint8))
    assert ok and bits_equal(out[10:], uv[10:]) and (st[10:] == 0).all() and (st[:10] != 0).any()
    # entries > kTracked are left untouched
    st_in = np.zeros(30, np.uint8)
    st_in[2], st_in[4] = 3, 2
    pred = uv + np.float32(0.5)
    ok, out, st = oracle.klt_track(p, rl, cl, uv, cur_uv=pred, status=st_in)
    assert st[2] == 3 and st[4] == 2 and bits_equal(out[[2, 4]], pred[[2, 4]])


def test_klt_recovers_known_shift(oracle):
    """Property: a pure integer translation is recovered by every variant (status kTracked, error << 1 px)."""
    ref = S.make_image(200, 260, seed=77)
    cur = np.roll(np.roll(ref, 3, axis=0), -2, axis=1)
    uv = S.detect_features(ref, 60, seed=1, border=40, border_fraction=0.0)
    rl, cl = oracle.pyramid_build(ref, 3), oracle.pyramid_build(cur, 3)
    for v in ("basic", "affine", "lssd"):
        for m in ("inverse", "direct", "fast"):
            ok, out, st = oracle.klt_track(po.make_params(v, m, half=7), rl, cl, uv)
            good = st == 1
            assert good.mean() > 0.8, (v, m, good.mean())
            err = np.abs(out[good] - (uv[good] + np.array([-2.0, 3.0], np.float32)))
            assert np.median(err) < 0.15, (v, m, np.median(err))


def test_ldlt_restatement(oracle):
    """SURVEY App. A.5 known answers: zero matrix -> 0, rank-1 [[4,2],[2,1]] b=(2,1) -> (0.5, 0); random SPD residuals."""
    import ctypes as C
    f = oracle.lib.ftko_ldlt_solve
    f.restype = None

    def solve(A, b):
        A = np.ascontiguousarray(A, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        x = np.zeros(len(b), np.float32)
        f(C.c_int32(len(b)), A.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p))
        return x

    assert (solve(np.zeros((2, 2)), [1, 2]) == 0).all()
    assert np.allclose(solve([[4, 2], [2, 1]], [2, 1]), [0.5, 0.0])
    rng = np.random.default_rng(0)
    for n in (2, 3, 6):
        for _ in range(200):
            M = rng.normal(size=(n, n + 2)).astype(np.float32)
            A = (M @ M.T).astype(np.float32)
            b = rng.normal(size=n).astype(np.float32)
            x = solve(A, b)
            assert np.linalg.norm(A.astype(np.float64) @ x - b) <= 2e-3 * (1 + np.linalg.norm(b)) * np.linalg.cond(A.astype(np.float64)) ** 0.5


def test_matchers_match_golden(oracle, matcher_golden):
    m = matcher_golden
    ok, idx = oracle.match_brief_force(m["brief_ref"], m["brief_cur"], 60.0)
    assert ok and (idx == m["brief_force_idx"]).all()
    ok, idx = oracle.match_brief_nearby(m["brief_ref"], m["brief_cur"], m["brief_pred"], m["brief_pos"], 50, 50, 60.0)
    assert ok and (idx == m["brief_nearby_idx"]).all()
    ok, muv, mst = oracle.match_brief_nearby_uv(m["brief_ref"], m["brief_cur"], m["brief_pred"], m["brief_pos"], 50, 50, 60.0)
    assert (mst == m["brief_nearby_st"]).all() and bits_equal(muv[mst == 1], m["brief_nearby_uv"][mst == 1])
    ok, idx = oracle.match_cosine_force(m["float_ref"], m["float_cur"], 0.1)
    assert ok and (idx == m["float_force_idx"]).all()
    ok, idx = oracle.match_cosine_nearby(m["float_ref"], m["float_cur"], m["float_pred"], m["float_pos"], 50, 50, 0.3)
    assert ok and (idx == m["float_nearby_idx"]).all()


def test_matchers_vs_reference_build(oracle, reflib):
    rb, cb, pred, pos, _ = S.make_brief_sets(150, 170, seed=21)
    for args in [(rb, cb, 60.0), (rb, cb, 200.0), (rb[:0], cb, 60.0)]:
        if args[0].shape[0] == 0:
            continue
        assert (oracle.match_brief_force(*args)[1] == reflib.match_brief_force(*args)[1]).all()
    a = oracle.match_brief_nearby(rb, cb, pred, pos, 30, 45, 70.0)
    b = reflib.match_brief_nearby(rb, cb, pred, pos, 30, 45, 70.0)
    assert a[0] == b[0] and (a[1] == b[1]).all()
    rf, cf = S.make_float_sets(90, 100, dim=64, seed=2)
    assert (oracle.match_cosine_force(rf, cf, 0.2)[1] == reflib.match_cosine_force(rf, cf, 0.2)[1]).all()


def test_matcher_semantics(oracle):
    """Strict '<' keeps the lowest j on ties; nothing matches at the default max distance 0; preset indices survive."""
    rng = np.random.default_rng(4)
    cur = rng.integers(0, 2, (6, 64), dtype=np.uint8)
    cur[4] = cur[1]
    ref = cur[[1, 3]].copy()
    ok, idx = oracle.match_brief_force(ref, cur, 10.0)
    assert ok and list(idx) == [1, 3]
    ok, idx = oracle.match_brief_force(ref, cur, 0.0)
    assert ok and list(idx) == [-1, -1]
    ok, idx = oracle.match_brief_force(ref, cur, 0.0, idx=np.array([5, 2], np.int32))
    assert list(idx) == [5, 2]
    ok, _ = oracle.match_brief_force(ref, cur[:0].reshape(0, 64), 10.0)
    assert not ok
    # nearby: the window gate excludes the exact duplicate, a farther candidate inside the window wins
    pos = np.array([[0, 0], [100, 100], [10, 10], [50, 50], [12, 12], [300, 300]], np.float32)
    pred = np.array([[11, 11], [52, 52]], np.float32)
    ok, idx = oracle.match_brief_nearby(ref, cur, pred, pos, 5, 5, 64.0)
    assert list(idx) == [4, 3]


def _mutual_scores_python(scores, min_score):
    """Pure-Python transcription of nn_feature_matcher.cpp:187-214 (small cases only)."""
    n_ref, n_cur = scores.shape
    col_best = []
    for j in range(n_cur):
        best, best_i = scores[0, j], 0
        for i in range(1, n_ref):
            if scores[i, j] > best:
                best, best_i = scores[i, j], i
        col_best.append(best_i)
    idx = np.full(n_ref, -1, np.int32)
    for i in range(n_ref):
        best, best_j = scores[i, 0], 0
        for j in range(1, n_cur):
            if scores[i, j] > best:
                best, best_j = scores[i, j], j
        if best < min_score or col_best[best_j] != i:
            continue
        idx[i] = best_j
    return idx


def test_mutual_scores_restatement(oracle):
    """The C restatement of the score-matrix post-processing against a literal Python loop, ties / NaN / -inf included."""
    rng = np.random.default_rng(77)
    for n_ref, n_cur in [(1, 1), (7, 5), (40, 61), (64, 64)]:
        s = rng.normal(-4, 3, (n_ref, n_cur)).astype(np.float32)
        s[rng.random(s.shape) < 0.1] = np.float32(-1.5)  # ties
        if n_ref > 5:
            s[2, :] = -np.inf
            s[3, 0] = np.nan
            s[0, 2] = np.nan
            s[5, 3] = np.nan
        ok, idx = oracle.mutual_scores(s, -3.0)
        assert ok
        with np.errstate(invalid="ignore"):
            assert np.array_equal(idx, _mutual_scores_python(s, np.float32(-3.0)))
    ok, _ = oracle.mutual_scores(np.zeros((3, 0), np.float32), -3.0)
    assert not ok


def test_mutual_scores_restatement_pinned_to_reference_match(oracle, reflib):
    """The restatement against the REFERENCE's own NNFeatureMatcher::Match (nn_feature_matcher.cpp:150-219 compiled in place; ONNX
    Runtime replaced by the stub of oracle/shim/onnx_run_time.h, whose session returns the injected score matrix): ties, NaN, -inf,
    every threshold.  This pins the mutual-score row of SURVEY 8(f) to the reference."""
    rng = np.random.default_rng(78)
    for n_ref, n_cur in [(1, 1), (5, 7), (40, 61), (64, 64), (300, 300), (257, 1031)]:
        s = rng.normal(-4, 3, (n_ref, n_cur)).astype(np.float32)
        s[rng.random(s.shape) < 0.1] = np.float32(-1.5)  # ties
        for i in range(min(n_ref, n_cur) // 2):  # planted mutual maxima
            s[i, (i * 7) % n_cur] = np.float32(1.0 + 0.01 * i)
        if n_ref > 5:
            s[2, :] = -np.inf
            s[3, 0] = np.nan
            s[0, 2] = np.nan
            s[5, 3] = np.nan
            s[4, 4], s[4, 6] = -0.0, 0.0
        for thr in (-3.0, -1.0, -100.0, 0.5):
            ok_r, exp = reflib.nn_match_scores(s, thr)
            ok_o, got = oracle.mutual_scores(s, thr)
            assert ok_r and ok_o and np.array_equal(got, exp), (n_ref, n_cur, thr, np.nonzero(got != exp)[0][:10])
        assert (exp >= 0).sum() >= min(1, n_ref // 8)


def test_reference_match_pairs_branch(reflib):
    """The fused-model branch of Match (nn_feature_matcher.cpp:160-178): index pairs scattered with bounds checks -- what the python
    NNFeatureMatcher.MatchPairs mirrors on the host (no kernel: n <= a few hundred pairs)."""
    import feature_tracker_b200.api as api
    pairs = np.array([[0, 3], [2, 1], [5, 9], [-1, 2], [3, 40], [99, 1], [4, -2], [2, 7]], np.int64)
    ok, exp = reflib.nn_match_pairs(pairs, 8, 10)
    assert ok and np.array_equal(exp, api.nn_match_pairs_host(pairs, 8, 10))


@pytest.mark.parametrize("levels,half,max_points", [(4, 6, 500), (3, 4, 40), (1, 7, 500), (5, 6, 25)])
def test_direct_method_restatement_vs_reference_build(oracle, reflib, levels, half, max_points):
    """the C restatement of DirectMethod::TrackFeatures equals the reference's own
    direct_method_tracker.cpp compiled in place, bit for bit (pose, projected positions, status)."""
    from feature_tracker_b200 import synthetic as S
    ref, cur, uv, K, pts = S.make_direct_method_scene(240, 320, 60, pair_id=40 + levels, border=10)
    pts = pts.copy()
    pts[3, 2] = -1.0     # behind the reference camera: skipped (:129)
    pts[7, 2] = 1e-7     # below kZeroFloat
    pts[11] = (50.0, 0.0, 1.0)  # projects far outside the image: status kOutside, contributes nothing
    rl, cl = oracle.pyramid_build(ref, levels), oracle.pyramid_build(cur, levels)
    q0 = np.array([0.9999, 0.002, -0.003, 0.004], np.float32)
    p0 = np.array([0.01, -0.02, 0.005], np.float32)
    for method in ("direct", "inverse", "fast"):
        prm = po.make_direct_params(half=half, max_points=max_points, method=method)
        for kwargs in ({}, {"cur_uv": uv + 0.5, "status": np.full(60, 2, np.uint8)}):
            a = oracle.direct_method_track(prm, rl, cl, K, pts, uv, q0, p0, **kwargs)
            b = reflib.direct_method_track(prm, rl, cl, K, pts, uv, q0, p0, **kwargs)
            assert a[0] and b[0]
            for x, y, name in zip(a[1:], b[1:], ("cur_uv", "q_rc", "p_rc", "status")):
                if x.dtype == np.uint8:
                    assert np.array_equal(x, y), (method, name)
                else:
                    assert bits_equal(x, y), (method, name, x[:4], y[:4])
    # the pose actually moved and most projections stay inside
    prm = po.make_direct_params(half=half, max_points=max_points)
    ok, cur_uv, q, p, st = oracle.direct_method_track(prm, rl, cl, K, pts, uv, [1, 0, 0, 0], [0, 0, 0])
    assert ok and (st == 1).sum() > 40 and np.abs(p).max() > 1e-3
    assert not oracle.direct_method_track(prm, rl, cl, K, np.zeros((0, 3)), np.zeros((0, 2)), [1, 0, 0, 0], [0, 0, 0])[0]


def test_direct_method_matches_golden(oracle):
    """The restatement against outputs of the reference itself (direct_method_tracker.cpp compiled in place) on the reference's
    own KITTI fixture, two frames with pose / positions / status carried over like test/test_direct_method.cpp:69-86."""
    g = dict(np.load(os.path.join(GOLDEN, "direct_method_golden.npz")))
    levels = int(g["levels"])
    ll = oracle.pyramid_build(g["left"], levels)
    q, p, cur_uv, st = np.array([1, 0, 0, 0], np.float32), np.zeros(3, np.float32), None, None
    for i in (1, 2):
        ok, cur_uv, q, p, st = oracle.direct_method_track(po.make_direct_params(), ll, oracle.pyramid_build(g[f"cur{i}"], levels), g["K"], g["p_c_in_ref"],
                                                          g["uv"], q, p, cur_uv=cur_uv, status=st)
        assert ok and bits_equal(q, g[f"q_{i}"]) and bits_equal(p, g[f"p_{i}"]) and bits_equal(cur_uv, g[f"uv_{i}"]) and np.array_equal(st, g[f"st_{i}"])
    assert 0.5 < p[2] / 2 < 1.0  # the car drives forward ~0.7 m per frame


def test_direct_method_world_frame_algebra(oracle, reflib):
    """The host-side quaternion algebra of the world-frame overload (feature_tracker_b200/quat.py, used by the product's
    DirectMethod.TrackFeaturesWorld) around the camera-frame core == the reference's own world-frame overload
    (direct_method_tracker.cpp:8-39), bit for bit."""
    from feature_tracker_b200 import quat
    ref, cur, uv, K, pts = S.make_direct_method_scene(240, 320, 70, pair_id=90, border=10)
    ref_q = np.array([0.96, 0.1, -0.2, 0.15], np.float32)
    ref_q /= np.float32(np.linalg.norm(ref_q))
    ref_p = np.array([1.5, -0.7, 3.0], np.float32)
    p_w = quat.rotate(ref_q, pts) + ref_p
    rl, cl = oracle.pyramid_build(ref, 4), oracle.pyramid_build(cur, 4)
    exp = reflib.direct_method_track_world(po.make_direct_params(), rl, cl, K, ref_q, ref_p, p_w, uv, ref_q, ref_p)
    q_cw = quat.inverse(ref_q)
    ok, cu, q, p, st = oracle.direct_method_track(po.make_direct_params(), rl, cl, K, quat.rotate(q_cw, p_w - ref_p), uv, quat.multiply(q_cw, ref_q),
                                                  quat.rotate(q_cw, ref_p - ref_p))
    assert ok and exp[0]
    assert bits_equal(quat.multiply(ref_q, q), exp[2]) and bits_equal(quat.rotate(ref_q, p) + ref_p, exp[3])
    assert bits_equal(cu, exp[1]) and np.array_equal(st, exp[4])


@pytest.mark.parametrize("levels,single", [(3, False), (1, True), (4, False)])
def test_dense_flow_restatement_vs_reference_build(oracle, reflib, levels, single):
    """the C restatement of DenseOpticalFlow::Track equals the reference's own dense_optical_flow.cpp compiled in
    place, bit for bit, for several Gaussian window sizes, both overloads, with and without an initial flow."""
    ref, cur, _, _ = S.make_pair(96, 131, 10, pair_id=5)
    rl, cl = oracle.pyramid_build(ref, levels), oracle.pyramid_build(cur, levels)
    for half in (2, 1, 0, 3):
        prm = po.make_dense_flow_params(half=half)
        flows = [None]
        if single:
            rng = np.random.default_rng(half)
            flows.append((rng.normal(0, 1, ref.shape).astype(np.float32), rng.normal(0, 1, ref.shape).astype(np.float32)))
        for flow in flows:
            a = oracle.dense_flow_track(prm, rl, cl, single_level=single, flow=flow)
            b = reflib.dense_flow_track(prm, rl, cl, single_level=single, flow=flow)
            assert a[0] and b[0] and bits_equal(a[1], b[1]) and bits_equal(a[2], b[2]), (half, flow is not None)
    assert np.abs(a[1]).mean() > 0.3  # there is motion in the pair


def _check_dense_flow_golden(track, g_flow, g_img):
    """`track(levels_ref, levels_cur, params, single)` -> (ok, flow_row, flow_col); compared with the reference's committed outputs."""
    import hashlib
    levels = int(g_img["levels"])
    rl = [g_img["ref"]] + [g_img[f"ref_l{l}"] for l in range(1, levels)]
    cl = [g_img["cur"]] + [g_img[f"cur_l{l}"] for l in range(1, levels)]
    for prm, single, key in ((po.make_dense_flow_params(), False, ""), (po.make_dense_flow_params(half=1, max_iter=4), True, "single_h1_")):
        ok, fr, fc = track(rl[:1] if single else rl, cl[:1] if single else cl, prm, single)
        assert ok
        row_key, col_key = ("flow_row_8", "flow_col_8") if not single else ("single_h1_row_8", "single_h1_col_8")
        assert bits_equal(fr[::8, ::8], g_flow[row_key]) and bits_equal(fc[::8, ::8], g_flow[col_key])
        digest = np.frombuffer(hashlib.sha256(fr.tobytes() + fc.tobytes()).digest(), np.uint8)
        assert np.array_equal(digest, g_flow[key + "sha256"])


def test_dense_flow_matches_golden(oracle, euroc_golden):
    """The restatement against outputs of the reference itself on its own EuRoC fixture pair (every 8th flow vector + a SHA-256 of
    the full field)."""
    g = dict(np.load(os.path.join(GOLDEN, "dense_flow_golden.npz")))
    _check_dense_flow_golden(lambda rl, cl, prm, single: oracle.dense_flow_track(prm, rl, cl, single_level=single), g, euroc_golden)

