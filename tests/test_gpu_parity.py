"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle on identical seeded inputs and
against the committed golden vectors (produced by the reference's own sources).

Bar (BASELINE.json north_star): positions within 1e-3 px with identical status, Hamming distances / match indices
bit-exact.  This implementation reproduces the reference's fp32 arithmetic and summation order, so the tests assert the
stronger property: positions BIT-IDENTICAL, status identical, indices identical."""
import numpy as np
import pytest

import feature_tracker_b200 as ft
from conftest import bits_equal
from feature_tracker_b200 import synthetic as S
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

VARIANTS = {"basic": ft.OpticalFlowBasicKlt, "affine": ft.OpticalFlowAffineKlt, "lssd": ft.OpticalFlowLssdKlt}
METHODS = {"inverse": ft.OpticalFlowMethod.kInverse, "direct": ft.OpticalFlowMethod.kDirect, "fast": ft.OpticalFlowMethod.kFast}
KLT_COMBOS = [(v, m, h) for v in ("basic", "affine", "lssd") for m in ("inverse", "direct", "fast") for h in (6, 7, 10)]
POS_TOL_PX = 1e-3  # the stated tolerance; the assertions below demand exact equality


def make_tracker(ctx, variant, method, half, half_col=None, max_points=500, predict=None, luminance=False):
    klt = VARIANTS[variant](ctx)
    o = klt.options()
    o.kPatchRowHalfSize = half
    o.kPatchColHalfSize = half if half_col is None else half_col
    o.kMethod = METHODS[method]
    o.kMaxTrackPointsNumber = max_points
    if predict is not None:
        klt._predict = np.array(predict, np.float32).reshape(2, 2)
    if variant == "lssd":
        klt.consider_patch_luminance = luminance
    return klt


def upload_levels(ctx, levels_list):
    """Builds a device pyramid batch from host-built levels (one list of level arrays per image), verbatim."""
    rows, cols = levels_list[0][0].shape
    pyr = ft.ImagePyramidBatch(ctx, rows, cols, len(levels_list[0]), len(levels_list))
    for i, lv in enumerate(levels_list):
        for l, a in enumerate(lv):
            pyr.SetLevel(i, l, a)
    return pyr


def assert_same(tag, got, exp):
    ok_g, uv_g, st_g = got
    ok_e, uv_e, st_e = exp
    assert ok_g == ok_e, tag
    bad_st = np.nonzero(st_g != st_e)[0]
    assert bad_st.size == 0, f"{tag}: status differs at {bad_st[:10]} gpu={st_g[bad_st[:10]]} oracle={st_e[bad_st[:10]]}"
    with np.errstate(invalid="ignore"):  # inf - inf on untouched garbage positions: the bit comparison below decides
        d = np.abs(uv_g.astype(np.float64) - uv_e.astype(np.float64))
    d = np.where(np.isnan(d), 0 if np.array_equal(np.isnan(uv_g), np.isnan(uv_e)) else np.inf, d)
    assert d.max() <= POS_TOL_PX, f"{tag}: max position error {d.max()} px"
    assert bits_equal(uv_g, uv_e), f"{tag}: positions within tolerance ({d.max()} px) but not bit-identical"


# ---- pyramid -------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,levels", [((480, 752), 4), ((720, 1280), 4), ((37, 53), 4), ((64, 64), 1), ((100, 131), 6), ((9, 9), 3)])
def test_pyramid_bit_exact(ctx, oracle, shape, levels):
    rng = np.random.default_rng(shape[0] * 7 + levels)
    imgs = rng.integers(0, 256, (3,) + shape, dtype=np.uint8)
    pyr = ft.ImagePyramidBatch(ctx, shape[0], shape[1], levels, 3)
    pyr.SetRawImages(imgs)
    pyr.CreateImagePyramid()
    for i in range(3):
        exp = oracle.pyramid_build(imgs[i], levels)
        for l in range(levels):
            assert (pyr.GetLevel(i, l) == exp[l]).all(), (i, l)


def test_pyramid_matches_golden(ctx, euroc_golden):
    g = euroc_golden
    pyr = ft.ImagePyramidBatch(ctx, 480, 752, 4, 2)
    pyr.SetRawImages(np.stack([g["ref"], g["cur"]]))
    pyr.CreateImagePyramid()
    for l in range(1, 4):
        assert (pyr.GetLevel(0, l) == g[f"ref_l{l}"]).all() and (pyr.GetLevel(1, l) == g[f"cur_l{l}"]).all()


# ---- KLT vs golden (reference-produced) ------------------------------------------------------------------------------
@pytest.mark.parametrize("variant,method,half", KLT_COMBOS)
def test_klt_matches_golden(ctx, euroc_golden, variant, method, half):
    g = euroc_golden
    pyr = ft.ImagePyramidBatch(ctx, 480, 752, 4, 2)
    pyr.SetRawImages(np.stack([g["ref"], g["cur"]]))
    pyr.CreateImagePyramid()
    klt = make_tracker(ctx, variant, method, half)
    got = klt.TrackFeatures(pyr, pyr, g["pts"], ref_image=0, cur_image=1)
    key = f"{variant}_{method}_h{half}"
    assert_same(key, got, (True, g[key + "_uv"], g[key + "_st"]))


def test_klt_golden_extras(ctx, euroc_golden):
    g = euroc_golden
    pyr = ft.ImagePyramidBatch(ctx, 480, 752, 4, 2)
    pyr.SetRawImages(np.stack([g["ref"], g["cur"]]))
    pyr.CreateImagePyramid()
    klt = make_tracker(ctx, "lssd", "fast", 6, luminance=True)
    assert_same("lssd lum", klt.TrackFeatures(pyr, pyr, g["pts"], ref_image=0, cur_image=1), (True, g["lssd_fast_h6_lum_uv"], g["lssd_fast_h6_lum_st"]))
    pred = g["pts"] + np.float32(3.0)
    for v in ("basic", "affine", "lssd"):
        klt = make_tracker(ctx, v, "fast", 6, predict=(0.9995, -0.03, 0.03, 0.9995))
        got = klt.TrackFeatures(pyr, pyr, g["pts"], cur_pixel_uv=pred, single_level=True, ref_image=0, cur_image=1)
        assert_same(v + " single", got, (True, g[f"{v}_fast_h6_single_uv"], g[f"{v}_fast_h6_single_st"]))


# ---- KLT vs oracle on seeded synthetic pairs (interior + border + far-outside features) -------------------------------
@pytest.mark.parametrize("variant", ["basic", "affine", "lssd"])
@pytest.mark.parametrize("method", ["inverse", "direct", "fast"])
def test_klt_synthetic_vs_oracle(ctx, oracle, variant, method):
    ref, cur, uv, _ = S.make_pair(240, 320, 150, pair_id=11, border=12)
    rng = np.random.default_rng(5)
    uv = np.concatenate([uv, np.stack([rng.uniform(-6, 326, 60), rng.uniform(-6, 246, 60)], 1).astype(np.float32)])
    rl, cl = oracle.pyramid_build(ref, 3), oracle.pyramid_build(cur, 3)
    pyr = upload_levels(ctx, [rl, cl])
    for half, half_col, lum, single in [(6, 6, False, False), (4, 5, True, False), (7, 7, False, True), (2, 9, False, False), (5, 5, False, False),
                                        (4, 4, False, True), (6, 6, False, True), (7, 7, False, False)]:
        predict = (0.9995, -0.03, 0.03, 0.9995) if single else (1, 0, 0, 1)
        p = po.make_params(variant, method, half=half, half_col=half_col, max_points=1000, luminance=lum, predict=predict)
        exp = oracle.klt_track(p, rl, cl, uv, single_level=single)
        klt = make_tracker(ctx, variant, method, half, half_col, max_points=1000, predict=predict, luminance=lum)
        got = klt.TrackFeatures(pyr, pyr, uv, single_level=single, ref_image=0, cur_image=1)
        assert_same(f"{variant}/{method}/h{half}x{half_col}/lum{lum}/single{single}", got, exp)


def test_klt_entry_semantics(ctx, oracle):
    ref, cur, uv, _ = S.make_pair(120, 160, 30, pair_id=5, border=10)
    rl, cl = oracle.pyramid_build(ref, 2), oracle.pyramid_build(cur, 2)
    pyr = upload_levels(ctx, [rl, cl])
    klt = make_tracker(ctx, "basic", "fast", 4, max_points=10)
    p = po.make_params("basic", "fast", half=4, max_points=10)
    ok, _, _ = klt.TrackFeatures(pyr, pyr, np.zeros((0, 2), np.float32), ref_image=0, cur_image=1)
    assert not ok  # optical_flow.cpp:8
    assert_same("cap", klt.TrackFeatures(pyr, pyr, uv, cur_pixel_uv=uv[:5] + 1, status=np.full(3, 4, np.uint8), ref_image=0, cur_image=1),
                oracle.klt_track(p, rl, cl, uv, cur_uv=uv[:5] + 1, status=np.full(3, 4, np.uint8)))
    st_in = np.zeros(30, np.uint8)
    st_in[2], st_in[4] = 3, 2
    pred = uv + np.float32(0.5)
    assert_same("skip", klt.TrackFeatures(pyr, pyr, uv, cur_pixel_uv=pred, status=st_in, ref_image=0, cur_image=1),
                oracle.klt_track(p, rl, cl, uv, cur_uv=pred, status=st_in))
    # kSse / kNeon take the reference's `default:` branch, i.e. behave exactly like kFast (basic_klt.cpp:31-34)
    for half in (4, 7):
        fast = make_tracker(ctx, "basic", "fast", half)
        sse = make_tracker(ctx, "basic", "fast", half)
        sse.options().kMethod = ft.OpticalFlowMethod.kSse
        a, b = fast.TrackFeatures(pyr, pyr, uv, ref_image=0, cur_image=1), sse.TrackFeatures(pyr, pyr, uv, ref_image=0, cur_image=1)
        assert (a[2] == b[2]).all() and bits_equal(a[1], b[1])
    # level mismatch -> false (optical_flow.cpp:9)
    other = ft.ImagePyramidBatch(ctx, 120, 160, 3, 1)
    ok, _, _ = klt.TrackFeatures(pyr, other, uv, ref_image=0, cur_image=0)
    assert not ok


@pytest.mark.gpu
@pytest.mark.parametrize("variant,method,half", [("basic", "inverse", 5), ("basic", "fast", 9), ("affine", "direct", 6), ("affine", "fast", 6),
                                                  ("affine", "inverse", 4), ("lssd", "inverse", 6), ("lssd", "fast", 5)])
def test_klt_untracked_neighbours_with_garbage_positions(ctx, oracle, variant, method, half):
    """Several features share a warp and run one instruction stream: a feature that is not tracked (entry status above kTracked, or beyond
    kMaxTrackPointsNumber) keeps its partner company with its results discarded.  Its positions may be anything -- NaN, infinities, 1e30 --
    and must neither disturb the partner nor be touched themselves (basic_klt.cpp:9,12,15)."""
    ref, cur, uv, _ = S.make_pair(150, 200, 61, pair_id=41, border=14)  # odd count: the last warp has a group without a feature
    rl, cl = oracle.pyramid_build(ref, 3), oracle.pyramid_build(cur, 3)
    pyr = upload_levels(ctx, [rl, cl])
    uv = uv.copy()
    pred = uv + np.float32(0.75)
    st_in = np.zeros(61, np.uint8)
    garbage = [np.nan, np.inf, -np.inf, 1e30, -1e30, 3e9, -7.5]
    for i in range(1, 61, 2):
        st_in[i] = 2 + (i // 2) % 3  # kLargeResidual / kOutside / kNumericError: never re-tracked
        g = np.float32(garbage[(i // 2) % len(garbage)])
        if i % 4 == 1:
            pred[i] = (g, -g)
        else:
            uv[i] = (g, np.float32(12.0))
            pred[i, 1] = g
    klt = make_tracker(ctx, variant, method, half, max_points=57)  # features 57..60 are beyond the cap as well
    p = po.make_params(variant, method, half=half, max_points=57)
    assert_same(f"{variant}/{method}", klt.TrackFeatures(pyr, pyr, uv, cur_pixel_uv=pred, status=st_in, ref_image=0, cur_image=1),
                oracle.klt_track(p, rl, cl, uv, cur_uv=pred, status=st_in))


def test_klt_batch_of_pairs(ctx, oracle):
    """Many frame pairs in one call (the sharding unit): ragged feature counts incl. an empty pair, device-built pyramids."""
    n_pairs, rows, cols, levels = 5, 120, 160, 3
    counts = [40, 0, 17, 64, 1]
    refs, curs, uvs = [], [], []
    for p in range(n_pairs):
        r, c, uv, _ = S.make_pair(rows, cols, max(counts[p], 1), pair_id=40 + p, border=10)
        refs.append(r), curs.append(c), uvs.append(uv[:counts[p]])
    pyr = ft.ImagePyramidBatch(ctx, rows, cols, levels, 2 * n_pairs)
    pyr.SetRawImages(np.stack(refs + curs))
    pyr.CreateImagePyramid()
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    all_uv = np.concatenate(uvs)
    for variant, method, half in [("basic", "inverse", 7), ("basic", "fast", 6), ("affine", "fast", 6), ("lssd", "inverse", 5)]:
        klt = make_tracker(ctx, variant, method, half, max_points=50)
        ok, cur_uv, st = klt.TrackFeaturesBatch(pyr, pyr, offsets, all_uv, ref_image=np.arange(n_pairs), cur_image=np.arange(n_pairs) + n_pairs)
        assert ok
        prm = po.make_params(variant, method, half=half, max_points=50)
        for p in range(n_pairs):
            if counts[p] == 0:
                continue
            exp = oracle.klt_track(prm, oracle.pyramid_build(refs[p], levels), oracle.pyramid_build(curs[p], levels), uvs[p])
            sl = slice(offsets[p], offsets[p + 1])
            assert_same(f"{variant}/{method} pair {p}", (True, cur_uv[sl], st[sl]), exp)


def test_track_image_pairs_pipelined(ctx, oracle):
    """ftk_track_image_pairs: host images in, chunked two-stream pipeline (40 pairs -> 8 chunks of 5), ragged feature counts."""
    n_pairs, rows, cols, levels = 40, 96, 128, 3
    rng = np.random.default_rng(12)
    uniq = [S.make_pair(rows, cols, 24, pair_id=80 + p, border=8) for p in range(5)]
    counts = rng.integers(0, 25, n_pairs)
    counts[3] = 0
    refs = np.stack([uniq[p % 5][0] for p in range(n_pairs)])
    curs = np.stack([uniq[p % 5][1] for p in range(n_pairs)])
    uvs = [uniq[p % 5][2][:counts[p]] for p in range(n_pairs)]
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    all_uv = np.concatenate(uvs)
    for variant, method, half in [("basic", "inverse", 7), ("basic", "fast", 6), ("lssd", "fast", 4)]:
        klt = make_tracker(ctx, variant, method, half)
        ok, cur_uv, st = klt.TrackImagePairs(levels, refs, curs, offsets, all_uv)
        assert ok
        prm = po.make_params(variant, method, half=half)
        cache = {}
        for p in range(n_pairs):
            if counts[p] == 0:
                continue
            u = p % 5
            if u not in cache:
                cache[u] = (oracle.pyramid_build(uniq[u][0], levels), oracle.pyramid_build(uniq[u][1], levels))
            exp = oracle.klt_track(prm, cache[u][0], cache[u][1], uvs[p])
            sl = slice(offsets[p], offsets[p + 1])
            assert_same(f"pipelined {variant}/{method} pair {p}", (True, cur_uv[sl], st[sl]), exp)
    ok, _, _ = klt.TrackImagePairs(levels, refs, curs, np.zeros(n_pairs + 1, np.int32), np.zeros((0, 2), np.float32))
    assert not ok


def test_track_image_pairs_multi(ctx, oracle):
    """ftk_track_image_pairs_multi: several trackers per upload; every tracker's slice equals its own ftk_track_image_pairs call."""
    n_pairs, rows, cols, levels = 20, 96, 128, 3
    uniq = [S.make_pair(rows, cols, 20, pair_id=180 + p, border=8) for p in range(4)]
    refs = np.stack([uniq[p % 4][0] for p in range(n_pairs)])
    curs = np.stack([uniq[p % 4][1] for p in range(n_pairs)])
    uvs = [uniq[p % 4][2] for p in range(n_pairs)]
    offsets = np.concatenate([[0], np.cumsum([u.shape[0] for u in uvs])]).astype(np.int32)
    all_uv = np.concatenate(uvs)
    trackers = [make_tracker(ctx, "affine", "direct", 6), make_tracker(ctx, "affine", "fast", 6), make_tracker(ctx, "basic", "inverse", 7)]
    trackers[2].forward_backward_max_error = 0.5
    ok, cur_uv, st = ft.OpticalFlow.TrackImagePairsMulti(trackers, levels, refs, curs, offsets, all_uv)
    assert ok and cur_uv.shape == (3, all_uv.shape[0], 2)
    for k, t in enumerate(trackers):
        ok1, uv1, st1 = t.TrackImagePairs(levels, refs, curs, offsets, all_uv)
        assert ok1 and bits_equal(cur_uv[k], uv1) and np.array_equal(st[k], st1), k
    prm = po.make_params("affine", "direct", half=6)
    exp = oracle.klt_track(prm, oracle.pyramid_build(uniq[1][0], levels), oracle.pyramid_build(uniq[1][1], levels), uvs[1])
    sl = slice(offsets[1], offsets[2])
    assert_same("multi affine direct pair 1", (True, cur_uv[0][sl], st[0][sl]), exp)


def test_klt_temporal_sequence_shares_pyramids(ctx, oracle):
    """in a sequence the cur frame of pair k is the ref frame of pair k+1.  One pyramid per frame, pairs
    (k, k+1) addressed through the ref_image / cur_image maps; results equal tracking each pair on its own."""
    rows, cols, levels, n_frames = 120, 160, 3, 4
    base = S.make_image(rows, cols, seed=500)
    frames = [base]
    for k in range(1, n_frames):
        frames.append(S.warp_image(frames[-1], seed=600 + k, max_shift=3.0, max_rot_deg=1.0)[0])
    pyr = ft.ImagePyramidBatch(ctx, rows, cols, levels, n_frames)
    pyr.SetRawImages(np.stack(frames))
    pyr.CreateImagePyramid()
    uvs = [S.detect_features(frames[k], 30, seed=k, border=12) for k in range(n_frames - 1)]
    offsets = np.arange(n_frames) * 30
    klt = make_tracker(ctx, "basic", "inverse", 6)
    ok, cur_uv, st = klt.TrackFeaturesBatch(pyr, pyr, offsets.astype(np.int32), np.concatenate(uvs), ref_image=np.arange(n_frames - 1),
                                            cur_image=np.arange(1, n_frames))
    assert ok
    prm = po.make_params("basic", "inverse", half=6)
    lv = [oracle.pyramid_build(f, levels) for f in frames]
    for k in range(n_frames - 1):
        exp = oracle.klt_track(prm, lv[k], lv[k + 1], uvs[k])
        assert_same(f"sequence pair {k}", (True, cur_uv[30 * k:30 * (k + 1)], st[30 * k:30 * (k + 1)]), exp)


def forward_backward_oracle(oracle, prm, ref_levels, cur_levels, uv, max_error):
    """The composition ftk_klt_params::forward_backward_max_error documents, out of two reference TrackFeatures calls."""
    ok, fwd_uv, fwd_st = oracle.klt_track(prm, ref_levels, cur_levels, uv)
    _, back_uv, back_st = oracle.klt_track(prm, cur_levels, ref_levels, fwd_uv, cur_uv=uv, status=fwd_st)
    d = back_uv - uv.astype(np.float32)
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32)
    good = (back_st == 1) & (d2 <= np.float32(max_error) * np.float32(max_error))
    st = fwd_st.copy()
    st[(fwd_st == 1) & ~good] = 2
    return ok, fwd_uv, st


def test_track_image_sequence_pipelined(ctx, oracle):
    """ftk_track_image_sequence: 41 host frames -> 40 pairs in 8 chunks; every frame uploaded once; equals pair-by-pair tracking."""
    rows, cols, levels, n_frames = 96, 128, 3, 41
    base = S.make_image(rows, cols, seed=900)
    frames = [base]
    for k in range(1, 6):
        frames.append(S.warp_image(frames[-1], seed=910 + k, max_shift=2.5, max_rot_deg=1.0)[0])
    cycle = frames + frames[-2:0:-1]  # 0 1 2 3 4 5 4 3 2 1 | 0 1 ...: consecutive frames always differ by one small warp
    seq = np.stack([cycle[k % len(cycle)] for k in range(n_frames)])
    rng = np.random.default_rng(3)
    counts = rng.integers(0, 20, n_frames - 1)
    counts[7] = 0
    uvs = [S.detect_features(seq[k], 20, seed=k % len(cycle), border=10)[:counts[k]] for k in range(n_frames - 1)]
    counts = np.array([u.shape[0] for u in uvs])
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    all_uv = np.concatenate(uvs)
    lv = {}
    for variant, method, half, fb in [("basic", "inverse", 7, 0.0), ("basic", "fast", 6, 0.5), ("affine", "direct", 5, 0.0)]:
        klt = make_tracker(ctx, variant, method, half)
        klt.forward_backward_max_error = fb
        ok, cur_uv, st = klt.TrackImageSequence(levels, seq, offsets, all_uv)
        assert ok
        prm = po.make_params(variant, method, half=half)
        for k in range(n_frames - 1):
            if counts[k] == 0:
                continue
            for f in (k, k + 1):
                if f % len(cycle) not in lv:
                    lv[f % len(cycle)] = oracle.pyramid_build(seq[f], levels)
            a, b = lv[k % len(cycle)], lv[(k + 1) % len(cycle)]
            exp = forward_backward_oracle(oracle, prm, a, b, uvs[k], fb) if fb > 0 else oracle.klt_track(prm, a, b, uvs[k])
            sl = slice(offsets[k], offsets[k + 1])
            assert_same(f"sequence {variant}/{method} fb={fb} pair {k}", (True, cur_uv[sl], st[sl]), exp)
    ok, _, _ = klt.TrackImageSequence(levels, seq[:1], np.zeros(1, np.int32), np.zeros((0, 2), np.float32))
    assert not ok  # fewer than two frames == empty input


@pytest.mark.parametrize("variant,method,half", [("basic", "inverse", 7), ("basic", "direct", 6), ("affine", "fast", 6), ("lssd", "inverse", 5)])
def test_klt_forward_backward_check(ctx, oracle, variant, method, half):
    """forward_backward_max_error: the backward pass and the consistency test equal the composition of two reference calls."""
    ref, cur, uv, _ = S.make_pair(240, 320, 300, pair_id=31)
    cur = cur.copy()
    cur[60:140, 100:220] = np.random.default_rng(5).integers(0, 255, (80, 120), dtype=np.uint8)  # an occluder: forward tracks land on noise
    pyr = ft.ImagePyramidBatch(ctx, 240, 320, 3, 2)
    pyr.SetRawImages(np.stack([ref, cur]))
    pyr.CreateImagePyramid()
    klt = make_tracker(ctx, variant, method, half)
    plain = klt.TrackFeatures(pyr, pyr, uv, ref_image=0, cur_image=1)
    klt.forward_backward_max_error = 0.7
    got = klt.TrackFeatures(pyr, pyr, uv, ref_image=0, cur_image=1)
    prm = po.make_params(variant, method, half=half)
    exp = forward_backward_oracle(oracle, prm, oracle.pyramid_build(ref, 3), oracle.pyramid_build(cur, 3), uv, 0.7)
    assert_same(f"forward-backward {variant}/{method}", got, exp)
    rejected = (plain[2] == 1) & (got[2] == 2)
    assert rejected.sum() > 0 and (got[2] == 1).sum() > 100  # the test rejects something and keeps the bulk
    assert bits_equal(got[1], plain[1])  # positions are the forward pass's


def test_klt_north_star_config(ctx, oracle):
    """BASELINE configs[0]: basic inverse, 4 levels, 15x15 patches, 200 features on a 752x480 pair."""
    ref, cur, uv, fwd = S.make_pair(480, 752, 200, pair_id=0)
    pyr = ft.ImagePyramidBatch(ctx, 480, 752, 4, 2)
    pyr.SetRawImages(np.stack([ref, cur]))
    pyr.CreateImagePyramid()
    klt = make_tracker(ctx, "basic", "inverse", 7)
    got = klt.TrackFeatures(pyr, pyr, uv, ref_image=0, cur_image=1)
    exp = oracle.klt_track(po.make_params("basic", "inverse", half=7), oracle.pyramid_build(ref, 4), oracle.pyramid_build(cur, 4), uv)
    assert_same("north star", got, exp)
    good = got[2] == 1
    assert good.mean() > 0.9 and np.median(np.linalg.norm(got[1][good] - fwd(uv)[good], axis=1)) < 0.3


def test_klt_lssd_c3_shape(ctx, oracle):
    """BASELINE configs[2] at reduced count: LSSD inverse, 21x21 patches, 1280x720."""
    ref, cur, uv, _ = S.make_pair(720, 1280, 300, pair_id=2)
    pyr = ft.ImagePyramidBatch(ctx, 720, 1280, 4, 2)
    pyr.SetRawImages(np.stack([ref, cur]))
    pyr.CreateImagePyramid()
    klt = make_tracker(ctx, "lssd", "inverse", 10, max_points=10000)
    got = klt.TrackFeatures(pyr, pyr, uv, ref_image=0, cur_image=1)
    exp = oracle.klt_track(po.make_params("lssd", "inverse", half=10, max_points=10000), oracle.pyramid_build(ref, 4), oracle.pyramid_build(cur, 4), uv)
    assert_same("lssd c3", got, exp)


# ---- descriptor matching -----------------------------------------------------------------------------------------------
def brief_matcher(ctx, max_dist, drow=40, dcol=40):
    m = ft.BriefMatcher(ctx)
    m.options().kMaxValidDescriptorDistance = max_dist
    m.options().kMaxValidPredictRowDistance = drow
    m.options().kMaxValidPredictColDistance = dcol
    return m


def test_matchers_match_golden(ctx, matcher_golden):
    g = matcher_golden
    m = brief_matcher(ctx, 60.0, 50, 50)
    rb, cb = ft.pack_brief(g["brief_ref"]), ft.pack_brief(g["brief_cur"])
    ok, idx = m.ForceMatch(rb, cb)
    assert ok and (idx == g["brief_force_idx"]).all()
    ok, idx = m.NearbyMatch(rb, cb, g["brief_pred"], g["brief_pos"])
    assert ok and (idx == g["brief_nearby_idx"]).all()
    ok, muv, mst = m.NearbyMatchUv(rb, cb, g["brief_pred"], g["brief_pos"])
    assert ok and (mst == g["brief_nearby_st"]).all() and bits_equal(muv[mst == 1], g["brief_nearby_uv"][mst == 1])
    c = ft.CosineMatcher(ctx)
    c.options().kMaxValidDescriptorDistance = 0.1
    ok, idx = c.ForceMatch(g["float_ref"], g["float_cur"])
    assert ok and (idx == g["float_force_idx"]).all()
    c.options().kMaxValidDescriptorDistance = 0.3
    c.options().kMaxValidPredictRowDistance = c.options().kMaxValidPredictColDistance = 50
    ok, idx = c.NearbyMatch(g["float_ref"], g["float_cur"], g["float_pred"], g["float_pos"])
    assert ok and (idx == g["float_nearby_idx"]).all()


@pytest.mark.parametrize("n_ref,n_cur,bits", [(1, 1, 256), (257, 300, 256), (1000, 777, 256), (300, 5000, 256), (64, 200, 128), (50, 60, 512), (40, 70, 96)])
def test_hamming_force_vs_oracle(ctx, oracle, n_ref, n_cur, bits):
    rb, cb, _, _, _ = S.make_brief_sets(n_ref, n_cur, bits=bits, seed=n_ref + n_cur)
    for max_dist in (bits * 0.23, bits * 2.0, 0.0):
        exp = oracle.match_brief_force(rb, cb, max_dist)
        got = brief_matcher(ctx, max_dist).ForceMatch(ft.pack_brief(rb), ft.pack_brief(cb))
        assert got[0] == exp[0] and (got[1] == exp[1]).all(), (n_ref, n_cur, bits, max_dist)


def test_hamming_semantics(ctx, oracle):
    rng = np.random.default_rng(4)
    cur = rng.integers(0, 2, (6, 64), dtype=np.uint8)
    cur[4] = cur[1]  # tie at distance 0: lowest j wins
    ref = cur[[1, 3]].copy()
    m = brief_matcher(ctx, 10.0)
    assert list(m.ForceMatch(ft.pack_brief(ref), ft.pack_brief(cur))[1]) == [1, 3]
    m0 = brief_matcher(ctx, 0.0)  # reference default: nothing matches
    assert list(m0.ForceMatch(ft.pack_brief(ref), ft.pack_brief(cur))[1]) == [-1, -1]
    assert list(m0.ForceMatch(ft.pack_brief(ref), ft.pack_brief(cur), np.array([5, 2], np.int32))[1]) == [5, 2]  # preset indices survive
    ok, _ = m.ForceMatch(ft.pack_brief(ref), np.zeros((0, 2), np.uint32))
    assert not ok  # descriptor_matcher.h:58
    pos = np.array([[0, 0], [100, 100], [10, 10], [50, 50], [12, 12], [300, 300]], np.float32)
    pred = np.array([[11, 11], [52, 52]], np.float32)
    mn = brief_matcher(ctx, 64.0, 5, 5)
    assert list(mn.NearbyMatch(ft.pack_brief(ref), ft.pack_brief(cur), pred, pos)[1]) == [4, 3]
    ok, _ = mn.NearbyMatch(ft.pack_brief(ref), ft.pack_brief(cur), pred[:1], pos)
    assert not ok  # descriptor_matcher.h:95


@pytest.mark.parametrize("bits", [256, 96])
def test_hamming_pairs_batch_vs_oracle(ctx, oracle, bits):
    """ftk_match_hamming_pairs: many small Force / NearbyMatch problems in one launch equal the per-pair reference loop, including empty
    pairs, duplicates (ties -> lowest index) and pre-filled index vectors."""
    rng = np.random.default_rng(bits)
    sizes = [(40, 50), (0, 30), (25, 0), (300, 280), (1, 1), (70, 33), (64, 64)]
    refs, curs, preds, poss = [], [], [], []
    for k, (nr, nc) in enumerate(sizes):
        rb, cb, pred, pos, _ = S.make_brief_sets(max(nr, 1), max(nc, 1), seed=70 + k, bits=bits)
        rb, cb, pred, pos = rb[:nr], cb[:nc], pred[:nr], pos[:nc]
        if nc > 4:
            cb[3] = cb[1]  # exact duplicates: the lower index must win
        refs.append(rb), curs.append(cb), preds.append(pred), poss.append(pos)
    ro = np.concatenate([[0], np.cumsum([len(r) for r in refs])]).astype(np.int32)
    co = np.concatenate([[0], np.cumsum([len(c) for c in curs])]).astype(np.int32)
    pack = lambda blocks: np.concatenate([ft.pack_brief(b) if len(b) else np.zeros((0, bits // 32), np.uint32) for b in blocks])
    R, Cc = pack(refs), pack(curs)
    P, Q = np.concatenate(preds), np.concatenate(poss)
    m = brief_matcher(ctx, 60.0 if bits == 256 else 30.0, 45, 55)
    prefill = rng.integers(-1, 5, ro[-1]).astype(np.int32)
    for nearby in (False, True):
        for pre in (None, prefill):
            ok, idx = m.MatchPairs(R, ro, Cc, co, P if nearby else None, Q if nearby else None, index_pairs_in_cur=pre)
            assert ok
            for k, (nr, nc) in enumerate(sizes):
                got = idx[ro[k]:ro[k + 1]]
                start = None if pre is None else pre[ro[k]:ro[k + 1]]
                if nc == 0 or nr == 0:  # the reference returns false and leaves the vector as it came
                    exp = np.full(nr, -1, np.int32) if start is None else start
                elif nearby:
                    exp = oracle.match_brief_nearby(refs[k], curs[k], preds[k], poss[k], 45, 55, m.options().kMaxValidDescriptorDistance, idx=start)[1]
                else:
                    exp = oracle.match_brief_force(refs[k], curs[k], m.options().kMaxValidDescriptorDistance, idx=start)[1]
                assert np.array_equal(got, exp), (bits, nearby, pre is not None, k)


@pytest.mark.parametrize("n_ref,n_cur,win", [(500, 600, (50, 50)), (1500, 1400, (30, 45)), (200, 3000, (5, 80)), (300, 300, (1000, 1000)), (100, 120, (0, 0))])
def test_hamming_nearby_vs_oracle(ctx, oracle, n_ref, n_cur, win):
    rb, cb, pred, pos, _ = S.make_brief_sets(n_ref, n_cur, seed=3 * n_ref + n_cur)
    pos[::17] = np.round(pos[::17])  # exact window-edge cases
    pred[::13] = np.round(pred[::13])
    if n_ref == 500:  # non-finite coordinates follow the reference's comparison semantics (NaN passes the gate)
        pos[5] = [np.nan, 100.0]
        pos[6] = [np.inf, 50.0]
        pred[9] = [np.nan, np.nan]
        pred[10] = [200.0, np.nan]
    exp = oracle.match_brief_nearby(rb, cb, pred, pos, win[0], win[1], 70.0)
    got = brief_matcher(ctx, 70.0, win[0], win[1]).NearbyMatch(ft.pack_brief(rb), ft.pack_brief(cb), pred, pos)
    assert got[0] == exp[0]
    bad = np.nonzero(got[1] != exp[1])[0]
    assert bad.size == 0, (bad[:10], got[1][bad[:10]], exp[1][bad[:10]])


@pytest.mark.parametrize("n_ref,n_cur,dim", [(1, 1, 256), (130, 170, 256), (400, 333, 256), (60, 90, 64), (33, 40, 100)])
def test_cosine_force_vs_oracle(ctx, oracle, n_ref, n_cur, dim):
    rf, cf = S.make_float_sets(n_ref, n_cur, dim=dim, seed=dim + n_ref)
    if n_ref > 100:
        cf[7] = cf[3]          # exact tie: lowest j wins
        rf[5] = 0.0            # zero-norm descriptor -> NaN distance -> never matches
    for max_dist in (0.1, 0.6):
        exp = oracle.match_cosine_force(rf, cf, max_dist)
        c = ft.CosineMatcher(ctx)
        c.options().kMaxValidDescriptorDistance = max_dist
        got = c.ForceMatch(rf, cf)
        assert got[0] == exp[0] and (got[1] == exp[1]).all(), (n_ref, n_cur, dim, max_dist, np.nonzero(got[1] != exp[1])[0][:10])


def test_cosine_tensor_path_decides_without_fallback(ctx, oracle):
    """The tcgen05 GEMM + top-2 must itself find the matches: on well-separated data no row may need the exact fall-back scan
    (otherwise a broken GEMM could hide behind it), and the result still equals the oracle's."""
    rf, cf = S.make_float_sets(700, 900, dim=256, seed=77)
    c = ft.CosineMatcher(ctx)
    c.options().kMaxValidDescriptorDistance = 0.1
    ok, idx = c.ForceMatch(rf, cf)
    assert ok and c.last_exact_scan_items() == 0
    assert (idx == oracle.match_cosine_force(rf, cf, 0.1)[1]).all() and (idx >= 0).mean() > 0.95
    # near-duplicates inside the error margin force the exact path for those rows only
    cf2 = cf.copy()
    cf2[10] = cf2[3] * np.float32(1.0001)
    ok, idx2 = c.ForceMatch(rf, cf2)
    assert 0 < c.last_exact_scan_items() < 20
    assert (idx2 == oracle.match_cosine_force(rf, cf2, 0.1)[1]).all()
    # dims that are not a multiple of 64, tiny problems, a dim above the tensor-path limit
    for n_ref, n_cur, dim in [(5, 3, 200), (129, 257, 65), (50, 60, 300)]:
        a, b = S.make_float_sets(n_ref, n_cur, dim=dim, seed=dim)
        ok, got = c.ForceMatch(a, b)
        assert (got == oracle.match_cosine_force(a, b, 0.1)[1]).all(), (n_ref, n_cur, dim)


def test_cosine_abnormal_norms_follow_the_reference(ctx, oracle):
    """Descriptors whose fp32 sum of squares overflows (norm = inf => dot / inf = 0 => d = 0.5), underflows (norm = 0 => NaN or +-inf)
    or is NaN: the tensor-core path must give the reference's answer for them too (they are evaluated exactly)."""
    rf, cf = S.make_float_sets(300, 400, dim=256, seed=77)
    rf, cf = rf.copy(), cf.copy()
    rf[5] *= np.float32(1e20)     # reference row with an infinite norm
    cf[7] *= np.float32(1e20)     # current column with an infinite norm: d = 0.5 to every finite row
    cf[9] *= np.float32(1e-30)    # squares underflow: norm 0
    rf[11] = 0.0                  # zero descriptor: NaN distances
    cf[13, 4] = np.nan
    rf[17] *= np.float32(1e-25)   # tiny but exactly representable scale: norm ~1e-25 (abnormal range, still a well-defined distance)
    c = ft.CosineMatcher(ctx)
    for max_dist in (0.1, 0.45, 0.6, 2.0):
        c.options().kMaxValidDescriptorDistance = max_dist
        ok, idx = c.ForceMatch(rf, cf)
        ok_e, exp = oracle.match_cosine_force(rf, cf, max_dist)
        assert ok == ok_e and np.array_equal(idx, exp), (max_dist, np.nonzero(idx != exp)[0][:10], idx[idx != exp][:10], exp[idx != exp][:10])
        assert c.last_exact_scan_items() >= 2  # rows 5, 11 and 17 went through the exact scan


def test_cosine_force_odd_dims_and_unaligned_device_pointers(ctx, oracle):
    """The tensor-core path reads descriptor rows as float4s when it may (dim % 4 == 0, 16-byte aligned bases) and as scalars
    otherwise: dims that are not multiples of 4, and device pointers that are only 4-byte aligned, give the reference's result too."""
    import ctypes as C
    import torch
    from feature_tracker_b200 import _capi
    from feature_tracker_b200.api import lib as ftk_lib
    L = ftk_lib()
    dev = torch.device("cuda", 0)
    vp = C.c_void_p
    fl = _capi.FLAG_DEVICE_POINTERS | _capi.FLAG_NO_INDEX_INPUT
    for dim, shift in ((1, 0), (3, 0), (37, 0), (101, 1), (250, 0), (256, 1), (256, 3), (128, 2), (64, 0)):
        rf, cf = S.make_float_sets(700, 900, dim=dim, seed=90 + dim + shift)
        exp = oracle.match_cosine_force(rf, cf, 0.1)[1]
        buf_r = torch.zeros(rf.size + 8, dtype=torch.float32, device=dev)
        buf_c = torch.zeros(cf.size + 8, dtype=torch.float32, device=dev)
        d_r, d_c = buf_r[shift:shift + rf.size], buf_c[shift:shift + cf.size]
        d_r.copy_(torch.from_numpy(rf.ravel()))
        d_c.copy_(torch.from_numpy(cf.ravel()))
        d_idx = torch.full((700,), -1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        ctx.check(L.ftk_match_cosine_force(ctx._h, vp(d_r.data_ptr()), 700, vp(d_c.data_ptr()), 900, dim, 0.1, vp(d_idx.data_ptr()), fl))
        ctx.synchronize()
        assert np.array_equal(d_idx.cpu().numpy(), exp), (dim, shift)


def test_cosine_large_shapes_tensor_path_equals_exact_kernel(monkeypatch):
    """Shapes far beyond BASELINE's: every persistent CTA of the tcgen05 kernel walks dozens of work items and changes its reference tile
    many times (pipeline parities wrap repeatedly).  The checker here is the library's own exact CUDA-core kernel (a context created with
    FTK_DISABLE_FASTPATH=1), which the smaller tests pin to the oracle; sizes the CPU oracle would need minutes for."""
    import ctypes as C
    import torch
    from feature_tracker_b200 import _capi
    from feature_tracker_b200.api import lib as ftk_lib
    L = ftk_lib()
    monkeypatch.delenv("FTK_DISABLE_FASTPATH", raising=False)
    fast = ft.Context(0)
    monkeypatch.setenv("FTK_DISABLE_FASTPATH", "1")
    exact = ft.Context(0)
    monkeypatch.delenv("FTK_DISABLE_FASTPATH", raising=False)
    dev = torch.device("cuda", 0)
    vp = C.c_void_p
    fl = _capi.FLAG_DEVICE_POINTERS | _capi.FLAG_NO_INDEX_INPUT
    for n_ref, n_cur, dim, max_dist in ((70001, 33000, 96, 0.1), (150000, 9000, 256, 0.1), (3000, 120000, 64, 0.6), (40000, 40000, 128, 0.05)):
        g = torch.Generator(device=dev).manual_seed(n_ref + dim)
        d_r = torch.randn((n_ref, dim), generator=g, device=dev, dtype=torch.float32)
        d_c = torch.randn((n_cur, dim), generator=g, device=dev, dtype=torch.float32)
        m = min(n_ref, n_cur)
        perm = torch.randperm(n_cur, generator=g, device=dev)
        d_c[perm[:m]] = d_r[:m] + 0.2 * torch.randn((m, dim), generator=g, device=dev, dtype=torch.float32)  # planted partners
        d_c[perm[1]] = d_c[perm[0]]  # an exact duplicate column: ties resolve to the lower index in both paths
        out = []
        for context in (fast, exact):
            d_idx = torch.full((n_ref,), -7, dtype=torch.int32, device=dev)
            torch.cuda.synchronize()
            context.check(L.ftk_match_cosine_force(context._h, vp(d_r.data_ptr()), n_ref, vp(d_c.data_ptr()), n_cur, dim, max_dist, vp(d_idx.data_ptr()), fl))
            context.synchronize()
            out.append(d_idx.cpu().numpy())
        assert np.array_equal(out[0], out[1]), (n_ref, n_cur, dim, int((out[0] != out[1]).sum()))
        assert (out[0] >= -1).all() and (out[0][:m] >= 0).mean() > 0.9


def test_cosine_two_devices_in_one_process(oracle):
    """One process, contexts on two GPUs: the tensor-core kernel's shared-memory opt-in is per device.  Skipped on a
    single-GPU box."""
    import ctypes as C
    from feature_tracker_b200.api import lib as ftk_lib
    try:
        other = ft.Context(1)
    except Exception:
        pytest.skip("needs a second GPU")
    rf, cf = S.make_float_sets(500, 700, dim=256, seed=78)
    exp = oracle.match_cosine_force(rf, cf, 0.1)[1]
    for context in (ft.Context(0), other, ft.Context(0)):
        c = ft.CosineMatcher(context)
        c.options().kMaxValidDescriptorDistance = 0.1
        ok, idx = c.ForceMatch(rf, cf)
        assert ok and np.array_equal(idx, exp)


def test_cosine_pair_batches_vs_oracle(ctx, oracle):
    """ftk_match_cosine_pairs: many independent ForceMatch / NearbyMatch problems (ragged sizes, an empty ref set, an empty cur set,
    exact duplicates whose distance rounds to 0 -> NearbyMatch's break) == the reference's per-pair calls."""
    rng = np.random.default_rng(91)
    sizes = [(40, 55), (0, 10), (33, 1), (7, 0), (120, 90), (64, 64)]
    refs, curs, preds, poss = [], [], [], []
    for q, (nr, nc) in enumerate(sizes):
        a, b = S.make_float_sets(max(nr, 1), max(nc, 1), dim=128, seed=200 + q)
        a, b = a[:nr].copy(), b[:nc].copy()
        if nr > 5 and nc > 9:
            a[3] = b[8]  # exact duplicate
        refs.append(a), curs.append(b)
        poss.append(np.stack([rng.uniform(0, 751, nc), rng.uniform(0, 479, nc)], 1).astype(np.float32))
        preds.append(np.stack([rng.uniform(0, 751, nr), rng.uniform(0, 479, nr)], 1).astype(np.float32))
        if nr > 5 and nc > 9:
            preds[-1][3] = poss[-1][8]
    ro = np.concatenate([[0], np.cumsum([x[0] for x in sizes])]).astype(np.int32)
    co = np.concatenate([[0], np.cumsum([x[1] for x in sizes])]).astype(np.int32)
    R, Cc = np.concatenate(refs), np.concatenate(curs)
    c = ft.CosineMatcher(ctx)
    c.options().kMaxValidDescriptorDistance = 0.45
    c.options().kMaxValidPredictRowDistance, c.options().kMaxValidPredictColDistance = 200, 300
    ok, idx = c.MatchPairs(R, ro, Cc, co)
    ok2, nidx = c.MatchPairs(R, ro, Cc, co, np.concatenate(preds), np.concatenate(poss))
    assert ok and ok2
    matched = 0
    for q, (nr, nc) in enumerate(sizes):
        got, ngot = idx[ro[q]:ro[q + 1]], nidx[ro[q]:ro[q + 1]]
        if nr == 0:
            continue
        if nc == 0:  # the reference returns false and touches nothing: the entries keep their initial -1
            assert (got == -1).all() and (ngot == -1).all()
            continue
        exp = oracle.match_cosine_force(refs[q], curs[q], 0.45)[1]
        nexp = oracle.match_cosine_nearby(refs[q], curs[q], preds[q], poss[q], 200, 300, 0.45)[1]
        assert np.array_equal(got, exp), (q, np.nonzero(got != exp)[0][:5])
        assert np.array_equal(ngot, nexp), (q, np.nonzero(ngot != nexp)[0][:5])
        matched += int((exp >= 0).sum())
    assert matched > 100


def test_cosine_nearby_vs_oracle(ctx, oracle):
    rf, cf = S.make_float_sets(300, 350, dim=256, seed=31)
    rng = np.random.default_rng(8)
    pos = np.stack([rng.uniform(0, 751, 350), rng.uniform(0, 479, 350)], 1).astype(np.float32)
    pred = np.stack([rng.uniform(0, 751, 300), rng.uniform(0, 479, 300)], 1).astype(np.float32)
    rf[10] = cf[20]  # exact duplicate -> distance that may round to exactly 0 (the `break` path)
    pred[10] = pos[20]
    exp = oracle.match_cosine_nearby(rf, cf, pred, pos, 120, 150, 0.45)
    c = ft.CosineMatcher(ctx)
    c.options().kMaxValidDescriptorDistance = 0.45
    c.options().kMaxValidPredictRowDistance, c.options().kMaxValidPredictColDistance = 120, 150
    got = c.NearbyMatch(rf, cf, pred, pos)
    assert got[0] == exp[0] and (got[1] == exp[1]).all()
    assert (exp[1] >= 0).sum() > 50


def assert_pose_same(tag, got, exp):
    ok_g, uv_g, q_g, p_g, st_g = got
    ok_e, uv_e, q_e, p_e, st_e = exp
    assert ok_g == ok_e, tag
    assert np.array_equal(st_g, st_e), f"{tag}: status differs at {np.nonzero(st_g != st_e)[0][:10]}"
    assert bits_equal(q_g, q_e) and bits_equal(p_g, p_e), f"{tag}: pose differs: gpu q={q_g} p={p_g} oracle q={q_e} p={p_e}"
    assert bits_equal(uv_g, uv_e), f"{tag}: projected positions differ (max {np.abs(uv_g - uv_e).max()} px)"


@pytest.mark.parametrize("shape,levels,half,n_feat,max_points", [((240, 320), 4, 6, 60, 500), ((240, 320), 3, 4, 150, 40), ((480, 752), 4, 6, 300, 500),
                                                                   ((120, 160), 1, 7, 25, 500), ((240, 320), 5, 3, 80, 500)])
def test_direct_method_vs_oracle(ctx, oracle, shape, levels, half, n_feat, max_points):
    """ftk_direct_method_track == DirectMethod::TrackFeatures of the reference, bit for bit."""
    rows, cols = shape
    ref, cur, uv, K, pts = S.make_direct_method_scene(rows, cols, n_feat, pair_id=60 + levels, border=10)
    n = uv.shape[0]
    pts = pts.copy()
    pts[1, 2] = -1.0
    pts[2, 2] = 1e-7
    pts[4] = (50.0, 0.0, 1.0)
    pyr = ft.ImagePyramidBatch(ctx, rows, cols, levels, 2)
    pyr.SetRawImages(np.stack([ref, cur]))
    pyr.CreateImagePyramid()
    rl, cl = oracle.pyramid_build(ref, levels), oracle.pyramid_build(cur, levels)
    dm = ft.DirectMethod(ctx)
    dm.options().kPatchRowHalfSize = dm.options().kPatchColHalfSize = half
    dm.options().kMaxTrackPointsNumber = max_points
    q0 = np.array([0.9999, 0.002, -0.003, 0.004], np.float32)
    p0 = np.array([0.01, -0.02, 0.005], np.float32)
    for method in ("direct", "inverse", "fast"):
        dm.options().kMethod = {"direct": 1, "inverse": 0, "fast": 2}[method]
        prm = po.make_direct_params(half=half, max_points=max_points, method=method)
        for kwargs in ({}, {"cur_uv": uv + 0.5, "status": np.full(n, 2, np.uint8)}):
            got = dm.TrackFeatures(pyr, pyr, K, pts, uv, q0, p0, cur_pixel_uv=kwargs.get("cur_uv"), status=kwargs.get("status"), ref_image=0, cur_image=1)
            exp = oracle.direct_method_track(prm, rl, cl, K, pts, uv, q0, p0, **kwargs)
            assert_pose_same(f"direct method {method} {sorted(kwargs)}", got, exp)
    dm.options().kMethod = 1
    ok, cur_uv, q, p, st = dm.TrackFeatures(pyr, pyr, K, pts, uv, [1, 0, 0, 0], [0, 0, 0], ref_image=0, cur_image=1)
    assert ok and (st == 1).mean() > 0.7 and np.abs(p).max() > 1e-3  # the pose moved, most projections stay inside


def test_direct_method_matches_golden(ctx):
    """The CUDA path against the committed outputs of the reference itself on its own KITTI fixture (1241x376, 5 levels, 300 features,
    pose / positions / status carried over two frames like test/test_direct_method.cpp:69-86)."""
    import os
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, "direct_method_golden.npz")))
    levels = int(g["levels"])
    rows, cols = g["left"].shape
    pyr = ft.ImagePyramidBatch(ctx, rows, cols, levels, 3)
    pyr.SetRawImages(np.stack([g["left"], g["cur1"], g["cur2"]]))
    pyr.CreateImagePyramid()
    dm = ft.DirectMethod(ctx)
    q, p, cur_uv, st = np.array([1, 0, 0, 0], np.float32), np.zeros(3, np.float32), None, None
    for i in (1, 2):
        ok, cur_uv, q, p, st = dm.TrackFeatures(pyr, pyr, g["K"], g["p_c_in_ref"], g["uv"], q, p, cur_pixel_uv=cur_uv, status=st, ref_image=0, cur_image=i)
        assert_pose_same(f"golden frame {i}", (ok, cur_uv, q, p, st), (True, g[f"uv_{i}"], g[f"q_{i}"], g[f"p_{i}"], g[f"st_{i}"]))


def test_direct_method_world_frame_overload(ctx, reflib):
    """direct_method_tracker.cpp:8-39: the world-frame overload (host-side quaternion algebra around the kernel) against the
    reference's own overload, bit for bit."""
    rows, cols, levels = 240, 320, 4
    ref, cur, uv, K, pts = S.make_direct_method_scene(rows, cols, 70, pair_id=90, border=10)
    ref_q = np.array([0.96, 0.1, -0.2, 0.15], np.float32)
    ref_q /= np.float32(np.linalg.norm(ref_q))
    ref_p = np.array([1.5, -0.7, 3.0], np.float32)
    from feature_tracker_b200 import quat
    p_w = quat.rotate(ref_q, pts) + ref_p  # world points that the reference camera sees at `pts`
    cur_q, cur_p = ref_q.copy(), ref_p.copy()  # prediction: the camera did not move
    pyr = ft.ImagePyramidBatch(ctx, rows, cols, levels, 2)
    pyr.SetRawImages(np.stack([ref, cur]))
    pyr.CreateImagePyramid()
    dm = ft.DirectMethod(ctx)
    got = dm.TrackFeaturesWorld(pyr, pyr, K, ref_q, ref_p, p_w, uv, cur_q, cur_p, ref_image=0, cur_image=1)
    exp = reflib.direct_method_track_world(po.make_direct_params(), reflib.pyramid_build(ref, levels), reflib.pyramid_build(cur, levels), K, ref_q, ref_p,
                                           p_w, uv, cur_q, cur_p)
    assert_pose_same("world-frame overload", got, exp)
    assert np.abs(got[3] - ref_p).max() > 1e-3


def test_direct_method_batch_of_pairs(ctx, oracle):
    """Several independent pose problems in one launch (one CTA per frame pair), ragged feature counts, image maps."""
    rows, cols, levels = 240, 320, 4
    scenes = [S.make_direct_method_scene(rows, cols, 40 + 15 * p, pair_id=70 + p, border=10) for p in range(5)]
    imgs = np.stack([s[0] for s in scenes] + [s[1] for s in scenes])
    pyr = ft.ImagePyramidBatch(ctx, rows, cols, levels, 10)
    pyr.SetRawImages(imgs)
    pyr.CreateImagePyramid()
    counts = [s[2].shape[0] for s in scenes]
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    dm = ft.DirectMethod(ctx)
    K = np.stack([s[3] for s in scenes])
    q0 = np.tile(np.array([1, 0, 0, 0], np.float32), (5, 1))
    p0 = np.zeros((5, 3), np.float32)
    ok, cur_uv, q, p, st = dm.TrackFeaturesBatch(pyr, pyr, offsets, K, np.concatenate([s[4] for s in scenes]), np.concatenate([s[2] for s in scenes]), q0, p0,
                                                 ref_image=np.arange(5), cur_image=np.arange(5, 10))
    assert ok
    prm = po.make_direct_params()
    for i, s in enumerate(scenes):
        exp = oracle.direct_method_track(prm, oracle.pyramid_build(s[0], levels), oracle.pyramid_build(s[1], levels), s[3], s[4], s[2], q0[i], p0[i])
        sl = slice(offsets[i], offsets[i + 1])
        assert_pose_same(f"pair {i}", (True, cur_uv[sl], q[i], p[i], st[sl]), exp)


@pytest.mark.parametrize("shape,levels,single", [((96, 131), 3, False), ((120, 160), 1, True), ((240, 320), 4, False), ((37, 53), 2, False)])
def test_dense_flow_vs_oracle(ctx, oracle, shape, levels, single):
    """ftk_dense_flow_track == DenseOpticalFlow::Track of the reference, bit for bit (both overloads, several
    Gaussian windows, initial flow for the GrayImage overload)."""
    rows, cols = shape
    ref, cur, _, _ = S.make_pair(rows, cols, 10, pair_id=5 + levels)
    pyr = ft.ImagePyramidBatch(ctx, rows, cols, levels, 2)
    pyr.SetRawImages(np.stack([ref, cur]))
    pyr.CreateImagePyramid()
    rl, cl = oracle.pyramid_build(ref, levels), oracle.pyramid_build(cur, levels)
    dof = ft.DenseOpticalFlow(ctx)
    for half in (2, 1, 0, 3):
        dof.options().kHalfPatchSize = half
        prm = po.make_dense_flow_params(half=half)
        flows = [None]
        if single:
            rng = np.random.default_rng(half)
            flows.append((rng.normal(0, 1, ref.shape).astype(np.float32), rng.normal(0, 1, ref.shape).astype(np.float32)))
        for flow in flows:
            ok, fr, fc = dof.Track(pyr, pyr, flow_rc=flow, single_level=single, ref_image=0, cur_image=1)
            eok, er, ec = oracle.dense_flow_track(prm, rl, cl, single_level=single, flow=flow)
            assert ok and eok
            for got, exp, name in ((fr, er, "row"), (fc, ec, "col")):
                bad = np.argwhere(got.view(np.uint32) != exp.view(np.uint32))
                assert bad.size == 0, f"half={half} {name} flow differs at {bad[:5].tolist()}: gpu {got[tuple(bad[0])]} oracle {exp[tuple(bad[0])]}"
    assert np.abs(er).mean() > 0.2


def test_dense_flow_and_direct_method_vs_reference_direct(ctx, reflib):
    """The two 8(f) trackers straight against oracle/_ref (dense_optical_flow.cpp and direct_method_tracker.cpp compiled in place), not via
    the C restatement."""
    rows, cols, levels = 120, 160, 3
    ref, cur, _, _ = S.make_pair(rows, cols, 10, pair_id=55)
    pyr = ft.ImagePyramidBatch(ctx, rows, cols, levels, 2)
    pyr.SetRawImages(np.stack([ref, cur]))
    pyr.CreateImagePyramid()
    rl, cl = reflib.pyramid_build(ref, levels), reflib.pyramid_build(cur, levels)
    ok, fr, fc = ft.DenseOpticalFlow(ctx).Track(pyr, pyr, ref_image=0, cur_image=1)
    eok, er, ec = reflib.dense_flow_track(po.make_dense_flow_params(), rl, cl)
    assert ok and eok and bits_equal(fr, er) and bits_equal(fc, ec)
    sref, scur, suv, K, pts = S.make_direct_method_scene(240, 320, 80, pair_id=56)
    pyr2 = ft.ImagePyramidBatch(ctx, 240, 320, 4, 2)
    pyr2.SetRawImages(np.stack([sref, scur]))
    pyr2.CreateImagePyramid()
    got = ft.DirectMethod(ctx).TrackFeatures(pyr2, pyr2, K, pts, suv, [1, 0, 0, 0], [0, 0, 0], ref_image=0, cur_image=1)
    exp = reflib.direct_method_track(po.make_direct_params(), reflib.pyramid_build(sref, 4), reflib.pyramid_build(scur, 4), K, pts, suv, [1, 0, 0, 0], [0, 0, 0])
    assert_pose_same("direct method vs _ref", got, exp)


def test_dense_flow_matches_golden(ctx, euroc_golden):
    """The CUDA path against the committed outputs of the reference itself on its own EuRoC fixture pair (752x480, 4 levels)."""
    import os
    from conftest import GOLDEN
    from test_oracle import _check_dense_flow_golden
    g = dict(np.load(os.path.join(GOLDEN, "dense_flow_golden.npz")))

    def track(rl, cl, prm, single):
        pyr = upload_levels(ctx, [rl, cl])
        dof = ft.DenseOpticalFlow(ctx)
        dof.options().kHalfPatchSize, dof.options().kMaxIteration = prm.half_patch_size, prm.max_iteration
        return dof.Track(pyr, pyr, single_level=single, ref_image=0, cur_image=1)
    _check_dense_flow_golden(track, g, euroc_golden)


def lightglue_like_scores(n_ref, n_cur, seed):
    """Log-assignment-like matrix: a planted partial permutation of strong scores over weak background, ties, -inf, NaN."""
    rng = np.random.default_rng(seed)
    s = rng.normal(-8.0, 2.0, (n_ref, n_cur)).astype(np.float32)
    n_match = min(n_ref, n_cur) * 2 // 3
    rows, cols = rng.permutation(n_ref)[:n_match], rng.permutation(n_cur)[:n_match]
    s[rows, cols] = rng.uniform(-4.0, -0.01, n_match).astype(np.float32)  # some below the -3 threshold
    s[rng.random(s.shape) < 0.3 / n_cur] = np.float32(-0.5)  # ties between strong entries (about every third row)
    if n_ref > 8 and n_cur > 8:
        s[1, :] = -np.inf
        s[:, 2] = -np.inf
        s[3, 0] = np.nan
        s[0, 5] = np.nan
        s[6, 7] = np.nan
        s[4, 4] = -0.0
        s[4, 6] = 0.0
    return s


@pytest.mark.parametrize("n_ref,n_cur", [(1, 1), (5, 3), (300, 257), (512, 2048), (2048, 2048), (1000, 4099), (3000, 130)])
def test_mutual_scores_vs_oracle(ctx, oracle, n_ref, n_cur):
    """ftk_match_mutual_scores == the reference's score-matrix post-processing (nn_feature_matcher.cpp:180-216), index for index."""
    s = lightglue_like_scores(n_ref, n_cur, seed=n_ref * 7 + n_cur)
    m = ft.NNFeatureMatcher(ctx)
    counts = []
    for thr in (-1.0, -100.0, -3.0):
        m.options().kMinValidMatchScore = thr
        ok, idx = m.MatchScores(s)
        ok_e, exp = oracle.mutual_scores(s, thr)
        assert ok and ok_e
        bad = np.nonzero(idx != exp)[0]
        assert bad.size == 0, f"rows {bad[:10]}: gpu {idx[bad[:10]]} oracle {exp[bad[:10]]}"
        counts.append(int((exp >= 0).sum()))
    if n_ref > 8:
        assert counts[1] > min(n_ref, n_cur) // 4 and counts[0] <= counts[1]  # planted matches found; the threshold only removes
    uv_ref = np.zeros((n_ref, 2), np.float32)
    uv_cur = np.arange(2 * n_cur, dtype=np.float32).reshape(n_cur, 2)
    ok, matched, st = m.Match(s, uv_ref, uv_cur)
    assert ok and np.array_equal(st == 1, exp >= 0) and np.array_equal(matched[exp >= 0], uv_cur[exp[exp >= 0]])
    ok, _ = m.MatchScores(np.zeros((4, 0), np.float32))
    assert not ok


@pytest.mark.parametrize("n_ref,n_cur", [(300, 300), (257, 1031), (2048, 2048)])
def test_mutual_scores_vs_reference_match(ctx, reflib, n_ref, n_cur):
    """ftk_match_mutual_scores against the reference's OWN NNFeatureMatcher::Match (nn_feature_matcher.cpp:150-219 compiled in place with
    the ONNX Runtime stub; the session returns the injected score matrix), index for index."""
    s = lightglue_like_scores(n_ref, n_cur, seed=n_ref * 11 + n_cur)
    m = ft.NNFeatureMatcher(ctx)
    for thr in (-3.0, -1.0, -100.0):
        m.options().kMinValidMatchScore = thr
        ok, idx = m.MatchScores(s)
        ok_e, exp = reflib.nn_match_scores(s, thr)
        assert ok and ok_e and np.array_equal(idx, exp), (thr, np.nonzero(idx != exp)[0][:10])
    assert (exp >= 0).sum() > min(n_ref, n_cur) // 4


def test_cross_check_force_match(ctx, oracle):
    """Cross-check matching = two ForceMatch calls (reference semantics each) + the mutual filter."""
    rng = np.random.default_rng(9)
    ref_bits = rng.integers(0, 2, (700, 256), dtype=np.uint8)
    cur_bits = ref_bits[rng.permutation(700)[:600]].copy()
    cur_bits[rng.random(cur_bits.shape) < 0.08] ^= 1
    cur_bits = np.concatenate([cur_bits, rng.integers(0, 2, (150, 256), dtype=np.uint8)])
    m = brief_matcher(ctx, 60.0)
    ok, idx = m.CrossCheckForceMatch(ft.pack_brief(ref_bits), ft.pack_brief(cur_bits))
    _, fwd = oracle.match_brief_force(ref_bits, cur_bits, 60.0)
    _, bwd = oracle.match_brief_force(cur_bits, ref_bits, 60.0)
    exp = np.where((fwd >= 0) & (bwd[np.clip(fwd, 0, None)] == np.arange(700)), fwd, -1)
    assert ok and np.array_equal(idx, exp) and (exp >= 0).sum() > 500


def test_hamming_c4_full_size_properties(ctx):
    """BASELINE configs[3] at full size (10k x 10k): size-independent properties instead of the O(N*M) CPU oracle --
    planted matches are recovered, results are idempotent, and a ref-row permutation permutes the result."""
    rb, cb, pred, pos, truth = S.make_brief_sets(10000, 10000, seed=99)
    pr, pc = ft.pack_brief(rb), ft.pack_brief(cb)
    m = brief_matcher(ctx, 60.0, 50, 50)
    ok, idx = m.ForceMatch(pr, pc)
    has = truth >= 0
    assert ok and (idx[has] == truth[has]).mean() > 0.999
    ok, idx2 = m.ForceMatch(pr, pc, idx.copy())
    assert (idx2 == idx).all()
    perm = np.random.default_rng(0).permutation(10000)
    ok, idx3 = m.ForceMatch(pr[perm], pc)
    assert (idx3 == idx[perm]).all()
    # distances of the reported matches really are the row minima (checked on a sample with numpy popcount)
    sample = np.random.default_rng(1).integers(0, 10000, 64)
    d = (rb[sample][:, None, :] != cb[None, :, :]).sum(2)
    exp = np.where(d.min(1) < 60, d.argmin(1), -1)
    assert (idx[sample] == exp).all()
    ok, nidx = m.NearbyMatch(pr, pc, pred, pos)
    assert ok and (nidx[has] == truth[has]).mean() > 0.95


# ---- BASELINE.json full sizes ---------------------------------------------------------------------------------------------
def test_klt_c2_full_batch_properties(ctx, oracle):
    """BASELINE configs[1] batch shape at full size (1000 frame pairs x 2000 features, 752x480, 4 levels) for the north-star tracker
    and the affine fast tracker: size-independent properties (identical pairs give identical results wherever they sit in the
    batch) + the first copy of every unique pair checked against the oracle."""
    n_pairs, n_feat, unique = 1000, 2000, 4
    pairs = [S.make_pair(480, 752, n_feat, pair_id=300 + u) for u in range(unique)]
    refs = np.stack([pairs[p % unique][0] for p in range(n_pairs)])
    curs = np.stack([pairs[p % unique][1] for p in range(n_pairs)])
    pyr = ft.ImagePyramidBatch(ctx, 480, 752, 4, 2 * n_pairs)
    pyr.SetRawImages(refs, first=0)
    pyr.SetRawImages(curs, first=n_pairs)
    del refs, curs
    pyr.CreateImagePyramid()
    uv = np.concatenate([pairs[p % unique][2] for p in range(n_pairs)])
    offsets = (np.arange(n_pairs + 1) * n_feat).astype(np.int32)
    lv = [(oracle.pyramid_build(pairs[u][0], 4), oracle.pyramid_build(pairs[u][1], 4)) for u in range(unique)]
    for variant, method, half, check in [("basic", "inverse", 7, n_feat), ("affine", "fast", 6, 300)]:
        klt = make_tracker(ctx, variant, method, half, max_points=n_feat)
        ok, cur_uv, st = klt.TrackFeaturesBatch(pyr, pyr, offsets, uv, ref_image=np.arange(n_pairs), cur_image=np.arange(n_pairs) + n_pairs)
        assert ok
        cur_uv = cur_uv.reshape(n_pairs, n_feat, 2)
        st = st.reshape(n_pairs, n_feat)
        for u in range(unique):  # every copy of a unique pair equals the first copy, bit for bit
            assert (cur_uv[u::unique].view(np.uint32) == cur_uv[u].view(np.uint32)).all() and (st[u::unique] == st[u]).all()
            prm = po.make_params(variant, method, half=half, max_points=n_feat)
            exp = oracle.klt_track(prm, lv[u][0], lv[u][1], pairs[u][2][:check])
            assert_same(f"C2 full batch {variant}/{method} pair {u}", (True, cur_uv[u][:check], st[u][:check]), exp)
        assert (st == 1).mean() > 0.95
    pyr.close()


def test_klt_c3_full_size(ctx, oracle):
    """BASELINE configs[2] at full size: LSSD inverse, 21x21 patches, 10 000 features on a 1280x720 pair; the oracle needs a few
    seconds for the 2 000 features that are compared one to one, the rest is covered by a permutation property."""
    ref, cur, uv, _ = S.make_pair(720, 1280, 10000, pair_id=7)
    pyr = ft.ImagePyramidBatch(ctx, 720, 1280, 4, 2)
    pyr.SetRawImages(np.stack([ref, cur]))
    pyr.CreateImagePyramid()
    klt = make_tracker(ctx, "lssd", "inverse", 10, max_points=10000)
    ok, cur_uv, st = klt.TrackFeatures(pyr, pyr, uv, ref_image=0, cur_image=1)
    assert ok and (st == 1).mean() > 0.9
    exp = oracle.klt_track(po.make_params("lssd", "inverse", half=10, max_points=10000), oracle.pyramid_build(ref, 4), oracle.pyramid_build(cur, 4), uv[:2000])
    assert_same("C3 full size", (True, cur_uv[:2000], st[:2000]), exp)
    perm = np.random.default_rng(3).permutation(10000)
    ok, cur2, st2 = klt.TrackFeatures(pyr, pyr, uv[perm], ref_image=0, cur_image=1)
    assert bits_equal(cur2, cur_uv[perm]) and (st2 == st[perm]).all()  # features are independent: order does not matter


def test_cosine_c5_full_size_properties(ctx):
    """BASELINE configs[4] at full size (20k x 20k x 256): planted matches recovered without the exact fall-back, idempotence,
    row-permutation equivariance, and a sampled check of the reported arg-min against float64 numpy."""
    rf, cf = S.make_float_sets(20000, 20000, seed=5)
    c = ft.CosineMatcher(ctx)
    c.options().kMaxValidDescriptorDistance = 0.1
    ok, idx = c.ForceMatch(rf, cf)
    assert ok and (idx >= 0).all() and c.last_exact_scan_items() == 0
    ok, idx2 = c.ForceMatch(rf, cf, idx.copy())
    assert (idx2 == idx).all()
    perm = np.random.default_rng(0).permutation(20000)
    ok, idx3 = c.ForceMatch(rf[perm], cf)
    assert (idx3 == idx[perm]).all()
    sample = np.random.default_rng(1).integers(0, 20000, 128)
    d = 0.5 - 0.5 * (rf[sample].astype(np.float64) @ cf.astype(np.float64).T)
    assert (d.argmin(1) == idx[sample]).all() and (d.min(1) < 0.1).all()


# ---- BASELINE configs[3] / configs[4] at FULL size, every index against the reference's own code (oracle/_ref) ------------------
def test_hamming_c4_full_size_bit_exact_vs_reference(ctx, reflib):
    """BASELINE configs[3]: BRIEF-256 ForceMatch 10k x 10k and NearbyMatch (window 50 / 50), ALL 10 000 indices of each against
    descriptor_matcher.h:55-79 / :90-124 with the ComputeDistance of test/test_descriptor_matcher_brief.cpp:33-45, compiled in place
    (ref rows split over the host threads; every slice is the reference's own single-threaded loop)."""
    rb, cb, pred, pos, _ = S.make_brief_sets(10000, 10000, seed=99)
    pr, pc = ft.pack_brief(rb), ft.pack_brief(cb)
    m = brief_matcher(ctx, 60.0, 50, 50)
    ok, idx = m.ForceMatch(pr, pc)
    ok_e, exp = reflib.match_rows_threaded("brief_force", rb, cb, 60.0)
    assert ok and ok_e and (exp >= 0).sum() > 5000
    assert int((idx != exp).sum()) == 0
    ok, nidx = m.NearbyMatch(pr, pc, pred, pos)
    ok_e, nexp = reflib.match_rows_threaded("brief_nearby", rb, cb, 60.0, pred_uv=pred, cur_uv=pos, max_drow=50, max_dcol=50)
    assert ok and ok_e and (nexp >= 0).sum() > 5000
    assert int((nidx != nexp).sum()) == 0


def test_cosine_c5_full_size_bit_exact_vs_reference(ctx, reflib):
    """BASELINE configs[4]: float-256 ForceMatch 20k x 20k, ALL 20 000 indices against descriptor_matcher.h:55-79 with the
    ComputeDistance of test/test_descriptor_matcher_superpoint.cpp:32-34 (sequential fp32 dot / norms), compiled in place."""
    rf, cf = S.make_float_sets(20000, 20000, seed=5)
    c = ft.CosineMatcher(ctx)
    c.options().kMaxValidDescriptorDistance = 0.1
    ok, idx = c.ForceMatch(rf, cf)
    ok_e, exp = reflib.match_rows_threaded("cosine_force", rf, cf, 0.1)
    assert ok and ok_e and (exp >= 0).sum() > 10000
    assert int((idx != exp).sum()) == 0


@pytest.mark.parametrize("variant,method", [(v, m) for v in ("basic", "affine", "lssd") for m in ("inverse", "direct", "fast")])
def test_klt_vs_reference_direct(ctx, reflib, variant, method):
    """Every tracker family straight against oracle/_ref (the reference's .cpp compiled in place), not via the C restatement:
    pyramid from the reference's CreateImagePyramid, bit-identical positions, identical status."""
    ref, cur, uv, _ = S.make_pair(240, 320, 150, pair_id=40, border=10)
    rl, cl = reflib.pyramid_build(ref, 4), reflib.pyramid_build(cur, 4)
    klt = make_tracker(ctx, variant, method, 6)
    pyr = ft.ImagePyramidBatch(ctx, 240, 320, 4, 2)
    pyr.SetRawImages(np.stack([ref, cur]))
    pyr.CreateImagePyramid()
    for l in range(4):
        assert (pyr.GetLevel(0, l) == rl[l]).all() and (pyr.GetLevel(1, l) == cl[l]).all()
    got = klt.TrackFeatures(pyr, pyr, uv, ref_image=0, cur_image=1)
    exp = reflib.klt_track(po.make_params(variant, method, half=6), rl, cl, uv)
    assert_same(f"{variant}/{method} vs _ref", got, exp)


def test_matchers_vs_reference_direct(ctx, reflib):
    """ForceMatch / NearbyMatch of both descriptor types straight against oracle/_ref on mid-sized sets."""
    rb, cb, pred, pos, _ = S.make_brief_sets(1500, 1700, seed=41)
    m = brief_matcher(ctx, 60.0, 50, 50)
    assert (m.ForceMatch(ft.pack_brief(rb), ft.pack_brief(cb))[1] == reflib.match_brief_force(rb, cb, 60.0)[1]).all()
    assert (m.NearbyMatch(ft.pack_brief(rb), ft.pack_brief(cb), pred, pos)[1] == reflib.match_brief_nearby(rb, cb, pred, pos, 50, 50, 60.0)[1]).all()
    rf, cf = S.make_float_sets(1200, 1300, seed=42)
    c = ft.CosineMatcher(ctx)
    c.options().kMaxValidDescriptorDistance = 0.1
    c.options().kMaxValidPredictRowDistance = c.options().kMaxValidPredictColDistance = 50
    assert (c.ForceMatch(rf, cf)[1] == reflib.match_cosine_force(rf, cf, 0.1)[1]).all()
    assert (c.NearbyMatch(rf, cf, pred[:1200], pos[:1300])[1] == reflib.match_cosine_nearby(rf, cf, pred[:1200], pos[:1300], 50, 50, 0.1)[1]).all()


# ---- feature detection + BRIEF (parity unpinned -- the checker is the oracle's restatement) ---------------
def make_detector(ctx, kind, half, thr, dist):
    det = (ft.FeaturePointHarrisDetector if kind == "harris" else ft.FeaturePointShiTomasDetector)(ctx)
    o = det.options()
    o.kHalfPatchSize, o.kMinValidResponse, o.kMinFeatureDistance = half, thr, dist
    return det


def single_image_pyramid(ctx, img, levels=1):
    pyr = ft.ImagePyramidBatch(ctx, img.shape[0], img.shape[1], levels, 1)
    pyr.SetRawImages(img[None])
    pyr.CreateImagePyramid()
    return pyr


@pytest.mark.parametrize("kind", ["harris", "shi_tomasi"])
@pytest.mark.parametrize("half", [1, 2, 3])
def test_detector_response_bit_exact(ctx, oracle, kind, half):
    for shape, seed in (((480, 752), 31), ((97, 131), 32), ((33, 65), 33), ((9, 9), 34), ((4, 40), 35)):
        img = S.make_image(*shape, seed=seed) if min(shape) >= 32 else np.random.default_rng(seed).integers(0, 256, shape, dtype=np.uint8)
        ok, exp = oracle.detect_response(po.make_detector_params(kind, half, 0.04, 40.0, 20), img)
        got = make_detector(ctx, kind, half, 40.0, 20).ComputeResponse(single_image_pyramid(ctx, img))
        assert ok and bits_equal(got, exp), (kind, half, shape)


@pytest.mark.parametrize("kind,thr,dist,needed,shape", [
    ("harris", 40.0, 20, 300, (480, 752)),       # the demo's values (test_descriptor_matcher_brief.cpp:60-61): every pixel is a candidate
    ("shi_tomasi", 40.0, 20, 300, (480, 752)),
    ("harris", 1e5, 9, 5000, (480, 752)),        # more wanted than exist: the full maximal set
    ("shi_tomasi", 300.0, 5, 1000, (240, 376)),
    ("harris", 5e5, 1, 4000, (97, 131)),         # d = 1: nothing blocks anything
    ("harris", 5e5, 0, 4000, (97, 131)),
    ("shi_tomasi", 100.0, 64, 50, (97, 131)),    # windows wider than two warps' worth of columns
    ("harris", 1e12, 20, 10, (97, 131)),         # no candidate at all
    ("harris", -1e30, 3, 100000, (61, 83)),      # every defined pixel, plateaus of equal response included
])
def test_detector_selection_vs_oracle(ctx, oracle, kind, thr, dist, needed, shape):
    img = S.make_image(*shape, seed=41 + dist)
    if thr < 0:
        img[:, : shape[1] // 2] = 77  # flat half: thousands of exactly equal responses, resolved by index
    pyr = single_image_pyramid(ctx, img, levels=3 if min(shape) >= 64 else 1)
    det = make_detector(ctx, kind, 1, thr, dist)
    rng = np.random.default_rng(5)
    existing = np.concatenate([rng.uniform(0, 1, (40, 2)) * [shape[1], shape[0]], [[np.nan, 3.0], [-2.0, 5.0], [shape[1], 1.0], [shape[1] - 0.5, shape[0] - 0.5]]]).astype(np.float32)
    for ex in (None, existing):
        ok_e, uv_e, resp_e = oracle.detect_features(po.make_detector_params(kind, 1, 0.04, thr, dist), img, needed, existing=ex)
        want = needed + (0 if ex is None else len(ex))
        ok_g, uv_g, resp_g = det.DetectGoodFeatures(pyr, want, ex, return_response=True)
        n0 = 0 if ex is None else len(ex)
        assert ok_g and ok_e
        assert np.array_equal(uv_g[n0:], uv_e), (kind, thr, dist, needed, len(uv_g) - n0, len(uv_e))
        assert bits_equal(resp_g, resp_e)
        if ex is not None:
            assert bits_equal(uv_g[:n0], ex)


@pytest.mark.parametrize("two_pass", [False, True])
def test_detector_matches_regression_fixture(ctx, euroc_golden, monkeypatch, two_pass):
    """Both selection paths: the fused shared-memory round kernel (default for windows that fit) and the two-kernel rounds."""
    import os
    if two_pass:
        monkeypatch.setenv("FTK_DETECT_TWO_PASS", "1")
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, "detector_golden.npz")))
    pyr = ft.ImagePyramidBatch(ctx, 480, 752, 4, 2)
    pyr.SetRawImages(np.stack([euroc_golden["ref"], euroc_golden["cur"]]))
    pyr.CreateImagePyramid()
    assert np.array_equal(ft.brief_pattern(256, 8, 0), g["pattern"])
    for i, name in enumerate(("ref", "cur")):
        for kind in ("harris", "shi_tomasi"):
            ok, uv, resp = make_detector(ctx, kind, 1, 40.0, 20).DetectGoodFeatures(pyr, 300, image_index=i, return_response=True)
            assert ok and np.array_equal(uv, g[f"{name}_{kind}_uv"]) and bits_equal(resp, g[f"{name}_{kind}_response"]), (name, kind)
        ok, desc, valid = ft.BriefDescriptor(ctx).Compute(pyr, g[f"{name}_harris_uv"], image_index=i)
        assert ok and np.array_equal(desc, g[f"{name}_brief"]) and np.array_equal(valid, g[f"{name}_brief_valid"]), name


@pytest.mark.parametrize("scratch_bytes", [None, "3000000", "1"])
def test_detector_batch_vs_oracle(ctx, oracle, monkeypatch, scratch_bytes):
    """ftk_detect_features_batch: all images of a batch in one call equal the per-image results, also when the scratch budget forces the
    batch through in chunks (two images per chunk / one image per chunk)."""
    if scratch_bytes is not None:
        monkeypatch.setenv("FTK_DETECT_SCRATCH_BYTES", scratch_bytes)
    rows, cols, n_img = 150, 200, 7
    imgs = np.stack([S.make_image(rows, cols, seed=60 + i) for i in range(n_img)])
    imgs[3] = 128          # no corner at all
    imgs[5, :, :100] = 9   # a flat half
    pyr = ft.ImagePyramidBatch(ctx, rows, cols, 2, n_img)
    pyr.SetRawImages(imgs)
    pyr.CreateImagePyramid()
    for kind, thr, dist, needed, first, count in (("harris", 1e4, 12, 40, 0, n_img), ("shi_tomasi", 50.0, 7, 1000, 2, 4), ("harris", 40.0, 40, 5, 0, n_img)):
        det = make_detector(ctx, kind, 1, thr, dist)
        ok, uvs, resps = det.DetectGoodFeaturesBatch(pyr, needed, first=first, count=count)
        assert ok and len(uvs) == count
        for i in range(count):
            ok_e, uv_e, resp_e = oracle.detect_features(po.make_detector_params(kind, 1, 0.04, thr, dist), imgs[first + i], needed)
            assert ok_e and np.array_equal(uvs[i], uv_e) and bits_equal(resps[i], resp_e), (kind, first + i, len(uvs[i]), len(uv_e))


def test_brief_vs_oracle(ctx, oracle):
    img = S.make_image(90, 120, seed=21)
    pyr = single_image_pyramid(ctx, img)
    rng = np.random.default_rng(3)
    uv = np.concatenate([rng.uniform(-5, 125, (3000, 2)).astype(np.float32),
                         np.array([[8.0, 8.0], [111.99, 81.99], [112.0, 40.0], [np.nan, 10], [7.99, 30.0], [1e30, 1e30], [-1e30, 5]], np.float32)])
    for n_bits, half, seed in ((256, 8, 0), (128, 4, 77), (32, 15, 1), (1024, 20, 9)):
        d = ft.BriefDescriptor(ctx)
        d.options().kLength, d.options().kHalfPatchSize, d.options().kPatternSeed = n_bits, half, seed
        assert np.array_equal(d.pattern(), oracle.brief_pattern(n_bits, half, seed))
        ok_g, desc_g, valid_g = d.Compute(pyr, uv)
        ok_e, desc_e, valid_e = oracle.describe_brief(img, uv, d.pattern(), half)
        assert ok_g and ok_e and np.array_equal(valid_g, valid_e) and np.array_equal(desc_g, desc_e), (n_bits, half)
    ok, desc, valid = ft.BriefDescriptor(ctx).Compute(pyr, np.zeros((0, 2), np.float32))
    assert ok and desc.shape[0] == 0
    with pytest.raises(ft.FtkError):  # a pair outside the stated patch
        ft.BriefDescriptor(ctx, pattern=np.full((32, 4), 9, np.int8)).Compute(pyr, uv[:4])


def test_front_end_batched_detect_describe_match(ctx, oracle):
    """The batched front end -- ftk_detect_features_batch -> ftk_describe_brief_batch -> ftk_match_hamming_pairs: a handful of launches
    for many frame pairs -- equals the per-image / per-pair oracle calls stage by stage."""
    rows, cols, n_pairs = 160, 208, 5
    pairs = [S.make_pair(rows, cols, 5, pair_id=500 + p) for p in range(n_pairs)]
    imgs = np.stack([pr[0] for pr in pairs] + [pr[1] for pr in pairs])  # refs then curs
    imgs[2] = 200  # a frame without corners: an empty ref set
    pyr = ft.ImagePyramidBatch(ctx, rows, cols, 3, 2 * n_pairs)
    pyr.SetRawImages(imgs)
    pyr.CreateImagePyramid()
    det = make_detector(ctx, "harris", 1, 2e4, 9)
    ok, feats, _ = det.DetectGoodFeaturesBatch(pyr, 120)
    prm = po.make_detector_params("harris", 1, 0.04, 2e4, 9)
    assert ok and all(np.array_equal(feats[i], oracle.detect_features(prm, imgs[i], 120)[1]) for i in range(2 * n_pairs))
    brief = ft.BriefDescriptor(ctx)
    ok, descs, valids = brief.ComputeBatch(pyr, feats)
    assert ok
    unpack = lambda d: np.unpackbits(np.ascontiguousarray(d).view(np.uint8).reshape(len(d), -1), axis=1, bitorder="little") if len(d) else np.zeros((0, 256), np.uint8)
    for i in range(2 * n_pairs):
        _, d_e, v_e = oracle.describe_brief(imgs[i], feats[i], brief.pattern(), 8)
        assert np.array_equal(descs[i], d_e[:len(feats[i])]) and np.array_equal(valids[i], v_e[:len(feats[i])]), i
    ro = np.concatenate([[0], np.cumsum([len(feats[p]) for p in range(n_pairs)])]).astype(np.int32)
    co = np.concatenate([[0], np.cumsum([len(feats[n_pairs + p]) for p in range(n_pairs)])]).astype(np.int32)
    m = brief_matcher(ctx, 70.0, 30, 30)
    ok, idx = m.MatchPairs(np.concatenate(descs[:n_pairs]), ro, np.concatenate(descs[n_pairs:]), co, np.concatenate(feats[:n_pairs]), np.concatenate(feats[n_pairs:]))
    assert ok and (idx >= 0).sum() > 50
    for p in range(n_pairs):
        got = idx[ro[p]:ro[p + 1]]
        if len(feats[p]) == 0 or len(feats[n_pairs + p]) == 0:
            assert (got == -1).all()
            continue
        exp = oracle.match_brief_nearby(unpack(descs[p]), unpack(descs[n_pairs + p]), feats[p], feats[n_pairs + p], 30, 30, 70.0)[1]
        assert np.array_equal(got, exp), p


def test_front_end_detect_describe_match_track(ctx, oracle, euroc_golden):
    """The reference demo's flow (test_descriptor_matcher_brief.cpp:57-95) on one device pyramid batch, each stage against the oracle
    fed with the previous stage's oracle output: detect in both frames -> BRIEF -> NearbyMatch, then KLT from the same pyramids."""
    ref, cur = euroc_golden["ref"], euroc_golden["cur"]
    pyr = ft.ImagePyramidBatch(ctx, 480, 752, 4, 2)
    pyr.SetRawImages(np.stack([ref, cur]))
    pyr.CreateImagePyramid()
    det = make_detector(ctx, "harris", 1, 40.0, 20)
    _, ref_uv = det.DetectGoodFeatures(pyr, 300, image_index=0)
    _, cur_uv = det.DetectGoodFeatures(pyr, 300, image_index=1)
    prm = po.make_detector_params("harris", 1, 0.04, 40.0, 20)
    assert np.array_equal(ref_uv, oracle.detect_features(prm, ref, 300)[1]) and np.array_equal(cur_uv, oracle.detect_features(prm, cur, 300)[1])
    brief = ft.BriefDescriptor(ctx)
    _, ref_desc, _ = brief.Compute(pyr, ref_uv, image_index=0)
    _, cur_desc, _ = brief.Compute(pyr, cur_uv, image_index=1)
    m = brief_matcher(ctx, 60.0, 50, 50)
    ok, idx = m.NearbyMatch(ref_desc, cur_desc, ref_uv, cur_uv)
    unpack = lambda d: np.unpackbits(d.view(np.uint8).reshape(len(d), -1), axis=1, bitorder="little")
    ok_e, idx_e = oracle.match_brief_nearby(unpack(ref_desc), unpack(cur_desc), ref_uv, cur_uv, 50, 50, 60.0)
    assert ok and ok_e and np.array_equal(idx, idx_e) and (idx >= 0).sum() >= 100
    klt = make_tracker(ctx, "basic", "fast", 6)
    got = klt.TrackFeatures(pyr, pyr, ref_uv, ref_image=0, cur_image=1)
    levels = [[ref] + [euroc_golden[f"ref_l{l}"] for l in range(1, 4)], [cur] + [euroc_golden[f"cur_l{l}"] for l in range(1, 4)]]
    assert_same("front end klt", got, oracle.klt_track(po.make_params("basic", "fast", 6), levels[0], levels[1], ref_uv))
