"""Generates the committed golden fixtures from the REFERENCE ITSELF (oracle/_ref/libftk_ref.so = the reference's own
.cpp files compiled in place against oracle/shim/).  Runs only where /root/reference exists; the outputs
(tests/golden/*.npz) travel with the repo.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from feature_tracker_b200 import synthetic as S  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = "/root/reference/example/optical_flow"

KLT_COMBOS = [(v, m, h) for v in ("basic", "affine", "lssd") for m in ("inverse", "direct", "fast") for h in (6, 7, 10)]


def klt_key(v, m, h, extra=""):
    return f"{v}_{m}_h{h}{extra}"


def main():
    R = po.RefLib()
    # ---- KLT on the reference's own EuRoC fixture pair -------------------------------------------------------
    ref = np.array(Image.open(os.path.join(REF_DIR, "ref_image.png")).convert("L"))
    cur = np.array(Image.open(os.path.join(REF_DIR, "cur_image.png")).convert("L"))
    # 150 corners (the reference demo detects <= 300 Harris corners, test/test_optical_flow.cpp:25,34-39) + 10 border points
    pts = S.detect_features(ref, 160, seed=42, min_distance=25, border=20, border_fraction=0.0625, jitter=False)
    levels = 4
    rl, cl = R.pyramid_build(ref, levels), R.pyramid_build(cur, levels)
    out = {"ref": ref, "cur": cur, "pts": pts, "levels": np.int32(levels)}
    for l in range(1, levels):
        out[f"ref_l{l}"] = rl[l]
        out[f"cur_l{l}"] = cl[l]
    for v, m, h in KLT_COMBOS:
        p = po.make_params(v, m, half=h, max_points=500)
        ok, uv, st = R.klt_track(p, rl, cl, pts)
        assert ok
        out[klt_key(v, m, h) + "_uv"] = uv
        out[klt_key(v, m, h) + "_st"] = st
    # lssd fast with luminance normalisation, single-level overloads with a prediction
    p = po.make_params("lssd", "fast", half=6, luminance=True)
    ok, uv, st = R.klt_track(p, rl, cl, pts)
    out["lssd_fast_h6_lum_uv"], out["lssd_fast_h6_lum_st"] = uv, st
    pred = pts + np.float32(3.0)
    for v in ("basic", "affine", "lssd"):
        p = po.make_params(v, "fast", half=6, predict=(0.9995, -0.03, 0.03, 0.9995))
        ok, uv, st = R.klt_track(p, rl, cl, pts, cur_uv=pred, single_level=True)
        out[f"{v}_fast_h6_single_uv"], out[f"{v}_fast_h6_single_st"] = uv, st
    np.savez_compressed(os.path.join(HERE, "euroc_klt_golden.npz"), **out)
    print("euroc_klt_golden.npz:", len(out), "arrays,", len(pts), "features")

    # ---- descriptor matching on seeded inputs ----------------------------------------------------------------
    m = {}
    rb, cb, pred, pos, _ = S.make_brief_sets(300, 320, seed=5)
    rb[7] = cb[11]  # an exact duplicate: distance 0 exercises the `break`
    cb[200] = cb[11]  # a tie at distance 0 further down: lowest j must win
    m["brief_ref"], m["brief_cur"], m["brief_pred"], m["brief_pos"] = rb, cb, pred, pos
    ok, m["brief_force_idx"] = R.match_brief_force(rb, cb, 60.0)
    ok, m["brief_nearby_idx"] = R.match_brief_nearby(rb, cb, pred, pos, 50, 50, 60.0)
    ok, muv, mst = R.match_brief_nearby_uv(rb, cb, pred, pos, 50, 50, 60.0)
    m["brief_nearby_uv"], m["brief_nearby_st"] = muv, mst
    rf, cf = S.make_float_sets(200, 220, dim=256, seed=6)
    fpos = np.stack([np.linspace(0, 700, 220), np.linspace(0, 400, 220)], 1).astype(np.float32)
    fpred = fpos[np.random.default_rng(1).integers(0, 220, 200)] + np.float32(5.0)
    m["float_ref"], m["float_cur"], m["float_pred"], m["float_pos"] = rf, cf, fpred, fpos
    ok, m["float_force_idx"] = R.match_cosine_force(rf, cf, 0.1)
    ok, m["float_nearby_idx"] = R.match_cosine_nearby(rf, cf, fpred, fpos, 50, 50, 0.3)
    np.savez_compressed(os.path.join(HERE, "matcher_golden.npz"), **m)
    print("matcher_golden.npz:", {k: v.shape for k, v in m.items() if k.endswith("idx")},
          "force matched", int((m["brief_force_idx"] >= 0).sum()), "nearby matched", int((m["brief_nearby_idx"] >= 0).sum()),
          "float force", int((m["float_force_idx"] >= 0).sum()), "float nearby", int((m["float_nearby_idx"] >= 0).sum()))

    # ---- direct-method pose tracker on the reference's own KITTI fixture (test/test_direct_method.cpp) ------------------------
    dm_dir = "/root/reference/example/direct_method"
    left = np.array(Image.open(os.path.join(dm_dir, "left.png")).convert("L"))
    disparity = np.array(Image.open(os.path.join(dm_dir, "disparity.png")).convert("L"))
    curs = [np.array(Image.open(os.path.join(dm_dir, f"00000{i}.png")).convert("L")) for i in (1, 2)]
    fx = fy = np.float32(718.856)
    cx, cy, baseline = np.float32(607.1928), np.float32(185.2157), np.float32(0.573)  # test_direct_method.cpp:15-20
    rng = np.random.default_rng(7)  # the demo draws 300 rand() pixels (:44-48); seeded here, zero-disparity pixels skipped
    uv = []
    while len(uv) < 300:
        c, r = int(rng.integers(0, left.shape[1])), int(rng.integers(0, left.shape[0]))
        if disparity[r, c] > 0:
            uv.append((c, r))
    uv = np.array(uv, np.float32)
    depth = (fx * baseline / disparity[uv[:, 1].astype(int), uv[:, 0].astype(int)].astype(np.float32)).astype(np.float32)
    p_w = (np.stack([(uv[:, 0] - cx) / fx, (uv[:, 1] - cy) / fy, np.ones(300, np.float32)], 1).astype(np.float32) * depth[:, None]).astype(np.float32)
    dm_levels = 5  # :39, :76
    d = {"left": left, "cur1": curs[0], "cur2": curs[1], "uv": uv, "p_c_in_ref": p_w, "K": np.array([fx, fy, cx, cy], np.float32),
         "levels": np.int32(dm_levels)}
    ll = R.pyramid_build(left, dm_levels)
    q, p = np.array([1, 0, 0, 0], np.float32), np.zeros(3, np.float32)  # q_ref = identity, p_ref = 0: world frame == reference camera frame
    cur_uv, st = None, None
    for i, cur_img in enumerate(curs, 1):  # the demo carries pose, positions and status from frame to frame (:69-86)
        ok, cur_uv, q, p, st = R.direct_method_track(po.make_direct_params(), ll, R.pyramid_build(cur_img, dm_levels), d["K"], p_w, uv, q, p,
                                                     cur_uv=cur_uv, status=st)
        assert ok
        d[f"q_{i}"], d[f"p_{i}"], d[f"uv_{i}"], d[f"st_{i}"] = q, p, cur_uv, st
        print(f"direct method frame {i}: q_rc = {q}, p_rc = {p}, inside = {int((st == 1).sum())}")
    np.savez_compressed(os.path.join(HERE, "direct_method_golden.npz"), **d)

    # ---- dense optical flow on the reference's EuRoC fixture pair (the committed images of euroc_klt_golden.npz) ---------------
    ok, fr, fc = R.dense_flow_track(po.make_dense_flow_params(), rl, cl)  # reference defaults, 4 levels (test_dense_optical_flow.cpp)
    assert ok
    import hashlib
    df = {"flow_row_8": fr[::8, ::8].copy(), "flow_col_8": fc[::8, ::8].copy(),
          "sha256": np.frombuffer(hashlib.sha256(fr.tobytes() + fc.tobytes()).digest(), np.uint8).copy()}
    ok, fr1, fc1 = R.dense_flow_track(po.make_dense_flow_params(half=1, max_iter=4), rl[:1], cl[:1], single_level=True)
    df["single_h1_row_8"], df["single_h1_col_8"] = fr1[::8, ::8].copy(), fc1[::8, ::8].copy()
    df["single_h1_sha256"] = np.frombuffer(hashlib.sha256(fr1.tobytes() + fc1.tobytes()).digest(), np.uint8).copy()
    np.savez_compressed(os.path.join(HERE, "dense_flow_golden.npz"), **df)
    print("dense_flow_golden.npz: mean |flow| =", float(np.abs(fr).mean()), float(np.abs(fc).mean()))


if __name__ == "__main__":
    main()
