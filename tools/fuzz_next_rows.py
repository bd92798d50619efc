#!/usr/bin/env python
"""Randomised parity sweep of the SURVEY 8(f) kernels (direct-method pose tracker, dense flow, forward-backward pass, sequence
pipeline, detector + BRIEF) against the C oracle.
    python tools/fuzz_next_rows.py [n_cases] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import feature_tracker_b200 as ft  # noqa: E402
from feature_tracker_b200 import synthetic as S  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype == np.float32:
        return np.array_equal(a.view(np.uint32), b.view(np.uint32)) or (np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a), np.nan_to_num(b)))
    return np.array_equal(a, b)


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    ctx = ft.Context(0)
    oracle = po.OracleLib()
    bad = 0
    for case in range(n_cases):
        kind = rng.choice(["direct_method", "dense_flow", "detector"])
        rows, cols = int(rng.integers(40, 200)), int(rng.integers(40, 260))
        levels = int(rng.integers(1, 5))
        while (min(rows, cols) >> (levels - 1)) < 8:
            levels -= 1
        if kind == "direct_method":
            n = int(rng.integers(1, 120))
            ref, cur, uv, K, pts = S.make_direct_method_scene(rows, cols, n, pair_id=2000 + case, border=3, depth=float(rng.uniform(1, 20)),
                                                              focal=float(rng.uniform(100, 600)))
            pts = pts.copy()
            kbad = max(1, uv.shape[0] // 6)
            pts[:kbad, 2] = rng.choice([-1.0, 0.0, 1e-7, 0.5], kbad)
            hr, hc = int(rng.integers(1, 8)), int(rng.integers(1, 8))
            max_points = int(rng.choice([500, max(1, uv.shape[0] // 2)]))
            q0 = np.array([1, 0, 0, 0], np.float32) + rng.normal(0, 0.01, 4).astype(np.float32)
            p0 = rng.normal(0, 0.05, 3).astype(np.float32)
            pred = uv + rng.normal(0, 1, uv.shape).astype(np.float32) if rng.random() < 0.4 else None
            st_in = rng.integers(0, 5, uv.shape[0]).astype(np.uint8) if rng.random() < 0.4 else None
            dm = ft.DirectMethod(ctx)
            o = dm.options()
            o.kPatchRowHalfSize, o.kPatchColHalfSize, o.kMaxTrackPointsNumber = hr, hc, max_points
            o.kMaxIteration = int(rng.integers(1, 16))
            pyr = ft.ImagePyramidBatch(ctx, rows, cols, levels, 2)
            pyr.SetRawImages(np.stack([ref, cur]))
            pyr.CreateImagePyramid()
            got = dm.TrackFeatures(pyr, pyr, K, pts, uv, q0, p0, cur_pixel_uv=pred, status=st_in, ref_image=0, cur_image=1)
            prm = po.make_direct_params(half=hr, half_col=hc, max_points=max_points, max_iter=o.kMaxIteration)
            exp = oracle.direct_method_track(prm, oracle.pyramid_build(ref, levels), oracle.pyramid_build(cur, levels), K, pts, uv, q0, p0, cur_uv=pred,
                                             status=st_in)
            ok = got[0] == exp[0] and all(same(g, e) for g, e in zip(got[1:], exp[1:]))
            desc = f"{2 * hr + 1}x{2 * hc + 1} n={uv.shape[0]}"
            pyr.close()
        elif kind == "detector":
            img = S.make_image(rows, cols, seed=4000 + case)
            if rng.random() < 0.3:  # quantise: plateaus of exactly equal responses
                img = (img // int(rng.choice([16, 64]))).astype(np.uint8) * 3
            dkind = str(rng.choice(["harris", "shi_tomasi"]))
            half, dist, needed = int(rng.integers(1, 4)), int(rng.choice([0, 1, 2, 5, 12, 20, 33, 70])), int(rng.choice([1, 30, 300, 100000]))
            thr = float(rng.choice([-1e20, 0.0, 40.0, 1e3, 1e5, 1e7]))
            existing = (rng.uniform(-0.1, 1.1, (int(rng.integers(1, 40)), 2)) * [cols, rows]).astype(np.float32) if rng.random() < 0.5 else None
            det = (ft.FeaturePointHarrisDetector if dkind == "harris" else ft.FeaturePointShiTomasDetector)(ctx)
            o = det.options()
            o.kHalfPatchSize, o.kMinValidResponse, o.kMinFeatureDistance = half, thr, dist
            pyr = ft.ImagePyramidBatch(ctx, rows, cols, levels, 1)
            pyr.SetRawImages(img[None])
            pyr.CreateImagePyramid()
            n0 = 0 if existing is None else len(existing)
            g_ok, g_uv, g_resp = det.DetectGoodFeatures(pyr, needed + n0, existing, return_response=True)
            e_ok, e_uv, e_resp = oracle.detect_features(po.make_detector_params(dkind, half, 0.04, thr, dist), img, needed, existing=existing)
            ok = g_ok == e_ok and same(g_uv[n0:], e_uv) and same(g_resp, e_resp)
            n_bits, bh = int(rng.choice([32, 128, 256, 512])), int(rng.integers(0, 12))
            uv = np.concatenate([e_uv[:200], (rng.uniform(-0.2, 1.2, (20, 2)) * [cols, rows]).astype(np.float32)])
            bd = ft.BriefDescriptor(ctx)
            bd.options().kLength, bd.options().kHalfPatchSize, bd.options().kPatternSeed = n_bits, bh, int(rng.integers(0, 1 << 31))
            g_ok, g_desc, g_valid = bd.Compute(pyr, uv)
            e_ok, e_desc, e_valid = oracle.describe_brief(img, uv, bd.pattern(), bh)
            ok = ok and g_ok == e_ok and same(g_desc, e_desc) and same(g_valid, e_valid)
            desc = f"{dkind} h={half} d={dist} thr={thr} needed={needed} existing={n0} found={len(e_uv)} brief={n_bits}/{bh}"
            pyr.close()
        else:
            ref, cur, _, _ = S.make_pair(rows, cols, 5, pair_id=3000 + case)
            half = int(rng.integers(0, 5))
            single = bool(rng.random() < 0.3)
            flow = (rng.normal(0, 2, ref.shape).astype(np.float32), rng.normal(0, 2, ref.shape).astype(np.float32)) if single and rng.random() < 0.5 else None
            dof = ft.DenseOpticalFlow(ctx)
            o = dof.options()
            o.kHalfPatchSize, o.kMaxIteration, o.kMaxDeltaFlowStep = half, int(rng.integers(1, 12)), float(rng.choice([0.25, 1.0, 4.0]))
            pyr = ft.ImagePyramidBatch(ctx, rows, cols, levels, 2)
            pyr.SetRawImages(np.stack([ref, cur]))
            pyr.CreateImagePyramid()
            got = dof.Track(pyr, pyr, flow_rc=flow, single_level=single, ref_image=0, cur_image=1)
            prm = po.make_dense_flow_params(max_iter=o.kMaxIteration, half=half, max_step=o.kMaxDeltaFlowStep)
            exp = oracle.dense_flow_track(prm, oracle.pyramid_build(ref, levels), oracle.pyramid_build(cur, levels), single_level=single, flow=flow)
            ok = got[0] == exp[0] and same(got[1], exp[1]) and same(got[2], exp[2])
            desc = f"half={half} single={single}"
            pyr.close()
        if not ok:
            bad += 1
            print(f"MISMATCH case {case}: {kind} {rows}x{cols} L{levels} {desc}")
    print(f"fuzz next rows: {n_cases} cases, {bad} mismatching")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
