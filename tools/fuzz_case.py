#!/usr/bin/env python
"""Re-runs one case of tools/fuzz_parity.py (same seed) and prints the features whose GPU result differs from the oracle's.
    python tools/fuzz_case.py <seed> <case>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import feature_tracker_b200 as ft  # noqa: E402
from feature_tracker_b200 import synthetic as S  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tools.fuzz_parity import METHODS, VARIANTS  # noqa: E402


def draw(rng, case):
    rows, cols = int(rng.integers(40, 300)), int(rng.integers(40, 400))
    levels = int(rng.integers(1, 6))
    while (min(rows, cols) >> (levels - 1)) < 8:
        levels -= 1
    variant = rng.choice(list(VARIANTS))
    method = rng.choice(list(METHODS))
    hr, hc = int(rng.integers(1, 11)), int(rng.integers(1, 11))
    if rng.random() < 0.5:
        hc = hr
    n = int(rng.integers(1, 80))
    ref, cur, uv, _ = S.make_pair(rows, cols, n, pair_id=1000 + case, border=2)
    uv = uv.copy()
    k = max(1, uv.shape[0] // 5)
    uv[:k, 0] = rng.uniform(-3, cols + 3, k)
    uv[:k, 1] = rng.uniform(-3, rows + 3, k)
    uv[k:2 * k] = np.round(uv[k:2 * k] * 2) / 2
    single = bool(rng.random() < 0.2)
    pred = uv + rng.normal(0, 2, uv.shape).astype(np.float32) if rng.random() < 0.4 else None
    st_in = rng.integers(0, 5, uv.shape[0]).astype(np.uint8) if rng.random() < 0.3 else None
    max_points = int(rng.choice([500, max(1, uv.shape[0] // 2)]))
    lum = bool(variant == "lssd" and rng.random() < 0.5)
    return dict(rows=rows, cols=cols, levels=levels, variant=variant, method=method, hr=hr, hc=hc, ref=ref, cur=cur, uv=uv, single=single, pred=pred,
                st_in=st_in, max_points=max_points, lum=lum)


def main():
    seed, target = int(sys.argv[1]), int(sys.argv[2])
    rng = np.random.default_rng(seed)
    for case in range(target + 1):
        c = draw(rng, case)
    print({k: v for k, v in c.items() if k not in ("ref", "cur", "uv", "pred", "st_in")})
    oracle = po.OracleLib()
    rl, cl = oracle.pyramid_build(c["ref"], c["levels"]), oracle.pyramid_build(c["cur"], c["levels"])
    prm = po.make_params(c["variant"], c["method"], half=c["hr"], half_col=c["hc"], max_points=c["max_points"], luminance=c["lum"])
    exp = oracle.klt_track(prm, rl, cl, c["uv"], cur_uv=c["pred"], status=c["st_in"], single_level=c["single"])
    if len(sys.argv) > 3 and sys.argv[3] == "cpu":
        print(exp[1], exp[2])
        return
    ctx = ft.Context(0)
    klt = VARIANTS[c["variant"]](ctx)
    o = klt.options()
    o.kPatchRowHalfSize, o.kPatchColHalfSize, o.kMethod, o.kMaxTrackPointsNumber = c["hr"], c["hc"], METHODS[c["method"]], c["max_points"]
    if c["variant"] == "lssd":
        klt.consider_patch_luminance = c["lum"]
    pyr = ft.ImagePyramidBatch(ctx, c["rows"], c["cols"], c["levels"], 2)
    pyr.SetRawImages(np.stack([c["ref"], c["cur"]]))
    pyr.CreateImagePyramid()
    got = klt.TrackFeatures(pyr, pyr, c["uv"], cur_pixel_uv=c["pred"], status=c["st_in"], single_level=c["single"], ref_image=0, cur_image=1)
    for i in range(c["uv"].shape[0]):
        if got[2][i] != exp[2][i] or not np.array_equal(got[1][i].view(np.uint32), exp[1][i].view(np.uint32)):
            print(f"feature {i}: ref_uv {c['uv'][i]} pred {None if c['pred'] is None else c['pred'][i]} st_in {None if c['st_in'] is None else c['st_in'][i]}"
                  f" | gpu {got[1][i]} st {got[2][i]} | oracle {exp[1][i]} st {exp[2][i]}")


if __name__ == "__main__":
    main()
