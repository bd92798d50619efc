#!/usr/bin/env python
"""Randomised parity sweep (GPU vs the C oracle) over tracker variants / methods / patch sizes / pyramid depths / image sizes,
with border and far-outside features, predictions and entry statuses.  Not part of the test-suite (minutes of CPU time);
    python tools/fuzz_parity.py [n_cases] [seed]
prints one line per mismatching case and a summary."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import feature_tracker_b200 as ft  # noqa: E402
from feature_tracker_b200 import synthetic as S  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

VARIANTS = {"basic": ft.OpticalFlowBasicKlt, "affine": ft.OpticalFlowAffineKlt, "lssd": ft.OpticalFlowLssdKlt}
METHODS = {"inverse": ft.OpticalFlowMethod.kInverse, "direct": ft.OpticalFlowMethod.kDirect, "fast": ft.OpticalFlowMethod.kFast}


def in_child(fn):
    """Runs fn() in a forked child and returns its pickled result, or None when the child died."""
    import pickle
    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:
        os.close(r)
        try:
            data = pickle.dumps(fn())
            with os.fdopen(w, "wb") as f:
                f.write(data)
        finally:
            os._exit(0)
    os.close(w)
    with os.fdopen(r, "rb") as f:
        data = f.read()
    _, status = os.waitpid(pid, 0)
    if status != 0 or not data:
        return None
    return pickle.loads(data)


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    ctx = ft.Context(0)
    oracle = po.OracleLib()
    bad = crashed = undefined = 0
    for case in range(n_cases):
        if os.environ.get("FUZZ_VERBOSE"):
            print("case", case, flush=True)
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
        k = max(1, uv.shape[0] // 5)  # features on / beyond the border, integer and half-integer positions
        uv[:k, 0] = rng.uniform(-3, cols + 3, k)
        uv[:k, 1] = rng.uniform(-3, rows + 3, k)
        uv[k:2 * k] = np.round(uv[k:2 * k] * 2) / 2
        single = bool(rng.random() < 0.2)
        pred = uv + rng.normal(0, 2, uv.shape).astype(np.float32) if rng.random() < 0.4 else None
        st_in = rng.integers(0, 5, uv.shape[0]).astype(np.uint8) if rng.random() < 0.3 else None
        max_points = int(rng.choice([500, max(1, uv.shape[0] // 2)]))
        klt = VARIANTS[variant](ctx)
        o = klt.options()
        o.kPatchRowHalfSize, o.kPatchColHalfSize, o.kMethod, o.kMaxTrackPointsNumber = hr, hc, METHODS[method], max_points
        lum = bool(variant == "lssd" and rng.random() < 0.5)
        if variant == "lssd":
            klt.consider_patch_luminance = lum
        pyr = ft.ImagePyramidBatch(ctx, rows, cols, levels, 2)
        pyr.SetRawImages(np.stack([ref, cur]))
        pyr.CreateImagePyramid()
        got = klt.TrackFeatures(pyr, pyr, uv, cur_pixel_uv=pred, status=st_in, single_level=single, ref_image=0, cur_image=1)
        prm = po.make_params(variant, method, half=hr, half_col=hc, max_points=max_points, luminance=lum)
        # The reference (and so the oracle) reads out of bounds when a tracker diverges to NaN positions (e.g. LSSD kFast with 3x3
        # patches): evaluate it in a forked child so that such a crash only skips the case.  The GPU result above is still computed.
        def run_oracle():
            oracle.lib.ftko_outside_reads.restype = C.c_longlong
            oracle.lib.ftko_outside_reads(C.c_int32(1))
            r = oracle.klt_track(prm, oracle.pyramid_build(ref, levels), oracle.pyramid_build(cur, levels), uv, cur_uv=pred, status=st_in, single_level=single)
            return r + (int(oracle.lib.ftko_outside_reads(C.c_int32(0))),)
        exp = in_child(run_oracle)
        if exp is None:
            crashed += 1
            print(f"reference crashed (undefined behaviour) on case {case}: {variant}/{method} {2 * hr + 1}x{2 * hc + 1}; GPU returned normally")
            pyr.close()
            continue
        same_st = np.array_equal(got[2], exp[2])
        same_uv = np.array_equal(got[1].view(np.uint32), exp[1].view(np.uint32)) or np.array_equal(np.nan_to_num(got[1]), np.nan_to_num(exp[1]))
        if not (got[0] == exp[0] and same_st and same_uv):
            if exp[3] > 0:  # the reference sampled outside the image without a bounds test: its result is undefined there
                undefined += 1
                print(f"reference read outside the image ({exp[3]} unchecked samples) on case {case}: {variant}/{method} {2 * hr + 1}x{2 * hc + 1}; results differ")
                pyr.close()
                continue
            bad += 1
            d = np.abs(got[1].astype(np.float64) - exp[1].astype(np.float64))
            print(f"MISMATCH case {case}: {variant}/{method} {2 * hr + 1}x{2 * hc + 1} {rows}x{cols} L{levels} single={single} n={uv.shape[0]} "
                  f"status_diff={int((got[2] != exp[2]).sum())} max_pos_diff={np.nanmax(d) if d.size else 0}")
        pyr.close()
    print(f"fuzz: {n_cases} cases, {bad} mismatching, {crashed} where the reference itself crashed, "
          f"{undefined} differing where the reference read outside the image (undefined behaviour)")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
