#!/usr/bin/env python
"""Randomised parity sweep of the descriptor matchers and the score-matrix post-processing (GPU vs the C oracle).
    python tools/fuzz_matchers.py [n_cases] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import feature_tracker_b200 as ft  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    ctx = ft.Context(0)
    oracle = po.OracleLib()
    bad = 0
    for case in range(n_cases):
        kind = rng.choice(["brief_force", "brief_nearby", "cos_force", "cos_nearby", "mutual"])
        n_ref, n_cur = int(rng.integers(1, 700)), int(rng.integers(1, 900))
        pos = np.stack([rng.uniform(0, 752, n_cur), rng.uniform(0, 480, n_cur)], 1).astype(np.float32)
        pred = (pos[rng.integers(0, n_cur, n_ref)] + rng.normal(0, 12, (n_ref, 2))).astype(np.float32)
        win = (int(rng.integers(0, 90)), int(rng.integers(0, 90)))
        idx_in = rng.integers(-1, n_cur, n_ref).astype(np.int32) if rng.random() < 0.3 else None
        if kind.startswith("brief"):
            bits = int(rng.choice([32, 64, 96, 128, 256, 512]))
            cur = rng.integers(0, 2, (n_cur, bits), dtype=np.uint8)
            ref = cur[rng.integers(0, n_cur, n_ref)].copy()
            flip = rng.random(ref.shape) < rng.choice([0.0, 0.05, 0.2, 0.5])
            ref[flip] ^= 1
            if n_cur > 3:
                cur[rng.integers(0, n_cur)] = cur[0]  # exact duplicates: ties at distance 0
            thr = float(rng.choice([0.5, 8.0, 40.0, 60.0, 1e9]))
            m = ft.BriefMatcher(ctx)
            m.options().kMaxValidDescriptorDistance = thr
            m.options().kMaxValidPredictRowDistance, m.options().kMaxValidPredictColDistance = win
            if kind == "brief_force":
                got = m.ForceMatch(ft.pack_brief(ref), ft.pack_brief(cur), idx_in)
                exp = oracle.match_brief_force(ref, cur, thr, idx=idx_in)
            else:
                got = m.NearbyMatch(ft.pack_brief(ref), ft.pack_brief(cur), pred, pos, idx_in)
                exp = oracle.match_brief_nearby(ref, cur, pred, pos, win[0], win[1], thr, idx=idx_in)
        elif kind.startswith("cos"):
            dim = int(rng.choice([1, 7, 16, 37, 64, 100, 128, 250, 256]))
            cur = rng.normal(0, 1, (n_cur, dim)).astype(np.float32)
            ref = (cur[rng.integers(0, n_cur, n_ref)] + rng.choice([0.0, 0.05, 0.3]) * rng.normal(0, 1, (n_ref, dim))).astype(np.float32)
            if rng.random() < 0.5:
                cur /= np.linalg.norm(cur, axis=1, keepdims=True)
            if n_cur > 3:
                cur[rng.integers(0, n_cur)] = cur[1]  # duplicate: exact tie
            thr = float(rng.choice([0.01, 0.1, 0.3, 2.0]))
            m = ft.CosineMatcher(ctx)
            m.options().kMaxValidDescriptorDistance = thr
            m.options().kMaxValidPredictRowDistance, m.options().kMaxValidPredictColDistance = win
            if kind == "cos_force":
                got = m.ForceMatch(ref, cur, idx_in)
                exp = oracle.match_cosine_force(ref, cur, thr, idx=idx_in)
            else:
                got = m.NearbyMatch(ref, cur, pred, pos, idx_in)
                exp = oracle.match_cosine_nearby(ref, cur, pred, pos, win[0], win[1], thr, idx=idx_in)
        else:
            s = rng.normal(-6, 3, (n_ref, n_cur)).astype(np.float32)
            s[rng.random(s.shape) < 0.01] = np.float32(-0.25)
            s[rng.random(s.shape) < 0.002] = -np.inf
            s[rng.random(s.shape) < 0.001] = np.nan
            thr = float(rng.choice([-3.0, -1.0, -50.0]))
            m = ft.NNFeatureMatcher(ctx)
            m.options().kMinValidMatchScore = thr
            got = m.MatchScores(s)
            exp = oracle.mutual_scores(s, thr)
        if not (got[0] == exp[0] and np.array_equal(got[1], exp[1])):
            bad += 1
            diff = np.nonzero(np.asarray(got[1]) != np.asarray(exp[1]))[0]
            print(f"MISMATCH case {case}: {kind} n_ref={n_ref} n_cur={n_cur} rows {diff[:8]} gpu {np.asarray(got[1])[diff[:8]]} oracle {np.asarray(exp[1])[diff[:8]]}")
    print(f"fuzz matchers: {n_cases} cases, {bad} mismatching")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
