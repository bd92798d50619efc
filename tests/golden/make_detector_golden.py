"""Detector / BRIEF regression fixture.  UNLIKE the other fixtures this one does not come from the reference: the detector and
descriptor classes live in the absent sibling repository Feature_Detector, so parity for SURVEY 8(f) rank 1 is unpinned.  The file
freezes what oracle/ftk_oracle.c's restatement of the published algorithm gives on the reference's own EuRoC image pair (taken
from euroc_klt_golden.npz), with the option values test/test_descriptor_matcher_brief.cpp:59-76 sets, so that later changes to
the oracle or the kernels cannot drift silently.

    python tests/golden/make_detector_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    g = np.load(os.path.join(HERE, "euroc_klt_golden.npz"))
    O = po.OracleLib()
    out = {}
    pattern = O.brief_pattern(256, 8, 0)
    out["pattern"] = pattern
    for name in ("ref", "cur"):
        img = g[name]
        for kind in ("harris", "shi_tomasi"):
            prm = po.make_detector_params(kind, 1, 0.04, 40.0, 20)
            ok, uv, resp = O.detect_features(prm, img, 300)
            assert ok
            out[f"{name}_{kind}_uv"], out[f"{name}_{kind}_response"] = uv, resp
        ok, desc, valid = O.describe_brief(img, out[f"{name}_harris_uv"], pattern, 8)
        assert ok
        out[f"{name}_brief"], out[f"{name}_brief_valid"] = desc, valid
    np.savez_compressed(os.path.join(HERE, "detector_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
